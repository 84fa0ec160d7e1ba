"""ctypes wrapper of the TEST-ONLY CPU logic simulator (libdpe_hostsim.so).

TEST INFRASTRUCTURE: only tests/ and oracle/make_stage_golden.py import this.  See
oracle/dpe_hostsim.cu for why it exists and why it is neither an oracle nor a fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


class StageParams(C.Structure):
    """dpe_stage_params of include/dpe_b200.h"""
    _fields_ = [("state", C.c_int), ("geom_consistency", C.c_int), ("use_apd", C.c_int),
                ("max_iterations", C.c_int), ("top_k", C.c_int), ("weak_peak_radius", C.c_int),
                ("rotate_time", C.c_int), ("ransac_threshold", C.c_float), ("geom_factor", C.c_float)]


FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WEAK, STRONG, UNKNOWN = 0, 1, 2


def stage_schedule(n_scales):
    """The reference's coarse-to-fine schedule (main.cpp:508-567): list of (scale_idx, StageParams)."""
    out = []
    for i in range(n_scales):
        p = StageParams(FIRST_INIT if i == 0 else REFINE_INIT, 0, int(i > 0), 3, 4, 6, 4, 0.005, 0.2)
        if i > 0:
            p.ransac_threshold = 0.01 - i * 0.00125
            p.rotate_time = min(2 ** i, 4)
        out.append((i, p))
        for j in range(3):
            q = StageParams(REFINE_ITER, 1, int(i > 0), 3, 4, max(4 - 2 * j, 2), min(2 ** i, 4),
                            0.01 - i * 0.00125, 0.2)
            out.append((i, q))
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        import subprocess
        so = HERE / "_ref" / "libdpe_hostsim.so"
        rc = subprocess.call(["make", "-s", "-C", str(HERE), "_ref/libdpe_hostsim.so"])
        if rc != 0 and not so.exists():
            raise RuntimeError("cannot build oracle/_ref/libdpe_hostsim.so")
        _lib = C.CDLL(str(so))
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ptr_array(arrs):
    T = C.POINTER(C.c_float) * len(arrs)
    return T(*[_fp(a) for a in arrs])


def resize_linear(img, dw, dh):
    src = np.ascontiguousarray(img, dtype=np.float32)
    dst = np.empty((dh, dw), np.float32)
    lib().dpe_hostsim_resize_linear(_fp(src), src.shape[1], src.shape[0], _fp(dst), dw, dh)
    return dst


def cost_eval(images, cams, full_wh, xy, planes, quant=1, centred=True):
    """images: [ref, src...] float32 HxW; cams: [(K,R,t)...]; returns n_pix x n_src costs."""
    imgs = [np.ascontiguousarray(i, np.float32) for i in images]
    H, W = imgs[0].shape
    n_src = len(imgs) - 1
    K = np.ascontiguousarray(np.stack([c[0] for c in cams]).reshape(-1), np.float32)
    R = np.ascontiguousarray(np.stack([c[1] for c in cams]).reshape(-1), np.float32)
    t = np.ascontiguousarray(np.stack([c[2] for c in cams]).reshape(-1), np.float32)
    xy = np.ascontiguousarray(xy, np.int32)
    planes = np.ascontiguousarray(planes, np.float32)
    out = np.empty((len(xy), n_src), np.float32)
    import os
    # cost arithmetic (dpe_core.cuh): centred by default here — this entry point serves the float64 comparisons
    if centred:
        os.environ["DPE_HOSTSIM_CENTRED"] = "1"
    else:
        os.environ.pop("DPE_HOSTSIM_CENTRED", None)
    lib().dpe_hostsim_cost_eval(W, H, full_wh[0], full_wh[1], _fp(imgs[0]), n_src, _ptr_array(imgs[1:]), _fp(K),
                                _fp(R), _fp(t), quant, len(xy), xy.ctypes.data_as(C.POINTER(C.c_int)), _fp(planes),
                                _fp(out))
    os.environ.pop("DPE_HOSTSIM_CENTRED", None)
    return out


def run_stage(images, cams, depth_range, full_wh, params, seed, view=0, stage_counter=0, prev=None,
              src_depths=None, edge=None, edge_low=None, label=None, quant=1):
    """One (view, stage) on the CPU.  prev = (planes HxWx4, state, selected) or None.
    Returns dict(planes, state, selected, depth, units)."""
    imgs = [np.ascontiguousarray(i, np.float32) for i in images]
    H, W = imgs[0].shape
    n_src = len(imgs) - 1
    K = np.ascontiguousarray(np.stack([c[0] for c in cams]).reshape(-1), np.float32)
    R = np.ascontiguousarray(np.stack([c[1] for c in cams]).reshape(-1), np.float32)
    t = np.ascontiguousarray(np.stack([c[2] for c in cams]).reshape(-1), np.float32)
    planes = np.empty((H, W, 4), np.float32)
    state = np.empty((H, W), np.uint8)
    sel = np.empty((H, W), np.uint32)
    depth = np.empty((H, W), np.float32)
    units = C.c_double(0)
    null_f = C.POINTER(C.c_float)()
    null_b = C.POINTER(C.c_uint8)()
    null_u = C.POINTER(C.c_uint32)()
    null_i = C.POINTER(C.c_int32)()
    if prev is not None:
        pp = np.ascontiguousarray(prev[0], np.float32)
        ps = np.ascontiguousarray(prev[1], np.uint8)
        pl = np.ascontiguousarray(prev[2], np.uint32)
        pH, pW = ps.shape
        a_pp, a_ps, a_pl = _fp(pp), ps.ctypes.data_as(C.POINTER(C.c_uint8)), pl.ctypes.data_as(C.POINTER(C.c_uint32))
    else:
        pH, pW = H, W
        a_pp, a_ps, a_pl = null_f, null_b, null_u
    sd = None
    if src_depths is not None:
        sdl = [np.ascontiguousarray(d, np.float32) for d in src_depths]
        sd = _ptr_array(sdl)
    e = np.ascontiguousarray(edge, np.uint8) if edge is not None else None
    el = np.ascontiguousarray(edge_low, np.uint8) if edge_low is not None else None
    lb = np.ascontiguousarray(label, np.int32) if label is not None else None
    rc = lib().dpe_hostsim_stage(
        W, H, full_wh[0], full_wh[1], n_src, _ptr_array(imgs), _fp(K), _fp(R), _fp(t),
        C.c_float(depth_range[0]), C.c_float(depth_range[1]), sd if sd is not None else C.POINTER(C.POINTER(C.c_float))(),
        a_pp, a_ps, a_pl, pW, pH,
        e.ctypes.data_as(C.POINTER(C.c_uint8)) if e is not None else null_b,
        el.ctypes.data_as(C.POINTER(C.c_uint8)) if el is not None else null_b,
        el.shape[1] if el is not None else 0, el.shape[0] if el is not None else 0,
        lb.ctypes.data_as(C.POINTER(C.c_int32)) if lb is not None else null_i,
        C.byref(params), C.c_uint64(seed), view, C.c_uint32(stage_counter), quant,
        _fp(planes), state.ctypes.data_as(C.POINTER(C.c_uint8)), sel.ctypes.data_as(C.POINTER(C.c_uint32)),
        _fp(depth), C.byref(units))
    if rc != 0:
        raise RuntimeError(f"hostsim stage failed: {rc}")
    return dict(planes=planes, state=state, selected=sel, depth=depth, units=units.value)


def run_stage_dbg(images, cams, depth_range, full_wh, params, seed, stop_step, prev=None, src_depths=None, edge=None,
                  edge_low=None, label=None, quant=1):
    """Like run_stage but stops after `stop_step` (0 anchors, 1 init, 2+3i strong, 3+3i fit, 4+3i weak, 11 final)
    and returns the raw state there: dict(planes HxWx4, costs, selected, state, fit HxWx4, radius, neighbours HxWx9x2,
    reliable)."""
    imgs = [np.ascontiguousarray(i, np.float32) for i in images]
    H, W = imgs[0].shape
    n_src = len(imgs) - 1
    K = np.ascontiguousarray(np.stack([c[0] for c in cams]).reshape(-1), np.float32)
    R = np.ascontiguousarray(np.stack([c[1] for c in cams]).reshape(-1), np.float32)
    t = np.ascontiguousarray(np.stack([c[2] for c in cams]).reshape(-1), np.float32)
    null_f = C.POINTER(C.c_float)()
    null_b = C.POINTER(C.c_uint8)()
    null_u = C.POINTER(C.c_uint32)()
    null_i = C.POINTER(C.c_int32)()
    if prev is not None:
        pp = np.ascontiguousarray(prev[0], np.float32)
        ps = np.ascontiguousarray(prev[1], np.uint8)
        pl = np.ascontiguousarray(prev[2], np.uint32)
        pH, pW = ps.shape
        a_pp, a_ps, a_pl = _fp(pp), ps.ctypes.data_as(C.POINTER(C.c_uint8)), pl.ctypes.data_as(C.POINTER(C.c_uint32))
    else:
        pH, pW = H, W
        a_pp, a_ps, a_pl = null_f, null_b, null_u
    sd = None
    if src_depths is not None:
        sdl = [np.ascontiguousarray(d, np.float32) for d in src_depths]
        sd = _ptr_array(sdl)
    e = np.ascontiguousarray(edge, np.uint8) if edge is not None else None
    el = np.ascontiguousarray(edge_low, np.uint8) if edge_low is not None else None
    lb = np.ascontiguousarray(label, np.int32) if label is not None else None
    out = dict(planes=np.zeros((H, W, 4), np.float32), costs=np.zeros((H, W), np.float32), selected=np.zeros((H, W), np.uint32),
               state=np.zeros((H, W), np.uint8), fit=np.zeros((H, W, 4), np.float32), radius=np.zeros((H, W), np.int32),
               neighbours=np.zeros((H, W, 9, 2), np.int16), reliable=np.zeros((H, W), np.uint8))
    fn = lib().dpe_hostsim_stage_dbg
    fn.restype = C.c_int
    rc = fn(W, H, full_wh[0], full_wh[1], n_src, _ptr_array(imgs), _fp(K), _fp(R), _fp(t),
            C.c_float(depth_range[0]), C.c_float(depth_range[1]), sd if sd is not None else C.POINTER(C.POINTER(C.c_float))(),
            a_pp, a_ps, a_pl, pW, pH,
            e.ctypes.data_as(C.POINTER(C.c_uint8)) if e is not None else null_b,
            el.ctypes.data_as(C.POINTER(C.c_uint8)) if el is not None else null_b,
            el.shape[1] if el is not None else 0, el.shape[0] if el is not None else 0,
            lb.ctypes.data_as(C.POINTER(C.c_int32)) if lb is not None else null_i,
            C.byref(params), C.c_uint64(seed), quant, int(stop_step),
            _fp(out["planes"]), _fp(out["costs"]), out["selected"].ctypes.data_as(C.POINTER(C.c_uint32)),
            out["state"].ctypes.data_as(C.POINTER(C.c_uint8)), _fp(out["fit"]),
            out["radius"].ctypes.data_as(C.POINTER(C.c_int32)), out["neighbours"].ctypes.data_as(C.POINTER(C.c_int16)),
            out["reliable"].ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise RuntimeError(f"hostsim stage failed: {rc}")
    return out
