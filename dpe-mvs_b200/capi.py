"""ctypes binding of the C ABI (include/dpe_b200.h, lib/libdpe_b200.so).

This is how tests and bench.py reach the CUDA path; there is no CPU fallback: without a GPU
`Context()` raises (DPE_ERR_NO_DEVICE), without the built library `load()` raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
import os
LIB_PATH = Path(os.environ["DPE_LIB"]) if os.environ.get("DPE_LIB") else HERE / "lib" / "libdpe_b200.so"

FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WEAK, STRONG, UNKNOWN = 0, 1, 2


class StageParams(C.Structure):
    """dpe_stage_params (include/dpe_b200.h)"""
    _fields_ = [("state", C.c_int), ("geom_consistency", C.c_int), ("use_apd", C.c_int),
                ("max_iterations", C.c_int), ("top_k", C.c_int), ("weak_peak_radius", C.c_int),
                ("rotate_time", C.c_int), ("ransac_threshold", C.c_float), ("geom_factor", C.c_float)]


# every symbol include/dpe_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "dpe_ctx_create", "dpe_ctx_destroy", "dpe_last_error", "dpe_kernel_launches", "dpe_scene_begin",
    "dpe_scene_set_view", "dpe_scene_set_pairs", "dpe_scene_set_prep", "dpe_scene_set_shard", "dpe_scene_commit",
    "dpe_run_stage", "dpe_stage_begin", "dpe_stage_wait_view", "dpe_stage_end", "dpe_stage_atlas", "dpe_view_slot", "dpe_stage_commit",
    "dpe_shard_range", "dpe_comm_get_unique_id", "dpe_comm_init_rank", "dpe_comm_init_all", "dpe_comm_reset_all",
    "dpe_scene_broadcast_images", "dpe_export_view", "dpe_scene_set_active", "dpe_stage_comm_ms", "dpe_debug_set_variants", "dpe_fuse_prepare", "dpe_fuse_set_color", "dpe_fuse_set_block", "dpe_fuse_broadcast_colors", "dpe_fuse_get_ply_records", "dpe_viz_render", "dpe_cost_eval", "dpe_geom_eval", "dpe_get_size",
    "dpe_get_maps", "dpe_set_count_evals", "dpe_eval_units", "dpe_stage_gpu_ms", "dpe_probe_tex_rate",
    "dpe_probe_fma_rate", "dpe_probe_tex_pattern", "dpe_probe_tex_weights", "dpe_run_pipeline", "dpe_set_profile", "dpe_get_profile", "dpe_set_view_order", "dpe_bench_ncc", "dpe_set_reference_race", "dpe_debug_read", "dpe_debug_stop_after", "dpe_debug_set_maps", "dpe_set_cost_arithmetic", "dpe_fuse_set_view", "dpe_fuse_run", "dpe_fuse_get",
]

_lib = None


def _nccl_hint():
    """Multi-GPU runs bind NCCL at run time (csrc/dpe_capi.cu: NcclApi).  When this interpreter has a pip-installed
    NCCL (PyTorch's, newer than the system's, same soname), point the library at that file so that a later
    `import torch` in the same process finds the NCCL it was built against already loaded."""
    import importlib.util
    import os
    if os.environ.get("DPE_NCCL_LIB"):
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["DPE_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def load(build=True):
    """Loads libdpe_b200.so (building it in-tree first if asked and needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if build and not os.environ.get("DPE_LIB"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("dpe_build", HERE / "build.py")
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        b.build_lib()
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first (python dpe-mvs_b200/build.py)")
    _nccl_hint()
    lib = C.CDLL(str(LIB_PATH))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    lib.dpe_ctx_create.argtypes = [C.POINTER(vp), ci]
    lib.dpe_ctx_destroy.argtypes = [vp]
    lib.dpe_ctx_destroy.restype = None
    lib.dpe_last_error.argtypes = [vp]
    lib.dpe_last_error.restype = C.c_char_p
    lib.dpe_kernel_launches.argtypes = [vp]
    lib.dpe_kernel_launches.restype = C.c_longlong
    lib.dpe_scene_begin.argtypes = [vp, ci, ci, ci, ci]
    lib.dpe_scene_set_view.argtypes = [vp, ci, vp, vp, vp, vp, cf, cf]
    lib.dpe_scene_set_pairs.argtypes = [vp, ci, vp, ci]
    lib.dpe_scene_set_prep.argtypes = [vp, ci, ci, vp, vp, C.c_size_t]
    lib.dpe_scene_set_shard.argtypes = [vp, ci, ci, ci]
    lib.dpe_scene_broadcast_images.argtypes = [vp, ci]
    lib.dpe_scene_set_active.argtypes = [vp, ci, ci]
    lib.dpe_shard_range.argtypes = [ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]
    lib.dpe_comm_get_unique_id.argtypes = [vp, C.c_size_t]
    lib.dpe_comm_init_rank.argtypes = [vp, vp, ci, ci]
    lib.dpe_comm_init_all.argtypes = [C.POINTER(vp), ci]
    lib.dpe_comm_reset_all.argtypes = []
    lib.dpe_comm_reset_all.restype = None
    lib.dpe_stage_begin.argtypes = [vp, ci, C.POINTER(StageParams), C.c_uint64]
    lib.dpe_stage_wait_view.argtypes = [vp, ci]
    lib.dpe_stage_end.argtypes = [vp]
    lib.dpe_view_slot.argtypes = [vp, ci, C.POINTER(ci)]
    lib.dpe_export_view.argtypes = [vp, ci, vp, vp, vp]
    lib.dpe_scene_commit.argtypes = [vp]
    lib.dpe_run_stage.argtypes = [vp, ci, C.POINTER(StageParams), C.c_uint64]
    lib.dpe_stage_atlas.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.dpe_stage_commit.argtypes = [vp]
    lib.dpe_cost_eval.argtypes = [vp, ci, ci, ci, vp, vp, ci, vp]
    lib.dpe_geom_eval.argtypes = [vp, ci, ci, ci, vp, vp, vp]
    lib.dpe_get_size.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci)]
    lib.dpe_get_maps.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.dpe_set_count_evals.argtypes = [vp, ci]
    lib.dpe_eval_units.argtypes = [vp]
    lib.dpe_eval_units.restype = C.c_double
    lib.dpe_stage_gpu_ms.argtypes = [vp]
    lib.dpe_stage_gpu_ms.restype = C.c_double
    lib.dpe_stage_comm_ms.argtypes = [vp]
    lib.dpe_stage_comm_ms.restype = C.c_double
    lib.dpe_probe_tex_rate.argtypes = [vp, ci, ci, ci, C.POINTER(C.c_double)]
    lib.dpe_probe_fma_rate.argtypes = [vp, ci, C.POINTER(C.c_double)]
    lib.dpe_probe_tex_pattern.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci, ci, C.POINTER(C.c_double)]
    lib.dpe_probe_tex_weights.argtypes = [vp, ci, vp]
    lib.dpe_set_profile.argtypes = [vp, ci]
    lib.dpe_set_view_order.argtypes = [vp, ci]
    lib.dpe_set_reference_race.argtypes = [vp, ci]
    lib.dpe_debug_set_variants.argtypes = [vp, ci]
    lib.dpe_set_cost_arithmetic.argtypes = [vp, ci]
    lib.dpe_fuse_set_view.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.dpe_fuse_prepare.argtypes = [vp]
    lib.dpe_fuse_set_color.argtypes = [vp, ci, vp]
    lib.dpe_fuse_set_block.argtypes = [vp, ci, vp]
    lib.dpe_fuse_broadcast_colors.argtypes = [vp, ci]
    lib.dpe_fuse_run.argtypes = [vp, ci, ci, C.POINTER(C.c_size_t)]
    lib.dpe_fuse_get.argtypes = [vp, vp, vp]
    lib.dpe_fuse_get_ply_records.argtypes = [vp, vp]
    lib.dpe_viz_render.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(ci), C.POINTER(ci)]
    lib.dpe_debug_read.argtypes = [vp, ci, vp, C.c_size_t]
    lib.dpe_debug_stop_after.argtypes = [vp, ci]
    lib.dpe_debug_set_maps.argtypes = [vp, ci, ci, vp, vp, vp, vp]
    lib.dpe_bench_ncc.argtypes = [vp, ci, ci, ci, ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.dpe_get_profile.argtypes = [vp, vp, vp, vp]
    lib.dpe_run_pipeline.argtypes = [C.c_char_p, ci, ci, ci, ci, ci, ci, ci, ci]
    _lib = lib
    return lib


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it; the launcher hands the bytes to every rank)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = load().dpe_comm_get_unique_id(buf, COMM_ID_BYTES)
    if rc != 0:
        raise DpeError(f"dpe_comm_get_unique_id failed with code {rc}")
    return buf.raw


def shard_range(n_problems, n_ranks, rank):
    """(first, count) of the block of reference views rank `rank` owns (dpe_shard_range)."""
    f, c = C.c_int(), C.c_int()
    if load().dpe_shard_range(n_problems, n_ranks, rank, C.byref(f), C.byref(c)) != 0:
        raise DpeError("dpe_shard_range: bad arguments")
    return f.value, c.value


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
            out.append((i, StageParams(REFINE_ITER, 1, int(i > 0), 3, 4, max(4 - 2 * j, 2), min(2 ** i, 4),
                                       0.01 - i * 0.00125, 0.2)))
    return out


def compute_round_num(width, height):
    """ComputeRoundNum (main.cpp:390-408)."""
    m = max(width, height)
    r = 1
    while m > 800:
        m //= 2
        r += 1
    return max(r, 2)


class DpeError(RuntimeError):
    pass


class Context:
    """Owns a dpe_ctx.  Methods mirror the C ABI one to one."""

    def __init__(self, gpu_index=0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.dpe_ctx_create(C.byref(self.h), gpu_index)
        if rc != 0:
            raise DpeError(f"dpe_ctx_create failed with code {rc} (no CUDA device / bad index)")

    def _ck(self, rc):
        if rc != 0:
            raise DpeError(f"code {rc}: {self.lib.dpe_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.dpe_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ---------------------------------------------------------------------------
    def scene_begin(self, n_views, width, height, n_scales):
        self._ck(self.lib.dpe_scene_begin(self.h, n_views, width, height, n_scales))
        self.n_views, self.n_scales = n_views, n_scales

    def set_view(self, view, gray, K, R, t, dmin, dmax):
        g = np.ascontiguousarray(gray, np.uint8)
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        self._ck(self.lib.dpe_scene_set_view(self.h, view, g.ctypes.data, K.ctypes.data, R.ctypes.data, t.ctypes.data,
                                             dmin, dmax))

    def set_pairs(self, view, src):
        s = np.ascontiguousarray(src, np.int32)
        self._ck(self.lib.dpe_scene_set_pairs(self.h, view, s.ctypes.data, len(s)))

    def set_prep(self, view, scale, edge=None, label=None):
        e = np.ascontiguousarray(edge, np.uint8) if edge is not None else None
        l = np.ascontiguousarray(label, np.int32) if label is not None else None
        n = e.size if e is not None else (l.size if l is not None else 0)
        if e is not None and l is not None and e.size != l.size:
            raise DpeError("edge and label arrays differ in size")
        self._ck(self.lib.dpe_scene_set_prep(self.h, view, scale, e.ctypes.data if e is not None else None,
                                             l.ctypes.data if l is not None else None, n))

    def set_shard(self, n_problems, rank, n_ranks):
        """This context is rank `rank` of n_ranks; it owns the block shard_range(n_problems, n_ranks, rank)."""
        self._ck(self.lib.dpe_scene_set_shard(self.h, n_problems, rank, n_ranks))

    def set_active(self, first, count):
        """Test / profiling hook: run stages over views [first, first + count) only."""
        self._ck(self.lib.dpe_scene_set_active(self.h, first, count))

    def comm_init_rank(self, unique_id: bytes, n_ranks, rank):
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._ck(self.lib.dpe_comm_init_rank(self.h, buf, n_ranks, rank))

    def broadcast_images(self, root=0):
        self._ck(self.lib.dpe_scene_broadcast_images(self.h, root))

    def set_cams(self, view, K, R, t, dmin, dmax):
        """Camera only; the image arrives by broadcast_images()."""
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        t = np.ascontiguousarray(t, np.float32).reshape(3)
        self._ck(self.lib.dpe_scene_set_view(self.h, view, None, K.ctypes.data, R.ctypes.data, t.ctypes.data, dmin, dmax))

    def commit(self):
        self._ck(self.lib.dpe_scene_commit(self.h))

    # -- stages --------------------------------------------------------------------------
    def run_stage(self, scale_idx, params, seed):
        self._ck(self.lib.dpe_run_stage(self.h, scale_idx, C.byref(params), C.c_uint64(seed)))

    def stage_begin(self, scale_idx, params, seed):
        self._ck(self.lib.dpe_stage_begin(self.h, scale_idx, C.byref(params), C.c_uint64(seed)))

    def stage_wait_view(self, view):
        self._ck(self.lib.dpe_stage_wait_view(self.h, view))

    def stage_end(self):
        self._ck(self.lib.dpe_stage_end(self.h))

    def view_slot(self, view):
        s = C.c_int()
        self._ck(self.lib.dpe_view_slot(self.h, view, C.byref(s)))
        return s.value

    def export_view(self, view, scale_idx, depth=True, normal=True, weak=True):
        """The payloads of depth.npy / normal.npy / weak.npy of a view."""
        w, h = self.size(scale_idx)
        d = np.empty((h, w), np.float32) if depth else None
        n = np.empty((h, w, 3), np.float32) if normal else None
        k = np.empty((h, w), np.int8) if weak else None
        self._ck(self.lib.dpe_export_view(self.h, view, d.ctypes.data if depth else None, n.ctypes.data if normal else None,
                                          k.ctypes.data if weak else None))
        return dict(depth=d, normal=n, weak=k)

    def stage_atlas(self):
        """(device pointer, bytes per slot, total bytes) of the atlas the last stage wrote."""
        p, slot, total = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._ck(self.lib.dpe_stage_atlas(self.h, C.byref(p), C.byref(slot), C.byref(total)))
        return p.value, slot.value, total.value

    def stage_commit(self):
        self._ck(self.lib.dpe_stage_commit(self.h))

    def size(self, scale_idx):
        w, h = C.c_int(), C.c_int()
        self._ck(self.lib.dpe_get_size(self.h, scale_idx, C.byref(w), C.byref(h)))
        return w.value, h.value

    def get_maps(self, view, scale_idx):
        w, h = self.size(scale_idx)
        depth = np.empty((h, w), np.float32)
        normal = np.empty((h, w, 3), np.float32)
        state = np.empty((h, w), np.uint8)
        sel = np.empty((h, w), np.uint32)
        self._ck(self.lib.dpe_get_maps(self.h, view, depth.ctypes.data, normal.ctypes.data, state.ctypes.data,
                                       sel.ctypes.data))
        return dict(depth=depth, normal=normal, state=state, selected=sel)

    # -- gate-1 hooks --------------------------------------------------------------------
    def cost_eval(self, view, scale_idx, xy, planes, n_src, mode=0):
        xy = np.ascontiguousarray(xy, np.int32)
        pl = np.ascontiguousarray(planes, np.float32)
        out = np.empty((len(xy), n_src), np.float32)
        self._ck(self.lib.dpe_cost_eval(self.h, view, scale_idx, len(xy), xy.ctypes.data, pl.ctypes.data, mode,
                                        out.ctypes.data))
        return out

    def geom_eval(self, view, scale_idx, xy, planes, n_src):
        xy = np.ascontiguousarray(xy, np.int32)
        pl = np.ascontiguousarray(planes, np.float32)
        out = np.empty((len(xy), n_src), np.float32)
        self._ck(self.lib.dpe_geom_eval(self.h, view, scale_idx, len(xy), xy.ctypes.data, pl.ctypes.data,
                                        out.ctypes.data))
        return out

    # -- counters / probes ---------------------------------------------------------------
    def set_count_evals(self, on):
        self._ck(self.lib.dpe_set_count_evals(self.h, int(on)))

    def eval_units(self):
        return self.lib.dpe_eval_units(self.h)

    def stage_gpu_ms(self):
        return self.lib.dpe_stage_gpu_ms(self.h)

    def stage_comm_ms(self):
        return self.lib.dpe_stage_comm_ms(self.h)

    def kernel_launches(self):
        return self.lib.dpe_kernel_launches(self.h)

    KERNEL_CLASSES = ["load", "edge_info", "nearest_strong", "gen_neighbours", "init", "strong_sweep", "fit_plane",
                      "weak_sweep", "extract", "median", "classify_refine", "finish"]

    def set_view_order(self, sequential):
        self._ck(self.lib.dpe_set_view_order(self.h, int(sequential)))

    def bench_ncc(self, view, variant, n_cand=8, reps=3):
        r, c = C.c_double(), C.c_double()
        self._ck(self.lib.dpe_bench_ncc(self.h, view, variant, n_cand, reps, C.byref(r), C.byref(c)))
        return r.value, c.value

    def fuse(self, maps, colours, blocks=None):
        """maps[v] = dict(depth, normal, state) at full resolution (None = no maps), colours[v] = HxWx3 uint8 BGR,
        blocks[v] = HxW uint8 block mask (<dense>/blocks/mask_<id>.jpg) or None.  Returns (xyz Nx3 float32, bgr Nx3 uint8)."""
        for v, b in enumerate(blocks or []):
            if b is not None:
                bb = np.ascontiguousarray(b, np.uint8)
                self._ck(self.lib.dpe_fuse_set_block(self.h, v, bb.ctypes.data))
        for v, m in enumerate(maps):
            if m is None:
                continue
            d = np.ascontiguousarray(m["depth"], np.float32); n = np.ascontiguousarray(m["normal"], np.float32)
            st = np.ascontiguousarray(m["state"], np.uint8); c = np.ascontiguousarray(colours[v], np.uint8)
            self._ck(self.lib.dpe_fuse_set_view(self.h, v, d.ctypes.data, n.ctypes.data, st.ctypes.data, c.ctypes.data))
        n = C.c_size_t()
        self._ck(self.lib.dpe_fuse_run(self.h, 0, self.n_views, C.byref(n)))
        xyz = np.empty((n.value, 3), np.float32); bgr = np.empty((n.value, 3), np.uint8)
        if n.value:
            self._ck(self.lib.dpe_fuse_get(self.h, xyz.ctypes.data, bgr.ctypes.data))
        return xyz, bgr

    def fuse_resident(self, colours):
        """Fuses the maps the last stage left on the GPU (one context owning all views); colours[v] = HxWx3 uint8 BGR."""
        self._ck(self.lib.dpe_fuse_prepare(self.h))
        keep = [np.ascontiguousarray(c, np.uint8) for c in colours]
        for v, c in enumerate(keep):
            self._ck(self.lib.dpe_fuse_set_color(self.h, v, c.ctypes.data))
        self._ck(self.lib.dpe_fuse_broadcast_colors(self.h, 0))
        n = C.c_size_t()
        self._ck(self.lib.dpe_fuse_run(self.h, 0, self.n_views, C.byref(n)))
        xyz = np.empty((n.value, 3), np.float32); bgr = np.empty((n.value, 3), np.uint8)
        if n.value:
            self._ck(self.lib.dpe_fuse_get(self.h, xyz.ctypes.data, bgr.ctypes.data))
        return xyz, bgr

    def set_cost_arithmetic(self, mode):
        """0 = centred (precise against float64), 1 = the reference's raw fp32 moments on a constant-folded
        homography (fast), 2 = the reference's arithmetic operation by operation (default; include/dpe_b200.h)."""
        self._ck(self.lib.dpe_set_cost_arithmetic(self.h, int(mode)))

    def debug_set_variants(self, mask):
        self._ck(self.lib.dpe_debug_set_variants(self.h, int(mask)))

    def set_reference_race(self, on):
        self._ck(self.lib.dpe_set_reference_race(self.h, int(on)))

    def debug_set_maps(self, view, scale_idx, planes4=None, state=None, selected=None, atlas_depth=None):
        keep = [np.ascontiguousarray(planes4, np.float32) if planes4 is not None else None,
                np.ascontiguousarray(state, np.uint8) if state is not None else None,
                np.ascontiguousarray(selected, np.uint32) if selected is not None else None,
                np.ascontiguousarray(atlas_depth, np.float32) if atlas_depth is not None else None]
        self._ck(self.lib.dpe_debug_set_maps(self.h, view, scale_idx, *[a.ctypes.data if a is not None else None for a in keep]))

    def debug_stop_after(self, step):
        self._ck(self.lib.dpe_debug_stop_after(self.h, int(step)))

    def debug_read(self, what, shape, dtype):
        out = np.empty(shape, dtype)
        self._ck(self.lib.dpe_debug_read(self.h, what, out.ctypes.data, out.nbytes))
        return out

    def set_profile(self, on):
        self._ck(self.lib.dpe_set_profile(self.h, int(on)))

    def get_profile(self):
        n = len(self.KERNEL_CLASSES)
        ms, units, launches = np.zeros(n), np.zeros(n), np.zeros(n, np.int64)
        self._ck(self.lib.dpe_get_profile(self.h, ms.ctypes.data, units.ctypes.data, launches.ctypes.data))
        return {k: dict(ms=float(ms[i]), units=float(units[i]), launches=int(launches[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def probe_tex_rate(self, w=2048, h=2048, iters=200):
        r = C.c_double()
        self._ck(self.lib.dpe_probe_tex_rate(self.h, w, h, iters, C.byref(r)))
        return r.value

    def probe_tex_pattern(self, fmt, layout, m=(1, 0, 0, 1), w=2048, h=2048, iters=200, threads=128, blocks_per_sm=4):
        r = C.c_double()
        mm = np.asarray(m, np.float32)
        self._ck(self.lib.dpe_probe_tex_pattern(self.h, fmt, layout, w, h, iters, mm.ctypes.data, threads, blocks_per_sm, C.byref(r)))
        return r.value

    def probe_fma_rate(self, iters=2000):
        r = C.c_double()
        self._ck(self.lib.dpe_probe_fma_rate(self.h, iters, C.byref(r)))
        return r.value

    def probe_tex_weights(self, n=4096):
        w = np.empty(n + 1, np.float32)
        self._ck(self.lib.dpe_probe_tex_weights(self.h, n, w.ctypes.data))
        return w


def upload_scene(ctx: Context, images, cams, depth_ranges, pairs, n_scales=None, shard=None, active=None):
    """images: list of HxW uint8; cams: list of (K,R,t); pairs: list of source-id lists;
    shard = (n_problems, rank, n_ranks); active = (first, count)."""
    H, W = images[0].shape
    if n_scales is None:
        n_scales = compute_round_num(W, H)
    ctx.scene_begin(len(images), W, H, n_scales)
    for v, (img, (K, R, t), (dmin, dmax)) in enumerate(zip(images, cams, depth_ranges)):
        ctx.set_view(v, img, K, R, t, dmin, dmax)
        ctx.set_pairs(v, pairs[v])
    if shard is not None:
        ctx.set_shard(*shard)
    if active is not None:
        ctx.set_active(*active)
    ctx.commit()
    return n_scales
