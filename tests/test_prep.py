"""Host prep (dpe-mvs_b200/csrc/host/prep.cpp) against the real OpenCV (cv2) and against the
oracle restatement of GetProblemEdges (oracle/prep_cv2.py).  CPU only."""
import ctypes as C

import cv2
import numpy as np
import pytest

import capi
import prep_cv2


@pytest.fixture(scope="module")
def lib():
    return capi.load()


def _img(rng, h, w, smooth=3):
    a = rng.normal(0, 1, (h, w)).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), smooth)
    a = (a - a.min()) / (a.max() - a.min())
    return (a * 255).astype(np.uint8)


def _resize_u8(lib, img, dw, dh):
    out = np.empty((dh, dw), np.uint8)
    lib.dpe_host_resize_u8(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], out.ctypes.data_as(C.c_void_p), dw, dh)
    return out


@pytest.mark.parametrize("shape,dsize", [((120, 160), (80, 60)), ((121, 161), (80, 60)), ((60, 80), (160, 120)),
                                         ((75, 100), (400, 300)), ((96, 128), (128, 96)), ((90, 131), (65, 45))])
def test_resize_u8_matches_cv2(lib, shape, dsize):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    ours = _resize_u8(lib, img, *dsize)
    ref = cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ours, ref)


@pytest.mark.parametrize("shape,dsize", [((120, 160), (80, 60)), ((120, 160), (40, 30)), ((121, 161), (81, 61))])
def test_resize_f32_matches_cv2(lib, shape, dsize):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, shape).astype(np.float32)
    out = np.empty((dsize[1], dsize[0]), np.float32)
    lib.dpe_host_resize_f32(img.ctypes.data_as(C.c_void_p), shape[1], shape[0], out.ctypes.data_as(C.c_void_p), dsize[0], dsize[1])
    ref = cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR)
    # exact for the 2^k decimations the pyramid uses; non-integer scales differ from OpenCV's SIMD
    # path only by float association (relative ~1e-5)
    tol = 0.0 if (shape[1] % dsize[0] == 0 and shape[0] % dsize[1] == 0) else 5e-3
    assert np.abs(out - ref).max() <= tol


@pytest.mark.parametrize("seed,low,high", [(0, 20, 60), (1, 33, 100), (2, 5, 15), (3, 60, 30)])
def test_canny_matches_cv2(lib, seed, low, high):
    rng = np.random.default_rng(seed)
    img = _img(rng, 150, 200, smooth=2 + seed)
    out = np.empty_like(img)
    lib.dpe_host_canny.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.dpe_host_canny(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], float(low), float(high), out.ctypes.data_as(C.c_void_p))
    ref = cv2.Canny(img, low, high, apertureSize=3, L2gradient=True)
    assert np.array_equal(out, ref)


def test_hough_and_line_match_cv2(lib):
    rng = np.random.default_rng(5)
    img = np.zeros((120, 160), np.uint8)
    for _ in range(6):
        p = rng.integers(5, 115, 4)
        cv2.line(img, (int(p[0]) + 20, int(p[1])), (int(p[2]) + 20, int(p[3])), 255, 1)
    img[rng.random(img.shape) < 0.01] = 255
    out = np.zeros((256, 4), np.int32)
    n = lib.dpe_host_hough(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], 12, 12, 12, out.ctypes.data_as(C.c_void_p), 256)
    ref = cv2.HoughLinesP(img, 1, np.pi / 180, 12, minLineLength=12, maxLineGap=12)
    ref = np.zeros((0, 4), np.int32) if ref is None else ref.reshape(-1, 4)
    assert n == len(ref)
    assert np.array_equal(out[:n], ref)
    # cv::line
    a = np.zeros((50, 60), np.uint8)
    b = a.copy()
    for (x0, y0, x1, y1) in [(3, 4, 50, 20), (50, 20, 3, 4), (10, 45, 12, 2), (5, 5, 5, 40), (2, 30, 58, 30), (40, 40, 10, 10)]:
        lib.dpe_host_line(a.ctypes.data_as(C.c_void_p), 60, 50, x0, y0, x1, y1)
        cv2.line(b, (x0, y0), (x1, y1), 255, 1)
    assert np.array_equal(a, b)


def _weak_scene_image():
    import synth
    spec = synth.make_scene("c4", scale=0.125, n_views=2)   # 378 x 252, weak planes + frames
    return synth.render_view(spec, 0)[0]


@pytest.mark.parametrize("scale_size", [1, 2])
def test_problem_edges_matches_oracle(lib, scale_size):
    img = np.ascontiguousarray(_weak_scene_image())
    h, w = img.shape
    _, e_ref, l_ref = prep_cv2.problem_edges(img, scale_size)
    e = np.empty(e_ref.shape, np.uint8)
    l = np.empty(l_ref.shape, np.int32)
    lib.dpe_host_problem_edges(img.ctypes.data_as(C.c_void_p), w, h, scale_size, e.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p))
    assert np.array_equal(e, e_ref)
    assert (l > 0).any(), "the weak-texture test image must contain labelled regions"
    assert np.array_equal(l, l_ref)


def test_degenerate_image_sizes_do_not_overrun():
    """Images a pixel or two wide, flat / binary / noisy, at scale sizes 1, 2, 4 — a target size can round to zero
    (cv::resize would throw) and the reference's border clean-up indexes column 1 and row 1 unconditionally
    (DPE.cpp:239-250).  The prep must come back with arrays of the rounded size; guard bytes behind them stay intact.
    (The same cases ran clean under -fsanitize=address,undefined.)"""
    lib = capi.load()
    rng = np.random.default_rng(11)
    for (h, w) in [(1, 1), (1, 5), (5, 1), (2, 2), (3, 2), (2, 7), (4, 4), (9, 3)]:
        for kind in range(3):
            img = (rng.integers(0, 256, (h, w)) if kind == 0 else np.full((h, w), 128) if kind == 1 else rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
            img = np.ascontiguousarray(img)
            for ss in (1, 2, 4):
                f = np.float32(1.0) / np.float32(ss)
                oc, orr = int(np.floor(float(np.float32(w) * f) + 0.5)), int(np.floor(float(np.float32(h) * f) + 0.5))     # std::round
                n = max(oc * orr, 0)
                e = np.full(n + 16, 0xA5, np.uint8)
                l = np.full(n + 16, 0x5A5A5A5A, np.int32)
                lib.dpe_host_problem_edges(img.ctypes.data_as(C.c_void_p), w, h, ss, e.ctypes.data_as(C.c_void_p), l.ctypes.data_as(C.c_void_p))
                assert (e[n:] == 0xA5).all() and (l[n:] == 0x5A5A5A5A).all(), (h, w, kind, ss)
                assert set(np.unique(e[:n])) <= {0, 255}
