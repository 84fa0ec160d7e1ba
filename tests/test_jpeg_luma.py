"""The host JPEG luma decoder (csrc/host/jpeg_luma.cpp: Huffman + integer dequantisation + jpeg_idct_islow) against
cv2.imread(..., IMREAD_GRAYSCALE), i.e. libjpeg-turbo — the pixels the reference feeds PatchMatch (DPE.cpp:745).
Bit-identical on greyscale and colour files, subsampled or not, any size, with restart intervals.  CPU only."""
import ctypes as C

import cv2
import numpy as np
import pytest

import capi


def _decode(path):
    lib = capi.load()
    raw = np.fromfile(path, np.uint8)
    out = np.zeros(1 << 24, np.uint8)
    w, h = C.c_int(), C.c_int()
    rc = lib.dpe_host_decode_luma_islow(raw.ctypes.data_as(C.c_void_p), C.c_long(raw.size), out.ctypes.data_as(C.c_void_p), C.c_long(out.size),
                                        C.byref(w), C.byref(h))
    if rc != 0:
        return None
    return out[:w.value * h.value].reshape(h.value, w.value)


def _image(h, w, color, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 128 + 60 * np.sin(xx / 7.3 + seed) * np.cos(yy / 5.1) + 30 * np.sin((xx + yy) / 23.0)
    img = base[..., None] + rng.normal(0, 12, (h, w, 3 if color else 1)) + (np.array([20, -10, 5]) if color else 0)
    img = np.clip(img, 0, 255).astype(np.uint8)
    img[: h // 8] = 255; img[-h // 8:] = 0          # saturated regions: exercises the range limit
    return img if color else img[..., 0]


CASES = [(120, 160, False, [cv2.IMWRITE_JPEG_QUALITY, 98]), (97, 131, False, [cv2.IMWRITE_JPEG_QUALITY, 75]),
         (120, 160, True, [cv2.IMWRITE_JPEG_QUALITY, 95]), (101, 67, True, [cv2.IMWRITE_JPEG_QUALITY, 50]),
         (64, 200, True, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 3]),
         (33, 45, False, [cv2.IMWRITE_JPEG_QUALITY, 100, cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
         (80, 96, True, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, 0x111111])]


@pytest.mark.parametrize("h,w,color,params", CASES)
def test_luma_decoder_equals_libjpeg(tmp_path, h, w, color, params):
    p = str(tmp_path / "t.jpg")
    try:
        ok = cv2.imwrite(p, _image(h, w, color, h + w), params)
    except cv2.error:
        pytest.skip("this OpenCV build does not know the parameter")
    assert ok
    want = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
    got = _decode(p)
    assert got is not None and got.shape == want.shape
    assert np.array_equal(got, want), (int(np.abs(got.astype(int) - want.astype(int)).max()), float((got != want).mean()))


def test_progressive_files_are_declined(tmp_path):
    p = str(tmp_path / "p.jpg")
    assert cv2.imwrite(p, _image(64, 64, True, 1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert _decode(p) is None        # the pipeline then falls back to nvJPEG


def test_frame_size_and_path_decode_need_no_gpu(tmp_path):
    """the size comes from the SOFn marker and baseline files decode on the host: neither touches nvJPEG (whose handle is
    only created for colour, progressive files and DPE_JPEG_DECODER=nvjpeg), so both work in a process without a GPU"""
    lib = capi.load()
    for i, (h, w, color, params) in enumerate([(97, 131, False, [cv2.IMWRITE_JPEG_QUALITY, 75]), (64, 200, True, [cv2.IMWRITE_JPEG_QUALITY, 90]),
                                               (48, 80, True, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])]):
        p = str(tmp_path / f"s{i}.jpg")
        assert cv2.imwrite(p, _image(h, w, color, 3 + i), params)
        ww, hh = C.c_int(), C.c_int()
        assert lib.dpe_host_jpeg_size(p.encode(), C.byref(ww), C.byref(hh)) == 0
        assert (hh.value, ww.value) == (h, w)
        if i < 2:
            out = np.zeros(h * w, np.uint8)
            assert lib.dpe_host_decode_gray(p.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(ww), C.byref(hh)) == 0
            assert np.array_equal(out.reshape(h, w), cv2.imread(p, cv2.IMREAD_GRAYSCALE))
    bad = tmp_path / "bad.jpg"
    bad.write_bytes(b"\xff\xd8\xff\xd9")
    assert lib.dpe_host_jpeg_size(str(bad).encode(), C.byref(ww), C.byref(hh)) != 0


def test_damaged_files_are_refused_or_decoded_within_bounds(tmp_path):
    """Truncated files, flipped bytes, runs of 0xFF, damaged headers (marker lengths, sampling factors, dimensions): the
    decoder returns an error or some image, never writes past the capacity it was given (guard bytes behind it stay
    intact), and still decodes the undamaged file afterwards.  The same corpus ran clean under
    -fsanitize=address,undefined (the IDCT is built with -fwrapv: damaged coefficients may wrap its 32-bit sums)."""
    lib = capi.load()
    rng = np.random.default_rng(7)
    w, h = C.c_int(), C.c_int()
    n_ok = n_bad = 0
    for it in range(400):
        color = bool(it & 1)
        params = [cv2.IMWRITE_JPEG_QUALITY, 60 + it % 40] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, 2] if it % 5 == 0 else [])
        img = _image(17 + it % 37, 23 + it % 29, color, it)
        ok, enc = cv2.imencode(".jpg", img, params)
        assert ok
        raw = enc.copy().ravel()
        mode = it % 4
        if mode == 0:
            raw = raw[: rng.integers(2, raw.size)]
        elif mode == 1:
            k = rng.integers(1, 8)
            raw[rng.integers(0, raw.size, k)] = rng.integers(0, 256, k)
        elif mode == 2:
            i = rng.integers(0, raw.size - 4)
            raw[i:i + 4] = 0xFF
        else:
            k = rng.integers(1, 4)
            raw[rng.integers(2, min(raw.size, 600), k)] = rng.integers(0, 256, k)
        raw = np.ascontiguousarray(raw)
        for cap in (img.shape[0] * img.shape[1], 257):
            out = np.full(cap + 64, 0xA5, np.uint8)
            rc = lib.dpe_host_decode_luma_islow(raw.ctypes.data_as(C.c_void_p), C.c_long(raw.size), out.ctypes.data_as(C.c_void_p),
                                                C.c_long(cap), C.byref(w), C.byref(h))
            assert (out[cap:] == 0xA5).all(), (it, cap)
            if rc == 0:
                assert w.value * h.value <= cap
            if cap > 257:
                n_ok += rc == 0
                n_bad += rc != 0
    assert n_ok > 50 and n_bad > 50, (n_ok, n_bad)      # both outcomes occur
    p = str(tmp_path / "good.jpg")
    assert cv2.imwrite(p, _image(64, 80, True, 3), [cv2.IMWRITE_JPEG_QUALITY, 90])
    assert np.array_equal(_decode(p), cv2.imread(p, cv2.IMREAD_GRAYSCALE))
