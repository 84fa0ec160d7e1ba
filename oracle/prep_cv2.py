"""TEST INFRASTRUCTURE — the reference's edge / label preparation restated in Python on top
of the REAL OpenCV (cv2 4.13, the only OpenCV in this image) for the library calls and
oracle/prep_oracle.c for the two hand-written routines.  It follows
  GetProblemEdges  /root/reference/csrc/DPE-MVS/main.cpp:331-388
  EdgeSegment      /root/reference/csrc/DPE-MVS/DPE.cpp:136-291
and is used (a) to pre-write DPE/<view>/edges_k.dmb / labels_k.dmb for the reference build
(which honours them as a cache, main.cpp:351-355, 370-374, so its stubbed Canny/Hough are
never reached) and (b) as the checker for the product's own C++ prep.
Only tests/, bench.py's reference arm and tools/ import this; the product never does.
"""
from __future__ import annotations

import ctypes as C
import struct
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
CV_8UC1, CV_32SC1, CV_32FC1, CV_32FC3 = 0, 4, 5, 21

_lib = None


def _oracle_lib():
    global _lib
    if _lib is None:
        so = HERE / "_ref" / "libprep_oracle.so"
        src = HERE / "prep_oracle.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            so.parent.mkdir(exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", str(src), "-o", str(so), "-lm"])
        _lib = C.CDLL(str(so))
    return _lib


def roberts(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    _oracle_lib().oracle_roberts(img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], out.ctypes.data_as(C.c_void_p))
    return out


def connect(img):
    """Returns (label int32 HxW, counts)."""
    img = np.ascontiguousarray(img, np.uint8)
    lab = np.empty(img.shape, np.int32)
    cnt = np.empty(img.size // 2 + 4, np.int32)
    n = _oracle_lib().oracle_connect(img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1],
                                     lab.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
    return lab, cnt[:n].copy()


def _cround(x):
    """std::round: half away from zero."""
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def _border_clean(dst):
    """DPE.cpp:239-250"""
    rows, cols = dst.shape
    for y in range(rows):
        if dst[y, 1] == 0:
            dst[y, 0] = 0
        if dst[y, cols - 2] == 0:
            dst[y, cols - 1] = 0
    for x in range(cols):
        if dst[1, x] == 0:
            dst[0, x] = 0
        if dst[rows - 2, x] == 0:
            dst[rows - 1, x] = 0
    return dst


def edge_segment(scale, src_image, mode, use_canny, hough=True):
    """EdgeSegment(scale, srcImage, mode, use_canny, high_res_img=true), DPE.cpp:136-291.
    mode 0 -> edge map (uint8 0/255), mode 1 -> label map (int32)."""
    import cv2
    robthr = 4
    rows, cols = src_image.shape
    weak_tex_num = int(1.0 * rows * cols / ((1024 << scale) << scale))
    if not use_canny:
        src_down = cv2.resize(src_image, (cols // 2, rows // 2), interpolation=cv2.INTER_LINEAR)
        src_down = cv2.resize(src_down, (src_down.shape[1] // 2, src_down.shape[0] // 2), interpolation=cv2.INTER_LINEAR)
        m = int(min(src_down.shape[1], src_down.shape[0]))
        houthr = int(m / 30.0)
        dst = roberts(src_down)
        _, dst = cv2.threshold(dst, robthr, 255, cv2.THRESH_BINARY)
        lab0, cnt0 = connect(dst)
        if hough:
            for k in range(1, len(cnt0)):
                if cnt0[k] < weak_tex_num:
                    continue
                inside = lab0 == k
                nb = np.zeros_like(inside)
                nb[:, 1:] |= inside[:, :-1]
                nb[:, :-1] |= inside[:, 1:]
                nb[1:, :] |= inside[:-1, :]
                nb[:-1, :] |= inside[1:, :]
                img_weak = np.where(nb & ~inside, 255, 0).astype(np.uint8)
                lines = cv2.HoughLinesP(img_weak, 1, np.pi / 180, houthr, minLineLength=houthr, maxLineGap=houthr)
                if lines is not None:
                    for l in lines.reshape(-1, 4):
                        cv2.line(dst, (int(l[0]), int(l[1])), (int(l[2]), int(l[3])), 255, 1)
    else:
        hist = np.bincount(src_image.reshape(-1), minlength=256).astype(np.float32)
        half = rows * cols // 2
        median_val, acc = -1, 0
        for i in range(255):
            acc = acc + int(hist[i])
            if acc > half:
                median_val = i
                break
        sigma = np.float32(0.67)
        threshold1 = int((np.float32(1) - sigma) * np.float32(median_val))
        threshold2 = median_val
        dst = cv2.Canny(src_image, threshold1, threshold2, apertureSize=3, L2gradient=True)
    if mode == 0:
        dst = cv2.resize(dst, (cols, rows), interpolation=cv2.INTER_LINEAR)
    else:
        factor = np.float32(1.0) / np.float32(1 << scale)
        new_cols, new_rows = _cround(float(np.float32(cols) * factor)), _cround(float(np.float32(rows) * factor))
        dst = cv2.resize(dst, (new_cols, new_rows), interpolation=cv2.INTER_LINEAR)
    _, dst = cv2.threshold(dst, robthr, 255, cv2.THRESH_BINARY)
    dst = _border_clean(np.ascontiguousarray(dst))
    if mode == 0:
        return dst
    lab, cnt = connect(dst)
    small = (cnt[lab] <= weak_tex_num) & (lab != 0)
    lab = lab.copy()
    lab[small] = -1
    return lab


def scaled_gray(gray_u8, scale_size):
    """main.cpp:338-346: float convert, bilinear resize, back to uint8 (saturate_cast rounds
    half to even)."""
    import cv2
    src = gray_u8.astype(np.float32)
    factor = np.float32(1.0) / np.float32(scale_size)
    new_cols = _cround(float(np.float32(src.shape[1]) * factor))
    new_rows = _cround(float(np.float32(src.shape[0]) * factor))
    scaled = cv2.resize(src, (new_cols, new_rows), interpolation=cv2.INTER_LINEAR)
    return np.clip(np.rint(scaled), 0, 255).astype(np.uint8), scaled


def problem_edges(gray_u8, scale_size, hough=True):
    """GetProblemEdges for one view and one scale_size (1, 2, 4...).  Returns (scale, edge, label)."""
    scale = 0
    while (1 << scale) < scale_size:
        scale += 1
    src_img, _ = scaled_gray(gray_u8, scale_size)
    edge = edge_segment(scale, src_img, 0, True)
    label = edge_segment(scale, gray_u8, 1, False, hough=hough)
    return scale, edge, label


def write_dmb(path, arr):
    """WriteBinMat, DPE.cpp:320-339."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype == np.uint8:
        t = CV_8UC1
    elif arr.dtype == np.int32:
        t = CV_32SC1
    elif arr.dtype == np.float32 and arr.ndim == 2:
        t = CV_32FC1
    elif arr.dtype == np.float32 and arr.ndim == 3 and arr.shape[2] == 3:
        t = CV_32FC3
    else:
        raise ValueError("unsupported dmb type")
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", 1, arr.shape[0], arr.shape[1], t))
        f.write(arr.tobytes())


def read_dmb(path):
    """ReadBinMat, DPE.cpp:293-318."""
    with open(path, "rb") as f:
        version, rows, cols, t = struct.unpack("<iiii", f.read(16))
        assert version == 1
        dt, cn = {CV_8UC1: (np.uint8, 1), CV_32SC1: (np.int32, 1), CV_32FC1: (np.float32, 1), CV_32FC3: (np.float32, 3)}[t]
        a = np.frombuffer(f.read(), dtype=dt)
    return a.reshape(rows, cols) if cn == 1 else a.reshape(rows, cols, cn)
