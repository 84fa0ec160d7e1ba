"""TEST INFRASTRUCTURE — float64 restatement of the reference's per-(pixel, hypothesis, view)
cost formulas.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product never does.

PARITY PINNING: the reference ships no tests, golden vectors or fixtures, so
this restatement is pinned against the reference's own device functions instead: the harness
oracle/ref_probe.cu #includes /root/reference/csrc/DPE-MVS/DPE.cu where it lies, calls
ComputeBilateralNCCOld / ComputeGeomConsistencyCost on a B200 for seeded inputs, and the
outputs are committed as tests/golden/ref_probe_*.npz (generating script:
oracle/make_golden.py).  tests/test_oracle_golden.py checks this file against them.

Each function cites the reference lines it follows (paths under csrc/DPE-MVS/).  Arithmetic
is float64 throughout; the texture unit is modelled explicitly (clamp addressing, bilinear
weights in 1.8 fixed point when quant=1 — SURVEY.md Q16).
"""
from __future__ import annotations

import numpy as np

COST_MAX = 2.0
SIGMA_SPATIAL = 5.0   # main.h:81
SIGMA_COLOR = 3.0     # main.h:82


def camera_center(R, t):
    """C = -R^T t (DPE.cu:455-462)."""
    return -(np.asarray(R, np.float64).T @ np.asarray(t, np.float64))


def compute_homography(Kr, Rr, tr, Ks, Rs, ts, plane):
    """ComputeHomography, DPE.cu:453-513.  plane = (nx, ny, nz, d), n.X + d = 0 in ref-camera
    coordinates.  Returns 3x3 H mapping reference pixels to source pixels."""
    Kr, Rr, tr, Ks, Rs, ts = (np.asarray(a, np.float64) for a in (Kr, Rr, tr, Ks, Rs, ts))
    n = np.asarray(plane[:3], np.float64)
    d = float(plane[3])
    Cr, Cs = camera_center(Rr, tr), camera_center(Rs, ts)
    R_rel = Rs @ Rr.T                          # :467-475
    t_rel = Rs @ (Cr - Cs)                     # :476-481
    Hm = R_rel - np.outer(t_rel, n) / d        # :483-491
    # multiply by Kr^-1 (zero skew, K[8] = 1 assumed by the reference)  :493-502
    tmp = np.empty((3, 3))
    tmp[:, 0] = Hm[:, 0] / Kr[0, 0]
    tmp[:, 1] = Hm[:, 1] / Kr[1, 1]
    tmp[:, 2] = -Hm[:, 0] * Kr[0, 2] / Kr[0, 0] - Hm[:, 1] * Kr[1, 2] / Kr[1, 1] + Hm[:, 2]
    H = np.empty((3, 3))                       # :504-512
    H[0] = Ks[0, 0] * tmp[0] + Ks[0, 2] * tmp[2]
    H[1] = Ks[1, 1] * tmp[1] + Ks[1, 2] * tmp[2]
    H[2] = Ks[2, 2] * tmp[2]
    return H


def corresponding_point(H, x, y):
    """ComputeCorrespondingPoint, DPE.cu:515-522 (vectorised over x, y)."""
    z = H[2, 0] * x + H[2, 1] * y + H[2, 2]
    return (H[0, 0] * x + H[0, 1] * y + H[0, 2]) / z, (H[1, 0] * x + H[1, 1] * y + H[1, 2]) / z


def tex2d_point(img, x, y):
    """tex2D<float>(img, x + 0.5, y + 0.5) for integer (x, y): the texel, clamp addressing
    (cudaAddressModeWrap + unnormalised coordinates degrades to clamp, DPE.cpp:929-933)."""
    H, W = img.shape
    return img[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)].astype(np.float64)


def tex2d_linear(img, u, v, quant=1):
    """tex2D<float>(img, u, v) with cudaFilterModeLinear, unnormalised coordinates, clamp:
    xB = u - 0.5, i = floor(xB), alpha = frac(xB); CUDA stores alpha, beta in 9-bit fixed
    point with 8 fractional bits (quant=1: rounded to nearest 1/256; 2: truncated; 0: exact)."""
    H, W = img.shape
    xb, yb = np.asarray(u, np.float64) - 0.5, np.asarray(v, np.float64) - 0.5
    x0, y0 = np.floor(xb), np.floor(yb)
    a, b = xb - x0, yb - y0
    if quant == 1:
        a, b = np.floor(a * 256 + 0.5) / 256, np.floor(b * 256 + 0.5) / 256
    elif quant == 2:
        a, b = np.floor(a * 256) / 256, np.floor(b * 256) / 256
    x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
    xa, xb_ = np.clip(x0, 0, W - 1), np.clip(x0 + 1, 0, W - 1)
    ya, yb_ = np.clip(y0, 0, H - 1), np.clip(y0 + 1, 0, H - 1)
    im = img.astype(np.float64)
    return (1 - b) * ((1 - a) * im[ya, xa] + a * im[ya, xb_]) + b * ((1 - a) * im[yb_, xa] + a * im[yb_, xb_])


def bilateral_weight(dx, dy, pix, center_pix):
    """ComputeBilateralWeight, DPE.cu:550-555."""
    return np.exp(-np.sqrt(dx * dx + dy * dy) / (2 * SIGMA_SPATIAL ** 2) - np.abs(pix - center_pix) / (2 * SIGMA_COLOR ** 2))


def _patch_ncc(ref_img, src_img, H, cx, cy, center_pix, radius, increment, quant):
    """The tap loop + finalisation shared by NCCOld (DPE.cu:724-774) and NCCNew (624-668)."""
    offs = np.arange(-radius, radius + 1, increment)
    di, dj = np.meshgrid(offs, offs, indexing="ij")      # i: x offset (outer), j: y offset
    rx, ry = cx + di, cy + dj
    ref_pix = tex2d_point(ref_img, rx, ry)
    sx, sy = corresponding_point(H, rx.astype(np.float64), ry.astype(np.float64))
    src_pix = tex2d_linear(src_img, sx + 0.5, sy + 0.5, quant)
    w = bilateral_weight(di.astype(np.float64), dj.astype(np.float64), ref_pix, center_pix)
    sw = w.sum()
    m_r, m_s = (w * ref_pix).sum() / sw, (w * src_pix).sum() / sw
    var_r = (w * ref_pix * ref_pix).sum() / sw - m_r * m_r
    var_s = (w * src_pix * src_pix).sum() / sw - m_s * m_s
    if var_r < 1e-5 or var_s < 1e-5:
        return COST_MAX
    cov = (w * ref_pix * src_pix).sum() / sw - m_r * m_s
    return float(max(0.0, min(COST_MAX, 1.0 - cov / np.sqrt(var_r * var_s))))


def bilateral_ncc_old(ref_img, src_img, ref_cam, src_cam, x, y, plane, quant=1):
    """ComputeBilateralNCCOld, DPE.cu:692-778.  cams = (K, R, t) at this scale.
    6x6 taps {-5,-3,-1,1,3,5}^2 (strong_radius 5, strong_increment 2; the centre pixel is
    never a tap)."""
    H = compute_homography(*ref_cam, *src_cam, plane)
    px, py = corresponding_point(H, float(x), float(y))
    sh, sw = src_img.shape
    if px >= sw or px < 0.0 or py >= sh or py < 0.0:         # :708-710
        return COST_MAX
    center = float(tex2d_point(ref_img, np.int64(x), np.int64(y)))
    return _patch_ncc(ref_img, src_img, H, x, y, center, 5, 2, quant)


def bilateral_ncc_new(ref_img, src_img, ref_cam, src_cam, x, y, plane, anchors, radius_p,
                      anchor_selected, view_bit, quant=1):
    """ComputeBilateralNCCNew, DPE.cu:557-690 (WEAK pixels).  anchors: 9 x (ax, ay) with
    anchors[0] = (x, y), absent = (-1, -1); radius_p = radius_cuda[p]; anchor_selected[k] =
    selected_views bitmask of anchor k."""
    H = compute_homography(*ref_cam, *src_cam, plane)
    px, py = corresponding_point(H, float(x), float(y))
    sh, sw = src_img.shape
    rh, rw = ref_img.shape
    if px >= sw or px < 0.0 or py >= sh or py < 0.0:
        return COST_MAX
    center = float(tex2d_point(ref_img, np.int64(x), np.int64(y)))
    center_cost, strong_cost, strong_count = 0.0, 0.0, 0
    for k in range(9):
        ax, ay = int(anchors[k][0]), int(anchors[k][1])
        if ax == -1 or ay == -1:
            continue
        qx, qy = corresponding_point(H, float(ax), float(ay))
        if qx < 0 or qy < 0 or qx >= rw or qy >= rh:          # :596 (reference-image size)
            if k != 0:
                if (int(anchor_selected[k]) >> view_bit) & 1:
                    strong_cost += COST_MAX
                    strong_count += 1
                continue
            return COST_MAX
        if k == 0:
            radius = int(radius_p)
            inc = max(2, int(2.0 * radius / 5.0))             # :619-622
        else:
            radius, inc = 5, 5                                # weak_radius, weak_increment
        c = _patch_ncc(ref_img, src_img, H, ax, ay, center, radius, inc, quant)
        if k == 0:
            center_cost = c
        else:
            strong_cost += c
            strong_count += 1
    if strong_count == 0:
        return center_cost
    strong_cost = min(strong_cost / strong_count, COST_MAX)
    return 0.25 * center_cost + 0.75 * strong_cost


def depth_from_plane(K, plane, x, y):
    """ComputeDepthfromPlaneHypothesis, DPE.cu:356-359."""
    K = np.asarray(K, np.float64)
    return -plane[3] * K[0, 0] / ((x - K[0, 2]) * plane[0] + (K[0, 0] / K[1, 1]) * (y - K[1, 2]) * plane[1] + K[0, 0] * plane[2])


def distance_to_origin(K, x, y, depth, normal):
    """GetDistance2Origin, DPE.cu:337-342."""
    K = np.asarray(K, np.float64)
    X = depth * np.array([(x - K[0, 2]) / K[0, 0], (y - K[1, 2]) / K[1, 1], 1.0])
    return -float(np.dot(np.asarray(normal, np.float64)[:3], X))


def point_on_world(K, R, t, x, y, depth):
    """Get3DPointonWorld_cu, DPE.cu:881-901 (camera centre c = -R^T t)."""
    K, R, t = (np.asarray(a, np.float64) for a in (K, R, t))
    X = np.array([depth * (x - K[0, 2]) / K[0, 0], depth * (y - K[1, 2]) / K[1, 1], depth])
    return R.T @ X + camera_center(R, t)


def project_on_camera(K, R, t, X):
    """ProjectonCamera_cu, DPE.cu:903-913."""
    K, R, t = (np.asarray(a, np.float64) for a in (K, R, t))
    tmp = R @ X + t
    d = K[2] @ tmp
    return (K[0] @ tmp) / d, (K[1] @ tmp) / d, d


def geom_consistency_cost(ref_cam, src_cam, src_depth, x, y, plane):
    """ComputeGeomConsistencyCost, DPE.cu:915-953."""
    depth = depth_from_plane(ref_cam[0], plane, x, y)
    Xw = point_on_world(*ref_cam, x, y, depth)
    u, v, _ = project_on_camera(*src_cam, Xw)
    H, W = src_depth.shape
    iu = int(np.clip(np.trunc(u), 0, W - 1)) if np.isfinite(u) else 0   # (int)src_pt.x + clamp
    iv = int(np.clip(np.trunc(v), 0, H - 1)) if np.isfinite(v) else 0
    sd = float(src_depth[iv, iu])
    if sd == 0.0:
        return 3.0
    Xs = point_on_world(*src_cam, u, v, sd)
    bx, by, _ = project_on_camera(*ref_cam, Xs)
    return float(min(3.0, np.hypot(x - bx, y - by)))


def scale_camera(K, w, h, full_w, full_h):
    """Intrinsics at a pyramid level, in float32 like DPE.cpp:804-817."""
    K = np.array(K, np.float32).copy()
    if (w, h) != (full_w, full_h):
        sx, sy = np.float32(w) / np.float32(full_w), np.float32(h) / np.float32(full_h)
        K[0, 0] *= sx; K[0, 2] *= sx; K[1, 1] *= sy; K[1, 2] *= sy
    return K
