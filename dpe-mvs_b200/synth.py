"""Synthetic textured-plane scenes with exact ground truth, written in the colmap2mvsnet
layout the reference consumes (images/%08d.jpg, cams/%08d_cam.txt, pair.txt; layout as
produced by src/DPE_MVS/colmap2mvsnet.py:454-494 and parsed by main.cpp:264-308 and
DPE.cpp:341-382).  DTU / ETH3D / Tanks&Temples data is not available offline, so every
configuration of BASELINE.json is rendered here in the named shape:

  c1   5 views  640x480   N=4   one textured slanted plane + floor
  c2   49 views 1600x1200 N=10  DTU-shape: 7x7 pose grid over a floor with panels
  c4   20 views 3024x2016 N=10  ETH3D-shape: >=60 % low-texture planes with textured frames
  c5   300 views 1920x1080 N=10 Tanks&Temples-shape: closed ring around a box on a floor

World: z-up, floor z = 0, finite rectangular panels.  Albedo per plane: band-limited noise,
6 octaves of random-phase sinusoids, mean 128, sigma ~40 (or ~1 for weak planes).
Rendering is analytic ray/plane intersection, 2x2 supersampled; depth and normal are
exact at the pixel centre.  Rendering uses torch (GPU when present) as plumbing only.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import torch

SEED = 20261018


@dataclass
class Plane:
    origin: np.ndarray          # a point on the plane
    e1: np.ndarray              # in-plane unit axes
    e2: np.ndarray
    half: tuple                 # half sizes along e1, e2 (inf for the floor)
    amp: float = 8.2            # amplitude of each sinusoid
    ramp: float = 0.0           # linear shading ramp (grey levels per world unit)
    base: float = 128.0
    tex: dict = field(default_factory=dict)

    @property
    def normal(self):
        n = np.cross(self.e1, self.e2)
        return n / np.linalg.norm(n)


@dataclass
class SceneSpec:
    name: str
    width: int
    height: int
    n_views: int
    n_src: int
    planes: list
    cams: list                  # list of (K, R, t)
    look_dist: float


def _look_at(C, target, up=np.array([0.0, 0.0, 1.0])):
    z = target - C
    z = z / np.linalg.norm(z)
    x = np.cross(z, up)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])      # rows: camera axes in world coords (X_cam = R X_world + t)
    t = -R @ C
    return R, t


def _make_tex(rng, f_max, amp):
    f1, f2, ph = [], [], []
    for o in range(6):
        f = f_max / (2 ** o)
        for _ in range(8):
            a = rng.uniform(0, 2 * math.pi)
            f1.append(f * math.cos(a))
            f2.append(f * math.sin(a))
            ph.append(rng.uniform(0, 2 * math.pi))
    return dict(f1=np.array(f1), f2=np.array(f2), ph=np.array(ph), amp=amp)


def make_scene(config: str, scale: float = 1.0, n_views: int | None = None) -> SceneSpec:
    """scale < 1 shrinks the image size (for CPU tests); geometry is unchanged."""
    rng = np.random.default_rng(SEED)
    cfg = config.lower()
    shapes = {"c1": (640, 480, 5, 4), "c2": (1600, 1200, 49, 10), "c4": (3024, 2016, 20, 10),
              "c5": (1920, 1080, 300, 10)}
    W, H, V, N = shapes[cfg]
    W, H = int(round(W * scale)), int(round(H * scale))
    if n_views is not None:
        V = n_views
        N = min(N, V - 1)
    f = 0.9 * W
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1.0]])
    dist = 3.0
    f_max = 0.9 * W / (4.0 * dist)
    ex, ey, ez = np.eye(3)
    planes = []
    target = np.array([0.0, 0.0, 0.3])
    cams = []
    if cfg == "c1":
        planes.append(Plane(np.zeros(3), ex, ey, (math.inf, math.inf)))
        # one slanted textured panel in front of the floor
        e1 = np.array([1.0, 0.0, 0.0]); e2 = np.array([0.0, math.sin(0.5), math.cos(0.5)])
        planes.append(Plane(np.array([0.0, 0.2, 0.5]), e1, e2, (0.9, 0.6)))
        for i in range(V):
            a = math.radians(-20 + 40 * i / max(V - 1, 1))
            C = target + dist * np.array([math.sin(a), -math.cos(a) * math.cos(0.6), math.sin(0.6)])
            C += rng.normal(0, 0.02, 3)
            cams.append((K, *_look_at(C, target)))
    elif cfg in ("c2", "c4"):
        weak = cfg == "c4"
        planes.append(Plane(np.zeros(3), ex, ey, (math.inf, math.inf), amp=0.25 if weak else 8.2,
                            ramp=2.0 if weak else 0.0))
        # panels: boxes' faces approximated by tilted rectangles
        specs = [((-0.8, 0.3, 0.45), 0.4, 0.2, (0.5, 0.45)), ((0.7, -0.2, 0.35), -0.5, 0.3, (0.45, 0.35)),
                 ((0.0, 0.9, 0.6), 0.0, 0.9, (0.8, 0.5)), ((-0.2, -0.8, 0.25), 0.8, 0.15, (0.4, 0.25))]
        for (o, yaw, tilt, half) in specs:
            e1 = np.array([math.cos(yaw), math.sin(yaw), 0.0])
            up = np.array([-math.sin(yaw) * math.sin(tilt), math.cos(yaw) * math.sin(tilt), math.cos(tilt)])
            planes.append(Plane(np.array(o), e1, up, half, amp=0.25 if weak else 8.2, ramp=3.0 if weak else 0.0,
                                base=150.0 if weak else 128.0))
            if weak:  # textured frame around each weak panel (so Canny finds the border)
                for sgn, ax in ((1, 0), (-1, 0), (1, 1), (-1, 1)):
                    hh = (0.04, half[1] + 0.04) if ax == 0 else (half[0] + 0.04, 0.04)
                    off = (e1 * sgn * (half[0] + 0.04)) if ax == 0 else (up * sgn * (half[1] + 0.04))
                    planes.append(Plane(np.array(o) + off + 0.002 * np.cross(e1, up), e1, up, hh))
        if weak:  # textured stripes on the floor bound the weak floor regions
            for k in range(-3, 4):
                planes.append(Plane(np.array([0.6 * k, 0.0, 0.002]), ex, ey, (0.03, 3.0)))
                planes.append(Plane(np.array([0.0, 0.6 * k, 0.002]), ex, ey, (3.0, 0.03)))
        g = int(round(math.sqrt(V)))
        idx = 0
        for i in range(V):
            if g * g == V:
                gi, gj = divmod(i, g)
                az = math.radians(-24 + 48 * gj / max(g - 1, 1))
                el = math.radians(35 + 30 * gi / max(g - 1, 1))
            else:
                az = math.radians(-30 + 60 * (i % 5) / 4.0)
                el = math.radians(35 + 30 * (i // 5) / max((V - 1) // 5, 1))
            C = target + dist * np.array([math.sin(az) * math.cos(el), -math.cos(az) * math.cos(el), math.sin(el)])
            C += rng.normal(0, 0.01, 3)
            cams.append((K, *_look_at(C, target)))
            idx += 1
    elif cfg == "c5":
        planes.append(Plane(np.zeros(3), ex, ey, (math.inf, math.inf)))
        # a box: 4 side faces + top
        bx, by, bz = 0.6, 0.4, 0.7
        planes.append(Plane(np.array([0, -by, bz / 2]), ex, ez, (bx, bz / 2)))
        planes.append(Plane(np.array([0, by, bz / 2]), -ex, ez, (bx, bz / 2)))
        planes.append(Plane(np.array([bx, 0, bz / 2]), ey, ez, (by, bz / 2)))
        planes.append(Plane(np.array([-bx, 0, bz / 2]), -ey, ez, (by, bz / 2)))
        planes.append(Plane(np.array([0, 0, bz]), ex, ey, (bx, by)))
        for i in range(V):
            az = 2 * math.pi * i / V
            el = math.radians(30 + 10 * math.sin(3 * az))
            C = target + dist * np.array([math.sin(az) * math.cos(el), -math.cos(az) * math.cos(el), math.sin(el)])
            cams.append((K, *_look_at(C, target)))
    else:
        raise ValueError(config)
    for p in planes:
        p.tex = _make_tex(rng, f_max, p.amp)
    return SceneSpec(cfg, W, H, V, N, planes, cams, dist)


@torch.no_grad()
def render_view(spec: SceneSpec, view: int, device=None):
    """Returns (gray uint8 HxW, depth float32 HxW, normal_world float32 HxWx3)."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dt = torch.float64
    K, R, t = spec.cams[view]
    Kt = torch.tensor(K, dtype=dt, device=dev)
    Rt = torch.tensor(R, dtype=dt, device=dev)
    C = torch.tensor(-R.T @ t, dtype=dt, device=dev)
    W, H = spec.width, spec.height
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt, device=dev), torch.arange(W, dtype=dt, device=dev), indexing="ij")
    img_acc = torch.zeros(H, W, dtype=dt, device=dev)
    depth_c = None
    normal_c = None
    offs = [(-0.25, -0.25), (0.25, -0.25), (-0.25, 0.25), (0.25, 0.25), (0.0, 0.0)]
    for oi, (ox, oy) in enumerate(offs):
        u = (xs + ox - Kt[0, 2]) / Kt[0, 0]
        v = (ys + oy - Kt[1, 2]) / Kt[1, 1]
        d_cam = torch.stack([u, v, torch.ones_like(u)], -1)
        d = d_cam @ Rt                                   # R^T d_cam, as row vectors
        best_t = torch.full((H, W), float("inf"), dtype=dt, device=dev)
        best_val = torch.zeros(H, W, dtype=dt, device=dev)
        best_n = torch.zeros(H, W, 3, dtype=dt, device=dev)
        for p in spec.planes:
            n = torch.tensor(p.normal, dtype=dt, device=dev)
            o = torch.tensor(p.origin, dtype=dt, device=dev)
            e1 = torch.tensor(p.e1, dtype=dt, device=dev)
            e2 = torch.tensor(p.e2, dtype=dt, device=dev)
            denom = d @ n
            tt = ((o - C) @ n) / denom
            X = C + tt[..., None] * d
            s1 = (X - o) @ e1
            s2 = (X - o) @ e2
            ok = (tt > 1e-6) & (tt < best_t)
            if math.isfinite(p.half[0]):
                ok &= (s1.abs() <= p.half[0]) & (s2.abs() <= p.half[1])
            if oi < 4:
                f1 = torch.tensor(p.tex["f1"], dtype=dt, device=dev)
                f2 = torch.tensor(p.tex["f2"], dtype=dt, device=dev)
                ph = torch.tensor(p.tex["ph"], dtype=dt, device=dev)
                val = torch.full((H, W), p.base, dtype=dt, device=dev) + p.ramp * (s1 + 0.5 * s2)
                for k0 in range(0, f1.numel(), 8):  # chunked to bound memory
                    arg = 2 * math.pi * (s1[..., None] * f1[k0:k0 + 8] + s2[..., None] * f2[k0:k0 + 8]) + ph[k0:k0 + 8]
                    val = val + p.tex["amp"] * torch.sin(arg).sum(-1)
                best_val = torch.where(ok, val, best_val)
            else:
                nn = torch.where((denom > 0)[..., None], -n.expand(H, W, 3), n.expand(H, W, 3))
                best_n = torch.where(ok[..., None], nn, best_n)
            best_t = torch.where(ok, tt, best_t)
        if oi < 4:
            img_acc += best_val
        else:
            depth_c = best_t
            normal_c = best_n
    img = (img_acc / 4.0).clamp(0, 255).round().to(torch.uint8)
    depth_c = torch.where(torch.isfinite(depth_c), depth_c, torch.zeros_like(depth_c))
    return img.cpu().numpy(), depth_c.float().cpu().numpy(), normal_c.float().cpu().numpy()


def select_pairs(spec: SceneSpec):
    """N nearest cameras by centre distance, score 100 (pair.txt format, main.cpp:264-308)."""
    Cs = np.stack([-R.T @ t for (_, R, t) in spec.cams])
    pairs = []
    for i in range(spec.n_views):
        d = np.linalg.norm(Cs - Cs[i], axis=1)
        d[i] = np.inf
        pairs.append([int(j) for j in np.argsort(d, kind="stable")[: spec.n_src]])
    return pairs


def write_cam(path, K, R, t, dmin, dmax):
    interval = (dmax - dmin) / 191.0
    with open(path, "w") as f:
        f.write("extrinsic\n")
        for r in range(3):
            f.write(" ".join(f"{R[r, c]:.10f}" for c in range(3)) + f" {t[r]:.10f}\n")
        f.write("0.0 0.0 0.0 1.0\n\nintrinsic\n")
        for r in range(3):
            f.write(" ".join(f"{K[r, c]:.10f}" for c in range(3)) + "\n")
        f.write(f"\n{dmin:.8f} {interval:.8f} 192 {dmax:.8f}\n")


def write_scene(spec: SceneSpec, out_dir, device=None, save_gt=True, jpeg_quality=98, sidecar=True, gt_every=1):
    """Writes the colmap2mvsnet layout; returns the source lists.  save_gt: True (depth + normal), "depth", or False;
    gt_every = K keeps the ground truth of every K-th view only (large scenes)."""
    import cv2
    out = Path(out_dir)
    (out / "images").mkdir(parents=True, exist_ok=True)
    (out / "cams").mkdir(exist_ok=True)
    if save_gt:
        (out / "gt").mkdir(exist_ok=True)
    for v in range(spec.n_views):
        img, depth, normal = render_view(spec, v, device)
        cv2.imwrite(str(out / "images" / f"{v:08d}.jpg"), img, [cv2.IMWRITE_JPEG_QUALITY, jpeg_quality])
        if sidecar:
            # decoded pixels as cv2 (libjpeg-turbo) sees them: fed to the reference build's
            # imread stand-in so both implementations can be given identical pixels
            dec = cv2.imread(str(out / "images" / f"{v:08d}.jpg"), cv2.IMREAD_GRAYSCALE)
            with open(out / "images" / f"{v:08d}.gray", "wb") as fg:   # int32 rows, int32 cols, pixels
                fg.write(np.array(dec.shape, np.int32).tobytes())
                fg.write(dec.tobytes())
        valid = depth[depth > 0]
        dmin = float(np.percentile(valid, 1)) * 0.75   # colmap2mvsnet.py:407-408
        dmax = float(np.percentile(valid, 99)) * 1.25
        K, R, t = spec.cams[v]
        write_cam(out / "cams" / f"{v:08d}_cam.txt", K, R, t, dmin, dmax)
        if save_gt and v % gt_every == 0:
            np.save(out / "gt" / f"{v:08d}_depth.npy", depth)
            if save_gt != "depth":
                np.save(out / "gt" / f"{v:08d}_normal.npy", normal)
    pairs = select_pairs(spec)
    with open(out / "pair.txt", "w") as f:
        f.write(f"{spec.n_views}\n")
        for v in range(spec.n_views):
            f.write(f"{v}\n{len(pairs[v])} " + " ".join(f"{j} 100.0" for j in pairs[v]) + "\n")
    return pairs


def read_cam(path):
    """Parser mirroring ReadCamera (DPE.cpp:341-382), 'TAT & ETH' 4-number depth line."""
    tok = Path(path).read_text().split()
    assert tok[0] == "extrinsic"
    E = np.array(tok[1:17], dtype=np.float64).reshape(4, 4)
    assert tok[17] == "intrinsic"
    K = np.array(tok[18:27], dtype=np.float64).reshape(3, 3)
    dmin, _interval, _num, dmax = (float(x) for x in tok[27:31])
    return K.astype(np.float32), E[:3, :3].astype(np.float32), E[:3, 3].astype(np.float32), dmin, dmax


def read_pairs(path):
    lines = Path(path).read_text().split("\n")
    n = int(lines[0].split()[0])
    out = []
    for i in range(n):
        ref = int(lines[1 + 2 * i].split()[0])
        tok = lines[2 + 2 * i].split()
        k = int(tok[0])
        src = [int(tok[1 + 2 * j]) for j in range(k) if float(tok[2 + 2 * j]) > 0.0]
        out.append((ref, src))
    return out
