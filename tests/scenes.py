"""Shared small synthetic scenes for the tests (rendered on the CPU, cached per process)."""
import functools

import numpy as np

import synth


@functools.lru_cache(maxsize=8)
def small_scene(config="c1", scale=0.25, n_views=None):
    spec = synth.make_scene(config, scale=scale, n_views=n_views)
    rv = [synth.render_view(spec, v, device="cpu") for v in range(spec.n_views)]
    grays = [r[0] for r in rv]
    cams = [tuple(np.asarray(a, np.float32) for a in c) for c in spec.cams]
    drs = []
    for r in rv:
        valid = r[1][r[1] > 0]
        drs.append((float(np.percentile(valid, 1)) * 0.75, float(np.percentile(valid, 99)) * 1.25))
    pairs = synth.select_pairs(spec)
    gt = [(r[1], r[2]) for r in rv]
    return spec, grays, cams, drs, pairs, gt


def seeded_hypotheses(spec, cams, gt, n, seed=0, margin=8, depth_sigma=0.01, normal_sigma=0.05):
    """Pixels + plane hypotheses around the ground truth of view 0 (camera coordinates)."""
    rng = np.random.default_rng(seed)
    W, H = spec.width, spec.height
    xy = np.stack([rng.integers(margin, W - margin, n), rng.integers(margin, H - margin, n)], 1).astype(np.int32)
    K, R, t = cams[0]
    planes = np.zeros((n, 4), np.float32)
    for i, (x, y) in enumerate(xy):
        d = float(gt[0][0][y, x]) * (1 + rng.normal(0, depth_sigma))
        nrm = R @ gt[0][1][y, x] + rng.normal(0, normal_sigma, 3)
        nrm /= np.linalg.norm(nrm)
        X = d * np.array([(x - K[0, 2]) / K[0, 0], (y - K[1, 2]) / K[1, 1], 1.0])
        planes[i] = [nrm[0], nrm[1], nrm[2], -float(nrm @ X)]
    return xy, planes
