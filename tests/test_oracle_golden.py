"""Pins the float64 oracle (oracle/ncc_oracle.py) to the reference's own device code: the
fixture tests/golden/ref_probe_c1.npz holds outputs of ComputeBilateralNCCOld and
ComputeGeomConsistencyCost compiled from /root/reference (oracle/ref_probe.cu) and run on a
B200 by oracle/make_golden.py.  CPU only.

Tolerances: the reference computes in fp32 with --use_fast_math and fetches source pixels
through the texture unit (bilinear weights in 1.8 fixed point, coordinates in fp32), the oracle
in float64 with the same weight quantisation; a coordinate that differs by 1e-4 px can move a
tap to the neighbouring 1/256 weight bin.  Measured: median |diff| ~5e-5.  Gate: median < 2e-4,
99th percentile < 3e-3, and every cost on the same side of the 2.0 "invalid" value unless the
projected centre is within 1e-3 px of the image border.
"""
from pathlib import Path

import numpy as np
import pytest

import ncc_oracle as O

FIX = Path(__file__).resolve().parent / "golden" / "ref_probe_c1.npz"


@pytest.fixture(scope="module")
def fx():
    if not FIX.exists():
        pytest.fail("tests/golden/ref_probe_c1.npz is missing: run oracle/make_golden.py on a GPU box and commit it")
    return np.load(FIX)


def test_ncc_old_matches_reference_device_code(fx):
    imgs = fx["images"].astype(np.float32)
    cams = [(fx["K"][i], fx["R"][i], fx["t"][i]) for i in range(4)]
    xy, planes, ref = fx["xy"], fx["planes"], fx["ref_ncc"]
    got = np.zeros_like(ref)
    for i, (x, y) in enumerate(xy):
        for s in range(3):
            got[i, s] = O.bilateral_ncc_old(imgs[0], imgs[s + 1], cams[0], cams[s + 1], int(x), int(y), planes[i].astype(np.float64), quant=1)
    invalid_ref, invalid_got = ref >= 2.0, got >= 2.0
    assert (invalid_ref == invalid_got).mean() > 0.995
    both = ~invalid_ref & ~invalid_got
    d = np.abs(got - ref)[both]
    assert both.sum() > 300
    assert np.median(d) < 2e-4, np.median(d)
    assert np.percentile(d, 99) < 3e-3, np.percentile(d, 99)
    assert d.max() < 2e-2, d.max()


def test_geom_cost_matches_reference_device_code(fx):
    cams = [(fx["K"][i], fx["R"][i], fx["t"][i]) for i in range(4)]
    xy, planes, ref, depths = fx["xy"], fx["planes"], fx["ref_geom"], fx["depths"]
    got = np.zeros_like(ref)
    for i, (x, y) in enumerate(xy):
        for s in range(3):
            got[i, s] = O.geom_consistency_cost(cams[0], cams[s + 1], depths[s + 1], int(x), int(y), planes[i].astype(np.float64))
    assert ((ref == 3.0) == (got == 3.0)).mean() > 0.99
    ok = (ref < 3.0) & (got < 3.0)
    d = np.abs(got - ref)[ok]
    assert ok.sum() > 100
    # fp32 world-coordinate round trip in the reference: ~1e-3 px
    assert np.median(d) < 2e-3 and np.percentile(d, 99) < 2e-2, (np.median(d), np.percentile(d, 99))
