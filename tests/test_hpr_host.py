"""The hidden-point-removal LP (cloudaae_b200/csrc/hpr_lp.cuh — the source the sm_100a kernel compiles)
run on the CPU through tests/hpr_host_harness.cpp and compared with scipy's Qhull, the very call the
reference makes (utils/hidden_point_removal.py:32).  Host-logic test: no GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases
from oracle import synthesis as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hpr") / "libhpr_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "cloudaae_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hpr_host_harness.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    lib.hpr_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.hpr_host.restype = ctypes.c_int

    def run(flipped):
        flipped = np.ascontiguousarray(flipped, np.float32)
        b, n, _ = flipped.shape
        flags = np.zeros((b, n), np.uint8)
        stats = np.zeros((b, 8), np.int64)
        assert lib.hpr_host(flipped.ctypes.data, b, n, flags.ctypes.data, stats.ctypes.data) == 0
        return flags.astype(bool), stats
    return run


def _clouds(b, seed, occluded):
    t, a, c = cases.ycb_poses()
    sel = np.random.default_rng(seed).integers(0, len(c), b)
    P = S.transform_object_model(cases.ycb_models()[c[sel]], a[sel], t[sel])
    if occluded:
        rng = np.random.default_rng(seed + 1)
        occ = S.spherical_occluder(t[sel][:, 2], rng.standard_normal((b, 2, 3)), rng.standard_normal((b, 2, 200, 3)))
        P = np.concatenate([P, occ], 1)
    flipped, _ = S.spherical_flip(P)
    return flipped[:, :-1]  # the viewpoint row is implicit in the LP formulation


@pytest.mark.parametrize("occluded", [True, False])
def test_lp_visibility_equals_qhull(harness, occluded):
    from scipy.spatial import ConvexHull
    b = 6
    flipped = _clouds(b, 5, occluded)
    flags, stats = harness(flipped)
    inter = union = 0
    for k in range(b):
        pts = np.concatenate([flipped[k], np.zeros((1, 3), np.float32)]).astype(np.float64)
        hv = np.zeros(flipped.shape[1] + 1, bool)
        hv[ConvexHull(pts).vertices] = True
        assert hv[-1]                                    # the viewpoint is always a hull vertex
        inter += (hv[:-1] & flags[k]).sum(); union += (hv[:-1] | flags[k]).sum()
    iou = inter / union
    per_point = stats[:, 1].sum() / stats[:, 0].sum()
    print(f"host LP vs Qhull IoU {iou:.6f} ({'occluded' if occluded else 'org'}): {per_point:.0f} constraint "
          f"evaluations/point in phase 1, survivors {stats[:, 2].mean():.0f}, re-solved {stats[:, 3].mean():.1f} "
          f"({stats[:, 4].sum() / max(stats[:, 3].sum(), 1):.0f} evaluations each, max rounds {stats[:, 5].max()}, "
          f"full-LP fallbacks {stats[:, 7].sum()}), fp64 checks in verification {stats[:, 6].mean():.0f}")
    assert iou >= 0.999


def test_lp_duplicates_keep_one_representative(harness):
    # class 17 stores 574 copies of point 0: at most one copy may be visible
    t, a, c = cases.ycb_poses()
    k = int(np.where(c == 17)[0][0])
    P = S.transform_object_model(cases.ycb_models()[[17]], a[[k]], t[[k]])
    flipped, _ = S.spherical_flip(P)
    flags, stats = harness(flipped[:, :-1])
    dup = np.where((P[0] == P[0, 0]).all(axis=1))[0]
    assert len(dup) > 500 and flags[0, dup].sum() <= 1
    assert stats[0, 0] == 2048 - (len(dup) - 1)
