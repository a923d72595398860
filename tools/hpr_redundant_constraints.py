"""CPU experiment behind DESIGN 4.3 / 7.1: how much of the hidden-point-removal LP work is spent on constraints that can
never bind?  If a plane through point i lies above every upper-hull VERTEX it lies above every point, so a point that is
proven hidden can be dropped from every other point's constraint set.  This script runs the kernel's own LP
(cloudaae_b200/csrc/hpr_lp.cuh through tests/hpr_host_harness.cpp — test infrastructure, CPU) on fixture clouds

  (1) as the kernel runs it (every point is a constraint), and
  (2) on the visible points alone (the best case: every hidden point already removed),

checks that (2) classifies every one of them visible again (the redundancy claim, on real data), and prints the phase-1
constraint evaluations per point for hidden / visible points and the size of the verification candidate set.
   python tools/hpr_redundant_constraints.py [clouds per problem, default 24]"""
import ctypes, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from oracle import synthesis as S

out = os.path.join(tempfile.mkdtemp(), "libhpr_host.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "cloudaae_b200", "csrc"),
                       os.path.join(ROOT, "tests", "hpr_host_harness.cpp"), "-o", out])
lib = ctypes.CDLL(out)
lib.hpr_host_iters_ids.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
lib.hpr_host_iters_ids.restype = ctypes.c_int


def run(flipped):
    flipped = np.ascontiguousarray(flipped, np.float32)
    b, n, _ = flipped.shape
    flags = np.zeros((b, n), np.uint8); stats = np.zeros((b, 8), np.int64)
    iters = np.zeros((b, n), np.int32); ids = np.zeros((b, n), np.int32)
    assert lib.hpr_host_iters_ids(flipped.ctypes.data, b, n, flags.ctypes.data, stats.ctypes.data, iters.ctypes.data, ids.ctypes.data) == 0
    per_point = np.full((b, n), -1, np.int64)          # phase-1 evaluations by ORIGINAL index (-1: duplicate, took no part)
    for k in range(b):
        m = ids[k] >= 0
        per_point[k, ids[k][m]] = iters[k][m]
    return flags.astype(bool), stats, per_point


def clouds(b, seed, occluded):
    t, a, c = cases.ycb_poses()
    sel = np.random.default_rng(seed).integers(0, len(c), b)
    P = S.transform_object_model(cases.ycb_models()[c[sel]], a[sel], t[sel])
    if occluded:
        rng = np.random.default_rng(seed + 1)
        occ = S.spherical_occluder(t[sel][:, 2], rng.standard_normal((b, 2, 3)), rng.standard_normal((b, 2, 200, 3)))
        P = np.concatenate([P, occ], 1)
    return S.spherical_flip(P)[0][:, :-1]


b = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for occluded in (False, True):
    F = clouds(b, 11, occluded)
    flags, stats, pp = run(F)
    took = pp >= 0
    hid, vis = took & ~flags, took & flags
    ev_hid, ev_vis = pp[hid].mean(), pp[vis].mean()
    ev2, cand1, cand2, still = [], [], [], 0
    for k in range(b):
        Fv = F[k][flags[k]][None]                       # the visible points alone
        f2, s2, p2 = run(Fv)
        still += int(f2.sum()); ev2.append(p2[p2 >= 0]); cand1.append(stats[k, 0]); cand2.append(s2[0, 0])
    ev2 = np.concatenate(ev2).mean()
    nvis = int(flags.sum())
    total1 = pp[took].sum()
    total2 = pp[hid].sum() + ev2 * nvis                 # hidden points still have to be proven hidden once
    print(f"{'occluded cloud (2449 pts)' if occluded else 'bare object (2048 pts)'}: {took.sum() / b:.0f} unique points per cloud, "
          f"{100 * hid.sum() / took.sum():.0f} % hidden\n"
          f"  phase-1 constraint evaluations per point: hidden {ev_hid:.0f}, visible {ev_vis:.0f}; visible with the hidden points removed {ev2:.0f}\n"
          f"  every visible point is still classified visible without them: {still} of {nvis}\n"
          f"  phase-1 total {total1 / b:.0f} -> {total2 / b:.0f} evaluations per cloud ({100 * (1 - total2 / total1):.0f} % fewer); "
          f"verification candidates per survivor {np.mean(cand1):.0f} -> {np.mean(cand2):.0f}")
