"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the public
operator API -> ctypes -> the C ABI of libcloudaae_b200.so.  Checkers: the CPU oracle, the golden
vectors, and the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref)."""
import numpy as np
import pytest
import torch

import cases
from oracle import ops as O

pytestmark = pytest.mark.gpu

from cloudaae_b200 import (farthest_point_sample, farthest_point_sample_gather, gather_point,  # noqa: E402
                           gather_point_grad, nn_distance, nn_distance_grad, prob_sample)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def np_(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ nn_distance forward
NND_SHAPES = [(32, 1024, 1024), (32, 256, 256), (1, 256, 256), (3, 1, 7), (2, 513, 2049), (5, 2449, 1024),
              (1, 2049, 2449), (2, 4100, 37), (7, 333, 1)]


@pytest.mark.parametrize("b,n,m", NND_SHAPES)
def test_nn_distance_bit_exact_vs_oracle(b, n, m):
    x1, x2 = cases.random_clouds(b * 7 + n, b, n), cases.random_clouds(b * 11 + m, b, m)
    d1, i1, d2, i2 = nn_distance(cu(x1), cu(x2))
    o = O.nn_distance(x1, x2, "gpu", threads=O.max_threads())
    assert i1.dtype == torch.int32 and d1.dtype == torch.float32 and d1.shape == (b, n) and i2.shape == (b, m)
    assert (np_(i1) == o[1]).all() and (np_(i2) == o[3]).all()          # indices bit-exact
    assert (np_(d1) == o[0]).all() and (np_(d2) == o[2]).all()          # distances bit-exact too


def test_nn_distance_golden_reference_smoke_arrays(golden_ops):
    x1, x2 = cases.nnd_smoke_inputs()
    d1, i1, d2, i2 = nn_distance(cu(x1), cu(x2))
    g = golden_ops
    assert (np_(i1) == g["nnd_smoke_gpu_idx1"]).all() and (np_(i2) == g["nnd_smoke_gpu_idx2"]).all()
    assert (np_(d1) == g["nnd_smoke_gpu_dist1"]).all() and (np_(d2) == g["nnd_smoke_gpu_dist2"]).all()
    # against what the reference's CPU kernel produced: within the north star's 1e-5 relative
    np.testing.assert_allclose(np_(d1), g["nnd_smoke_cpu_dist1"], rtol=1e-5)
    np.testing.assert_allclose(np_(d2), g["nnd_smoke_cpu_dist2"], rtol=1e-5)


@pytest.mark.parametrize("b,n,m", [(32, 1024, 1024), (4, 2449, 2049), (2, 16384, 1024), (3, 5, 600)])
def test_nn_distance_equals_reference_cuda_kernel(b, n, m):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    x1, x2 = cu(cases.random_clouds(1, b, n)), cu(cases.random_clouds(2, b, m))
    mine = nn_distance(x1, x2)
    ref = O.ref_gpu_nn_distance(x1, x2)
    for a, r in zip(mine, ref):
        assert torch.equal(a, r)


def test_nn_distance_exact_ties_pick_lowest_index():
    base = cases.random_clouds(6, 2, 600)
    x2 = np.concatenate([base, base, base], axis=1)  # duplicates across the 512 / 2048 tile edges
    x1 = cases.random_clouds(5, 2, 300)
    _, i1, _, i2 = nn_distance(cu(x1), cu(x2))
    o = O.nn_distance(x1, x2, "gpu")
    assert (np_(i1) < 600).all() and (np_(i1) == o[1]).all() and (np_(i2) == o[3]).all()
    # identical clouds: every point's nearest neighbour is the first copy of itself
    d1, i1, d2, i2 = nn_distance(cu(x2), cu(x2))
    assert (np_(d1) == 0).all() and (np_(i1) == np.tile(np.arange(600), 3)[None]).all()


def test_nn_distance_empty_and_ragged():
    z = torch.zeros(2, 0, 3, device="cuda")
    x = cu(cases.random_clouds(0, 2, 17))
    d1, i1, d2, i2 = nn_distance(x, z)
    assert d1.shape == (2, 17) and (d1 == 0).all() and (i1 == 0).all() and d2.shape == (2, 0)
    d1, i1, d2, i2 = nn_distance(z, x)
    assert d1.shape == (2, 0) and (d2 == 0).all() and (i2 == 0).all()
    e = torch.zeros(0, 5, 3, device="cuda")
    assert nn_distance(e, e)[0].shape == (0, 5)
    # non-contiguous input views are accepted
    big = cu(cases.random_clouds(3, 2, 64))
    v = big[:, ::2, :]
    a = nn_distance(v, big)
    b_ = nn_distance(v.contiguous(), big)
    assert torch.equal(a[1], b_[1])


def test_nn_distance_linearity_properties_at_full_size():
    """Size-independent properties at the train-step size (b=128, n=m=1024)."""
    x1, x2 = cu(cases.random_clouds(21, 128, 1024)), cu(cases.random_clouds(22, 128, 1024))
    d1, i1, d2, i2 = nn_distance(x1, x2)
    # swapping the arguments swaps the outputs exactly (d is bit-symmetric)
    e1, j1, e2, j2 = nn_distance(x2, x1)
    assert torch.equal(d1, e2) and torch.equal(i1, j2) and torch.equal(d2, e1) and torch.equal(i2, j1)
    # the reported distance is the distance to the reported index
    nb = torch.gather(x2, 1, i1.long()[..., None].expand(-1, -1, 3))
    assert torch.allclose(((x1 - nb) ** 2).sum(-1), d1, rtol=1e-5, atol=1e-9)
    # the minimum is no larger than the distance to any sampled candidate
    assert (d1 <= ((x1[:, :, None, :] - x2[:, None, :64, :]) ** 2).sum(-1).min(-1).values * (1 + 1e-5) + 1e-12).all()
    # permuting the candidates permutes the indices (no exact ties in random data)
    perm = torch.randperm(1024, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    _, k1, _, _ = nn_distance(x1, x2[:, perm])
    assert torch.equal(perm[k1.long()].int(), i1)


# ------------------------------------------------------------------ nn_distance backward
@pytest.mark.parametrize("b,n,m", [(32, 1024, 1024), (2, 300, 2449), (1, 256, 256), (2, 16384, 1024), (2, 20000, 50)])
def test_nn_distance_grad_vs_oracle(b, n, m):
    x1, x2 = cases.random_clouds(31, b, n), cases.random_clouds(32, b, m)
    rng = np.random.default_rng(7)
    gd1 = rng.standard_normal((b, n)).astype(np.float32)
    gd2 = rng.standard_normal((b, m)).astype(np.float32)
    _, i1, _, i2 = O.nn_distance(x1, x2, "gpu", threads=O.max_threads())
    g1, g2 = nn_distance_grad(cu(x1), cu(x2), cu(gd1), cu(i1), cu(gd2), cu(i2))
    o1, o2 = O.nn_distance_grad(x1, x2, gd1, i1, gd2, i2)
    # products are bit-identical; only the fp32 summation order differs (atomics) -> 1e-5 relative
    scale1 = np.abs(o1).max() + 1e-12
    scale2 = np.abs(o2).max() + 1e-12
    assert np.abs(np_(g1) - o1).max() <= 1e-5 * scale1
    assert np.abs(np_(g2) - o2).max() <= 1e-5 * scale2
    # points that receive no cross term have a single addend: bit-exact
    lone1 = np.ones((b, n), bool)
    for i in range(b):
        lone1[i, np.unique(i2[i])] = False
    assert (np_(g1)[lone1] == o1[lone1]).all()


def test_nn_distance_autograd_matches_reference_gradient_definition(golden_ops):
    x1n, x2n = cases.nnd_smoke_inputs()
    x1 = cu(x1n).requires_grad_(True)
    x2 = cu(x2n).requires_grad_(True)
    gd1, gd2 = cases.nnd_smoke_grads()
    d1, i1, d2, i2 = nn_distance(x1, x2)
    assert not i1.requires_grad and not i2.requires_grad
    ((d1 * cu(gd1)).sum() + (d2 * cu(gd2)).sum()).backward()
    g = golden_ops  # produced by the reference's CPU NnDistanceGrad kernel
    if (np_(i1) == g["nnd_smoke_cpu_idx1"]).all() and (np_(i2) == g["nnd_smoke_cpu_idx2"]).all():
        s1 = np.abs(g["nnd_smoke_cpu_gxyz1"]).max()
        s2 = np.abs(g["nnd_smoke_cpu_gxyz2"]).max()
        assert np.abs(np_(x1.grad) - g["nnd_smoke_cpu_gxyz1"]).max() <= 1e-5 * s1
        assert np.abs(np_(x2.grad) - g["nnd_smoke_cpu_gxyz2"]).max() <= 1e-5 * s2
    # loss = sum(dist1): gradient only through dist1
    x1.grad = None; x2.grad = None
    nn_distance(x1, x2)[0].sum().backward()
    nb = torch.gather(x2.detach(), 1, i1.long()[..., None].expand(-1, -1, 3))
    assert torch.allclose(x1.grad, 2 * (x1.detach() - nb), rtol=1e-6, atol=1e-7)


def test_nn_distance_grad_equals_reference_cuda_kernel_where_deterministic():
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    b, n, m = 8, 1024, 1024
    x1, x2 = cu(cases.random_clouds(41, b, n)), cu(cases.random_clouds(42, b, m))
    d1, i1, d2, i2 = nn_distance(x1, x2)
    gd1, gd2 = torch.randn_like(d1), torch.randn_like(d2)
    mine = nn_distance_grad(x1, x2, gd1, i1, gd2, i2)
    ref = O.ref_gpu_nn_distance_grad(x1, x2, gd1, i1, gd2, i2)
    for a, r in zip(mine, ref):
        assert (a - r).abs().max() <= 1e-5 * r.abs().max()


# ------------------------------------------------------------------ farthest point sampling
@pytest.mark.parametrize("n,m", [(2048, 256), (1024, 256), (2449, 256), (2049, 1024), (256, 256), (100, 37),
                                 (513, 64), (3000, 128), (5000, 64), (8192, 32), (9000, 48)])
def test_fps_bit_exact_vs_oracle(n, m):
    b = 6
    x = cases.random_clouds(n + m, b, n)
    got = farthest_point_sample(m, cu(x))
    assert got.dtype == torch.int32 and got.shape == (b, m)
    assert (np_(got) == O.fps(x, m, threads=O.max_threads())).all()


def test_fps_golden_ycb_and_known_answer(golden_ops):
    got = np_(farthest_point_sample(256, cu(cases.fps_ycb_inputs())))
    assert (got == golden_ops["fps_ycb"]).all()
    # the reference fixture is stored in FPS order: FPS from seed 0 is the identity prefix
    ident = np_(farthest_point_sample(512, cu(cases.ycb_models())))
    assert (ident == np.arange(512)[None]).all()


def test_fps_tie_rule(golden_ops):
    x = cases.fps_ties_inputs()
    got = np_(farthest_point_sample(64, cu(x)))
    assert (got == golden_ops["fps_ties"]).all()
    assert (got >= 1024).any()  # not "lowest index": (k mod 512) decides first
    # padded-visible style duplicates (convexHull pads with repeats of visible points), n = 2449
    rng = np.random.default_rng(3)
    base = cases.random_clouds(9, 4, 900)
    pad = np.stack([base[i, rng.integers(0, 900, 2449 - 900)] for i in range(4)])
    xp = np.concatenate([base, pad], axis=1)
    assert (np_(farthest_point_sample(256, cu(xp))) == O.fps(xp, 256, threads=4)).all()


def test_fps_equals_reference_cuda_kernel():
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    for x, m in ((cases.fps_ycb_inputs(), 256), (cases.fps_ties_inputs(), 64), (cases.random_clouds(5, 40, 1024), 256),
                 (cases.random_clouds(6, 3, 3500), 100)):
        xc = cu(x)
        assert torch.equal(farthest_point_sample(m, xc), O.ref_gpu_fps(xc, m))


def test_fps_edge_cases():
    x = cu(cases.random_clouds(3, 2, 10))
    idx = np_(farthest_point_sample(16, x))
    assert sorted(idx[0, :10].tolist()) == list(range(10)) and (idx[:, 10:] == 0).all()
    assert farthest_point_sample(4, torch.zeros(0, 8, 3, device="cuda")).shape == (0, 4)
    one = np_(farthest_point_sample(1, x))
    assert (one == 0).all()
    # all points identical -> always index 0
    same = torch.ones(2, 700, 3, device="cuda")
    assert (farthest_point_sample(9, same) == 0).all()
    # properties at the microbench size: unique indices, first = 0, min pairwise distance non-increasing
    xb = cu(cases.random_clouds(8, 32, 2048))
    ib = farthest_point_sample(256, xb)
    assert (ib[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == 256 for r in np_(ib))


def test_fps_gather_fused_equals_composition():
    x = cu(cases.random_clouds(12, 5, 1024)).requires_grad_(True)
    idx, xyz = farthest_point_sample_gather(256, x)
    assert torch.equal(xyz, gather_point(x, farthest_point_sample(256, x)))
    xyz.sum().backward()
    assert x.grad.sum().item() == pytest.approx(5 * 256 * 3)
    idx2, xyz2 = farthest_point_sample_gather(256, x.detach())
    assert torch.equal(idx, idx2) and torch.equal(xyz2, xyz.detach())


# ------------------------------------------------------------------ gather / gather grad
def test_gather_point_and_grad():
    b, n, m = 7, 1024, 256
    xn = cases.random_clouds(13, b, n)
    idx = np.random.default_rng(1).integers(0, n, (b, m)).astype(np.int32)
    x = cu(xn).requires_grad_(True)
    out = gather_point(x, cu(idx))
    assert (np_(out) == O.gather(xn, idx)).all()
    og = np.random.default_rng(2).standard_normal((b, m, 3)).astype(np.float32)
    out.backward(cu(og))
    want = O.gather_grad(xn.shape, idx, og)
    np.testing.assert_allclose(np_(x.grad), want, rtol=1e-6, atol=1e-6)
    # unique indices (the FPS case): single writer per slot -> bit exact
    fidx = farthest_point_sample(m, x.detach())
    g = gather_point_grad(x.detach(), fidx, cu(og))
    assert (np_(g) == O.gather_grad(xn.shape, np_(fidx), og)).all()
    if O.have_ref():
        assert torch.equal(out.detach(), O.ref_gpu_gather(x.detach(), cu(idx)))
        assert torch.equal(g, O.ref_gpu_gather_grad(x.detach(), fidx, cu(og)))
    assert gather_point(x.detach(), torch.zeros(b, 0, dtype=torch.int32, device="cuda")).shape == (b, 0, 3)


# ------------------------------------------------------------------ prob_sample
@pytest.mark.parametrize("n,m", [(5, 64), (1000, 300), (8192, 100), (20000, 257)])
def test_prob_sample_bit_exact(n, m):
    rng = np.random.default_rng(n)
    p = rng.uniform(0, 1, (3, n)).astype(np.float32)
    r = rng.uniform(0, 1, (3, m)).astype(np.float32)
    got = prob_sample(cu(p), cu(r))
    assert (np_(got) == O.prob_sample(p, r)).all()
    if O.have_ref():
        assert torch.equal(got, O.ref_gpu_prob_sample(cu(p), cu(r)))


def test_ops_are_cuda_graph_capturable():
    x1, x2 = cu(cases.random_clouds(51, 4, 512)), cu(cases.random_clouds(52, 4, 512))
    eager = nn_distance(x1, x2)
    fe = farthest_point_sample(64, x1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        nn_distance(x1, x2); farthest_point_sample(64, x1)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = nn_distance(x1, x2)
        f = farthest_point_sample(64, x1)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[1], eager[1]) and torch.equal(out[0], eager[0]) and torch.equal(f, fe)
