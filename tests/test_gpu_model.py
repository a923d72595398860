"""GPU parity of the network, losses, gradients and optimiser against the plain-PyTorch float64
restatement of the reference graph (oracle/model_ref.py).  Tolerance: the north star's 1e-3 relative
for features, poses, losses (fp32 here, so the observed errors are ~1e-5)."""
import numpy as np
import pytest
import torch

import cases
from oracle import model_ref as MR

pytestmark = pytest.mark.gpu

from cloudaae_b200 import _capi  # noqa: E402
from cloudaae_b200.models import pointnet_ycb_23_decoder_4 as M  # noqa: E402
from cloudaae_b200.train import CloudAAETrainer  # noqa: E402

RTOL = 1e-3


def rel_err(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def l2_err(a, b):
    """Norm-wise relative error.  Used for gradients: one ReLU / max-pool decision on an activation
    that is zero to within fp32 rounding moves a single addend (d_emb/N) between the two sides, which
    is a 1e-3-sized max-norm blip at the tiny batch sizes used here but negligible norm-wise."""
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _gemm(ta, tb, Mm, N, K, A, lda, B, ldb, C, ldc, bias=None, acc=0):
    lib = _capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.caae_gemm_f32(ta, tb, Mm, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, C.data_ptr(), ldc,
                                  None if bias is None else bias.data_ptr(), acc, st), "gemm")


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("Mm,N,K", [(128, 64, 16), (300, 130, 77), (1, 3, 1024), (64, 256, 4096), (33, 17, 5000)])
def test_gemm_f32_all_layouts(ta, tb, Mm, N, K):
    g = torch.Generator("cuda").manual_seed(Mm * 7 + N)
    A = torch.randn((K, Mm) if ta else (Mm, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.full((Mm, N + 3), 7.0, device="cuda")
    _gemm(ta, tb, Mm, N, K, A, A.shape[1], B, B.shape[1], C, N + 3, bias)
    want = (A.double().T if ta else A.double()) @ (B.double().T if tb else B.double()) + bias.double()
    assert rel_err(C[:, :N], want) < 1e-5
    assert (C[:, N:] == 7.0).all()  # leading-dimension padding untouched
    C2 = C.clone()
    _gemm(ta, tb, Mm, N, K, A, A.shape[1], B, B.shape[1], C2, N + 3, None, 1)
    assert rel_err(C2[:, :N], want + want - bias.double()) < 1e-5


def _knn(x, c, k):
    b, n, ld = x.shape
    idx = torch.empty(b, n, k, dtype=torch.int32, device="cuda")
    _capi.check(_capi.lib().caae_knn(b, n, c, k, x.data_ptr(), ld, idx.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), "knn")
    return idx


@pytest.mark.parametrize("b,n,c,ld,k", [(4, 256, 3, 24, 10), (3, 256, 64, 320, 10), (2, 100, 64, 64, 10),
                                        (2, 600, 16, 16, 7), (1, 64, 128, 128, 20), (2, 10, 3, 3, 10)])
def test_knn_matches_reference_topk(b, n, c, ld, k):
    g = torch.Generator("cuda").manual_seed(n + c)
    x = torch.randn(b, n, ld, device="cuda", generator=g)
    idx = _knn(x, c, k).cpu().long()
    xd = x[:, :, :c].double().cpu()
    adj = MR.pairwise_xyz_distance(xd.unsqueeze(2) if c != 3 else xd)
    want = MR.knn(adj, k)
    same = (idx == want)
    assert same.float().mean() > 0.995
    # any disagreement must be a floating-point near-tie of the selected distances
    dsel = torch.gather(adj, 2, idx)
    dwant = torch.gather(adj, 2, want)
    scale = adj.abs().max()
    assert ((dsel - dwant).abs() <= 1e-5 * scale)[~same].all()
    assert (idx[:, :, 0] == torch.arange(n)[None]).float().mean() > 0.99  # the point itself comes first
    # ascending order of the (fp32) distances the kernel saw
    assert (dsel[:, :, 1:] - dsel[:, :, :-1] >= -1e-5 * scale).all()


def test_knn_exact_ties_pick_lower_index():
    x = torch.zeros(1, 64, 3, device="cuda")
    x[0, :, 0] = torch.arange(64, device="cuda").float() // 2  # pairs of identical points on a line
    idx = _knn(x, 3, 4).cpu()
    want = MR.knn(MR.pairwise_xyz_distance(x.double().cpu()), 4)
    assert (idx == want).all()


def _setup(model, b, n, seed=3):
    layers = MR.DGCNN_LAYERS if model == "dgcnn" else MR.pn_layers(24, n)
    if model == "dgcnn":
        layers = list(layers)
        layers[7] = ("dgcnn_output", 1024, n * 12, False)
    p32 = MR.init_params(layers, seed=seed, perturb=True)
    v = M.Variables(M.dgcnn_layers(n, 24) if model == "dgcnn" else M.pn_layers(n, 24), device="cuda", seed=None)
    v.load_state_dict(p32)
    p64 = {k: t.double().clone() for k, t in p32.items()}
    rng = np.random.default_rng(seed)
    clouds = cases.posed_ycb_clouds(1)[rng.integers(0, 21, b)]
    cls = torch.from_numpy(rng.integers(0, 21, b).astype(np.int32))
    visible = torch.from_numpy(np.ascontiguousarray(clouds[:, :n + 50])).float()
    target = torch.from_numpy(np.ascontiguousarray(clouds[:, :4 * n])).float()
    noise = torch.from_numpy((rng.standard_normal((b, n, 3)) * 0.004 / 3).astype(np.float32))
    t, a, _ = cases.ycb_poses()
    sel = rng.integers(0, len(t), b)
    return v, p64, visible, target, cls, torch.from_numpy(t[sel]), torch.from_numpy(a[sel]), noise


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("model,b,n", [("dgcnn", 8, 256), ("dgcnn", 3, 128), ("pn", 8, 256), ("dgcnn", 1, 256)])
def test_train_forward_losses_and_gradients(model, b, n, precision):
    v, p64, visible, target, cls, trans, axag, noise = _setup(model, b, n)
    tr = CloudAAETrainer(batch_size=b, num_point=n, model=model, variables=v, precision=precision)
    dev = lambda t: t.cuda().contiguous()  # noqa: E731
    bn_decay = 0.9375
    tr.decay.fill_(bn_decay)
    losses = tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise))
    tr.backward(dev(target))
    torch.cuda.synchronize()

    # ---- oracle (float64, literal TF graph); neighbour indices injected to take near-ties out
    x64, mean64 = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
    assert rel_err(tr.x, x64) < 1e-5 and rel_err(tr.mean, mean64) < 1e-6
    params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p64.items()}
    ema = {}
    override = [i.view(b, n, -1).cpu().long() for i in tr.engine.idx] if model == "dgcnn" else None
    amax = tr.engine.argmax.cpu().long() if model == "pn" else None
    total, aux = MR.train_losses(params, x64, mean64, target.double(), trans.double(), axag.double(), bn_decay,
                                 ema_updates=ema, nn_idx_override=override, model=model, argmax_override=amax)
    total.backward()
    if model == "pn":  # the kernel's argmax rows hold the maximum (up to fp32 near-ties)
        ep = aux["end_points"]
        assert rel_err(ep["embedding"], ep["pre_pool_max"]) < 1e-5

    if model == "dgcnn":  # the kernel's own kNN agrees with the reference selection
        _, _, _, ep = MR.get_model_dgcnn_mean_6d(x64, p64, True, True, 10, bn_decay)
        agree = np.mean([(o == w).float().mean().item() for o, w in zip(override, ep["nn_idx"])])
        assert agree > 0.98  # layers 2-4 select in fp32 feature space; near-ties flip (see test_knn_*)

    assert rel_err(tr.engine.emb, aux["end_points"]["embedding"]) < RTOL
    assert rel_err(tr.recon, aux["recon"]) < RTOL
    assert rel_err(tr.engine.fc_y[tr.engine.branches[1][-1]], aux["rot_pred"]) < RTOL
    assert rel_err(tr.trans_pred, aux["trans_pred"]) < RTOL
    got = losses.cpu().double()
    for i, key in enumerate(("chamfer", "trans", "rot"), start=1):
        assert abs(got[i] - aux[key].item()) <= RTOL * abs(aux[key].item()), key
    assert abs(got[0] - total.item()) <= RTOL * abs(total.item())

    # ---- gradients of every trainable variable
    worst = {}
    for name in v.trainable_names():
        g_ref = params[name].grad
        g = v.grad_of(name)
        if name.endswith("/biases") and (name.rsplit("/", 1)[0] + "/bn/gamma") in v:
            # bias in front of a training-mode BN: mathematically zero gradient
            assert g.abs().max().item() == 0.0
            assert g_ref.abs().max().item() < 1e-6 * max(1.0, params[name.replace("biases", "weights")].grad.abs().max().item())
            continue
        worst[name] = l2_err(g, g_ref)
    # fp32 path: 1e-3.  tf32 path (dgcnn_agg / pn_conv5 GEMMs on the tensor cores): forward features,
    # poses and losses stay inside 1e-3 (asserted above), but the FC layers' batch norm over only `b`
    # near-identical embeddings (random-init network) divides by a tiny batch std and amplifies the
    # 2e-4 TF32 perturbation of the embedding ~100x in the gradients behind it; 5e-2 bounds that.
    gtol = RTOL if precision == "fp32" else 5e-2
    if b * n < 1024:
        # 384 points: ONE ReLU decision on a pre-activation that is zero to within fp32 rounding moves 1/3840 of a
        # channel's batch-norm backward statistics, i.e. every gradient element behind it, by ~1e-3 (observed 1.7e-3
        # in 3 of 3072 elements for one seed, 4e-5 for the others — tools/debug_grad_b3.py; kNN agreement 1.0000 and
        # features 5e-6 in the same run).  The benchmarked batch is asserted at 1e-3 in test_gpu_model_b128.py.
        gtol = max(gtol, 2.5e-3)
    bad = {k: e for k, e in worst.items() if e > gtol}
    assert not bad, bad

    # ---- EMA update  shadow = d*shadow + (1-d)*batch
    for name, want in ema.items():
        assert rel_err(v[name], want) < RTOL, name


def test_adam_matches_tf_formula_and_step_state():
    b, n = 4, 256
    v, p64, visible, target, cls, trans, axag, noise = _setup("dgcnn", b, n)
    tr = CloudAAETrainer(batch_size=b, num_point=n, variables=v, precision="fp32")
    dev = lambda t: t.cuda().contiguous()  # noqa: E731
    args = (dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise))
    p0 = v.flat.clone()
    m = torch.zeros_like(p0); vv = torch.zeros_like(p0)
    want = p0.double().cpu()
    m, vv = m.double().cpu(), vv.double().cpu()
    for step in range(1, 4):
        tr.train_step(*args)
        g = v.grad.double().cpu()
        want, m, vv = MR.adam_step(want, g, m, vv, step)
        # bn_decay schedule: 0.5, then min(0.99, 1 - 0.5*0.5^floor(step*B/40))
        assert tr.decay.item() == pytest.approx(MR.bn_decay_schedule(step - 1, b), rel=1e-6)
        assert tr.state[0].item() == step
        assert rel_err(v.flat, want) < 1e-5
        want = v.flat.double().cpu()  # re-sync so errors do not compound through the next gradient
        m, vv = tr.adam_m.double().cpu(), tr.adam_v.double().cpu()


def test_eval_mode_uses_moving_averages_and_api_contract():
    b, n = 2, 256
    v, p64, visible, target, cls, trans, axag, noise = _setup("dgcnn", b, n, seed=5)
    x64, _ = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
    x = x64.float().cuda()
    recon, rot, tvec, ep = M.get_model_dgcnn_mean_6d(x, False, False, 10, variables=v)
    assert recon.shape == (b, 4 * n, 3) and rot.shape == (b, 3) and tvec.shape == (b, 3)
    assert ep["layer_before_embedding"].shape == (b, n, 1, 1024) and ep["embedding"].shape == (b, 1024)
    override = [i.cpu().long() for i in ep["nn_idx"]]
    r64, rot64, t64, ep64 = MR.get_model_dgcnn_mean_6d(x64, p64, False, False, 10, nn_idx_override=override)
    assert rel_err(recon, r64) < RTOL and rel_err(rot, rot64) < RTOL and rel_err(tvec, t64) < RTOL
    assert rel_err(ep["layer_before_embedding"], ep64["layer_before_embedding"]) < RTOL
    assert rel_err(ep["embedding"], ep64["embedding"]) < RTOL
    # mixed flags (encoder in training mode, FC in inference mode), bn_decay as a float
    ema_before = v["dgcnn_fc1/bn/ema_mean"].clone()
    enc_before = v["dgcnn1/bn/ema_mean"].clone()
    M.get_model_dgcnn_mean_6d(x, True, False, 10, bn_decay=0.5, variables=v)
    assert torch.equal(v["dgcnn_fc1/bn/ema_mean"], ema_before)
    assert not torch.equal(v["dgcnn1/bn/ema_mean"], enc_before)
    # PointNet variant
    vp, pp64, *_ = _setup("pn", b, n, seed=6)
    recon, rot, tvec, ep = M.get_model_pn(x, False, variables=vp)
    r64, rot64, t64, ep64 = MR.get_model_pn(x64, pp64, False)
    assert rel_err(recon, r64) < RTOL and rel_err(ep["embedding"], ep64["embedding"]) < RTOL


def test_autograd_through_public_model_api():
    b, n = 2, 256
    v, p64, visible, target, cls, trans, axag, noise = _setup("dgcnn", b, n, seed=8)
    x64, _ = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
    x = x64.float().cuda()
    v.flat.requires_grad_(True)
    recon, rot, tvec, ep = M.get_model_dgcnn_mean_6d(x, True, True, 10, bn_decay=0.9, variables=v)
    loss = (recon ** 2).sum() + rot.sum() + (tvec * 2).sum()
    loss.backward()
    params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p64.items()}
    override = [i.cpu().long() for i in ep["nn_idx"]]
    r64, rot64, t64, _ = MR.get_model_dgcnn_mean_6d(x64, params, True, True, 10, 0.9, nn_idx_override=override)
    ((r64 ** 2).sum() + rot64.sum() + (t64 * 2).sum()).backward()
    flat_grad = v.flat.grad
    for name in ("dgcnn1/weights", "dgcnn4/bn/gamma", "dgcnn_agg/weights", "dgcnn_output/biases", "dgcnn_rot_fc2/weights"):
        o, shape = v.index[name]
        g = flat_grad[o:o + int(np.prod(shape))].view(shape)
        assert l2_err(g, params[name].grad) < RTOL, name


def test_cuda_graph_replay_equals_eager():
    b, n = 4, 256
    v1, _, visible, target, cls, trans, axag, noise = _setup("dgcnn", b, n, seed=9)
    v2, *_ = _setup("dgcnn", b, n, seed=9)
    dev = lambda t: t.cuda().contiguous()  # noqa: E731
    args = (dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise))
    eager = CloudAAETrainer(batch_size=b, num_point=n, variables=v1)
    graph = CloudAAETrainer(batch_size=b, num_point=n, variables=v2)
    graph.capture(*args)
    for it in range(3):
        le = eager.train_step(*args).clone()
        lg = graph.replay().clone()
        torch.cuda.synchronize()
        # step 1: same parameters, deterministic forward -> identical losses.  Later steps: fp32 atomics
        # reorder the EdgeConv scatter sums and early Adam updates are sign-like, so trajectories drift.
        assert torch.allclose(le, lg, rtol=1e-6 if it == 0 else 2e-2, atol=1e-6), (it, le, lg)
    assert eager.state[0].item() == graph.state[0].item() == 3
    assert l2_err(v2.flat, v1.flat) < 1e-2


def test_fused_gemm_statistics_equal_the_separate_statistics_pass():
    """At B*N >= 18944 rows the dgcnn_agg forward GEMM sums the batch-norm statistics of its own output tiles
    (caae_gemm_tf32_stats); the BN coefficients and the losses must equal those of the separate caae_col_stats pass."""
    b, n = 80, 256
    v, p64, visible, target, cls, trans, axag, noise = _setup("dgcnn", b, n, seed=9)
    dev = lambda t: t.cuda().contiguous()  # noqa: E731
    got = []
    for fused in (True, False):
        tr = CloudAAETrainer(batch_size=b, num_point=n, model="dgcnn", variables=v, precision="tf32")
        tr.engine.fused_stats = fused
        tr.decay.fill_(0.9)
        before = _capi.COUNTER[0]
        losses = tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise)).clone()
        launches = _capi.COUNTER[0] - before
        torch.cuda.synchronize()
        bn = tr.engine.bn["dgcnn_agg"]
        got.append((losses, {k: bn[k].clone() for k in ("scale", "shift", "mean", "invstd")}, launches))
    (l1, bn1, n1), (l0, bn0, n0) = got
    assert n1 == n0 - 1                                          # the statistics pass is gone
    for k in bn1:
        assert torch.allclose(bn1[k], bn0[k], rtol=1e-4, atol=1e-6), k
    assert torch.allclose(l1, l0, rtol=1e-3, atol=1e-6), (l1, l0)


def _knn_ffma(x, c, k):
    b, n, ld = x.shape
    idx = torch.empty(b, n, k, dtype=torch.int32, device="cuda")
    _capi.check(_capi.lib().caae_knn_ffma(b, n, c, k, x.data_ptr(), ld, idx.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "knn_ffma")
    return idx


@pytest.mark.parametrize("b,n,c,ld,k", [(128, 256, 64, 320, 10), (128, 256, 3, 24, 10), (5, 200, 64, 64, 10), (3, 256, 24, 24, 16),
                                        (2, 17, 8, 8, 10), (4, 256, 32, 32, 24)])
@pytest.mark.parametrize("kind", ["features", "clustered", "duplicates", "padded3", "padded40"])
def test_knn_tensor_core_screen_is_bit_identical_to_the_ffma_kernel(b, n, c, ld, k, kind):
    """caae_knn (tcgen05 Gram-matrix screen + exact re-rank of a shortlist, knn_tc.cu) must return exactly the indices
    of the all-pairs fp32 kernel: on post-ReLU-like features with a large common mean (the hard case for the error
    bound: norms >> neighbour distances), on tight clusters, and with mass duplicates (shortlist overflow -> fix-up)."""
    g = torch.Generator("cuda").manual_seed(n * 7 + c + k)
    if kind == "features":
        x = torch.relu(torch.randn(b, n, ld, device="cuda", generator=g) * 0.3 + 1.0)
    elif kind == "clustered":
        centers = torch.randn(b, 8, ld, device="cuda", generator=g)
        x = centers[:, torch.arange(n, device="cuda") % 8] + 1e-3 * torch.randn(b, n, ld, device="cuda", generator=g)
    elif kind == "duplicates":
        x = torch.randn(b, n, ld, device="cuda", generator=g)
        x[:, n // 3:] = x[:, :1]                        # two thirds of every cloud are copies of point 0
    else:
        # convexHull()'s padding (utils/hidden_point_removal.py:38-40): V visible points + random repeats of them
        V = min(int(kind[6:]), n)
        x = torch.randn(b, n, ld, device="cuda", generator=g)
        pick = torch.randint(0, V, (b, n - V), device="cuda", generator=g)
        x[:, V:] = torch.gather(x[:, :V], 1, pick[:, :, None].expand(-1, -1, ld))
    got, want = _knn(x, c, k), _knn_ffma(x, c, k)
    assert torch.equal(got, want)


def test_knn_routing_of_padded_clouds_covers_the_batch_exactly():
    """caae_knn_classify flags the clouds with >= n/8 repeated rows; caae_knn_part(1) + caae_knn_part(2) together must
    equal the all-pairs kernel on every cloud, and each part must leave the other part's clouds untouched."""
    b, n, c, k = 12, 256, 64, 10
    g = torch.Generator("cuda").manual_seed(5)
    x = torch.relu(torch.randn(b, n, c, device="cuda", generator=g) * 0.2 + 1.0)
    for i, V in ((1, 3), (4, 40), (7, 100), (9, 230)):      # 253, 216, 156, 26 repeated rows
        pick = torch.randint(0, V, (n - V,), device="cuda", generator=g)
        x[i, V:] = x[i, pick]
    lib, st = _capi.lib(), torch.cuda.current_stream().cuda_stream
    flags = torch.full((b,), -7, dtype=torch.int32, device="cuda")
    _capi.check(lib.caae_knn_classify(b, n, c, x.data_ptr(), c, flags.data_ptr(), st), "classify")
    assert flags.tolist() == [0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0]   # 26 repeats < n/8 stay on the tensor-core path
    want = _knn_ffma(x, c, k)
    for part in (1, 2):
        idx = torch.full((b, n, k), -1, dtype=torch.int32, device="cuda")
        _capi.check(lib.caae_knn_part(part, flags.data_ptr(), b, n, c, k, x.data_ptr(), c, idx.data_ptr(), st), "part")
        mine = (flags == (1 if part == 2 else 0))
        assert torch.equal(idx[mine], want[mine])
        assert (idx[~mine] == -1).all()


@pytest.mark.gpu
def test_backward_after_a_second_forward_of_the_same_shape_raises():
    """The activations live in one workspace per (variables, batch, num_point): a backward pass against activations that a
    later forward overwrote, or a second backward, must fail loudly instead of returning wrong gradients."""
    b, n = 2, 256
    v, _, visible, _, cls, _, _, noise = _setup("dgcnn", b, n, seed=9)
    x64, _ = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
    x = x64.float().cuda()
    v.flat.requires_grad_(True)
    recon1, _, _, _ = M.get_model_dgcnn_mean_6d(x, True, True, 10, bn_decay=0.9, variables=v)
    recon2, _, _, _ = M.get_model_dgcnn_mean_6d(x * 0.5, True, True, 10, bn_decay=0.9, variables=v)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        recon1.sum().backward()
    recon2.sum().backward(retain_graph=True)          # the latest forward is fine ...
    with pytest.raises(RuntimeError, match="already consumed"):
        recon2.sum().backward()                         # ... once


@pytest.mark.gpu
def test_edge_recording_path_matches_the_two_gather_path(monkeypatch):
    """CLOUDAAE_EDGE_REC=1: the forward EdgeConv apply pass records (positive-neighbour count, centred sum of their
    pre-activations) and the backward batch-norm sums come from a streaming pass (caae_edge_bwd_stats) instead of the
    second gather pass (caae_edge_bwd_reduce).  Same losses, same gradients."""
    from cloudaae_b200 import _capi
    b, n = 4, 256
    outs = []
    for rec in ("0", "1"):
        monkeypatch.setenv("CLOUDAAE_EDGE_REC", rec)
        v, _, visible, target, cls, trans, axag, noise = _setup("dgcnn", b, n, seed=11)
        tr = CloudAAETrainer(batch_size=b, num_point=n, model="dgcnn", variables=v, precision="fp32")
        assert tr.engine.edge_rec == (rec == "1")
        dev = lambda t: t.cuda().contiguous()  # noqa: E731
        before = dict(_capi.CALLS)
        losses = tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise)).clone()
        tr.backward(dev(target))
        torch.cuda.synchronize()
        took = {k: c - before.get(k, 0) for k, c in _capi.CALLS.items()}
        assert (took.get("caae_edge_bwd_stats", 0) == 4) == (rec == "1") and (took.get("caae_edge_bwd_reduce", 0) == 4) == (rec == "0")
        outs.append((losses, tr.v.grad.clone(), tr.engine.hcat.clone()))
    (l0, g0, h0), (l1, g1, h1) = outs
    assert rel_err(h1, h0) < 1e-5                       # forward activations: sc*sum(z-mu) + cnt*(sc mu + sh) vs sum relu(sc z + sh)
    assert torch.allclose(l0, l1, rtol=1e-5, atol=1e-7), (l0, l1)
    assert l2_err(g1, g0) < 1e-3, l2_err(g1, g0)        # the north-star tolerance (fp32 atomics reorder sums in both runs)
