import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import model_ref as MR
import test_gpu_model as T
from cloudaae_b200.train import CloudAAETrainer
model, b, n = "dgcnn", 3, 128
for seed in (3, 4):
    v, p64, visible, target, cls, trans, axag, noise = T._setup(model, b, n, seed=seed)
    tr = CloudAAETrainer(batch_size=b, num_point=n, model=model, variables=v, precision="fp32")
    dev = lambda t: t.cuda().contiguous()
    tr.decay.fill_(0.9375)
    tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise))
    tr.backward(dev(target)); torch.cuda.synchronize()
    x64, mean64 = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
    override = [i.view(b, n, -1).cpu().long() for i in tr.engine.idx]
    params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p64.items()}
    # layer by layer, keeping the intermediate features (and their gradients)
    net = x64.clone().requires_grad_(True); nets = []
    feats = []
    for li, scope in enumerate(("dgcnn1", "dgcnn2", "dgcnn3", "dgcnn4")):
        adj = MR.pairwise_xyz_distance(net)
        want = MR.knn(adj, 10)
        idx = override[li]
        dsel = torch.gather(adj, 2, idx); dwant = torch.gather(adj, 2, want)
        uniq = all(len(set(idx[i, j].tolist())) == 10 for i in range(b) for j in range(n))
        print(f"seed {seed} layer {li+1}: knn agreement {(idx == want).float().mean():.4f}, max |d_sel - d_want| {(dsel - dwant).abs().max():.2e} (scale {adj.abs().max():.2e}), unique {uniq}, sorted {(dsel[:, :, 1:] - dsel[:, :, :-1] >= -1e-9 * adj.abs().max()).all().item()}")
        edge = MR.get_edge_feature(net, idx)
        net = MR.conv2d_1x1(edge, params, scope, True, 0.9375, None).mean(dim=-2, keepdim=True)
        net.retain_grad(); nets.append(net)
        got = tr.engine.hcat[:, tr.engine.offs[li]:tr.engine.offs[li] + tr.engine.couts[li]].double().cpu().view(b, n, 1, -1)
        print(f"   feature rel err {T.rel_err(got, net):.2e}")
