import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import model_ref as MR
import test_gpu_model as T
from cloudaae_b200.train import CloudAAETrainer
b, n = 8, 256
v, p64, visible, target, cls, trans, axag, noise = T._setup("dgcnn", b, n)
tr = CloudAAETrainer(batch_size=b, num_point=n, model="dgcnn", variables=v)
dev = lambda t: t.cuda().contiguous()
tr.decay.fill_(0.9375)
tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise))
torch.cuda.synchronize()
y = tr.engine.yagg.clone().double()
bn = tr.engine.bn["dgcnn_agg"]
sc, sh, mu, istd = [bn[k].double() for k in ("scale", "shift", "mean", "invstd")]
print("mean err", (y.mean(0) - mu).abs().max().item(), "var->invstd err", ((y.var(0, unbiased=False) + 1e-3).rsqrt() - istd).abs().max().item() / istd.abs().max().item())
tr.backward(dev(target)); torch.cuda.synchronize()
d_emb = tr.engine.d_emb.double()
mask = (y * sc + sh) > 0
dy = mask * (d_emb / n).repeat_interleave(n, 0)
dbeta = dy.sum(0)
got = v.grad_of("dgcnn_agg/bn/beta").double()
print("dbeta kernel vs torch-on-my-tensors:", ((got - dbeta).abs().max() / dbeta.abs().max()).item())
yhat = (y - mu) * istd
dgamma = (dy * yhat).sum(0)
print("dgamma:", ((v.grad_of("dgcnn_agg/bn/gamma").double() - dgamma).abs().max() / dgamma.abs().max()).item())
# oracle
x64, mean64 = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
override = [i.view(b, n, -1).cpu().long() for i in tr.engine.idx]
params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p64.items()}
total, aux = MR.train_losses(params, x64, mean64, target.double(), trans.double(), axag.double(), 0.9375, nn_idx_override=override)
total.backward()
ref = params["dgcnn_agg/bn/beta"].grad
err = (got.cpu() - ref).abs()
w = err.argmax().item()
print("vs oracle worst ch", w, got[w].item(), ref[w].item(), "max|ref|", ref.abs().max().item())
lb = aux["end_points"]["layer_before_embedding"].reshape(b * n, 1024)
act_ref = (lb > 0)
print("mask mismatches:", (act_ref != mask.cpu()).sum().item(), "of", mask.numel(), "in worst ch:", (act_ref[:, w] != mask.cpu()[:, w]).sum().item())
print("emb rel err", T.rel_err(tr.engine.emb, aux["end_points"]["embedding"]))
