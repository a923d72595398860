import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import model_ref as MR
import test_gpu_model as T
from cloudaae_b200.train import CloudAAETrainer
for model in ("dgcnn", "pn"):
    b, n = 8, 256
    v, p64, visible, target, cls, trans, axag, noise = T._setup(model, b, n)
    tr = CloudAAETrainer(batch_size=b, num_point=n, model=model, variables=v)
    dev = lambda t: t.cuda().contiguous()
    tr.decay.fill_(0.9375)
    tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise))
    tr.backward(dev(target)); torch.cuda.synchronize()
    x64, mean64 = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
    override = [i.view(b, n, -1).cpu().long() for i in tr.engine.idx] if model == "dgcnn" else None
    amax = tr.engine.argmax.cpu().long() if model == "pn" else None
    for dt in (torch.float64, torch.float32):
        params = {k: t.to(dt).clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p64.items()}
        total, aux = MR.train_losses(params, x64.to(dt), mean64.to(dt), target.to(dt), trans.to(dt), axag.to(dt), 0.9375,
                                     nn_idx_override=override, model=model, argmax_override=amax)
        total.backward()
        errs = {n_: T.rel_err(v.grad_of(n_), params[n_].grad) for n_ in v.trainable_names() if not (n_.endswith("biases") and (n_.rsplit("/",1)[0]+"/bn/gamma") in v)}
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        print(model, dt, "total", total.item(), "worst:", [(k, f"{e:.2e}") for k, e in top])
