import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import model_ref as MR
import test_gpu_model as T
from cloudaae_b200.train import CloudAAETrainer
for model, b in (("dgcnn", 32), ("dgcnn", 64)):
    n = 256
    for prec in ("fp32", "tf32"):
        v, p64, visible, target, cls, trans, axag, noise = T._setup(model, b, n)
        tr = CloudAAETrainer(batch_size=b, num_point=n, model=model, variables=v, precision=prec)
        dev = lambda t: t.cuda().contiguous()
        tr.decay.fill_(0.9375)
        losses = tr.forward_losses(dev(visible), dev(target), dev(cls), dev(trans), dev(axag), dev(noise)).clone()
        tr.backward(dev(target)); torch.cuda.synchronize()
        x64, mean64 = MR.prepare_input(visible.double(), cls, noise.double(), num_point=n)
        override = [i.view(b, n, -1).cpu().long() for i in tr.engine.idx] if model == "dgcnn" else None
        amax = tr.engine.argmax.cpu().long() if model == "pn" else None
        params = {k: t.clone().requires_grad_(not k.endswith(("ema_mean", "ema_var"))) for k, t in p64.items()}
        total, aux = MR.train_losses(params, x64, mean64, target.double(), trans.double(), axag.double(), 0.9375,
                                     nn_idx_override=override, model=model, argmax_override=amax)
        total.backward()
        print(model, prec, "emb", f"{T.rel_err(tr.engine.emb, aux['end_points']['embedding']):.2e}", "recon", f"{T.rel_err(tr.recon, aux['recon']):.2e}",
              "rot", f"{T.rel_err(tr.engine.fc_y[tr.engine.branches[1][-1]], aux['rot_pred']):.2e}", "trans", f"{T.rel_err(tr.trans_pred, aux['trans_pred']):.2e}",
              "losses", [f"{abs(a-b_)/abs(b_):.1e}" for a, b_ in zip(losses.tolist(), [total.item(), aux['chamfer'].item(), aux['trans'].item(), aux['rot'].item()])])
        errs = {n_: T.l2_err(v.grad_of(n_), params[n_].grad) for n_ in v.trainable_names() if not (n_.endswith("biases") and (n_.rsplit("/",1)[0]+"/bn/gamma") in v)}
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print("   grads worst l2:", [(k, f"{e:.1e}") for k, e in top])
