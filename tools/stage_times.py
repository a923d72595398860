"""Per-stage time of the train step: each stage captured as its OWN CUDA graph and replayed (so the
numbers include the same launch-gap behaviour as the real single-graph step, without eager overhead).
Debug aid for the GPU box; never a bench value.

    python tools/stage_times.py [--batch 128] [--iters 30]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cloudaae_b200 import _capi  # noqa: E402
from cloudaae_b200.models.pointnet_ycb_23_decoder_4 import NUM_CLASS, _Engine  # noqa: E402
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz  # noqa: E402
from cloudaae_b200.train import CloudAAETrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=bench.TRAIN_B)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--detail", action="store_true", help="also time single encoder calls")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N = args.batch, bench.TRAIN_N
    tr = CloudAAETrainer(batch_size=B, num_point=N, device=dev, seed=0)
    syn = SegmentSynthesizer(load_models_xyz(device=dev), B, N, seed=1234)
    bt = {k: torch.from_numpy(v).to(dev) for k, v in bench.pose_batches(B, seed=0, pool=1)[0].items()}
    c, ax, tl = (bt[k] for k in bench.TRAIN_KEYS)
    nmain = 9
    tot = 0.0
    for si, (name, ms, launches) in enumerate(measure(tr, syn, c, ax, tl, args.iters, args.detail)):
        if name != "whole_step" and si < nmain:
            tot += ms
        print(f"{name:26s} {ms * 1000:9.1f} us   {launches:4d} C-ABI launches", flush=True)
        if si == nmain - 1:
            print(f"{'sum of stages':22s} {tot * 1000:9.1f} us")


def measure(tr, syn, c, ax, tl, iters=30, detail=True):
    """[(name, ms, C-ABI launches)] for the stages of one train step (first 9 entries) and, with detail, for single
    kernels of the encoder — each captured as its own CUDA graph and replayed `iters` times (warm caches)."""
    dev = tr.dev
    B, N = tr.B, tr.N
    eng, p, M = tr.engine, _Engine._p, tr.M
    # one eager step so that every buffer holds sane values
    tr.train_step_online(syn, c, ax, tl)
    torch.cuda.synchronize()
    vis, tgt, noise = syn.visible, syn.target, syn.noise

    def st_synth():
        syn.synthesize(c, ax, tl)

    def st_prepare():
        tr._c("caae_step_begin", p(tr.state), B)
        tr._c("caae_prepare_input", B, N, vis.shape[1], p(vis), p(noise), p(c), NUM_CLASS, p(tr.x), p(tr.mean))

    def st_encoder_fwd():
        saved = eng.forward_fc
        eng.forward_fc = lambda *a, **k: None
        try:
            eng.forward(tr.x, True, True, tr.decay)
        finally:
            eng.forward_fc = saved

    def st_fc_fwd():
        eng.forward_fc(True, tr.decay)

    def st_losses():
        recon, rot, trans = (eng.fc_y[br[-1]] for br in eng.branches)
        tr._c("caae_pose_losses", B, p(rot), p(ax), p(trans), p(tr.mean), p(tl), 1.0 / B, 10.0 / B, p(tr.per_rot),
              p(tr.per_trans), p(tr.d_rot), p(tr.d_trans), p(tr.trans_pred))
        tr._c("caae_add_cloud_vec", B, M, p(recon), p(tr.mean), p(tr.recon))
        tr._c("caae_nn_distance", B, M, p(tr.recon), M, p(tgt), p(tr.dist1), p(tr.idx1), p(tr.dist2), p(tr.idx2))
        tr._c("caae_loss_reduce", B * M, p(tr.dist1), p(tr.dist2), B, p(tr.per_trans), p(tr.per_rot), p(tr.losses))
        tr._c("caae_nn_distance_grad", B, M, p(tr.recon), M, p(tgt), p(tr.gconst), p(tr.idx1), p(tr.gconst), p(tr.idx2),
              p(tr.d_recon), p(tr.d_target))

    def st_fc_bwd():
        saved = eng.backward_encoder
        eng.backward_encoder = lambda: None
        try:
            eng.backward(tr.d_recon.view(B, 3 * M), tr.d_rot, tr.d_trans)
        finally:
            eng.backward_encoder = saved

    def st_encoder_bwd():
        eng.backward_encoder()

    def st_adam():
        tr.apply_gradients()

    def st_whole():
        tr.train_step_online(syn, c, ax, tl)

    def detail_stages():
        """Single C-ABI calls of the encoder (layer 2: cout 64, layer 4: cout 128, dgcnn_agg), warm caches."""
        R, k = B * N, eng.k
        out = []
        for l in (1, 3):
            scope, ci, co = f"dgcnn{l + 1}", eng.cins[l], eng.couts[l]
            feat = eng.hcat[:, eng.offs[l - 1]:]
            bn = eng.bn[scope]
            d_out = eng.d_hcat[:, eng.offs[l]:]
            args = (B, N, k, co, p(eng.pq[l]), 2 * co, p(eng.idx[l]), p(bn["scale"]), p(bn["shift"]), p(bn["mean"]),
                    p(bn["invstd"]))
            nparts = eng.lib.caae_edge_parts(B, N, k, co, 2 * co)
            out += [
                (f"L{l + 1} knn (tensor-core part)", lambda feat=feat, l=l: eng._c("caae_knn_part", 1, p(eng.knn_flags), B, N, 64, k, p(feat), 320, p(eng.idx[l]))),
                (f"L{l + 1} knn (all-pairs part: padded clouds)", lambda feat=feat, l=l: eng._c("caae_knn_part", 2, p(eng.knn_flags), B, N, 64, k, p(feat), 320, p(eng.idx[l]))),
                (f"L{l + 1} proj gemm", lambda feat=feat, l=l, ci=ci, co=co: eng._gemm_fwd(R, 2 * co, ci, feat, eng.hcat_lo[:, eng.offs[l - 1]:] if eng.x3 else None, 320, eng.wf[l], eng.wf_lo[l], 2 * co, eng.pq[l], 2 * co, eng.bf[l])),
                (f"L{l + 1} edge_stats", lambda l=l, co=co: eng._c("caae_edge_stats", B, N, k, co, p(eng.pq[l]), 2 * co, p(eng.idx[l]), p(eng.parts))),
                (f"L{l + 1} bn_finalize", lambda scope=scope, nparts=nparts: eng._bn_coeffs(scope, True, nparts, R * k, tr.decay)),
                (f"L{l + 1} edge_apply", lambda l=l, co=co, bn=bn, scope=scope, nparts=nparts: eng._c("caae_edge_apply_fused", B, N, k, co, p(eng.pq[l]), 2 * co, p(eng.idx[l]), p(eng.parts), nparts, float(R * k), p(eng.v[f"{scope}/bn/gamma"]), p(eng.v[f"{scope}/bn/beta"]), None, None, p(tr.decay), p(bn["scale"]), p(bn["shift"]), p(bn["mean"]), p(bn["invstd"]), p(eng.hcat[:, eng.offs[l]:]), 320, p(eng.hcat_lo[:, eng.offs[l]:]), *((p(eng.pos_cnt_e[:, eng.offs[l]:]), p(eng.pos_sum_e[:, eng.offs[l]:]), 320) if eng.edge_rec else (None, None, 0)))),
                (f"L{l + 1} edge_bwd_stats", lambda d_out=d_out, bn=bn, l=l, co=co: eng._c("caae_edge_bwd_stats", B, N, k, co, 2 * co, p(d_out), 320, p(eng.pos_cnt_e[:, eng.offs[l]:]), p(eng.pos_sum_e[:, eng.offs[l]:]), 320, p(bn["invstd"]), p(eng.parts))),
                (f"L{l + 1} edge_bwd_reduce" + (" (replaced by edge_bwd_stats)" if eng.edge_rec else ""), lambda args=args, d_out=d_out: eng._c("caae_edge_bwd_reduce", *args, p(d_out), 320, p(eng.parts))),
                (f"L{l + 1} edge_bwd_apply", lambda args=args, d_out=d_out, bn=bn, l=l, co=co: eng._c("caae_edge_bwd_apply", *args, p(bn["coef"]), p(d_out), 320, p(eng.d_pq[l]), 2 * co)),
                (f"L{l + 1} wgrad gemm", lambda feat=feat, l=l, ci=ci, co=co: eng._gemm(1, 0, ci, 2 * co, R, feat, 320, eng.d_pq[l], 2 * co, eng.d_wf[l], 2 * co)),
                (f"L{l + 1} dgrad gemm", lambda l=l, ci=ci, co=co: eng._gemm(0, 1, R, ci, 2 * co, eng.d_pq[l], 2 * co, eng.wf[l], 2 * co, eng.d_hcat[:, eng.offs[l - 1]:], 320, None, 1)),
            ]
        bn = eng.bn["dgcnn_agg"]
        W = tr.v["dgcnn_agg/weights"]
        out += [
            ("L1 knn (xyz, tensor-core part)", lambda: eng._c("caae_knn_part", 1, p(eng.knn_flags), B, N, 3, k, p(tr.x), eng.D, p(eng.idx[0]))),
            ("L1 knn (xyz, all-pairs part)", lambda: eng._c("caae_knn_part", 2, p(eng.knn_flags), B, N, 3, k, p(tr.x), eng.D, p(eng.idx[0]))),
            ("knn classify", lambda: eng._c("caae_knn_classify", B, N, eng.D, p(tr.x), eng.D, p(eng.knn_flags))),
            ("agg gemm fwd (+stats)", lambda: eng._dense_fwd("dgcnn_agg", eng.hcat, 320, R, True, tr.decay, eng.yagg, None, x_lo=eng.hcat_lo if eng.x3 else None)),
            ("agg bn_act_pool", lambda: eng._c("caae_bn_act_pool", B, N, 1024, p(eng.yagg), 1024, p(bn["scale"]), p(bn["shift"]), 0, p(eng.emb), None, p(eng.pos_cnt), p(eng.pos_sum))),
            ("agg bn_bwd (finalize + apply)", lambda: (
                eng._c("caae_bn_pool_bwd_finalize", 1024, B, N, p(eng.d_emb), 1024, 1.0 / N, p(eng.pos_cnt), p(eng.pos_sum), p(bn["mean"]), p(bn["invstd"]),
                       p(tr.v["dgcnn_agg/bn/gamma"]), p(bn["coef"]), p(tr.v.grad_of("dgcnn_agg/bn/gamma")), p(tr.v.grad_of("dgcnn_agg/bn/beta"))),
                eng._c("caae_bn_act_bwd_apply", R, 1024, p(eng.yagg), 1024, p(bn["scale"]), p(bn["shift"]), p(bn["mean"]), p(bn["invstd"]), p(bn["coef"]),
                       p(eng.d_emb), 1024, N, 1.0 / N, 1, None, p(eng.yagg), 1024))),
            ("agg wgrad gemm", lambda: eng._dense_wgrad("dgcnn_agg", eng.hcat, 320, R, eng.yagg, False)),
            ("agg dgrad gemm", lambda: eng._gemm(0, 1, R, 320, 1024, eng.yagg, 1024, W, 1024, eng.d_hcat, 320)),
            ("nn_distance fwd", lambda: tr._c("caae_nn_distance", B, M, p(tr.recon), M, p(tgt), p(tr.dist1), p(tr.idx1), p(tr.dist2), p(tr.idx2))),
            ("nn_distance bwd", lambda: tr._c("caae_nn_distance_grad", B, M, p(tr.recon), M, p(tgt), p(tr.gconst), p(tr.idx1), p(tr.gconst), p(tr.idx2), p(tr.d_recon), p(tr.d_target))),
            ("hpr_select (both problems)", lambda: syn._c("caae_hpr_select_pair", B, syn.nm + syn.no, p(syn.flip_all), syn.N, p(syn.pad_u), p(syn.visible), p(syn.num_vis), syn.nm, p(syn.flip_org), 4 * syn.N, p(syn.pad_u_org), p(syn.target), p(syn.num_vis_org), p(syn.points), syn.nm + syn.no)),
        ]
        return out

    stages = [("synthesis", st_synth), ("prepare_input", st_prepare), ("encoder_fwd", st_encoder_fwd),
              ("fc_fwd", st_fc_fwd), ("losses+chamfer_bwd", st_losses), ("fc_bwd", st_fc_bwd),
              ("encoder_bwd", st_encoder_bwd), ("adam", st_adam), ("whole_step", st_whole)]
    if detail:
        stages += detail_stages()
    results = []
    for si, (name, fn) in enumerate(stages):
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        before = _capi.COUNTER[0]
        with torch.cuda.graph(g):
            fn()
        launches = _capi.COUNTER[0] - before
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        results.append((name, a.elapsed_time(b) / iters, launches))
        del g
    return results


if __name__ == "__main__":
    main()
