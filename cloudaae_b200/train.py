"""Training step of train_cloudAAE_ycbv.py (graph section :189-273) on one B200, data-parallel over N.

``CloudAAETrainer.train_step`` runs, on the current CUDA stream and without host synchronisation
(so the whole step can be captured in a CUDA graph):

  step/bn_decay state -> input prep (noise, mean-normalise, one-hot) -> get_model_dgcnn_mean_6d
  forward -> recon + mean -> chamfer nn_distance -> pose losses (rotation in float64) -> total loss
  -> backward of everything -> [NCCL allreduce of the flat gradient] -> TF-style Adam.

Data parallelism: one process per GPU, per-GPU batch fixed, batch-norm statistics
stay per replica (the reference has a single replica and no sync-BN), ONE allreduce(sum) of the flat
fp32 gradient buffer per step followed by a 1/world scale inside the Adam kernel.
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import _capi
from .models.pointnet_ycb_23_decoder_4 import NUM_CLASS, Variables, _Engine, dgcnn_layers, pn_layers
from .parallel import BucketedAllReduce, broadcast_variables


class CloudAAETrainer:
    def __init__(self, batch_size: int = 128, num_point: int = 256, k_neighbor: int = 10, learning_rate: float = 0.0008,
                 model: str = "dgcnn", device="cuda", seed: int = 0, process_group=None,
                 variables: Variables | None = None, precision: str | None = None):
        self.B, self.N, self.k = batch_size, num_point, k_neighbor
        self.lr, self.beta1, self.beta2, self.eps = learning_rate, 0.9, 0.999, 1e-8
        self.dev = torch.device(device)
        self.D = 3 + NUM_CLASS
        layers = dgcnn_layers(num_point, self.D) if model == "dgcnn" else pn_layers(num_point, self.D)
        self.v = variables if variables is not None else Variables(layers, device=self.dev, seed=seed)
        self.engine = _Engine(self.v, model, batch_size, num_point, self.D, k_neighbor if model == "dgcnn" else 0,
                              precision=precision)
        self.lib = _capi.lib()
        self.pg = process_group
        self.world = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        f32 = dict(dtype=torch.float32, device=self.dev)
        B, M = batch_size, 4 * num_point
        self.M = M
        self.adam_m = torch.zeros_like(self.v.flat)
        self.adam_v = torch.zeros_like(self.v.flat)
        self.state = torch.zeros(4, dtype=torch.int32, device=self.dev)  # {step, adam_t, bn_decay(float bits)}
        self.decay = self.state.view(torch.float32)[2:3]
        self.x = torch.empty(B, num_point, self.D, **f32)
        self.mean = torch.empty(B, 3, **f32)
        self.recon = torch.empty(B, M, 3, **f32)
        self.dist1 = torch.empty(B, M, **f32); self.dist2 = torch.empty(B, M, **f32)
        self.idx1 = torch.empty(B, M, dtype=torch.int32, device=self.dev)
        self.idx2 = torch.empty(B, M, dtype=torch.int32, device=self.dev)
        self.gconst = torch.full((B, M), 1000.0 / (B * M), **f32)  # d(1000*mean(dist1+dist2))/d dist
        self.d_recon = torch.empty(B, M, 3, **f32)
        self.d_target = torch.empty(B, M, 3, **f32)
        self.per_rot = torch.empty(B, dtype=torch.float64, device=self.dev)
        self.per_trans = torch.empty(B, **f32)
        self.d_rot = torch.empty(B, 3, **f32); self.d_trans = torch.empty(B, 3, **f32)
        self.trans_pred = torch.empty(B, 3, **f32)
        self.losses = torch.zeros(4, **f32)  # total, chamfer, trans, rot
        self._loss_stream = torch.cuda.Stream(self.dev, priority=-1) if self.engine.concurrent else None
        self._loss_pending = False
        self._graph = None
        self._static = None
        self.launches_per_step = 0
        # gradient exchange in three buckets, each started the moment its gradients are final: 2 = everything after the
        # encoder (FC decoder + pose heads, 94 % of the bytes: after the FC backward), 1 = the last encoder convolution
        # (dgcnn_agg / pn_conv5, 1.3 MB: right after its weight gradient, at the START of the encoder backward), 0 = the
        # first four encoder layers (0.1 MB: at the end of backward — the only exchange left exposed in front of Adam).
        enc_last = "dgcnn_agg" if model == "dgcnn" else "pn_conv5_encoder"
        first = self.v.index[f"{enc_last}/weights"][0]
        split = self.v.index[f"{enc_last}/bn/gamma"][0] + 1024
        split = ((split + 31) // 32) * 32
        self.reducer = BucketedAllReduce(self.v.grad, [0, first, split, self.v.grad.numel()], group=process_group)
        self.engine.after_fc_backward = (lambda: self.reducer.start(2)) if self.world > 1 else None
        self.engine.after_last_conv_wgrad = (lambda: self.reducer.start(1)) if self.world > 1 else None
        if self.world > 1:
            broadcast_variables(self.v.flat, self.v.ema, group=process_group)

    # ------------------------------------------------------------------ save / restore (train_cloudAAE_ycbv.py:423-430)
    def state_dict(self):
        """Everything a restart needs: parameters and moving averages, Adam moments, and the device-side step state
        (step count, Adam t, bn_decay) — so bias correction and the bn_decay staircase continue where they stopped.
        In data-parallel runs the batch-norm moving averages are per replica (the reference has one replica); save
        rank 0's, or average them across ranks first (`average_ema`)."""
        return {"variables": self.v.state_dict(), "adam_m": self.adam_m.detach().clone(), "adam_v": self.adam_v.detach().clone(),
                "state": self.state.detach().clone()}

    def load_state_dict(self, sd):
        with torch.no_grad():
            self.v.load_state_dict(sd["variables"])
            self.adam_m.copy_(sd["adam_m"].to(self.dev)); self.adam_v.copy_(sd["adam_v"].to(self.dev))
            self.state.copy_(sd["state"].to(self.dev))

    def average_ema(self):
        """Average the per-replica batch-norm moving averages across the process group (call before saving)."""
        if self.world > 1:
            torch.distributed.all_reduce(self.v.ema, group=self.pg)
            self.v.ema.div_(self.world)

    # ------------------------------------------------------------------
    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _c(self, name, *args):
        _capi.check(getattr(self.lib, name)(*args, self._st()), name)

    def forward_losses(self, visible, target, class_id, translation, axisangle, noise, start_heads_backward=False):
        """Forward pass + losses (+ the loss gradients w.r.t. the network outputs).  start_heads_backward:
        launch the pose heads' backward pass as soon as their loss gradients exist (side streams), so it
        overlaps the chamfer forward/backward of the decoder branch; `backward` must follow."""
        B, N, M = self.B, self.N, self.M
        p = _Engine._p
        assert visible.is_contiguous() and target.is_contiguous() and target.shape == (B, M, 3)
        assert class_id.dtype == torch.int32
        self._c("caae_prepare_input", B, N, visible.shape[1], p(visible), p(noise), p(class_id), NUM_CLASS, p(self.x),
                p(self.mean))
        recon, rot, trans, _, _ = self.engine.forward(self.x, True, True, self.decay)
        self._c("caae_pose_losses", B, p(rot), p(axisangle), p(trans), p(self.mean), p(translation), 1.0 / B, 10.0 / B,
                p(self.per_rot), p(self.per_trans), p(self.d_rot), p(self.d_trans), p(self.trans_pred))
        if start_heads_backward:
            self.engine.backward_heads_async(self.d_rot, self.d_trans)
        self._c("caae_add_cloud_vec", B, M, p(recon), p(self.mean), p(self.recon))
        self._c("caae_nn_distance", B, M, p(self.recon), M, p(target), p(self.dist1), p(self.idx1), p(self.dist2),
                p(self.idx2))
        # the reported loss values feed nothing downstream: reduce them next to the backward pass
        side = self._loss_stream
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(self.dev))
            self._loss_pending = True
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            self._c("caae_loss_reduce", B * M, p(self.dist1), p(self.dist2), B, p(self.per_trans), p(self.per_rot),
                    p(self.losses))
        if not start_heads_backward:
            self._join_loss()
        return self.losses

    def _join_loss(self):
        if self._loss_pending:
            torch.cuda.current_stream(self.dev).wait_stream(self._loss_stream)
            self._loss_pending = False

    def backward(self, target):
        B, M = self.B, self.M
        p = _Engine._p
        self._c("caae_nn_distance_grad", B, M, p(self.recon), M, p(target), p(self.gconst), p(self.idx1),
                p(self.gconst), p(self.idx2), p(self.d_recon), p(self.d_target))
        self.engine.backward(self.d_recon.view(B, 3 * M), self.d_rot, self.d_trans)

    def apply_gradients(self):
        p = _Engine._p
        self._join_loss()
        if self.world > 1:
            self.reducer.start(0)
            self.reducer.finish()
        self._c("caae_adam_tf", self.v.flat.numel(), p(self.v.flat), p(self.v.grad), p(self.adam_m), p(self.adam_v),
                p(self.state), self.lr, self.beta1, self.beta2, self.eps, 1.0 / self.world)

    def train_step(self, visible, target, class_id, translation, axisangle, noise=None):
        """One optimisation step.  visible f32[B,P>=N,3] (first N rows are the network input),
        target f32[B,4N,3] (first 4N un-occluded visible points), class_id i32[B],
        translation/axisangle f32[B,3] (labels), noise f32[B,N,3] or None.
        Returns the device tensor [total, chamfer, trans, rot] (no host sync)."""
        self._c("caae_step_begin", _Engine._p(self.state), self.B)
        self.forward_losses(visible, target, class_id, translation, axisangle, noise, start_heads_backward=True)
        self.backward(target)
        self.apply_gradients()
        return self.losses

    # ------------------------------------------------------------------ on-line synthesis
    def train_step_online(self, synthesizer, class_id, axisangle, translation):
        """train_cloudAAE_ycbv.py's full step: pose records -> on-line synthesis (pose transform, occluder,
        hidden point removal, visible-prefix selection, sensor noise) -> train_step."""
        visible, target, noise = synthesizer.synthesize(class_id, axisangle, translation)
        return self.train_step(visible, target, class_id, translation, axisangle, noise)

    # ------------------------------------------------------------------ CUDA graph
    def _capture(self, fn, static, warmup):
        self._replay_fn = None
        self._join_streams = ()
        snap = (self.v.flat.clone(), self.v.ema.clone(), self.adam_m.clone(), self.adam_v.clone(), self.state.clone())
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn(*static)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self._graph = torch.cuda.CUDAGraph()
        before = _capi.COUNTER[0]
        # capture on a HIGH-priority stream: kernel nodes inherit it, so the train step's kernels are
        # dispatched ahead of the low-priority synthesis branch of the pipelined graph
        self._capture_stream = torch.cuda.Stream(self.dev, priority=-1)
        with torch.cuda.graph(self._graph, stream=self._capture_stream):
            fn(*static)
        self.launches_per_step = _capi.COUNTER[0] - before  # C-ABI launches captured into one step
        # warm-up and capture must not count as training steps
        for dst, src in zip((self.v.flat, self.v.ema, self.adam_m, self.adam_v, self.state), snap):
            dst.copy_(src)
        self._static = static
        return static

    def capture(self, visible, target, class_id, translation, axisangle, noise=None, warmup: int = 2):
        """Capture train_step on static input buffers; afterwards `replay()` runs one step per call
        on whatever those buffers hold.  Optimiser/BN state advances exactly as in eager mode."""
        static = tuple(t.clone() if t is not None else None
                       for t in (visible, target, class_id, translation, axisangle, noise))
        return self._capture(self.train_step, static, warmup)

    def capture_online(self, synthesizer, class_id, axisangle, translation, warmup: int = 2):
        """Capture synthesis + train_step in ONE graph; static inputs are the pose records only."""
        static = (class_id.clone(), axisangle.clone(), translation.clone())
        return self._capture(lambda c, a, t: self.train_step_online(synthesizer, c, a, t), static, warmup)

    def capture_online_pipelined(self, synthesizer, class_id, axisangle, translation, warmup: int = 2):
        """Like capture_online, with the data pipeline's prefetch(1) (train_cloudAAE_ycbv.py:115): ONE graph
        whose two parallel branches are the train step on batch i (synthesized during the previous replay)
        and the on-line synthesis of batch i+1 from the pose records currently in the static inputs.  The
        hidden-point-removal kernel (one cloud per CTA, uneven cloud costs) and the many small model
        kernels fill each other's idle SMs.  The first replay trains on the batch synthesized here from
        the arguments.  Every replay still performs one full synthesis and one full train step."""
        B = self.B
        rec = torch.empty(7 * B, dtype=torch.float32, device=self.dev)   # pose records of the NEXT batch
        static = (rec[:B].view(torch.int32), rec[B:4 * B].view(B, 3), rec[4 * B:].view(B, 3))
        for dst, src in zip(static, (class_id, axisangle, translation)):
            dst.copy_(src)
        rec_prev = rec.clone()                                         # records the pending batch was made from
        rec_cur = rec.clone()                                          # ... of the batch being trained on
        cur = torch.empty_like(synthesizer.out_flat)
        n3 = B * self.N * 3
        cur_vis, cur_tgt, cur_noise = cur[:n3].view(B, self.N, 3), cur[n3:5 * n3].view(B, 4 * self.N, 3), \
            cur[5 * n3:].view(B, self.N, 3)
        c_cur, a_cur, t_cur = rec_cur[:B].view(torch.int32), rec_cur[B:4 * B].view(B, 3), rec_cur[4 * B:].view(B, 3)
        side = torch.cuda.Stream(self.dev, priority=0 if os.environ.get("CLOUDAAE_SYNTH_PRIORITY", "low") == "low" else -1)

        def prime():
            synthesizer.synthesize(*static)
            rec_prev.copy_(rec)

        def fn(c, a, t):
            main = torch.cuda.current_stream(self.dev)
            cur.copy_(synthesizer.out_flat)       # hand over the batch synthesized during the previous replay
            rec_cur.copy_(rec_prev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                rec_prev.copy_(rec)
                synthesizer.synthesize(c, a, t)
            self.train_step(cur_vis, cur_tgt, c_cur, t_cur, a_cur, cur_noise)
            main.wait_stream(side)

        prime()
        out = self._capture(fn, static, warmup)
        prime()   # warm-up and capture consumed the primed batch's slot; make the first replay well-defined
        self.prime_pipeline = prime   # re-synthesize the pending batch from the static records (tests, restarts)
        # the graph holds raw addresses of these buffers: they must outlive it
        self._pipeline_keep = (rec, rec_prev, rec_cur, cur, side, synthesizer)
        return out

    def capture_online_decoupled(self, synthesizer, class_id, axisangle, translation, depth: int = 2, warmup: int = 2):
        """On-line synthesis and training as TWO CUDA graphs on two streams with a `depth`-slot hand-over queue
        (the reference's parallel map + prefetch, train_cloudAAE_ycbv.py:96-117): `replay()` trains on the oldest
        queued batch (high-priority stream) and enqueues the synthesis of one new batch from the pose records
        currently in the static inputs (low-priority stream) into the slot it has just freed.  Unlike the
        single-graph pipeline there is no barrier at the step boundary: the long tail of the hidden-point-
        removal kernel (one cloud per CTA, cloud costs spread 5x) runs under the NEXT train step instead of
        idling the GPU.  Every replay still performs one full synthesis and one full train step.
        `prime_pipeline(batches)` fills the queue (list of `depth` record tuples; default: the static records)."""
        B, D = self.B, int(depth)
        assert D >= 1
        rec = torch.empty(7 * B, dtype=torch.float32, device=self.dev)   # static inputs: records of the batch to synthesize
        static = (rec[:B].view(torch.int32), rec[B:4 * B].view(B, 3), rec[4 * B:].view(B, 3))
        for dst, src in zip(static, (class_id, axisangle, translation)):
            dst.copy_(src)
        rec_syn = rec.clone()                                            # what the synthesis graph reads
        rec_cur = rec.clone()                                            # records of the batch being trained on
        cur = torch.empty_like(synthesizer.out_flat)
        slots_x = [torch.empty_like(synthesizer.out_flat) for _ in range(D)]
        slots_r = [torch.empty_like(rec) for _ in range(D)]
        n3 = B * self.N * 3
        cur_vis, cur_tgt, cur_noise = cur[:n3].view(B, self.N, 3), cur[n3:5 * n3].view(B, 4 * self.N, 3), \
            cur[5 * n3:].view(B, self.N, 3)
        views = lambda r: (r[:B].view(torch.int32), r[B:4 * B].view(B, 3), r[4 * B:].view(B, 3))  # noqa: E731
        c_cur, a_cur, t_cur = views(rec_cur)
        c_syn, a_syn, t_syn = views(rec_syn)
        s_syn = torch.cuda.Stream(self.dev, priority=0)
        s_train = torch.cuda.Stream(self.dev, priority=-1)
        main = torch.cuda.current_stream(self.dev)

        # ---- synthesis graph (low-priority stream)
        s_syn.wait_stream(main)
        with torch.cuda.stream(s_syn):
            for _ in range(warmup):
                synthesizer.synthesize(c_syn, a_syn, t_syn)
        main.wait_stream(s_syn)
        synth_graph = torch.cuda.CUDAGraph()
        before = _capi.COUNTER[0]
        with torch.cuda.graph(synth_graph, stream=s_syn):
            synthesizer.synthesize(c_syn, a_syn, t_syn)
        synth_launches = _capi.COUNTER[0] - before
        # ---- train graph (high-priority stream, via the common helper)
        cur.copy_(synthesizer.out_flat); rec_cur.copy_(rec_syn)
        self._capture(lambda: self.train_step(cur_vis, cur_tgt, c_cur, t_cur, a_cur, cur_noise), (), warmup)
        self.launches_per_step += synth_launches
        st = {"k": 0, "ready": [None] * D, "fresh": True}

        def prime(batches=None):
            torch.cuda.synchronize(self.dev)
            for j in range(D):
                src = static if batches is None else batches[j]
                for dst, x in zip((c_syn, a_syn, t_syn), src):
                    dst.copy_(x)
                synthesizer.synthesize(c_syn, a_syn, t_syn)
                slots_x[j].copy_(synthesizer.out_flat); slots_r[j].copy_(rec_syn)
            st["k"], st["ready"], st["fresh"] = 0, [None] * D, True

        def replay():
            main = torch.cuda.current_stream(self.dev)
            slot = st["k"] % D
            if st["fresh"]:                      # the queue was filled on the caller's stream
                s_train.wait_stream(main)
                st["fresh"] = False
            if st["ready"][slot] is not None:
                s_train.wait_event(st["ready"][slot])
            with torch.cuda.stream(s_train):
                cur.copy_(slots_x[slot]); rec_cur.copy_(slots_r[slot])
                consumed = torch.cuda.Event(); consumed.record(s_train)
                self._graph.replay()
                done = torch.cuda.Event(); done.record(s_train)
            s_syn.wait_stream(main)              # the records the caller has just loaded
            s_syn.wait_event(consumed)           # the slot is free
            with torch.cuda.stream(s_syn):
                rec_syn.copy_(rec)
                taken = torch.cuda.Event(); taken.record(s_syn)
                synth_graph.replay()
                slots_x[slot].copy_(synthesizer.out_flat); slots_r[slot].copy_(rec_syn)
                ready = torch.cuda.Event(); ready.record(s_syn)
            st["ready"][slot] = ready
            main.wait_event(taken)               # the caller may overwrite the static records again
            main.wait_event(done)                # ... and sees this step's losses / parameters
            st["k"] += 1
            return self.losses

        prime()
        self.prime_pipeline = prime
        self._replay_fn = replay
        self._join_streams = (s_syn, s_train)
        self.pipeline_depth = D
        self._pipeline_keep = (rec, rec_syn, rec_cur, cur, slots_x, slots_r, s_syn, s_train, synth_graph, synthesizer, st)
        return static

    def join(self):
        """Order the caller's stream after everything `replay()` has enqueued so far, including synthesis
        work still running ahead on the side streams of the decoupled pipeline."""
        main = torch.cuda.current_stream(self.dev)
        for s_ in getattr(self, "_join_streams", ()):
            main.wait_stream(s_)

    def replay(self):
        fn = getattr(self, "_replay_fn", None)
        if fn is not None:
            return fn()
        self._graph.replay()
        return self.losses
