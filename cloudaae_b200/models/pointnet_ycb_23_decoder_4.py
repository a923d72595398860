"""CloudAAE networks — drop-in for the reference's ``models/pointnet_ycb_23_decoder_4.py``.

``get_model_dgcnn_mean_6d(point_cloud, is_training_pl_encoder, is_training, k_neighbor, bn_decay=None)``
(:327-455, the network both reference scripts run) and ``get_model_pn(point_cloud, is_training,
bn_decay=None)`` (:23-89) keep the reference's names, argument order and return tuple
``(net_recon, net_rot, net_trans, end_points)``.  Tensors are torch CUDA tensors; every layer runs on
the hand-written sm_100a kernels behind the ``caae_*`` C ABI (no torch.nn, no cuBLAS on this path).

TensorFlow keeps variables in named scopes of a graph; here a :class:`Variables` store plays that
role (same names: ``dgcnn1/weights``, ``dgcnn_agg/bn/gamma`` ...), holding all trainable parameters
in ONE flat fp32 buffer — the operand of the fused Adam kernel and of the data-parallel allreduce.
"""
from __future__ import annotations

import contextlib
import math
import os
from collections import OrderedDict

import torch

from .. import _capi

NUM_CLASS = 21
_ALIGN = 32  # floats; every parameter tensor starts on a 128-byte boundary

# (scope, fan_in, fan_out, has_bn) in the reference's creation order
DGCNN_LAYERS = [
    ("dgcnn1", 48, 64, True), ("dgcnn2", 128, 64, True), ("dgcnn3", 128, 64, True), ("dgcnn4", 128, 128, True),
    ("dgcnn_agg", 320, 1024, True), ("dgcnn_fc1", 1024, 1024, True), ("dgcnn_fc2", 1024, 1024, True),
    ("dgcnn_output", 1024, 3072, False),
    ("dgcnn_rot_fc1", 1024, 512, True), ("dgcnn_rot_fc2", 512, 256, True), ("dgcnn_output_rot", 256, 3, False),
    ("dgcnn_trans_fc1", 1024, 512, True), ("dgcnn_trans_fc2", 512, 256, True), ("dgcnn_output_trans", 256, 3, False),
]


def dgcnn_layers(num_point: int = 256, point_dim: int = 24):
    layers = list(DGCNN_LAYERS)
    layers[0] = ("dgcnn1", 2 * point_dim, 64, True)
    layers[7] = ("dgcnn_output", 1024, num_point * 12, False)
    return layers


def pn_layers(num_point: int = 256, point_dim: int = 24):
    return [
        ("pn_conv1_encoder", point_dim, 64, True), ("pn_conv2_encoder", 64, 64, True),
        ("pn_conv3_encoder", 64, 64, True), ("pn_conv4_encoder", 64, 128, True), ("pn_conv5_encoder", 128, 1024, True),
        ("pn_fc1_decoder", 1024, 1024, True), ("pn_fc2_decoder", 1024, 1024, True),
        ("pn_output", 1024, num_point * 12, False),
        ("pn_rot_fc1", 1024, 512, True), ("pn_rot_fc2", 512, 256, True), ("pn_output_rot", 256, 3, False),
        ("pn_trans_fc1", 1024, 512, True), ("pn_trans_fc2", 512, 256, True), ("pn_output_trans", 256, 3, False),
    ]


class Variables:
    """Named parameter store over flat buffers (the analogue of TF variable scopes + Saver).

    ``flat``/``grad``: all trainable parameters / their gradients (weights, biases, bn/beta, bn/gamma),
    ``ema``: the non-trainable moving averages (bn/ema_mean, bn/ema_var; start at 0 like TF's
    ExponentialMovingAverage shadows of tensors).  Initialisation follows utils/tf_util.py:42-43
    (Xavier-uniform weights), :164-165 (zero biases) and :488-491 (gamma 1, beta 0).
    """

    def __init__(self, layers, device="cuda", seed: int | None = 0):
        self.layers = list(layers)
        self.device = torch.device(device)
        self.index = OrderedDict()
        self.ema_index = OrderedDict()
        off = eoff = 0

        def _add(table, name, shape, o):
            table[name] = (o, tuple(shape))
            return o + ((math.prod(shape) + _ALIGN - 1) // _ALIGN) * _ALIGN

        for scope, fin, fout, has_bn in self.layers:
            off = _add(self.index, f"{scope}/weights", (fin, fout), off)
            off = _add(self.index, f"{scope}/biases", (fout,), off)
            if has_bn:
                off = _add(self.index, f"{scope}/bn/beta", (fout,), off)
                off = _add(self.index, f"{scope}/bn/gamma", (fout,), off)
                eoff = _add(self.ema_index, f"{scope}/bn/ema_mean", (fout,), eoff)
                eoff = _add(self.ema_index, f"{scope}/bn/ema_var", (fout,), eoff)
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.ema = torch.zeros(max(eoff, 1), dtype=torch.float32, device=self.device)
        self.num_trainable = sum(math.prod(s) for _, s in self.index.values())
        if seed is not None:
            self.initialize(seed)

    # -- access
    def _view(self, buf, table, name):
        o, shape = table[name]
        return buf[o:o + math.prod(shape)].view(shape)

    def __getitem__(self, name):
        if name in self.index:
            return self._view(self.flat, self.index, name)
        return self._view(self.ema, self.ema_index, name)

    def __contains__(self, name):
        return name in self.index or name in self.ema_index

    def grad_of(self, name):
        return self._view(self.grad, self.index, name)

    def names(self):
        return list(self.index) + list(self.ema_index)

    def trainable_names(self):
        return list(self.index)

    # -- init / io
    def initialize(self, seed: int):
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            self.flat.zero_(); self.ema.zero_()
            for scope, fin, fout, has_bn in self.layers:
                limit = math.sqrt(6.0 / (fin + fout))
                w = (torch.rand(fin, fout, generator=g, dtype=torch.float64) * 2 - 1) * limit
                self[f"{scope}/weights"].copy_(w.float())
                if has_bn:
                    self[f"{scope}/bn/gamma"].fill_(1.0)

    def load_state_dict(self, state):
        with torch.no_grad():
            for name in self.names():
                if name in state:
                    self[name].copy_(torch.as_tensor(state[name]).to(torch.float32).reshape(self[name].shape))

    def state_dict(self):
        return OrderedDict((n, self[n].detach().clone()) for n in self.names())

    def grads_dict(self):
        return OrderedDict((n, self.grad_of(n).detach().clone()) for n in self.index)


# ---------------------------------------------------------------------------------------------
class _Engine:
    """Forward/backward orchestration over the C ABI for one (model, B, N, k) configuration.

    Workspaces are allocated once; a forward followed by its backward may be captured in a CUDA
    graph.  Activations are [B*N, C] row-major; the four EdgeConv outputs live side by side in the
    320-wide concat buffer ``hcat`` (no concat copy), and EdgeConv layer l reads its input features
    from the slice layer l-1 wrote.
    """

    def __init__(self, variables: Variables, model: str, batch: int, num_point: int, point_dim: int, k: int = 10,
                 precision: str | None = None):
        self.precision = precision or os.environ.get("CLOUDAAE_GEMM", "tf32")
        assert self.precision in ("tf32", "fp32")
        self.v = variables
        self.model = model
        self.B, self.N, self.D, self.k = batch, num_point, point_dim, k
        self.R = batch * num_point
        self.dev = variables.device
        self.lib = _capi.lib()
        f32 = dict(dtype=torch.float32, device=self.dev)
        R, B = self.R, self.B
        self.scopes = {s: (fin, fout, bn) for s, fin, fout, bn in variables.layers}
        self.bn = {}
        for s, (fin, fout, bn) in self.scopes.items():
            if bn:
                self.bn[s] = {k_: torch.empty(fout, **f32) for k_ in ("scale", "shift", "mean", "invstd")}
                self.bn[s]["coef"] = torch.empty(3 * fout, **f32)
        nparts = max(B * ((num_point + 31) // 32), self.lib.caae_col_parts(R), 1)
        self.parts = torch.empty(nparts * 2 * 1024, dtype=torch.float64, device=self.dev)
        self.default_decay = torch.full((1,), 0.9, **f32)  # tf_util.py:494: decay defaults to 0.9
        self.decay_scalar = torch.empty(1, **f32)
        self.emb = torch.empty(B, 1024, **f32)
        self.d_emb = torch.empty(B, 1024, **f32)
        if model == "dgcnn":
            self.cins = [point_dim, 64, 64, 64]
            self.couts = [64, 64, 64, 128]
            self.offs = [0, 64, 128, 192]
            self.hcat = torch.empty(R, 320, **f32)
            self.hcat_lo = torch.empty(R, 320, **f32)   # hcat - tf32(hcat): low part for the split-precision forward GEMMs
            self.d_hcat = torch.empty(R, 320, **f32)
            self.idx = [torch.empty(R, k, dtype=torch.int32, device=self.dev) for _ in range(4)]
            self.pq = [torch.empty(R, 2 * c, **f32) for c in self.couts]
            self.d_pq = [torch.empty(R, 2 * c, **f32) for c in self.couts]   # per layer: wgrad of layer l overlaps layer l-1
            self.wf = [torch.empty(ci, 2 * co, **f32) for ci, co in zip(self.cins, self.couts)]
            self.wf_lo = [torch.empty(ci, 2 * co, **f32) for ci, co in zip(self.cins, self.couts)]
            self.bf = [torch.empty(2 * co, **f32) for co in self.couts]
            self.d_wf = [torch.empty(ci, 2 * co, **f32) for ci, co in zip(self.cins, self.couts)]
            self.yagg = torch.empty(R, 1024, **f32)
            self.pos_cnt = torch.empty(B, 1024, **f32)   # ReLU-mask statistics of the pooled layer (forward -> BN backward)
            self.pos_sum = torch.empty(B, 1024, **f32)
            # EdgeConv layers: per (point, channel) positive-neighbour count and sum of their pre-activations (hcat's pitch)
            self.pos_cnt_e = torch.empty(R, 320, dtype=torch.uint8, device=self.dev)
            self.pos_sum_e = torch.empty(R, 320, **f32)
        else:
            self.enc = ["pn_conv1_encoder", "pn_conv2_encoder", "pn_conv3_encoder", "pn_conv4_encoder",
                        "pn_conv5_encoder"]
            self.enc_y = [torch.empty(R, self.scopes[s][1], **f32) for s in self.enc]
            self.enc_a = [torch.empty(R, self.scopes[s][1], **f32) for s in self.enc[:-1]]
            self.enc_d = [torch.empty(R, self.scopes[s][1], **f32) for s in self.enc[:-1]]
            self.argmax = torch.empty(B, 1024, dtype=torch.int32, device=self.dev)
        # low parts of the forward-GEMM operands that no producer kernel writes directly (weights; pn activations)
        conv_scopes = ["dgcnn_agg"] if model == "dgcnn" else self.enc
        self.w_lo = {s_: torch.empty(self.scopes[s_][0], self.scopes[s_][1], **f32) for s_ in conv_scopes
                     if R * self.scopes[s_][0] * self.scopes[s_][1] >= (1 << 28)}
        self.a_lo = {}
        p = "dgcnn" if model == "dgcnn" else "pn"
        self.prefix = p
        fc1, fc2, out = (f"{p}_fc1", f"{p}_fc2", f"{p}_output") if p == "dgcnn" else \
            ("pn_fc1_decoder", "pn_fc2_decoder", "pn_output")
        self.branches = [[fc1, fc2, out], [f"{p}_rot_fc1", f"{p}_rot_fc2", f"{p}_output_rot"],
                         [f"{p}_trans_fc1", f"{p}_trans_fc2", f"{p}_output_trans"]]
        # The FC GEMMs are split-K (M = batch): their outputs accumulate into zero-filled buffers.  All of them
        # live in ONE flat buffer that a single fill kernel clears off the critical path (side stream, start of
        # forward) instead of one zero-fill launch in front of every GEMM of the chain.
        self.fc_y, self.fc_a, self.fc_d = {}, {}, {}
        n_acc = sum(B * self.scopes[s][1] * (2 if self.scopes[s][2] else 1) for br in self.branches for s in br) + 3 * B * 1024
        self.fc_acc_flat = torch.zeros(n_acc, **f32)
        off = 0

        def _take(rows, cols):
            nonlocal off
            t = self.fc_acc_flat[off:off + rows * cols].view(rows, cols)
            off += rows * cols
            return t

        for br in self.branches:
            for s in br:
                fout = self.scopes[s][1]
                self.fc_y[s] = _take(B, fout)
                if self.scopes[s][2]:
                    self.fc_a[s] = torch.empty(B, fout, **f32)
                    self.fc_d[s] = _take(B, fout)
        self.x0 = None
        self.trained = (False, False)
        self.after_fc_backward = None
        self.after_last_conv_wgrad = None
        # Independent kernel chains run on forked side streams (and become parallel branches of the
        # captured CUDA graph): the three FC branches, weight gradients next to the data-gradient chain,
        # the EdgeConv projection next to the kNN search of the same layer.  Every launch of those chains
        # is far too small to fill 148 SMs on its own.  CLOUDAAE_STREAMS=0 serialises everything.
        self.concurrent = os.environ.get("CLOUDAAE_STREAMS", "1") != "0" and self.dev.type == "cuda"
        self.fused_stats = os.environ.get("CLOUDAAE_FUSED_STATS", "1") != "0"
        self.fuse_finalize = os.environ.get("CLOUDAAE_FUSE_FINALIZE", "1") != "0"
        self.wgrad_floor = 1 << int(os.environ.get("CLOUDAAE_WGRAD_FLOOR", "26"))
        # EdgeConv (CLOUDAAE_EDGE_REC=1): the forward apply pass records (positive-neighbour count, centred sum of their
        # pre-activations) per (point, channel); the batch-norm backward sums then are a streaming pass instead of a second
        # staged gather over the k-neighbour tensor (caae_edge_bwd_stats vs caae_edge_bwd_reduce).  Measured on B200: the
        # backward pass gets 8 / 16 us per layer shorter, the recording apply pass 2 / 9 us longer, the pipelined step is
        # unchanged within noise (1.859 vs 1.860 ms) — so the two-gather form stays the default.
        self.edge_rec = os.environ.get("CLOUDAAE_EDGE_REC", "0") == "1"
        self.edge_recorded = [False] * 4
        # forward GEMMs on the tensor cores: split-precision by default (CLOUDAAE_TF32X3=0: single TF32 pass)
        self.x3 = self.precision == "tf32" and os.environ.get("CLOUDAAE_TF32X3", "1") != "0"
        hi = dict(device=self.dev, priority=-1)   # the model's streams outrank the synthesis branch of a pipelined graph
        self.s_branch = [torch.cuda.Stream(**hi) for _ in range(2)] if self.concurrent else []
        self.s_wgrad = [torch.cuda.Stream(**hi) for _ in range(3)] if self.concurrent else []
        self.s_enc = torch.cuda.Stream(**hi) if self.concurrent else None
        self.s_knn = torch.cuda.Stream(**hi) if self.concurrent else None
        self.knn_flags = torch.zeros(batch, dtype=torch.int32, device=self.dev)
        self.pool_parts = torch.empty(4 * batch * 1024, dtype=torch.float32, device=self.dev)   # pooled-epilogue scratch
        self.d_emb_br = _take(3 * B, 1024).view(3, B, 1024)
        self._heads_pending = False
        # FC stack on the tensor cores (split-precision product): low parts of the weights (one pass over the FC range
        # of the flat parameter buffer per forward), of the activations (written by the BN kernels) and of the gradients
        fc_scopes = [s for br in self.branches for s in br]
        self.fc_tc = self.x3 and os.environ.get("CLOUDAAE_FC_TC", "1") != "0" and B >= 64
        self.fc_lo_start = min(variables.index[f"{s}/weights"][0] for s in fc_scopes)
        self.flat_lo = torch.empty_like(variables.flat) if self.fc_tc else None
        self.emb_lo = torch.empty(B, 1024, **f32)
        self.fc_a_lo = {s: torch.empty_like(t) for s, t in self.fc_a.items()}
        self.fc_d_lo = {s: torch.empty_like(t) for s, t in self.fc_d.items()}
        self.d_out_lo = [torch.empty(B, self.scopes[br[-1]][1], **f32) for br in self.branches]

    # -- helpers
    def _st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _c(self, name, *args):
        _capi.check(getattr(self.lib, name)(*args, self._st()), name)

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def _fork(self, s):
        """Stream `s` continues from everything issued so far on the current stream."""
        s.wait_stream(torch.cuda.current_stream(self.dev))

    def _join(self, s):
        torch.cuda.current_stream(self.dev).wait_stream(s)

    def _on(self, s):
        return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()

    def _w_lo(self, scope):
        o, shape = self.v.index[f"{scope}/weights"]
        return self.flat_lo[o:o + shape[0] * shape[1]].view(shape)

    def _fc_gemm(self, ta, tb, M, N, K, A, A_lo, lda, Bm, B_lo, ldb, C, ldc, bias=None, acc=0):
        """A contraction of the FC stack (M or K = batch).  Tensor cores with the split-precision product when the
        output is at least a tile wide — fp32-grade, so the batch norm over `batch` rows behind it sees the same
        values as the FFMA path; the 3-wide pose outputs and tiny batches stay on the FFMA kernel."""
        if (self.fc_tc and A_lo is not None and B_lo is not None and N >= 128 and M >= 64 and K >= 64 and
                self.lib.caae_gemm_tf32_supported(ta, tb, M, N, K, self._p(A), lda, self._p(Bm), ldb)):
            self._c("caae_gemm_tf32x3", ta, tb, M, N, K, self._p(A), self._p(A_lo), lda, self._p(Bm), self._p(B_lo), ldb,
                    self._p(C), ldc, self._p(bias), acc, None)
        else:
            self._gemm(ta, tb, M, N, K, A, lda, Bm, ldb, C, ldc, bias, acc)

    def _fc_fwd(self, scope, x, x_lo, training, decay):
        """One fully connected layer on [B, fin]: y = x W + b, then training-mode BN + ReLU in one launch
        (tf_util.fully_connected, :321-365).  Returns (activation, its low part)."""
        fin, fout, has_bn = self.scopes[scope]
        B, v = self.B, self.v
        y = self.fc_y[scope]
        self._fc_gemm(0, 0, B, fout, fin, x, x_lo, fin, v[f"{scope}/weights"], self._w_lo(scope) if self.fc_tc else None, fout,
                      y, fout, v[f"{scope}/biases"], 1)
        if not has_bn:
            return None, None
        bn, a = self.bn[scope], self.fc_a[scope]
        a_lo = self.fc_a_lo[scope] if self.fc_tc else None
        if training:
            self._c("caae_fc_bn_fwd", B, fout, self._p(y), fout, self._p(v[f"{scope}/bn/gamma"]),
                    self._p(v[f"{scope}/bn/beta"]), self._p(v[f"{scope}/bn/ema_mean"]), self._p(v[f"{scope}/bn/ema_var"]),
                    self._p(decay), self._p(bn["scale"]), self._p(bn["shift"]), self._p(bn["mean"]),
                    self._p(bn["invstd"]), 1, self._p(a), fout, self._p(a_lo))
        else:
            self._bn_coeffs(scope, False, 1, B, decay)
            self._c("caae_bn_act", B, fout, self._p(y), fout, self._p(bn["scale"]), self._p(bn["shift"]), 1,
                    self._p(a), fout)
            if a_lo is not None:
                self._split(a, fout, B, fout, a_lo, fout)
        return a, a_lo

    def _fc_bn_bwd(self, scope, d):
        """BN + ReLU backward of an FC layer, in place over d [B, fout]; bn/gamma, bn/beta gradients."""
        C = self.scopes[scope][1]
        bn, v = self.bn[scope], self.v
        self._c("caae_fc_bn_bwd", self.B, C, self._p(self.fc_y[scope]), C, self._p(bn["scale"]), self._p(bn["shift"]),
                self._p(bn["mean"]), self._p(bn["invstd"]), self._p(v[f"{scope}/bn/gamma"]), 1, self._p(d), C,
                self._p(d), C, self._p(v.grad_of(f"{scope}/bn/gamma")), self._p(v.grad_of(f"{scope}/bn/beta")),
                self._p(self.fc_d_lo[scope] if self.fc_tc else None))

    def _branch_backward(self, bi, d_out):
        """Backward of FC branch `bi` on the current stream; its weight gradients on their own stream.
        Leaves d(embedding) of the branch in d_emb_br[bi]."""
        B = self.B
        s3, s2, s1 = self.branches[bi][2], self.branches[bi][1], self.branches[bi][0]
        f3, f2, f1 = self.scopes[s3], self.scopes[s2], self.scopes[s1]
        w = self.s_wgrad[bi] if self.concurrent else None
        d_out = d_out.contiguous()
        d2, d1 = self.fc_d[s2], self.fc_d[s1]
        tc = self.fc_tc
        lo = lambda t: t if tc else None  # noqa: E731
        d_out_lo = None
        if tc and f3[1] >= 128:           # the decoder's 3072-wide output gradient (the pose heads' are 3 wide: FFMA)
            d_out_lo = self.d_out_lo[bi]
            self._split(d_out, f3[1], B, f3[1], d_out_lo, f3[1])
        W3, W2, W1 = (self.v[f"{s_}/weights"] for s_ in (s3, s2, s1))
        W3l, W2l, W1l = ((self._w_lo(s_) if tc else None) for s_ in (s3, s2, s1))
        # linear output layer
        if w is not None: self._fork(w)
        with self._on(w):
            self._fc_wgrad(s3, self.fc_a[s2], lo(self.fc_a_lo[s2]), f3[0], B, d_out, d_out_lo, True)
        self._fc_gemm(0, 1, B, f3[0], f3[1], d_out, d_out_lo, f3[1], W3, W3l, f3[1], d2, f3[0], None, 1)
        # fc2: BN+ReLU backward, wgrad, dgrad
        self._fc_bn_bwd(s2, d2)
        if w is not None: self._fork(w)
        with self._on(w):
            self._fc_wgrad(s2, self.fc_a[s1], lo(self.fc_a_lo[s1]), f2[0], B, d2, lo(self.fc_d_lo[s2]), False)
        self._fc_gemm(0, 1, B, f2[0], f2[1], d2, lo(self.fc_d_lo[s2]), f2[1], W2, W2l, f2[1], d1, f2[0], None, 1)
        # fc1
        self._fc_bn_bwd(s1, d1)
        if w is not None: self._fork(w)
        with self._on(w):
            self._fc_wgrad(s1, self.emb, lo(self.emb_lo), f1[0], B, d1, lo(self.fc_d_lo[s1]), False)
        self._fc_gemm(0, 1, B, f1[0], f1[1], d1, lo(self.fc_d_lo[s1]), f1[1], W1, W1l, f1[1], self.d_emb_br[bi], f1[0], None, 1)
        if w is not None: self._join(w)

    def backward_heads_async(self, d_rot, d_trans):
        """Start the backward pass of the two pose heads (on side streams when enabled): their loss
        gradients exist long before the chamfer gradient of the decoder branch.  `backward` joins them."""
        assert self.trained == (True, True), "backward is defined for training-mode batch norm"
        for bi, d in ((1, d_rot), (2, d_trans)):
            s = self.s_branch[bi - 1] if self.concurrent else None
            if s is not None: self._fork(s)
            with self._on(s):
                self._branch_backward(bi, d)
        self._heads_pending = True

    def _gemm(self, ta, tb, M, N, K, A, lda, Bm, ldb, C, ldc, bias=None, acc=0, floor=1 << 28):
        """Dense contraction.  precision 'tf32': the tcgen05 tensor-core kernel (TF32 multiply, fp32
        accumulate) for the LARGE contractions — in practice the three dgcnn_agg / pn_conv5 GEMMs
        (forward, data gradient, weight gradient: >90 % of the step's FLOPs).  Everything else, and
        everything with precision 'fp32', runs on the FFMA kernel: the FC stack (M = batch) is
        weight-bandwidth bound, and its batch-norm backward over only `batch` rows amplifies TF32
        rounding of the pre-activations far beyond the 1e-3 parity budget."""
        fn = "caae_gemm_f32"
        if self._tensor_core(ta, tb, M, N, K, A, lda, Bm, ldb, floor):
            fn = "caae_gemm_tf32"
        self._c(fn, ta, tb, M, N, K, self._p(A), lda, self._p(Bm), ldb, self._p(C), ldc, self._p(bias), acc)

    def _tensor_core(self, ta, tb, M, N, K, A, lda, Bm, ldb, floor=1 << 28):
        # large contractions only (>= 2^28 MACs with >= 64 output columns): besides dgcnn_agg these are the
        # EdgeConv projections and their data / weight gradients (M or K = B*N rows)
        # (never the FC stack: none of its dimensions is the B*N row count)
        return bool(self.precision == "tf32" and M * N * K >= floor and N >= 64 and max(M, K) >= 8192 and
                    self.lib.caae_gemm_tf32_supported(ta, tb, M, N, K, self._p(A), lda, self._p(Bm), ldb))

    def _split(self, x, ldx, rows, cols, lo, ldlo):
        """lo = x - tf32(x): the low part of a split-precision GEMM operand."""
        self._c("caae_split_tf32", rows, cols, self._p(x), ldx, self._p(lo), ldlo)

    def _gemm_fwd(self, M, N, K, A, A_lo, lda, Bm, B_lo, ldb, C, ldc, bias, parts=None):
        """Forward contraction C = A B + bias.  On the tensor cores it is the split-precision product
        (caae_gemm_tf32x3: A B + A_lo B + A B_lo): a single TF32 pass perturbs the encoder features by ~5e-4, which
        the batch norm over the batch axis of the FC layers amplifies beyond the 1e-3 parity budget of the pose
        outputs (measured at B = 128: rot 2.1e-3, gradients 2e-2), and flips 6-14 % of the layer 3/4 neighbours.
        Returns True when the tensor-core path (and `parts`, if given) was taken."""
        if self.x3 and A_lo is not None and B_lo is not None and self._tensor_core(0, 0, M, N, K, A, lda, Bm, ldb):
            self._c("caae_gemm_tf32x3", 0, 0, M, N, K, self._p(A), self._p(A_lo), lda, self._p(Bm), self._p(B_lo), ldb,
                    self._p(C), ldc, self._p(bias), 0, self._p(parts))
            return True
        if self.x3:   # forward values must stay fp32-grade: without the low parts the FFMA kernel, never a single TF32 pass
            self._c("caae_gemm_f32", 0, 0, M, N, K, self._p(A), lda, self._p(Bm), ldb, self._p(C), ldc, self._p(bias), 0)
        else:
            self._gemm(0, 0, M, N, K, A, lda, Bm, ldb, C, ldc, bias)
        return False

    def _bn_coeffs(self, scope, training, nparts, count, decay):
        v, bn = self.v, self.bn[scope]
        C = self.scopes[scope][1]
        if training:
            self._c("caae_bn_finalize", C, self._p(self.parts), nparts, float(count), self._p(v[f"{scope}/bn/gamma"]),
                    self._p(v[f"{scope}/bn/beta"]), self._p(v[f"{scope}/bn/ema_mean"]),
                    self._p(v[f"{scope}/bn/ema_var"]), self._p(decay), self._p(bn["scale"]), self._p(bn["shift"]),
                    self._p(bn["mean"]), self._p(bn["invstd"]))
        else:
            self._c("caae_bn_eval_coeffs", C, self._p(v[f"{scope}/bn/gamma"]), self._p(v[f"{scope}/bn/beta"]),
                    self._p(v[f"{scope}/bn/ema_mean"]), self._p(v[f"{scope}/bn/ema_var"]), self._p(bn["scale"]),
                    self._p(bn["shift"]))

    def _pooled_fwd(self, scope, x, x_lo, ldx, training, want_activation, mode):
        """Inference form of the last encoder convolution: moving-average batch norm is affine, so bias + BN + ReLU + the
        mean (mode 1) / max (mode 2) over a cloud's points are reduced in the GEMM epilogue and the [B*N, 1024]
        activation is never stored (caae_gemm_tf32_pool).  False when the configuration needs the stored activation
        (training statistics, end_points['layer_before_embedding'], num_point != 256, FFMA precision)."""
        fin, fout, _ = self.scopes[scope]
        W = self.v[f"{scope}/weights"]
        if (training or want_activation or self.N != 256 or self.precision != "tf32" or self.R * fin * fout < (1 << 28) or
                os.environ.get("CLOUDAAE_POOL_EPILOGUE", "1") == "0" or
                not self.lib.caae_gemm_tf32_supported(0, 0, self.R, fout, fin, self._p(x), ldx, self._p(W), fout)):
            return False
        bn = self.bn[scope]
        self._bn_coeffs(scope, False, 1, self.R, None)
        w_lo = self.w_lo.get(scope) if (self.x3 and x_lo is not None) else None
        self._c("caae_gemm_tf32_pool", self.R, fout, fin, self._p(x), self._p(x_lo if w_lo is not None else None), ldx, self._p(W),
                self._p(w_lo), fout, self._p(self.v[f"{scope}/biases"]), self._p(bn["scale"]), self._p(bn["shift"]), mode, self.N,
                self._p(self.pool_parts), self._p(self.emb))
        return True

    def _dense_fwd(self, scope, x, ldx, R, training, decay, y, a, x_lo=None):
        """y = x W + b; (BN + ReLU -> a) when the layer has BN (tf_util.conv2d 1x1 / fully_connected)."""
        fin, fout, has_bn = self.scopes[scope]
        W, bias = self.v[f"{scope}/weights"], self.v[f"{scope}/biases"]
        # tall TF32 contractions followed by training-mode BN (dgcnn_agg, pn_conv5): the persistent GEMM sums the
        # columns of each output tile while it sits in shared memory, which saves the separate 134 MB statistics pass
        fused = 0
        if (has_bn and training and self.fused_stats and self.precision == "tf32" and R * fout * fin >= (1 << 28) and
                self.lib.caae_gemm_tf32_supported(0, 0, R, fout, fin, self._p(x), ldx, self._p(W), fout)):
            fused = self.lib.caae_gemm_tf32_stats_parts(R, fout, fin, fout)
            if fused * 2 * fout > self.parts.numel():
                fused = 0
        if self.x3 and scope in self.w_lo and self._tensor_core(0, 0, R, fout, fin, x, ldx, W, fout):
            if x_lo is None:   # no producer wrote the low part of this activation: one elementwise pass
                if scope not in self.a_lo:
                    self.a_lo[scope] = torch.empty(R, fin, dtype=torch.float32, device=self.dev)
                x_lo = self.a_lo[scope]
                self._split(x, ldx, R, fin, x_lo, fin)
            self._gemm_fwd(R, fout, fin, x, x_lo, ldx, W, self.w_lo[scope], fout, y, fout, bias,
                           self.parts if fused else None)
        elif fused:
            self._c("caae_gemm_tf32_stats", R, fout, fin, self._p(x), ldx, self._p(W), fout, self._p(y), fout,
                    self._p(bias), self._p(self.parts))
        else:
            self._gemm(0, 0, R, fout, fin, x, ldx, W, fout, y, fout, bias)
        if has_bn:
            if training and not fused:
                self._c("caae_col_stats", R, fout, self._p(y), fout, self._p(self.parts))
            self._bn_coeffs(scope, training, fused if fused else self.lib.caae_col_parts(R), R, decay)
            if a is not None:
                bn = self.bn[scope]
                self._c("caae_bn_act", R, fout, self._p(y), fout, self._p(bn["scale"]), self._p(bn["shift"]), 1,
                        self._p(a), fout)

    def _bn_bwd(self, scope, R, y, d_out, ldo, group, gscale, argmax, d_y):
        """training-mode BN + ReLU backward on [R, C]: d_out (per group row) -> d_y; bn/beta, bn/gamma grads."""
        C = self.scopes[scope][1]
        bn, v = self.bn[scope], self.v
        self._c("caae_bn_act_bwd_reduce", R, C, self._p(y), C, self._p(bn["scale"]), self._p(bn["shift"]),
                self._p(bn["mean"]), self._p(bn["invstd"]), self._p(d_out), ldo, group, float(gscale), 1,
                self._p(argmax), self._p(self.parts))
        self._c("caae_bn_bwd_finalize", C, self._p(self.parts), self.lib.caae_col_parts(R), float(R),
                self._p(v[f"{scope}/bn/gamma"]), self._p(bn["invstd"]), self._p(bn["coef"]),
                self._p(v.grad_of(f"{scope}/bn/gamma")), self._p(v.grad_of(f"{scope}/bn/beta")))
        self._c("caae_bn_act_bwd_apply", R, C, self._p(y), C, self._p(bn["scale"]), self._p(bn["shift"]),
                self._p(bn["mean"]), self._p(bn["invstd"]), self._p(bn["coef"]), self._p(d_out), ldo, group,
                float(gscale), 1, self._p(argmax), self._p(d_y), C)

    def _fc_wgrad(self, scope, x, x_lo, ldx, R, d_y, d_y_lo, bias_grad):
        fin, fout, _ = self.scopes[scope]
        self._fc_gemm(1, 0, fin, fout, R, x, x_lo, ldx, d_y, d_y_lo, fout, self.v.grad_of(f"{scope}/weights"), fout)
        if bias_grad:
            self._c("caae_colsum", R, fout, self._p(d_y), fout, self._p(self.v.grad_of(f"{scope}/biases")))

    def _dense_wgrad(self, scope, x, ldx, R, d_y, bias_grad):
        fin, fout, _ = self.scopes[scope]
        self._gemm(1, 0, fin, fout, R, x, ldx, d_y, fout, self.v.grad_of(f"{scope}/weights"), fout)
        if bias_grad:
            self._c("caae_colsum", R, fout, self._p(d_y), fout, self._p(self.v.grad_of(f"{scope}/biases")))

    # -- forward
    def forward(self, x, train_enc: bool, train_fc: bool, decay=None, want_before_embedding=False):
        """x f32[B,N,D] contiguous.  decay: 1-element device tensor (bn_decay) or None (-> 0.9)."""
        B, N, R, D, k = self.B, self.N, self.R, self.D, self.k
        assert x.shape == (B, N, D) and x.is_contiguous() and x.dtype == torch.float32
        decay = self.default_decay if decay is None else decay
        self.x0 = x
        self.trained = (train_enc, train_fc)
        before = None
        se = self.s_enc
        if se is not None: self._fork(se)
        with self._on(se):   # accumulation targets of the FC stack's split-K GEMMs (forward and backward)
            self._c("caae_fill_f32", self.fc_acc_flat.numel(), self._p(self.fc_acc_flat), 0.0)
            if self.fc_tc:       # low parts of every FC weight: one pass over the FC range of the flat parameter buffer
                n_fc = self.v.flat.numel() - self.fc_lo_start
                self._split(self.v.flat[self.fc_lo_start:], n_fc, 1, n_fc, self.flat_lo[self.fc_lo_start:], n_fc)
            if self.x3:          # low parts of the conv weights that take the split-precision tensor-core product
                for s_, lo in self.w_lo.items():
                    fin_, fout_, _ = self.scopes[s_]
                    self._split(self.v[f"{s_}/weights"], fout_, fin_, fout_, lo, fout_)
        if self.model == "dgcnn":
            feat, feat_lo, ldf, cknn = x, None, D, 3
            # clouds padded with repeats of their visible points (identical rows stay identical in every layer): flagged on
            # a side stream next to the first layer's search and used for routing from the second layer on
            sk0 = self.s_knn
            if sk0 is not None: self._fork(sk0)
            with self._on(sk0):
                self._c("caae_knn_classify", B, N, D, self._p(x), D, self._p(self.knn_flags))
            with self._on(se):   # the folded weights (and the low parts of the weights) depend on the parameters only
                for l in range(4):
                    self._c("caae_edge_fold_weights", self.cins[l], self.couts[l], self._p(self.v[f"dgcnn{l + 1}/weights"]),
                            self._p(self.v[f"dgcnn{l + 1}/biases"]), self._p(self.wf[l]), self._p(self.bf[l]), self.couts[l])
                    if self.x3 and l > 0:
                        self._split(self.wf[l], 2 * self.couts[l], self.cins[l], 2 * self.couts[l], self.wf_lo[l], 2 * self.couts[l])

            for l in range(4):
                scope = f"dgcnn{l + 1}"
                ci, co = self.cins[l], self.couts[l]
                # the projection [P|Q] = X Wf and the kNN search read the same features: run them side by side
                if se is not None: self._fork(se)
                with self._on(se):
                    self._gemm_fwd(R, 2 * co, ci, feat, feat_lo, ldf, self.wf[l], self.wf_lo[l], 2 * co, self.pq[l], 2 * co,
                                   self.bf[l])
                # heavily padded clouds (flagged once per forward, below) take the all-pairs kernel next to the
                # tensor-core kernel of the others: disjoint halves of idx[l]
                sk = self.s_knn
                if l == 0:
                    self._c("caae_knn", B, N, cknn, k, self._p(feat), ldf, self._p(self.idx[l]))
                    if sk is not None: self._join(sk)     # the flags are ready long before the second layer
                else:
                    if sk is not None: self._fork(sk)
                    with self._on(sk):
                        self._c("caae_knn_part", 2, self._p(self.knn_flags), B, N, cknn, k, self._p(feat), ldf, self._p(self.idx[l]))
                    self._c("caae_knn_part", 1, self._p(self.knn_flags), B, N, cknn, k, self._p(feat), ldf, self._p(self.idx[l]))
                    if sk is not None: self._join(sk)
                if se is not None: self._join(se)
                out = self.hcat[:, self.offs[l]:]
                feat_lo = self.hcat_lo[:, self.offs[l]:] if self.x3 else None   # written by the same kernel
                bn = self.bn[scope]
                nparts = self.lib.caae_edge_parts(B, N, k, co, 2 * co)
                rec = train_enc and self.edge_rec and nparts == B and self.fuse_finalize and k <= 255   # (the fused entry records)
                self.edge_recorded[l] = rec
                rec_args = (self._p(self.pos_cnt_e[:, self.offs[l]:]), self._p(self.pos_sum_e[:, self.offs[l]:]), 320) if rec \
                    else (None, None, 0)
                if train_enc:
                    self._c("caae_edge_stats", B, N, k, co, self._p(self.pq[l]), 2 * co, self._p(self.idx[l]),
                            self._p(self.parts))
                if train_enc and nparts == B and self.fuse_finalize:
                    # statistics -> coefficients inside the apply kernel (no finalize launch on the dependent chain)
                    v = self.v
                    self._c("caae_edge_apply_fused", B, N, k, co, self._p(self.pq[l]), 2 * co, self._p(self.idx[l]),
                            self._p(self.parts), nparts, float(R * k), self._p(v[f"{scope}/bn/gamma"]), self._p(v[f"{scope}/bn/beta"]),
                            self._p(v[f"{scope}/bn/ema_mean"]), self._p(v[f"{scope}/bn/ema_var"]), self._p(decay),
                            self._p(bn["scale"]), self._p(bn["shift"]), self._p(bn["mean"]), self._p(bn["invstd"]),
                            self._p(out), 320, self._p(feat_lo), *rec_args)
                else:
                    self._bn_coeffs(scope, train_enc, nparts, R * k, decay)
                    self._c("caae_edge_apply", B, N, k, co, self._p(self.pq[l]), 2 * co, self._p(self.idx[l]),
                            self._p(bn["scale"]), self._p(bn["shift"]), self._p(out), 320, self._p(feat_lo))
                feat, ldf, cknn = out, 320, co
            scope = "dgcnn_agg"
            bn = self.bn[scope]
            if not self._pooled_fwd(scope, self.hcat, self.hcat_lo if self.x3 else None, 320, train_enc, want_before_embedding, 1):
                self._dense_fwd(scope, self.hcat, 320, R, train_enc, decay, self.yagg, None,
                                x_lo=self.hcat_lo if self.x3 else None)
                self._c("caae_bn_act_pool", B, N, 1024, self._p(self.yagg), 1024, self._p(bn["scale"]),
                        self._p(bn["shift"]), 0, self._p(self.emb), None,
                        self._p(self.pos_cnt if train_enc else None), self._p(self.pos_sum if train_enc else None))
            if want_before_embedding:
                before = torch.empty(R, 1024, dtype=torch.float32, device=self.dev)
                self._c("caae_bn_act", R, 1024, self._p(self.yagg), 1024, self._p(bn["scale"]), self._p(bn["shift"]),
                        1, self._p(before), 1024)
        else:
            inp, ldi = x, D
            pooled = False
            for i, scope in enumerate(self.enc):
                last = i == len(self.enc) - 1
                if last and not train_enc and self.x3 and scope in self.w_lo:
                    if scope not in self.a_lo:
                        self.a_lo[scope] = torch.empty(R, ldi, dtype=torch.float32, device=self.dev)
                    self._split(inp, ldi, R, ldi, self.a_lo[scope], ldi)
                    pooled = self._pooled_fwd(scope, inp, self.a_lo[scope], ldi, train_enc, False, 2)
                if pooled:
                    break
                self._dense_fwd(scope, inp, ldi, R, train_enc, decay, self.enc_y[i], None if last else self.enc_a[i])
                if not last:
                    inp, ldi = self.enc_a[i], self.scopes[scope][1]
            if not pooled:
                bn = self.bn[self.enc[-1]]
                self._c("caae_bn_act_pool", B, N, 1024, self._p(self.enc_y[-1]), 1024, self._p(bn["scale"]),
                        self._p(bn["shift"]), 1, self._p(self.emb), self._p(self.argmax), None, None)
        if se is not None: self._join(se)
        self.forward_fc(train_fc, decay)
        outs = [self.fc_y[br[-1]] for br in self.branches]
        return outs[0], outs[1], outs[2], self.emb, before

    def forward_fc(self, train_fc: bool, decay):
        """embedding -> the three FC branches (decoder, rotation head, translation head)."""
        if self.fc_tc:
            self._split(self.emb, 1024, self.B, 1024, self.emb_lo, 1024)
        for bi in (1, 2, 0):   # heads on side streams, the decoder on the caller's
            st = self.s_branch[bi - 1] if (self.concurrent and bi > 0) else None
            if st is not None: self._fork(st)
            with self._on(st):
                inp, inp_lo = self.emb, (self.emb_lo if self.fc_tc else None)
                for s in self.branches[bi]:
                    inp, inp_lo = self._fc_fwd(s, inp, inp_lo, train_fc, decay)
        for st in self.s_branch:
            self._join(st)

    # -- backward (training-mode BN only, like the reference's training graph)
    def backward(self, d_recon, d_rot, d_trans, d_emb_extra=None):
        """Gradients of all trainable variables into ``variables.grad``.  d_* are f32 [B, fout]."""
        assert self.trained == (True, True), "backward is defined for training-mode batch norm"
        B, N, R, k = self.B, self.N, self.R, self.k
        if not self._heads_pending:
            self.backward_heads_async(d_rot, d_trans)
        self._heads_pending = False
        self._branch_backward(0, d_recon)
        for st in self.s_branch:
            self._join(st)
        self._c("caae_add3", B * 1024, self._p(self.d_emb_br[0]), self._p(self.d_emb_br[1]), self._p(self.d_emb_br[2]),
                self._p(self.d_emb))
        if d_emb_extra is not None:
            self.d_emb.add_(d_emb_extra)
        if self.after_fc_backward is not None:  # every FC / head gradient is final: start its allreduce
            self.after_fc_backward()
        self.backward_encoder()

    def backward_encoder(self):
        """d(embedding) in self.d_emb -> gradients of the encoder variables."""
        B, N, R, k = self.B, self.N, self.R, self.k
        if self.model == "dgcnn":
            scope = "dgcnn_agg"
            # mean-pool + ReLU + BN backward, in place over the pre-activation.  The two reductions over the 134 MB
            # pre-activation collapse to per-(cloud, channel) terms the forward pool pass recorded.
            bn, v = self.bn[scope], self.v
            self._c("caae_bn_pool_bwd_finalize", 1024, B, N, self._p(self.d_emb), 1024, 1.0 / N, self._p(self.pos_cnt),
                    self._p(self.pos_sum), self._p(bn["mean"]), self._p(bn["invstd"]), self._p(v[f"{scope}/bn/gamma"]),
                    self._p(bn["coef"]), self._p(v.grad_of(f"{scope}/bn/gamma")), self._p(v.grad_of(f"{scope}/bn/beta")))
            self._c("caae_bn_act_bwd_apply", R, 1024, self._p(self.yagg), 1024, self._p(bn["scale"]), self._p(bn["shift"]),
                    self._p(bn["mean"]), self._p(bn["invstd"]), self._p(bn["coef"]), self._p(self.d_emb), 1024, N,
                    1.0 / N, 1, None, self._p(self.yagg), 1024)
            se = self.s_enc   # weight gradients next to the data-gradient chain
            if se is not None: self._fork(se)
            with self._on(se):
                self._dense_wgrad(scope, self.hcat, 320, R, self.yagg, False)
                if self.after_last_conv_wgrad is not None:   # dgcnn_agg's gradients are final: start their allreduce
                    self.after_last_conv_wgrad()
            self._gemm(0, 1, R, 320, 1024, self.yagg, 1024, self.v[f"{scope}/weights"], 1024, self.d_hcat, 320)
            for l in (3, 2, 1, 0):
                scope = f"dgcnn{l + 1}"
                ci, co = self.cins[l], self.couts[l]
                bn = self.bn[scope]
                d_out = self.d_hcat[:, self.offs[l]:]
                args = (B, N, k, co, self._p(self.pq[l]), 2 * co, self._p(self.idx[l]), self._p(bn["scale"]),
                        self._p(bn["shift"]), self._p(bn["mean"]), self._p(bn["invstd"]))
                if self.edge_recorded[l]:
                    self._c("caae_edge_bwd_stats", B, N, k, co, 2 * co, self._p(d_out), 320,
                            self._p(self.pos_cnt_e[:, self.offs[l]:]), self._p(self.pos_sum_e[:, self.offs[l]:]), 320,
                            self._p(bn["invstd"]), self._p(self.parts))
                else:
                    self._c("caae_edge_bwd_reduce", *args, self._p(d_out), 320, self._p(self.parts))
                d_pq, d_wf = self.d_pq[l], self.d_wf[l]
                nparts = self.lib.caae_edge_parts(B, N, k, co, 2 * co)
                if nparts == B and self.fuse_finalize:
                    self._c("caae_edge_bwd_apply_fused", *args, self._p(self.parts), nparts, float(R * k),
                            self._p(self.v[f"{scope}/bn/gamma"]), self._p(bn["coef"]), self._p(self.v.grad_of(f"{scope}/bn/gamma")),
                            self._p(self.v.grad_of(f"{scope}/bn/beta")), self._p(d_out), 320, self._p(d_pq), 2 * co)
                else:
                    self._c("caae_bn_bwd_finalize", co, self._p(self.parts), nparts, float(R * k),
                            self._p(self.v[f"{scope}/bn/gamma"]), self._p(bn["invstd"]), self._p(bn["coef"]),
                            self._p(self.v.grad_of(f"{scope}/bn/gamma")), self._p(self.v.grad_of(f"{scope}/bn/beta")))
                    self._c("caae_edge_bwd_apply", *args, self._p(bn["coef"]), self._p(d_out), 320, self._p(d_pq), 2 * co)
                feat, ldf = (self.x0, self.D) if l == 0 else (self.hcat[:, self.offs[l - 1]:], 320)
                # dWf = X^T dPQ, then unfold to the reference's [2C, cout] weight
                if se is not None: self._fork(se)
                with self._on(se):
                    # the first layer's [24, 128] weight gradient is a thin contraction over all B*N rows: on the FFMA
                    # kernel its 128-row tiles are 80 % padding (35 us, step 1.891 -> 1.869 ms on tcgen05; CLOUDAAE_WGRAD_FLOOR=28 restores the FFMA route)
                    self._gemm(1, 0, ci, 2 * co, R, feat, ldf, d_pq, 2 * co, d_wf, 2 * co, floor=self.wgrad_floor)
                    self._c("caae_edge_unfold_wgrad", ci, co, self._p(d_wf), 2 * co,
                            self._p(self.v.grad_of(f"{scope}/weights")))
                if l > 0:  # d(net_{l-1}) += dPQ Wf^T, accumulated into its slice of d_hcat
                    self._gemm(0, 1, R, ci, 2 * co, d_pq, 2 * co, self.wf[l], 2 * co,
                               self.d_hcat[:, self.offs[l - 1]:], 320, None, 1)
            if se is not None: self._join(se)
        else:
            last = len(self.enc) - 1
            scope = self.enc[last]
            self._bn_bwd(scope, R, self.enc_y[last], self.d_emb, 1024, N, 1.0, self.argmax, self.enc_y[last])
            d_y = self.enc_y[last]
            for i in range(last, -1, -1):
                scope = self.enc[i]
                fin, fout, _ = self.scopes[scope]
                inp, ldi = (self.x0, self.D) if i == 0 else (self.enc_a[i - 1], fin)
                self._dense_wgrad(scope, inp, ldi, R, d_y, False)
                if i == last and self.after_last_conv_wgrad is not None:
                    self.after_last_conv_wgrad()
                if i > 0:
                    self._gemm(0, 1, R, fin, fout, d_y, fout, self.v[f"{scope}/weights"], fout, self.enc_d[i - 1], fin)
                    prev = self.enc[i - 1]
                    self._bn_bwd(prev, R, self.enc_y[i - 1], self.enc_d[i - 1], fin, 1, 1.0, None, self.enc_d[i - 1])
                    d_y = self.enc_d[i - 1]


# ---------------------------------------------------------------------------------------------
_DEFAULT_VARIABLES: dict = {}
_ENGINES: dict = {}


def get_variables(model: str, num_point: int = 256, point_dim: int = 24, device="cuda", seed: int = 0) -> Variables:
    """The default variable store of a model (TF's default graph collection), created on first use."""
    key = (model, num_point, point_dim, str(torch.device(device)))
    if key not in _DEFAULT_VARIABLES:
        layers = dgcnn_layers(num_point, point_dim) if model == "dgcnn" else pn_layers(num_point, point_dim)
        _DEFAULT_VARIABLES[key] = Variables(layers, device=device, seed=seed)
    return _DEFAULT_VARIABLES[key]


def reset_default_variables():
    _DEFAULT_VARIABLES.clear()
    _ENGINES.clear()


def _engine_for(variables: Variables, model: str, b: int, n: int, d: int, k: int) -> _Engine:
    key = (id(variables), model, b, n, d, k, os.environ.get("CLOUDAAE_GEMM", "tf32"), os.environ.get("CLOUDAAE_TF32X3", "1"))
    if key not in _ENGINES:
        _ENGINES[key] = _Engine(variables, model, b, n, d, k)
    return _ENGINES[key]


class _ModelFn(torch.autograd.Function):
    """Whole-network autograd node: backward fills ``variables.grad`` and hands it to autograd as the
    gradient of the flat parameter buffer."""

    @staticmethod
    def forward(ctx, point_cloud, flat, engine, train_enc, train_fc, decay, want_before):
        recon, rot, trans, emb, before = engine.forward(point_cloud.contiguous(), train_enc, train_fc, decay,
                                                        want_before)
        ctx.engine = engine
        # the activations live in the engine's shared workspace: stamp them, so that a backward pass that comes after
        # another forward (or a second backward) of the same shape fails loudly instead of using the wrong activations
        engine.generation = getattr(engine, "generation", 0) + 1
        ctx.generation = engine.generation
        ctx.mark_non_differentiable(*([before] if before is not None else []))
        outs = (recon.clone(), rot.clone(), trans.clone(), emb.clone())
        return outs + ((before,) if before is not None else (torch.empty(0, device=flat.device),))

    @staticmethod
    def backward(ctx, d_recon, d_rot, d_trans, d_emb, _d_before):
        e = ctx.engine
        if getattr(e, "generation", 0) != ctx.generation:
            raise RuntimeError("get_model backward: the engine workspace of this (variables, batch, num_point) shape was "
                               "overwritten by a later forward pass or already consumed by a backward pass; run "
                               "forward -> backward one call at a time per shape")
        e.generation += 1   # consumed: the backward pass normalises activations in place
        z = lambda g, ref: torch.zeros_like(ref) if g is None else g  # noqa: E731
        e.backward(z(d_recon, e.fc_y[e.branches[0][-1]]), z(d_rot, e.fc_y[e.branches[1][-1]]),
                   z(d_trans, e.fc_y[e.branches[2][-1]]), d_emb)
        return None, e.v.grad.clone(), None, None, None, None, None


def _decay_tensor(engine: _Engine, bn_decay):
    if bn_decay is None:
        return None
    if torch.is_tensor(bn_decay):
        return bn_decay.to(device=engine.dev, dtype=torch.float32).reshape(1)
    engine.decay_scalar.fill_(float(bn_decay))
    return engine.decay_scalar


def _check_cloud(point_cloud):
    if point_cloud.dim() != 3:
        raise _capi.InvalidArgumentError("point_cloud must be (batch_size, num_point, point_dim)")
    if point_cloud.dtype != torch.float32:
        raise _capi.InvalidArgumentError("point_cloud must be float32")
    _capi.require_cuda(point_cloud, "get_model")


def get_model_dgcnn_mean_6d(point_cloud, is_training_pl_encoder, is_training, k_neighbor, bn_decay=None,
                            variables: Variables | None = None):
    """DGCNN-style encoder (4 EdgeConv layers, kNN k=k_neighbor, mean aggregation, 320->1024 conv,
    mean pool), FC decoder to 4N points and two 3-vector pose heads.

    point_cloud f32[B,N,D] (xyz + class one-hot).  is_training_pl_encoder / is_training: batch-norm
    mode of the encoder convs / of the FC layers (training: batch statistics + EMA update with
    bn_decay, default 0.9; inference: EMA statistics).  Returns (net_recon [B,4N,3], net_rot [B,3],
    net_trans [B,3], end_points{'layer_before_embedding' [B,N,1,1024], 'embedding' [B,1024]}).
    """
    _check_cloud(point_cloud)
    b, n, d = point_cloud.shape
    v = variables if variables is not None else get_variables("dgcnn", n, d, point_cloud.device)
    eng = _engine_for(v, "dgcnn", b, n, d, int(k_neighbor))
    recon, rot, trans, emb, before = _ModelFn.apply(point_cloud, v.flat, eng, bool(is_training_pl_encoder),
                                                    bool(is_training), _decay_tensor(eng, bn_decay), True)
    end_points = {"layer_before_embedding": before.view(b, n, 1, 1024), "embedding": emb,
                  "nn_idx": [i.view(b, n, -1) for i in eng.idx]}
    return recon.view(b, n * 4, 3), rot, trans, end_points


def get_model_pn(point_cloud, is_training, bn_decay=None, variables: Variables | None = None):
    """PointNet variant: per-point MLP 64-64-64-128-1024 (BN + ReLU), max pool over points, the same
    FC decoder and pose heads.  Returns (net_recon, net_rot, net_trans, end_points{'embedding'})."""
    _check_cloud(point_cloud)
    b, n, d = point_cloud.shape
    v = variables if variables is not None else get_variables("pn", n, d, point_cloud.device)
    eng = _engine_for(v, "pn", b, n, d, 0)
    recon, rot, trans, emb, _ = _ModelFn.apply(point_cloud, v.flat, eng, bool(is_training), bool(is_training),
                                               _decay_tensor(eng, bn_decay), False)
    return recon.view(b, n * 4, 3), rot, trans, {"embedding": emb}
