"""Plain-PyTorch restatement of the reference's model, losses and optimiser — TEST INFRASTRUCTURE ONLY.

Follows the TensorFlow graph literally (no algebraic refactoring), in whatever dtype the inputs
carry (tests use float64 on CPU for a tight bound, float32 for "what TF computes"):
  models/pointnet_ycb_23_decoder_4.py:23-89 (get_model_pn), :327-455 (get_model_dgcnn_mean_6d)
  utils/tf_util.py:111-179 (conv2d), :321-365 (fully_connected), :473-511 (batch_norm_template),
                   :597-618 (pairwise_xyz_distance), :621-632 (knn), :635-669 (get_edge_feature)
  losses/chamfer_loss.py, losses/trans_distance.py, losses/angular_distance_taylor.py
  train_cloudAAE_ycbv.py:196-273 (bn_decay schedule, input prep, total loss, Adam)
PINNED against the reference's own code: oracle/ref_py executes /root/reference/models/pointnet_ycb_23_decoder_4.py,
utils/tf_util.py and losses/*.py in place (eager TensorFlow stand-in over torch) and
tests/test_ref_py_pins_model_oracle.py requires this file to reproduce their outputs, losses, moving-average updates
and the gradient of every trainable variable to 1e-10 — plus the vectors that run committed under
tests/golden/ref_py_golden.npz for boxes without /root/reference.  (TensorFlow 1.12 itself is not installable here, so
the arithmetic of TF's kernels is represented by float64 torch ops.)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

BN_EPS = 1e-3  # tf.nn.batch_normalization(..., 1e-3), tf_util.py:510

# (scope, fan_in, fan_out, has_bn) in creation order — get_model_dgcnn_mean_6d
DGCNN_LAYERS = [
    ("dgcnn1", 48, 64, True), ("dgcnn2", 128, 64, True), ("dgcnn3", 128, 64, True), ("dgcnn4", 128, 128, True),
    ("dgcnn_agg", 320, 1024, True), ("dgcnn_fc1", 1024, 1024, True), ("dgcnn_fc2", 1024, 1024, True),
    ("dgcnn_output", 1024, 3072, False),
    ("dgcnn_rot_fc1", 1024, 512, True), ("dgcnn_rot_fc2", 512, 256, True), ("dgcnn_output_rot", 256, 3, False),
    ("dgcnn_trans_fc1", 1024, 512, True), ("dgcnn_trans_fc2", 512, 256, True), ("dgcnn_output_trans", 256, 3, False),
]


def pn_layers(point_dim: int, num_point: int):
    return [
        ("pn_conv1_encoder", point_dim, 64, True), ("pn_conv2_encoder", 64, 64, True),
        ("pn_conv3_encoder", 64, 64, True), ("pn_conv4_encoder", 64, 128, True), ("pn_conv5_encoder", 128, 1024, True),
        ("pn_fc1_decoder", 1024, 1024, True), ("pn_fc2_decoder", 1024, 1024, True),
        ("pn_output", 1024, num_point * 12, False),
        ("pn_rot_fc1", 1024, 512, True), ("pn_rot_fc2", 512, 256, True), ("pn_output_rot", 256, 3, False),
        ("pn_trans_fc1", 1024, 512, True), ("pn_trans_fc2", 512, 256, True), ("pn_output_trans", 256, 3, False),
    ]


def init_params(layers, seed: int = 1234, dtype=torch.float32, perturb: bool = True):
    """Xavier-uniform weights (tf_util.py:42-43), zero biases, gamma 1 / beta 0; EMA mean 0 / var 0
    as TF initialises them.  With perturb=True the biases, gamma/beta and EMA statistics get seeded
    noise so parity tests exercise every term (the trained checkpoint blob is not shipped)."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for scope, fin, fout, has_bn in layers:
        limit = math.sqrt(6.0 / (fin + fout))
        p[f"{scope}/weights"] = ((torch.rand(fin, fout, generator=g, dtype=torch.float64) * 2 - 1) * limit).to(dtype)
        p[f"{scope}/biases"] = torch.zeros(fout, dtype=dtype)
        if perturb:
            p[f"{scope}/biases"] = (torch.randn(fout, generator=g, dtype=torch.float64) * 0.05).to(dtype)
        if has_bn:
            p[f"{scope}/bn/gamma"] = torch.ones(fout, dtype=dtype)
            p[f"{scope}/bn/beta"] = torch.zeros(fout, dtype=dtype)
            p[f"{scope}/bn/ema_mean"] = torch.zeros(fout, dtype=dtype)
            p[f"{scope}/bn/ema_var"] = torch.zeros(fout, dtype=dtype)
            if perturb:
                p[f"{scope}/bn/gamma"] = (1 + 0.1 * torch.randn(fout, generator=g, dtype=torch.float64)).to(dtype)
                p[f"{scope}/bn/beta"] = (0.1 * torch.randn(fout, generator=g, dtype=torch.float64)).to(dtype)
                p[f"{scope}/bn/ema_mean"] = (0.1 * torch.randn(fout, generator=g, dtype=torch.float64)).to(dtype)
                p[f"{scope}/bn/ema_var"] = (0.5 + torch.rand(fout, generator=g, dtype=torch.float64)).to(dtype)
    return p


def trainable_names(params):
    return [k for k in params if not k.endswith(("ema_mean", "ema_var"))]


# ------------------------------------------------------------------ tf_util restatements
def batch_norm_template(x, params, scope, moments_dims, is_training, bn_decay, ema_updates):
    """tf_util.py:473-511.  Training: batch mean / biased variance, EMA update
    shadow = decay*shadow + (1-decay)*batch (decay defaults to 0.9).  Eval: EMA statistics."""
    gamma, beta = params[f"{scope}/bn/gamma"], params[f"{scope}/bn/beta"]
    if is_training:
        mean = x.mean(dim=moments_dims)
        var = ((x - mean) ** 2).mean(dim=moments_dims)
        decay = 0.9 if bn_decay is None else bn_decay
        if ema_updates is not None:
            ema_updates[f"{scope}/bn/ema_mean"] = decay * params[f"{scope}/bn/ema_mean"] + (1 - decay) * mean.detach()
            ema_updates[f"{scope}/bn/ema_var"] = decay * params[f"{scope}/bn/ema_var"] + (1 - decay) * var.detach()
    else:
        mean, var = params[f"{scope}/bn/ema_mean"], params[f"{scope}/bn/ema_var"]
    inv = torch.rsqrt(var + BN_EPS) * gamma  # tf.nn.batch_normalization
    return x * inv + (beta - mean * inv)


def conv2d_1x1(x, params, scope, is_training, bn_decay, ema_updates, bn=True, relu=True):
    """tf_util.conv2d with a [1,1] kernel on BxHxWxC (tf_util.py:111-179): matmul + bias + BN + ReLU."""
    y = torch.matmul(x, params[f"{scope}/weights"]) + params[f"{scope}/biases"]
    if bn:
        y = batch_norm_template(y, params, scope, [0, 1, 2], is_training, bn_decay, ema_updates)
    return torch.relu(y) if relu else y


def fully_connected(x, params, scope, is_training=None, bn_decay=None, ema_updates=None, bn=False, relu=True):
    """tf_util.fully_connected (tf_util.py:321-365)."""
    y = torch.matmul(x, params[f"{scope}/weights"]) + params[f"{scope}/biases"]
    if bn:
        y = batch_norm_template(y, params, scope, [0], is_training, bn_decay, ema_updates)
    return torch.relu(y) if relu else y


def pairwise_xyz_distance(pc):
    """tf_util.py:597-618.  For a [B,N,C] input the slice [:,:,0:3] keeps xyz; for the [B,N,1,C]
    feature tensors of layers 2-4 the same slice hits the size-1 axis, so ALL C channels are used."""
    if pc.dim() == 3:
        pc = pc[:, :, 0:3]
    else:
        pc = pc[:, :, 0:3].squeeze(2)
    inner = -2 * torch.matmul(pc, pc.transpose(1, 2))
    sq = (pc ** 2).sum(-1, keepdim=True)
    return sq + inner + sq.transpose(1, 2)


def knn(adj, k):
    """tf_util.py:621-632: top_k(-adj, k) indices (sorted; ties -> lower index, as TopKV2)."""
    neg = -adj
    # torch.topk does not promise index order on ties; a stable sort does
    return torch.sort(neg, dim=-1, descending=True, stable=True).indices[..., :k]


def get_edge_feature(pc, nn_idx):
    """tf_util.py:635-669: concat(x_i tiled k, x_nn - x_i) -> [B,N,k,2C]."""
    if pc.dim() == 4:
        pc = pc.squeeze(2)
    b, n, c = pc.shape
    k = nn_idx.shape[-1]
    flat = pc.reshape(b * n, c)
    base = (torch.arange(b, device=pc.device) * n).view(b, 1, 1)
    nb = flat[(nn_idx + base).reshape(-1)].view(b, n, k, c)
    central = pc.unsqueeze(2).expand(b, n, k, c)
    return torch.cat([central, nb - central], dim=-1)


# ------------------------------------------------------------------ models
def get_model_dgcnn_mean_6d(point_cloud, params, is_training_pl_encoder, is_training, k_neighbor, bn_decay=None,
                            ema_updates=None, nn_idx_override=None):
    """models/pointnet_ycb_23_decoder_4.py:327-455.  nn_idx_override: optional list of four
    [B,N,k] index tensors replacing the kNN result (used to take neighbour near-ties out of a
    floating-point comparison)."""
    b, n, _ = point_cloud.shape
    end_points = {}
    nets = []
    net = point_cloud
    used_idx = []
    for li, scope in enumerate(("dgcnn1", "dgcnn2", "dgcnn3", "dgcnn4")):
        adj = pairwise_xyz_distance(net)
        idx = knn(adj, k_neighbor) if nn_idx_override is None else nn_idx_override[li]
        used_idx.append(idx)
        edge = get_edge_feature(net, idx)
        net = conv2d_1x1(edge, params, scope, is_training_pl_encoder, bn_decay, ema_updates)
        net = net.mean(dim=-2, keepdim=True)
        nets.append(net)
    net = conv2d_1x1(torch.cat(nets, dim=-1), params, "dgcnn_agg", is_training_pl_encoder, bn_decay, ema_updates)
    end_points["layer_before_embedding"] = net
    net = net.mean(dim=1, keepdim=True).reshape(b, -1)
    end_points["embedding"] = net
    end_points["nn_idx"] = used_idx
    emb = net
    net = fully_connected(emb, params, "dgcnn_fc1", is_training, bn_decay, ema_updates, bn=True)
    net = fully_connected(net, params, "dgcnn_fc2", is_training, bn_decay, ema_updates, bn=True)
    net = fully_connected(net, params, "dgcnn_output", relu=False)
    net_recon = net.reshape(b, n * 4, 3)
    r = fully_connected(emb, params, "dgcnn_rot_fc1", is_training, bn_decay, ema_updates, bn=True)
    r = fully_connected(r, params, "dgcnn_rot_fc2", is_training, bn_decay, ema_updates, bn=True)
    net_rot = fully_connected(r, params, "dgcnn_output_rot", relu=False)
    t = fully_connected(emb, params, "dgcnn_trans_fc1", is_training, bn_decay, ema_updates, bn=True)
    t = fully_connected(t, params, "dgcnn_trans_fc2", is_training, bn_decay, ema_updates, bn=True)
    net_trans = fully_connected(t, params, "dgcnn_output_trans", relu=False)
    return net_recon, net_rot, net_trans, end_points


def get_model_pn(point_cloud, params, is_training, bn_decay=None, ema_updates=None, argmax_override=None):
    """models/pointnet_ycb_23_decoder_4.py:23-89.  conv1 has kernel [1,point_dim] over the input
    expanded to [B,N,D,1], i.e. a per-point D->64 linear map; max-pool over the N points."""
    b, n, _ = point_cloud.shape
    end_points = {}
    net = point_cloud.unsqueeze(2)  # [B,N,1,D] — equivalent view of the [1,D] VALID convolution
    for scope in ("pn_conv1_encoder", "pn_conv2_encoder", "pn_conv3_encoder", "pn_conv4_encoder", "pn_conv5_encoder"):
        net = conv2d_1x1(net, params, scope, is_training, bn_decay, ema_updates)
    pre_pool = net.reshape(b, n, -1)
    if argmax_override is None:
        net = pre_pool.max(dim=1).values
    else:  # route through the given rows (takes max-pool near-ties out of a floating-point comparison)
        net = torch.gather(pre_pool, 1, argmax_override.view(b, 1, -1)).reshape(b, -1)
    end_points["pre_pool_max"] = pre_pool.max(dim=1).values
    end_points["embedding"] = net
    emb = net
    net = fully_connected(emb, params, "pn_fc1_decoder", is_training, bn_decay, ema_updates, bn=True)
    net = fully_connected(net, params, "pn_fc2_decoder", is_training, bn_decay, ema_updates, bn=True)
    net = fully_connected(net, params, "pn_output", relu=False)
    net_recon = net.reshape(b, n * 4, 3)
    r = fully_connected(emb, params, "pn_rot_fc1", is_training, bn_decay, ema_updates, bn=True)
    r = fully_connected(r, params, "pn_rot_fc2", is_training, bn_decay, ema_updates, bn=True)
    net_rot = fully_connected(r, params, "pn_output_rot", relu=False)
    t = fully_connected(emb, params, "pn_trans_fc1", is_training, bn_decay, ema_updates, bn=True)
    t = fully_connected(t, params, "pn_trans_fc2", is_training, bn_decay, ema_updates, bn=True)
    net_trans = fully_connected(t, params, "pn_output_trans", relu=False)
    return net_recon, net_rot, net_trans, end_points


# ------------------------------------------------------------------ losses
def chamfer_get_loss(pred, label, chunk: int = 8):
    """losses/chamfer_loss.py:8-14 with a brute-force nn_distance (squared distances, first argmin).
    The [b,n,m,3] difference tensor is built `chunk` clouds at a time (3.2 GB in float64 at b=128, n=m=1024)."""
    d1, d2 = [], []
    for s in range(0, pred.shape[0], chunk):
        d = ((pred[s:s + chunk].unsqueeze(2) - label[s:s + chunk].unsqueeze(1)) ** 2).sum(-1)
        d1.append(d.min(dim=2).values); d2.append(d.min(dim=1).values)
    per = torch.cat(d1) + torch.cat(d2)
    return per.mean(), per


def get_translation_error(pred, label):
    """losses/trans_distance.py:4-9."""
    per = torch.sqrt(((label - pred) ** 2).sum(dim=1))
    return per.mean(), per


def skew_symmetric(a):
    z = torch.zeros_like(a[:, 0])
    return torch.stack([torch.stack([z, -a[:, 2], a[:, 1]], 1), torch.stack([a[:, 2], z, -a[:, 0]], 1),
                        torch.stack([-a[:, 1], a[:, 0], z], 1)], 1)


def exponential_map(axag, eps=1e-2):
    """losses/angular_distance_taylor.py:30-66 (float64 in the reference)."""
    ss = skew_symmetric(axag)
    theta_sq = (axag ** 2).sum(dim=1)
    small = theta_sq < eps
    safe = torch.where(small, torch.ones_like(theta_sq), theta_sq)
    theta = torch.sqrt(safe)
    t4, t6, t8 = theta_sq ** 2, theta_sq ** 3, theta_sq ** 4
    term1 = torch.where(small, 1 - theta_sq / 6 + t4 / 120 - t6 / 5040 + t8 / 362880, torch.sin(theta) / theta)
    term2 = torch.where(small, 0.5 - theta_sq / 24 + t4 / 720 - t6 / 40320 + t8 / 3628800,
                        (1 - torch.cos(theta)) / safe)
    eye = torch.eye(3, dtype=axag.dtype).unsqueeze(0)
    return eye + term1[:, None, None] * ss + term2[:, None, None] * torch.matmul(ss, ss)


def get_rotation_error(pred, label):
    """losses/angular_distance_taylor.py:102-116: theta = acos(clip((tr(R_l R_p^T) - 1)/2, +-0.9999999))."""
    r = torch.matmul(exponential_map(label), exponential_map(pred).transpose(1, 2))
    tr = (r.diagonal(dim1=1, dim2=2).sum(-1) - 1) / 2
    theta = torch.acos(torch.clamp(tr, -0.9999999, 0.9999999))
    return theta.mean(), theta


def bn_decay_schedule(step: int, batch_size: int) -> float:
    """train_cloudAAE_ycbv.py:166-169,196-202: min(0.99, 1 - 0.5*0.5^floor(step*B/40))."""
    return min(0.99, 1.0 - 0.5 * (0.5 ** math.floor(step * batch_size / 40.0)))


def prepare_input(visible, class_id, noise, num_point=256, num_class=21):
    """train_cloudAAE_ycbv.py:206-226: slice, add noise, subtract the per-cloud mean, append one-hot."""
    v = visible[:, :num_point, :] + noise
    mean = v.mean(dim=1)
    onehot = torch.nn.functional.one_hot(class_id.long(), num_class).to(v.dtype)
    x = torch.cat([v - mean.unsqueeze(1), onehot.unsqueeze(1).expand(-1, num_point, -1)], dim=2)
    return x, mean


def train_losses(params, x, mean, target, translation, axisangle, bn_decay, k=10, ema_updates=None,
                 nn_idx_override=None, model="dgcnn", argmax_override=None):
    """train_cloudAAE_ycbv.py:228-268: model, chamfer on recon+mean, translation L2, rotation geodesic
    (float64), total = 1000*chamfer + 10*trans + rot."""
    if model == "dgcnn":
        recon, rot, trans_res, ep = get_model_dgcnn_mean_6d(x, params, True, True, k, bn_decay, ema_updates,
                                                            nn_idx_override)
    else:
        recon, rot, trans_res, ep = get_model_pn(x, params, True, bn_decay, ema_updates, argmax_override)
    xyz_recon = recon + mean.unsqueeze(1)
    trans_pred = trans_res + mean
    xyz_loss, _ = chamfer_get_loss(xyz_recon, target)
    trans_loss, _ = get_translation_error(trans_pred, translation)
    axag_loss, _ = get_rotation_error(rot.double(), axisangle.double())
    axag_loss = axag_loss.to(x.dtype)
    total = 1000 * xyz_loss + 10 * trans_loss + axag_loss
    return total, {"chamfer": xyz_loss, "trans": trans_loss, "rot": axag_loss, "recon": xyz_recon, "rot_pred": rot,
                   "trans_pred": trans_pred, "end_points": ep}


def adam_step(param, grad, m, v, step, lr=0.0008, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    return param - lr_t * m / (torch.sqrt(v) + eps), m, v
