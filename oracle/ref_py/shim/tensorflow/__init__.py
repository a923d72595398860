"""A minimal EAGER stand-in for the TensorFlow-1 API surface the reference's model / loss files use — TEST
INFRASTRUCTURE ONLY (oracle/ref_py executes /root/reference/models/pointnet_ycb_23_decoder_4.py, utils/tf_util.py and
losses/*.py IN PLACE through it, so oracle/model_ref.py can be pinned against the reference's own code).

Every op is its PyTorch equivalent on torch tensors (float64 in the tests).  Graph concepts collapse to eager ones:
variable scopes are a name stack, ``tf.get_variable`` looks the tensor up in the parameter dict the harness installed
(``install``), ``tf.cond`` evaluates its predicate, ``ExponentialMovingAverage`` keeps its shadows in the harness' dict.
Semantics follow TensorFlow 1.12: tf.nn.moments = biased variance; tf.nn.batch_normalization =
x * (rsqrt(var + eps) * gamma) + (beta - mean * rsqrt(var + eps) * gamma); tf.nn.top_k = sorted, ties to the lower
index; conv2d = NHWC, VALID / stride 1 only.
"""
import contextlib
import math as _math

import torch as _t

float16, float32, float64, int32, int64 = _t.float16, _t.float32, _t.float64, _t.int32, _t.int64
AUTO_REUSE = "AUTO_REUSE"


class _Dim(int):
    @property
    def value(self):
        return int(self)


class _Shape(list):
    def as_list(self):
        return [int(d) for d in self]


def _get_shape(self):
    return _Shape(_Dim(d) for d in self.shape)


_t.Tensor.get_shape = _get_shape      # tensors in this process answer .get_shape()[i].value like tf.Tensor

# ---- harness state ---------------------------------------------------------------------------------------------
_STATE = {"params": None, "ema": None, "ema_updates": None, "scopes": [], "dtype": _t.float64, "top_k_override": None,
          "created": []}


def install(params, ema=None, ema_updates=None, dtype=_t.float64, top_k_override=None):
    """params: {'dgcnn1/weights': tensor, ...}; ema: {'dgcnn1/bn/ema_mean': ...}; ema_updates: dict that receives the
    moving-average updates; top_k_override: optional iterator of index tensors that replaces successive tf.nn.top_k
    results (takes neighbour near-ties out of a floating-point comparison)."""
    _STATE.update(params=params, ema=ema if ema is not None else params, ema_updates=ema_updates, scopes=[], dtype=dtype,
                  top_k_override=iter(top_k_override) if top_k_override is not None else None, created=[])


def _scope_name():
    return "/".join(s for s in _STATE["scopes"] if s)


class _Scope:
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None, **_):
    if isinstance(name_or_scope, _Scope):          # re-entering the current scope (tf.get_variable_scope())
        yield name_or_scope
        return
    _STATE["scopes"].append(str(name_or_scope))
    try:
        yield _Scope(_scope_name())
    finally:
        _STATE["scopes"].pop()


def get_variable_scope():
    return _Scope(_scope_name())


@contextlib.contextmanager
def device(_):
    yield


@contextlib.contextmanager
def control_dependencies(_):
    yield


def get_variable(name=None, shape=None, initializer=None, dtype=None, trainable=None, **_):
    full = (_scope_name() + "/" + name) if _scope_name() else name
    _STATE["created"].append(full)
    p = _STATE["params"]
    if full not in p:
        raise KeyError(f"reference graph asked for variable {full!r}, which the harness did not provide")
    v = p[full]
    if shape is None and isinstance(initializer, _t.Tensor):
        shape = list(initializer.shape)
    if shape is not None and list(v.shape) != [int(s) for s in shape]:
        if v.numel() != int(_math.prod(int(s) for s in shape)):
            raise ValueError(f"{full}: harness tensor {tuple(v.shape)} cannot take the reference's shape {shape}")
        v = v.reshape([int(s) for s in shape])     # [fan_in, cout] -> [kh, kw, cin, cout] (row-major, same order)
    return v


def Variable(value, **_):
    return _t.as_tensor(value, dtype=_STATE["dtype"])


def placeholder(dtype, shape=None, **_):
    raise RuntimeError("the eager stand-in has no placeholders: pass tensors")


def constant(value, dtype=None, shape=None, **_):
    t = _t.as_tensor(value, dtype=dtype if dtype is not None else _STATE["dtype"])
    return t.expand(list(shape)).clone() if shape is not None else t


def constant_initializer(value=0.0, **_):
    return ("constant", value)


def truncated_normal_initializer(stddev=1.0, **_):
    return ("truncated_normal", stddev)


class _Namespace:
    pass


contrib = _Namespace()
contrib.layers = _Namespace()
contrib.layers.xavier_initializer = lambda **_: ("xavier",)
summary = _Namespace()
summary.histogram = lambda *a, **k: None
summary.scalar = lambda *a, **k: None


def add_to_collection(*_a, **_k):
    return None


def no_op(*_a, **_k):
    return None


def identity(x, **_):
    return x


def cond(pred, true_fn, false_fn, **_):
    return true_fn() if bool(pred) else false_fn()


# ---- array ops -------------------------------------------------------------------------------------------------
def shape(x, **_):
    return [int(d) for d in x.shape]


def reshape(x, shp, **_):
    return x.reshape([int(s) for s in shp])


def expand_dims(x, axis=None, dim=None, **_):
    return x.unsqueeze(axis if axis is not None else dim)


def squeeze(x, axis=None, **_):
    return x.squeeze() if axis is None else x.squeeze(axis)


def transpose(x, perm=None, **_):
    return x.permute(*perm) if perm is not None else x.t()


def matrix_transpose(x, **_):
    return x.transpose(-1, -2)


def concat(values, axis, **_):
    return _t.cat(list(values), dim=axis)


def tile(x, multiples, **_):
    return x.repeat(*[int(m) for m in multiples])


def range(*args, **_):       # noqa: A001 (tf.range)
    return _t.arange(*[int(a) for a in args])


def gather(params, indices, **_):
    return params[indices.long()]


def zeros(shp, dtype=None, **_):
    return _t.zeros([int(s) for s in shp], dtype=dtype if dtype is not None else _STATE["dtype"])


def eye(n, batch_shape=None, dtype=None, **_):
    e = _t.eye(int(n), dtype=dtype if dtype is not None else _STATE["dtype"])
    return e.expand(*[int(b) for b in batch_shape], int(n), int(n)).clone() if batch_shape is not None else e


def where(c, a, b, **_):
    return _t.where(c, a, b)


def less(a, b, **_):
    return a < b


def argmax(x, axis=None, **_):
    return _t.argmax(x, dim=axis)


def cast(x, dtype, **_):
    return x.to(dtype)


def random_normal(shp, mean=0.0, stddev=1.0, dtype=None, **_):
    return _t.randn([int(s) for s in shp], dtype=dtype if dtype is not None else _STATE["dtype"]) * stddev + mean


# ---- math ------------------------------------------------------------------------------------------------------
def matmul(a, b, **_):
    return _t.matmul(a, b)


def multiply(a, b, **_):
    return a * b


def square(x, **_):
    return x * x


sqrt, sin, cos, acos = (lambda x, **_: _t.sqrt(x)), (lambda x, **_: _t.sin(x)), (lambda x, **_: _t.cos(x)), \
    (lambda x, **_: _t.acos(x))


def clip_by_value(x, lo, hi, **_):
    return _t.clamp(x, lo, hi)


def trace(x, **_):
    return x.diagonal(dim1=-2, dim2=-1).sum(-1)


def _reduce(fn):
    def op(x, axis=None, keep_dims=False, keepdims=None, **_):
        kd = keep_dims if keepdims is None else keepdims
        if axis is None:
            return fn(x)
        return fn(x, dim=axis, keepdim=kd)
    return op


reduce_sum = _reduce(_t.sum)
reduce_mean = _reduce(_t.mean)


def reduce_max(x, axis=None, keep_dims=False, keepdims=None, **_):
    kd = keep_dims if keepdims is None else keepdims
    return x.max() if axis is None else x.max(dim=axis, keepdim=kd).values


# ---- tf.nn -----------------------------------------------------------------------------------------------------
nn = _Namespace()
nn.relu = lambda x, **_: _t.relu(x)
nn.bias_add = lambda x, b, **_: x + b
nn.l2_loss = lambda x, **_: (x * x).sum() / 2


def _moments(x, axes, name=None, keep_dims=False, **_):
    mean = x.mean(dim=list(axes), keepdim=keep_dims)
    var = ((x - x.mean(dim=list(axes), keepdim=True)) ** 2).mean(dim=list(axes), keepdim=keep_dims)   # biased
    mean._moment_kind, var._moment_kind = "mean", "var"
    return mean, var


def _batch_normalization(x, mean, variance, offset, scale, variance_epsilon, **_):
    inv = _t.rsqrt(variance + variance_epsilon)
    if scale is not None:
        inv = inv * scale
    return x * inv + ((offset - mean * inv) if offset is not None else (-mean * inv))


def _conv2d(x, kernel, strides, padding, **_):
    if str(padding) != "VALID" or list(strides) != [1, 1, 1, 1]:
        raise NotImplementedError("stand-in conv2d: VALID padding, stride 1 (all the reference's hot path uses)")
    # NHWC x [kh, kw, cin, cout] -> NCHW conv -> NHWC
    y = _t.nn.functional.conv2d(x.permute(0, 3, 1, 2), kernel.permute(3, 2, 0, 1))
    return y.permute(0, 2, 3, 1)


def _top_k(x, k=1, sorted=True, **_):   # noqa: A002
    ov = _STATE["top_k_override"]
    idx = _t.sort(x, dim=-1, descending=True, stable=True).indices[..., :k]   # ties -> lower index (TopKV2)
    if ov is not None:
        idx = next(ov).to(idx.device).long()
    return _t.gather(x, -1, idx), idx


def _max_pool(x, ksize, strides, padding, **_):
    if str(padding) != "VALID":
        raise NotImplementedError
    y = _t.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), kernel_size=(ksize[1], ksize[2]), stride=(strides[1], strides[2]))
    return y.permute(0, 2, 3, 1)


nn.moments, nn.batch_normalization, nn.conv2d, nn.top_k, nn.max_pool = _moments, _batch_normalization, _conv2d, _top_k, _max_pool
nn.dropout = lambda x, keep_prob=None, **_: x


# ---- tf.train --------------------------------------------------------------------------------------------------
train = _Namespace()


class _EMA:
    """tf.train.ExponentialMovingAverage of batch-norm moments: shadow = decay * shadow + (1 - decay) * value, shadows
    start at 0 (tensor averages are not zero-debiased).  Shadows live in the harness' dict under
    '<scope>/ema_mean' / '<scope>/ema_var' (the scope is the layer's 'bn' scope)."""

    def __init__(self, decay, **_):
        self.decay = decay
        self.scope = _scope_name()

    def _key(self, t):
        return f"{self.scope}/ema_{t._moment_kind}"

    def apply(self, tensors):
        upd = _STATE["ema_updates"]
        for t in tensors:
            k = self._key(t)
            if upd is not None:
                d = float(self.decay)
                upd[k] = d * _STATE["ema"][k] + (1 - d) * t.detach()
        return None

    def average(self, t):
        return _STATE["ema"][self._key(t)]


train.ExponentialMovingAverage = _EMA


# ---- ops of the synthesis utilities (utils/hidden_point_removal.py, generate_occluder.py, sample_pose_in_frustum.py) ----
# These run in float32 in the reference's tf.data map chain; the stand-in keeps the dtype of its inputs.
def norm(x, ord="euclidean", axis=None, **_):   # noqa: A002
    return _t.sqrt((x * x).sum(dim=axis))


def stack(values, axis=0, **_):
    return _t.stack(list(values), dim=axis)


def zeros_like(x, **_):
    return _t.zeros_like(x)


def convert_to_tensor(x, **_):
    return _t.as_tensor(x)


def py_func(func, inp, Tout, **_):
    """tf.py_func: run the Python function on NumPy copies of the inputs (hidden_point_removal.py:49)."""
    res = func(*[t.detach().cpu().numpy() for t in inp])
    return tuple(_t.as_tensor(r) for r in res)


math = _Namespace()
math.reduce_max = reduce_max
math.pow = lambda a, b, **_: _t.pow(_t.as_tensor(a, dtype=b.dtype if _t.is_tensor(b) else _t.float32), b)
math.divide = lambda a, b, **_: a / b
math.tan = lambda x, **_: _t.tan(x)
linalg = _Namespace()
linalg.cross = lambda a, b, **_: _t.linalg.cross(a, b)
dtypes = _Namespace()
dtypes.cast = cast

random = _Namespace()
_DRAWS = {"queue": None}


def script_normal_draws(draws):
    """The next tf.random.normal calls return mean + stddev * z with z taken from `draws` in call order (TensorFlow's
    generator cannot be reproduced; every random draw of the parity boundary is an explicit input)."""
    _DRAWS["queue"] = iter(draws)


def _random_normal(shp, mean=0.0, stddev=1.0, dtype=None, **_):
    z = next(_DRAWS["queue"]).reshape([int(s) for s in shp])
    return _t.as_tensor(mean, dtype=z.dtype) + _t.as_tensor(stddev, dtype=z.dtype) * z


random.normal = _random_normal
