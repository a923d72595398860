"""TensorFlow-1 checkpoint (TensorBundle) reader — `tf.train.Saver` files without TensorFlow.

The reference saves and restores its network with ``tf.train.Saver`` (train_cloudAAE_ycbv.py:276, 423-430;
evaluate_cloudAAE_ycbv.py:495-499): ``model.ckpt.index`` + ``model.ckpt.data-00000-of-00001``.  This module
reads that pair and maps the reference's variable scopes onto a
:class:`cloudaae_b200.models.pointnet_ycb_23_decoder_4.Variables` store, so a trained reference
checkpoint drives `get_model_dgcnn_mean_6d` / `get_model_pn` here.  A minimal writer produces the same
format (used by the tests and to hand weights back to a TensorFlow reader).

On-disk format of ``.index`` (a LevelDB-style sorted table):
  footer (last 48 bytes): varint64 (offset, size) of the metaindex block and of the index block,
      zero padding, 8-byte magic 0xdb4775248b80fb57 (little endian);
  block: prefix-compressed entries  varint shared | varint non_shared | varint value_len | key suffix |
      value,  then uint32 restart offsets and uint32 num_restarts;  1-byte compression tag + 4-byte masked
      crc32c follow the block (not counted in its handle's size);
  index block: values are (offset, size) handles of the data blocks;
  data block entry: key = variable name ("" = BundleHeaderProto), value = BundleEntryProto
      {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset, 5: size, 6: crc32c}.
Tensor bytes live at `offset` in ``<prefix>.data-<shard:05d>-of-<num_shards:05d>``, little endian.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

_MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointFormatError(ValueError):
    pass


# ---- primitives -------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> tuple[int, int]:
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointFormatError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointFormatError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf: bytes):
    """Yield (field number, wire type, value) of one protobuf message (varint / fixed / length-delimited)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise CheckpointFormatError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli) through the native library's host routine (caae_crc32c)."""
    from .. import _capi
    return int(_capi.lib().caae_crc32c(0, data, len(data)))


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- table reading ------------------------------------------------------------------------------
def _block_entries(buf: bytes, offset: int, size: int):
    block = buf[offset:offset + size]
    if len(block) != size or size < 4:
        raise CheckpointFormatError("block handle outside the file")
    if offset + size < len(buf) and buf[offset + size] != 0:
        raise CheckpointFormatError("compressed table blocks are not supported (TensorFlow writes them raw)")
    num_restarts = struct.unpack_from("<I", block, size - 4)[0]
    end = size - 4 - 4 * num_restarts
    if end < 0:
        raise CheckpointFormatError("corrupt restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _parse_entry(value: bytes) -> dict:
    e = {"dtype": None, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    for field, wt, v in _proto_fields(value):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            dims = []
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
    return e


def read_index(index_path: str) -> tuple[dict, "OrderedDict[str, dict]"]:
    """Parse ``<prefix>.index``.  Returns (header {num_shards, ...}, name -> entry) in file (= sorted) order."""
    buf = open(index_path, "rb").read()
    if len(buf) < 48:
        raise CheckpointFormatError("file shorter than a table footer")
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != _MAGIC:
        raise CheckpointFormatError("bad table magic: not a TensorBundle index")
    pos = 0
    _, pos = _varint(footer, pos); _, pos = _varint(footer, pos)          # metaindex handle (unused)
    idx_off, pos = _varint(footer, pos); idx_size, pos = _varint(footer, pos)
    header, entries = {"num_shards": 1}, OrderedDict()
    for _, handle in _block_entries(buf, idx_off, idx_size):
        off, p2 = _varint(handle, 0)
        size, _ = _varint(handle, p2)
        for key, value in _block_entries(buf, off, size):
            if key == b"":
                for field, _, v in _proto_fields(value):
                    if field == 1:
                        header["num_shards"] = v
                continue
            entries[key.decode()] = _parse_entry(value)
    return header, entries


def load_checkpoint(prefix: str, names=None, verify_crc: bool = False) -> "OrderedDict[str, np.ndarray]":
    """All (or the named) tensors of the checkpoint `prefix` (e.g. ``.../model.ckpt``)."""
    header, entries = read_index(prefix + ".index")
    out, shards = OrderedDict(), {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["dtype"] not in _DTYPES:
            raise CheckpointFormatError(f"{name}: unsupported dtype code {e['dtype']}")
        path = f"{prefix}.data-{e['shard_id']:05d}-of-{header['num_shards']:05d}"
        if path not in shards:
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path}: the tensor data shard is missing (the reference repository ships "
                                        f"only the .index/.meta files of its trained network)")
            shards[path] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = bytes(shards[path][e["offset"]:e["offset"] + e["size"]])
        dt = np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")
        want = int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize
        if len(raw) != e["size"] or e["size"] != want:
            raise CheckpointFormatError(f"{name}: {e['size']} bytes stored, shape {e['shape']} needs {want}")
        if verify_crc and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise CheckpointFormatError(f"{name}: crc32c mismatch")
        out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return out


# ---- writing (single data block per table block is enough for ~100 variables) ---------------------
def _block(entries) -> bytes:
    body, restarts = bytearray(), []
    for key, value in entries:                      # no prefix sharing: every entry is a restart point
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(value)) + key + value
    for r in restarts or [0]:
        body += struct.pack("<I", r)
    body += struct.pack("<I", max(len(restarts), 1))
    return bytes(body)


def _msg(field: int, wt: int, payload) -> bytes:
    tag = _put_varint((field << 3) | wt)
    if wt == 0:
        return tag + _put_varint(payload)
    if wt == 2:
        return tag + _put_varint(len(payload)) + payload
    if wt == 5:
        return tag + payload
    raise ValueError(wt)


def save_checkpoint(prefix: str, tensors: "dict[str, np.ndarray]") -> None:
    """Write `tensors` as a one-shard TensorBundle (``prefix.index`` + ``prefix.data-00000-of-00001``)."""
    data, items = bytearray(), []
    for name in sorted(tensors):
        arr = np.asarray(tensors[name])
        if arr.ndim:                                   # (np.ascontiguousarray would turn a scalar into shape (1,))
            arr = np.ascontiguousarray(arr)
        code = _DTYPE_CODES.get(arr.dtype)
        if code is None:
            raise TypeError(f"{name}: dtype {arr.dtype} has no TensorFlow code here")
        raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
        shape = b"".join(_msg(2, 2, _msg(1, 0, int(d))) for d in arr.shape)
        entry = _msg(1, 0, code) + _msg(2, 2, shape) + (_msg(4, 0, len(data)) if len(data) else b"") + \
            _msg(5, 0, len(raw)) + _msg(6, 5, struct.pack("<I", masked_crc32c(raw)))
        items.append((name.encode(), entry))
        data += raw
    header = _msg(1, 0, 1) + _msg(3, 2, _msg(1, 0, 1))       # num_shards = 1, version {producer: 1}
    blob = bytearray()

    def put(block: bytes) -> bytes:
        off = len(blob)
        blob.extend(block + b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    data_handle = put(_block([(b"", header)] + items))
    meta_handle = put(_block([]))
    index_handle = put(_block([(items[-1][0] if items else b"" , data_handle)]))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(blob) + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# ---- mapping onto the Variables store --------------------------------------------------------------
def tf_name_map(variables, name_scope: str = "6d_pose", available=None) -> "OrderedDict[str, str]":
    """our name -> TensorFlow variable name for every tensor of a Variables store.

    Weights, biases, bn/beta and bn/gamma keep the reference's scope names (utils/tf_util.py:42-43, 164-165,
    488-491).  The moving averages are the shadow variables `ema.average(batch_mean / batch_var)` of
    batch_norm_template (:493-509); TF names them after the moments tensors, including the NAME scope the
    model was built under — ``<scope>/bn/<name_scope>/<scope>/bn/moments/Squeeze/ExponentialMovingAverage``
    (mean) and ``.../Squeeze_1/...`` (variance).  The shipped checkpoint and evaluate_cloudAAE_ycbv.py:436 use
    '6d_pose', train_cloudAAE_ycbv.py:223 uses 'decoder'; with `available` (the names in a checkpoint) the
    scope actually present is picked up."""
    import re

    m = OrderedDict()
    for name in variables.names():
        scope, leaf = name.split("/", 1)
        if leaf in ("bn/ema_mean", "bn/ema_var"):
            sq = "Squeeze" if leaf == "bn/ema_mean" else "Squeeze_1"
            mid = f"{name_scope}/" if name_scope else ""
            tf_name = f"{scope}/bn/{mid}{scope}/bn/moments/{sq}/ExponentialMovingAverage"
            if available is not None and tf_name not in available:
                pat = re.compile(rf"^{re.escape(scope)}/bn/(.*/)?{re.escape(scope)}/bn/moments/{sq}/ExponentialMovingAverage$")
                hits = [a for a in available if pat.match(a)]
                if hits:
                    tf_name = hits[0]
            m[name] = tf_name
        else:
            m[name] = name
    return m


def import_tf_checkpoint(variables, prefix: str, strict: bool = True) -> list[str]:
    """Load a reference checkpoint into `variables` (conv kernels [1,1,cin,cout] / [1,D,1,cout] flatten to
    [fan_in, cout]).  Returns the names that were loaded; with strict=True a missing or mis-shaped tensor raises."""
    import torch

    _, entries = read_index(prefix + ".index")
    mapping = tf_name_map(variables, available=entries)
    wanted = {tf: ours for ours, tf in mapping.items() if tf in entries}
    missing = [ours for ours, tf in mapping.items() if tf not in entries]
    if strict and missing:
        raise KeyError(f"checkpoint {prefix} lacks {len(missing)} variables, e.g. {missing[:4]}")
    tensors = load_checkpoint(prefix, names=set(wanted))
    loaded = []
    with torch.no_grad():
        for tf_name, arr in tensors.items():
            ours = wanted[tf_name]
            dst = variables[ours]
            if int(np.prod(arr.shape)) != dst.numel() or (arr.ndim >= 1 and arr.shape[-1] != dst.shape[-1]):
                if strict:
                    raise CheckpointFormatError(f"{tf_name}: shape {arr.shape} does not fit {tuple(dst.shape)}")
                continue
            dst.copy_(torch.from_numpy(arr.astype(np.float32).reshape(tuple(dst.shape))))
            loaded.append(ours)
    return loaded


def export_tf_checkpoint(variables, prefix: str, name_scope: str = "6d_pose", global_step: float = 0.0,
                         optimizer: dict | None = None) -> None:
    """Write `variables` under the reference's TensorFlow names and kernel shapes ([1,1,cin,cout] for the
    encoder convolutions, [fan_in, cout] for the fully connected layers).

    A TF-1 ``tf.train.Saver()`` restores EVERY global variable of the graph it was built in, so the bundle also
    carries what the reference's graphs hold besides the network: the float32 scalar ``Variable`` (the ``batch`` /
    global-step variable of evaluate_cloudAAE_ycbv.py:411 and train_cloudAAE_ycbv.py:191) and, when `optimizer` =
    {"adam_m": flat tensor, "adam_v": flat tensor, "t": steps taken} is given, the training graph's Adam state:
    ``beta1_power``, ``beta2_power`` and the ``<var>/Adam`` / ``<var>/Adam_1`` slots (train…:259-262)."""
    mapping = tf_name_map(variables, name_scope)
    conv = {s for s, *_ in variables.layers if ("fc" not in s and "output" not in s)}

    def shaped(ours, arr):
        scope = ours.split("/", 1)[0]
        if ours.endswith("/weights") and scope in conv:
            return arr.reshape(1, 1, *arr.shape)
        return arr

    out = {}
    for ours, tf_name in mapping.items():
        out[tf_name] = shaped(ours, variables[ours].detach().cpu().numpy().astype(np.float32))
    out["Variable"] = np.asarray(global_step, np.float32)
    if optimizer is not None:
        t = int(optimizer.get("t", 0))
        out["beta1_power"] = np.asarray(0.9 ** (t + 1), np.float32)      # TF keeps beta^(t+1): initial value beta
        out["beta2_power"] = np.asarray(0.999 ** (t + 1), np.float32)
        for slot, key in (("Adam", "adam_m"), ("Adam_1", "adam_v")):
            flat = optimizer[key].detach().cpu().numpy().astype(np.float32)
            for ours in variables.trainable_names():
                off, shape = variables.index[ours]
                out[f"{mapping[ours]}/{slot}"] = shaped(ours, flat[off:off + int(np.prod(shape))].reshape(shape))
    save_checkpoint(prefix, out)


def import_optimizer_state(variables, prefix: str):
    """The other direction for a restart: global step and Adam slots of a TF-1 checkpoint, when it holds them.
    Returns {"global_step": float | None, "t": int | None, "adam_m": flat tensor | None, "adam_v": ...} laid out like
    ``variables.flat`` (load into CloudAAETrainer.load_state_dict)."""
    import torch

    _, entries = read_index(prefix + ".index")
    mapping = tf_name_map(variables, available=entries)
    want = {"Variable", "beta1_power", "beta2_power"} & set(entries)
    slots = {}
    for ours in variables.trainable_names():
        for slot in ("Adam", "Adam_1"):
            name = f"{mapping.get(ours, ours)}/{slot}"
            if name in entries:
                slots[name] = (ours, slot)
    tensors = load_checkpoint(prefix, names=want | set(slots))
    res = {"global_step": None, "t": None, "adam_m": None, "adam_v": None}
    if "Variable" in tensors:
        res["global_step"] = float(np.asarray(tensors["Variable"]).reshape(-1)[0])
    if "beta1_power" in tensors:
        b1p = float(np.asarray(tensors["beta1_power"]).reshape(-1)[0])
        res["t"] = max(int(round(np.log(b1p) / np.log(0.9))) - 1, 0)
    if slots:
        m = torch.zeros_like(variables.flat, device="cpu"); v = torch.zeros_like(variables.flat, device="cpu")
        for name, (ours, slot) in slots.items():
            off, shape = variables.index[ours]
            dst = m if slot == "Adam" else v
            dst[off:off + int(np.prod(shape))] = torch.from_numpy(np.asarray(tensors[name], np.float32).reshape(-1))
        res["adam_m"], res["adam_v"] = m, v
    return res
