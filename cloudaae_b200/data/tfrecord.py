"""TFRecord + ``tf.train.Example`` reader that needs no TensorFlow.

The reference reads its two fixtures with ``tf.python_io.tf_record_iterator`` /
``tf.data.TFRecordDataset`` + ``tf.parse_single_example``
(train_cloudAAE_ycbv.py:40-65).  This module parses the same bytes directly:

* TFRecord framing: ``uint64 len | uint32 crc(len) | payload | uint32 crc(payload)``
* ``Example{1: Features{1: map<string, Feature>}}`` with
  ``Feature{1: BytesList | 2: FloatList | 3: Int64List}``, each list in field 1.

Only what the two fixtures use is implemented (packed floats, packed or
unpacked int64 varints, bytes).  CRCs are not verified.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Union

import numpy as np

FeatureValue = Union[np.ndarray, List[bytes]]


def iter_records(path: str) -> Iterator[bytes]:
    """Yield the payload of every record of a TFRecord file."""
    with open(path, "rb") as f:
        data = f.read()
    pos, end = 0, len(data)
    while pos < end:
        if pos + 12 > end:
            raise ValueError(f"{path}: truncated record header at byte {pos}")
        (length,) = struct.unpack_from("<Q", data, pos)
        pos += 12  # length + its crc
        if pos + length + 4 > end:
            raise ValueError(f"{path}: truncated record payload at byte {pos}")
        yield data[pos:pos + length]
        pos += length + 4


def _varint(buf: bytes, pos: int):
    result = 0
    shift = 0
    while True:
        byte = buf[pos]
        pos += 1
        result |= (byte & 0x7F) << shift
        if not byte & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) over one protobuf message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            value, pos = _varint(buf, pos)
        elif wire == 1:
            value = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            length, pos = _varint(buf, pos)
            value = buf[pos:pos + length]
            pos += length
        elif wire == 5:
            value = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wire}")
        yield field, wire, value


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_feature(buf: bytes) -> FeatureValue:
    for field, _, value in _fields(buf):
        if field == 1:  # BytesList
            return [v for f, _, v in _fields(value) if f == 1]
        if field == 2:  # FloatList
            chunks = []
            for f, wire, v in _fields(value):
                if f != 1:
                    continue
                chunks.append(np.frombuffer(v, dtype="<f4"))  # packed, or one fixed32
            return np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
        if field == 3:  # Int64List
            out: List[int] = []
            for f, wire, v in _fields(value):
                if f != 1:
                    continue
                if wire == 0:
                    out.append(_signed64(v))
                else:
                    p = 0
                    while p < len(v):
                        x, p = _varint(v, p)
                        out.append(_signed64(x))
            return np.asarray(out, dtype=np.int64)
    return np.zeros(0, np.float32)


def parse_example(payload: bytes) -> Dict[str, FeatureValue]:
    """Decode one serialized ``tf.train.Example`` into ``{key: ndarray | [bytes]}``."""
    out: Dict[str, FeatureValue] = {}
    for field, _, features in _fields(payload):
        if field != 1:
            continue
        for f, _, entry in _fields(features):
            if f != 1:
                continue
            key, feat = None, None
            for ef, _, ev in _fields(entry):
                if ef == 1:
                    key = ev.decode("utf-8")
                elif ef == 2:
                    feat = ev
            if key is not None and feat is not None:
                out[key] = _parse_feature(feat)
    return out


def read_object_models(path: str) -> np.ndarray:
    """``obj_models.tfrecords`` -> float32 [num_class, 2048, 6] ordered by label.

    Mirrors ``read_and_decode_obj_model`` (train_cloudAAE_ycbv.py:40-54).
    """
    models, labels = [], []
    for rec in iter_records(path):
        ex = parse_example(rec)
        models.append(np.asarray(ex["model"], np.float32).reshape(2048, 6))
        labels.append(int(ex["label"][0]))
    order = np.argsort(np.asarray(labels), kind="stable")
    return np.stack(models)[order]


def read_pose_records(path: str, limit: int | None = None):
    """``<cls>_syn.tfrecords`` -> (translation f32[R,3], axisangle f32[R,3], class_id i64[R]).

    Mirrors ``decode`` (train_cloudAAE_ycbv.py:57-65).
    """
    t, a, c = [], [], []
    for i, rec in enumerate(iter_records(path)):
        if limit is not None and i >= limit:
            break
        ex = parse_example(rec)
        t.append(np.asarray(ex["translation"], np.float32))
        a.append(np.asarray(ex["axisangle"], np.float32))
        c.append(int(ex["class_id"][0]))
    return (np.stack(t).astype(np.float32), np.stack(a).astype(np.float32),
            np.asarray(c, np.int64))


# ---- evaluation records (<seq>_pcnn.tfrecord) -----------------------------------------------------------
# One record = one YCB-Video frame (evaluate_cloudAAE_ycbv.py:125-160, `decode`): raw uint8 image bytes +
# image_shape, raw uint16 depth bytes + depth_shape, raw uint8 label bytes + label_shape, per-class
# quaternions f32[21,4] (w, x, y, z) / translations f32[21,3] / class_one_hot i64[21], seq_id, frame_id and
# the camera (fx, fy, cx, cy, factor_depth).  The frames feed cloudaae_b200.evaluate_cloudAAE_ycbv.SegmentFrontEnd.

NUM_CLASS_YCB = 21


def decode_eval_frame(payload: bytes) -> Dict[str, Union[np.ndarray, int, float]]:
    """One serialized evaluation Example -> dict of arrays / scalars (mirrors `decode`, evaluate…:125-160,
    including the drop of a fourth image channel, :150-151)."""
    ex = parse_example(payload)

    def scalar(key, cast):
        return cast(np.asarray(ex[key]).reshape(-1)[0])

    image_shape = tuple(int(v) for v in ex["image_shape"])
    depth_shape = tuple(int(v) for v in ex["depth_shape"])
    label_shape = tuple(int(v) for v in ex["label_shape"])
    image = np.frombuffer(ex["image"][0], dtype=np.uint8).reshape(image_shape)
    if image.shape[2] == 4:
        image = image[:, :, :3]
    return {
        "image": image,
        "depth": np.frombuffer(ex["depth"][0], dtype="<u2").reshape(depth_shape),
        "label": np.frombuffer(ex["label"][0], dtype=np.uint8).reshape(label_shape),
        "quaternions": np.asarray(ex["quaternions"], np.float32).reshape(NUM_CLASS_YCB, 4),
        "translations": np.asarray(ex["translations"], np.float32).reshape(NUM_CLASS_YCB, 3),
        "class_one_hot": np.asarray(ex["class_one_hot"], np.int64).reshape(NUM_CLASS_YCB),
        "seq_id": scalar("seq_id", int), "frame_id": scalar("frame_id", int),
        "fx": scalar("fx", float), "fy": scalar("fy", float), "cx": scalar("cx", float), "cy": scalar("cy", float),
        "factor_depth": scalar("factor_depth", float),
    }


def read_eval_frames(path: str, limit: int | None = None, target_class: int | None = None):
    """`<seq>_pcnn.tfrecord` -> list of frame dicts; `target_class` keeps the frames showing that class
    (the dataset filter of evaluate…:315)."""
    out = []
    for rec in iter_records(path):
        fr = decode_eval_frame(rec)
        if target_class is not None and fr["class_one_hot"][target_class] != 1:
            continue
        out.append(fr)
        if limit is not None and len(out) >= limit:
            break
    return out


def frames_to_front_end_inputs(frames, target_class: int | None = None):
    """Stack decoded frames into the arrays SegmentFrontEnd takes — depth u16[F,h,w], label u8[F,h,w],
    intrinsics f32[F,5] — plus the segment list of `split_samples` (evaluate…:187-216): one (frame, class)
    pair per class present in a frame (only `target_class` when given, evaluate…:319), with its ground-truth
    quaternion and translation."""
    depth = np.stack([f["depth"] for f in frames]).astype(np.uint16)
    label = np.stack([f["label"] for f in frames]).astype(np.uint8)
    intr = np.array([[f["fx"], f["fy"], f["cx"], f["cy"], f["factor_depth"]] for f in frames], np.float32)
    frame_of_seg, class_of_seg, quat, trans = [], [], [], []
    for i, f in enumerate(frames):
        for c in np.flatnonzero(f["class_one_hot"]):
            if target_class is not None and c != target_class:
                continue
            frame_of_seg.append(i); class_of_seg.append(int(c))
            quat.append(f["quaternions"][c]); trans.append(f["translations"][c])
    return {"depth": depth, "label": label, "intrinsics": intr,
            "frame_of_seg": np.asarray(frame_of_seg, np.int32), "class_of_seg": np.asarray(class_of_seg, np.int32),
            "quaternion": np.asarray(quat, np.float32).reshape(-1, 4), "translation": np.asarray(trans, np.float32).reshape(-1, 3)}


def quat2axag(quaternion: np.ndarray) -> np.ndarray:
    """(w, x, y, z) quaternions [B,4] -> axis-angle vectors angle * axis, float32 [B,3] — `quat2axag_batch` +
    `quat2axag_tf` (evaluate…:66-79).  The reference calls transforms3d.quaternions.quat2axangle (not in its tree,
    no version pinned); its published algorithm: normalise, axis = (x,y,z)/|(x,y,z)|, angle = 2 acos(clamp(w)) in
    [0, 2 pi], identity -> axis (1,0,0), angle 0."""
    q = np.asarray(quaternion, np.float64).reshape(-1, 4)
    out = np.zeros((q.shape[0], 3), np.float32)
    eps = np.finfo(np.float64).eps
    for k in range(q.shape[0]):
        w, x, y, z = q[k]
        nq = w * w + x * x + y * y + z * z
        if not np.isfinite(nq):
            out[k] = np.nan
            continue
        if nq < eps ** 2:
            continue
        s = np.sqrt(nq)
        w, x, y, z = w / s, x / s, y / s, z / s
        len2 = x * x + y * y + z * z
        if len2 < (3 * eps) ** 2:
            continue
        theta = 2.0 * np.arccos(max(min(w, 1.0), -1.0))
        axis = np.array([x, y, z]) / np.sqrt(len2)
        # float32 as in the reference: axag4 is a float32 array, then angle * axis
        out[k] = np.float32(theta) * axis.astype(np.float32)
    return out


# ---- writer (fixtures, exports) ---------------------------------------------------------------------------

def _enc_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_field(field: int, payload: bytes) -> bytes:
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(features: Dict[str, FeatureValue]) -> bytes:
    """{key: float array | int array | bytes | [bytes]} -> serialized tf.train.Example (packed lists)."""
    entries = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, list):
            feat = _enc_field(1, b"".join(_enc_field(1, b) for b in v))
        else:
            a = np.asarray(v)
            if a.dtype.kind == "f":
                feat = _enc_field(2, _enc_field(1, a.astype("<f4").tobytes()))
            else:
                feat = _enc_field(3, _enc_field(1, b"".join(_enc_varint(int(x)) for x in a.reshape(-1))))
        entries += _enc_field(1, _enc_field(1, key.encode("utf-8")) + _enc_field(2, feat))
    return _enc_field(1, entries)


def write_records(path: str, payloads) -> None:
    """TFRecord framing with valid masked CRC-32C words (tf.data / tf_record_iterator verify them)."""
    from .tf_checkpoint import masked_crc32c
    with open(path, "wb") as f:
        for p in payloads:
            head = struct.pack("<Q", len(p))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + p + struct.pack("<I", masked_crc32c(p)))


def encode_eval_frame(frame) -> bytes:
    """Inverse of decode_eval_frame."""
    img, dep, lab = np.ascontiguousarray(frame["image"], np.uint8), np.ascontiguousarray(frame["depth"], "<u2"), \
        np.ascontiguousarray(frame["label"], np.uint8)
    return encode_example({
        "image": img.tobytes(), "image_shape": np.asarray(img.shape, np.int64),
        "depth": dep.tobytes(), "depth_shape": np.asarray(dep.shape, np.int64),
        "label": lab.tobytes(), "label_shape": np.asarray(lab.shape, np.int64),
        "quaternions": np.asarray(frame["quaternions"], np.float32).reshape(-1),
        "translations": np.asarray(frame["translations"], np.float32).reshape(-1),
        "class_one_hot": np.asarray(frame["class_one_hot"], np.int64),
        "seq_id": np.asarray([frame["seq_id"]], np.int64), "frame_id": np.asarray([frame["frame_id"]], np.int64),
        "fx": np.asarray([frame["fx"]], np.float32), "fy": np.asarray([frame["fy"]], np.float32),
        "cx": np.asarray([frame["cx"]], np.float32), "cy": np.asarray([frame["cy"]], np.float32),
        "factor_depth": np.asarray([frame["factor_depth"]], np.float32),
    })
