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
