"""Fixture reader: TFRecord framing + tf.Example wire format (no TensorFlow)."""
import os
import struct

import numpy as np
import pytest

from cloudaae_b200.data import tfrecord


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _example(feats):
    entries = b""
    for key, val in feats.items():
        if val.dtype == np.float32:
            feat = _ld(2, _ld(1, val.astype("<f4").tobytes()))
        else:
            feat = _ld(3, _ld(1, b"".join(_varint(int(x) & ((1 << 64) - 1)) for x in val)))
        entries += _ld(1, _ld(1, key.encode()) + _ld(2, feat))
    return _ld(1, entries)


def _write(path, payloads):
    with open(path, "wb") as f:
        for p in payloads:
            f.write(struct.pack("<Q", len(p)) + b"\0\0\0\0" + p + b"\0\0\0\0")


def test_roundtrip_pose_records(tmp_path):
    rng = np.random.default_rng(0)
    t = rng.standard_normal((5, 3)).astype(np.float32)
    a = rng.standard_normal((5, 3)).astype(np.float32)
    path = str(tmp_path / "x.tfrecords")
    _write(path, [_example({"translation": t[i], "axisangle": a[i], "class_id": np.array([7 + i], np.int64)})
                  for i in range(5)])
    tt, aa, cc = tfrecord.read_pose_records(path)
    assert (tt == t).all() and (aa == a).all() and (cc == 7 + np.arange(5)).all()
    assert len(tfrecord.read_pose_records(path, limit=2)[2]) == 2


def test_negative_and_unpacked_int64(tmp_path):
    # a single int64 may be written unpacked (wire type 0)
    feat = _ld(3, _varint((1 << 3) | 0) + _varint((-3) & ((1 << 64) - 1)))
    ex = _ld(1, _ld(1, _ld(1, b"v") + _ld(2, feat)))
    assert tfrecord.parse_example(ex)["v"].tolist() == [-3]


def test_truncated_file_raises(tmp_path):
    path = str(tmp_path / "bad.tfrecords")
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", 100) + b"\0\0\0\0" + b"abc")
    with pytest.raises(ValueError):
        list(tfrecord.iter_records(path))


def test_object_models_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    models = rng.standard_normal((3, 2048, 6)).astype(np.float32)
    path = str(tmp_path / "m.tfrecords")
    _write(path, [_example({"model": models[i].reshape(-1), "label": np.array([i], np.int64)}) for i in (2, 0, 1)])
    got = tfrecord.read_object_models(path)
    assert got.shape == (3, 2048, 6) and (got == models).all()


def test_committed_fixtures_match_reference(have_reference_tree):
    if not have_reference_tree:
        pytest.skip("/root/reference not present (GPU box)")
    import cases
    ref = tfrecord.read_object_models("/root/reference/object_model_tfrecord/obj_models.tfrecords")
    assert (cases.ycb_models() == ref[:, :, :3]).all()
    t, a, c = tfrecord.read_pose_records("/root/reference/ycb_video_data_tfRecords/train_syn/3_syn.tfrecords", limit=4)
    T, A, C = cases.ycb_poses()
    per = len(C) // 21
    assert (T[3 * per:3 * per + 4] == t).all() and (A[3 * per:3 * per + 4] == a).all() and (c == 3).all()
