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


# ---- evaluation records (<seq>_pcnn.tfrecord) and the quaternion conversion ---------------------------------

def _eval_frame(seed, classes, four_channel=False):
    import cases
    from cloudaae_b200.data.synthetic_frames import YCBV_INTRINSICS, render_frame
    rng = np.random.default_rng(seed)
    clouds = cases.posed_ycb_clouds(seed % 4)
    depth, label = render_frame(clouds[list(classes)], list(classes), h=120, w=160, splat=1, seed=seed, n_stray=20)
    one_hot = np.zeros(21, np.int64); one_hot[list(classes)] = 1
    q = rng.standard_normal((21, 4)).astype(np.float32)
    return {"image": rng.integers(0, 255, (120, 160, 4 if four_channel else 3), dtype=np.uint8), "depth": depth, "label": label,
            "quaternions": q / np.linalg.norm(q, axis=1, keepdims=True), "translations": rng.standard_normal((21, 3)).astype(np.float32),
            "class_one_hot": one_hot, "seq_id": 48 + seed, "frame_id": 7 * seed + 1,
            "fx": float(YCBV_INTRINSICS[0]), "fy": float(YCBV_INTRINSICS[1]), "cx": float(YCBV_INTRINSICS[2]),
            "cy": float(YCBV_INTRINSICS[3]), "factor_depth": float(YCBV_INTRINSICS[4])}


def test_eval_frame_records_round_trip(tmp_path):
    frames = [_eval_frame(0, (0, 3)), _eval_frame(1, (3, 7, 9), four_channel=True), _eval_frame(2, (5,))]
    path = str(tmp_path / "0048_pcnn.tfrecord")
    tfrecord.write_records(path, [tfrecord.encode_eval_frame(f) for f in frames])
    back = tfrecord.read_eval_frames(path)
    assert len(back) == 3
    for f, b in zip(frames, back):
        assert (b["depth"] == f["depth"]).all() and b["depth"].dtype == np.uint16
        assert (b["label"] == f["label"]).all() and (b["image"] == f["image"][:, :, :3]).all()   # 4th channel dropped
        assert (b["quaternions"] == f["quaternions"]).all() and (b["translations"] == f["translations"]).all()
        assert (b["class_one_hot"] == f["class_one_hot"]).all()
        assert (b["seq_id"], b["frame_id"]) == (f["seq_id"], f["frame_id"])
        assert b["fx"] == pytest.approx(f["fx"], rel=1e-7) and b["factor_depth"] == 10000.0
    # the dataset filter of evaluate…:315 and the record framing: valid masked CRC-32C words
    assert [b["seq_id"] for b in tfrecord.read_eval_frames(path, target_class=3)] == [48, 49]
    assert len(tfrecord.read_eval_frames(path, limit=1)) == 1
    from cloudaae_b200.data.tf_checkpoint import masked_crc32c
    raw = open(path, "rb").read()
    (n,) = struct.unpack_from("<Q", raw, 0)
    assert struct.unpack_from("<I", raw, 8)[0] == masked_crc32c(raw[:8])
    assert struct.unpack_from("<I", raw, 12 + n)[0] == masked_crc32c(raw[12:12 + n])


def test_frames_to_front_end_inputs_lists_one_segment_per_visible_class():
    frames = [_eval_frame(0, (0, 3)), _eval_frame(1, (3, 7, 9))]
    x = tfrecord.frames_to_front_end_inputs(frames)
    assert x["depth"].shape == (2, 120, 160) and x["depth"].dtype == np.uint16 and x["label"].dtype == np.uint8
    assert x["intrinsics"].shape == (2, 5) and x["intrinsics"].dtype == np.float32
    assert x["frame_of_seg"].tolist() == [0, 0, 1, 1, 1] and x["class_of_seg"].tolist() == [0, 3, 3, 7, 9]
    assert (x["quaternion"][2] == frames[1]["quaternions"][3]).all() and (x["translation"][4] == frames[1]["translations"][9]).all()
    only3 = tfrecord.frames_to_front_end_inputs(frames, target_class=3)
    assert only3["frame_of_seg"].tolist() == [0, 1] and only3["class_of_seg"].tolist() == [3, 3]
    # the oracle front end accepts exactly these arrays
    from oracle import evaluation as E
    org, flt, pix, mean = E.segment_extract(x["depth"][1], x["label"][1], x["intrinsics"][1], 7, 0.2)
    assert len(org) > 0 and len(flt) <= len(org)


def test_quat2axag_matches_scipy_rotations():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    q = rng.standard_normal((64, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[0] = [1, 0, 0, 0]                       # identity
    q[1] = [-0.2, 0.5, 0.1, -0.8]             # w < 0: angle above pi, as transforms3d returns it
    q[2] = 3.0 * q[2]                         # not normalised
    a = tfrecord.quat2axag(q.astype(np.float32))
    assert a.dtype == np.float32 and a.shape == (64, 3) and (a[0] == 0).all()
    assert np.linalg.norm(a[1]) > np.pi
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    want = Rotation.from_quat(qn[:, [1, 2, 3, 0]]).as_matrix()
    got = Rotation.from_rotvec(a.astype(np.float64)).as_matrix()
    np.testing.assert_allclose(got, want, atol=2e-6)
    assert np.isnan(tfrecord.quat2axag(np.array([[np.inf, 0, 0, 0]]))).all()
