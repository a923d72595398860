"""TensorFlow-1 checkpoint reader / writer (cloudaae_b200/data/tf_checkpoint.py) — the tf.train.Saver pair the
reference writes (train_cloudAAE_ycbv.py:276,423-430) and restores (evaluate_cloudAAE_ycbv.py:495-499)."""
import os

import numpy as np
import pytest
import torch

from cloudaae_b200.data import tf_checkpoint as T
from cloudaae_b200.models.pointnet_ycb_23_decoder_4 import Variables, dgcnn_layers, pn_layers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_model.ckpt.index")


def test_reads_the_reference_snapshot_index():
    header, entries = T.read_index(GOLDEN)
    assert header["num_shards"] == 1 and len(entries) == 175
    assert list(entries) == sorted(entries)                       # a sorted table
    assert entries["dgcnn1/weights"]["shape"] == (1, 1, 48, 64)
    assert entries["dgcnn_agg/weights"]["shape"] == (1, 1, 320, 1024)
    assert entries["dgcnn_output/weights"]["shape"] == (1024, 3072)
    trainable = [n for n in entries if "Adam" not in n and "ExponentialMovingAverage" not in n and "/" in n]
    assert sum(int(np.prod(entries[n]["shape"])) for n in trainable) == 6936518
    # tensors are laid out back to back in one shard
    off = 0
    for e in sorted(entries.values(), key=lambda e: e["offset"]):
        assert e["offset"] == off and e["dtype"] == 1
        off += e["size"]


def test_every_variable_of_the_store_maps_onto_the_reference_names():
    v = Variables(dgcnn_layers(256, 24), device="cpu", seed=0)
    _, entries = T.read_index(GOLDEN)
    mapping = T.tf_name_map(v, available=entries)
    assert len(mapping) == 72
    for ours, tf_name in mapping.items():
        assert tf_name in entries, (ours, tf_name)
        assert int(np.prod(entries[tf_name]["shape"])) == v[ours].numel()
    assert mapping["dgcnn1/bn/ema_var"] == "dgcnn1/bn/6d_pose/dgcnn1/bn/moments/Squeeze_1/ExponentialMovingAverage"
    with pytest.raises(FileNotFoundError, match="data shard is missing"):
        T.import_tf_checkpoint(v, GOLDEN[:-len(".index")])


@pytest.mark.parametrize("layers", [dgcnn_layers(256, 24), pn_layers(256, 24)])
def test_export_import_round_trip(tmp_path, layers):
    src = Variables(layers, device="cpu", seed=5)
    with torch.no_grad():
        src.ema.copy_(torch.rand_like(src.ema))
    prefix = str(tmp_path / "model.ckpt")
    T.export_tf_checkpoint(src, prefix)
    header, entries = T.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and list(entries) == sorted(entries)
    first = layers[0][0]
    assert entries[f"{first}/weights"]["shape"] == (1, 1, layers[0][1], layers[0][2])   # conv kernels keep TF's rank 4
    tensors = T.load_checkpoint(prefix, verify_crc=True)
    assert np.array_equal(tensors[f"{first}/weights"].reshape(layers[0][1], layers[0][2]), src[f"{first}/weights"].numpy())
    dst = Variables(layers, device="cpu", seed=6)
    loaded = T.import_tf_checkpoint(dst, prefix)
    assert sorted(loaded) == sorted(src.names())
    assert torch.equal(dst.flat, src.flat) and torch.equal(dst.ema, src.ema)
    # a different name scope of the moving averages is found by pattern
    T.export_tf_checkpoint(src, prefix, name_scope="decoder")
    dst2 = Variables(layers, device="cpu", seed=7)
    T.import_tf_checkpoint(dst2, prefix)
    assert torch.equal(dst2.ema, src.ema)


def test_corrupt_files_are_rejected(tmp_path):
    p = tmp_path / "x.index"
    p.write_bytes(b"\x00" * 64)
    with pytest.raises(T.CheckpointFormatError, match="magic"):
        T.read_index(str(p))
    v = Variables(dgcnn_layers(256, 24), device="cpu", seed=1)
    prefix = str(tmp_path / "m.ckpt")
    T.export_tf_checkpoint(v, prefix)
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[100] ^= 0xFF
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(T.CheckpointFormatError, match="crc32c"):
        T.load_checkpoint(prefix, verify_crc=True)
    assert T.masked_crc32c(b"123456789") == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_exported_bundle_holds_everything_a_tf1_saver_restores(tmp_path):
    """tf.train.Saver() restores every global variable of its graph: the evaluation graph needs the float32 scalar
    'Variable' (evaluate_cloudAAE_ycbv.py:411), the training graph also beta1_power / beta2_power and the Adam slots —
    the exported key set must cover the reference snapshot's own index."""
    v = Variables(dgcnn_layers(256, 24), device="cpu", seed=2)
    m = torch.rand_like(v.flat); vv = torch.rand_like(v.flat)
    prefix = str(tmp_path / "model.ckpt")
    T.export_tf_checkpoint(v, prefix, global_step=1234.0, optimizer={"adam_m": m, "adam_v": vv, "t": 7})
    _, ours = T.read_index(prefix + ".index")
    _, ref = T.read_index(GOLDEN)
    assert set(ref) <= set(ours), sorted(set(ref) - set(ours))[:5]
    for name in ref:                                             # same shapes and dtypes as TensorFlow wrote them
        assert ours[name]["shape"] == ref[name]["shape"] and ours[name]["dtype"] == ref[name]["dtype"], name
    st = T.import_optimizer_state(Variables(dgcnn_layers(256, 24), device="cpu", seed=3), prefix)
    assert st["global_step"] == 1234.0 and st["t"] == 7
    for name in ("dgcnn_agg/weights", "dgcnn_output/biases", "dgcnn1/bn/gamma"):
        off, shape = v.index[name]
        k = int(np.prod(shape))
        assert torch.equal(st["adam_m"][off:off + k], m[off:off + k]) and torch.equal(st["adam_v"][off:off + k], vv[off:off + k])
