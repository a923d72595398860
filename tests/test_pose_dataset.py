"""Record side of the training input pipeline (full-permutation shuffle, drop_remainder batching, rank slices)."""
import os

import numpy as np
import pytest

import cases
from cloudaae_b200.data import tfrecord
from cloudaae_b200.data.pose_dataset import PoseRecordDataset


def _ds():
    t, a, c = cases.ycb_poses()
    return PoseRecordDataset(c, a, t)


def test_epoch_is_a_permutation_cut_into_full_batches():
    ds = _ds()
    n, b = len(ds), 128
    assert ds.steps_per_epoch(b) == n // b
    seen = []
    for bt in ds.epoch(b, seed=3):
        assert bt["class_id"].shape == (b,) and bt["class_id"].dtype == np.int32
        assert bt["axisangle"].shape == (b, 3) and bt["translation"].dtype == np.float32
        seen.append(np.concatenate([bt["class_id"][:, None].astype(np.float64), bt["axisangle"], bt["translation"]], 1))
    seen = np.concatenate(seen)
    assert len(seen) == (n // b) * b                                   # remainder dropped
    full = np.concatenate([ds.class_id[:, None].astype(np.float64), ds.axisangle, ds.translation], 1)
    # the drawn records are a sub-multiset of the record set: nothing invented, nothing drawn more often than it exists
    from collections import Counter
    have, drawn = Counter(r.tobytes() for r in full), Counter(r.tobytes() for r in seen)
    assert all(have[k] >= v for k, v in drawn.items())
    # shuffled, reproducible, different per epoch
    first = next(iter(ds.epoch(b, seed=3)))["translation"]
    assert (first == next(iter(ds.epoch(b, seed=3)))["translation"]).all()
    assert not (first == next(iter(ds.epoch(b, seed=3, epoch=1)))["translation"]).all()
    assert not (first == ds.translation[:b]).all()
    assert (next(iter(ds.epoch(b, seed=3, shuffle=False)))["translation"] == ds.translation[:b]).all()


def test_ranks_take_disjoint_slices_of_the_same_global_batches():
    ds = _ds()
    b, world = 32, 4
    per_rank = [list(ds.epoch(b, seed=9, rank=r, world=world)) for r in range(world)]
    assert all(len(p) == ds.steps_per_epoch(b, world) == len(ds) // (b * world) for p in per_rank)
    single = list(ds.epoch(b * world, seed=9))
    for s in range(len(single)):
        glob = np.concatenate([per_rank[r][s]["translation"] for r in range(world)])
        assert (glob == single[s]["translation"]).all()                # rank r = rows [r*B, (r+1)*B) of the global batch
    with pytest.raises(ValueError):
        next(ds.epoch(b, seed=0, rank=4, world=4))


def test_from_tfrecords_reads_the_reference_layout(tmp_path):
    t, a, c = cases.ycb_poses()
    for cls in (0, 1):
        m = c == cls
        recs = [tfrecord.encode_example({"translation": t[i], "axisangle": a[i], "class_id": np.asarray([c[i]], np.int64)})
                for i in np.flatnonzero(m)[:50]]
        tfrecord.write_records(str(tmp_path / f"{cls}_syn.tfrecords"), recs)
    ds = PoseRecordDataset.from_tfrecords(str(tmp_path))
    assert len(ds) == 100 and (ds.class_id[:50] == 0).all() and (ds.class_id[50:] == 1).all()
    assert (ds.translation[:50] == t[c == 0][:50]).all() and (ds.axisangle[50:] == a[c == 1][:50]).all()
    assert len(PoseRecordDataset.from_tfrecords(str(tmp_path), limit_per_file=7)) == 14
    z = os.path.join(cases.GOLDEN, "ycb_poses.npz")
    assert len(PoseRecordDataset.from_npz(z)) == len(c)
    with pytest.raises(FileNotFoundError):
        PoseRecordDataset.from_tfrecords(str(tmp_path / "nothing"))


def test_pinned_ring_yields_the_same_batches():
    import torch
    ds = _ds()
    try:
        torch.empty(1).pin_memory()
    except RuntimeError:
        pytest.skip("pinned memory needs a CUDA runtime")
    want = list(ds.epoch(64, seed=1))[:5]
    for i, bt in enumerate(ds.pinned_batches(64, seed=1)):
        if i == 5:
            break
        assert bt["translation"].is_pinned() and (bt["translation"].numpy() == want[i]["translation"]).all()
        assert (bt["class_id"].numpy() == want[i]["class_id"]).all()



def test_pinned_ring_never_refills_a_slot_before_its_release_event():
    """The ring re-uses a slot `depth` batches later; it must wait for the consumer's release event (recorded after the
    asynchronous host-to-device copy) — and fall back to a full synchronisation when a slot was never released."""
    import torch
    ds = _ds()
    want = list(ds.epoch(16, seed=3))[:6]
    waited = []

    class FakeEvent:
        def __init__(self, tag): self.tag = tag
        def synchronize(self): waited.append(self.tag)

    it = ds.pinned_batches(16, seed=3, depth=2)
    for i in range(6):
        bt = next(it)
        assert callable(bt["release"])
        assert (bt["class_id"].numpy() == want[i]["class_id"]).all() and (bt["axisangle"].numpy() == want[i]["axisangle"]).all()
        bt["release"]()                      # CPU build: no event; the slot is simply handed back
    if torch.cuda.is_available():
        it = ds.pinned_batches(16, seed=3, depth=2)
        seen = []
        for i in range(5):
            bt = next(it)
            dev = {k: bt[k].cuda(non_blocking=True) for k in ("class_id", "axisangle", "translation")}
            bt["release"]()
            seen.append(dev)
        torch.cuda.synchronize()
        for i, dev in enumerate(seen):       # no torn or overwritten batch
            assert (dev["class_id"].cpu().numpy() == want[i]["class_id"]).all()
            assert (dev["translation"].cpu().numpy() == want[i]["translation"]).all()
