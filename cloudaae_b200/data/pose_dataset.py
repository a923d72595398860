"""The training input pipeline's record side — ``TFRecordDataset(train_filenames).shuffle(3042462 * 13)`` …
``.batch(batch_size, drop_remainder=True)`` (train_cloudAAE_ycbv.py:177, 114) — as host arrays.

The reference shuffles with a buffer larger than the data set, i.e. every epoch is one uniform permutation of all
pose records (381,553 in ``ycb_video_data_tfRecords/train_syn``), cut into batches with the remainder dropped
(2,980 steps per epoch at batch 128).  Everything downstream of the records (pose transform, occluder, hidden
point removal) happens on the GPU (cloudaae_b200.synthesis), so a batch here is just 28 bytes per segment:
class id, axis-angle, translation.  For data-parallel training the GLOBAL batch is drawn from the permutation
and rank r takes rows [r*B, (r+1)*B) of it: the ranks see disjoint records and an epoch still visits every
record at most once.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, Iterator, Optional, Sequence

import numpy as np

from . import tfrecord

KEYS = ("class_id", "axisangle", "translation")


class PoseRecordDataset:
    def __init__(self, class_id: np.ndarray, axisangle: np.ndarray, translation: np.ndarray):
        n = len(class_id)
        if axisangle.shape != (n, 3) or translation.shape != (n, 3):
            raise ValueError("PoseRecordDataset expects class_id [R], axisangle [R,3], translation [R,3]")
        self.class_id = np.ascontiguousarray(class_id, np.int32)
        self.axisangle = np.ascontiguousarray(axisangle, np.float32)
        self.translation = np.ascontiguousarray(translation, np.float32)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_tfrecords(cls, paths: Sequence[str] | str, limit_per_file: Optional[int] = None) -> "PoseRecordDataset":
        """`paths`: files, or a directory holding the reference's ``<class>_syn.tfrecords``."""
        if isinstance(paths, str):
            paths = sorted(glob.glob(os.path.join(paths, "*.tfrecords"))) if os.path.isdir(paths) else [paths]
        if not paths:
            raise FileNotFoundError("PoseRecordDataset.from_tfrecords: no record file")
        t, a, c = zip(*(tfrecord.read_pose_records(p, limit_per_file) for p in paths))
        return cls(np.concatenate(c), np.concatenate(a), np.concatenate(t))

    @classmethod
    def from_npz(cls, path: str) -> "PoseRecordDataset":
        z = np.load(path)
        return cls(z["class_id"], z["axisangle"], z["translation"])

    def __len__(self) -> int:
        return len(self.class_id)

    def steps_per_epoch(self, batch_size: int, world: int = 1) -> int:
        return len(self) // (batch_size * world)

    # ------------------------------------------------------------------ iteration
    def epoch(self, batch_size: int, seed: int, epoch: int = 0, rank: int = 0, world: int = 1,
              shuffle: bool = True) -> Iterator[Dict[str, np.ndarray]]:
        """One pass: a uniform permutation (seeded by (seed, epoch), identical on every rank), global batches of
        batch_size * world records with the remainder dropped, this rank's slice of each."""
        if not 0 <= rank < world:
            raise ValueError("rank outside [0, world)")
        n = len(self)
        order = np.random.default_rng([seed, epoch]).permutation(n) if shuffle else np.arange(n)
        g = batch_size * world
        for s in range(n // g):
            sel = order[s * g + rank * batch_size: s * g + (rank + 1) * batch_size]
            yield {"class_id": self.class_id[sel], "axisangle": self.axisangle[sel], "translation": self.translation[sel]}

    def pinned_batches(self, batch_size: int, seed: int, epoch: int = 0, rank: int = 0, world: int = 1, depth: int = 2):
        """The same batches in a small ring of pinned host tensors (what `static.copy_(…, non_blocking=True)` of a
        captured training step wants).

        A slot is re-used `depth` batches later, and the host runs far ahead of the GPU (a step takes ~2 ms, refilling a
        slot microseconds), so the consumer MUST tell the ring when its asynchronous host-to-device copy of a slot has
        been enqueued: call ``batch["release"]()`` right after the ``copy_(..., non_blocking=True)`` calls, on the
        stream that performs them.  It records a CUDA event; the slot is refilled only after that event has
        completed.  A slot that was never released is refilled only after a full device synchronisation — slow, but
        never torn (class_id of one batch with the pose of another)."""
        import torch
        cuda = torch.cuda.is_available()
        ring = [{"class_id": torch.empty(batch_size, dtype=torch.int32).pin_memory() if cuda else torch.empty(batch_size, dtype=torch.int32),
                 "axisangle": torch.empty(batch_size, 3).pin_memory() if cuda else torch.empty(batch_size, 3),
                 "translation": torch.empty(batch_size, 3).pin_memory() if cuda else torch.empty(batch_size, 3)}
                for _ in range(max(depth, 1))]
        state = [{"event": None, "handed_out": False} for _ in ring]

        def make_release(st):
            def release(stream=None):
                if cuda:
                    ev = torch.cuda.Event()
                    ev.record(stream if stream is not None else torch.cuda.current_stream())
                    st["event"] = ev
                st["handed_out"] = False
            return release

        for i, bt in enumerate(self.epoch(batch_size, seed, epoch, rank, world)):
            slot, st = ring[i % len(ring)], state[i % len(ring)]
            if st["event"] is not None:
                st["event"].synchronize()             # the copy that read this slot has finished
                st["event"] = None
            elif st["handed_out"] and cuda:
                torch.cuda.synchronize()              # consumer never released the slot: be safe, not fast
            for k in KEYS:
                slot[k].copy_(torch.from_numpy(bt[k]))
            st["handed_out"] = True
            slot["release"] = make_release(st)
            yield slot
