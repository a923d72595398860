"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on the
B200 box, gloo in the CPU tests).

The reference is single-process / single-GPU (train_cloudAAE_ycbv.py:189); what is added here is
exactly what SURVEY.md §8(e) calls for:
  * every cloud is independent, so FPS / gather / nn_distance / synthesis / inference shard the
    segment list by rank with NO collective (`shard_range`);
  * training is data parallel with ONE exchange: a sum-allreduce of the flat fp32 gradient buffer,
    issued as two buckets so the first (the FC stack and pose heads — 94 % of the bytes, finished
    early in the backward pass) overlaps the encoder backward (`BucketedAllReduce`).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, stop) of `total` independent segments for `rank`.
    The first total % world ranks get one extra segment."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_variables(flat: torch.Tensor, ema: torch.Tensor, group=None, src: int = 0) -> None:
    """Identical initial state on every rank (weights and BN moving averages)."""
    dist.broadcast(flat, src=src, group=group)
    dist.broadcast(ema, src=src, group=group)


class BucketedAllReduce:
    """Sum-allreduce of a flat gradient buffer in buckets [b0,b1), [b1,b2), ... that become ready at
    different points of the backward pass.  On CUDA each bucket runs on a side stream that waits for
    the producing stream, so the collective overlaps the remaining backward kernels; `finish()` makes
    the caller's stream wait for every bucket.  Works in eager mode and under CUDA-graph capture."""

    def __init__(self, flat_grad: torch.Tensor, boundaries: list[int], group=None):
        assert boundaries[0] == 0 and boundaries[-1] == flat_grad.numel() and sorted(boundaries) == list(boundaries)
        self.flat = flat_grad
        self.views = [flat_grad[a:b] for a, b in zip(boundaries[:-1], boundaries[1:])]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        # the collective outranks everything: at the default (low) priority its CTAs would queue behind the
        # not-yet-dispatched CTAs of the synthesis branch of a pipelined graph and the exchange would start late
        prio = int(os.environ.get("CLOUDAAE_NCCL_PRIORITY", "-2"))
        self.side = torch.cuda.Stream(flat_grad.device, priority=prio) if self.cuda else None
        self._pending = []

    def start(self, bucket: int) -> None:
        """Call when every gradient inside `bucket` has been written on the current stream."""
        if self.world == 1:
            return
        view = self.views[bucket]
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
                done = torch.cuda.Event()
                done.record(self.side)
            self._pending.append(done)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self) -> None:
        if self.cuda:
            cur = torch.cuda.current_stream(self.flat.device)
            for ev in self._pending:
                cur.wait_event(ev)
        self._pending = []
