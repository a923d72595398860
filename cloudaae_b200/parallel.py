"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on the
B200 box, gloo in the CPU tests).

The reference is single-process / single-GPU (train_cloudAAE_ycbv.py:189); what is added here is
exactly what SURVEY.md §8(e) calls for:
  * every cloud is independent, so FPS / gather / nn_distance / synthesis / inference shard the
    segment list by rank with NO collective (`shard_range`);
  * training is data parallel with ONE exchange: a sum-allreduce of the flat fp32 gradient buffer,
    issued as three buckets in backward order so the first (the FC stack and pose heads — 94 % of the
    bytes, finished early in the backward pass) overlaps the encoder backward (`BucketedAllReduce`);
  * the step is bound by SM work, not by the exchange, so the NCCL communicator is capped at a few CTAs
    (`init_nccl`): the 26 MB bucket has the whole encoder backward (~0.5 ms) to finish in.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def nccl_options(max_ctas: int | None = None):
    """ProcessGroupNCCL options with the communicator's CTA count capped (None: CLOUDAAE_NCCL_MAX_CTAS, default 8;
    0: no cap -> returns None, NCCL chooses)."""
    if max_ctas is None:
        max_ctas = int(os.environ.get("CLOUDAAE_NCCL_MAX_CTAS", "8"))
    if max_ctas <= 0:
        return None
    options = dist.ProcessGroupNCCL.Options()
    options.config.max_ctas = max_ctas
    options.config.min_ctas = 1
    return options


def init_nccl(local_rank: int, max_ctas: int | None = None) -> None:
    """torch.distributed over NCCL for one process per GPU, with the communicator capped at `max_ctas` CTAs
    (default 8, CLOUDAAE_NCCL_MAX_CTAS overrides, 0 = NCCL's own choice).  NCCL picks 24 NVLS channels for the
    26 MB gradient bucket on an 8-GPU NVSwitch box, i.e. 24 SMs taken from the backward pass the allreduce runs next
    to — and the step is bound by SM work.  Measured at 8 x B200, batch 128 per GPU
    (tools/gpu_r2_dp_ab.sh): 2.002 ms/step uncapped, 1.963 ms with 8 CTAs (single GPU: 1.869 ms)."""
    options = nccl_options(max_ctas)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=options)


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
