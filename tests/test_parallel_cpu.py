"""Host-side multi-GPU logic on CPU: world_size-2 gloo process groups (rendezvous on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cloudaae_b200.parallel import BucketedAllReduce, broadcast_variables, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 128, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))       # contiguous, no overlap
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1                                      # balanced
    assert shard_range(4096, 3, 8) == (1536, 2048)                                   # BASELINE config 5: 512 per GPU
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # identical initial state
        flat = torch.full((1000,), float(rank + 1))
        ema = torch.full((10,), float(rank + 5))
        broadcast_variables(flat, ema)
        assert (flat == 1).all() and (ema == 5).all()
        # bucketed gradient exchange: sum over ranks, buckets started out of order
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        red = BucketedAllReduce(g, [0, 320, 1000])
        red.start(1)
        assert (g[:320] == torch.arange(320) * (rank + 1)).all()  # bucket 0 untouched so far
        red.start(0)
        red.finish()
        want = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(g, want)
        # data-parallel step semantics: averaged gradient == gradient of the concatenated batch
        torch.manual_seed(0)
        w = torch.randn(5, 3, dtype=torch.float64)
        xs = torch.randn(world, 4, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
        x = xs[rank]
        grad_local = 2 * x.T @ (x @ w) / x.shape[0]
        buf = grad_local.clone().reshape(-1)
        r2 = BucketedAllReduce(buf, [0, buf.numel()])
        r2.start(0); r2.finish()
        full = xs.reshape(-1, 5)
        want = 2 * full.T @ (full @ w) / full.shape[0]
        assert torch.allclose(buf.reshape(5, 3) / world, want)
        # sharded inference needs no collective: every rank handles a disjoint slice
        a, b = shard_range(11, rank, world)
        got = torch.zeros(11)
        got[a:b] = 1
        dist.all_reduce(got)
        assert (got == 1).all()
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_nccl_options_cap_the_communicator_ctas(monkeypatch):
    """The data-parallel step is bound by SM work, so the NCCL communicator is capped at a few CTAs: default 8,
    CLOUDAAE_NCCL_MAX_CTAS overrides, 0 leaves the choice to NCCL."""
    from cloudaae_b200 import parallel
    if not hasattr(torch.distributed, "ProcessGroupNCCL"):
        pytest.skip("torch built without NCCL")
    monkeypatch.delenv("CLOUDAAE_NCCL_MAX_CTAS", raising=False)
    o = parallel.nccl_options()
    assert o.config.max_ctas == 8 and o.config.min_ctas == 1
    monkeypatch.setenv("CLOUDAAE_NCCL_MAX_CTAS", "4")
    assert parallel.nccl_options().config.max_ctas == 4
    monkeypatch.setenv("CLOUDAAE_NCCL_MAX_CTAS", "0")
    assert parallel.nccl_options() is None
    assert parallel.nccl_options(16).config.max_ctas == 16
