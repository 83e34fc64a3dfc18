"""CPU test of the multi-GPU host logic: samples shard over ranks with no data-path collective;
only the per-rank throughput numbers are reduced.  world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seekdeep_b200.qluster import shard_samples


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nSamples, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_samples(nSamples, rank, world)
    # every rank "processes" its samples: here, counts reads with a deterministic per-sample size
    reads = torch.tensor([float(sum(1000 + 7 * s for s in mine)), float(len(mine))], dtype=torch.float64)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(reads, op=dist.ReduceOp.SUM)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((reads.tolist(), ms.item(), gathered))
    dist.destroy_process_group()


def test_sample_sharding_two_ranks():
    world, nSamples = 2, 9
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nSamples, out)) for r in range(world)]
    for p in procs:
        p.start()
    reads, ms, gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allSamples = sorted(s for part in gathered for s in part)
    assert allSamples == list(range(nSamples))  # every sample exactly once
    assert set(gathered[0]).isdisjoint(gathered[1]) and abs(len(gathered[0]) - len(gathered[1])) <= 1
    assert reads[0] == sum(1000 + 7 * s for s in range(nSamples)) and reads[1] == nSamples
    assert ms == 11.0  # max over ranks is what bench.py reports


def test_shard_edge_cases():
    assert shard_samples(0, 0, 4) == []
    assert shard_samples(3, 3, 4) == []
    assert sorted(sum((shard_samples(10, r, 4) for r in range(4)), [])) == list(range(10))
