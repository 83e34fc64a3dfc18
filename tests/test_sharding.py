"""CPU test of the multi-GPU host logic: samples shard over ranks with no data-path collective;
only the per-rank throughput numbers are reduced.  world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seekdeep_b200.qluster import shard_pair_blocks, shard_samples


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


def _pair_worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard_pair_blocks(total, rank, world)
    # stand-in for this rank's sdgpu_all_vs_all block: a checksum over its pair indices
    mine = torch.arange(first, first + n, dtype=torch.int64)
    acc = torch.stack([mine.sum(), torch.tensor(n, dtype=torch.int64)])
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    blocks = [None] * world
    dist.all_gather_object(blocks, (first, n))
    if rank == 0:
        out.put((acc.tolist(), blocks))
    dist.destroy_process_group()


def test_pair_block_sharding_two_ranks():
    """Config 3's pair-block partition over 2 ranks: disjoint contiguous blocks that cover the
    upper triangle exactly once."""
    world, nSeqs = 2, 301
    total = nSeqs * (nSeqs - 1) // 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pair_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    acc, blocks = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert acc == [total * (total - 1) // 2, total]
    assert blocks[0][0] == 0 and blocks[0][0] + blocks[0][1] == blocks[1][0] and blocks[1][0] + blocks[1][1] == total


def test_pair_block_edge_cases():
    for total in (0, 1, 7, 8, 199990000):
        for world in (1, 2, 4, 8):
            blocks = [shard_pair_blocks(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
