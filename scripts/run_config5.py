"""Config 5 (scaling sweep): a batch of samples, 50 000 reads x 250 bp from 20 haplotypes each (config 1's
generator at the `useCutOff` size), qluster per sample.  Samples are sharded over ranks (rank r takes
samples r, r + N, ...: qlusterCmds.txt + runMultipleCommands, no data-path collective); each rank runs its
share through sdq_run_many with a few worker contexts on its GPU.
  python scripts/run_config5.py SAMPLES_PER_GPU [workers]              one GPU
  torchrun --nproc-per-node G ... scripts/run_config5.py SAMPLES_PER_GPU [workers]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seekdeep_b200 as sd  # noqa: E402
from tests import synth  # noqa: E402
from seekdeep_b200.qluster import qluster_many  # noqa: E402

perGpu = int(sys.argv[1]) if len(sys.argv) > 1 else 48
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
nSamples = perGpu * world
mine = sd.shard_samples(nSamples, rank, world)
samples = []
for s in mine:
    rng = np.random.default_rng(500 + s)
    haps = synth.make_haplotypes(rng, 20, 250, 1, 6)
    seqs, quals, _ = synth.fast_reads(rng, haps, 50_000, 0.005)
    samples.append(synth.pack(seqs, quals))
params = sd.make_params()
qluster_many(params, samples[:2], device=local, nWorkers=workers, want_clusters=False)  # warm-up
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res = qluster_many(params, samples, device=local, nWorkers=workers, want_clusters=False)
dt = time.perf_counter() - t0
reads = float(sum(len(s[2]) - 1 for s in samples))
cells = float(sum(r["stats"]["cellUpdates"] for r in res))
clusters = float(sum(r["stats"]["finalClusters"] for r in res))
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    v = torch.tensor([reads, cells, clusters], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(v, op=dist.ReduceOp.SUM)
    dt = t.item()
    reads, cells, clusters = v.tolist()
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps({"config": "C5: samples of 50k reads x 250 bp, 20 haplotypes, 0.5 % subs", "n_gpus": world, "samples": nSamples,
                      "workers_per_gpu": workers, "wall_s": round(dt, 3), "reads_per_s": reads / dt, "samples_per_s": nSamples / dt,
                      "gcups_effective": cells / dt / 1e9, "final_clusters": int(clusters)}), flush=True)
