"""Config 3 (all-vs-all of unique 400-bp haplotypes) timing helper.
  python scripts/run_config3.py N                       one GPU
  torchrun --nproc-per-node G ... scripts/run_config3.py N    pair blocks sharded over G GPUs
Every rank holds all sequences and computes its own contiguous block of the (i < j) pair list
(sdgpu_all_vs_all firstPair / nPairs); results stay on the rank, only timings are reduced."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seekdeep_b200 as sd  # noqa: E402
from tests import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
seqs, quals = synth.config3(n=n)
bases, q, offsets = synth.pack(seqs, quals)
total = n * (n - 1) // 2
first, mine = sd.shard_pair_blocks(total, rank, world)
with sd.GpuAligner(sd.make_params(), device=local) as al:
    al.upload_seqs(bases, q, offsets)
    al.all_vs_all(firstPair=first, nPairs=min(100000, mine))  # warm-up
    al.profile_enable(True)
    al.profile_read(reset=True)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = al.all_vs_all(firstPair=first, nPairs=mine)
    dt = time.perf_counter() - t0
    prof = al.profile_read()
cells, kms, score = float(prof["cellUpdates"]), float(prof["alignKernelMs"]), float(out["alnScore"].astype(np.int64).sum())
if world > 1:
    t = torch.tensor([dt, kms], dtype=torch.float64, device="cuda")
    s = torch.tensor([cells, float(len(out)), score], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dt, kms = t.tolist()
    cells, npairs, score = s.tolist()
    dist.destroy_process_group()
else:
    npairs = float(len(out))
if rank == 0:
    print(json.dumps({"n": n, "n_gpus": world, "pairs": int(npairs), "wall_s": round(dt, 3), "pairs_per_s": npairs / dt,
                      "gcups_wall": cells / dt / 1e9, "gcups_kernel_per_gpu": cells / world / (kms * 1e-3) / 1e9,
                      "score_checksum": int(score)}))
