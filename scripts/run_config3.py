"""Config 3 (all-vs-all of unique 400-bp haplotypes) timing helper: python scripts/run_config3.py N"""
import json
import sys
import time

import numpy as np

import seekdeep_b200 as sd
from seekdeep_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
seqs, quals = synth.config3(n=n)
bases, q, offsets = synth.pack(seqs, quals)
lens = np.diff(offsets).astype(np.int64)
with sd.GpuAligner(sd.make_params()) as al:
    al.upload_seqs(bases, q, offsets)
    al.all_vs_all(nPairs=min(100000, n * (n - 1) // 2))  # warm-up
    al.profile_enable(True)
    al.profile_read(reset=True)
    t0 = time.perf_counter()
    out = al.all_vs_all()
    dt = time.perf_counter() - t0
    prof = al.profile_read()
print(json.dumps({"n": n, "pairs": len(out), "wall_s": round(dt, 3), "pairs_per_s": len(out) / dt,
                  "gcups_wall": prof["cellUpdates"] / dt / 1e9, "gcups_kernel": prof["cellUpdates"] / (prof["alignKernelMs"] * 1e-3) / 1e9,
                  "d2h_bytes": prof["d2hBytes"], "mean_score": float(out["alnScore"].mean())}))
