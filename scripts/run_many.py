"""Multi-sample throughput on one GPU: python scripts/run_many.py nSamples readsPerSample workers..."""
import json
import sys
import time

import seekdeep_b200 as sd
from tests import synth
from seekdeep_b200.qluster import qluster_many

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nR = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
workers = [int(x) for x in sys.argv[3:]] or [1, 2]
samples = []
for s in range(nS):
    haps, seqs, quals, _ = synth.config2_fast(seed=100 + s, nReads=nR) if nR != 50000 else synth.config2_fast(seed=100 + s, nReads=nR, length=250, nHaps=20)
    samples.append(synth.pack(seqs, quals))
params = sd.make_params()
qluster_many(params, samples[:1], nWorkers=1, want_clusters=False)  # warm-up (library load, first allocations)
for w in workers:
    t0 = time.perf_counter()
    res = qluster_many(params, samples, nWorkers=w, want_clusters=False)
    dt = time.perf_counter() - t0
    cells = sum(r["stats"]["cellUpdates"] for r in res)
    print(json.dumps({"samples": nS, "reads_per_sample": nR, "workers": w, "wall_s": round(dt, 3), "reads_per_s": nS * nR / dt,
                      "gcups": cells / dt / 1e9}), flush=True)
