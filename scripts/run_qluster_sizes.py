"""Exploration helper: run the GPU qluster on C1 / C2-shaped samples of growing size and print stats."""
import json
import sys
import time

import numpy as np

import seekdeep_b200 as sd
from tests import synth

sizes = [int(x) for x in sys.argv[1:]] or [10000]
for n in sizes:
    t0 = time.time()
    if n == 10000:
        haps, seqs, quals, _ = synth.config1()
    else:
        haps, seqs, quals, _ = synth.config2_fast(nReads=n)
    bases, q, offsets = synth.pack(seqs, quals)
    tg = time.time() - t0
    with sd.GpuAligner(sd.make_params()) as al:
        t0 = time.time()
        res = sd.qluster(al, bases, q, offsets, want_clusters=True)
        dt = time.time() - t0
    st = res["stats"]
    hs = {h.tobytes().decode() for h in haps}
    top = [(int(c), s in hs) for s, _, c in res["clusters"][:8]]
    print(json.dumps({"reads": n, "gen_s": round(tg, 1), "wall_s": round(dt, 2), "reads_per_s": round(n / dt, 1),
                      "gcups_overall": round(st["cellUpdates"] / dt / 1e9, 1), **st, "top": top}), flush=True)
