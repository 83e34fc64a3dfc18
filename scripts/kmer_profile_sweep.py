"""Sweep of the k = 5 profile build's launch knobs (SD_PROF_STORE / SD_PROF_CTAS / SD_PROF_WARPS): one process per setting
(the knobs are read once per process).  Prints GB/s of algorithmic bytes on the C2 sample."""
import os
import subprocess
import sys

CHILD = r'''
import numpy as np, seekdeep_b200 as sd
from seekdeep_b200 import synth
haps, seqs, quals, _ = synth.config2_fast(nReads=200000)
bases, q, offsets = synth.pack(seqs, quals)
lens = np.diff(offsets).astype(np.int64)
with sd.GpuAligner(sd.make_params()) as al:
    al.upload_seqs(bases, q, offsets)
    out = []
    for k in (5, 6):
        al.build_kmer_profiles(k); al.sync()
        ts = []
        for _ in range(9):
            al.timer_start(); al.build_kmer_profiles(k); ts.append(al.timer_stop())
        algo = float(lens.sum() + len(lens) * (16 + 4 ** k * 2))
        out.append(f"k={k} best {min(ts):.4f} ms median {sorted(ts)[4]:.4f} ms {algo / min(ts) / 1e6:.0f} GB/s")
    print(" | ".join(out))
'''
if __name__ == "__main__":
    for store, ctas, warps in ((1, 32, 0), (0, 32, 0), (1, 64, 0), (1, 32, 4)):
        if True:
            if True:
                env = dict(os.environ, SD_PROF_STORE=str(store), SD_PROF_CTAS=str(ctas), SD_PROF_WARPS=str(warps))
                r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
                print(f"store={store} ctas/SM={ctas} warps={warps or 8}: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
