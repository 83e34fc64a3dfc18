"""k-mer kernels microbenchmark (device-resident): profile build + reads x clusters compare grid."""
import sys
import numpy as np
import seekdeep_b200 as sd
from tests import synth

nReads = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
haps, seqs, quals, _ = synth.config2_fast(nReads=nReads)
bases, q, offsets = synth.pack(seqs, quals)
lens = np.diff(offsets).astype(np.int64)
with sd.GpuAligner(sd.make_params()) as al:
    al.upload_seqs(bases, q, offsets)
    for k in (4, 5, 6):
        al.build_kmer_profiles(k)
        al.sync()
        best = 1e9
        for _ in range(5):
            al.timer_start()
            al.build_kmer_profiles(k)
            best = min(best, al.timer_stop())
        P = 4 ** k * 2
        algo = float(lens.sum() + len(lens) * (16 + P))
        print(f"k={k} profile build: {best:.3f} ms  {algo / best / 1e6:.0f} GB/s algorithmic")
        nR, nC = len(lens), 100
        dR, dC = al.dev_alloc(nR * 4), al.dev_alloc(nC * 4)
        dS, dF = al.dev_alloc(nR * nC * 4), al.dev_alloc(nR * nC * 8)
        al.h2d(dR, np.arange(nR, dtype=np.uint32))
        al.h2d(dC, np.arange(nC, dtype=np.uint32))
        bestc = 1e9
        for _ in range(4):
            al.timer_start()
            al._check(al._lib.sdgpu_kmer_compare_all_dev(al._ctx, dR, nR, dC, nC, dS, dF))
            bestc = min(bestc, al.timer_stop())
        algoc = float((nR + nC) * P + nR * nC * 12)
        print(f"k={k} compare {nR} x {nC}: {bestc:.3f} ms  {nR * nC / bestc / 1e6:.2f} Gpairs/s  {algoc / bestc / 1e6:.0f} GB/s algorithmic")
        for p in (dR, dC, dS, dF):
            al.dev_free(p)
