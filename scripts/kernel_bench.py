"""Kernel-only microbenchmark: one collapse-pass pair batch resident in HBM, int32 vs packed fp16 forward
pass, with / without the k-mer check.  python scripts/kernel_bench.py [reads] [only_mode] [only_flags]"""
import sys
import time

import numpy as np

import seekdeep_b200 as sd
from seekdeep_b200 import _abi
from tests import synth

nReads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
onlyMode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
onlyFlags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
haps, seqs, quals, _ = synth.config2_fast(nReads=nReads)
uniq = {}
for s, q in zip(seqs, quals):
    uniq.setdefault(s.tobytes(), [q, 0])[1] += 1
items = sorted(uniq.items(), key=lambda kv: (-kv[1][1], kv[0]))
useqs = [np.frombuffer(k, np.uint8) for k, _ in items]
uquals = [v[0] for _, v in items]
cnts = np.array([v[1] for _, v in items], np.float64)
bases, q, offsets = synth.pack(useqs, uquals)
nU = len(useqs)
refIdx = np.tile(np.arange(100, dtype=np.uint32), nU - 100)
readIdx = np.repeat(np.arange(100, nU, dtype=np.uint32), 100)
lens = np.diff(offsets).astype(np.int64)
cells = int((lens[refIdx] * lens[readIdx]).sum())
n = len(refIdx)
with sd.GpuAligner(sd.make_params()) as al:
    al.upload_seqs(bases, q, offsets, cnts)
    al.index_kmers(9, 1, True)
    dRef, dRead, dOut = al.dev_alloc(n * 4), al.dev_alloc(n * 4), al.dev_alloc(n * 104)
    al.h2d(dRef, refIdx)
    al.h2d(dRead, readIdx)
    for mode in ((onlyMode,) if onlyMode else (1, 2)):
        al.set_kernel(mode)
        for flags in ((onlyFlags,) if onlyFlags else (2, 3)):
            al.align_profile_dev(dRef, dRead, n, dOut, flags)
            al.sync()
            al.timer_start()
            for _ in range(3):
                al.align_profile_dev(dRef, dRead, n, dOut, flags)
            ms = al.timer_stop() / 3
            print(f"mode {mode} ({'int32' if mode == 1 else 'fp16x2'}) flags {flags} ({'checkKmer' if flags & 1 else 'no k-mer check'}): "
                  f"{ms:.1f} ms  {cells / ms / 1e6:.0f} GCUPS  {n / ms / 1e3:.2f} Mpairs/s", flush=True)
