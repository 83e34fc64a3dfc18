"""C2 with --illuminaAllowHomopolyers (weighHomopolymers on: fractional homopolymer-weighted indel tallies,
clusterDownSetUp.cpp:270-280): one sdq_run of the bench's sample, timing + a few result properties."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seekdeep_b200 as sd  # noqa: E402
from tests import synth  # noqa: E402

nReads = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
haps, seqs, quals, which = synth.config2_fast(nReads=nReads)
bases, q, offsets = synth.pack(seqs, quals)
out = {}
for weigh in (False, True):
    with sd.GpuAligner(sd.make_params(weighHomopolymers=weigh)) as al:
        sd.qluster(al, bases, q, offsets, want_clusters=False)  # warm-up
        t0 = time.perf_counter()
        res = sd.qluster(al, bases, q, offsets, want_clusters=True)
        dt = time.perf_counter() - t0
    st = res["stats"]
    hapSet = {bytes(h).decode() for h in haps}
    top = [c for c in res["clusters"] if c[2] >= 20]
    out["weighHomopolymers" if weigh else "default"] = {
        "wall_s": round(dt, 3), "reads_per_s": nReads / dt, "final_clusters": st["finalClusters"],
        "clusters_with_20+_reads": len(top), "of_which_exact_haplotypes": sum(c[0] in hapSet for c in top),
        "pairs_computed": st["pairsComputed"], "chimeric": st["chimericClusters"]}
print(json.dumps(out))
