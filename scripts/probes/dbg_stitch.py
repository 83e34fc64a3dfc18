import numpy as np, sys
sys.path.insert(0, '/root/repo')
from seekdeep_b200 import stitch as ST
from seekdeep_b200.aligner import GpuAligner
from tests import oracle_binding
from tests.test_stitch import make_pairs
oracle = oracle_binding.load()
rng = np.random.default_rng(4)
r1, r2 = make_pairs(rng, 1500)
sp = ST.ProcessParams(r1Trim=7, r2Trim=4, windowSize=8, windowStep=2, avgQualCutOff=22.0)
pars = ST.pair_aligner_params()
with GpuAligner(pars) as al:
    got = ST.process_pairs(al, r1, r2, sp)
st, ov, seqs, retried = oracle.stitch_pairs(pars, sp, r1, r2)
off = got["offsets"]
nbad = 0
for i in range(len(st)):
    lo, hi = int(off[i]), int(off[i + 1])
    g = got["bases"][lo:hi].tobytes()
    if g != seqs[i][0]:
        nbad += 1
        if nbad <= 3:
            l1 = int(r1[2][i+1]-r1[2][i]); l2 = int(r2[2][i+1]-r2[2][i])
            print(i, "status", st[i], "shift", ov["shift"][i], "bases", ov["basesInAln"][i], "l1", l1, "l2", l2, "lenGot", len(g), "lenWant", len(seqs[i][0]))
            print(" got ", g[-60:]); print(" want", seqs[i][0][-60:])
            print(" got0", g[:60]); print(" wan0", seqs[i][0][:60])
print("bad", nbad, "of", len(st), "retried", retried)
