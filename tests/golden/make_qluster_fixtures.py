"""Freezes the oracle's qluster results at BASELINE sizes into tests/golden/qluster_*.npz so the
GPU parity tests can compare sd.qluster bit for bit at sizes the oracle cannot run inside a test
(config 1 at full size: ~2 min of one core; the C2 generator shrunk to 20 000 reads: ~15 min).

Run once in the authoring container:  python tests/golden/make_qluster_fixtures.py [name ...]
Inputs are regenerated from the seeded generators at test time; the fixture carries a SHA-256 of
the packed input so a generator drift is reported as such, not as a parity failure.

Stored per fixture: inputSha, membership (uint32 per input read), cluster counts, consensus
sequences and qualities (concatenated + offsets), chimera records, the oracle's unique count."""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import seekdeep_b200 as sd  # noqa: E402
from tests import synth  # noqa: E402
from seekdeep_b200.iterpars import CollapseIterations  # noqa: E402
from tests import oracle_binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case_inputs(name):
    """(bases, quals, offsets, params, pars, iters) of a named fixture; shared with the tests."""
    if name == "c1_full":  # BASELINE config 1: 10k reads x 250 bp, 20 haplotypes, --illumina defaults
        _, seqs, quals, _ = synth.config1()
        weigh = False
    elif name == "c2_20k":  # the C2 generator (200 haplotypes x 300 bp, 1 % subs + homopolymer indels) at 20k reads
        _, seqs, quals, _ = synth.config2_fast(nReads=20_000)
        weigh = False
    elif name == "c2_20k_hp":  # same reads with --illuminaAllowHomopolyers (fractional indel weights)
        _, seqs, quals, _ = synth.config2_fast(nReads=20_000)
        weigh = True
    elif name == "c2_5k":  # small sibling of the above for quick runs
        _, seqs, quals, _ = synth.config2_fast(nReads=5_000, nHaps=60)
        weigh = False
    else:
        raise KeyError(name)
    bases, q, offsets = synth.pack(seqs, quals)
    params = sd.make_params(weighHomopolymers=weigh)
    iters = CollapseIterations.gen_illumina_default_pars(100)
    return bases, q, offsets, params, sd.QlusterPars(), iters


def input_sha(bases, quals, offsets):
    h = hashlib.sha256()
    for a in (bases, quals, offsets):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def pack_result(res):
    seqs = [np.frombuffer(s.encode(), np.uint8) for s, _, _ in res["clusters"]]
    quals = [np.asarray(q, np.uint8) for _, q, _ in res["clusters"]]
    off = np.zeros(len(seqs) + 1, np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    chi = np.array([(c,) + tuple(v) for c, v in sorted(res["chimeras"].items())], np.uint32).reshape(-1, 5)
    return dict(membership=np.asarray(res["membership"], np.uint32),
                cnts=np.array([c for _, _, c in res["clusters"]], np.float64),
                seqs=np.concatenate(seqs) if seqs else np.zeros(0, np.uint8),
                quals=np.concatenate(quals) if quals else np.zeros(0, np.uint8), offsets=off, chimeras=chi)


def main(names):
    oracle = oracle_binding.load()
    for name in names:
        bases, q, offsets, params, pars, iters = case_inputs(name)
        t0 = time.time()
        want = oracle.qluster(params, pars.to_c(), iters.iters, iters.iters, bases, q, offsets)
        dt = time.time() - t0
        out = pack_result(want)
        out["inputSha"] = np.frombuffer(input_sha(bases, q, offsets).encode(), np.uint8)
        out["nUnique"] = np.array([want["nUnique"]], np.uint64)
        out["nAlign"] = np.array([want["nAlign"]], np.uint64)
        path = os.path.join(HERE, f"qluster_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(offsets) - 1} reads, {want['nUnique']} unique, {len(want['clusters'])} clusters, "
              f"{want['nAlign']} alignments, oracle {dt:.1f} s -> {path} ({os.path.getsize(path)} bytes)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c2_5k", "c1_full", "c2_20k", "c2_20k_hp"])
