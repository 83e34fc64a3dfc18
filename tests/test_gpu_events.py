"""sdgpu_align_events: comp_.distances_.mismatches_ / lowKmerMismatches_ / alignmentGaps_ (the maps qluster's
--writeOutFinalAllByAllComparison dumps, clusterDown.cpp:759-800) as event lists, identical to the oracle's walk over
the gapped strings: kinds, columns, positions, bases, qualities, k-mer counts."""
import numpy as np
import pytest

from seekdeep_b200 import _abi
from seekdeep_b200.aligner import GpuAligner
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _check(oracle, params, seqs, quals, bases, q, offsets, cnts, refIdx, readIdx, flags, kIndex):
    km = None
    with GpuAligner(params) as al:
        al.upload_seqs(bases, q, offsets, cnts)
        if kIndex:
            al.index_kmers(*kIndex)
            km = oracle.index_kmers(bases, offsets, cnts, *kIndex)
        try:
            got, events = al.align_events(refIdx, readIdx, checkKmer=bool(flags & 1), usingQuality=bool(flags & 2),
                                          doingMatchQuality=bool(flags & 4), eventCap=6)  # small: the wrapper has to grow it
            plain = al.align_profile(refIdx, readIdx, checkKmer=bool(flags & 1), usingQuality=bool(flags & 2),
                                     doingMatchQuality=bool(flags & 4))
            assert np.array_equal(got, plain)  # the scalar part is sdgpu_align_profile's
            nGapEv = 0
            for i, (a, b) in enumerate(zip(refIdx, readIdx)):
                wc, wev = oracle.align_events(params, seqs[a], quals[a], seqs[b], quals[b], flags, kmaps=km)
                assert len(events[i]) == len(wev), (i, len(events[i]), len(wev))
                assert np.array_equal(events[i], wev), (i, events[i], wev)
                for name in _abi.INT_FIELDS:
                    assert got[name][i] == wc[name], (name, i)
                nGapEv += int((wev["kind"] >= _abi.EV_GAP_IN_REF).sum())
        finally:
            if km:
                oracle.free_kmaps(km)
    return nGapEv


@pytest.mark.parametrize("pname", ["qluster", "asym", "countEnds"])
def test_align_events_match_oracle(oracle, pname):
    rng = np.random.default_rng(len(pname) * 7 + 1)
    params = H.params_for(pname)
    bases, q, offsets, cnts, seqs, quals = H.make_store(rng, 40, 180)
    n = 300
    refIdx = rng.integers(0, 40, n).astype(np.uint32)
    readIdx = rng.integers(0, 40, n).astype(np.uint32)
    nGaps = _check(oracle, params, seqs, quals, bases, q, offsets, cnts, refIdx, readIdx, _abi.USING_QUALITY | _abi.CHECK_KMER, (9, 1, True))
    assert nGaps > 0
    _check(oracle, params, seqs, quals, bases, q, offsets, cnts, refIdx[:100], readIdx[:100], _abi.USING_QUALITY, None)
    _check(oracle, params, seqs, quals, bases, q, offsets, cnts, refIdx[:100], readIdx[:100], _abi.CHECK_KMER, (7, 2, False))


def test_align_events_unrelated_and_ragged(oracle):
    """Unrelated sequences of very different lengths: dozens of gap runs per pair (the first, 16-run gap list is too
    small and the call sizes a second one), end gaps on both sides, sequences of one and two bases."""
    rng = np.random.default_rng(3)
    seqs = [H.rand_seq(rng, int(n)) for n in rng.integers(1, 260, 32)]
    seqs[0] = H.rand_seq(rng, 1)
    seqs[1] = H.rand_seq(rng, 2)
    quals = [H.rand_quals(rng, len(s)) for s in seqs]
    bases, q, offsets = np.concatenate(seqs), np.concatenate(quals), np.zeros(33, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    refIdx, readIdx = np.meshgrid(np.arange(32, dtype=np.uint32), np.arange(32, dtype=np.uint32))
    nGaps = _check(oracle, H.params_for("qluster"), seqs, quals, bases, q, offsets, None, refIdx.ravel(), readIdx.ravel(),
                   _abi.USING_QUALITY, None)
    assert nGaps > 16 * 32


def test_align_gap_lists_match_align_profile(oracle):
    """sdgpu_align_gap_lists (what the consensus rebuild asks for): packed records only for the pairs with gap runs, equal
    to the gap lists sdgpu_align_profile returns densely; lists longer than gapCap are flagged by their full run count."""
    rng = np.random.default_rng(19)
    bases, q, offsets, cnts, seqs, quals = H.make_store(rng, 60, 200)
    n = 3000
    refIdx = rng.integers(0, 60, n).astype(np.uint32)
    readIdx = rng.integers(0, 60, n).astype(np.uint32)
    with GpuAligner(H.params_for("qluster")) as al:
        al.upload_seqs(bases, q, offsets, cnts)
        for gapCap in (2, 8):
            dense, gaps = al.align_profile(refIdx, readIdx, usingQuality=False, gapCap=gapCap)
            packed = al.align_gap_lists(refIdx, readIdx, gapCap=gapCap)
            assert set(packed) == set(np.nonzero(dense["nGapRuns"])[0].tolist())
            for p, (nRuns, g) in packed.items():
                assert nRuns == dense["nGapRuns"][p]
                assert np.array_equal(g, gaps[p, :min(nRuns, gapCap)])
            if gapCap == 2:
                assert any(nRuns > 2 for nRuns, _ in packed.values())
        assert al.align_gap_lists(refIdx[:0], readIdx[:0]) == {}



def test_cxx_all_by_all_example_matches_python_api():
    """examples/all_by_all_cabi.cpp (the --writeOutFinalAllByAllComparison tables of clusterDown.cpp:727-805 from one
    sdgpu_align_events call, C++ straight onto the C ABI) prints what the Python wrapper returns for the same clusters."""
    import subprocess
    import __graft_entry__ as g
    exe = g.build_example("all_by_all_cabi")
    rng = np.random.default_rng(12)
    root = H.rand_seq(rng, 150)
    seqs = [root]
    for i in range(5):
        s = H.mutate(rng, root, int(rng.integers(1, 5)), int(i % 2), int((i + 1) % 2), hp=False)
        seqs.append(s)
    names = [f"clus{i}" for i in range(len(seqs))]
    args = []
    for nme, s in zip(names, seqs):
        args += [nme, s.tobytes().decode()]
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    counts, events = r.stdout.split("--\n")
    counts = [l.split("\t") for l in counts.strip().splitlines()[1:]]
    events = [l.split("\t") for l in events.rstrip("\n").splitlines()[1:]]
    # the same store through the Python wrapper
    n = len(seqs)
    quals = [np.where(np.arange(len(s)) % 7 == 3, 12, 35).astype(np.uint8) for s in seqs]
    offsets = np.zeros(n + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    cnts = np.array([10.0 * (n - i) for i in range(n)])
    refIdx, readIdx = [], []
    for s in range(n - 1, -1, -1):
        for o in range(s):
            refIdx.append(o)
            readIdx.append(s)
    with GpuAligner(H.params_for("qluster")) as al:
        al.upload_seqs(np.concatenate(seqs), np.concatenate(quals), offsets, cnts)
        al.index_kmers(9, 1, True)
        comp, evs = al.align_events(refIdx, readIdx, checkKmer=True, usingQuality=True)
    assert len(counts) == len(refIdx)
    for p, row in enumerate(counts):
        assert row[0] == names[refIdx[p]] and row[1] == names[readIdx[p]] and int(row[2]) == comp["alnScore"][p]
        assert float(row[3]) == comp["eventBasedIdentity"][p] and float(row[4]) == comp["eventBasedIdentityHq"][p]
        assert float(row[5]) == comp["oneBaseIndel"][p] and float(row[6]) == comp["twoBaseIndel"][p]
        assert [int(x) for x in row[8:11]] == [comp["lqMismatches"][p], comp["hqMismatches"][p], comp["lowKmerMismatches"][p]]
    nMis = sum(int(((e["kind"] >= 1) & (e["kind"] <= 3)).sum()) for e in evs)
    nGap = sum(int(((e["kind"] == _abi.EV_GAP_IN_REF) | (e["kind"] == _abi.EV_GAP_IN_READ)).sum()) for e in evs)
    assert sum(1 for e in events if e[2] in ("mismatch", "lowKmerMismatch")) == nMis > 0
    assert sum(1 for e in events if e[2].startswith("gap-")) == nGap > 0
    assert all(len(e) == 11 for e in events)
