"""CPU tests of the oracle itself: hand-derived micro-cases for every tie / boundary rule,
optimality against an independent textbook DP, and traceback self-consistency.
(Parity to upstream njhseq is unpinned: the reference ships no golden vectors — DESIGN.md.)"""
import numpy as np
import pytest

from seekdeep_b200 import _abi
from seekdeep_b200.scoring import GapScoringParameters, SubstituteMatrix, make_params
from tests import helpers as H


def test_identical(oracle):
    p = make_params()
    s = "ACGTTGCAAGGCTTAACCGG"
    q = np.full(len(s), 38, np.uint8)
    comp, gaps, a, b = oracle.align_profile(p, s, q, s, q)
    assert comp.alnScore == 2 * len(s) and len(gaps) == 0 and a == s and b == s
    assert comp.highQualityMatches == len(s) and comp.hqMismatches == 0 and comp.numGaps == 0
    assert comp.eventBasedIdentity == 1.0 and comp.refCoverage == 1.0 and comp.queryCoverage == 1.0


def test_tie_prefers_gap_over_diagonal_all_costly(oracle):
    # A="AA", B="A", every gap 5/1: both placements score -3; needleMaximum picks 'U' at the
    # corner (u == d > l), so the read's base pairs with A[0] and the gap trails.
    p = make_params(gaps=GapScoringParameters.from_strings("5,1", "5,1", "5,1"))
    score, gaps = oracle.align_global(p, "AA", "A")
    assert score == -3
    assert [(g["pos"], g["size"], g["gapInA"]) for g in gaps] == [(1, 1, 0)]


def test_free_right_gap(oracle):
    p = make_params()  # qluster: right gaps free
    comp, gaps, a, b = oracle.align_profile(p, "AA", None, "A", None, flags=0)
    assert comp.alnScore == 2 and a == "AA" and b == "A-"
    assert comp.numGaps == 0 and comp.oneBaseIndel == 0  # end gap not counted
    assert comp.queryCoverage == 1.0 and comp.refCoverage == 0.5


def test_internal_one_base_deletion(oracle):
    p = make_params()
    ref = "ACGTTGCAAGGCTTAACCGGATCG"
    read = ref[:10] + ref[11:]
    comp, gaps, a, b = oracle.align_profile(p, ref, None, read, None, flags=0)
    assert comp.alnScore == 2 * len(read) - 5
    assert comp.oneBaseIndel == 1 and comp.twoBaseIndel == 0 and comp.largeBaseIndel == 0 and comp.numGaps == 1
    assert a == ref and b.replace("-", "") == read and b.count("-") == 1
    assert len(gaps) == 1 and gaps[0]["gapInA"] == 0 and gaps[0]["size"] == 1


def test_indel_size_classes(oracle):
    p = make_params()
    rng = np.random.default_rng(5)
    ref = H.rand_seq(rng, 60).tobytes().decode()
    for size, field in ((1, "oneBaseIndel"), (2, "twoBaseIndel"), (3, "largeBaseIndel"), (5, "largeBaseIndel")):
        read = ref[:30] + ref[30 + size:]
        comp, gaps, a, b = oracle.align_profile(p, ref, None, read, None, flags=0)
        assert getattr(comp, field) == 1, (size, a, b)
        assert comp.oneBaseIndel + comp.twoBaseIndel + comp.largeBaseIndel == 1


def test_quality_classes(oracle):
    p = make_params(primaryQual=25, secondaryQual=20, qualThresWindow=2)
    ref = "ACGTTGCAAGGCTTAACCGGATCG"
    read = ref[:12] + "A" + ref[13:]  # T->A mismatch at 12
    assert read != ref
    hi = np.full(len(ref), 38, np.uint8)
    comp, *_ = oracle.align_profile(p, ref, hi, read, hi)
    assert (comp.hqMismatches, comp.lqMismatches) == (1, 0)
    q = hi.copy(); q[12] = 25  # not strictly above primary
    comp, *_ = oracle.align_profile(p, ref, hi, read, q)
    assert (comp.hqMismatches, comp.lqMismatches) == (0, 1)
    q = hi.copy(); q[12] = 26
    comp, *_ = oracle.align_profile(p, ref, hi, read, q)
    assert (comp.hqMismatches, comp.lqMismatches) == (1, 0)
    q = hi.copy(); q[10] = 20  # neighbour inside the window, not strictly above secondary
    comp, *_ = oracle.align_profile(p, ref, hi, read, q)
    assert (comp.hqMismatches, comp.lqMismatches) == (0, 1)
    q = hi.copy(); q[9] = 2  # outside the window
    comp, *_ = oracle.align_profile(p, ref, hi, read, q)
    assert (comp.hqMismatches, comp.lqMismatches) == (1, 0)
    q = hi.copy(); q[14] = 20  # trailing neighbour
    comp, *_ = oracle.align_profile(p, ref, q, read, hi)
    assert (comp.hqMismatches, comp.lqMismatches) == (0, 1)
    comp, *_ = oracle.align_profile(p, ref, q, read, hi, flags=0)  # not using quality: all HQ
    assert (comp.hqMismatches, comp.lqMismatches) == (1, 0)


def test_homopolymer_weighting(oracle):
    ref = "ACGTCAAAAAGTCCGTAGGCT"
    read = "ACGTCAAAAGTCCGTAGGCT"  # one A fewer in a 5-run
    hi = np.full(30, 38, np.uint8)
    p = make_params(weighHomopolymers=False)
    comp, *_ = oracle.align_profile(p, ref, hi[:len(ref)], read, hi[:len(read)])
    assert comp.oneBaseIndel == 1.0
    p = make_params(weighHomopolymers=True)
    comp, gaps, a, b = oracle.align_profile(p, ref, hi[:len(ref)], read, hi[:len(read)])
    # run of 5 vs run of 4: |5-4| / (5+4)
    assert comp.oneBaseIndel == pytest.approx(1.0 / 9.0, abs=0)
    assert comp.numGaps == 1


def test_low_kmer_mismatch(oracle):
    rng = np.random.default_rng(9)
    ref = H.rand_seq(rng, 80)
    read = ref.copy()
    read[40] = H.BASES[(np.nonzero(H.BASES == ref[40])[0][0] + 1) % 4]
    bases = np.concatenate([ref, read])
    offsets = np.array([0, 80, 160], np.uint64)
    cnts = np.array([10.0, 1.0])
    km = oracle.index_kmers(bases, offsets, cnts, 9, 1, byPos=True)
    try:
        hi = np.full(80, 38, np.uint8)
        p = make_params()
        comp, *_ = oracle.align_profile(p, ref.tobytes(), hi, read.tobytes(), hi,
                                        flags=_abi.USING_QUALITY | _abi.CHECK_KMER, kmaps=km)
        # the read's k-mer around the mismatch occurs once (count 1 <= runCutOff 1)
        assert (comp.hqMismatches, comp.lowKmerMismatches) == (0, 1)
        comp, *_ = oracle.align_profile(p, ref.tobytes(), hi, read.tobytes(), hi, flags=_abi.USING_QUALITY, kmaps=km)
        assert (comp.hqMismatches, comp.lowKmerMismatches) == (1, 0)
        assert oracle.kmaps_count(km, ref[:9].tobytes(), 0) == 11
        assert oracle.kmaps_count(km, read[36:45].tobytes(), 36) == 1
    finally:
        oracle.free_kmaps(km)


def test_kmer_compare(oracle):
    shared, frac = oracle.kmer_compare("AAAAAAA", "AAAAA", 3)
    assert shared == 3 and frac == 1.0  # min(5,3) shared 3-mers / (5-3+1)
    shared, frac = oracle.kmer_compare("ACGTACGT", "TTTTTTTT", 3)
    assert shared == 0 and frac == 0.0
    shared, frac = oracle.kmer_compare("ACGTACGTAA", "ACGTACGTCC", 5)
    assert shared == 4 and frac == pytest.approx(4 / 6)


@pytest.mark.parametrize("pname", sorted(H.PARAM_SETS))
def test_score_optimal_and_traceback_consistent(oracle, pname):
    """Oracle score == independent textbook DP; the gapped strings it reports score the same."""
    p = H.params_for(pname)
    gp = H.PARAM_SETS[pname]["gaps"]
    mat = np.ctypeslib.as_array(p.scoreMat).reshape(128, 128)
    rng = np.random.default_rng(hash(pname) % 2**32)
    for trial in range(40):
        la = int(rng.integers(1, 40))
        a = H.rand_seq(rng, la)
        if trial % 2:
            b = H.mutate(rng, a, int(rng.integers(0, 4)), int(rng.integers(0, 3)), int(rng.integers(0, 3)))
        else:
            b = H.rand_seq(rng, int(rng.integers(1, 40)))
        if len(b) == 0:
            continue
        comp, gaps, alnA, alnB = oracle.align_profile(p, a.tobytes(), None, b.tobytes(), None, flags=0)
        want = H.gotoh_score(list(a), list(b), gp, mat)
        assert comp.alnScore == want, (pname, trial, a.tobytes(), b.tobytes())
        assert alnA.replace("-", "").encode() == a.tobytes() and alnB.replace("-", "").encode() == b.tobytes()
        assert len(alnA) == len(alnB) == comp.alnLen
        assert all(not (x == "-" and y == "-") for x, y in zip(alnA, alnB))
        assert H.score_of_alignment(alnA, alnB, gp, mat) == comp.alnScore, (alnA, alnB)


def test_gap_list_matches_gapped_strings(oracle):
    p = make_params()
    rng = np.random.default_rng(77)
    for _ in range(50):
        a = H.rand_seq(rng, int(rng.integers(20, 60)))
        b = H.mutate(rng, a, 2, 1, 1)
        comp, gaps, alnA, alnB = oracle.align_profile(p, a.tobytes(), None, b.tobytes(), None, flags=0)
        sa, sb = a.tobytes().decode(), b.tobytes().decode()
        for g in gaps:  # descending positions: inserting in list order never shifts later ones
            if g["gapInA"]:
                sa = sa[:g["pos"]] + "-" * int(g["size"]) + sa[g["pos"]:]
            else:
                sb = sb[:g["pos"]] + "-" * int(g["size"]) + sb[g["pos"]:]
        assert (sa, sb) == (alnA, alnB)
        assert comp.nGapRuns == len(gaps)


def test_identity_fields(oracle):
    p = make_params()
    ref = "ACGTTGCAAGGCTTAACCGGATCGTTAGC"
    read = ref[:8] + "T" + ref[9:20] + ref[21:]
    hi = np.full(len(ref), 38, np.uint8)
    comp, *_ = oracle.align_profile(p, ref, hi, read, hi[:len(read)])
    ev = comp.highQualityMatches + comp.hqMismatches + comp.lqMismatches + comp.lowKmerMismatches + comp.numGaps
    assert comp.overLappingEvents == ev
    assert comp.eventBasedIdentity == comp.highQualityMatches / ev
    assert comp.basesInAln == len(read)


def test_mark_chimeras_crossover(oracle):
    """Two abundant parents, their crossover and a one-SNP variant: only the crossover is flagged,
    with (front parent, back parent) and crossover mismatch positions that bracket the cut."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations
    rng = np.random.default_rng(11)
    bases, quals, offsets, (p0, p1, chi, var), cut = H.chimera_sample(rng)
    its = CollapseIterations.gen_illumina_default_pars(100)
    pars = sd.QlusterPars()
    res = oracle.qluster(sd.make_params(), pars.to_c(), its.iters, its.iters, bases, quals, offsets)
    seqs = [c[0] for c in res["clusters"]]
    assert seqs == [bytes(s).decode() for s in (p0, p1, chi, var)]
    assert list(res["chimeras"]) == [2]
    front, back, fpos, bpos = res["chimeras"][2]
    assert (front, back) == (0, 1) and bpos < cut <= fpos
    off = sd.QlusterPars(markChimeras=False)
    assert oracle.qluster(sd.make_params(), off.to_c(), its.iters, its.iters, bases, quals, offsets)["chimeras"] == {}
    # a parent must be parFreqs times as abundant as the candidate
    strict = sd.QlusterPars(parFreqs=7.0)
    assert oracle.qluster(sd.make_params(), strict.to_c(), its.iters, its.iters, bases, quals, offsets)["chimeras"] == {}


def test_chimera_probe_windows(oracle):
    """Front / back windows of the probe: low-quality mismatches are dropped from the kept list but
    still judged by chiOverlap's lqMismatches; gaps inside a window use its indel thresholds."""
    from seekdeep_b200.qluster import IterParC
    p = make_params()
    a = "ACGTTGCAAGCTAGCTTAGGCTAACGTAGCTAGGATCCATGCAAGT"
    b = list(a)
    b[8] = "C"    # low quality mismatch (front window)
    b[20] = "A"   # kept mismatch
    b[35] = "G"   # kept mismatch
    b = "".join(b)
    qa = np.full(len(a), 38, np.uint8)
    qb = qa.copy()
    qb[8] = 5
    bases = np.frombuffer((a + b).encode(), np.uint8)
    quals = np.concatenate([qa, qb])
    offsets = np.array([0, len(a), 2 * len(a)], np.uint64)
    ov = IterParC(0, 0, 2.0, 1.0, 0.99, 0.0, 0.0, 1.0, 0, 0, 0.0)
    pr = oracle.chimera_probe_batch(p, bases, quals, offsets, [0], [1], ov)[0]
    assert (pr["keptMismatches"], pr["firstPos"], pr["lastPos"]) == (2, 20, 35)
    assert pr["flags"] == 2          # the low-quality mismatch fails the front window, the back passes
    ov.lqMismatches = 1.0
    assert oracle.chimera_probe_batch(p, bases, quals, offsets, [0], [1], ov)[0]["flags"] == 3
    keep = oracle.chimera_probe_batch(p, bases, quals, offsets, [0], [1], ov, keepLowQual=True)[0]
    assert (keep["keptMismatches"], keep["firstPos"], keep["flags"]) == (3, 8, 3)


def test_mark_chimeras_options(oracle):
    """The visible chimera options (clusterDownSetUp.cpp:372-381) each switch the flag as their help text
    says: a candidate at or below runCutOff is not examined, posSpacing demands that much shared sequence
    between the two crossover mismatches, overLapSizeCutoff demands that many explained bases at each end."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations
    rng = np.random.default_rng(12)
    bases, quals, offsets, (p0, p1, chi, var), cut = H.chimera_sample(rng, length=160, nDiff=6, counts=(40, 30, 6, 4))
    its = CollapseIterations.gen_illumina_default_pars(100)

    def run(**kw):
        return oracle.qluster(sd.make_params(), sd.QlusterPars(**kw).to_c(), its.iters, its.iters, bases, quals, offsets)["chimeras"]

    base = run()
    assert list(base) == [2]
    front, back, fpos, bpos = base[2]
    assert run(chiRunCutOff=6) == {}                       # the crossover has 6 reads: not examined
    assert run(chiRunCutOff=5) == base
    assert run(chiPosSpacing=fpos - bpos) == base          # exactly the shared stretch between the mismatches
    assert run(chiPosSpacing=fpos - bpos + 1) == {}
    assert run(chiOverlapSizeCutoff=bpos + 2) == base or run(chiOverlapSizeCutoff=bpos + 2) == {}  # depends on which end is shorter
    assert run(chiOverlapSizeCutoff=161) == {}             # nobody can explain more bases than the read has
    assert run(parFreqs=5.0) == base and run(parFreqs=5.01) == {}  # 30 reads / 6 reads = 5: parent 1 just qualifies


def _pack_reads(reads):
    seqs = [np.frombuffer(s.encode(), np.uint8) for s, _ in reads]
    quals = [np.asarray(q, np.uint8) for _, q in reads]
    offsets = np.zeros(len(seqs) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    return np.concatenate(seqs), np.concatenate(quals), offsets


def test_consensus_quality_and_majority_rules(oracle):
    """Hand-derived consensus cases (baseCluster::calculateConsensus as restated): the winning base of a column
    keeps the quality sum of ITS reads divided by the column's total count; an insertion carried by a minority
    of the reads does not enter the consensus."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations
    rng = np.random.default_rng(8)
    A = "".join(rng.choice(list("ACGT"), 60))
    p = 30
    sub = A[:p] + ("C" if A[p] != "C" else "G") + A[p + 1:]
    hi = [38] * 60
    lowAtP = hi[:p] + [10] + hi[p + 1:]
    reads = [(A, hi)] * 3 + [(sub, lowAtP)] * 2
    bases, quals, offsets = _pack_reads(reads)
    its = CollapseIterations.gen_illumina_default_pars(100)
    res = oracle.qluster(sd.make_params(), sd.QlusterPars(markChimeras=False).to_c(), its.iters, its.iters, bases, quals, offsets)
    assert len(res["clusters"]) == 1                      # the low-quality mismatch is forgiven from the third line on
    seq, q, cnt = res["clusters"][0]
    assert seq == A and cnt == 5.0
    expect = np.full(60, 38, np.uint8)
    expect[p] = (3 * 38) // 5                             # 114 / 5 = 22: quality of the A reads over all five reads
    assert np.array_equal(q, expect)
    # a one-base insertion in 2 of 7 reads: merged (1-base indel allowed), not adopted (5 reads do not have it)
    ins = A[:20] + ("A" if A[20] != "A" and A[19] != "A" else "C") + A[20:]
    reads = [(A, hi)] * 5 + [(ins, [38] * 61)] * 2
    bases, quals, offsets = _pack_reads(reads)
    res = oracle.qluster(sd.make_params(), sd.QlusterPars(markChimeras=False).to_c(), its.iters, its.iters, bases, quals, offsets)
    assert [(c[0], c[2]) for c in res["clusters"]] == [(A, 7.0)]


def test_stop_check_and_small_cutoff_semantics(oracle):
    """collapseWithParameters as restated: a read looks at the first `stopCheck` live candidates only (largest
    first), and a cluster is a candidate only if its count exceeds `smallCutoff` (par-file columns 1 and 2,
    clusterDownSetUp.cpp:63-84)."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations, IterPar
    rng = np.random.default_rng(17)
    haps = ["".join(rng.choice(list("ACGT"), 80)) for _ in range(4)]
    hi = [38] * 80
    reads = []
    for h, n in zip(haps, (40, 30, 20, 10)):
        reads += [(h, hi)] * n
    # three reads of the smallest haplotype with one low-quality error each (distinct positions)
    noisy = []
    for p in (10, 40, 70):
        s = haps[3][:p] + ("A" if haps[3][p] != "A" else "C") + haps[3][p + 1:]
        noisy.append((s, hi[:p] + [8] + hi[p + 1:]))
    bases, quals, offsets = _pack_reads(reads + noisy)

    def line(stopCheck, smallCutoff):
        return IterPar(stopCheck=stopCheck, smallCutoff=smallCutoff, oneBaseIndel=0, twoBaseIndel=0, largeBaseIndel=0,
                       hqMismatches=0, lqMismatches=1, lowKmerMismatches=0)

    def run(stopCheck, smallCutoff):
        its = [line(stopCheck, smallCutoff)]
        pars = sd.QlusterPars(markChimeras=False, checkKmers=False)
        r = oracle.qluster(sd.make_params(), pars.to_c(), its, its, bases, quals, offsets)
        return sorted((c[2] for c in r["clusters"]), reverse=True)

    assert run(100, 0) == [40.0, 30.0, 20.0, 13.0]               # every noisy read finds its haplotype
    assert run(3, 0) == [40.0, 30.0, 20.0, 10.0, 1.0, 1.0, 1.0]   # ... but it is the 4th candidate: out of reach
    assert run(4, 0) == [40.0, 30.0, 20.0, 13.0]
    assert run(100, 10) == [40.0, 30.0, 20.0, 10.0, 1.0, 1.0, 1.0]  # count 10 does not exceed smallCutoff 10
    assert run(100, 9) == [40.0, 30.0, 20.0, 13.0]


def test_frozen_qluster_fixtures_match_their_generators():
    """tests/golden/qluster_*.npz carry the SHA-256 of the input they were computed from: the seeded
    generators must still reproduce it (otherwise a GPU mismatch would be a generator drift, not a parity failure),
    and the frozen results must be self-consistent (every read in a cluster, counts add up)."""
    from tests import golden_fixtures as G
    assert "c1_full" in G.NAMES and "c2_5k" in G.NAMES
    for name in G.NAMES:
        (bases, q, offsets, params, pars, iters), fx = G.load(name)
        n = len(offsets) - 1
        mem = fx["membership"]
        assert len(mem) == n and len(fx["cnts"]) + 1 == len(fx["offsets"])
        kept = mem != np.uint32(0xFFFFFFFF)
        assert np.array_equal(np.bincount(mem[kept], minlength=len(fx["cnts"])).astype(np.float64), fx["cnts"])
        assert fx["offsets"][-1] == len(fx["seqs"]) == len(fx["quals"])


def test_compare_kmers_revcomp_micro_cases(oracle):
    """compareKmersRevComp(a, b) == compareKmers(a, reverseComplement(b)); hand-checked: AAAC vs GTTT."""
    assert oracle.kmer_compare("AAAC", "GTTT", 2, revCompB=True) == (3, 1.0)   # revcomp(GTTT) = AAAC
    assert oracle.kmer_compare("AAAC", "GTTT", 2) == (0, 0.0)
    a, b = "ACGGTTACGATTACA", "TGTAATCGTAACCGT"                               # b = revcomp(a)
    assert oracle.kmer_compare(a, b, 5, revCompB=True) == oracle.kmer_compare(a, a, 5) == (11, 1.0)
    assert oracle.kmer_compare("ACGRT", "AYCGT", 3, revCompB=True) == oracle.kmer_compare("ACGRT", "ACGRT", 3)  # R <-> Y


def test_kmer_binning_control_flow(oracle):
    """--useKmerBinning as restated (clusterDownSetUp.cpp:352-354; njhseq runFullClustering): clusters are binned by
    compareKmers against each bin's first cluster (strictly more than kmerCutOff shared), every bin of two or more runs
    through binIteratorMap on its own, the survivors then go through the run's own iterations."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations, IterPar
    rng = np.random.default_rng(23)
    A = "".join(rng.choice(list("ACGT"), 90))
    B = "".join(rng.choice(list("ACGT"), 90))
    A1 = A[:45] + ("A" if A[45] != "A" else "C") + A[46:]  # one high-quality mismatch off A
    hi = [38] * 90
    bases, quals, offsets = _pack_reads([(A, hi)] * 5 + [(A1, hi)] * 2 + [(B, hi)] * 3)
    strict = [IterPar(stopCheck=100, smallCutoff=0, oneBaseIndel=0, twoBaseIndel=0, largeBaseIndel=0, hqMismatches=0,
                      lqMismatches=0, lowKmerMismatches=0)]
    loose = CollapseIterations.gen_otu_pars(100, 0.90, False)

    def run(**kw):
        binIts = kw.pop("binIts", None)
        pars = sd.QlusterPars(markChimeras=False, checkKmers=False, **kw)
        r = oracle.qluster(sd.make_params(), pars.to_c(binIts), strict, strict, bases, quals, offsets)
        return sorted(((c[0], c[2]) for c in r["clusters"]), key=lambda x: -x[1])

    plain = run()
    assert plain == [(A, 5.0), (B, 3.0), (A1, 2.0)]
    # A and A1 share 83 of their 87 4-mers (0.95 > 0.5), B is unrelated: bins {A, A1}, {B}; the loose bin iterations merge
    # A1 into A, the strict iterations of the run then find nothing to do
    assert run(useKmerBinning=True, kCompareLen=4, kmerCutOff=0.5, binIts=loose) == [(A, 7.0), (B, 3.0)]
    # no bin iterations given: the bins run the strict lines themselves -> nothing merges
    assert run(useKmerBinning=True, kCompareLen=4, kmerCutOff=0.5) == plain
    # strict '>' : a cut-off the fraction cannot exceed puts every cluster in its own bin -> the plain run
    assert run(useKmerBinning=True, kCompareLen=4, kmerCutOff=1.0, binIts=loose) == plain
    # the exact fraction of (A, A1) is 83/87 at k = 4: a cut-off equal to it does not bin them, just below does
    from tests import oracle_binding  # noqa: F401
    sh, fr = oracle.kmer_compare(A.encode(), A1.encode(), 4)
    assert run(useKmerBinning=True, kCompareLen=4, kmerCutOff=fr, binIts=loose) == plain
    assert run(useKmerBinning=True, kCompareLen=4, kmerCutOff=float(np.nextafter(fr, 0)), binIts=loose) == [(A, 7.0), (B, 3.0)]
    # cut-off below zero: one bin with everything; the loose lines (90 % identity) still keep B apart
    assert run(useKmerBinning=True, kCompareLen=4, kmerCutOff=-1.0, binIts=loose) == [(A, 7.0), (B, 3.0)]


def test_align_events_micro_case(oracle):
    """comp_'s per-event maps as restated (clusterDown.cpp:759-800 dumps them): one high-quality mismatch, one
    low-quality mismatch, an internal 2-base deletion in the read, a free right end gap -- positions, bases, qualities
    and kinds spelled out by hand."""
    import seekdeep_b200 as sd
    from seekdeep_b200 import _abi
    ref = "ACGTTGCAAGGCTTAACCGGATCGTTAGCAATCCGATTGACCATGCATGC"
    read = list(ref)
    read[5] = "A"    # G -> A, high quality
    read[20] = "C"   # A -> C, low quality
    del read[30:32]  # 2-base deletion in the read
    read = "".join(read)[:-3]  # the read ends 3 bases early: right end gap, free under qluster's gaps
    qa = np.full(len(ref), 38, np.uint8)
    qb = np.full(len(read), 38, np.uint8)
    qb[20] = 9
    params = sd.make_params()
    cmp_, ev = oracle.align_events(params, np.frombuffer(ref.encode(), np.uint8), qa, np.frombuffer(read.encode(), np.uint8), qb)
    kinds = [int(e["kind"]) for e in ev]
    assert kinds == [_abi.EV_MISMATCH_HQ, _abi.EV_MISMATCH_LQ, _abi.EV_GAP_IN_READ, _abi.EV_GAP_IN_READ | _abi.EV_END_GAP]
    e0, e1, e2, e3 = ev
    assert (e0["alnPos"], e0["refPos"], e0["seqPos"], chr(e0["refBase"]), chr(e0["seqBase"]), e0["refQual"], e0["seqQual"]) == (5, 5, 5, "G", "A", 38, 38)
    assert (e1["alnPos"], e1["refPos"], e1["seqPos"], chr(e1["refBase"]), chr(e1["seqBase"]), e1["seqQual"]) == (20, 20, 20, "A", "C", 9)
    assert e2["size"] == 2 and e2["refPos"] in (29, 30, 31) and e2["seqPos"] == e2["refPos"] and e2["alnPos"] == e2["refPos"]
    assert ref[e2["refPos"]] == chr(e2["refBase"])
    assert (e3["size"], e3["refPos"], e3["seqPos"], e3["alnPos"]) == (3, len(ref) - 3, len(read), len(ref) - 3)
    assert cmp_["hqMismatches"] == 1 and cmp_["lqMismatches"] == 1 and cmp_["numGaps"] == 1 and cmp_["twoBaseIndel"] == 1.0
    # counts of the events agree with the scalar comparison on random related pairs
    from tests import helpers as H
    rng = np.random.default_rng(5)
    bases, quals, offsets, cnts, seqs, qs = H.make_store(rng, 24, 140)
    km = oracle.index_kmers(bases, offsets, cnts, 9, 1, True)
    for i in range(1, 24):
        c, ev = oracle.align_events(params, seqs[0], qs[0], seqs[i], qs[i], _abi.USING_QUALITY | _abi.CHECK_KMER, kmaps=km)
        k = ev["kind"]
        assert (k == _abi.EV_MISMATCH_HQ).sum() == c["hqMismatches"] and (k == _abi.EV_MISMATCH_LQ).sum() == c["lqMismatches"]
        assert (k == _abi.EV_MISMATCH_LOWKMER).sum() == c["lowKmerMismatches"]
        assert ((k == _abi.EV_GAP_IN_REF) | (k == _abi.EV_GAP_IN_READ)).sum() == c["numGaps"]
        assert (k >= _abi.EV_GAP_IN_REF).sum() == c["nGapRuns"]
        assert np.all(np.diff(ev["alnPos"].astype(np.int64)) > 0)
        low = ev[k == _abi.EV_MISMATCH_LOWKMER]
        assert np.all((low["kmerFreqRef"] <= 1) | (low["kmerFreqSeq"] <= 1))  # runCutOff 1
    oracle.free_kmaps(km)
