"""Paired-end stitching: PairedReadProcessor::processPairedEnd on the GPU path vs the oracle's
string-based restatement of PairedReadProcessor.cpp:287-664, plus hand-checked micro-cases of the oracle itself."""
import numpy as np
import pytest

from seekdeep_b200 import stitch as ST
from seekdeep_b200.aligner import GpuAligner, pack_seqs
from tests import helpers as H

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def make_pairs(rng, n, ampLen=(120, 420), readLen=150, errRate=0.004, lowTail=True):
    """Fragments of random length sequenced from both ends: r1 = first readLen bases, r2 (already reverse-complemented
    back to the forward strand, like PairedRead::mateSeqBase_) = last readLen bases; substitution errors, qualities
    dropping towards the 3' ends (r1's end, r2's start); a few pairs from unrelated fragments (no overlap)."""
    r1s, q1s, r2s, q2s = [], [], [], []
    for i in range(n):
        L = int(rng.integers(*ampLen))
        frag = H.rand_seq(rng, L)
        other = H.rand_seq(rng, L) if i % 17 == 0 else frag  # every 17th pair does not belong together
        rl1, rl2 = min(readLen, L), min(readLen, L)
        if i % 11 == 0:  # read-through: reads longer than the fragment, adaptor-like tails
            rl1 = rl2 = L
            a = np.concatenate([frag, H.rand_seq(rng, int(rng.integers(5, 40)))])
            b = np.concatenate([H.rand_seq(rng, int(rng.integers(5, 40))), other])
        else:
            a = frag[:rl1].copy()
            b = other[L - rl2:].copy()
        qa = rng.integers(30, 41, len(a)).astype(np.uint8)
        qb = rng.integers(30, 41, len(b)).astype(np.uint8)
        if lowTail and i % 3 == 0:  # quality collapses towards the 3' ends
            k = int(rng.integers(10, 60))
            qa[-k:] = rng.integers(2, 20, k)
            qb[:k] = rng.integers(2, 20, k)
            for x in rng.choice(k, k // 3, replace=False):  # ... and so does accuracy
                a[len(a) - 1 - x] = H.BASES[rng.integers(4)]
                b[x] = H.BASES[rng.integers(4)]
        for s, q in ((a, qa), (b, qb)):
            err = np.nonzero(rng.random(len(s)) < errRate)[0]
            for p in err:
                s[p] = H.BASES[rng.integers(4)]
                q[p] = rng.integers(2, 30)
        r1s.append(a), q1s.append(qa), r2s.append(b), q2s.append(qb)
    return pack_seqs([s.tobytes() for s in r1s], q1s), pack_seqs([s.tobytes() for s in r2s], q2s)


def test_oracle_stitch_micro_cases(oracle):
    """Hand-checked: a 30-base overlap between two 50-base mates of an 70-base fragment -> R1ENDSINR2, the combined read
    is the fragment; a mismatch in the overlap takes the base of the higher quality (r1 on ties); unrelated mates fail."""
    frag = b"ACGGTCAGTTAGCCATGCAATCGGATCCTAGGCATTACGGATCAGGCTTAACGTACCGTTAGGCATCGAA"
    assert len(frag) == 70
    r1, r2 = frag[:50], frag[20:]
    q1, q2 = np.full(50, 38, np.uint8), np.full(50, 37, np.uint8)
    pars, sp = ST.pair_aligner_params(), ST.ProcessParams()
    st, ov, seqs, _ = oracle.stitch_pairs(pars, sp, pack_seqs([r1], [q1]), pack_seqs([r2], [q2]))
    assert st[0] == ST.R1ENDSINR2 and ov["shift"][0] == 20 and ov["basesInAln"][0] == 30 and ov["alnScore"][0] == 60
    assert seqs[0][0] == frag and list(seqs[0][1][:20]) == [38] * 20 and list(seqs[0][1][20:50]) == [38] * 30 and list(seqs[0][1][50:]) == [37] * 20
    r2m = bytearray(r2)
    r2m[5] = ord("A") if r2m[5] != ord("A") else ord("C")  # a mismatch at fragment position 25
    q2m = q2.copy()
    q2m[5] = 39                                             # the mate is surer of it
    st, ov, seqs, _ = oracle.stitch_pairs(pars, sp, pack_seqs([r1], [q1]), pack_seqs([bytes(r2m)], [q2m]))
    assert st[0] == ST.NOOVERLAP and ov["eventBasedIdentityHq"][0] == 29 / 30  # one HQ mismatch in 30 columns: below 1 - errorAllowed
    sp = ST.ProcessParams(errorAllowed=0.05)
    st, ov, seqs, _ = oracle.stitch_pairs(pars, sp, pack_seqs([r1], [q1]), pack_seqs([bytes(r2m)], [q2m]))
    assert st[0] == ST.R1ENDSINR2 and ov["hqMismatches"][0] == 1 and seqs[0][0][25] == r2m[5] and seqs[0][1][25] == 39
    q2m[5] = 38                                             # a tie: r1 keeps its base
    st, ov, seqs, _ = oracle.stitch_pairs(pars, sp, pack_seqs([r1], [q1]), pack_seqs([bytes(r2m)], [q2m]))
    assert seqs[0][0][25] == frag[25]
    sp = ST.ProcessParams()
    # read-through: both mates cover the whole fragment and run into adaptor: only the fragment comes back
    st, ov, seqs, _ = oracle.stitch_pairs(pars, sp, pack_seqs([frag + b"TTTTGGGGCC"], [np.full(80, 38, np.uint8)]),
                                          pack_seqs([b"CCAAGGTTAC" + frag], [np.full(80, 38, np.uint8)]))
    assert st[0] == ST.R1BEGINSINR2 and ov["shift"][0] == -10 and seqs[0][0] == frag
    st, ov, seqs, _ = oracle.stitch_pairs(pars, sp, pack_seqs([b"ACGT" * 12], [np.full(48, 38, np.uint8)]),
                                          pack_seqs([b"GGCATTCAAGCTAGCTTTAGGACCCATAGGAC"], [np.full(32, 38, np.uint8)]))
    assert st[0] == ST.NOOVERLAP and seqs[0][0] == b""


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["defaults", "noTrim", "strict", "r1r2Trim"])
def test_stitching_matches_oracle(oracle, variant):
    rng = np.random.default_rng({"defaults": 1, "noTrim": 2, "strict": 3, "r1r2Trim": 4}[variant])
    r1, r2 = make_pairs(rng, 1500)
    sp = {"defaults": ST.ProcessParams(), "noTrim": ST.ProcessParams(trimLowQaulWindows=False),
          "strict": ST.ProcessParams(errorAllowed=0.0, minOverlap=25, hqMismatchCutOff=0, lqMismatchCutOff=3, hardMismatchCutOff=3),
          "r1r2Trim": ST.ProcessParams(r1Trim=7, r2Trim=4, windowSize=8, windowStep=2, avgQualCutOff=22.0)}[variant]
    pars = ST.pair_aligner_params()
    with GpuAligner(pars) as al:
        got = ST.process_pairs(al, r1, r2, sp)
    st, ov, seqs, retried = oracle.stitch_pairs(pars, sp, r1, r2)
    assert np.array_equal(got["status"], st)
    for name in ST.OVERLAP_DTYPE.names:
        bad = np.nonzero(got["overlaps"][name] != ov[name])[0]
        assert bad.size == 0, f"{name}: pair {bad[0]} got {got['overlaps'][name][bad[0]]} want {ov[name][bad[0]]}"
    assert int(got["counts"][6]) == retried
    off = got["offsets"]
    for i in range(len(st)):
        lo, hi = int(off[i]), int(off[i + 1])
        assert got["bases"][lo:hi].tobytes() == seqs[i][0], i
        assert np.array_equal(got["quals"][lo:hi], seqs[i][1]), i
    assert [int(x) for x in got["counts"][:6]] == [int((st == k).sum()) for k in range(6)]
    seen = set(np.unique(st).tolist())
    assert {ST.NOOVERLAP, ST.R1BEGINSINR2, ST.R1ENDSINR2} <= seen
    if variant == "defaults":
        assert retried > 0


@pytest.mark.gpu
def test_overlap_kernel_general_penalties_and_containment(oracle):
    """sdgpu_align_no_internal_gaps on its own: end-gap penalties that are not zero move the best shift; a mate contained
    in the other (R1ALLINR2 / R2ALLINR1 geometry); k-mer flag classification; longest supported reads."""
    from seekdeep_b200.scoring import GapScoringParameters, make_params
    rng = np.random.default_rng(8)
    base = H.rand_seq(rng, 900)
    seqs = [base, base[100:400].copy(), base[350:].copy(), H.mutate(rng, base[200:700], 6), base[:1024 - 200].copy(), H.rand_seq(rng, 1024)]
    quals = [H.rand_quals(rng, len(s), 0.3) for s in seqs]
    b, q, off = pack_seqs([s.tobytes() for s in seqs], quals)
    a = np.array([0, 1, 0, 2, 3, 0, 4, 5, 1], np.uint32)
    c = np.array([1, 0, 2, 0, 0, 3, 2, 0, 2], np.uint32)
    for gaps in (GapScoringParameters.from_strings("10,1", "0,0", "0,0"), GapScoringParameters(7, 2, 3, 1, 2, 1, 4, 2, 6, 1)):
        params = make_params(gaps=gaps, qualThresWindow=0)
        with GpuAligner(params) as al:
            al.upload_seqs(b, q, off)
            al.index_kmers(7, 1, True)
            got = ST.align_no_internal_gaps(al, a, c, checkKmer=True, doingMatchQuality=True)
        r1 = pack_seqs([seqs[i].tobytes() for i in a], [quals[i] for i in a])
        r2 = pack_seqs([seqs[i].tobytes() for i in c], [quals[i] for i in c])
        _, ov, _, _ = oracle.stitch_pairs(params, ST.ProcessParams(trimLowQaulWindows=False), r1, r2)
        for name in ("alnScore", "shift", "basesInAln"):
            assert np.array_equal(got[name], ov[name]), (name, got[name], ov[name])
        assert np.array_equal(got["hqMismatches"] + got["lqMismatches"] + got["lowKmerMismatches"], ov["hqMismatches"] + ov["lqMismatches"])
        assert np.array_equal(got["highQualityMatches"] + got["lowQualityMatches"], ov["highQualityMatches"])
    assert got["shift"][0] == 100 and got["shift"][1] == -100 and got["basesInAln"][0] == 300
