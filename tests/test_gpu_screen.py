"""The screen pass (screen_kernel.cu) and the device-generated (reads x clusters) grid: the verdict
masks must be exactly those of the bit-recording kernels and of the oracle, whatever the screen
decides to finish by itself.  The stores below are built to sit on the screen's decision boundary:
equal-length relatives (finished by the screen), equal-length pairs whose optimum has gaps or ties
with a gapped alignment (tandem repeats, homopolymer shifts, a deletion compensated by an insertion),
pairs that need the 16 / 32 / 64-diagonal band, and pairs of unequal length."""
import numpy as np
import pytest

from seekdeep_b200 import _abi
from seekdeep_b200.aligner import GpuAligner, pack_seqs
from seekdeep_b200.iterpars import CollapseIterations, IterPar
from seekdeep_b200.scoring import GapScoringParameters, SubstituteMatrix, make_params
from tests import helpers as H

pytestmark = pytest.mark.gpu


def boundary_store(rng, length=240):
    seqs = []
    root = H.rand_seq(rng, length)
    rep2 = np.frombuffer(("AC" * (length // 2 + 1)).encode(), np.uint8)[:length].copy()
    rep7 = np.frombuffer(("GATTACA" * (length // 7 + 1)).encode(), np.uint8)[:length].copy()
    hp = np.frombuffer(("AAAAAACCCCCCGGGGGTTTTTT" * (length // 23 + 1)).encode(), np.uint8)[:length].copy()
    for base in (root, rep2, rep7, hp):
        seqs.append(base)
        for k in range(26):  # substitutions only: 0 .. 25 of them (bands of 16, 32, 64 and beyond)
            seqs.append(H.mutate(rng, base, k))
        for _ in range(10):  # same length, but a deletion here and an insertion there
            s = base.copy()
            p, r = sorted(rng.choice(np.arange(5, length - 5), 2, replace=False))
            s = np.delete(s, p)
            s = np.insert(s, r, H.BASES[rng.integers(4)])
            seqs.append(H.mutate(rng, s, int(rng.integers(0, 4))))
        for sh in (1, 2, 3, 7):  # rotated copies: same length, best alignment is shifted (end gaps)
            seqs.append(np.roll(base, sh))
        for _ in range(6):  # unequal lengths
            seqs.append(H.mutate(rng, base, int(rng.integers(0, 3)), int(rng.integers(0, 2)), 1))
    # substitutions next to homopolymer runs: a mismatch can also be written as two gaps
    for k in range(12):
        s = hp.copy()
        for p in rng.choice(np.arange(1, length - 1), 1 + k % 4, replace=False):
            s[p] = s[p - 1]
        seqs.append(s)
    return seqs


TABLE = CollapseIterations.gen_illumina_default_pars(100).iters[:11] + [
    IterPar(100, 3, 2, 1, 0.99, 1, 6, 1), IterPar(100, 0, 5, 5, 5, 30, 30, 30),
    IterPar(100, 0, onPerId=True, onHqPerId=True, perId=0.97), IterPar(100, 0, onPerId=True, onHqPerId=False, perId=0.95)]


@pytest.mark.parametrize("pname", ["qluster", "allEndsCost", "allEndsFree", "asym", "countEnds", "degenHomopolymer", "stitch"])
def test_screen_pass_is_exact(oracle, pname):
    rng = np.random.default_rng(len(pname) * 13 + 1)
    seqs = boundary_store(rng)
    quals = [H.rand_quals(rng, len(s), 0.2) for s in seqs]
    bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
    n = len(seqs)
    refIdx, readIdx = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(n), np.arange(n))]
    params = H.params_for(pname)
    with GpuAligner(params) as al:
        al.upload_seqs(bases, q, offsets)
        al.index_kmers(9, 1, True)
        al.set_pass_table(TABLE)
        al.profile_enable(True)
        al.profile_read(reset=True)
        got = al.align_pass(refIdx, readIdx, checkKmer=True)
        prof = al.profile_read()
        al.set_screen(1)
        plain = al.align_pass(refIdx, readIdx, checkKmer=True)
        profPlain = al.profile_read()
        al.set_band(1)
        full = al.align_pass(refIdx, readIdx, checkKmer=True)
    km = oracle.index_kmers(bases, offsets, None, 9, 1, True)
    want = oracle.align_profile_batch(params, bases, q, offsets, refIdx, readIdx, _abi.USING_QUALITY | _abi.CHECK_KMER, kmaps=km)
    oracle.free_kmaps(km)
    wantMasks = H.masks_from_comparisons(want, TABLE)
    for name, m in (("screen", got), ("band", plain), ("full", full)):
        bad = np.nonzero(m != wantMasks)[0]
        assert bad.size == 0, f"{name}: pair {bad[0]} (ref {refIdx[bad[0]]}, read {readIdx[bad[0]]}) got {m[bad[0]]:#x} want {wantMasks[bad[0]]:#x}"
    assert prof["screenPairs"] == len(refIdx) and profPlain["screenPairs"] == 0
    # the screen finished a good share of the pairs, but by far not all of them
    assert 0.03 * len(refIdx) < prof["screenCertified"] < 0.9 * len(refIdx)
    assert prof["screenCells"] > 0 and prof["screenCertCells"] > 0
    assert prof["cellUpdates"] == profPlain["cellUpdates"]  # every pair is accounted for once, whoever finished it
    assert int((wantMasks != 0).sum()) > 0 and len(np.unique(wantMasks)) > 3


def test_screen_every_certified_pair_is_gapless(oracle):
    """Soundness seen from the other side: with a table line only the gapless, mismatch-free alignment of
    equal-length sequences can pass ... and one that everything passes, the masks tell which pairs the
    oracle aligned without gaps; the screen may only have finished pairs among those."""
    rng = np.random.default_rng(5)
    seqs = boundary_store(rng, 180)
    quals = [np.full(len(s), 38, np.uint8) for s in seqs]
    bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
    n = len(seqs)
    refIdx, readIdx = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(n), np.arange(n))]
    params = make_params()
    want, gaps = oracle.align_profile_batch(params, bases, q, offsets, refIdx, readIdx, _abi.USING_QUALITY, gapCap=4)
    with GpuAligner(params) as al:
        al.upload_seqs(bases, q, offsets)
        al.set_pass_table([IterPar(100, 0, 999, 999, 999, 999, 999, 999)])
        al.profile_enable(True)
        al.profile_read(reset=True)
        masks = al.align_pass(refIdx, readIdx)
        prof = al.profile_read()
    assert (masks == 1).all()
    gapless = int((want["nGapRuns"] == 0).sum())
    assert 0 < prof["screenCertified"] <= gapless < len(refIdx)


def test_grid_pass_matches_pair_list(oracle):
    """sdgpu_align_pass_grid: pairs generated on the device, only non-zero masks returned; a sequence that is
    both a read and a cluster is not compared with itself; more pairs than one 2 M-pair chunk."""
    rng = np.random.default_rng(11)
    haps = [H.rand_seq(rng, 150)]
    for _ in range(11):
        haps.append(H.mutate(rng, haps[0], int(rng.integers(1, 6))))
    seqs = []
    for _ in range(3000):
        h = haps[int(rng.integers(len(haps)))]
        seqs.append(H.mutate(rng, h, int(rng.integers(0, 4)), 0, int(rng.random() < 0.1)))
    seqs = haps + seqs
    quals = [H.rand_quals(rng, len(s), 0.1) for s in seqs]
    bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
    clus = np.arange(12, dtype=np.uint32)
    reads = np.concatenate([np.arange(len(seqs) - 1, -1, -1)]).astype(np.uint32)  # includes the clusters themselves
    table = CollapseIterations.gen_illumina_default_pars(100).iters[:11]
    params = make_params()
    with GpuAligner(params) as al:
        al.upload_seqs(bases, q, offsets)
        al.index_kmers(9, 1, True)
        al.set_pass_table(table)
        hits = al.align_pass_grid(clus, reads, checkKmer=True)
        refIdx = np.tile(clus, len(reads))
        readIdx = np.repeat(reads, len(clus))
        masks = al.align_pass(refIdx, readIdx, checkKmer=True)
        al.set_screen(1)
        hitsPlain = al.align_pass_grid(clus, reads, checkKmer=True)
        # a grid wide enough for several chunks: every sequence against the first 800
        big = al.align_pass_grid(np.arange(800, dtype=np.uint32), reads, checkKmer=True)
        bigMasks = al.align_pass(np.tile(np.arange(800, dtype=np.uint32), len(reads)), np.repeat(reads, 800), checkKmer=True)
    masks = masks.reshape(len(reads), len(clus))
    masks[reads[:, None] == clus[None, :]] = 0  # self pairs are skipped by the grid
    for h in (hits, hitsPlain):
        dense = np.zeros_like(masks)
        assert len(np.unique(h["read"].astype(np.uint64) * 64 + h["cluster"])) == len(h)  # no pair twice
        dense[h["read"], h["cluster"]] = h["mask"]
        assert (h["mask"] != 0).all()
        assert np.array_equal(dense, masks)
    assert 0 < len(hits) < masks.size
    bigMasks = bigMasks.reshape(len(reads), 800)
    bigMasks[reads[:, None] == np.arange(800)[None, :]] = 0
    dense = np.zeros_like(bigMasks)
    dense[big["read"], big["cluster"]] = big["mask"]
    assert len(reads) * 800 > (1 << 21) and np.array_equal(dense, bigMasks)
    # and against the oracle on a slice
    sl = slice(0, 4000)
    want = oracle.align_profile_batch(params, bases, q, offsets, refIdx[sl], readIdx[sl], _abi.USING_QUALITY | _abi.CHECK_KMER,
                                      kmaps=(km := oracle.index_kmers(bases, offsets, None, 9, 1, True)))
    oracle.free_kmaps(km)
    wm = H.masks_from_comparisons(want, table)
    wm[refIdx[sl] == readIdx[sl]] = 0
    assert np.array_equal(masks.ravel()[sl], wm)


def test_screen_fuzz_vs_bit_kernels():
    """Random penalties / scores / lengths / divergence: screen on == screen off, pair for pair."""
    rng = np.random.default_rng(2026)
    for trial in range(60):
        go = int(rng.integers(1, 12))
        ge = int(rng.integers(0, go + 1))
        ends = [(int(o), int(rng.integers(0, o + 1))) for o in rng.integers(0, 9, 4)]
        gaps = GapScoringParameters(go, ge, ends[0][0], ends[0][1], ends[1][0], ends[1][1], ends[2][0], ends[2][1], ends[3][0], ends[3][1])
        match, mis = int(rng.integers(1, 6)), -int(rng.integers(1, 7))
        params = make_params(gaps=gaps, scoring=SubstituteMatrix.create_degen_score_matrix_case_insensitive(match, mis)
                             if trial % 3 == 0 else SubstituteMatrix.create_score_matrix(match, mis))
        L = int(rng.choice([40, 41, 63, 64, 65, 97, 130, 255, 300, 420]))
        base = H.rand_seq(rng, L) if trial % 4 else np.frombuffer(("ACGGT" * (L // 5 + 1)).encode(), np.uint8)[:L].copy()
        seqs = [base] + [H.mutate(rng, base, int(rng.integers(0, 20))) for _ in range(30)]
        for _ in range(6):
            s = np.delete(base, int(rng.integers(L)))
            seqs.append(np.insert(s, int(rng.integers(L - 1)), H.BASES[rng.integers(4)]))
        seqs.append(np.roll(base, 1))
        seqs.append(base[:-1])
        quals = [H.rand_quals(rng, len(s), 0.3) for s in seqs]
        bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
        n = len(seqs)
        refIdx, readIdx = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(n), np.arange(n))]
        with GpuAligner(params) as al:
            al.upload_seqs(bases, q, offsets)
            al.set_pass_table(TABLE)
            a = al.align_pass(refIdx, readIdx)
            al.set_screen(1)
            b = al.align_pass(refIdx, readIdx)
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (trial, go, ge, ends, match, mis, L, int(refIdx[bad[0]]), int(readIdx[bad[0]]))
