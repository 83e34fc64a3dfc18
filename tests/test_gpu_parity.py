"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
Integer / index fields bit-exact; float fields within 1e-6 relative (north_star tolerance)."""
import numpy as np
import pytest

from seekdeep_b200 import _abi
from seekdeep_b200.aligner import GpuAligner, SdgpuError, pack_seqs
from seekdeep_b200.scoring import GapScoringParameters, make_params
from tests import helpers as H

pytestmark = pytest.mark.gpu


def run_both(oracle, params, bases, quals, offsets, cnts, refIdx, readIdx, flags, gapCap=0, kmer=None):
    km = None
    with GpuAligner(params) as al:
        al.upload_seqs(bases, quals, offsets, cnts)
        if kmer:
            al.index_kmers(**kmer)
            km = oracle.index_kmers(bases, offsets, cnts, kmer["kLength"], kmer["runCutOff"], kmer.get("kmersByPosition", True),
                                    kmer.get("expandKmerPos", False), kmer.get("expandKmerSize", 0))
        try:
            got = al.align_profile(refIdx, readIdx, checkKmer=bool(flags & 1), usingQuality=bool(flags & 2),
                                   doingMatchQuality=bool(flags & 4), gapCap=gapCap)
            want = oracle.align_profile_batch(params, bases, quals, offsets, refIdx, readIdx, flags, kmaps=km, gapCap=gapCap)
        finally:
            if km:
                oracle.free_kmaps(km)
    return got, want


@pytest.mark.parametrize("pname", sorted(H.PARAM_SETS))
@pytest.mark.parametrize("length", [37, 150, 250, 301])
def test_align_profile_matches_oracle(oracle, pname, length):
    rng = np.random.default_rng(length * 131 + len(pname))
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 48, length)
    n = 400
    refIdx = rng.integers(0, 48, n).astype(np.uint32)
    readIdx = rng.integers(0, 48, n).astype(np.uint32)
    (got, gg), (want, wg) = run_both(oracle, H.params_for(pname), bases, quals, offsets, cnts, refIdx, readIdx,
                                     _abi.USING_QUALITY, gapCap=24)
    H.assert_comparisons_equal(got, want)
    assert np.array_equal(gg, wg)
    for name in _abi.FLOAT_FIELDS:  # expected stronger than the 1e-6 bar: identical bits
        assert np.array_equal(got[name], want[name]), name


def test_unrelated_sequences_and_ragged_lengths(oracle):
    rng = np.random.default_rng(11)
    seqs = [H.rand_seq(rng, int(n)) for n in rng.integers(1, 330, 64)]
    seqs[0] = H.rand_seq(rng, 1)
    seqs[1] = H.rand_seq(rng, 2)
    quals = [H.rand_quals(rng, len(s)) for s in seqs]
    bases, q, offsets = H.np.concatenate(seqs), np.concatenate(quals), np.zeros(65, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    refIdx, readIdx = np.meshgrid(np.arange(64, dtype=np.uint32), np.arange(64, dtype=np.uint32))
    refIdx, readIdx = refIdx.ravel(), readIdx.ravel()
    for pname in ("qluster", "asym", "countEnds"):
        (got, gg), (want, wg) = run_both(oracle, H.params_for(pname), bases, q, offsets, None, refIdx, readIdx,
                                         _abi.USING_QUALITY | _abi.MATCH_QUALITY, gapCap=8)
        H.assert_comparisons_equal(got, want)
        assert np.array_equal(gg, wg)
        assert (got["status"] & _abi.ST_GAPLIST_TRUNCATED).any()  # unrelated pairs overflow 8 runs


def test_check_kmer_matches_oracle(oracle):
    rng = np.random.default_rng(21)
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 64, 250, lowFrac=0.05)
    n = 600
    refIdx = rng.integers(0, 8, n).astype(np.uint32)
    readIdx = rng.integers(0, 64, n).astype(np.uint32)
    for kw in (dict(kLength=9, runCutOff=1), dict(kLength=9, runCutOff=12, expandKmerPos=True, expandKmerSize=2),
               dict(kLength=5, runCutOff=30, kmersByPosition=False)):
        got, want = run_both(oracle, H.params_for("qluster"), bases, quals, offsets, cnts, refIdx, readIdx,
                             _abi.USING_QUALITY | _abi.CHECK_KMER, kmer=kw)
        H.assert_comparisons_equal(got, want)
        assert got["lowKmerMismatches"].sum() > 0


def test_degenerate_bases_and_n(oracle):
    rng = np.random.default_rng(31)
    alpha = np.frombuffer(b"ACGTNRYacgt", dtype=np.uint8)
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 32, 120, alphabet=alpha)
    refIdx = rng.integers(0, 32, 300).astype(np.uint32)
    readIdx = rng.integers(0, 32, 300).astype(np.uint32)
    for pname in ("qluster", "degenHomopolymer"):
        got, want = run_both(oracle, H.params_for(pname), bases, quals, offsets, cnts, refIdx, readIdx, _abi.USING_QUALITY,
                             kmer=None)
        H.assert_comparisons_equal(got, want)


def test_kmer_index_lookup(oracle):
    rng = np.random.default_rng(41)
    bases, quals, offsets, cnts, seqs, _ = H.make_store(rng, 40, 90)
    for kw in (dict(k=9, cut=1, byPos=True, ex=False, es=0), dict(k=7, cut=1, byPos=True, ex=True, es=3),
               dict(k=6, cut=1, byPos=False, ex=False, es=0)):
        km = oracle.index_kmers(bases, offsets, cnts, kw["k"], kw["cut"], kw["byPos"], kw["ex"], kw["es"])
        with GpuAligner(make_params()) as al:
            al.upload_seqs(bases, quals, offsets, cnts)
            al.index_kmers(kw["k"], kw["cut"], kw["byPos"], kw["ex"], kw["es"])
            kmers, pos = [], []
            for _ in range(500):
                s = seqs[int(rng.integers(40))]
                p = int(rng.integers(0, len(s) - kw["k"] + 1))
                kmers.append(s[p:p + kw["k"]].tobytes())
                pos.append(max(0, p + int(rng.integers(-4, 5))))
            got = al.kmer_index_lookup(kmers, pos)
        want = np.array([oracle.kmaps_count(km, k, p) for k, p in zip(kmers, pos)], np.uint32)
        oracle.free_kmaps(km)
        assert np.array_equal(got, want)


def test_kmer_compare_matches_oracle(oracle):
    rng = np.random.default_rng(51)
    bases, quals, offsets, cnts, seqs, _ = H.make_store(rng, 50, 200, jitter=40, family=False)
    fam = H.make_store(rng, 30, 200)
    seqs = seqs + fam[4]
    b2, q2, off2 = pack_seqs([s.tobytes() for s in seqs], [np.zeros(len(s), np.uint8) for s in seqs])
    with GpuAligner(make_params()) as al:
        al.upload_seqs(b2, q2, off2)
        for k in (3, 5, 7):
            al.build_kmer_profiles(k)
            a = rng.integers(0, len(seqs), 400).astype(np.uint32)
            b = rng.integers(0, len(seqs), 400).astype(np.uint32)
            shared, frac = al.compare_kmers(a, b)
            for i in range(400):
                ws, wf = oracle.kmer_compare(seqs[a[i]].tobytes(), seqs[b[i]].tobytes(), k)
                assert shared[i] == ws and abs(frac[i] - wf) <= 1e-6 * max(abs(wf), 1e-300) and frac[i] == wf
            reads = np.arange(50, 80, dtype=np.uint32)
            clus = np.arange(0, 12, dtype=np.uint32)
            sh, fr = al.compare_kmers_all(reads, clus)
            for i, r in enumerate(reads):
                for j, c in enumerate(clus):
                    ws, wf = oracle.kmer_compare(seqs[r].tobytes(), seqs[c].tobytes(), k)
                    assert sh[i, j] == ws and fr[i, j] == wf


def test_kmer_compare_revcomp_matches_oracle(oracle):
    """kmerInfo(seq, k, true) + compareKmersRevComp (extractorPairedEnd.cpp:819-824): a's k-mers against the k-mers of
    b's reverse complement, identical to the oracle (uint32 count, bit-identical double) for A,C,G,T stores, stores
    with N and IUPAC codes; the forward comparison is untouched; a character without a complement is refused."""
    rng = np.random.default_rng(31)
    comp = bytes.maketrans(b"ACGTNRYKM", b"TGCANYRMK")
    for alphabet, ks in ((b"ACGT", (3, 5, 7)), (b"ACGTN", (4,)), (b"ACGTNRYKM", (3,))):
        al_ = np.frombuffer(alphabet, np.uint8)
        seqs = [H.rand_seq(rng, int(n), al_) for n in rng.integers(30, 260, 50)]
        # mates: reverse complements of mutated copies, so that the rev-comp comparison has something to find
        seqs += [np.frombuffer(H.mutate(rng, s, 2).tobytes().translate(comp)[::-1], np.uint8) for s in seqs[:30]]
        b, q, off = pack_seqs([s.tobytes() for s in seqs], [np.zeros(len(s), np.uint8) for s in seqs])
        a = np.concatenate([np.arange(30), rng.integers(0, len(seqs), 200)]).astype(np.uint32)
        c = np.concatenate([np.arange(50, 80), rng.integers(0, len(seqs), 200)]).astype(np.uint32)
        with GpuAligner(make_params()) as al:
            al.upload_seqs(b, q, off)
            for k in ks:
                al.build_kmer_profiles(k)
                fwd = al.compare_kmers(a, c)
                rev = al.compare_kmers(a, c, revCompB=True)
                for i in range(len(a)):
                    ws, wf = oracle.kmer_compare(seqs[a[i]].tobytes(), seqs[c[i]].tobytes(), k)
                    rs, rf = oracle.kmer_compare(seqs[a[i]].tobytes(), seqs[c[i]].tobytes(), k, revCompB=True)
                    assert (fwd[0][i], fwd[1][i]) == (ws, wf) and (rev[0][i], rev[1][i]) == (rs, rf), (alphabet, k, i)
                if k >= 5:
                    assert rev[1][:30].mean() > 0.5 > fwd[1][:30].mean()  # the mates only match in reverse complement
    with GpuAligner(make_params()) as al:
        al.upload_seqs(*pack_seqs(["ACGTXACGTT", "ACGTTACGTA"], [np.zeros(10, np.uint8)] * 2))
        al.build_kmer_profiles(3)
        with pytest.raises(SdgpuError) as ei:
            al.compare_kmers([0], [1], revCompB=True)
        assert ei.value.code == _abi.E_ALPHABET


def test_cxx_gapped_strings_from_gap_lists(oracle):
    """examples/gapped_strings_cabi.cpp (INTEGRATION level 2, compiled): alignObjectA_ / alignObjectB_ rebuilt from the gap
    lists the C ABI returns must be the oracle's rearrangeGlobal strings -- internal gaps, end gaps, several runs."""
    import subprocess
    import __graft_entry__ as g
    exe = g.build_example("gapped_strings_cabi")
    rng = np.random.default_rng(12)
    base = H.rand_seq(rng, 90)
    pairs = [(base, H.mutate(rng, base, 2, 1, 1)), (base, base[7:]), (base[:-9], base), (H.mutate(rng, base, 0, 2, 2, hp=True), base),
             (base, np.concatenate([H.rand_seq(rng, 5), base[:40], base[46:]]))]
    args = []
    for a, b in pairs:
        args += [a.tobytes().decode(), b.tobytes().decode()]
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln.split("\t") for ln in r.stdout.strip().splitlines()]
    assert len(lines) == len(pairs)
    params = make_params()
    for (a, b), (score, alnA, alnB) in zip(pairs, lines):
        comp, gaps, wa, wb = oracle.align_profile(params, a.tobytes(), np.full(len(a), 30, np.uint8), b.tobytes(), np.full(len(b), 30, np.uint8))
        assert int(score) == comp.alnScore and alnA == wa and alnB == wb
    assert any("-" in ln[1] for ln in lines) and any("-" in ln[2] for ln in lines)


def test_append_and_errors():
    p = make_params()
    with GpuAligner(p) as al:
        b, q, off = pack_seqs(["ACGTACGT", "ACGAACGT"], [np.full(8, 30, np.uint8)] * 2)
        al.upload_seqs(b, q, off)
        first = al.append_seqs(*pack_seqs(["ACGTTACGT"], [np.full(9, 30, np.uint8)]))
        assert first == 2 and al.num_seqs() == 3
        out = al.align_profile([0, 0], [1, 2])
        assert out["alnScore"][0] == 2 * 7 - 2 and out["oneBaseIndel"][1] == 1
        with pytest.raises(SdgpuError) as ei:
            al.align_profile([0], [7])
        assert ei.value.code == _abi.E_ARG
        with pytest.raises(SdgpuError) as ei:
            al.align_profile([0], [1], checkKmer=True)
        assert ei.value.code == _abi.E_STATE
        long = "ACGT" * 200  # 800 columns: strip-mined in two column blocks
        al.append_seqs(*pack_seqs([long], [np.full(800, 30, np.uint8)]))
        out = al.align_profile([0, 3], [3, 0])
        assert out["alnLen"][0] == 800 and out["alnLen"][1] == 800


def test_rejected_upload_leaves_the_store_and_alphabet_alone(oracle):
    """A batch that fails validation (17th distinct character, byte >= 128, empty sequence) must not change the
    ctx: the previous store, its alphabet codes and scoring stay usable and later uploads of characters that
    the rejected batch introduced still get correct scores."""
    p = H.params_for("degenHomopolymer")
    seqs = ["ACGTRYACGTTGCA", "ACGTRYACGATGCA"]
    b, q, off = pack_seqs(seqs, [np.full(len(s), 35, np.uint8) for s in seqs])
    with GpuAligner(p) as al:
        al.upload_seqs(b, q, off)
        before = al.align_profile([0], [1])
        many = "ACGTNRYSWKMBDHVXZ"  # 12 new characters on top of A,C,G,T,N,R,Y: the 17th does not fit
        for bad in (pack_seqs([many], [np.full(len(many), 30, np.uint8)]),
                    (np.frombuffer(b"ACG\xc8", np.uint8).copy(), np.full(4, 30, np.uint8), np.array([0, 4], np.uint64))):
            with pytest.raises(SdgpuError) as ei:
                al.upload_seqs(*bad)
            assert ei.value.code == _abi.E_ALPHABET
            with pytest.raises(SdgpuError):
                al.append_seqs(*bad)
        with pytest.raises(SdgpuError):
            al.upload_seqs(np.frombuffer(b"ACGT", np.uint8).copy(), np.full(4, 30, np.uint8), np.array([0, 4, 4], np.uint64))
        assert al.num_seqs() == 2
        after = al.align_profile([0], [1])
        assert after.tobytes() == before.tobytes()
        # S, W, K were named by the rejected batch: they must score like the matrix says once they do arrive
        extra = ["ACGTSWACGTKGCA", "ACGTGAACGTTGCA"]
        al.append_seqs(*pack_seqs(extra, [np.full(14, 35, np.uint8)] * 2))
        got = al.align_profile([2, 0], [3, 2])
    allSeqs = seqs + extra
    bb, qq, oo = pack_seqs(allSeqs, [np.full(len(s), 35, np.uint8) for s in allSeqs])
    want = oracle.align_profile_batch(p, bb, qq, oo, np.array([2, 0], np.uint32), np.array([3, 2], np.uint32))
    H.assert_comparisons_equal(got, want)


def test_kmer_index_size_limit():
    """The index build sorts with a signed 32-bit item count: 2^31 or more (position, k-mer) occurrences are refused
    loudly instead of silently producing an empty table (every mismatch would turn into a lowKmer mismatch)."""
    rng = np.random.default_rng(3)
    seqs = [H.rand_seq(rng, 300).tobytes() for _ in range(6000)]
    b, q, off = pack_seqs(seqs, [np.full(300, 30, np.uint8)] * len(seqs))
    with GpuAligner(make_params()) as al:
        al.upload_seqs(b, q, off)
        with pytest.raises(SdgpuError) as ei:
            al.index_kmers(9, 1, True, expandKmerPos=True, expandKmerSize=1300)  # 6000 x 292 x ~1450 keys > 2^31
        assert ei.value.code == _abi.E_UNSUPPORTED and "2^31" in str(ei.value)
        al.index_kmers(9, 1, True, expandKmerPos=True, expandKmerSize=2)
        assert al.kmer_index_lookup([seqs[0][5:14]], [4])[0] >= 1


def test_full_size_properties():
    """BASELINE config-2 shaped batch (300 bp) without the oracle: size-independent properties."""
    rng = np.random.default_rng(61)
    bases, quals, offsets, cnts, seqs, _ = H.make_store(rng, 2000, 300)
    n = 40000
    a = rng.integers(0, 2000, n).astype(np.uint32)
    b = rng.integers(0, 2000, n).astype(np.uint32)
    b[:200] = a[:200]
    with GpuAligner(make_params()) as al:
        al.upload_seqs(bases, quals, offsets, cnts)
        out = al.align_profile(a, b)
        again = al.align_profile(a, b)
    assert out.tobytes() == again.tobytes()  # deterministic / idempotent
    lens = np.diff(offsets).astype(np.int64)
    assert np.array_equal(out["alnScore"][:200], 2 * lens[a[:200]])  # self-alignment
    assert (out["hqMismatches"][:200] + out["lqMismatches"][:200] + out["numGaps"][:200] == 0).all()
    assert (out["alnLen"] >= np.maximum(lens[a], lens[b])).all() and (out["alnLen"] <= lens[a] + lens[b]).all()
    cols = out["highQualityMatches"] + out["lowQualityMatches"] + out["hqMismatches"] + out["lqMismatches"] + out["lowKmerMismatches"]
    assert (cols <= np.minimum(lens[a], lens[b])).all()
    assert (out["alnScore"] <= 2 * np.minimum(lens[a], lens[b])).all()
    assert ((out["eventBasedIdentity"] >= 0) & (out["eventBasedIdentity"] <= 1)).all()


def test_align_pass_matches_pass_error_profile(oracle):
    """On-device comparison::passErrorProfile over a par table == thresholds applied on the host
    to the oracle's errorProfiles (rows a5/a6)."""
    from seekdeep_b200.iterpars import CollapseIterations, IterPar
    rng = np.random.default_rng(71)
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 64, 200)
    n = 800
    refIdx = rng.integers(0, 64, n).astype(np.uint32)
    readIdx = rng.integers(0, 64, n).astype(np.uint32)
    iters = CollapseIterations.gen_illumina_default_pars(100).iters[:11]
    iters += [IterPar(100, 3, 0.99, 0.5, 0.25, 1, 2, 0), IterPar(100, 0, onPerId=True, onHqPerId=True, perId=0.985),
              IterPar(100, 0, onPerId=True, onHqPerId=False, perId=0.97)]
    params = H.params_for("countEnds")  # weighHomopolymers: fractional indel tallies
    with GpuAligner(params) as al:
        al.upload_seqs(bases, quals, offsets, cnts)
        al.set_pass_table(iters)
        masks = al.align_pass(refIdx, readIdx)
    want = oracle.align_profile_batch(params, bases, quals, offsets, refIdx, readIdx, _abi.USING_QUALITY)
    for l, it in enumerate(iters):
        if it.onPerId:
            ok = (want["eventBasedIdentityHq"] if it.onHqPerId else want["eventBasedIdentity"]) >= it.perId
        else:
            ok = ((it.oneBaseIndel >= want["oneBaseIndel"]) & (it.twoBaseIndel >= want["twoBaseIndel"]) &
                  (it.largeBaseIndel >= want["largeBaseIndel"]) & (it.hqMismatches >= want["hqMismatches"]) &
                  (it.lqMismatches >= want["lqMismatches"]) & (it.lowKmerMismatches >= want["lowKmerMismatches"]))
        got = ((masks >> np.uint64(l)) & np.uint64(1)).astype(bool)
        assert np.array_equal(got, ok), l
    assert (masks >> np.uint64(len(iters)) == 0).all()
    assert 0 < int(((masks & np.uint64(1)) != 0).sum()) < n


def test_all_vs_all_matches_oracle(oracle):
    """Config-3 shaped case at oracle-checkable size: every unordered pair of 70 related 400-mers,
    computed as two disjoint pair blocks (the multi-GPU pair-block sharding) + a seqIdx subset."""
    from tests import synth
    seqs, quals = synth.config3(seed=3, n=70, length=400, nRoots=4)
    bases, q, offsets = synth.pack(seqs, quals)
    n = len(seqs)
    ii, jj = np.triu_indices(n, 1)
    params = make_params()
    want = oracle.align_profile_batch(params, bases, q, offsets, ii.astype(np.uint32), jj.astype(np.uint32), _abi.USING_QUALITY)
    with GpuAligner(params) as al:
        al.upload_seqs(bases, q, offsets)
        total = n * (n - 1) // 2
        cut = total // 3
        got = np.concatenate([al.all_vs_all(firstPair=0, nPairs=cut), al.all_vs_all(firstPair=cut, nPairs=total - cut)])
        sub = np.array([5, 9, 2, 40, 33], np.uint32)
        gsub = al.all_vs_all(seqIdx=sub)
        with pytest.raises(SdgpuError):
            al.all_vs_all(firstPair=total, nPairs=1)
    assert len(got) == total
    for f in ("alnScore", "hqMismatches", "lqMismatches", "lowKmerMismatches", "numGaps"):
        assert np.array_equal(got[f].astype(np.int64), want[f].astype(np.int64)), f
    for f in ("oneBaseIndel", "twoBaseIndel", "largeBaseIndel", "eventBasedIdentity"):
        assert np.array_equal(got[f], want[f].astype(np.float32)), f  # exactly the double rounded to float ...
        w = want[f]
        assert (np.abs(got[f].astype(np.float64) - w) <= 1e-6 * np.abs(w)).all(), f  # ... i.e. within the contract's 1e-6 relative
    si, sj = np.triu_indices(len(sub), 1)
    wsub = oracle.align_profile_batch(params, bases, q, offsets, sub[si], sub[sj], _abi.USING_QUALITY)
    assert np.array_equal(gsub["alnScore"], wsub["alnScore"]) and np.array_equal(gsub["hqMismatches"], wsub["hqMismatches"])


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("pname", ["qluster", "asym", "countEnds", "stitch"])
def test_both_forward_kernels_match_oracle(oracle, mode, pname):
    """The int32 forward pass (mode 1) and the packed fp16 two-pairs-per-warp pass (mode 2) are both
    exact: same scores, gap lists and errorProfiles as the oracle on ragged ACGT(N) inputs."""
    rng = np.random.default_rng(1000 + mode)
    alpha = np.frombuffer(b"ACGTACGTACGTACGTN", dtype=np.uint8)
    seqs = []
    for fam in range(4):
        # 'asym' extends gaps at 2 per base: keep its fp16 value range exact by using shorter reads
        root = H.rand_seq(rng, int(rng.integers(60, 150 if pname == "asym" else 330)), alpha)
        seqs.append(root)
        for _ in range(11):
            seqs.append(H.mutate(rng, root, int(rng.integers(0, 6)), int(rng.integers(0, 2)), int(rng.integers(0, 2)), hp=True))
    seqs += [H.rand_seq(rng, 1), H.rand_seq(rng, 2), H.rand_seq(rng, 3)]
    quals = [H.rand_quals(rng, len(s)) for s in seqs]
    bases, q, offsets = np.concatenate(seqs), np.concatenate(quals), np.zeros(len(seqs) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    n = 701  # odd: the packed kernel's last work item holds a single pair
    refIdx = rng.integers(0, len(seqs), n).astype(np.uint32)
    readIdx = rng.integers(0, len(seqs), n).astype(np.uint32)
    params = H.params_for(pname)
    with GpuAligner(params) as al:
        al.set_kernel(mode)
        al.upload_seqs(bases, q, offsets)
        al.index_kmers(7, 2, True)
        got, gg = al.align_profile(refIdx, readIdx, checkKmer=True, usingQuality=True, gapCap=40)
        assert al.last_kernel() == mode
    km = oracle.index_kmers(bases, offsets, None, 7, 2, True)
    want, wg = oracle.align_profile_batch(params, bases, q, offsets, refIdx, readIdx, _abi.USING_QUALITY | _abi.CHECK_KMER,
                                          kmaps=km, gapCap=40)
    oracle.free_kmaps(km)
    H.assert_comparisons_equal(got, want)
    assert np.array_equal(gg, wg)


def test_packed_kernel_refuses_out_of_range():
    """fp16 exactness is a precondition, checked on the host: large scores / extra symbols fall back
    to the int32 pass automatically and fail loudly when the packed pass is forced."""
    from seekdeep_b200.scoring import SubstituteMatrix
    rng = np.random.default_rng(5)
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 8, 120)
    idx = np.arange(8, dtype=np.uint32)
    with GpuAligner(make_params(scoring=SubstituteMatrix.create_score_matrix(20, -20))) as al:
        al.upload_seqs(bases, quals, offsets, cnts)
        al.align_profile(idx, idx[::-1].copy())
        assert al.last_kernel() == 1
        al.set_kernel(2)
        with pytest.raises(SdgpuError) as ei:
            al.align_profile(idx, idx[::-1].copy())
        assert ei.value.code == _abi.E_UNSUPPORTED
    with GpuAligner(make_params()) as al:
        al.upload_seqs(bases, quals, offsets, cnts)
        al.align_profile(idx, idx[::-1].copy())
        assert al.last_kernel() == 2
        al.append_seqs(*pack_seqs(["ACGTRYACGT"], [np.full(10, 30, np.uint8)]))  # R, Y: beyond A,C,G,T,N
        al.align_profile(idx, idx[::-1].copy())
        assert al.last_kernel() == 1


def test_wide_reads_strip_mined(oracle):
    """Reads wider than 512 columns run through the strip-mined int32 pass (column blocks of 512 with
    edge buffers between them): same scores, gap lists and errorProfiles as the oracle."""
    rng = np.random.default_rng(77)
    seqs = []
    for length in (530, 1024, 1025, 1500):
        root = H.rand_seq(rng, length)
        seqs.append(root)
        for _ in range(3):
            seqs.append(H.mutate(rng, root, int(rng.integers(0, 12)), int(rng.integers(0, 3)), int(rng.integers(0, 3)), hp=True))
    seqs.append(H.rand_seq(rng, 40))
    seqs.append(H.rand_seq(rng, 700))
    quals = [H.rand_quals(rng, len(s)) for s in seqs]
    bases, q, offsets = np.concatenate(seqs), np.concatenate(quals), np.zeros(len(seqs) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    n = len(seqs)
    refIdx, readIdx = np.meshgrid(np.arange(n, dtype=np.uint32), np.arange(n, dtype=np.uint32))
    refIdx, readIdx = refIdx.ravel(), readIdx.ravel()
    for pname in ("qluster", "asym"):
        params = H.params_for(pname)
        with GpuAligner(params) as al:
            al.upload_seqs(bases, q, offsets)
            got, gg = al.align_profile(refIdx, readIdx, gapCap=64)
            assert al.last_kernel() == 1
        want, wg = oracle.align_profile_batch(params, bases, q, offsets, refIdx, readIdx, _abi.USING_QUALITY, gapCap=64)
        H.assert_comparisons_equal(got, want)
        assert np.array_equal(gg, wg)


def test_random_parameter_fuzz(oracle):
    """Differential fuzz over the whole parameter space of the boundary: random gap penalties (all ten,
    including extend > open and zeros), random score matrices (asymmetric, positive mismatches),
    random quality thresholds / windows / flags, ragged related and unrelated sequences."""
    from seekdeep_b200.scoring import GapScoringParameters, SubstituteMatrix
    rng = np.random.default_rng(20261019)
    for trial in range(24):
        gp = GapScoringParameters(*[int(x) for x in rng.integers(0, 9, 10)])
        m = np.full((128, 128), int(rng.integers(-4, 0)), np.int32)
        letters = b"ACGTN"
        sub = rng.integers(-4, 5, (5, 5))
        for a in range(5):
            for b in range(5):
                m[letters[a], letters[b]] = sub[a, b]
            m[letters[a], letters[a]] = int(rng.integers(1, 5))
        params = make_params(gaps=gp, scoring=SubstituteMatrix(m), primaryQual=int(rng.integers(0, 41)),
                             secondaryQual=int(rng.integers(0, 41)), qualThresWindow=int(rng.integers(0, 6)),
                             countEndGaps=bool(rng.integers(2)), weighHomopolymers=bool(rng.integers(2)))
        alpha = np.frombuffer(b"ACGT" * 6 + b"N", dtype=np.uint8)
        seqs = []
        for fam in range(3):
            root = H.rand_seq(rng, int(rng.integers(5, 140)), alpha)
            seqs.append(root)
            for _ in range(6):
                seqs.append(H.mutate(rng, root, int(rng.integers(0, 5)), int(rng.integers(0, 3)), int(rng.integers(0, 3)),
                                     hp=bool(rng.integers(2))))
        quals = [rng.integers(0, 42, len(s)).astype(np.uint8) for s in seqs]
        bases, q, offsets = np.concatenate(seqs), np.concatenate(quals), np.zeros(len(seqs) + 1, np.uint64)
        offsets[1:] = np.cumsum([len(s) for s in seqs])
        cnts = rng.integers(1, 9, len(seqs)).astype(np.float64)
        n = 150
        refIdx = rng.integers(0, len(seqs), n).astype(np.uint32)
        readIdx = rng.integers(0, len(seqs), n).astype(np.uint32)
        flags = int(rng.integers(0, 8))
        kmer = dict(kLength=int(rng.integers(2, 8)), runCutOff=int(rng.integers(0, 6)), kmersByPosition=bool(rng.integers(2)),
                    expandKmerPos=bool(rng.integers(2)), expandKmerSize=int(rng.integers(0, 4)))
        for mode in (1, 0):  # int32 pass, then automatic (packed fp16 whenever its range check allows)
            with GpuAligner(params) as al:
                al.set_kernel(mode)
                al.upload_seqs(bases, q, offsets, cnts)
                al.index_kmers(**kmer)
                got, gg = al.align_profile(refIdx, readIdx, checkKmer=bool(flags & 1), usingQuality=bool(flags & 2),
                                           doingMatchQuality=bool(flags & 4), gapCap=30)
            if mode == 1:
                km = oracle.index_kmers(bases, offsets, cnts, kmer["kLength"], kmer["runCutOff"], kmer["kmersByPosition"],
                                        kmer["expandKmerPos"], kmer["expandKmerSize"])
                want, wg = oracle.align_profile_batch(params, bases, q, offsets, refIdx, readIdx, flags, kmaps=km, gapCap=30)
                oracle.free_kmaps(km)
            H.assert_comparisons_equal(got, want)
            assert np.array_equal(gg, wg), (trial, mode)


@pytest.mark.parametrize("pname", ["qluster", "degenHomopolymer", "countEnds", "asym"])
@pytest.mark.parametrize("length", [90, 250, 300, 600])
def test_chimera_probe_matches_oracle(oracle, pname, length):
    """sdgpu_chimera_probe_pairs vs the oracle's markChimeras probe (windowed profileAlignment):
    kept-mismatch counts, first/last positions and the front/back chiOverlap verdicts, with and
    without low-quality mismatches kept, loose and strict overlap thresholds."""
    from seekdeep_b200.qluster import IterParC
    rng = np.random.default_rng(length + len(pname))
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 40, length, lowFrac=0.25)
    n = 600
    par = rng.integers(0, 40, n).astype(np.uint32)
    cand = rng.integers(0, 40, n).astype(np.uint32)
    params = H.params_for(pname)
    seen = set()
    with GpuAligner(params) as al:
        al.upload_seqs(bases, quals, offsets, cnts)
        for ov in (IterParC(0, 0, 2.0, 1.0, 0.99, 0.0, 0.0, 1.0, 0, 0, 0.0), IterParC(0, 0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0, 0, 0.0),
                   IterParC(0, 0, 0.6, 5.0, 5.0, 0.0, 3.0, 0.0, 0, 0, 0.0)):
            for keep in (False, True):
                got = al.chimera_probe(par, cand, ov, keepLowQualityMismatches=keep)
                want = oracle.chimera_probe_batch(params, bases, quals, offsets, par, cand, ov, keepLowQual=keep)
                for name in got.dtype.names:
                    bad = np.nonzero(got[name] != want[name])[0]
                    assert bad.size == 0, f"{name} keep={keep}: pair {bad[0]} got {got[name][bad[0]]} want {want[name][bad[0]]}"
                seen |= set(np.unique(got["flags"]).tolist())
    assert seen == {0, 1, 2, 3}  # every verdict combination was exercised


def test_kmer_profiles_ragged_lengths_both_paths(oracle):
    """Profile build over lengths 1..1300 (one to four words per lane, several 512-position rounds,
    sequences shorter than k) for the ACGT-only 2-bit path and, with an N in the store, the general
    sigma^k path: self-comparison must equal len-k+1 and random pairs must match the oracle."""
    rng = np.random.default_rng(77)
    lens = np.concatenate([np.arange(1, 40), rng.integers(40, 700, 60), [127, 128, 129, 131, 132, 133, 256, 257, 259, 260,
                                                                        384, 385, 388, 389, 511, 512, 513, 516, 517, 1023, 1300]])
    for withN in (False, True):
        seqs = [H.rand_seq(rng, int(n)) for n in lens]
        rel = [H.mutate(rng, s, 3, 1, 1) for s in seqs[40:80]]  # related copies: many shared k-mers
        seqs = seqs + rel
        if withN:
            seqs[50] = seqs[50].copy()
            seqs[50][len(seqs[50]) // 2] = ord("N")
        b, q, off = pack_seqs([s.tobytes() for s in seqs], [np.zeros(len(s), np.uint8) for s in seqs])
        n = len(seqs)
        with GpuAligner(make_params()) as al:
            al.upload_seqs(b, q, off)
            for k in (2, 4, 5, 6, 7) if not withN else (3, 5):
                al.build_kmer_profiles(k)
                me = np.arange(n, dtype=np.uint32)
                shared, frac = al.compare_kmers(me, me)
                L = np.array([len(s) for s in seqs])
                assert np.array_equal(shared, np.where(L >= k, L - k + 1, 0))
                a = np.concatenate([rng.integers(0, n, 300), np.arange(40, 80)]).astype(np.uint32)
                c = np.concatenate([rng.integers(0, n, 300), np.arange(len(lens), len(lens) + 40)]).astype(np.uint32)
                shared, frac = al.compare_kmers(a, c)
                for i in range(len(a)):
                    ws, wf = oracle.kmer_compare(seqs[a[i]].tobytes(), seqs[c[i]].tobytes(), k)
                    assert shared[i] == ws and frac[i] == wf, (withN, k, i, len(seqs[a[i]]), len(seqs[c[i]]))


@pytest.mark.parametrize("pname", ["qluster", "allEndsCost", "allEndsFree", "asym", "countEnds", "degenHomopolymer"])
def test_banded_pass_is_exact(oracle, pname):
    """The banded first pass + redo must return exactly what the full-matrix kernels and the oracle
    return: near-identical pairs (certified in the band), pairs with indels of 1..45 bases and length
    differences beyond the band (redone), low-complexity sequences where ties abound."""
    rng = np.random.default_rng(len(pname) * 7 + 3)
    seqs = []
    root = H.rand_seq(rng, 280)
    lowc = np.frombuffer(("ACACACACAC" * 8 + "AAAAAAAAAATTTTTTTTTT" * 6 + "GATTACA" * 12).encode(), np.uint8).copy()
    for base in (root, lowc):
        seqs.append(base)
        for _ in range(40):  # SNP / small-indel relatives
            seqs.append(H.mutate(rng, base, int(rng.integers(0, 8)), int(rng.integers(0, 2)), int(rng.integers(0, 2)), hp=True))
        for size in (1, 5, 12, 20, 27, 30, 31, 33, 38, 45):  # one long deletion / insertion at a random place
            p = int(rng.integers(20, len(base) - 60))
            seqs.append(np.concatenate([base[:p], base[p + size:]]))
            seqs.append(np.concatenate([base[:p], H.rand_seq(rng, size), base[p:]]))
        for shift in (3, 17, 29, 40):  # end-shifted copies: long end gaps
            seqs.append(base[shift:])
            seqs.append(np.concatenate([H.rand_seq(rng, shift), base]))
    quals = [H.rand_quals(rng, len(s)) for s in seqs]
    bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
    n = len(seqs)
    refIdx, readIdx = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(n), np.arange(n))]
    params = H.params_for(pname)
    with GpuAligner(params) as al:
        al.upload_seqs(bases, q, offsets)
        al.profile_enable(True)
        al.profile_read(reset=True)
        got, gg = al.align_profile(refIdx, readIdx, usingQuality=True, gapCap=12)
        prof = al.profile_read()
        al.set_band(1)
        full, fg = al.align_profile(refIdx, readIdx, usingQuality=True, gapCap=12)
        profFull = al.profile_read()
    assert prof["bandCertifiedCells"] > 0.15 * prof["cellUpdates"]     # the band finished a good share ...
    assert prof["bandCertifiedCells"] < prof["cellUpdates"]            # ... and some pairs were redone in full
    assert prof["evaluatedCells"] < prof["cellUpdates"] and profFull["evaluatedCells"] == profFull["cellUpdates"]
    assert profFull["bandCells"] == 0
    for name in got.dtype.names:
        assert np.array_equal(got[name], full[name]), name
    assert np.array_equal(gg, fg)
    sel = rng.choice(len(refIdx), 1500, replace=False)
    want, wg = oracle.align_profile_batch(params, bases, q, offsets, refIdx[sel], readIdx[sel], _abi.USING_QUALITY, gapCap=12)
    H.assert_comparisons_equal(got[sel], want)
    assert np.array_equal(gg[sel], wg)


def test_banded_pass_edge_geometries(oracle):
    """Band edge cases: 1..40-base sequences (no interior steps at all), length differences around the
    +-30 limit, near-identical 2.5 kb pairs (longer than the shared-memory staging: general cells with global
    loads), and a launch of dissimilar pairs followed by similar ones (the skip-the-band heuristic)."""
    rng = np.random.default_rng(99)
    params = H.params_for("qluster")
    seqs = [H.rand_seq(rng, int(n)) for n in list(range(1, 41)) + [64, 65, 66, 97, 128]]
    base = H.rand_seq(rng, 200)
    for d in (28, 29, 30, 31, 32, 40):          # length difference at / beyond the band's limit
        seqs.append(base[d:])
        seqs.append(np.concatenate([base, H.rand_seq(rng, d)]))
    seqs.append(base)
    long0 = H.rand_seq(rng, 2500)
    seqs += [long0, H.mutate(rng, long0, 9, 1, 1), H.mutate(rng, long0, 3, 0, 2), long0[7:]]
    quals = [H.rand_quals(rng, len(s)) for s in seqs]
    bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
    n = len(seqs)
    nShort = n - 4
    refIdx, readIdx = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(nShort), np.arange(nShort))]
    longPairs = np.array([(i, j) for i in range(nShort, n) for j in range(nShort, n)], np.uint32)
    longPairs = np.tile(longPairs, (5, 1))      # >= 64 pairs so the launch takes the banded pass
    for ri, rj, mostlyCertified in ((refIdx, readIdx, False), (longPairs[:, 0].copy(), longPairs[:, 1].copy(), True)):
        with GpuAligner(params) as al:  # fresh ctx each: the first launch's many redo pairs would switch the band off
            al.upload_seqs(bases, q, offsets)
            al.profile_enable(True)
            al.profile_read(reset=True)
            got, gg = al.align_profile(ri, rj, usingQuality=True, gapCap=16)
            prof = al.profile_read()
        assert prof["bandCells"] > 0
        want, wg = oracle.align_profile_batch(params, bases, q, offsets, ri, rj, _abi.USING_QUALITY, gapCap=16)
        H.assert_comparisons_equal(got, want)
        assert np.array_equal(gg, wg)
        if mostlyCertified:  # the 2.5 kb relatives all certify
            assert prof["bandCertifiedCells"] > 0.9 * prof["cellUpdates"]
    with GpuAligner(params) as al:
        al.profile_enable(True)
        # dissimilar launches switch the band off for the following ones, a similar launch gets it back later
        unrel = [H.rand_seq(rng, 150) for _ in range(40)]
        b2, q2, o2 = pack_seqs([s.tobytes() for s in unrel], [H.rand_quals(rng, 150) for _ in unrel])
        al.upload_seqs(b2, q2, o2)
        ri, rj = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(40), np.arange(40))]
        keep = ri != rj
        ri, rj = ri[keep], rj[keep]
        want = oracle.align_profile_batch(params, b2, q2, o2, ri, rj, _abi.USING_QUALITY)
        sawSkip = False
        for _ in range(4):
            al.profile_read(reset=True)
            got = al.align_profile(ri, rj, usingQuality=True)
            sawSkip |= al.profile_read()["bandCells"] == 0
            H.assert_comparisons_equal(got, want)
        assert sawSkip


def test_band_not_used_when_bound_is_unsound(oracle):
    """maxSub <= 0 (no positive score) or penalties beyond 4096: the certificate does not apply, the
    launch must go straight to the full-matrix kernels and still match the oracle."""
    rng = np.random.default_rng(5)
    bases, quals, offsets, cnts, _, _ = H.make_store(rng, 24, 120)
    ri = rng.integers(0, 24, 300).astype(np.uint32)
    rj = rng.integers(0, 24, 300).astype(np.uint32)
    from seekdeep_b200.scoring import SubstituteMatrix
    for kw in (dict(scoring=SubstituteMatrix.create_degen_score_matrix_case_insensitive(0, -3)),
               dict(gaps=GapScoringParameters(5000, 1, 5000, 1, 0, 0, 5000, 1, 0, 0))):
        params = make_params(**kw)
        with GpuAligner(params) as al:
            al.upload_seqs(bases, quals, offsets, cnts)
            al.profile_enable(True)
            al.profile_read(reset=True)
            got = al.align_profile(ri, rj, usingQuality=True)
            assert al.profile_read()["bandCells"] == 0
        want = oracle.align_profile_batch(params, bases, quals, offsets, ri, rj, _abi.USING_QUALITY)
        H.assert_comparisons_equal(got, want)


def test_banded_vs_full_matrix_fuzz():
    """Certificate stress test without the oracle: random gap penalties (zeros included), match / mismatch
    scores, lengths, divergence and indel sizes; the banded pass + redo must reproduce the full-matrix
    kernels field for field, gap lists included, and must actually certify pairs in most configurations."""
    from seekdeep_b200.scoring import SubstituteMatrix
    rng = np.random.default_rng(2026)
    usedBand = 0
    for trial in range(240):
        L = int(rng.choice([40, 90, 150, 250, 300, 420, 700]))
        g = [int(x) for x in rng.integers(0, 9, 10)]
        for k in (0, 2, 4, 6, 8):  # extend <= open keeps the parameter sets realistic
            g[k + 1] = min(g[k + 1], g[k])
        match, mismatch = int(rng.integers(1, 6)), -int(rng.integers(1, 7))
        params = make_params(gaps=GapScoringParameters(*g), scoring=SubstituteMatrix.create_degen_score_matrix_case_insensitive(match, mismatch),
                             weighHomopolymers=bool(rng.integers(2)), countEndGaps=bool(rng.integers(2)))
        root = H.rand_seq(rng, L)
        seqs = [root]
        for _ in range(23):
            s = H.mutate(rng, root, int(rng.integers(0, 1 + L // 12)), int(rng.integers(0, 3)), int(rng.integers(0, 3)), hp=bool(rng.integers(2)))
            if rng.random() < 0.3:  # one long indel
                p, size = int(rng.integers(5, max(6, len(s) - 40))), int(rng.integers(2, 36))
                s = np.concatenate([s[:p], s[p + size:]]) if rng.random() < 0.5 else np.concatenate([s[:p], H.rand_seq(rng, size), s[p:]])
            seqs.append(s)
        quals = [H.rand_quals(rng, len(s)) for s in seqs]
        bases, q, offsets = pack_seqs([s.tobytes() for s in seqs], quals)
        ri, rj = [a.ravel().astype(np.uint32) for a in np.meshgrid(np.arange(24), np.arange(24))]
        with GpuAligner(params) as al:
            al.upload_seqs(bases, q, offsets)
            al.profile_enable(True)
            al.profile_read(reset=True)
            got, gg = al.align_profile(ri, rj, usingQuality=True, gapCap=16)
            usedBand += al.profile_read()["bandCertifiedCells"] > 0
            al.set_band(1)
            full, fg = al.align_profile(ri, rj, usingQuality=True, gapCap=16)
        for name in got.dtype.names:
            bad = np.nonzero(got[name] != full[name])[0]
            assert bad.size == 0, (trial, name, g, match, mismatch, L, int(ri[bad[0]]), int(rj[bad[0]]))
        assert np.array_equal(gg, fg), (trial, g, match, mismatch, L)
    assert usedBand >= 160
