"""End-to-end parity of the collapse loop: GPU-driven sdq_run vs the oracle's sequential
qluster on the same seeded reads — cluster count, membership, consensus sequences, consensus
qualities and counts must be identical (north_star: bit-identical clusters and haplotypes)."""
import numpy as np
import pytest

import seekdeep_b200 as sd
from tests import synth
from seekdeep_b200.iterpars import CollapseIterations
from tests import oracle_binding

pytestmark = pytest.mark.gpu


def oracle_opts(p: sd.QlusterPars):
    return p.to_c()  # sdo_qluster_opts has sdq_opts' layout


def assert_same(got, want):
    assert len(got["clusters"]) == len(want["clusters"])
    assert np.array_equal(got["membership"], want["membership"])
    for (gs, gq, gc), (ws, wq, wc) in zip(got["clusters"], want["clusters"]):
        assert gs == ws and gc == wc and np.array_equal(gq, wq)
    assert got["chimeras"] == want["chimeras"]


def run_pair(oracle, params, pars, iters, bases, quals, offsets, init=None):
    with sd.GpuAligner(params) as al:
        got = sd.qluster(al, bases, quals, offsets, pars, iters, init)
    want = oracle.qluster(params, oracle_opts(pars), (init or iters).iters, iters.iters, bases, quals, offsets)
    return got, want


@pytest.mark.parametrize("seed,nReads,length,nHaps,weigh", [(1, 1500, 120, 6, False), (2, 2500, 250, 12, False),
                                                             (3, 1200, 150, 8, True)])
def test_qluster_matches_oracle(oracle, seed, nReads, length, nHaps, weigh):
    rng = np.random.default_rng(seed)
    haps = synth.make_haplotypes(rng, nHaps, length, 1, 5)
    seqs, quals, _ = synth.fast_reads(rng, haps, nReads, 0.004, hpIndelRate=0.002 if weigh else 0.0)
    bases, q, offsets = synth.pack(seqs, quals)
    params = sd.make_params(weighHomopolymers=weigh)
    iters = CollapseIterations.gen_illumina_default_pars(100)
    got, want = run_pair(oracle, params, sd.QlusterPars(), iters, bases, q, offsets)
    assert_same(got, want)
    st = got["stats"]
    assert st["uniqueSeqs"] == want["nUnique"] and st["finalClusters"] == len(want["clusters"])
    assert st["pairsComputed"] > 0 and st["gpuBatches"] > 0 and st["cellUpdates"] > 0
    # sanity of the result itself: the abundant haplotypes are recovered exactly
    top = {s for s, _, c in got["clusters"] if c >= 0.02 * nReads}
    assert top <= {h.tobytes().decode() for h in haps}


def test_qluster_options(oracle):
    rng = np.random.default_rng(9)
    haps = synth.make_haplotypes(rng, 5, 100, 2, 6)
    seqs, quals, _ = synth.fast_reads(rng, haps, 800, 0.006)
    seqs.append(np.frombuffer(b"ACGTACGTAC", np.uint8))  # dropped: len <= smallReadSize
    quals.append(np.full(10, 30, np.uint8))
    bases, q, offsets = synth.pack(seqs, quals)
    params = sd.make_params()
    for pars, iters in (
            (sd.QlusterPars(startWithSingles=True), CollapseIterations.gen_illumina_default_pars(20)),
            (sd.QlusterPars(leaveOutSinglets=True, dontRecalLowFreqMismatchAndReRun=True), CollapseIterations.gen_illumina_default_pars(100)),
            (sd.QlusterPars(checkKmers=False, kmersByPosition=False, runCutOff=3), CollapseIterations.gen_illumina_default_pars(3)),
            (sd.QlusterPars(converge=True), CollapseIterations.gen_illumina_default_pars(4)),
            (sd.QlusterPars(runCutOff=1), CollapseIterations.gen_otu_pars(100, 0.97, True)),
            (sd.QlusterPars(expandKmerPos=True, expandKmerSize=2), CollapseIterations.gen_otu_pars(100, 0.98, False))):
        got, want = run_pair(oracle, params, pars, iters, bases, q, offsets)
        assert_same(got, want)
        assert got["membership"][-1] == np.uint32(0xFFFFFFFF)


def test_qluster_tiny_and_empty(oracle):
    params = sd.make_params()
    iters = CollapseIterations.gen_illumina_default_pars(100)
    s = np.frombuffer(b"ACGTTGCAAGGCTTAACCGGATCGTTAGCAAT", np.uint8)
    for seqs in ([s], [s, s, s], [s, s, s[:-1].copy(), s[:-1].copy()]):
        quals = [np.full(len(x), 35, np.uint8) for x in seqs]
        bases, q, offsets = synth.pack(seqs, quals)
        got, want = run_pair(oracle, params, sd.QlusterPars(), iters, bases, q, offsets)
        assert_same(got, want)
    with sd.GpuAligner(params) as al:
        got = sd.qluster(al, np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert got["clusters"] == [] and got["stats"]["uniqueSeqs"] == 0


def test_qluster_many_equals_single_runs(oracle):
    """Several samples on one GPU through concurrent worker contexts give exactly the per-sample
    results of separate runs (samples never interact, whatever the interleaving)."""
    from seekdeep_b200.qluster import qluster_many
    params = sd.make_params()
    samples = []
    for seed in (11, 12, 13, 14, 15):
        rng = np.random.default_rng(seed)
        haps = synth.make_haplotypes(rng, 5, 110 + seed, 1, 5)
        seqs, quals, _ = synth.fast_reads(rng, haps, 600, 0.005)
        samples.append(synth.pack(seqs, quals))
    many = qluster_many(params, samples, nWorkers=3)
    assert len(many) == len(samples)
    with sd.GpuAligner(params) as al:
        for got, (b, q, o) in zip(many, samples):
            single = sd.qluster(al, b, q, o)
            assert_same(got, single)
    want = oracle.qluster(params, oracle_opts(sd.QlusterPars()), CollapseIterations.gen_illumina_default_pars(100).iters,
                          CollapseIterations.gen_illumina_default_pars(100).iters, *samples[0])
    assert_same(many[0], want)


@pytest.mark.parametrize("name", ["c2_5k", "c1_full", "c2_20k", "c2_20k_hp"])
def test_qluster_streamed_pass_matches_frozen_oracle(name, monkeypatch):
    """The big (read x window) grids are consumed chunk by chunk -- verdicts filed and the sequential pass run over the
    finished reads while the GPU works on the next chunk (sdgpu_align_pass_grid_stream).  At the fixtures' sizes that
    path does not trigger by itself, so the test lowers its thresholds: every whole-window grid is streamed, in chunks
    of 40 000 pairs.  Results must still be the oracle's, bit for bit."""
    from tests import golden_fixtures as G
    if name not in G.NAMES:
        pytest.skip(f"fixture {name} has not been generated")
    monkeypatch.setenv("SDQ_STREAM_MIN_PAIRS", "1")
    monkeypatch.setenv("SDGPU_GRID_CHUNK_PAIRS", "40000")
    (bases, q, offsets, params, pars, iters), fx = G.load(name)
    with sd.GpuAligner(params) as al:
        got = sd.qluster(al, bases, q, offsets, pars, iters, want_clusters=False)
    G.assert_result_equals_fixture(got, fx)
    assert got["stats"]["streamedGrids"] >= 3


@pytest.mark.parametrize("mode", ["default", "noScreen", "noBand"])
@pytest.mark.parametrize("name", ["c2_5k", "c1_full", "c2_20k", "c2_20k_hp"])
def test_qluster_matches_frozen_oracle_at_baseline_sizes(name, mode):
    """BASELINE config 1 at full size and the config 2 generator at 5k / 20k reads (with and without
    homopolymer weighting) against the oracle's frozen results (tests/golden/make_qluster_fixtures.py):
    membership, consensus sequences / qualities / counts, chimera flags bit for bit -- through the screen
    pass, with it off (banded + full-matrix kernels) and with screen and band off (full-matrix only)."""
    from tests import golden_fixtures as G
    if name not in G.NAMES:
        pytest.skip(f"fixture {name} has not been generated")
    (bases, q, offsets, params, pars, iters), fx = G.load(name)
    with sd.GpuAligner(params) as al:
        if mode != "default":
            al.set_screen(1)
        if mode == "noBand":
            al.set_band(1)
        al.profile_enable(True)
        got = sd.qluster(al, bases, q, offsets, pars, iters, want_clusters=False)
        prof = al.profile_read()
    G.assert_result_equals_fixture(got, fx)
    assert (prof["screenCertified"] > 0) == (mode == "default")
    assert (prof["bandCells"] > 0) == (mode != "noBand")


def test_population_clustering_matches_oracle(oracle):
    """processClusters' population step (processClusters.cpp:226, sdq_population_cluster): six samples drawn from a
    shared haplotype pool are collapsed one by one (qluster), then their final clusters are collapsed across samples
    with an OTU table; input clusters keep their counts and are not deduplicated, k-mer checks are off.  Membership
    (sample cluster -> population cluster), population consensus sequences / qualities / counts must equal the oracle's."""
    rng = np.random.default_rng(77)
    pool = synth.make_haplotypes(rng, 9, 180, 1, 6)
    params = sd.make_params()
    iters = CollapseIterations.gen_illumina_default_pars(100)
    results = []
    with sd.GpuAligner(params) as al:
        for s in range(6):
            haps = [pool[i] for i in rng.choice(len(pool), 4, replace=False)]
            seqs, quals, _ = synth.fast_reads(rng, haps, 700, 0.004)
            results.append(sd.qluster(al, *synth.pack(seqs, quals), sd.QlusterPars(), iters, want_clusters=False))
        popIters = CollapseIterations.gen_illumina_default_pars(100)
        got = sd.population_cluster(al, results, sd.QlusterPars(), popIters)
        otu = sd.population_cluster(al, results, sd.QlusterPars(), CollapseIterations.gen_otu_pars(100, 0.99, True))
    bases, quals, offsets, cnts = got["input"]
    pars = sd.QlusterPars()
    want = oracle.population_cluster(params, pars.to_c(), popIters.iters, bases, quals, offsets, cnts)
    assert_same(got, want)
    wantOtu = oracle.population_cluster(params, pars.to_c(), CollapseIterations.gen_otu_pars(100, 0.99, True).iters, bases, quals, offsets, cnts)
    assert_same(otu, wantOtu)
    n = len(offsets) - 1
    assert len(got["sampleOfInput"]) == n == sum(len(r["packed"]["cnts"]) for r in results)
    assert len(got["clusters"]) < n  # the shared haplotypes were merged across samples ...
    assert sum(c for _, _, c in got["clusters"]) == cnts.sum()  # ... and no read was lost
    # every abundant pool haplotype that was drawn is a population cluster exactly once
    popSeqs = [s for s, _, _ in got["clusters"]]
    assert len(set(popSeqs)) == len(popSeqs)
    assert sum(h.tobytes().decode() in popSeqs for h in pool) >= 6


def test_config1_full_size_properties():
    """BASELINE config 1 at full size (10k reads x 250 bp, 20 haplotypes), without the oracle:
    size-independent properties of the collapse — conservation of reads, determinism, exact recovery
    of the abundant haplotypes, consensus/membership consistency."""
    haps, seqs, quals, which = synth.config1()
    bases, q, offsets = synth.pack(seqs, quals)
    with sd.GpuAligner(sd.make_params()) as al:
        a = sd.qluster(al, bases, q, offsets)
        b = sd.qluster(al, bases, q, offsets)
    assert np.array_equal(a["membership"], b["membership"]) and len(a["clusters"]) == len(b["clusters"])
    for (s1, q1, c1), (s2, q2, c2) in zip(a["clusters"], b["clusters"]):
        assert s1 == s2 and c1 == c2 and np.array_equal(q1, q2)
    n = len(seqs)
    mem = a["membership"]
    assert mem.max() < len(a["clusters"]) and (mem != np.uint32(0xFFFFFFFF)).all()
    counts = np.bincount(mem, minlength=len(a["clusters"]))
    assert counts.sum() == n
    assert [int(c) for _, _, c in a["clusters"]] == counts.tolist()          # cnt == members
    assert counts.tolist() == sorted(counts.tolist(), reverse=True)          # sorted by total count
    hs = {h.tobytes().decode(): k for k, h in enumerate(haps)}
    big = [(s, c) for s, _, c in a["clusters"] if c >= 0.01 * n]
    assert len(big) >= 15 and all(s in hs for s, _ in big)                   # abundant clusters are true haplotypes
    # reads assigned to an abundant cluster overwhelmingly come from that haplotype
    for ci, (s, _) in enumerate(big):
        src = which[mem == ci]
        assert (src == hs[s]).mean() > 0.90  # a 1-SNP neighbour's erroneous reads can land here legitimately
    assert a["stats"]["uniqueSeqs"] == len({x.tobytes() for x in seqs})


def test_mark_chimeras_matches_oracle(oracle):
    """collapser::markChimeras at the end of the run: the crossover of two abundant parents is
    flagged with the same parents / positions as the oracle, noisy reads included."""
    from tests import helpers as H
    rng = np.random.default_rng(21)
    bases, quals, offsets, (p0, p1, chi, var), cut = H.chimera_sample(rng, length=240, counts=(300, 220, 40, 25))
    # sequencing noise on top: substitutions at low quality
    bases = bases.copy()
    quals = quals.copy()
    hit = rng.random(len(bases)) < 0.003
    bases[hit] = H.BASES[rng.integers(0, 4, int(hit.sum()))]
    quals[hit] = rng.integers(2, 15, int(hit.sum()))
    its = CollapseIterations.gen_illumina_default_pars(100)
    for pars in (sd.QlusterPars(), sd.QlusterPars(chiKeepLowQaulityMismatches=True, chiPosSpacing=3),
                 sd.QlusterPars(markChimeras=False), sd.QlusterPars(parFreqs=1.0, chiRunCutOff=0)):
        got, want = run_pair(oracle, sd.make_params(), pars, its, bases, quals, offsets)
        assert_same(got, want)
        if pars.markChimeras:
            assert got["stats"]["chimeraPairs"] > 0
            flagged = [got["clusters"][c][0] for c in got["chimeras"]]
            assert bytes(chi).decode() in flagged
            assert got["stats"]["chimericClusters"] == len(got["chimeras"])
        else:
            assert got["chimeras"] == {} and got["stats"]["chimeraPairs"] == 0


def test_cxx_caller_matches_python_api(tmp_path):
    """examples/qluster_cabi.cpp (C++ straight onto the C ABI, its own parameter set-up) must print the same
    clusters the Python wrapper returns for the same packed sample."""
    import subprocess
    import __graft_entry__ as g
    from tests import helpers as H
    exe = g.build_example()
    rng = np.random.default_rng(33)
    bases, quals, offsets, (p0, p1, chi, var), cut = H.chimera_sample(rng, length=220, counts=(200, 150, 30, 20))
    bases = bases.copy()
    quals = quals.copy()
    hit = rng.random(len(bases)) < 0.004
    bases[hit] = H.BASES[rng.integers(0, 4, int(hit.sum()))]
    quals[hit] = rng.integers(2, 15, int(hit.sum()))
    path = tmp_path / "sample.bin"
    with open(path, "wb") as f:
        f.write(np.uint32(len(offsets) - 1).tobytes())
        f.write(np.ascontiguousarray(offsets, np.uint64).tobytes())
        f.write(bases.tobytes())
        f.write(quals.tobytes())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [line.split("\t") for line in r.stdout.strip().splitlines()]
    with sd.GpuAligner(sd.make_params()) as al:
        want = sd.qluster(al, bases, quals, offsets)
    assert len(got) == len(want["clusters"])
    for c, (cnt, chim, seq) in enumerate(got):
        assert float(cnt) == want["clusters"][c][2] and seq == want["clusters"][c][0]
        assert (chim == "1") == (c in want["chimeras"])
    assert any(chim == "1" for _, chim, _ in got)


def test_stop_check_small_cutoff_and_consensus_micro_cases(oracle):
    """The hand-derived oracle micro-cases (tests/test_oracle.py) through the GPU-driven loop: stopCheck windows of
    3 / 4 candidates, smallCutoff 9 / 10, the consensus quality rule and the minority insertion."""
    from seekdeep_b200.iterpars import IterPar
    from tests.test_oracle import _pack_reads
    rng = np.random.default_rng(17)
    haps = ["".join(rng.choice(list("ACGT"), 80)) for _ in range(4)]
    hi = [38] * 80
    reads = []
    for h, n in zip(haps, (40, 30, 20, 10)):
        reads += [(h, hi)] * n
    for p in (10, 40, 70):
        s = haps[3][:p] + ("A" if haps[3][p] != "A" else "C") + haps[3][p + 1:]
        reads.append((s, hi[:p] + [8] + hi[p + 1:]))
    bases, quals, offsets = _pack_reads(reads)
    pars = sd.QlusterPars(markChimeras=False, checkKmers=False)
    for stopCheck, smallCutoff, expect in ((100, 0, 4), (3, 0, 7), (4, 0, 4), (100, 10, 7), (100, 9, 4)):
        iters = CollapseIterations([IterPar(stopCheck=stopCheck, smallCutoff=smallCutoff, lqMismatches=1)])
        got, want = run_pair(oracle, sd.make_params(), pars, iters, bases, quals, offsets)
        assert_same(got, want)
        assert len(got["clusters"]) == expect, (stopCheck, smallCutoff)
    A = haps[0][:60]
    p = 30
    sub = A[:p] + ("C" if A[p] != "C" else "G") + A[p + 1:]
    q60 = [38] * 60
    bases, quals, offsets = _pack_reads([(A, q60)] * 3 + [(sub, q60[:p] + [10] + q60[p + 1:])] * 2)
    its = CollapseIterations.gen_illumina_default_pars(100)
    got, want = run_pair(oracle, sd.make_params(), sd.QlusterPars(markChimeras=False), its, bases, quals, offsets)
    assert_same(got, want)
    assert len(got["clusters"]) == 1 and got["clusters"][0][1][p] == 22 and got["clusters"][0][2] == 5.0


def test_qluster_from_fastq_file(oracle, tmp_path):
    """sdq_run_fastq (FASTQ ingest + collapse behind the C ABI) == the collapse of the same reads handed over packed."""
    import gzip
    from seekdeep_b200.qluster import qluster_fastq
    rng = np.random.default_rng(41)
    haps = synth.make_haplotypes(rng, 5, 130, 1, 5)
    seqs, quals, _ = synth.fast_reads(rng, haps, 900, 0.005)
    bases, q, offsets = synth.pack(seqs, quals)
    path = tmp_path / "sample.fastq.gz"
    with gzip.open(path, "wt") as f:
        for i, (s, ql) in enumerate(zip(seqs, quals)):
            f.write(f"@r{i}\n{s.tobytes().decode()}\n+\n{''.join(chr(33 + int(x)) for x in ql)}\n")
    params = sd.make_params()
    iters = CollapseIterations.gen_illumina_default_pars(100)
    with sd.GpuAligner(params) as al:
        got = qluster_fastq(al, path, sd.QlusterPars(), iters)
        packed = sd.qluster(al, bases, q, offsets, sd.QlusterPars(), iters)
    assert_same(got, packed)
    want = oracle.qluster(params, oracle_opts(sd.QlusterPars()), iters.iters, iters.iters, bases, q, offsets)
    assert_same(got, want)


def test_c2_full_size_paths_agree(monkeypatch):
    """BASELINE config 2 at its full size (200 000 reads x 300 bp: the benchmarked sample; hours for the oracle): the
    result must not depend on which exact path produced the verdicts or on how the host loop consumed them -- big grid
    streamed into the sequential pass (the default at this size) vs. consumed after the fact, screen pass on vs. off
    (certified band + full-matrix kernels instead).  The paths agree with the oracle at 20 000 reads
    (test_qluster_matches_frozen_oracle_at_baseline_sizes); here they have to agree with each other where the big
    iteration's grid has nine chunks and 178 799 live clusters."""
    haps, seqs, quals, _ = synth.config2_fast(nReads=200000)
    bases, q, offsets = synth.pack(seqs, quals)
    params = sd.make_params()
    iters = CollapseIterations.gen_illumina_default_pars(100)

    def run(screen=True):
        with sd.GpuAligner(params) as al:
            if not screen:
                al.set_screen(1)
            return sd.qluster(al, bases, q, offsets, sd.QlusterPars(), iters, want_clusters=False)

    streamed = run()
    assert streamed["stats"]["streamedGrids"] >= 1
    monkeypatch.setenv("SDQ_STREAM_MIN_PAIRS", str(1 << 62))
    plain = run()
    assert plain["stats"]["streamedGrids"] == 0
    noScreen = run(screen=False)
    for other in (plain, noScreen):
        assert np.array_equal(streamed["membership"], other["membership"])
        for k in ("bases", "quals", "offsets", "cnts"):
            assert np.array_equal(streamed["packed"][k], other["packed"][k]), k
        assert streamed["chimeras"] == other["chimeras"]
    assert streamed["stats"]["finalClusters"] == plain["stats"]["finalClusters"] > 1000
