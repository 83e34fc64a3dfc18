"""qluster --useKmerBinning / --kmerCutOff / --kCompareLen (clusterDownSetUp.cpp:352-354): the sorted k-mer lists and the
seed x read masks against the oracle's compareKmers, and the binned collapse end to end against the oracle's restatement."""
import numpy as np
import pytest

import seekdeep_b200 as sd
from tests import synth
from seekdeep_b200.aligner import GpuAligner, pack_seqs
from seekdeep_b200.iterpars import CollapseIterations
from seekdeep_b200.scoring import make_params
from tests.test_gpu_qluster import assert_same, oracle_opts

pytestmark = pytest.mark.gpu


def _store(rng):
    """Ragged lengths, families of near-identical sequences, low-complexity sequences (repeated k-mers), sequences
    shorter than k, N and an IUPAC code, one sequence longer than what a lane keeps in registers."""
    seqs = []
    for fam in range(6):
        root = synth.random_seq(rng, int(rng.integers(120, 320)))
        seqs.append(root)
        for _ in range(7):
            seqs.append(synth.mutate_snps(rng, root, int(rng.integers(1, 12))))
    seqs.append(np.frombuffer(b"ACACACACACACACACACACACACACACACACACACACACACAC", np.uint8))
    seqs.append(np.frombuffer(b"ACACACACACACACACACACACACACTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", np.uint8))
    seqs.append(np.frombuffer(b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", np.uint8))
    seqs.append(np.frombuffer(b"AAAAAAAAAAAAAAAAAAAAAAAAAAA", np.uint8))
    seqs.append(np.frombuffer(b"ACGTA", np.uint8))
    seqs.append(np.frombuffer(b"ACGTACGTNACGTRACGTTGCANNACGT", np.uint8))
    long = synth.random_seq(rng, 700)
    seqs.append(long)
    seqs.append(synth.mutate_snps(rng, long, 9))
    seqs.append(np.concatenate([long[:300], long[:300]]))  # every k-mer of its first half twice
    return seqs


def test_kmer_bin_masks_match_oracle(oracle):
    rng = np.random.default_rng(77)
    seqs = _store(rng)
    n = len(seqs)
    b, q, off = pack_seqs([s.tobytes() for s in seqs], [np.zeros(len(s), np.uint8) for s in seqs])
    order = rng.permutation(n).astype(np.uint32)
    seeds = order[:min(64, n)]
    reads = np.arange(n, dtype=np.uint32)
    with GpuAligner(make_params()) as al:
        al.upload_seqs(b, q, off)
        with pytest.raises(sd.SdgpuError):
            al.kmer_bin_masks(seeds, reads, 0.5)  # no lists yet
        for k in (1, 4, 10, 16):
            al.build_kmer_lists(k)
            want_sh = np.zeros((n, len(seeds)), np.uint32)
            want_fr = np.zeros((n, len(seeds)), np.float64)
            for i, r in enumerate(reads):
                for j, c in enumerate(seeds):
                    want_sh[i, j], want_fr[i, j] = oracle.kmer_compare(seqs[c].tobytes(), seqs[r].tobytes(), k)
            # cut-offs on and next to values the fractions take: the comparison is a strict >
            cuts = [0.0, 0.5, 0.8, 1.0] + [float(x) for x in rng.choice(np.unique(want_fr), 6)]
            for cut in cuts:
                masks, shared = al.kmer_bin_masks(seeds, reads, cut, want_shared=True)
                assert np.array_equal(shared, want_sh)
                want_m = np.zeros(n, np.uint64)
                for j in range(len(seeds)):
                    want_m |= (want_fr[:, j] > cut).astype(np.uint64) << np.uint64(j)
                assert np.array_equal(masks, want_m)
            # symmetric in its arguments, like compareKmers
            m2, sh2 = al.kmer_bin_masks(reads[:64], seeds, 0.3, want_shared=True)
            assert np.array_equal(sh2, want_sh[:64].T)
        # few seeds / one read / no seeds
        al.build_kmer_lists(10)
        m = al.kmer_bin_masks(seeds[:3], reads[:1], 0.2)
        assert m.shape == (1,)
        assert np.array_equal(al.kmer_bin_masks(seeds[:0], reads, 0.2), np.zeros(n, np.uint64))
        with pytest.raises(sd.SdgpuError):
            al.kmer_bin_masks(np.arange(65, dtype=np.uint32) % n, reads, 0.2)
        with pytest.raises(sd.SdgpuError):
            al.build_kmer_lists(17)
        # a sequence appended after the lists were built is not covered
        first = al.append_seqs(*pack_seqs([seqs[0].tobytes()], [np.zeros(len(seqs[0]), np.uint8)]))
        with pytest.raises(sd.SdgpuError):
            al.kmer_bin_masks(seeds[:2], np.array([first], np.uint32), 0.2)
        al.build_kmer_lists(10)
        m = al.kmer_bin_masks(np.array([0], np.uint32), np.array([first], np.uint32), 0.99)
        assert m[0] == 1  # identical to sequence 0: fraction 1.0


def _multi_family_sample(rng, nFam, nReads, length, err=0.004):
    """Reads from several unrelated haplotype families (k-mer bins), each family a root + close variants."""
    seqs, quals = [], []
    for f in range(nFam):
        haps = synth.make_haplotypes(rng, 4, length + 7 * (f % 6), 1, 4)
        s, q, _ = synth.fast_reads(rng, haps, nReads // nFam, err)
        seqs += s
        quals += q
    perm = rng.permutation(len(seqs))
    return synth.pack([seqs[i] for i in perm], [quals[i] for i in perm])


@pytest.mark.parametrize("seed,nFam,k,cut", [(5, 3, 10, 0.80), (6, 5, 6, 0.55), (7, 2, 12, 0.95), (8, 70, 10, 0.80)])
def test_qluster_kmer_binning_matches_oracle(oracle, seed, nFam, k, cut):
    rng = np.random.default_rng(seed)
    # (70 families: more bins than one round's 64 candidate seeds; kept small, the oracle aligns every read against up
    #  to 100 unrelated clusters once the bins are merged back)
    bases, q, offsets = _multi_family_sample(rng, nFam, 1400 if nFam < 10 else 560, 110 if nFam < 10 else 90)
    params = sd.make_params()
    iters = CollapseIterations.gen_illumina_default_pars(100)
    pars = sd.QlusterPars(useKmerBinning=True, kCompareLen=k, kmerCutOff=cut)
    with sd.GpuAligner(params) as al:
        got = sd.qluster(al, bases, q, offsets, pars, iters)
        plain = sd.qluster(al, bases, q, offsets, sd.QlusterPars(), iters)
    want = oracle.qluster(params, oracle_opts(pars), iters.iters, iters.iters, bases, q, offsets)
    assert_same(got, want)
    assert got["stats"]["kmerBins"] >= nFam and got["stats"]["kmerPairs"] > 0
    assert plain["stats"]["kmerBins"] == 0
    # every read ends somewhere, and binning happened before any alignment
    assert sum(c for _, _, c in got["clusters"]) == sum(c for _, _, c in plain["clusters"])


def test_qluster_kmer_binning_with_bin_pars(oracle):
    """pars.binIteratorMap (--binPar) differs from the run's iterations: its lines join the verdict table."""
    rng = np.random.default_rng(21)
    bases, q, offsets = _multi_family_sample(rng, 4, 1600, 130, err=0.006)
    params = sd.make_params()
    iters = CollapseIterations.gen_illumina_default_pars(100)
    binIters = CollapseIterations.gen_otu_pars(100, 0.99, True)
    for extra in ({}, {"converge": True}, {"startWithSingles": True}):
        pars = sd.QlusterPars(useKmerBinning=True, kCompareLen=8, kmerCutOff=0.7, **extra)
        with sd.GpuAligner(params) as al:
            got = sd.qluster(al, bases, q, offsets, pars, iters, binIteratorMap=binIters)
        want = oracle.qluster(params, pars.to_c(binIters), iters.iters, iters.iters, bases, q, offsets)
        assert_same(got, want)
    with sd.GpuAligner(params) as al:
        with pytest.raises(sd.SdgpuError):
            sd.qluster(al, bases, q, offsets, sd.QlusterPars(useKmerBinning=True, kCompareLen=17), iters)
