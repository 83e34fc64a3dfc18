"""Seeded synthetic amplicon reads for the BASELINE.json configs (C1..C5): test and bench data, not product code.
No datasets ship with the reference and there is no network, so every workload is generated."""
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_seq(rng, n):
    return BASES[rng.integers(0, 4, n)]


def mutate_snps(rng, seq, nSnps):
    s = seq.copy()
    pos = rng.choice(len(s), size=nSnps, replace=False)
    for p in pos:
        s[p] = rng.choice(BASES[BASES != s[p]])
    return s


def make_haplotypes(rng, nHaps, length, minSnps=1, maxSnps=6):
    root = random_seq(rng, length)
    haps = [root]
    while len(haps) < nHaps:
        h = mutate_snps(rng, root, int(rng.integers(minSnps, maxSnps + 1)))
        if not any(np.array_equal(h, x) for x in haps):
            haps.append(h)
    return haps


def sample_reads(rng, haps, nReads, subRate, zipf=1.2, hpIndelRate=0.0, qMean=36.0, qSd=3.0, errQMean=12.0,
                 errQSd=6.0):
    """Reads = haplotype + per-base substitutions (+ optional homopolymer +-1 indels for runs
    >= 3 with P = hpIndelRate * runLen).  Erroneous bases get low qualities."""
    w = 1.0 / np.arange(1, len(haps) + 1) ** zipf
    w /= w.sum()
    which = rng.choice(len(haps), size=nReads, p=w)
    seqs, quals = [], []
    for r in range(nReads):
        s = haps[which[r]].copy()
        q = np.clip(np.rint(rng.normal(qMean, qSd, len(s))), 2, 40).astype(np.uint8)
        err = rng.random(len(s)) < subRate
        ne = int(err.sum())
        if ne:
            idx = np.nonzero(err)[0]
            for p in idx:
                s[p] = rng.choice(BASES[BASES != s[p]])
            q[idx] = np.clip(np.rint(rng.normal(errQMean, errQSd, ne)), 2, 40).astype(np.uint8)
        if hpIndelRate > 0:
            out_s, out_q = [], []
            i = 0
            n = len(s)
            while i < n:
                j = i
                while j < n and s[j] == s[i]:
                    j += 1
                run = j - i
                seg_s, seg_q = s[i:j], q[i:j]
                if run >= 3 and rng.random() < hpIndelRate * run:
                    if rng.random() < 0.5:
                        seg_s, seg_q = seg_s[:-1], seg_q[:-1]
                    else:
                        seg_s = np.append(seg_s, seg_s[0])
                        seg_q = np.append(seg_q, np.uint8(max(2, int(rng.normal(errQMean, errQSd)))))
                out_s.append(seg_s)
                out_q.append(seg_q)
                i = j
            s = np.concatenate(out_s)
            q = np.concatenate(out_q).astype(np.uint8)
        seqs.append(s)
        quals.append(q)
    return seqs, quals, which


def pack(seqs, quals):
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    return np.concatenate(seqs).astype(np.uint8), np.concatenate(quals).astype(np.uint8), offsets


def config1(seed=1, nReads=10_000, length=250, nHaps=20):
    """C1: one MiSeq sample, 10k reads of a 250-bp amplicon from 20 haplotypes, 0.5 %/base subs."""
    rng = np.random.default_rng(seed)
    haps = make_haplotypes(rng, nHaps, length, 1, 6)
    seqs, quals, which = sample_reads(rng, haps, nReads, 0.005)
    return haps, seqs, quals, which


def config2(seed=2, nReads=200_000, length=300, nHaps=200):
    """C2: 200k reads x 300 bp from 200 haplotypes, 1 % subs + homopolymer indels."""
    rng = np.random.default_rng(seed)
    haps = make_haplotypes(rng, nHaps, length, 1, 8)
    seqs, quals, which = sample_reads(rng, haps, nReads, 0.01, hpIndelRate=0.002)
    return haps, seqs, quals, which


def config3(seed=3, n=20_000, length=400, nRoots=50):
    """C3: unique 400-mers = roots x derived by 0-12 SNPs + 0-2 indels (all-vs-all input)."""
    rng = np.random.default_rng(seed)
    roots = [random_seq(rng, length) for _ in range(nRoots)]
    out, seen = [], set()
    while len(out) < n:
        s = mutate_snps(rng, roots[int(rng.integers(nRoots))], int(rng.integers(0, 13)))
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(1, len(s) - 1))
            if rng.random() < 0.5:
                s = np.delete(s, p)
            else:
                s = np.insert(s, p, BASES[rng.integers(4)])
        key = s.tobytes()
        if key not in seen:
            seen.add(key)
            out.append(s)
    quals = [np.full(len(s), 35, np.uint8) for s in out]
    return out, quals


def fast_reads(rng, haps, nReads, subRate, zipf=1.2, hpIndelRate=0.0, qMean=36.0, qSd=3.0, errQMean=12.0, errQSd=6.0):
    """Vectorised sample_reads (same error model) for the 200k-read configs."""
    w = 1.0 / np.arange(1, len(haps) + 1) ** zipf
    w /= w.sum()
    which = rng.choice(len(haps), size=nReads, p=w)
    seqs = [None] * nReads
    quals = [None] * nReads
    for h, hap in enumerate(haps):
        idx = np.nonzero(which == h)[0]
        if idx.size == 0:
            continue
        L = len(hap)
        S = np.tile(hap, (idx.size, 1))
        Q = np.clip(np.rint(rng.normal(qMean, qSd, S.shape)), 2, 40).astype(np.uint8)
        err = rng.random(S.shape) < subRate
        ne = int(err.sum())
        if ne:
            shift = rng.integers(1, 4, ne)
            code = np.searchsorted(BASES, S[err])  # BASES is sorted: A C G T
            S[err] = BASES[(code + shift) % 4]
            Q[err] = np.clip(np.rint(rng.normal(errQMean, errQSd, ne)), 2, 40).astype(np.uint8)
        runs = []
        if hpIndelRate > 0:
            i = 0
            while i < L:
                j = i
                while j < L and hap[j] == hap[i]:
                    j += 1
                if j - i >= 3:
                    runs.append((i, j - i))
                i = j
        ev = None
        if runs:
            p = np.array([hpIndelRate * ln for _, ln in runs])
            ev = rng.random((idx.size, len(runs))) < p
        for k, r in enumerate(idx):
            s, q = S[k], Q[k]
            if ev is not None and ev[k].any():
                for ri in np.nonzero(ev[k])[0][::-1]:  # right to left keeps earlier offsets valid
                    st, ln = runs[ri]
                    if rng.random() < 0.5:
                        s = np.delete(s, st + ln - 1)
                        q = np.delete(q, st + ln - 1)
                    else:
                        s = np.insert(s, st + ln, hap[st])
                        q = np.insert(q, st + ln, np.uint8(max(2, min(40, int(rng.normal(errQMean, errQSd))))))
            seqs[r] = s
            quals[r] = q
    return seqs, quals, which


def config2_fast(seed=2, nReads=200_000, length=300, nHaps=200):
    rng = np.random.default_rng(seed)
    haps = make_haplotypes(rng, nHaps, length, 1, 8)
    seqs, quals, which = fast_reads(rng, haps, nReads, 0.01, hpIndelRate=0.002)
    return haps, seqs, quals, which
