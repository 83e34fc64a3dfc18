"""Shared helpers for the parity tests: seeded random pair generators and independent checks."""
import numpy as np

from seekdeep_b200 import _abi
from seekdeep_b200.scoring import GapScoringParameters, SubstituteMatrix, make_params

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def rand_seq(rng, n, alphabet=BASES):
    return alphabet[rng.integers(0, len(alphabet), n)]


def mutate(rng, s, nSub=0, nIns=0, nDel=0, hp=False):
    s = s.copy()
    for _ in range(nSub):
        p = int(rng.integers(len(s)))
        s[p] = rng.choice(BASES[BASES != s[p]]) if s[p] in BASES else BASES[0]
    for _ in range(nDel):
        if len(s) > 2:
            p = int(rng.integers(len(s)))
            ln = int(rng.integers(1, 4))
            s = np.delete(s, slice(p, min(len(s) - 1, p + ln)))
    for _ in range(nIns):
        p = int(rng.integers(len(s) + 1))
        ln = int(rng.integers(1, 4))
        ins = np.full(ln, s[min(p, len(s) - 1)], np.uint8) if hp else rand_seq(rng, ln)
        s = np.insert(s, p, ins)
    return s


def rand_quals(rng, n, lowFrac=0.15):
    q = rng.integers(26, 41, n).astype(np.uint8)
    low = rng.random(n) < lowFrac
    q[low] = rng.integers(2, 26, int(low.sum())).astype(np.uint8)
    return q


def make_store(rng, nSeqs, length, jitter=0, family=True, lowFrac=0.15, alphabet=BASES):
    """A store of related sequences (one root, mutated copies) + qualities + counts."""
    root = rand_seq(rng, length, alphabet)
    seqs, quals = [], []
    for k in range(nSeqs):
        if family and k > 0:
            s = mutate(rng, root, int(rng.integers(0, 6)), int(rng.integers(0, 2)), int(rng.integers(0, 2)),
                       hp=bool(rng.integers(2)))
        else:
            s = rand_seq(rng, length + int(rng.integers(-jitter, jitter + 1)) if jitter else length, alphabet) if k else root
        seqs.append(s)
        quals.append(rand_quals(rng, len(s), lowFrac))
    offsets = np.zeros(nSeqs + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    cnts = rng.integers(1, 50, nSeqs).astype(np.float64)
    return np.concatenate(seqs), np.concatenate(quals), offsets, cnts, seqs, quals


def assert_comparisons_equal(got, want, rel=1e-6):
    """Integer fields bit-exact; float fields within `rel` (north_star: 1e-6 relative)."""
    assert got.shape == want.shape
    for name in _abi.INT_FIELDS:
        bad = np.nonzero(got[name] != want[name])[0]
        assert bad.size == 0, f"{name}: first mismatch at pair {bad[0]}: got {got[name][bad[0]]} want {want[name][bad[0]]}"
    for name in _abi.FLOAT_FIELDS:
        g, w = got[name], want[name]
        tol = rel * np.maximum(np.abs(w), 1e-300)
        bad = np.nonzero(np.abs(g - w) > tol)[0]
        assert bad.size == 0, f"{name}: first mismatch at pair {bad[0]}: got {g[bad[0]]!r} want {w[bad[0]]!r}"


def gotoh_score(a, b, gp, mat):
    """Independent optimal-score reference: textbook 3-state affine global DP with -inf
    boundaries and the ten end-gap penalties (pure Python; small inputs only)."""
    NEG = -10**9
    la, lb = len(a), len(b)
    U = [[NEG] * (lb + 1) for _ in range(la + 1)]
    L = [[NEG] * (lb + 1) for _ in range(la + 1)]
    D = [[NEG] * (lb + 1) for _ in range(la + 1)]
    D[0][0] = 0
    for i in range(1, la + 1):
        U[i][0] = -gp.gapLeftQueryOpen - (i - 1) * gp.gapLeftQueryExtend
    for j in range(1, lb + 1):
        L[0][j] = -gp.gapLeftRefOpen - (j - 1) * gp.gapLeftRefExtend
    for i in range(1, la + 1):
        for j in range(1, lb + 1):
            go, ge = (gp.gapRightQueryOpen, gp.gapRightQueryExtend) if j == lb else (gp.gapOpen, gp.gapExtend)
            U[i][j] = max(U[i - 1][j] - ge, L[i - 1][j] - go, D[i - 1][j] - go)
            go, ge = (gp.gapRightRefOpen, gp.gapRightRefExtend) if i == la else (gp.gapOpen, gp.gapExtend)
            L[i][j] = max(U[i][j - 1] - go, L[i][j - 1] - ge, D[i][j - 1] - go)
            D[i][j] = mat[a[i - 1]][b[j - 1]] + max(U[i - 1][j - 1], L[i - 1][j - 1], D[i - 1][j - 1])
    return max(U[la][lb], L[la][lb], D[la][lb])


def score_of_alignment(alnA, alnB, gp, mat):
    """Score of a concrete gapped alignment under the ten penalties (checks the traceback)."""
    n = len(alnA)
    la = sum(c != "-" for c in alnA)
    lb = sum(c != "-" for c in alnB)
    score, i, j, k = 0, 0, 0, 0
    while k < n:
        if alnA[k] == "-" or alnB[k] == "-":
            inA = alnA[k] == "-"
            s = alnA if inA else alnB
            e = k
            while e < n and s[e] == "-" and (alnB if inA else alnA)[e] != "-":
                e += 1
            size = e - k
            if inA:  # gap in ref: horizontal moves in row i
                if i == 0:
                    go, ge = gp.gapLeftRefOpen, gp.gapLeftRefExtend
                elif i == la:
                    go, ge = gp.gapRightRefOpen, gp.gapRightRefExtend
                else:
                    go, ge = gp.gapOpen, gp.gapExtend
                j += size
            else:
                if j == 0:
                    go, ge = gp.gapLeftQueryOpen, gp.gapLeftQueryExtend
                elif j == lb:
                    go, ge = gp.gapRightQueryOpen, gp.gapRightQueryExtend
                else:
                    go, ge = gp.gapOpen, gp.gapExtend
                i += size
            score -= go + (size - 1) * ge
            k = e
        else:
            score += mat[ord(alnA[k])][ord(alnB[k])]
            i += 1
            j += 1
            k += 1
    return score


PARAM_SETS = {
    "qluster": dict(gaps=GapScoringParameters.qluster_default()),
    "allEndsCost": dict(gaps=GapScoringParameters.from_strings("5,1", "5,1", "5,1")),
    "allEndsFree": dict(gaps=GapScoringParameters.from_strings("5,1", "0,0", "0,0")),
    "stitch": dict(gaps=GapScoringParameters.from_strings("10,1", "0,0", "0,0"), qualThresWindow=0),
    "asym": dict(gaps=GapScoringParameters(7, 2, 3, 1, 0, 0, 4, 2, 6, 1)),
    "degenHomopolymer": dict(gaps=GapScoringParameters.qluster_default(),
                             scoring=SubstituteMatrix.create_degen_score_matrix_case_insensitive(2, -2),
                             weighHomopolymers=True),
    "countEnds": dict(gaps=GapScoringParameters.qluster_default(), countEndGaps=True, weighHomopolymers=True),
}


def params_for(name):
    return make_params(**PARAM_SETS[name])


def chimera_sample(rng, length=200, nDiff=8, counts=(60, 45, 8, 6), qual=38):
    """Reads of two parents, their crossover (front of parent 0, back of parent 1) and a one-SNP
    variant of parent 0, all error-free at high quality: (bases, quals, offsets, seqs)."""
    p0 = rand_seq(rng, length)
    p1 = p0.copy()
    pos = np.sort(rng.choice(np.arange(10, length - 10), nDiff, replace=False))
    for x in pos:
        p1[x] = BASES[(list(BASES).index(p1[x]) + 1 + int(rng.integers(3))) % 4]
    cut = int((pos[nDiff // 2 - 1] + pos[nDiff // 2]) // 2)
    chi = np.concatenate([p0[:cut], p1[cut:]])
    var = p0.copy()
    x = int(pos[-1]) + 3
    var[x] = BASES[(list(BASES).index(var[x]) + 1) % 4]
    seqs = []
    for s, c in zip((p0, p1, chi, var), counts):
        seqs += [s] * c
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    offsets = np.zeros(len(seqs) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.concatenate(seqs)
    return bases, np.full(len(bases), qual, np.uint8), offsets, (p0, p1, chi, var), cut


def masks_from_comparisons(comps, iters):
    """comparison::passErrorProfile of every par line applied on the host to errorProfiles
    (COMPARISON_DTYPE records): the uint64 verdict masks sdgpu_align_pass* must return."""
    masks = np.zeros(len(comps), np.uint64)
    for l, it in enumerate(iters):
        if it.onPerId:
            ok = (comps["eventBasedIdentityHq"] if it.onHqPerId else comps["eventBasedIdentity"]) >= it.perId
        else:
            ok = ((it.oneBaseIndel >= comps["oneBaseIndel"]) & (it.twoBaseIndel >= comps["twoBaseIndel"]) &
                  (it.largeBaseIndel >= comps["largeBaseIndel"]) & (it.hqMismatches >= comps["hqMismatches"]) &
                  (it.lqMismatches >= comps["lqMismatches"]) & (it.lowKmerMismatches >= comps["lowKmerMismatches"]))
        masks |= ok.astype(np.uint64) << np.uint64(l)
    return masks
