"""CPU tests of the host-side mirrors: par-file parsing, iteration tables, score matrices."""
import os

import numpy as np
import pytest

from seekdeep_b200.iterpars import CollapseIterations
from seekdeep_b200.scoring import GapScoringParameters, SubstituteMatrix, make_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_illumina_default_equals_par_file():
    ref = CollapseIterations.from_file(os.path.join(GOLDEN, "illumina_lkmer1.par"))
    gen = CollapseIterations.gen_illumina_default_pars(100)
    assert ref.iters == gen.iters and len(gen) == 22


def test_fractional_par_values():
    its = CollapseIterations.from_file(os.path.join(GOLDEN, "454_it_lkmer2_largeHpr.par"))
    assert len(its) == 22 and its.iters[3].largeBaseIndel == 0.99 and its.iters[10].lowKmerMismatches == 2


def test_otu_pars():
    its = CollapseIterations.from_file(os.path.join(GOLDEN, "otu_99.par"), onPerId=True, onHqPerId=True)
    assert its.iters == CollapseIterations.gen_otu_pars(100, 0.99, True).iters


def test_round_trip_write():
    gen = CollapseIterations.gen_illumina_default_pars(100)
    assert CollapseIterations.parse(gen.write_pars()).iters == gen.iters


def test_bad_par_line():
    with pytest.raises(ValueError):
        CollapseIterations.parse("100:3:1:0\n")


def test_gap_strings():
    g = GapScoringParameters.qluster_default()
    assert (g.gapOpen, g.gapExtend, g.gapLeftRefOpen, g.gapRightQueryOpen, g.gapRightRefExtend) == (5, 1, 5, 0, 0)


def test_matrices():
    m = SubstituteMatrix.create_score_matrix(2, -2).mat
    assert m[ord("A"), ord("A")] == 2 and m[ord("A"), ord("a")] == -2 and m[ord("N"), ord("A")] == -2
    d = SubstituteMatrix.create_degen_score_matrix_case_insensitive(2, -2).mat
    assert d[ord("A"), ord("a")] == 2 and d[ord("N"), ord("A")] == 2 and d[ord("R"), ord("G")] == 2
    assert d[ord("R"), ord("C")] == -2 and d[ord("Y"), ord("R")] == -2 and d[ord("B"), ord("D")] == 2
    assert np.array_equal(d, d.T)
    p = make_params(scoring=SubstituteMatrix(d))
    assert p.scoreMat[ord("R") * 128 + ord("G")] == 2


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours) on a shrunken sample: one JSON
    line with the contract keys, `impl`, a `cpu_baseline` describing the run and a zero-copy `e2e`."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SD_BENCH_CPU_READS="40")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "qluster_reads_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit 0 without work
    r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=600)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_host_base_translation_matches_table_lookup():
    """sd_translate_bases (the upload's inner loop: 32 bases per step through two byte shuffles when a block holds only
    A, C, G, T, N, the code table byte by byte otherwise) against a plain table look-up: every length around the block
    size, blocks with degenerate / lower-case / unknown characters at every position."""
    import ctypes as C
    import numpy as np
    from seekdeep_b200 import _abi
    lib = _abi.load()
    fn = lib.sd_translate_bases
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_bool)]
    codeOf = np.full(256, -1, np.int16)
    for i, ch in enumerate(b"ACGTNRYa"):
        codeOf[ch] = i
    rng = np.random.default_rng(8)
    for trial in range(300):
        n = int(rng.integers(0, 200))
        src = np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, n)].copy()
        kind = trial % 4
        if n and kind:  # sprinkle other characters: coded ones (R, Y, a) and, every fourth trial, one without a code
            for p in rng.integers(0, n, int(rng.integers(1, 4))):
                src[p] = rng.choice(np.frombuffer(b"RYa", np.uint8))
            if kind == 3:
                src[int(rng.integers(0, n))] = ord("x")
        dst = np.zeros(n + 1, np.uint8)
        dst[n] = 0xAB  # guard
        mc, unk = C.c_int(-1), C.c_bool(False)
        fn(src.ctypes.data, dst.ctypes.data, n, codeOf.ctypes.data, C.byref(mc), C.byref(unk))
        want = codeOf[src]
        assert np.array_equal(dst[:n], want.astype(np.uint8)) and dst[n] == 0xAB
        assert mc.value == (int(want.max()) if n else -1)
        assert unk.value == bool((want < 0).any())
