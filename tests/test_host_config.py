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
