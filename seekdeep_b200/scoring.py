"""Host-side mirrors of njhseq's gapScoringParameters and substituteMatrix as the reference
uses them (clusterDown.cpp:217-218, clusterDownSetUp.cpp:402-405, extractorSetUp.cpp:138-150,
extractorPairedEnd.cpp:338-346).  They only fill the POD `sdgpu_params`."""
from dataclasses import dataclass

import numpy as np

from . import _abi

IUPAC = {
    "A": "A", "C": "C", "G": "G", "T": "T", "U": "T", "N": "ACGT", "R": "AG", "Y": "CT", "S": "CG",
    "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG",
}


@dataclass
class GapScoringParameters:
    """Ten penalties; the 2/4/6-argument constructors of the reference set query and ref sides alike."""
    gapOpen: int = 5
    gapExtend: int = 1
    gapLeftQueryOpen: int = 5
    gapLeftQueryExtend: int = 1
    gapRightQueryOpen: int = 5
    gapRightQueryExtend: int = 1
    gapLeftRefOpen: int = 5
    gapLeftRefExtend: int = 1
    gapRightRefOpen: int = 5
    gapRightRefExtend: int = 1

    @classmethod
    def from_strings(cls, gap="5,1", gapLeft="5,1", gapRight="0,0"):
        """`pars_.gap_ / gapLeft_ / gapRight_` strings (clusterDownSetUp.cpp:402-404)."""
        go, ge = (int(x) for x in gap.split(","))
        lo, le = (int(x) for x in gapLeft.split(","))
        ro, re_ = (int(x) for x in gapRight.split(","))
        return cls(go, ge, lo, le, ro, re_, lo, le, ro, re_)

    @classmethod
    def qluster_default(cls):
        return cls.from_strings("5,1", "5,1", "0,0")

    def to_c(self):
        g = _abi.GapPars()
        for name, _ in _abi.GapPars._fields_:
            setattr(g, name, int(getattr(self, name)))
        return g


class SubstituteMatrix:
    """128x128 int32 matrix indexed [refChar][readChar] like njhseq's `mat_`."""

    def __init__(self, mat):
        self.mat = np.ascontiguousarray(mat, dtype=np.int32).reshape(_abi.MAT_DIM, _abi.MAT_DIM)

    @classmethod
    def create_score_matrix(cls, match=2, mismatch=-2):
        """substituteMatrix(match, mismatch): identical characters match (case-sensitive)."""
        m = np.full((128, 128), mismatch, dtype=np.int32)
        np.fill_diagonal(m, match)
        return cls(m)

    @classmethod
    def create_degen_score_matrix_case_insensitive(cls, match=2, mismatch=-2):
        """createDegenScoreMatrixCaseInsensitive (clusterDown.cpp:218): IUPAC codes match when
        their base sets intersect; upper and lower case are interchangeable."""
        m = np.full((128, 128), mismatch, dtype=np.int32)
        np.fill_diagonal(m, match)
        for a, sa in IUPAC.items():
            for b, sb in IUPAC.items():
                v = match if set(sa) & set(sb) else mismatch
                for x in (a, a.lower()):
                    for y in (b, b.lower()):
                        m[ord(x), ord(y)] = v
        return cls(m)


def make_params(gaps=None, scoring=None, primaryQual=25, secondaryQual=20, qualThresWindow=2,
                countEndGaps=False, weighHomopolymers=False):
    """Everything the 7-argument aligner constructor receives (clusterDown.cpp:420-426).
    Defaults are qluster --illumina (clusterDownSetUp.cpp:254-288, 402-405)."""
    gaps = gaps or GapScoringParameters.qluster_default()
    scoring = scoring or SubstituteMatrix.create_score_matrix(2, -2)
    p = _abi.Params()
    p.gaps = gaps.to_c()
    flat = scoring.mat.reshape(-1)
    C = _abi.C
    C.memmove(p.scoreMat, flat.ctypes.data, flat.nbytes)
    p.primaryQual = primaryQual
    p.secondaryQual = secondaryQual
    p.qualThresWindow = qualThresWindow
    p.countEndGaps = int(countEndGaps)
    p.weighHomopolymers = int(weighHomopolymers)
    return p
