"""Host-side mirror of njhseq::PairedReadProcessor over the C ABI (include/sdgpu_stitch.h).

Reference: src/SeekDeep/objects/IlluminaUtils/PairedReadProcessor.{hpp,cpp} (ProcessParams :104-133,
processPairedEnd :287-664) as driven by extractorPairedEnd.cpp:338-348, 420-444 — the aligner it is given has
gap 10,1 with every end gap free, match 2 / mismatch -2 and qualThresWindow_ 0.  Mates are passed
reverse-complemented, like PairedRead::mateSeqBase_ with mateRComplemented_."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .aligner import GpuAligner, SdgpuError, _ptr
from .scoring import GapScoringParameters, SubstituteMatrix, make_params

OVERLAP_DTYPE = np.dtype([("alnScore", "<i4"), ("shift", "<i4"), ("basesInAln", "<u4"), ("hqMismatches", "<u4"),
                          ("lqMismatches", "<u4"), ("lowKmerMismatches", "<u4"), ("highQualityMatches", "<u4"),
                          ("lowQualityMatches", "<u4"), ("eventBasedIdentity", "<f8"), ("eventBasedIdentityHq", "<f8")])
assert OVERLAP_DTYPE.itemsize == 48
NOOVERLAP, R1BEGINSINR2, R1ENDSINR2, PERFECTOVERLAP, R1ALLINR2, R2ALLINR1 = range(6)


class ProcessParamsC(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("minOverlap", "hardMismatchCutOff", "lqMismatchCutOff", "hqMismatchCutOff", "r1Trim",
                                          "r2Trim", "trimLowQualWindows", "windowSize", "windowStep", "reserved")] + \
               [(n, C.c_double) for n in ("errorAllowed", "avgQualCutOff", "percentAfterTrimCutOff")]


@dataclass
class ProcessParams:
    """PairedReadProcessor::ProcessParams with its defaults (PairedReadProcessor.hpp:104-133)."""
    minOverlap: int = 10
    errorAllowed: float = 0.01
    hardMismatchCutOff: int = 20
    lqMismatchCutOff: int = 10
    hqMismatchCutOff: int = 10
    r1Trim: int = 0
    r2Trim: int = 0
    trimLowQaulWindows: bool = True
    windowSize: int = 5
    windowStep: int = 1
    avgQualCutOff: float = 25.0
    percentAfterTrimCutOff: float = 0.70

    def to_c(self):
        return ProcessParamsC(self.minOverlap, self.hardMismatchCutOff, self.lqMismatchCutOff, self.hqMismatchCutOff, self.r1Trim,
                              self.r2Trim, int(self.trimLowQaulWindows), self.windowSize, self.windowStep, 0, self.errorAllowed,
                              self.avgQualCutOff, self.percentAfterTrimCutOff)


def pair_aligner_params(primaryQual=20, secondaryQual=15):
    """The aligner extractorPairedEnd builds for stitching (extractorPairedEnd.cpp:338-348)."""
    return make_params(gaps=GapScoringParameters.from_strings("10,1", "0,0", "0,0"), scoring=SubstituteMatrix.create_score_matrix(2, -2),
                       primaryQual=primaryQual, secondaryQual=secondaryQual, qualThresWindow=0)


_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    v = C.c_void_p
    lib.sdgpu_align_no_internal_gaps.restype = C.c_int
    lib.sdgpu_align_no_internal_gaps.argtypes = [v, v, v, C.c_uint32, C.c_uint32, v]
    lib.sdst_process_pairs.restype = C.c_int
    lib.sdst_process_pairs.argtypes = [v, C.POINTER(ProcessParamsC), v, v, v, v, v, v, C.c_uint32, C.POINTER(v)]
    lib.sdst_free.argtypes = [v]
    lib.sdst_last_error.restype = C.c_char_p
    lib.sdst_total_len.restype = C.c_uint64
    lib.sdst_total_len.argtypes = [v]
    lib.sdst_fetch.restype = C.c_int
    lib.sdst_fetch.argtypes = [v, v, v, v, v, v]
    lib.sdst_counts.restype = C.c_int
    lib.sdst_counts.argtypes = [v, v]
    _bound = True


def align_no_internal_gaps(aligner: GpuAligner, r1Idx, r2Idx, checkKmer=False, usingQuality=True, doingMatchQuality=False):
    """≙ alignRegGlobalNoInternalGaps(seq[r1Idx[i]], seq[r2Idx[i]]) + profileAlignment(...) -> OVERLAP_DTYPE[n]."""
    lib = aligner._lib
    _bind(lib)
    r1Idx = np.ascontiguousarray(r1Idx, np.uint32)
    r2Idx = np.ascontiguousarray(r2Idx, np.uint32)
    flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0) | \
            (_abi.MATCH_QUALITY if doingMatchQuality else 0)
    out = np.zeros(len(r1Idx), OVERLAP_DTYPE)
    aligner._check(lib.sdgpu_align_no_internal_gaps(aligner._ctx, _ptr(r1Idx), _ptr(r2Idx), len(r1Idx), flags, _ptr(out)))
    return out


def process_pairs(aligner: GpuAligner, r1, r2, pars: ProcessParams = None):
    """PairedReadProcessor::processPairedEnd for a batch: r1 / r2 = (bases, quals, offsets), mates reverse-complemented.
    -> dict(status u32[n], overlaps OVERLAP_DTYPE[n], bases, quals, offsets u64[n + 1], counts)."""
    lib = aligner._lib
    _bind(lib)
    pars = (pars or ProcessParams()).to_c()
    b1, q1, o1 = (np.ascontiguousarray(a, t) for a, t in zip(r1, (np.uint8, np.uint8, np.uint64)))
    b2, q2, o2 = (np.ascontiguousarray(a, t) for a, t in zip(r2, (np.uint8, np.uint8, np.uint64)))
    n = len(o1) - 1
    h = C.c_void_p()
    rc = lib.sdst_process_pairs(aligner._ctx, C.byref(pars), _ptr(b1), _ptr(q1), _ptr(o1), _ptr(b2), _ptr(q2), _ptr(o2), n, C.byref(h))
    if rc != _abi.OK:
        raise SdgpuError(rc, (lib.sdst_last_error() or b"").decode())
    try:
        tot = lib.sdst_total_len(h)
        res = {"status": np.zeros(n, np.uint32), "overlaps": np.zeros(n, OVERLAP_DTYPE), "bases": np.zeros(tot, np.uint8),
               "quals": np.zeros(tot, np.uint8), "offsets": np.zeros(n + 1, np.uint64)}
        lib.sdst_fetch(h, _ptr(res["status"]), _ptr(res["overlaps"]), _ptr(res["bases"]), _ptr(res["quals"]), _ptr(res["offsets"]))
        counts = np.zeros(7, np.uint64)
        lib.sdst_counts(h, _ptr(counts))
        res["counts"] = counts
        return res
    finally:
        lib.sdst_free(h)
