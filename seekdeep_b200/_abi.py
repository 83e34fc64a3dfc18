"""ctypes view of include/sdgpu.h.

PyTorch / Python are harness glue only: the product is libsdgpu.so (CUDA, sm_100a) behind the
C ABI.  Loading fails loudly when the library has not been built; there is no CPU fallback.
"""
import ctypes as C
import os

MAT_DIM = 128
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdgpu.so")

OK, E_ARG, E_CUDA, E_UNSUPPORTED, E_ALPHABET, E_OVERFLOW, E_STATE = range(7)
CHECK_KMER, USING_QUALITY, MATCH_QUALITY = 1, 2, 4
ST_GAPLIST_TRUNCATED = 1


class GapPars(C.Structure):
    """njhseq gapScoringParameters (ten penalties; extractorSetUp.cpp:138-150)."""
    _fields_ = [(n, C.c_int32) for n in (
        "gapOpen", "gapExtend", "gapLeftQueryOpen", "gapLeftQueryExtend", "gapRightQueryOpen",
        "gapRightQueryExtend", "gapLeftRefOpen", "gapLeftRefExtend", "gapRightRefOpen",
        "gapRightRefExtend")]


class Params(C.Structure):
    _fields_ = [("gaps", GapPars), ("scoreMat", C.c_int32 * (MAT_DIM * MAT_DIM)),
                ("primaryQual", C.c_uint32), ("secondaryQual", C.c_uint32),
                ("qualThresWindow", C.c_uint32), ("countEndGaps", C.c_uint32),
                ("weighHomopolymers", C.c_uint32), ("reserved", C.c_uint32)]


class Gap(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("size", C.c_uint32), ("gapInA", C.c_uint32)]


class Comparison(C.Structure):
    _fields_ = [("alnScore", C.c_int32)] + [(n, C.c_uint32) for n in (
        "hqMismatches", "lqMismatches", "lowKmerMismatches", "highQualityMatches",
        "lowQualityMatches", "numGaps", "basesInAln", "overLappingEvents", "alnLen", "nGapRuns",
        "status")] + [(n, C.c_double) for n in (
            "oneBaseIndel", "twoBaseIndel", "largeBaseIndel", "eventBasedIdentity",
            "eventBasedIdentityHq", "refCoverage", "queryCoverage")]


INT_FIELDS = [n for n, t in Comparison._fields_ if t is not C.c_double]
FLOAT_FIELDS = [n for n, t in Comparison._fields_ if t is C.c_double]

# numpy structured dtype with the same layout (104 bytes)
import numpy as _np  # noqa: E402

COMPARISON_DTYPE = _np.dtype([("alnScore", "<i4")] + [(n, "<u4") for n in INT_FIELDS[1:]] +
                             [(n, "<f8") for n in FLOAT_FIELDS])
GAP_DTYPE = _np.dtype([("pos", "<u4"), ("size", "<u4"), ("gapInA", "<u4")])
PAIR_SUMMARY_DTYPE = _np.dtype([("alnScore", "<i4"), ("hqMismatches", "<u2"), ("lqMismatches", "<u2"),
                                ("lowKmerMismatches", "<u2"), ("numGaps", "<u2"), ("oneBaseIndel", "<f4"),
                                ("twoBaseIndel", "<f4"), ("largeBaseIndel", "<f4"), ("eventBasedIdentity", "<f4")])
CHIMERA_PROBE_DTYPE = _np.dtype([("alnScore", "<i4")] + [(n, "<u4") for n in (
    "keptMismatches", "numGaps", "firstPos", "lastPos", "firstAlnPos", "lastAlnPos", "flags")])
CHI_FRONT_PASS, CHI_BACK_PASS = 1, 2
EVENT_DTYPE = _np.dtype([("kind", "<u4"), ("alnPos", "<u4"), ("refPos", "<u4"), ("seqPos", "<u4"), ("size", "<u4"),
                         ("kmerFreqRef", "<u4"), ("kmerFreqSeq", "<u4"), ("refBase", "u1"), ("seqBase", "u1"),
                         ("refQual", "u1"), ("seqQual", "u1")])
assert EVENT_DTYPE.itemsize == 32
EV_MISMATCH_HQ, EV_MISMATCH_LQ, EV_MISMATCH_LOWKMER, EV_GAP_IN_REF, EV_GAP_IN_READ, EV_END_GAP = 1, 2, 3, 4, 5, 0x10
GRID_HIT_DTYPE = _np.dtype([("read", "<u4"), ("cluster", "<u4"), ("mask", "<u8")])
assert GRID_HIT_DTYPE.itemsize == 16
assert CHIMERA_PROBE_DTYPE.itemsize == 32
assert PAIR_SUMMARY_DTYPE.itemsize == 28
assert COMPARISON_DTYPE.itemsize == C.sizeof(Comparison) == 104
assert GAP_DTYPE.itemsize == C.sizeof(Gap) == 12

_u8p, _u32p, _u64p, _f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_double))
_ctx = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/sdgpu.h declares
SIGNATURES = {
    "sdgpu_abi_version": (C.c_int, []),
    "sdgpu_device_count": (C.c_int, []),
    "sdgpu_create": (C.c_int, [C.c_int, C.POINTER(Params), C.POINTER(_ctx)]),
    "sdgpu_destroy": (None, [_ctx]),
    "sdgpu_last_error": (C.c_char_p, [_ctx]),
    "sdgpu_set_params": (C.c_int, [_ctx, C.POINTER(Params)]),
    "sdgpu_get_params": (C.c_int, [_ctx, C.POINTER(Params)]),
    "sdgpu_upload_seqs": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "sdgpu_upload_seqs_gather": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "sdgpu_append_seqs": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, _u32p]),
    "sdgpu_get_alphabet": (C.c_int, [_ctx, C.c_void_p, _u32p]),
    "sdgpu_num_seqs": (C.c_int, [_ctx, _u32p]),
    "sdgpu_update_cnts": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32]),
    "sdgpu_build_kmer_index": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32]),
    "sdgpu_kmer_index_lookup": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "sdgpu_build_kmer_profiles": (C.c_int, [_ctx, C.c_uint32]),
    "sdgpu_kmer_compare": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "sdgpu_kmer_compare_all": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "sdgpu_build_kmer_lists": (C.c_int, [_ctx, C.c_uint32]),
    "sdgpu_kmer_bin_masks": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p]),
    "sdgpu_align_profile": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]),
    "sdgpu_align_gap_lists": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), _u64p]),
    "sdgpu_align_events": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "sdgpu_set_pass_table": (C.c_int, [_ctx, C.c_void_p, C.c_uint32]),
    "sdgpu_align_pass": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "sdgpu_align_pass_grid": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), _u64p]),
    "sdgpu_align_pass_grid_stream": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "sdgpu_set_screen": (C.c_int, [_ctx, C.c_int]),
    "sdgpu_align_no_internal_gaps": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "sdgpu_chimera_probe_pairs": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]),
    "sdgpu_kmer_compare_all_dev": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "sdgpu_all_vs_all": (C.c_int, [_ctx, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]),
    "sdgpu_align_pass_submit": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "sdgpu_align_pass_wait": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_void_p), _u32p]),
    "sdgpu_align_profile_dev": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "sdgpu_dev_alloc": (C.c_int, [_ctx, C.c_uint64, C.POINTER(C.c_void_p)]),
    "sdgpu_dev_free": (C.c_int, [_ctx, C.c_void_p]),
    "sdgpu_memcpy_h2d": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint64]),
    "sdgpu_memcpy_d2h": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint64]),
    "sdgpu_sync": (C.c_int, [_ctx]),
    "sdgpu_timer_start": (C.c_int, [_ctx]),
    "sdgpu_timer_stop": (C.c_int, [_ctx, C.POINTER(C.c_float)]),
    "sdgpu_counters": (C.c_int, [_ctx, _u64p, _u64p]),
    "sdgpu_set_kernel": (C.c_int, [_ctx, C.c_int]),
    "sdgpu_last_kernel": (C.c_int, [_ctx]),
    "sdgpu_set_band": (C.c_int, [_ctx, C.c_int]),
    "sdgpu_profile_enable": (C.c_int, [_ctx, C.c_int]),
    "sdgpu_profile_read": (C.c_int, [_ctx, C.c_void_p, C.c_int]),
    "sdgpu_measure_int32_peak": (C.c_int, [_ctx, C.c_int, _f64p]),
}

_lib = None


def load():
    """dlopen libsdgpu.so (built in-tree by seekdeep_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C seekdeep_b200/csrc). There is no CPU fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
