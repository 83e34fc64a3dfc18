"""ctypes binding of oracle/_build/libsdoracle.so — the CPU checker (test infrastructure).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

from seekdeep_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libsdoracle.so")


# POD layouts are shared with the product's headers (include/sdgpu_qluster.h ≙ oracle/sd_oracle.h)
from seekdeep_b200.qluster import IterParC, QlusterOptsC  # noqa: E402


def iter_pars_to_c(iters):
    arr = (IterParC * len(iters))()
    for i, it in enumerate(iters):
        arr[i] = IterParC(it.stopCheck, it.smallCutoff, it.oneBaseIndel, it.twoBaseIndel, it.largeBaseIndel,
                          it.hqMismatches, it.lqMismatches, it.lowKmerMismatches, int(it.onPerId), int(it.onHqPerId),
                          it.perId)
    return arr


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        v = C.c_void_p
        lib.sdo_align_global.restype = C.c_int
        lib.sdo_align_global.argtypes = [C.POINTER(_abi.Params), v, C.c_uint32, v, C.c_uint32, C.POINTER(C.c_int32), v, C.c_uint32]
        lib.sdo_align_profile.restype = C.c_int
        lib.sdo_align_profile.argtypes = [C.POINTER(_abi.Params), v, v, v, C.c_uint32, v, v, C.c_uint32, C.c_uint32,
                                          C.POINTER(_abi.Comparison), v, C.c_uint32, C.c_char_p, C.c_char_p]
        lib.sdo_align_profile_batch.restype = C.c_int
        lib.sdo_align_profile_batch.argtypes = [C.POINTER(_abi.Params), v, v, v, v, v, v, C.c_uint32, C.c_uint32, v, v, C.c_uint32]
        lib.sdo_index_kmers.restype = v
        lib.sdo_index_kmers.argtypes = [v, v, v, v, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32]
        lib.sdo_free_kmaps.argtypes = [v]
        lib.sdo_kmaps_count.restype = C.c_uint32
        lib.sdo_kmaps_count.argtypes = [v, v, C.c_uint32]
        lib.sdo_kmer_compare.restype = C.c_int
        lib.sdo_kmer_compare.argtypes = [v, C.c_uint32, v, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        lib.sdo_kmer_compare_revcomp.restype = C.c_int
        lib.sdo_kmer_compare_revcomp.argtypes = lib.sdo_kmer_compare.argtypes
        lib.sdo_qluster.restype = v
        lib.sdo_qluster.argtypes = [C.POINTER(_abi.Params), C.POINTER(QlusterOptsC), v, C.c_uint32, v, C.c_uint32, v, v, v, C.c_uint32]
        lib.sdo_qluster_free.argtypes = [v]
        for name, res in (("num_clusters", C.c_uint32), ("num_unique", C.c_uint32), ("num_alignments", C.c_uint64),
                          ("num_cells", C.c_uint64)):
            f = getattr(lib, "sdo_qluster_" + name)
            f.restype = res
            f.argtypes = [v]
        lib.sdo_qluster_cluster_len.restype = C.c_uint32
        lib.sdo_qluster_cluster_len.argtypes = [v, C.c_uint32]
        lib.sdo_qluster_cluster_cnt.restype = C.c_double
        lib.sdo_qluster_cluster_cnt.argtypes = [v, C.c_uint32]
        lib.sdo_qluster_cluster_seq.argtypes = [v, C.c_uint32, v, v]
        lib.sdo_qluster_membership.argtypes = [v, v]
        lib.sdo_qluster_cluster_chimeric.restype = C.c_int
        lib.sdo_qluster_cluster_chimeric.argtypes = [v, C.c_uint32, v, v]
        lib.sdo_chimera_probe_batch.restype = C.c_int
        lib.sdo_chimera_probe_batch.argtypes = [C.POINTER(_abi.Params), v, v, v, v, v, v, C.c_uint32, C.c_uint32,
                                                C.POINTER(IterParC), C.c_int, v]

    def stitch_pairs(self, params, stitchPars, r1, r2):
        """PairedReadProcessor::processPairedEnd restated: (status, overlaps, list of (seq bytes, quals)), retried count."""
        from seekdeep_b200 import stitch as ST
        lib = self.lib
        v = C.c_void_p
        lib.sdo_stitch_pairs.restype = v
        lib.sdo_stitch_pairs.argtypes = [C.POINTER(_abi.Params), C.POINTER(ST.ProcessParamsC), v, v, v, v, v, v, C.c_uint32]
        lib.sdo_stitch_free.argtypes = [v]
        lib.sdo_stitch_retried.restype = C.c_uint64
        lib.sdo_stitch_retried.argtypes = [v]
        lib.sdo_stitch_status.restype = C.c_uint32
        lib.sdo_stitch_status.argtypes = [v, C.c_uint32]
        lib.sdo_stitch_overlap.argtypes = [v, C.c_uint32, v]
        lib.sdo_stitch_len.restype = C.c_uint32
        lib.sdo_stitch_len.argtypes = [v, C.c_uint32]
        lib.sdo_stitch_seq.argtypes = [v, C.c_uint32, v, v]
        b1, q1, o1 = (np.ascontiguousarray(a, t) for a, t in zip(r1, (np.uint8, np.uint8, np.uint64)))
        b2, q2, o2 = (np.ascontiguousarray(a, t) for a, t in zip(r2, (np.uint8, np.uint8, np.uint64)))
        n = len(o1) - 1
        sp = stitchPars.to_c()
        h = lib.sdo_stitch_pairs(C.byref(params), C.byref(sp), _p(b1), _p(q1), _p(o1), _p(b2), _p(q2), _p(o2), n)
        try:
            status = np.array([lib.sdo_stitch_status(h, i) for i in range(n)], np.uint32)
            overlaps = np.zeros(n, ST.OVERLAP_DTYPE)
            seqs = []
            for i in range(n):
                lib.sdo_stitch_overlap(h, i, C.c_void_p(overlaps[i:i + 1].ctypes.data))
                ln = lib.sdo_stitch_len(h, i)
                s, q = np.zeros(ln, np.uint8), np.zeros(ln, np.uint8)
                if ln:
                    lib.sdo_stitch_seq(h, i, _p(s), _p(q))
                seqs.append((s.tobytes(), q))
            return status, overlaps, seqs, lib.sdo_stitch_retried(h)
        finally:
            lib.sdo_stitch_free(h)

    # one pair, strings in
    def align_global(self, params, a, b, gapCap=64):
        a = a.encode() if isinstance(a, str) else bytes(a)
        b = b.encode() if isinstance(b, str) else bytes(b)
        score = C.c_int32()
        gaps = np.zeros(gapCap, _abi.GAP_DTYPE)
        n = self.lib.sdo_align_global(C.byref(params), a, len(a), b, len(b), C.byref(score), _p(gaps), gapCap)
        assert n >= 0
        return score.value, gaps[:min(n, gapCap)]

    def align_profile(self, params, a, qa, b, qb, flags=_abi.USING_QUALITY, kmaps=None, gapCap=64):
        a = a.encode() if isinstance(a, str) else bytes(a)
        b = b.encode() if isinstance(b, str) else bytes(b)
        qa = None if qa is None else np.ascontiguousarray(qa, np.uint8)
        qb = None if qb is None else np.ascontiguousarray(qb, np.uint8)
        comp = _abi.Comparison()
        gaps = np.zeros(gapCap, _abi.GAP_DTYPE)
        alnA = C.create_string_buffer(len(a) + len(b) + 1)
        alnB = C.create_string_buffer(len(a) + len(b) + 1)
        rc = self.lib.sdo_align_profile(C.byref(params), kmaps, a, _p(qa), len(a), b, _p(qb), len(b), flags,
                                        C.byref(comp), _p(gaps), gapCap, alnA, alnB)
        assert rc == 0
        return comp, gaps[:min(comp.nGapRuns, gapCap)], alnA.value.decode(), alnB.value.decode()

    def align_profile_batch(self, params, bases, quals, offsets, refIdx, readIdx, flags=_abi.USING_QUALITY, kmaps=None,
                            gapCap=0):
        bases = np.ascontiguousarray(bases, np.uint8)
        quals = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        refIdx = np.ascontiguousarray(refIdx, np.uint32)
        readIdx = np.ascontiguousarray(readIdx, np.uint32)
        n = len(refIdx)
        out = np.zeros(n, _abi.COMPARISON_DTYPE)
        gaps = np.zeros((n, gapCap), _abi.GAP_DTYPE) if gapCap else None
        rc = self.lib.sdo_align_profile_batch(C.byref(params), kmaps, _p(bases), _p(quals), _p(offsets), _p(refIdx),
                                              _p(readIdx), n, flags, _p(out), _p(gaps), gapCap)
        assert rc == 0
        return (out, gaps) if gapCap else out

    def align_events(self, params, a, qa, b, qb, flags=_abi.USING_QUALITY, kmaps=None):
        """(comparison, events) of one pair: comp_'s mismatches_ / lowKmerMismatches_ / alignmentGaps_ in column order."""
        self.lib.sdo_align_events.restype = C.c_int
        self.lib.sdo_align_events.argtypes = [C.POINTER(_abi.Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        qa = None if qa is None else np.ascontiguousarray(qa, np.uint8)
        qb = None if qb is None else np.ascontiguousarray(qb, np.uint8)
        cmp_ = np.zeros(1, _abi.COMPARISON_DTYPE)
        cap = len(a) + len(b) + 2
        ev = np.zeros(cap, _abi.EVENT_DTYPE)
        n = self.lib.sdo_align_events(C.byref(params), kmaps, _p(a), _p(qa), len(a), _p(b), _p(qb), len(b), flags, _p(cmp_),
                                      _p(ev), cap)
        assert 0 <= n <= cap
        return cmp_[0], ev[:n]

    def chimera_probe_batch(self, params, bases, quals, offsets, parentIdx, candIdx, chiOverlap, keepLowQual=False,
                            flags=_abi.USING_QUALITY, kmaps=None):
        bases = np.ascontiguousarray(bases, np.uint8)
        quals = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        parentIdx = np.ascontiguousarray(parentIdx, np.uint32)
        candIdx = np.ascontiguousarray(candIdx, np.uint32)
        out = np.zeros(len(parentIdx), _abi.CHIMERA_PROBE_DTYPE)
        rc = self.lib.sdo_chimera_probe_batch(C.byref(params), kmaps, _p(bases), _p(quals), _p(offsets), _p(parentIdx),
                                              _p(candIdx), len(parentIdx), flags, C.byref(chiOverlap),
                                              int(keepLowQual), _p(out))
        assert rc == 0
        return out

    def index_kmers(self, bases, offsets, cnts, kLen, runCutOff, byPos=True, expandPos=False, expandSize=0, seqIdx=None):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        cnts = None if cnts is None else np.ascontiguousarray(cnts, np.float64)
        idx = None if seqIdx is None else np.ascontiguousarray(seqIdx, np.uint32)
        n = len(offsets) - 1 if idx is None else len(idx)
        return self.lib.sdo_index_kmers(_p(bases), _p(offsets), _p(cnts), _p(idx), n, kLen, runCutOff, int(byPos),
                                        int(expandPos), expandSize)

    def free_kmaps(self, k):
        self.lib.sdo_free_kmaps(k)

    def kmaps_count(self, k, kmer, pos):
        kmer = kmer.encode() if isinstance(kmer, str) else bytes(kmer)
        return self.lib.sdo_kmaps_count(k, kmer, pos)

    def kmer_compare(self, a, b, k, revCompB=False):
        a = a.encode() if isinstance(a, str) else bytes(a)
        b = b.encode() if isinstance(b, str) else bytes(b)
        s, f = C.c_uint32(), C.c_double()
        (self.lib.sdo_kmer_compare_revcomp if revCompB else self.lib.sdo_kmer_compare)(a, len(a), b, len(b), k, C.byref(s), C.byref(f))
        return s.value, f.value

    def population_cluster(self, params, opts, popPars, bases, quals, offsets, cnts):
        """doPopulationClustering's collapse over packed input clusters (counts per cluster, no dedupe)."""
        self.lib.sdo_population_cluster.restype = C.c_void_p
        self.lib.sdo_population_cluster.argtypes = [C.POINTER(_abi.Params), C.POINTER(QlusterOptsC), C.c_void_p, C.c_uint32,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        bases = np.ascontiguousarray(bases, np.uint8)
        quals = np.ascontiguousarray(quals, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        cnts = np.ascontiguousarray(cnts, np.float64)
        pp = iter_pars_to_c(list(popPars))
        n = len(offsets) - 1
        h = self.lib.sdo_population_cluster(C.byref(params), C.byref(opts), pp, len(pp), _p(bases), _p(quals), _p(offsets), _p(cnts), n)
        return self._qluster_result(h, n)

    def qluster(self, params, opts, initPars, pars, bases, quals, offsets):
        bases = np.ascontiguousarray(bases, np.uint8)
        quals = np.ascontiguousarray(quals, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        ip, pp = iter_pars_to_c(list(initPars)), iter_pars_to_c(list(pars))
        n = len(offsets) - 1
        h = self.lib.sdo_qluster(C.byref(params), C.byref(opts), ip, len(ip), pp, len(pp), _p(bases), _p(quals),
                                 _p(offsets), n)
        return self._qluster_result(h, n)

    def _qluster_result(self, h, n):
        try:
            nc = self.lib.sdo_qluster_num_clusters(h)
            res = {"nUnique": self.lib.sdo_qluster_num_unique(h), "nAlign": self.lib.sdo_qluster_num_alignments(h),
                   "nCells": self.lib.sdo_qluster_num_cells(h), "clusters": []}
            for c in range(nc):
                ln = self.lib.sdo_qluster_cluster_len(h, c)
                s = np.zeros(ln, np.uint8)
                q = np.zeros(ln, np.uint8)
                self.lib.sdo_qluster_cluster_seq(h, c, _p(s), _p(q))
                res["clusters"].append((s.tobytes().decode(), q, self.lib.sdo_qluster_cluster_cnt(h, c)))
            res["chimeras"] = {}
            for c in range(nc):
                par, pos = (C.c_uint32 * 2)(), (C.c_uint32 * 2)()
                if self.lib.sdo_qluster_cluster_chimeric(h, c, par, pos):
                    res["chimeras"][c] = (par[0], par[1], pos[0], pos[1])
            mem = np.zeros(n, np.uint32)
            self.lib.sdo_qluster_membership(h, _p(mem))
            res["membership"] = mem
            return res
        finally:
            self.lib.sdo_qluster_free(h)


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def load():
    if not os.path.exists(LIB):
        build()
    return Oracle(C.CDLL(LIB))
