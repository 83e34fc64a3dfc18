"""GpuAligner — Python mirror of the reference-side `aligner` usage over the C ABI.

The reference drives one pair at a time through stateful members
(`alignerObj.alignCacheGlobal(ref, read); alignerObj.profileAlignment(ref, read, ...)` then
reads `alignerObj.comp_`, clusterDown.cpp:739-753).  The GPU path takes pair *batches* and
returns one `comparison` record per pair; names and argument meaning follow the reference.
"""
import ctypes as C

import numpy as np

from . import _abi


class SdgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sdgpu error {code}: {msg}")
        self.code = code


def pack_seqs(seqs, quals=None):
    """list of str/bytes (+ list of uint8 arrays) -> (bases u8, quals u8|None, offsets u64)."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    q = None
    if quals is not None:
        q = np.concatenate([np.asarray(x, dtype=np.uint8) for x in quals]) if len(quals) else np.zeros(0, np.uint8)
        if q.size != bases.size:
            raise ValueError("quals and bases differ in total length")
    return bases, q, offsets


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GpuAligner:
    """One context per host thread / GPU, like one njhseq aligner per thread (AlignerPool)."""

    def __init__(self, params, device=0):
        self._lib = _abi.load()
        self._ctx = C.c_void_p()
        self.params = params
        rc = self._lib.sdgpu_create(device, C.byref(params), C.byref(self._ctx))
        if rc != _abi.OK:
            msg = self._lib.sdgpu_last_error(None)
            self._ctx = None
            raise SdgpuError(rc, msg.decode() if msg else "")

    # -- plumbing
    def _check(self, rc):
        if rc != _abi.OK:
            raise SdgpuError(rc, self._lib.sdgpu_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.sdgpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- sequence store
    def upload_seqs(self, bases, quals, offsets, cnts=None):
        bases = np.ascontiguousarray(bases, np.uint8)
        quals = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        cnts = None if cnts is None else np.ascontiguousarray(cnts, np.float64)
        self._check(self._lib.sdgpu_upload_seqs(self._ctx, _ptr(bases), _ptr(quals), _ptr(offsets), _ptr(cnts),
                                                len(offsets) - 1))

    def append_seqs(self, bases, quals, offsets, cnts=None):
        bases = np.ascontiguousarray(bases, np.uint8)
        quals = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        cnts = None if cnts is None else np.ascontiguousarray(cnts, np.float64)
        first = C.c_uint32()
        self._check(self._lib.sdgpu_append_seqs(self._ctx, _ptr(bases), _ptr(quals), _ptr(offsets), _ptr(cnts),
                                                len(offsets) - 1, C.byref(first)))
        return first.value

    def num_seqs(self):
        n = C.c_uint32()
        self._check(self._lib.sdgpu_num_seqs(self._ctx, C.byref(n)))
        return n.value

    def set_params(self, params):
        self.params = params
        self._check(self._lib.sdgpu_set_params(self._ctx, C.byref(params)))

    # -- k-mers
    def index_kmers(self, kLength, runCutOff, kmersByPosition=True, expandKmerPos=False, expandKmerSize=0, seqIdx=None):
        """≙ alignerObj.kMaps_ = indexKmers(clusters, kLength, runCutOff, kmersByPosition, expandKmerPos, expandKmerSize)."""
        idx = None if seqIdx is None else np.ascontiguousarray(seqIdx, np.uint32)
        self._check(self._lib.sdgpu_build_kmer_index(self._ctx, _ptr(idx), 0 if idx is None else len(idx), kLength,
                                                     runCutOff, int(kmersByPosition), int(expandKmerPos), expandKmerSize))

    def kmer_index_lookup(self, kmers, positions):
        k = np.frombuffer(b"".join(x.encode() if isinstance(x, str) else x for x in kmers), np.uint8).copy()
        pos = np.ascontiguousarray(positions, np.uint32)
        out = np.zeros(len(pos), np.uint32)
        self._check(self._lib.sdgpu_kmer_index_lookup(self._ctx, _ptr(k), _ptr(pos), len(pos), _ptr(out)))
        return out

    def build_kmer_profiles(self, kLen):
        """≙ kmerInfo(seq, kLen, false) for every stored sequence."""
        self._check(self._lib.sdgpu_build_kmer_profiles(self._ctx, kLen))

    def compare_kmers(self, a, b, revCompB=False):
        """≙ [seq[a[i]].compareKmers(seq[b[i]]) for i] -> (shared u32[n], frac f64[n]); revCompB:
        compareKmersRevComp (a's k-mers against the k-mers of b's reverse complement)."""
        a = np.ascontiguousarray(a, np.uint32)
        b = np.ascontiguousarray(b, np.uint32)
        shared = np.zeros(len(a), np.uint32)
        frac = np.zeros(len(a), np.float64)
        self._check(self._lib.sdgpu_kmer_compare(self._ctx, _ptr(a), _ptr(b), len(a), int(revCompB), _ptr(shared), _ptr(frac)))
        return shared, frac

    def compare_kmers_all(self, reads, clusters):
        reads = np.ascontiguousarray(reads, np.uint32)
        clusters = np.ascontiguousarray(clusters, np.uint32)
        shared = np.zeros((len(reads), len(clusters)), np.uint32)
        frac = np.zeros((len(reads), len(clusters)), np.float64)
        self._check(self._lib.sdgpu_kmer_compare_all(self._ctx, _ptr(reads), len(reads), _ptr(clusters), len(clusters),
                                                     _ptr(shared), _ptr(frac)))
        return shared, frac

    def build_kmer_lists(self, kLen):
        """≙ kmerInfo(seq, kLen, false) for every stored sequence, any kLen <= 16: sorted k-mer lists on the device."""
        self._check(self._lib.sdgpu_build_kmer_lists(self._ctx, int(kLen)))

    def kmer_bin_masks(self, seeds, reads, cutOff, want_shared=False):
        """≙ seed.compareKmers(read).second > cutOff for every (read, seed): uint64 mask per read (bit c = seed c);
        with want_shared also the shared k-mer counts [nReads, nSeeds] (qluster --useKmerBinning)."""
        seeds = np.ascontiguousarray(seeds, np.uint32)
        reads = np.ascontiguousarray(reads, np.uint32)
        masks = np.zeros(len(reads), np.uint64)
        shared = np.zeros((len(reads), len(seeds)), np.uint32) if want_shared else None
        self._check(self._lib.sdgpu_kmer_bin_masks(self._ctx, _ptr(seeds), len(seeds), _ptr(reads), len(reads), float(cutOff),
                                                   _ptr(masks), _ptr(shared)))
        return (masks, shared) if want_shared else masks

    # -- alignment + errorProfile
    def align_profile(self, refIdx, readIdx, checkKmer=False, usingQuality=True, doingMatchQuality=False, gapCap=0):
        """≙ for each pair: alignCacheGlobal(ref, read); profileAlignment(ref, read, checkKmer,
        usingQuality, doingMatchQuality) -> comp_.  Returns a structured array (COMPARISON_DTYPE)
        and, if gapCap > 0, the gap runs (GAP_DTYPE[nPairs, gapCap])."""
        refIdx = np.ascontiguousarray(refIdx, np.uint32)
        readIdx = np.ascontiguousarray(readIdx, np.uint32)
        n = len(refIdx)
        flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0) | \
                (_abi.MATCH_QUALITY if doingMatchQuality else 0)
        out = np.zeros(n, _abi.COMPARISON_DTYPE)
        gaps = np.zeros((n, gapCap), _abi.GAP_DTYPE) if gapCap else None
        self._check(self._lib.sdgpu_align_profile(self._ctx, _ptr(refIdx), _ptr(readIdx), n, flags, _ptr(out),
                                                  _ptr(gaps), gapCap))
        return (out, gaps) if gapCap else out

    def align_gap_lists(self, refIdx, readIdx, gapCap=8):
        """≙ alignCacheGlobal per pair, alnInfoGlobal::gapInfos_ only, packed (sdgpu_align_gap_lists): {pair: (nRuns, gaps)}
        for the pairs whose alignment has gap runs; gaps is GAP_DTYPE[min(nRuns, gapCap)] in traceback order."""
        refIdx = np.ascontiguousarray(refIdx, np.uint32)
        readIdx = np.ascontiguousarray(readIdx, np.uint32)
        recs, nWords = C.c_void_p(), C.c_uint64()
        self._check(self._lib.sdgpu_align_gap_lists(self._ctx, _ptr(refIdx), _ptr(readIdx), len(refIdx), gapCap, C.byref(recs),
                                                    C.byref(nWords)))
        out = {}
        if nWords.value:
            w = np.ctypeslib.as_array(C.cast(recs, C.POINTER(C.c_uint32)), (nWords.value,)).copy()
            i = 0
            while i < len(w):
                n = int(w[i + 1] & 0x7fffffff)
                kept = min(n, gapCap)
                g = np.zeros(kept, _abi.GAP_DTYPE)
                body = w[i + 2:i + 2 + 3 * kept].reshape(kept, 3)
                g["pos"], g["size"], g["gapInA"] = body[:, 0], body[:, 1], body[:, 2]
                out[int(w[i])] = (n, g)
                i += 2 + 3 * kept
        return out

    def align_events(self, refIdx, readIdx, checkKmer=False, usingQuality=True, doingMatchQuality=False, eventCap=64):
        """≙ alignCacheGlobal + profileAlignment, then comp_.distances_.mismatches_ / lowKmerMismatches_ /
        alignmentGaps_ as event lists (clusterDown.cpp:759-800).  Returns (comparisons, [events of pair i])."""
        refIdx = np.ascontiguousarray(refIdx, np.uint32)
        readIdx = np.ascontiguousarray(readIdx, np.uint32)
        n = len(refIdx)
        flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0) | \
                (_abi.MATCH_QUALITY if doingMatchQuality else 0)
        out = np.zeros(n, _abi.COMPARISON_DTYPE)
        nEv = np.zeros(n, np.uint32)
        while True:
            ev = np.zeros((n, eventCap), _abi.EVENT_DTYPE)
            self._check(self._lib.sdgpu_align_events(self._ctx, _ptr(refIdx), _ptr(readIdx), n, flags, _ptr(out), _ptr(ev),
                                                     eventCap, _ptr(nEv)))
            if n == 0 or nEv.max() <= eventCap:
                return out, [ev[i, :nEv[i]] for i in range(n)]
            eventCap = int(nEv.max())

    # -- device-resident benchmarking helpers
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self._lib.sdgpu_dev_alloc(self._ctx, nbytes, C.byref(p)))
        return p

    def dev_free(self, p):
        self._check(self._lib.sdgpu_dev_free(self._ctx, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self._lib.sdgpu_memcpy_h2d(self._ctx, dptr, _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        self._check(self._lib.sdgpu_memcpy_d2h(self._ctx, _ptr(arr), dptr, arr.nbytes))

    def align_profile_dev(self, dRef, dRead, nPairs, dOut, flags=_abi.USING_QUALITY):
        self._check(self._lib.sdgpu_align_profile_dev(self._ctx, dRef, dRead, nPairs, flags, dOut))

    def sync(self):
        self._check(self._lib.sdgpu_sync(self._ctx))

    def timer_start(self):
        self._check(self._lib.sdgpu_timer_start(self._ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self._lib.sdgpu_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    def set_pass_table(self, iters):
        """Upload par-file lines (IterPar list) for on-device comparison::passErrorProfile."""
        from .qluster import iter_pars_to_c
        arr, n = iter_pars_to_c(iters)
        self._check(self._lib.sdgpu_set_pass_table(self._ctx, C.cast(arr, C.c_void_p), n))

    def align_pass(self, refIdx, readIdx, checkKmer=False, usingQuality=True, doingMatchQuality=False):
        """Like align_profile, but returns uint64 masks: bit l <=> errorProfile passes table line l."""
        refIdx = np.ascontiguousarray(refIdx, np.uint32)
        readIdx = np.ascontiguousarray(readIdx, np.uint32)
        flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0) | \
                (_abi.MATCH_QUALITY if doingMatchQuality else 0)
        out = np.zeros(len(refIdx), np.uint64)
        self._check(self._lib.sdgpu_align_pass(self._ctx, _ptr(refIdx), _ptr(readIdx), len(refIdx), flags, _ptr(out)))
        return out

    def align_pass_grid(self, clusterIdx, readIdx, checkKmer=False, usingQuality=True):
        """Every read against every cluster (the collapse loop's pair shape; a sequence is never compared with
        itself): GRID_HIT_DTYPE records (read ordinal, cluster ordinal, mask) of the non-zero verdict masks only."""
        clusterIdx = np.ascontiguousarray(clusterIdx, np.uint32)
        readIdx = np.ascontiguousarray(readIdx, np.uint32)
        flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0)
        hits, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.sdgpu_align_pass_grid(self._ctx, _ptr(clusterIdx), len(clusterIdx), _ptr(readIdx), len(readIdx),
                                                    flags, C.byref(hits), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, _abi.GRID_HIT_DTYPE)
        buf = (C.c_uint8 * (n.value * 16)).from_address(hits.value)
        return np.frombuffer(buf, _abi.GRID_HIT_DTYPE).copy()

    def set_screen(self, mode):
        """0 = screen pass in front of the verdict-mask launches (default), 1 = off."""
        self._check(self._lib.sdgpu_set_screen(self._ctx, int(mode)))

    def chimera_probe(self, parentIdx, candIdx, chiOverlap, keepLowQualityMismatches=False, checkKmer=False,
                      usingQuality=True):
        """collapser::markChimeras' per-(parent, candidate) probe (clusterDown.cpp:576): CHIMERA_PROBE_DTYPE[nPairs].
        `chiOverlap` is an IterPar-like object / ctypes struct holding chiOpts_.chiOverlap_'s thresholds."""
        parentIdx = np.ascontiguousarray(parentIdx, np.uint32)
        candIdx = np.ascontiguousarray(candIdx, np.uint32)
        flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0)
        out = np.zeros(len(parentIdx), _abi.CHIMERA_PROBE_DTYPE)
        self._check(self._lib.sdgpu_chimera_probe_pairs(self._ctx, _ptr(parentIdx), _ptr(candIdx), len(parentIdx), flags,
                                                        C.cast(C.pointer(chiOverlap), C.c_void_p),
                                                        int(keepLowQualityMismatches), _ptr(out)))
        return out

    def all_vs_all(self, seqIdx=None, n=None, firstPair=0, nPairs=None, checkKmer=False, usingQuality=True):
        """Upper-triangle all-vs-all alignment (config 3).  Returns PAIR_SUMMARY_DTYPE[nPairs] for the pair
        block [firstPair, firstPair + nPairs) in row-major (i < j) order."""
        idx = None if seqIdx is None else np.ascontiguousarray(seqIdx, np.uint32)
        n = (self.num_seqs() if idx is None else len(idx)) if n is None else n
        total = n * (n - 1) // 2
        nPairs = total - firstPair if nPairs is None else nPairs
        flags = (_abi.CHECK_KMER if checkKmer else 0) | (_abi.USING_QUALITY if usingQuality else 0)
        out = np.zeros(nPairs, _abi.PAIR_SUMMARY_DTYPE)
        self._check(self._lib.sdgpu_all_vs_all(self._ctx, _ptr(idx), n, flags, firstPair, nPairs, _ptr(out)))
        return out

    def set_kernel(self, mode):
        """0 = automatic, 1 = int32 forward pass only, 2 = packed fp16 forward pass only."""
        self._check(self._lib.sdgpu_set_kernel(self._ctx, mode))

    def last_kernel(self):
        return self._lib.sdgpu_last_kernel(self._ctx)

    def profile_enable(self, on=True):
        self._check(self._lib.sdgpu_profile_enable(self._ctx, int(on)))

    def set_band(self, mode):
        """0 = banded first pass + exact redo (default), 1 = full-matrix kernels only."""
        self._check(self._lib.sdgpu_set_band(self._ctx, int(mode)))

    def profile_read(self, reset=True):
        class Prof(C.Structure):
            _fields_ = [("alignKernelMs", C.c_double), ("alignKernelLaunches", C.c_uint64), ("kernelLaunches", C.c_uint64),
                        ("cellUpdates", C.c_uint64), ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64),
                        ("evaluatedCells", C.c_uint64), ("bandCells", C.c_uint64), ("bandCertifiedCells", C.c_uint64),
                        ("bandKernelMs", C.c_double), ("screenKernelMs", C.c_double), ("screenPairs", C.c_uint64),
                        ("screenCertified", C.c_uint64), ("screenCells", C.c_uint64), ("screenCertCells", C.c_uint64),
                        ("screenScanMs", C.c_double)]
        p = Prof()
        self._check(self._lib.sdgpu_profile_read(self._ctx, C.byref(p), int(reset)))
        return {k: getattr(p, k) for k, _ in Prof._fields_}

    def measure_int32_peak(self, which=0):
        v = C.c_double()
        self._check(self._lib.sdgpu_measure_int32_peak(self._ctx, which, C.byref(v)))
        return v.value

    def counters(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.sdgpu_counters(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value
