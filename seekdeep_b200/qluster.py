"""Host-side mirror of `SeekDeep qluster` for one sample over the C ABI (include/sdgpu_qluster.h).

Reference: src/SeekDeepPrograms/SeekDeepProgram/clusterDown/clusterDown.cpp:230-563 (call
sequence) and clusterDownSetUp.cpp:254-312 (technology defaults).  Sharding of many samples over
GPUs follows the reference's own model: one independent qluster per sample
(SeekDeepUtilsRunner_setUpTarAmpAnalysis.cpp:998-1001, runMultipleCommands)."""
import ctypes as C
import os
import sys
import time
from dataclasses import dataclass

import numpy as np

from . import _abi
from .aligner import GpuAligner, SdgpuError, _ptr
from .iterpars import CollapseIterations


class IterParC(C.Structure):
    _fields_ = [("stopCheck", C.c_uint32), ("smallCutoff", C.c_uint32), ("oneBaseIndel", C.c_double),
                ("twoBaseIndel", C.c_double), ("largeBaseIndel", C.c_double), ("hqMismatches", C.c_double),
                ("lqMismatches", C.c_double), ("lowKmerMismatches", C.c_double), ("onPerId", C.c_uint32),
                ("onHqPerId", C.c_uint32), ("perId", C.c_double)]


class QlusterOptsC(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("kLen", "runCutOff", "kmersByPosition", "expandKmerPos", "expandKmerSize",
                                          "checkKmers", "startWithSingles", "leaveOutSinglets", "dontRecalLowFreq",
                                          "smallReadSize", "markChimeras", "chiKeepLowQual", "chiRunCutOff",
                                          "chiOverlapSizeCutoff", "chiPosSpacing", "converge")] + \
               [("chiParentFreqs", C.c_double), ("chiOverlap", IterParC), ("kmerCutOff", C.c_double),
                ("useKmerBinning", C.c_uint32), ("kCompareLen", C.c_uint32), ("binPars", C.c_void_p),
                ("nBinPars", C.c_uint64)]


class StatsC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("inputReads", "uniqueSeqs", "finalClusters", "pairsComputed", "pairsConsumed",
                                          "consensusAligns", "cellUpdates", "gpuBatches", "onDemandPairs",
                                          "chimeraPairs", "chimericClusters", "passingPairs", "kmerPairs",
                                          "kmerBins", "streamedGrids")] + \
               [("secondsGpu", C.c_double), ("secondsTotal", C.c_double), ("msResident", C.c_double)]


@dataclass
class QlusterPars:
    """The subset of clusterDownPars / seqSetUp options that shapes the hot path; defaults are
    `SeekDeep qluster --illumina` (kLength 9, runCutOff 1, k-mers by position, check k-mers)."""
    kLength: int = 9
    runCutOff: int = 1
    kmersByPosition: bool = True
    expandKmerPos: bool = False
    expandKmerSize: int = 5
    checkKmers: bool = True
    startWithSingles: bool = False
    leaveOutSinglets: bool = False
    dontRecalLowFreqMismatchAndReRun: bool = False
    converge: bool = False  # --converge: keep clustering at each iteration until there is no more collapsing
    smallReadSize: int = 18  # 2 * kLength (clusterDownSetUp.cpp:406)
    # pars_.chiOpts_ (clusterDownSetUp.cpp:372-381): on unless --noMarkChimeras, parentFreqs 2
    markChimeras: bool = True
    parFreqs: float = 2.0
    chiKeepLowQaulityMismatches: bool = False
    chiPosSpacing: int = 1
    chiRunCutOff: int = 1
    chiOverlapSizeCutoff: int = 5
    # chiOverlap_: njhseq's ChimeraOpts defaults + `largeBaseIndel_ = .99` (clusterDown.cpp:572)
    chiOverlap: tuple = (2.0, 1.0, 0.99, 0.0, 0.0, 1.0)  # one, two, large indel; hq, lq, lowKmer mismatches
    # colOpts_.kmerBinOpts_ (clusterDownSetUp.cpp:352-354); 10 / 0.80 are njhseq's defaults as far as memory serves
    useKmerBinning: bool = False
    kCompareLen: int = 10
    kmerCutOff: float = 0.80

    def to_c(self, binIteratorMap=None):
        """`binIteratorMap` (pars.binIteratorMap, --binPar) is attached as opts.binPars; the array is kept alive on the
        returned struct."""
        o = self.chiOverlap
        c = self._to_c_base(o)
        c.kmerCutOff = float(self.kmerCutOff)
        c.useKmerBinning = int(self.useKmerBinning)
        c.kCompareLen = int(self.kCompareLen)
        if binIteratorMap is not None:
            arr, n = iter_pars_to_c(binIteratorMap)
            c._binKeep = arr
            c.binPars = C.cast(arr, C.c_void_p)
            c.nBinPars = n
        return c

    def _to_c_base(self, o):
        return QlusterOptsC(self.kLength, self.runCutOff, int(self.kmersByPosition), int(self.expandKmerPos),
                            self.expandKmerSize, int(self.checkKmers), int(self.startWithSingles),
                            int(self.leaveOutSinglets), int(self.dontRecalLowFreqMismatchAndReRun), self.smallReadSize,
                            int(self.markChimeras), int(self.chiKeepLowQaulityMismatches), self.chiRunCutOff,
                            self.chiOverlapSizeCutoff, self.chiPosSpacing, int(self.converge), float(self.parFreqs),
                            IterParC(0, 0, o[0], o[1], o[2], o[3], o[4], o[5], 0, 0, 0.0))


def iter_pars_to_c(iters):
    iters = list(iters)
    arr = (IterParC * max(1, len(iters)))()
    for i, it in enumerate(iters):
        arr[i] = IterParC(it.stopCheck, it.smallCutoff, it.oneBaseIndel, it.twoBaseIndel, it.largeBaseIndel,
                          it.hqMismatches, it.lqMismatches, it.lowKmerMismatches, int(it.onPerId), int(it.onHqPerId),
                          it.perId)
    return arr, len(iters)


_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return
    v = C.c_void_p
    lib.sdq_run.restype = C.c_int
    lib.sdq_run.argtypes = [v, C.POINTER(QlusterOptsC), v, C.c_uint32, v, C.c_uint32, v, v, v, C.c_uint32, C.POINTER(v)]
    lib.sdq_free.argtypes = [v]
    lib.sdq_last_error.restype = C.c_char_p
    lib.sdq_get_stats.argtypes = [v, C.POINTER(StatsC)]
    lib.sdq_num_clusters.restype = C.c_uint32
    lib.sdq_num_clusters.argtypes = [v]
    lib.sdq_cluster_len.restype = C.c_uint32
    lib.sdq_cluster_len.argtypes = [v, C.c_uint32]
    lib.sdq_cluster_cnt.restype = C.c_double
    lib.sdq_cluster_cnt.argtypes = [v, C.c_uint32]
    lib.sdq_cluster_seq.argtypes = [v, C.c_uint32, v, v]
    lib.sdq_membership.argtypes = [v, v]
    lib.sdq_population_cluster_packed.restype = C.c_int
    lib.sdq_population_cluster_packed.argtypes = [v, C.POINTER(QlusterOptsC), v, C.c_uint32, v, v, v, v, C.c_uint32, C.POINTER(v)]
    lib.sdq_clusters_total_len.restype = C.c_uint64
    lib.sdq_clusters_total_len.argtypes = [v]
    lib.sdq_clusters_packed.restype = C.c_int
    lib.sdq_clusters_packed.argtypes = [v, v, v, v, v]
    lib.sdq_cluster_chimeric.restype = C.c_int
    lib.sdq_cluster_chimeric.argtypes = [v, C.c_uint32, v, v]
    _bound = True


def qluster(aligner: GpuAligner, bases, quals, offsets, pars: QlusterPars = None, iteratorMap: CollapseIterations = None,
            intialParameters: CollapseIterations = None, want_clusters=True, binIteratorMap: CollapseIterations = None):
    """Run the collapse of one sample on `aligner`'s GPU.  `iteratorMap` / `intialParameters` are
    the reference's names (clusterDownPars, setUpPars.hpp:126-127); both default to
    genIlluminaDefaultPars(100) like `--illumina`."""
    lib = aligner._lib
    _bind(lib)
    pars = pars or QlusterPars()
    iteratorMap = iteratorMap or CollapseIterations.gen_illumina_default_pars(100)
    intialParameters = intialParameters or iteratorMap
    bases = np.ascontiguousarray(bases, np.uint8)
    quals = np.ascontiguousarray(quals, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    opts = pars.to_c(binIteratorMap)
    ip, nip = iter_pars_to_c(intialParameters)
    pp, npp = iter_pars_to_c(iteratorMap)
    h = C.c_void_p()
    t0 = time.perf_counter()
    rc = lib.sdq_run(aligner._ctx, C.byref(opts), ip, nip, pp, npp, _ptr(bases), _ptr(quals), _ptr(offsets), n, C.byref(h))
    if rc != _abi.OK:
        raise SdgpuError(rc, (lib.sdq_last_error() or b"").decode())
    t1 = time.perf_counter()
    try:
        res = _result_to_dict(lib, h, n, want_clusters)
    finally:
        t2 = time.perf_counter()
        lib.sdq_free(h)
    if os.environ.get("SDQ_TIMING"):
        print(f"[sdq] python: sdq_run {t1 - t0:.3f} fetch {t2 - t1:.3f} free {time.perf_counter() - t2:.3f}", file=sys.stderr)
    return res


def _result_to_dict(lib, h, n, want_clusters):
    st = StatsC()
    lib.sdq_get_stats(h, C.byref(st))
    res = {"stats": {k: getattr(st, k) for k, _ in StatsC._fields_}, "clusters": []}
    # every consensus sequence, its qualities and its count in one call (sdq_clusters_packed)
    nc = lib.sdq_num_clusters(h)
    tot = lib.sdq_clusters_total_len(h)
    pb, pq = np.zeros(tot, np.uint8), np.zeros(tot, np.uint8)
    po, pc = np.zeros(nc + 1, np.uint64), np.zeros(nc, np.float64)
    lib.sdq_clusters_packed(h, _ptr(pb), _ptr(pq), _ptr(po), _ptr(pc))
    res["packed"] = {"bases": pb, "quals": pq, "offsets": po, "cnts": pc}
    if want_clusters:  # per-cluster Python objects (tests); the packed arrays above are the bulk form
        blob = pb.tobytes().decode()
        for c in range(nc):
            lo, hi = int(po[c]), int(po[c + 1])
            res["clusters"].append((blob[lo:hi], pq[lo:hi].copy(), float(pc[c])))
    res["chimeras"] = {}  # cluster -> (frontParent, backParent, frontPos, backPos), chiParentsInfo.txt's columns
    if st.chimericClusters:  # (a per-cluster ctypes call: only worth it when something was flagged)
        par, pos = (C.c_uint32 * 2)(), (C.c_uint32 * 2)()
        for c in range(lib.sdq_num_clusters(h)):
            if lib.sdq_cluster_chimeric(h, c, par, pos):
                res["chimeras"][c] = (par[0], par[1], pos[0], pos[1])
    mem = np.zeros(n, np.uint32)
    lib.sdq_membership(h, _ptr(mem))
    res["membership"] = mem
    return res


def qluster_many(params, samples, device=0, nWorkers=2, pars: QlusterPars = None, iteratorMap: CollapseIterations = None,
                 intialParameters: CollapseIterations = None, want_clusters=True):
    """Many independent samples on one GPU (`sdq_run_many`): `samples` is a list of (bases, quals, offsets);
    nWorkers host threads, each with its own ctx, overlap one sample's host phases with another's GPU batches."""
    lib = _abi.load()
    _bind(lib)
    v = C.c_void_p
    lib.sdq_run_many.restype = C.c_int
    lib.sdq_run_many.argtypes = [C.c_int, C.POINTER(_abi.Params), C.POINTER(QlusterOptsC), v, C.c_uint32, v, C.c_uint32,
                                 C.c_uint32, v, v, v, v, C.c_uint32, v]
    pars = pars or QlusterPars()
    iteratorMap = iteratorMap or CollapseIterations.gen_illumina_default_pars(100)
    intialParameters = intialParameters or iteratorMap
    opts = pars.to_c()
    ip, nip = iter_pars_to_c(intialParameters)
    pp, npp = iter_pars_to_c(iteratorMap)
    keep = [(np.ascontiguousarray(b, np.uint8), np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(o, np.uint64))
            for b, q, o in samples]
    n = len(keep)
    bp = (C.c_void_p * n)(*[k[0].ctypes.data for k in keep])
    qp = (C.c_void_p * n)(*[k[1].ctypes.data for k in keep])
    op = (C.c_void_p * n)(*[k[2].ctypes.data for k in keep])
    nr = (C.c_uint32 * n)(*[len(k[2]) - 1 for k in keep])
    out = (C.c_void_p * n)()
    rc = lib.sdq_run_many(device, C.byref(params), C.byref(opts), ip, nip, pp, npp, n, bp, qp, op, nr, nWorkers, out)
    if rc != _abi.OK:
        raise SdgpuError(rc, (lib.sdq_last_error() or b"").decode())
    results = []
    for s in range(n):
        try:
            results.append(_result_to_dict(lib, out[s], nr[s], want_clusters))
        finally:
            lib.sdq_free(out[s])
    return results


def population_cluster(aligner: GpuAligner, samples, pars: QlusterPars = None, popIteratorMap: CollapseIterations = None,
                       want_clusters=True):
    """sampColl.doPopulationClustering(...) (processClusters.cpp:226): the final clusters of several samples (`samples` =
    qluster() results, or (bases, quals, offsets, cnts) tuples of packed clusters) collapsed across samples with
    popIteratorMap, k-mer checks off.  Returns a qluster()-shaped result plus "sampleOfInput" / "clusterOfInput": for
    every input cluster, in sample-major order, where it came from; result["membership"][i] is its population cluster."""
    lib = aligner._lib
    _bind(lib)
    pars = pars or QlusterPars()
    popIteratorMap = popIteratorMap or CollapseIterations.gen_illumina_default_pars(100)
    packs = []
    for smp in samples:
        if isinstance(smp, dict):
            pk = smp["packed"]
            packs.append((pk["bases"], pk["quals"], pk["offsets"], pk["cnts"]))
        else:
            packs.append(smp)
    bases = np.concatenate([np.asarray(p[0], np.uint8) for p in packs]) if packs else np.zeros(0, np.uint8)
    quals = np.concatenate([np.asarray(p[1], np.uint8) for p in packs]) if packs else np.zeros(0, np.uint8)
    cnts = np.concatenate([np.asarray(p[3], np.float64) for p in packs]) if packs else np.zeros(0, np.float64)
    lens = np.concatenate([np.diff(np.asarray(p[2], np.uint64)) for p in packs]) if packs else np.zeros(0, np.uint64)
    offsets = np.zeros(len(lens) + 1, np.uint64)
    offsets[1:] = np.cumsum(lens)
    n = len(lens)
    opts = pars.to_c()
    pp, npp = iter_pars_to_c(popIteratorMap)
    h = C.c_void_p()
    rc = lib.sdq_population_cluster_packed(aligner._ctx, C.byref(opts), pp, npp, _ptr(bases), _ptr(quals), _ptr(offsets), _ptr(cnts), n,
                                           C.byref(h))
    if rc != _abi.OK:
        raise SdgpuError(rc, (lib.sdq_last_error() or b"").decode())
    try:
        res = _result_to_dict(lib, h, n, want_clusters)
    finally:
        lib.sdq_free(h)
    res["sampleOfInput"] = np.concatenate([np.full(len(p[2]) - 1, s, np.uint32) for s, p in enumerate(packs)]) if packs else np.zeros(0, np.uint32)
    res["clusterOfInput"] = np.concatenate([np.arange(len(p[2]) - 1, dtype=np.uint32) for p in packs]) if packs else np.zeros(0, np.uint32)
    res["input"] = (bases, quals, offsets, cnts)
    return res


def shard_samples(nSamples, rank, world):
    """Static round-robin of independent samples over ranks (no data-path collective)."""
    return list(range(rank, nSamples, world))


def shard_pair_blocks(totalPairs, rank, world):
    """Contiguous block [firstPair, firstPair + nPairs) of an all-vs-all pair list for this rank
    (sdgpu_all_vs_all's firstPair / nPairs): blocks are disjoint, cover every pair once and differ
    by at most one pair; sequences are replicated, results stay on the rank that computed them
    (no data-path collective)."""
    base, extra = divmod(int(totalPairs), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def read_fastq(path):
    """sdq_read_fastq: a plain or gzip FASTQ file as packed (bases, quals, offsets, names) -- what qluster() takes."""
    lib = _abi.load()
    v = C.c_void_p
    lib.sdq_read_fastq.restype = C.c_int
    lib.sdq_read_fastq.argtypes = [C.c_char_p, C.POINTER(v)]
    lib.sdq_reads_free.argtypes = [v]
    lib.sdq_reads_count.restype = C.c_uint32
    lib.sdq_reads_count.argtypes = [v]
    for f in ("bases", "quals", "offsets"):
        getattr(lib, "sdq_reads_" + f).restype = v
        getattr(lib, "sdq_reads_" + f).argtypes = [v]
    lib.sdq_reads_name.restype = C.c_char_p
    lib.sdq_reads_name.argtypes = [v, C.c_uint32]
    lib.sdq_fastq_last_error.restype = C.c_char_p
    h = v()
    rc = lib.sdq_read_fastq(str(path).encode(), C.byref(h))
    if rc != _abi.OK:
        raise SdgpuError(rc, (lib.sdq_fastq_last_error() or b"").decode())
    try:
        n = lib.sdq_reads_count(h)
        offsets = np.ctypeslib.as_array(C.cast(lib.sdq_reads_offsets(h), C.POINTER(C.c_uint64)), (n + 1,)).copy()
        tot = int(offsets[-1])
        if tot:
            bases = np.ctypeslib.as_array(C.cast(lib.sdq_reads_bases(h), C.POINTER(C.c_uint8)), (tot,)).copy()
            quals = np.ctypeslib.as_array(C.cast(lib.sdq_reads_quals(h), C.POINTER(C.c_uint8)), (tot,)).copy()
        else:
            bases, quals = np.zeros(0, np.uint8), np.zeros(0, np.uint8)
        names = [lib.sdq_reads_name(h, i).decode() for i in range(n)]
    finally:
        lib.sdq_reads_free(h)
    return bases, quals, offsets, names


def qluster_fastq(aligner: GpuAligner, path, pars: QlusterPars = None, iteratorMap: CollapseIterations = None,
                  intialParameters: CollapseIterations = None, want_clusters=True):
    """`SeekDeep qluster --fastq path`: sdq_run_fastq = sdq_read_fastq + sdq_run, everything behind the C ABI."""
    lib = aligner._lib
    _bind(lib)
    v = C.c_void_p
    lib.sdq_run_fastq.restype = C.c_int
    lib.sdq_run_fastq.argtypes = [v, C.POINTER(QlusterOptsC), v, C.c_uint32, v, C.c_uint32, C.c_char_p, C.POINTER(v)]
    lib.sdq_fastq_last_error.restype = C.c_char_p
    pars = pars or QlusterPars()
    iteratorMap = iteratorMap or CollapseIterations.gen_illumina_default_pars(100)
    intialParameters = intialParameters or iteratorMap
    opts = pars.to_c()
    ip, nip = iter_pars_to_c(intialParameters)
    pp, npp = iter_pars_to_c(iteratorMap)
    h = v()
    rc = lib.sdq_run_fastq(aligner._ctx, C.byref(opts), ip, nip, pp, npp, str(path).encode(), C.byref(h))
    if rc != _abi.OK:
        raise SdgpuError(rc, (lib.sdq_fastq_last_error() or lib.sdq_last_error() or b"").decode())
    try:
        st = StatsC()
        lib.sdq_get_stats(h, C.byref(st))
        return _result_to_dict(lib, h, int(st.inputReads), want_clusters)
    finally:
        lib.sdq_free(h)
