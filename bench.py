#!/usr/bin/env python
"""bench.py — headline benchmark of the qluster hot path on B200 (driver contract in the task).

Workload (default, `--workload c2`) = BASELINE.json configs[1]: one synthetic sample,
200k reads x 300 bp from 200 haplotypes at 1 % substitutions + homopolymer indels, qluster with the
Illumina iteration table (etc/SeekDeepParameterFiles/illumina_lkmer1 == genIlluminaDefaultPars(100)),
qualities 25/20, gaps 5,1 / left 5,1 / right 0,0, k-mer checks on (k = 9, by position).

A "step" is one full qluster collapse of one sample per GPU: dedupe -> k-mer index -> three
runFullClustering passes (k-mer table, screen pass / affine-gap alignment + traceback, errorProfile,
par-file verdicts on the GPU; the sequential merge/consensus logic on the host, in C++).

  value : input reads / s, timed with CUDA events on the library's stream from the moment the
          sample's unique reads are resident in HBM (sdq_stats.msResident)
  e2e   : input reads / s through the C-ABI call with host buffers (sdq_run wall time: dedupe, uploads,
          index lists H2D, hit records / gap lists D2H, and the final clusters -- consensus sequences,
          qualities, counts, membership -- fetched, all inside)
  gcups : DP cell updates / s (second half of BASELINE.json's metric) over the same step
Multi-GPU: samples shard over ranks with no data-path collective (one sample per GPU per step,
weak scaling); only the timings are reduced (MAX) for reporting.

`--workload c5` runs BASELINE configs[4] instead (384 samples x 50k reads x 250 bp sharded over the ranks,
several worker contexts per GPU through sdq_run_many); `--impl reference` times the CPU restatement of the
reference (the oracle port; the SeekDeep binary cannot be built here) on a bounded sample of the C2 generator.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_CELL = 13  # algorithmic INT32 ops of one 3-state affine cell update
# The CPU arms run the oracle on samples of the SAME generator shrunk to this many reads (the 200k-read sample
# would take hours per core); the GPU arm runs the identical small samples too (`same_input`).
CPU_SAMPLE_READS = int(os.environ.get("SD_BENCH_CPU_READS", 200))
CPU_SAMPLE_SEED0 = 5000
# DRAM read + write bytes per launch of the dominant kernel (screenDpKernel<8>), from the committed
# `ncu --set full` capture profiles/r2_screen_dp_kernel_ncu_full.csv (dram__bytes_read.sum + dram__bytes_write.sum of
# one launch of the bench at full size).  The kernel writes no traceback bits: its
# algorithmic HBM bytes are the two sequences and the 16-byte work item of each pair.
DP_KERNEL_DRAM_BYTES_PER_LAUNCH = 22_322_944  # 22.32 MB read + 768 B written, one screenDpKernel<8> launch of 6.53 ms (~640 k pairs)
DP_KERNEL_DRAM_SOURCE = "profiles/r2_screen_dp_kernel_ncu_full.csv"


def make_sample(seed, nReads):
    from tests import synth
    haps, seqs, quals, _ = synth.config2_fast(seed=seed, nReads=nReads)
    return synth.pack(seqs, quals)


def make_c5_sample(seed, nReads=50_000):
    """One sample of BASELINE configs[4]: the C1 generator (20 haplotypes x 250 bp, 0.5 % substitutions) at 50k reads."""
    from tests import synth
    rng = np.random.default_rng(seed)
    haps = synth.make_haplotypes(rng, 20, 250, 1, 6)
    seqs, quals, _ = synth.fast_reads(rng, haps, nReads, 0.005)
    return synth.pack(seqs, quals)


def workload_config(nReads):
    """Identical in both arms.  `reference_arm_sample` says what the CPU arm can afford of this workload."""
    return {"workload": f"SeekDeep qluster single sample, {nReads} reads x 300 bp from 200 haplotypes at 1% error with "
                        "homopolymer indels, Illumina par table (illumina_lkmer1, 22 iterations x 3 passes) — BASELINE configs[1]",
            "reads_per_sample": nReads, "samples_per_gpu_per_step": 1, "stopCheck": 100, "gap": "5,1 left 5,1 right 0,0",
            "qualThres": "25,20", "kLength": 9,
            "per_rank_sample": "identical copy of the seed-2 sample on every rank (fixed per-GPU work)",
            "reference_arm_sample": f"the CPU arm (--impl reference, cpu_baseline) runs the same generator shrunk to {CPU_SAMPLE_READS} reads "
                                    "per sample, one single-threaded qluster per host core; the GPU arm reports the identical small samples "
                                    "under same_input",
            "l2": "no flush: a step streams 107 MB of sequences and flags through ~60 grid launches whose pair lists (20.8 M pairs) "
                  "are generated on the device; inputs are far larger than the 126 MB L2"}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_qluster(seed, want_result=False):
    """One single-threaded oracle qluster (the CPU restatement of the reference) on a bounded sample."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations
    from tests import oracle_binding
    o = oracle_binding.load()
    bases, quals, offsets = make_sample(seed, CPU_SAMPLE_READS)
    its = CollapseIterations.gen_illumina_default_pars(100).iters
    t0 = time.perf_counter()
    r = o.qluster(sd.make_params(), sd.QlusterPars().to_c(), its, its, bases, quals, offsets)
    dt = time.perf_counter() - t0
    if want_result:
        return dt, CPU_SAMPLE_READS, r["nCells"], r
    return dt, CPU_SAMPLE_READS, r["nCells"]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  The SeekDeep binary cannot be built here
    (njhseq un-vendored), so this is the oracle port, parallelised exactly like the reference: one single-threaded
    qluster per sample, as many concurrent samples as host cores (runMultipleCommands --numThreads,
    SeekDeepUtilsRunner.cpp:294-356).  Each step = one bounded sample per core (a few seconds)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(cpu_qluster, [CPU_SAMPLE_SEED0 + (step * cores + c) % 64 for c in range(cores)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append((dt, sum(r[1] for r in res), sum(r[2] for r in res)))
    tot = sum(t[0] for t in times)
    reads = sum(t[1] for t in times)
    cells = sum(t[2] for t in times)
    v = reads / tot
    line = {"impl": "reference", "metric": "qluster_reads_per_s", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, len(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "gcups": cells / tot / 1e9,
            "config": workload_config(args.reads),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": f"per step: {cores} concurrent single-threaded oracle qlusters (one per host core), each on a "
                                       f"{CPU_SAMPLE_READS}-read sample of the C2 generator (seeds {CPU_SAMPLE_SEED0}+); NOT the 200k-read sample, "
                                       "which would take hours per core"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def same_input_block(sd, local):
    """The reference arm's own small samples on the GPU (like-for-like input): reads/s through sdq_run_many with
    host buffers, and the clusters of the first samples checked against the oracle's in this very run."""
    from seekdeep_b200.qluster import qluster_many
    n = 16
    samples = [make_sample(CPU_SAMPLE_SEED0 + c, CPU_SAMPLE_READS) for c in range(n)]
    qluster_many(sd.make_params(), samples[:4], device=local, nWorkers=4, want_clusters=False)  # contexts warm
    t0 = time.perf_counter()
    rs = qluster_many(sd.make_params(), samples, device=local, nWorkers=4, want_clusters=False)
    dt = time.perf_counter() - t0
    checked, cpu_s = 0, 0.0
    for c in range(2):
        cdt, _, _, want = cpu_qluster(CPU_SAMPLE_SEED0 + c, want_result=True)
        cpu_s += cdt
        got = rs[c]
        wseq = "".join(s for s, _, _ in want["clusters"])
        ok = (np.array_equal(got["membership"], want["membership"]) and got["packed"]["bases"].tobytes().decode() == wseq and
              np.array_equal(got["packed"]["cnts"], np.array([x for _, _, x in want["clusters"]])))
        if not ok:
            raise RuntimeError(f"same_input: GPU clusters of sample {c} differ from the oracle's")
        checked += 1
    return {"samples": n, "reads_per_sample": CPU_SAMPLE_READS, "value": n * CPU_SAMPLE_READS / dt, "unit": "reads/s", "workers": 4,
            "ms": dt * 1e3, "oracle_checked_samples": checked, "oracle_reads_per_s_one_core": checked * CPU_SAMPLE_READS / cpu_s,
            "note": "identical inputs to --impl reference's samples; such small samples are launch- and host-bound on a GPU "
                    "(59 launches per sample), the number is what the reference arm's ratio can honestly be compared with"}


def kmer_roofline(al, bases, quals, offsets, peaks):
    """Secondary roofline (HBM-bound prefilter kernel): dense k = 5 profile build over the sample."""
    lens = np.diff(offsets).astype(np.int64)
    al.upload_seqs(bases, quals, offsets)
    al.build_kmer_profiles(5)
    al.sync()
    best = 1e9
    for _ in range(5):
        al.timer_start()
        al.build_kmer_profiles(5)
        best = min(best, al.timer_stop())
    algo = float(lens.sum() + len(lens) * (16 + 1024 * 2))  # codes + offsets in, uint16[4^5] profile out, per read
    gbs = algo / (best * 1e-3) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {"bound": "hbm", "kernel": "kmerProfileKernel4(k=5)", "achieved": gbs, "peak": peak, "unit": "GB/s",
           "frac": gbs / peak, "traffic": None, "ms": best, "algorithmic_bytes": algo,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
    # reads x clusters compareKmers grid (every read vs the 100 first sequences), device-resident
    nR, nC = len(lens), min(100, len(lens))
    dR, dC = al.dev_alloc(nR * 4), al.dev_alloc(nC * 4)
    dS, dF = al.dev_alloc(nR * nC * 4), al.dev_alloc(nR * nC * 8)
    al.h2d(dR, np.arange(nR, dtype=np.uint32))
    al.h2d(dC, np.arange(nC, dtype=np.uint32))
    bestc = 1e9
    for _ in range(4):
        al.timer_start()
        al._check(al._lib.sdgpu_kmer_compare_all_dev(al._ctx, dR, nR, dC, nC, dS, dF))
        bestc = min(bestc, al.timer_stop())
    for p in (dR, dC, dS, dF):
        al.dev_free(p)
    algoc = float((nR + nC) * 1024 * 2 + nR * nC * 12)  # (R + C) * P + R * C * 12 B
    out["compare"] = {"kernel": "kmerCompareGridKernel(reads x 100 clusters)", "ms": bestc, "pairs_per_s": nR * nC / (bestc * 1e-3),
                      "achieved": algoc / (bestc * 1e-3) / 1e9, "unit": "GB/s", "frac": algoc / (bestc * 1e-3) / 1e9 / peak,
                      "algorithmic_bytes": algoc,
                      "note": "cluster tile in shared memory, each read profile fetched from HBM once: HBM traffic is the compulsory "
                              "bytes, the kernel is bound by min/add SIMD on the ALU pipe + LDS, not by HBM"}
    # a1: indexKmers over the whole sample (k = 9, by position): key generation + radix sort (CUB) + run-length counting + the
    # per-base low-frequency flags; bytes = codes in, one 8-byte key + 4-byte weight per k-mer written and read back by the sort
    al.index_kmers(9, 1, True)
    al.sync()
    besti = 1e9
    for _ in range(3):
        al.timer_start()
        al.index_kmers(9, 1, True)
        besti = min(besti, al.timer_stop())
    nk = float(np.maximum(lens - 8, 0).sum())
    out["index_build"] = {"kernel": "kmerKeysKernel + cub::DeviceRadixSort + ReduceByKey + kmerPosStartKernel + lowKmerFlagKernel (k=9, by position)",
                          "ms": besti, "kmers_per_s": nk / (besti * 1e-3), "algorithmic_bytes": float(lens.sum()) * 2 + nk * 12 * 2,
                          "achieved": (float(lens.sum()) * 2 + nk * 12 * 2) / (besti * 1e-3) / 1e9, "unit": "GB/s",
                          "note": "the 8-pass radix sort moves each 12-byte record 16 times; it stays a library sort: one build is "
                                  "12 ms of a 300 ms step and runs twice per sample"}
    # qluster --useKmerBinning on the same sample: sorted k-mer lists (k = 10) of every read, then one round of the binning
    # comparison (64 candidate seeds x every read, 8 bytes per read back)
    al.build_kmer_lists(10)
    al.sync()
    bestl = 1e9
    for _ in range(3):
        al.timer_start()
        al.build_kmer_lists(10)
        bestl = min(bestl, al.timer_stop())
    algol = float(lens.sum() + len(lens) * 16 + np.maximum(lens - 9, 0).sum() * 8)  # codes + offsets in, one uint64 key per k-mer out
    seeds = np.arange(min(64, len(lens)), dtype=np.uint32)
    reads = np.arange(len(lens), dtype=np.uint32)
    al.kmer_bin_masks(seeds, reads, 0.8)
    t0 = time.perf_counter()
    al.kmer_bin_masks(seeds, reads, 0.8)
    dtm = time.perf_counter() - t0
    out["binning"] = {"kernel": "kmerListKernel(k=10) + kmerBinKernel(64 seeds x reads)", "list_ms": bestl,
                      "list_achieved": algol / (bestl * 1e-3) / 1e9, "unit": "GB/s", "list_bound": "shared-memory sort network, not HBM",
                      "list_algorithmic_bytes": algol, "masks_ms_host_call": dtm * 1e3, "mask_pairs_per_s": len(seeds) * len(reads) / dtm,
                      "note": "lists: per-sequence bitonic sort in shared memory, HBM traffic = codes in + 8 B per k-mer out; masks: one "
                              "binary search per read k-mer against seed lists staged in shared memory (LDS-bound), timed through the host "
                              "call (index lists up, 8 bytes per read back)"}
    return out


def bind_cpus(local, world):
    """One rank per GPU: keep each rank's host threads (collapse loop, prep / consensus workers) on its own
    slice of the cores NVML reports as local to its GPU, so ranks do not migrate across NUMA nodes or pile
    onto the same cores."""
    try:
        import pynvml
        pynvml.nvmlInit()
        allowed = sorted(os.sched_getaffinity(0))
        words = (max(allowed) + 64) // 64
        sets = []
        for d in range(world):
            m = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(d), words)
            cpus = [64 * w + b for w, x in enumerate(m) for b in range(64) if (x >> b) & 1]
            sets.append(tuple(c for c in cpus if c in allowed) or tuple(allowed))
        mine = sets[local]
        peers = [d for d in range(world) if sets[d] == mine]
        per = max(1, len(mine) // len(peers))
        k = peers.index(local)
        os.sched_setaffinity(0, set(mine[k * per:(k + 1) * per]))
    except Exception as e:  # binding is an optimisation, never a requirement
        print(f"[bench] cpu binding skipped: {e}", file=sys.stderr)


def run_c5(args, rank, world, local, torch, dist, sd):
    """BASELINE configs[4]: 384 samples x 50k reads x 250 bp, sharded over the ranks by sample (static round-robin,
    SeekDeepUtilsRunner.cpp:294-356 / setUpTarAmpAnalysis.cpp:998-1001), several worker contexts per GPU."""
    from seekdeep_b200.qluster import qluster_many, shard_samples
    nSamples = int(os.environ.get("SD_BENCH_C5_SAMPLES", 384))
    distinct = [make_c5_sample(500 + s) for s in range(4)]  # four seeded samples, replicated (generating 384 takes minutes)
    mine = [distinct[s % len(distinct)] for s in shard_samples(nSamples, rank, world)]
    workers = max(1, min(8, host_cores()))
    params = sd.make_params()
    reads_mine = sum(len(s[2]) - 1 for s in mine)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        qluster_many(params, mine[:max(workers, 1)], device=local, nWorkers=workers, want_clusters=False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    wall = 0.0
    cells = launches = 0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        rs = qluster_many(params, mine, device=local, nWorkers=workers, want_clusters=False)
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
        cells += sum(r["stats"]["cellUpdates"] for r in rs)
        launches += sum(r["stats"]["gpuBatches"] for r in rs)
    barrier()
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([wall * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(reads_mine), float(cells), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        K = args.steps
        ms = t.item()
        v = tot[0].item() * K / (ms * 1e-3)
        line = {"metric": "qluster_reads_per_s", "value": v, "unit": "reads/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": f"Scaling sweep: {nSamples}-sample batch, 50k reads x 250 bp each (C1 generator), qluster — BASELINE configs[4]",
                           "samples": nSamples, "reads_per_sample": 50000, "worker_contexts_per_gpu": workers,
                           "distinct_samples": len(distinct), "timing": "host wall clock around sdq_run_many (many concurrent contexts per GPU: "
                                                                        "no single stream to put CUDA events on), max over ranks",
                           "l2": "inputs of one step (hundreds of MB per rank) exceed L2"},
                "gcups": tot[1].item() / (ms * 1e-3) / 1e9,
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                        "note": "value IS end to end here: host buffers in, clusters out"},
                "gpu_launches": int(tot[2].item()), "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sdgpu")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("SD_BENCH_READS", 200000)), help="reads per sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (cpu_baseline, same_input's oracle check)")
    ap.add_argument("--no-multi", action="store_true", help="skip the many-samples-per-GPU throughput measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import seekdeep_b200 as sd

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        bind_cpus(local, world)
    if args.workload == "c5":
        run_c5(args, rank, world, local, torch, dist, sd)
        if world > 1:
            dist.destroy_process_group()
        return
    # one sample per GPU per step, sharded by sample; every rank gets its own copy of the SAME sample so that
    # per-GPU work is exactly fixed as N grows (weak scaling) and the step time is not set by whichever rank drew
    # the heaviest random sample
    bases, quals, offsets = make_sample(2, args.reads)
    nReads = len(offsets) - 1
    al = sd.GpuAligner(sd.make_params(), device=local)
    al.profile_enable(True)

    def barrier():
        al.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        sd.qluster(al, bases, quals, offsets, want_clusters=False)
    barrier()
    al.profile_read(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    wall = resident_ms = 0.0
    stats = None
    d2h_result = 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res = sd.qluster(al, bases, quals, offsets, want_clusters=False)  # (packed clusters + membership are fetched)
        wall += time.perf_counter() - t0
        stats = res["stats"]
        resident_ms += stats["msResident"]
        d2h_result = int(res["membership"].nbytes + sum(a.nbytes for a in res["packed"].values()))
    barrier()
    sampler.stop_flag = True
    sampler.join()
    prof = al.profile_read(reset=True)

    t = torch.tensor([resident_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(nReads), float(prof["cellUpdates"]), float(prof["alignKernelLaunches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    res_ms_max, wall_ms_max = t.tolist()
    readsAll, cellsAll, launchesAll = tot.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        rates = [al.measure_int32_peak(k) for k in range(6)]
        names = ["add", "max", "fused add-max", "max3", "max+imad (two pipes)", "add-max+imad (two pipes)"]
        int32_peak = max(rates)  # the best rate of any mix, single-pipe or two-pipe
        peak_gcups = int32_peak / OPS_PER_CELL / 1e9
        K = args.steps
        # dominant kernel: the score-only banded DP of the screen pass (screenDpKernel<2|4|8>)
        dp_ms = prof["screenKernelMs"] - prof["screenScanMs"]
        dp_gcups = prof["screenCells"] / max(1e-9, dp_ms * 1e-3) / 1e9
        # all alignment kernels together: cells really evaluated (screen bands, banded pass, full matrices) per kernel second;
        # the BASELINE GCUPS metric (sum of lenA x lenB over aligned pairs / kernel time) is "effective"
        all_gcups = prof["evaluatedCells"] / (prof["alignKernelMs"] * 1e-3) / 1e9
        effective_gcups = prof["cellUpdates"] / (prof["alignKernelMs"] * 1e-3) / 1e9
        full_ms = prof["alignKernelMs"] - prof["bandKernelMs"] - prof["screenKernelMs"]
        full_cells = prof["evaluatedCells"] - prof["bandCells"] - prof["screenCells"]
        line = {
            "metric": "qluster_reads_per_s", "value": readsAll * K / (res_ms_max * 1e-3), "unit": "reads/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": res_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": dict(workload_config(nReads)),
            "run": {"unique_seqs": stats["uniqueSeqs"], "final_clusters": stats["finalClusters"],
                    "pairs_computed_per_step": stats["pairsComputed"], "passing_pairs_per_step": stats["passingPairs"],
                    "consensus_aligns_per_step": stats["consensusAligns"], "chimera_probes_per_step": stats["chimeraPairs"],
                    "chimeric_clusters": stats["chimericClusters"]},
            "gcups": cellsAll / (res_ms_max * 1e-3) / 1e9,
            "e2e": {"value": readsAll * K / (wall_ms_max * 1e-3), "unit": "reads/s",
                    "h2d_bytes_per_step": int(prof["h2dBytes"] // K), "d2h_bytes_per_step": int(prof["d2hBytes"] // K) + d2h_result,
                    "gcups": cellsAll / (wall_ms_max * 1e-3) / 1e9, "ms_per_step": wall_ms_max / K,
                    "result_fetched": "membership + packed consensus sequences, qualities, counts (sdq_clusters_packed)"},
            "gpu_launches": int(prof["kernelLaunches"]),
            "clocks": sampler.summary(),
            "roofline": {"bound": "int32", "kernel": "screenDpKernel<2|4|8> (score-only banded DP of the screen pass)",
                         "achieved": dp_gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": dp_gcups / peak_gcups,
                         "traffic": DP_KERNEL_DRAM_BYTES_PER_LAUNCH, "traffic_source": DP_KERNEL_DRAM_SOURCE,
                         "ops_per_cell": OPS_PER_CELL, "kernel_ms_per_step": dp_ms / K, "cells_per_step": prof["screenCells"] / K,
                         "kernel_share_of_step": dp_ms / resident_ms,
                         "peak_source": f"measured on this GPU in this run: best of six INT32 microbenchmarks = {int32_peak / 1e12:.2f} T thread-instr/s ("
                                        + ", ".join(f"{n} {r / 1e12:.2f}" for n, r in zip(names, rates)) + f") / {OPS_PER_CELL} algorithmic ops per cell update",
                         "single_alu_pipe_peak_gcups": rates[2] / OPS_PER_CELL / 1e9,
                         "all_alignment_kernels": {
                             "achieved": all_gcups, "frac": all_gcups / peak_gcups, "effective_gcups": effective_gcups,
                             "kernel_ms_per_step": prof["alignKernelMs"] / K, "kernel_share_of_step": prof["alignKernelMs"] / resident_ms,
                             "launches": int(prof["alignKernelLaunches"]),
                             "evaluated_over_effective_cells": prof["evaluatedCells"] / max(1, prof["cellUpdates"]),
                             "screen_pass": {"pairs_per_step": prof["screenPairs"] / K,
                                             "share_of_pairs_finished": prof["screenCertified"] / max(1, prof["screenPairs"]),
                                             "scan_kernel_ms_per_step": prof["screenScanMs"] / K, "dp_kernel_ms_per_step": dp_ms / K},
                             "banded_pass": {"share_of_pair_cells_certified": prof["bandCertifiedCells"] / max(1, prof["cellUpdates"]),
                                             "band_kernel_ms_per_step": prof["bandKernelMs"] / K,
                                             "band_kernel_gcups": prof["bandCells"] / max(1e-9, prof["bandKernelMs"] * 1e-3) / 1e9},
                             "full_matrix": {"kernel_ms_per_step": full_ms / K,
                                             "gcups": full_cells / max(1e-9, full_ms * 1e-3) / 1e9}}},
        }
        try:
            line["roofline_kmer"] = kmer_roofline(al, bases, quals, offsets, peaks)
        except Exception as e:  # the secondary number must not hide the headline
            line["roofline_kmer"] = {"error": str(e)}
        if world == 1 and not args.no_multi:
            # throughput mode for many-sample runs (configs 4/5): 3 worker contexts share this GPU so one
            # sample's host-only phases overlap another's alignment batches (sdq_run_many)
            try:
                from seekdeep_b200.qluster import qluster_many
                samples = [(bases, quals, offsets)] + [make_sample(50 + s, args.reads) for s in range(3)]
                t0 = time.perf_counter()
                rs = qluster_many(sd.make_params(), samples, device=local, nWorkers=3, want_clusters=False)
                dtm = time.perf_counter() - t0
                line["multi_sample"] = {"samples": len(samples), "workers": 3, "value": sum(len(s[2]) - 1 for s in samples) / dtm,
                                        "unit": "reads/s", "gcups": sum(r["stats"]["cellUpdates"] for r in rs) / dtm / 1e9,
                                        "note": "wall time of one sdq_run_many call, worker contexts created inside it"}
            except Exception as e:
                line["multi_sample"] = {"error": str(e)}
        if world == 1 and not args.no_cpu:
            dt, n, c = cpu_qluster(CPU_SAMPLE_SEED0)
            line["cpu_baseline"] = {"value": n / dt, "unit": "reads/s", "cores": 1, "kind": "port", "gcups": c / dt / 1e9,
                                    "sample": f"one single-threaded oracle qluster on a {n}-read sample of the same generator ({dt:.1f} s)"}
            try:
                line["same_input"] = same_input_block(sd, local)
            except Exception as e:
                line["same_input"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    al.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
