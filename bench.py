#!/usr/bin/env python
"""bench.py — headline benchmark of the qluster hot path on B200 (driver contract in the task).

Workload = BASELINE.json configs[1]: one synthetic sample, 200k reads x 300 bp
from 200 haplotypes at 1 % substitutions + homopolymer indels, qluster with the Illumina
iteration table (etc/SeekDeepParameterFiles/illumina_lkmer1 == genIlluminaDefaultPars(100)),
qualities 25/20, gaps 5,1 / left 5,1 / right 0,0, k-mer checks on (k = 9, by position).

A "step" is one full qluster collapse of one sample per GPU: dedupe -> k-mer index -> three
runFullClustering passes (k-mer prefilter table, affine-gap alignment + traceback, errorProfile,
par-file verdicts on the GPU; the sequential merge/consensus logic on the host, in C++).

  value : input reads / s, timed with CUDA events on the library's stream from the moment the
          sample's unique reads are resident in HBM (sdq_stats.msResident)
  e2e   : input reads / s through the C-ABI call with host buffers (sdq_run wall time: dedupe,
          uploads, pair lists H2D, verdict masks / gap lists D2H all inside)
  gcups : DP cell updates / s (second half of BASELINE.json's metric) over the same step
Multi-GPU: samples shard over ranks with no data-path collective (one sample per GPU per step,
weak scaling); only the timings are reduced (MAX) for reporting.
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
CPU_SAMPLE_READS = int(os.environ.get("SD_BENCH_CPU_READS", 600))  # reads of the bounded CPU sample (tests shrink it)
# DRAM bytes per evaluated cell of the dominant kernel (alignBandKernel): its only bulk traffic is the
# decision-bit slab, 128 B per 6 anti-diagonals x 32 lanes = 12.8 KB for the ~19.2 k band cells of a 300 x 300
# pair; the ncu --set full capture profiles/r1m_band_kernel_ncu_full.csv shows 3.37 GB written + 0.04 GB read
# for a launch whose slabs total 3.4 GB (the full-matrix strip kernel measured 1.40 B/cell, profiles/r1b_*)
TRAFFIC_BYTES_PER_CELL = 0.67


def make_sample(seed, nReads):
    from seekdeep_b200 import synth
    haps, seqs, quals, _ = synth.config2_fast(seed=seed, nReads=nReads)
    return synth.pack(seqs, quals)


def workload_config(nReads):
    return {"workload": f"SeekDeep qluster single sample, {nReads} reads x 300 bp from 200 haplotypes at 1% error with "
                        "homopolymer indels, Illumina par table (illumina_lkmer1, 22 iterations x 3 passes) — BASELINE configs[1]",
            "reads_per_sample": nReads, "samples_per_gpu_per_step": 1, "stopCheck": 100, "gap": "5,1 left 5,1 right 0,0",
            "qualThres": "25,20", "kLength": 9,
            "per_rank_sample": "identical copy of the seed-2 sample on every rank (fixed per-GPU work)",
            "l2": "no flush: a step streams 107 MB of sequences plus 330 MB of pair lists and verdict masks through ~70 "
                  "alignment launches and rewrites the 60 MB of per-warp decision-bit slabs in each, far beyond the 126 MB L2"}


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


def cpu_qluster(seed):
    """One single-threaded oracle qluster (the CPU restatement of the reference) on a bounded sample."""
    import seekdeep_b200 as sd
    from seekdeep_b200.iterpars import CollapseIterations
    from tests import oracle_binding
    from tests.test_gpu_qluster import oracle_opts
    o = oracle_binding.load()
    bases, quals, offsets = make_sample(seed, CPU_SAMPLE_READS)
    its = CollapseIterations.gen_illumina_default_pars(100).iters
    t0 = time.perf_counter()
    r = o.qluster(sd.make_params(), oracle_opts(sd.QlusterPars()), its, its, bases, quals, offsets)
    dt = time.perf_counter() - t0
    return dt, CPU_SAMPLE_READS, r["nCells"]


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  The SeekDeep binary
    cannot be built here (njhseq un-vendored), so this is the oracle port, parallelised exactly
    like the reference: one single-threaded qluster per sample, as many concurrent samples as
    host cores (runMultipleCommands --numThreads, SeekDeepUtilsRunner.cpp:294-356)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(cpu_qluster, [5000 + step * cores + c for c in range(cores)])
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
                             "sample": f"per step: {cores} concurrent single-threaded oracle qlusters, each on a {CPU_SAMPLE_READS}-read "
                                       "sample of the same generator (the full 200k-read sample would take hours per core)"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
    out = {"bound": "hbm", "kernel": "kmerProfileKernel(k=5)", "achieved": gbs, "peak": peak, "unit": "GB/s",
           "frac": gbs / peak, "traffic": None, "ms": best, "algorithmic_bytes": algo,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
    # reads x clusters compareKmers grid (every read vs the 100 first sequences), device-resident
    import ctypes as C
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
    out["compare"] = {"kernel": "kmerCompareKernel(reads x 100 clusters)", "ms": bestc, "pairs_per_s": nR * nC / (bestc * 1e-3),
                      "achieved": algoc / (bestc * 1e-3) / 1e9, "unit": "GB/s", "frac": algoc / (bestc * 1e-3) / 1e9 / peak,
                      "algorithmic_bytes": algoc,
                      "note": "cluster tile in shared memory, each read profile fetched from HBM once: HBM traffic is the compulsory "
                              "bytes, the kernel is bound by min/add SIMD on the ALU pipe + LDS, not by HBM"}
    return out


def bind_cpus(local, world):
    """One rank per GPU: keep each rank's host threads (collapse loop, mask filer, prep workers) on its own
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sdgpu")
    ap.add_argument("--reads", type=int, default=int(os.environ.get("SD_BENCH_READS", 200000)), help="reads per sample")
    ap.add_argument("--no-cpu", action="store_true")
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
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res = sd.qluster(al, bases, quals, offsets, want_clusters=False)
        wall += time.perf_counter() - t0
        stats = res["stats"]
        resident_ms += stats["msResident"]
        d2h_result = int(res["membership"].nbytes)
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
        rates = [al.measure_int32_peak(k) for k in range(4)]
        int32_peak = max(rates)
        peak_gcups = int32_peak / OPS_PER_CELL / 1e9
        # roofline numerator = cell updates the kernels really evaluated (band cells of the banded first pass,
        # whole matrices of the pairs redone in full); the GCUPS metric (sum of lenA x lenB over aligned
        # pairs / kernel time) is reported beside it as "effective"
        kernel_gcups = prof["evaluatedCells"] / (prof["alignKernelMs"] * 1e-3) / 1e9
        effective_gcups = prof["cellUpdates"] / (prof["alignKernelMs"] * 1e-3) / 1e9
        full_ms = prof["alignKernelMs"] - prof["bandKernelMs"] - prof["screenKernelMs"]
        full_cells = prof["evaluatedCells"] - prof["bandCells"] - prof["screenCells"]
        K = args.steps
        line = {
            "metric": "qluster_reads_per_s", "value": readsAll * K / (res_ms_max * 1e-3), "unit": "reads/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": res_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": dict(workload_config(nReads), unique_seqs=stats["uniqueSeqs"], final_clusters=stats["finalClusters"],
                           pairs_computed_per_step=stats["pairsComputed"], consensus_aligns_per_step=stats["consensusAligns"],
                           chimera_probes_per_step=stats["chimeraPairs"], chimeric_clusters=stats["chimericClusters"]),
            "gcups": cellsAll / (res_ms_max * 1e-3) / 1e9,
            "e2e": {"value": readsAll * K / (wall_ms_max * 1e-3), "unit": "reads/s",
                    "h2d_bytes_per_step": int(prof["h2dBytes"] // K), "d2h_bytes_per_step": int(prof["d2hBytes"] // K) + d2h_result,
                    "gcups": cellsAll / (wall_ms_max * 1e-3) / 1e9, "ms_per_step": wall_ms_max / K},
            "gpu_launches": int(prof["kernelLaunches"]),
            "clocks": sampler.summary(),
            "roofline": {"bound": "int32", "kernel": "alignBandKernel (+ alignProfileKernel redo)", "achieved": kernel_gcups, "peak": peak_gcups,
                         "unit": "GCUPS", "frac": kernel_gcups / peak_gcups,
                         # dram read+write bytes per launch: bytes per evaluated cell (see TRAFFIC_BYTES_PER_CELL) x cells per launch
                         "traffic": TRAFFIC_BYTES_PER_CELL * prof["evaluatedCells"] / max(1, prof["alignKernelLaunches"]),
                         "effective_gcups": effective_gcups,
                         "banded_pass": {"share_of_pair_cells_certified": prof["bandCertifiedCells"] / max(1, prof["cellUpdates"]),
                                         "evaluated_over_effective_cells": prof["evaluatedCells"] / max(1, prof["cellUpdates"]),
                                         "band_kernel_ms_per_step": prof["bandKernelMs"] / K,
                                         "band_kernel_gcups": prof["bandCells"] / max(1e-9, prof["bandKernelMs"] * 1e-3) / 1e9,
                                         "full_kernel_ms_per_step": full_ms / K,
                                         "full_kernel_gcups": full_cells / max(1e-9, full_ms * 1e-3) / 1e9},
                         "screen_pass": {"pairs_per_step": prof["screenPairs"] / K, "share_of_pairs_finished": prof["screenCertified"] / max(1, prof["screenPairs"]),
                                         "kernel_ms_per_step": prof["screenKernelMs"] / K,
                                         "gcups": prof["screenCells"] / max(1e-9, prof["screenKernelMs"] * 1e-3) / 1e9,
                                         "cells_per_step": prof["screenCells"] / K},
                         "algorithmic_hbm_bytes_per_cell": 0.67,
                         "launches": int(prof["alignKernelLaunches"]), "avg_launch_ms": prof["alignKernelMs"] / max(1, prof["alignKernelLaunches"]),
                         "kernel_share_of_step": prof["alignKernelMs"] / resident_ms,
                         "peak_source": f"measured on this GPU in this run: {int32_peak / 1e12:.2f} T INT32 thread-instr/s "
                                        f"(add / max / fused add-max / max3 = {', '.join(f'{r / 1e12:.2f}' for r in rates)}) / "
                                        f"{OPS_PER_CELL} algorithmic ops per cell update",
                         "single_alu_pipe_peak_gcups": rates[2] / OPS_PER_CELL / 1e9},
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
        if not args.no_cpu:
            dt, n, c = cpu_qluster(2)
            line["cpu_baseline"] = {"value": n / dt, "unit": "reads/s", "cores": 1, "kind": "port", "gcups": c / dt / 1e9,
                                    "sample": f"one single-threaded oracle qluster on a {n}-read sample of the same generator ({dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    al.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
