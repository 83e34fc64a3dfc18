#!/usr/bin/env python
"""bench.py — headline benchmark of the qluster hot path on B200 (driver contract in the task).

Workload (BASELINE.json configs[1]): one synthetic sample of 300-bp amplicon
reads from 200 haplotypes at 1 % substitutions + homopolymer indels.  A "step" is one pass of
the hot path over one collapse-iteration batch of candidate (read, cluster) pairs: every
unique read against the `stopCheck` = 100 most abundant clusters, each pair going through
affine-gap global alignment with traceback and the errorProfile reduction.

  value  : reads/s with everything resident in HBM (CUDA events on the library's stream)
  e2e    : same metric through the host-buffer C-ABI calls (sequence upload, pair-list H2D,
           comparison structs D2H inside the timed region)
  gcups  : DP cell updates / s of the alignment kernel (second half of BASELINE.json's metric)
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


def make_workload(seed, nReads, stopCheck):
    """Unique reads of a C2-shaped sample, sorted by abundance; pairs = reads x top clusters."""
    from seekdeep_b200 import synth
    haps, seqs, quals, _ = synth.config2_fast(seed=seed, nReads=nReads)
    uniq = {}
    for s, q in zip(seqs, quals):
        k = s.tobytes()
        if k in uniq:
            uniq[k][1] += 1
        else:
            uniq[k] = [q, 1]
    items = sorted(uniq.items(), key=lambda kv: (-kv[1][1], kv[0]))
    useqs = [np.frombuffer(k, np.uint8) for k, _ in items]
    uquals = [v[0] for _, v in items]
    cnts = np.array([v[1] for _, v in items], np.float64)
    bases, q, offsets = synth.pack(useqs, uquals)
    nU = len(useqs)
    nClus = min(stopCheck, nU)
    reads = np.arange(nClus, nU, dtype=np.uint32)
    refIdx = np.tile(np.arange(nClus, dtype=np.uint32), len(reads))
    readIdx = np.repeat(reads, nClus)
    return dict(bases=bases, quals=q, offsets=offsets, cnts=cnts, refIdx=refIdx, readIdx=readIdx, nReadsInput=nReads,
                nUnique=nU, nClus=nClus)


def pair_cells(w):
    lens = np.diff(w["offsets"]).astype(np.int64)
    return int((lens[w["refIdx"]] * lens[w["readIdx"]]).sum())


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


def cpu_oracle_batch(args):
    """Worker for the CPU arms: time the oracle on a slice of pairs (1 thread)."""
    seed, nPairs, flags = args
    from tests import oracle_binding
    import seekdeep_b200 as sd
    o = oracle_binding.load()
    w = make_workload(seed, 4000, 100)
    sel = np.random.default_rng(seed).choice(len(w["refIdx"]), nPairs, replace=False)
    t0 = time.perf_counter()
    o.align_profile_batch(sd.make_params(), w["bases"], w["quals"], w["offsets"], w["refIdx"][sel], w["readIdx"][sel], flags)
    dt = time.perf_counter() - t0
    lens = np.diff(w["offsets"]).astype(np.int64)
    cells = int((lens[w["refIdx"][sel]] * lens[w["readIdx"][sel]]).sum())
    return dt, nPairs, cells


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The real binary cannot
    be built (njhseq un-vendored), so this times the oracle port on all host cores, one
    single-threaded worker per core like `runMultipleCommands --numThreads` (SeekDeepUtilsRunner.cpp:294-356)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    pairsPerWorker = 3000
    stopCheck = 100
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(cpu_oracle_batch, [(1000 + step * cores + c, pairsPerWorker, 2) for c in range(cores)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append((dt, sum(r[1] for r in res), sum(r[2] for r in res)))
    tot = sum(t[0] for t in times)
    pairs = sum(t[1] for t in times)
    cells = sum(t[2] for t in times)
    readsPerS = pairs / stopCheck / tot
    line = {"impl": "reference", "metric": "qluster_collapse_pass_reads_per_s", "value": readsPerS, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, len(times)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "gcups": cells / tot / 1e9,
            "config": workload_config(stopCheck),
            "cpu_baseline": {"value": readsPerS, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": f"{pairsPerWorker} (read, cluster) pairs per core per step of the same C2-shaped workload"},
            "e2e": {"value": readsPerS, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(stopCheck):
    return {"workload": "SeekDeep qluster single sample, 300-bp amplicon reads from 200 haplotypes at 1% error with "
                        "homopolymer indels (BASELINE configs[1]); one collapse pass = every unique read vs the "
                        f"{stopCheck} most abundant clusters: affine NW + traceback + errorProfile per pair",
            "stopCheck": stopCheck, "gap": "5,1 left 5,1 right 0,0", "qualThres": "25,20",
            "l2": "no flush: per-launch working set (traceback-bit slabs + outputs, >300 MB) exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sdgpu")
    ap.add_argument("--reads", type=int, default=int(os.environ.get("SD_BENCH_READS", 20000)),
                    help="input reads of the sample per step and per GPU")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import seekdeep_b200 as sd
    from seekdeep_b200 import _abi

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stopCheck = 100
    w = make_workload(2 + rank, args.reads, stopCheck)  # one sample per GPU (weak scaling, sharded by sample)
    nPairs = len(w["refIdx"])
    cells = pair_cells(w)
    nReadsStep = len(np.unique(w["readIdx"]))
    al = sd.GpuAligner(sd.make_params(), device=local)
    al.upload_seqs(w["bases"], w["quals"], w["offsets"], w["cnts"])
    dRef, dRead = al.dev_alloc(nPairs * 4), al.dev_alloc(nPairs * 4)
    dOut = al.dev_alloc(nPairs * _abi.COMPARISON_DTYPE.itemsize)
    al.h2d(dRef, w["refIdx"])
    al.h2d(dRead, w["readIdx"])
    l0, _ = al.counters()

    def barrier():
        al.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident arm
    for _ in range(args.warmup):
        al.align_profile_dev(dRef, dRead, nPairs, dOut)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    al.timer_start()
    for _ in range(args.steps):
        al.align_profile_dev(dRef, dRead, nPairs, dOut)
    ms = al.timer_stop()
    barrier()
    l1, _ = al.counters()
    # ---- end-to-end arm: host buffers in, host structs out, sequence upload included
    al2 = sd.GpuAligner(sd.make_params(), device=local)
    for _ in range(2):
        al2.upload_seqs(w["bases"], w["quals"], w["offsets"], w["cnts"])
        al2.align_profile(w["refIdx"], w["readIdx"])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        al2.upload_seqs(w["bases"], w["quals"], w["offsets"], w["cnts"])
        out = al2.align_profile(w["refIdx"], w["readIdx"])
    al2.sync()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    barrier()
    h2d = int(w["bases"].nbytes + w["quals"].nbytes + w["offsets"].nbytes + w["cnts"].nbytes + 2 * w["refIdx"].nbytes)
    d2h = int(out.nbytes)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t.tolist()
    tot = torch.tensor([float(nReadsStep), float(cells), float(nPairs)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    readsAll, cellsAll, pairsAll = tot.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        rates = [al.measure_int32_peak(k) for k in range(4)]
        int32_peak = max(rates)
        gcups_gpu0 = cells * args.steps / (ms * 1e-3) / 1e9
        peak_gcups = int32_peak / OPS_PER_CELL / 1e9
        line = {
            "metric": "qluster_collapse_pass_reads_per_s", "value": readsAll * args.steps / (ms_max * 1e-3), "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": dict(workload_config(stopCheck), reads_per_step_per_gpu=nReadsStep, pairs_per_step_per_gpu=nPairs,
                           input_reads_per_gpu=w["nReadsInput"], unique_per_gpu=w["nUnique"]),
            "gcups": cellsAll * args.steps / (ms_max * 1e-3) / 1e9,
            "pairs_per_s": pairsAll * args.steps / (ms_max * 1e-3),
            "e2e": {"value": readsAll * args.steps / (e2e_ms_max * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "gcups": cellsAll * args.steps / (e2e_ms_max * 1e-3) / 1e9},
            "gpu_launches": int(l1 - l0) - args.warmup,
            "clocks": sampler.summary(),
            "roofline": {"bound": "int32", "kernel": "alignProfileKernel", "achieved": gcups_gpu0, "peak": peak_gcups,
                         "unit": "GCUPS", "frac": gcups_gpu0 / peak_gcups, "traffic": None,
                         "peak_source": f"measured on this GPU: {int32_peak / 1e12:.2f} T INT32 instr/s "
                                        f"(add/max/addmax/max3 = {', '.join(f'{r / 1e12:.2f}' for r in rates)}) / {OPS_PER_CELL} ops per cell",
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
        }
        if not args.no_cpu:
            dt, n, c = cpu_oracle_batch((2, 15000, 2))
            line["cpu_baseline"] = {"value": n / stopCheck / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                                    "gcups": c / dt / 1e9,
                                    "sample": f"{n} (read, cluster) pairs of the same workload through the single-threaded oracle"}
        print(json.dumps(line), flush=True)
    al.close()
    al2.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
