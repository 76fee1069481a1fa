#!/usr/bin/env python
"""Headline benchmark: batched MPC solves/sec at N=30 for a 65,536-instance batch per GPU (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (the per-step NLP solve of mpc/optimizer.py:319-400) over one synthetic batch of
B = 65,536 independent instances per GPU (weak scaling: every rank owns its own batch slice, no data-path collective).
  value      whole-job solves/sec with inputs already resident in HBM (CUDA events on the solve stream, max over ranks)
  e2e        the same metric through the public host API (NumPy in -> kmpc_solve_host -> NumPy out), H2D/D2H inside
  roofline   dominant kernel vs the measured HBM peak on ALGORITHMIC bytes (1,288 B per cold-start solve),
             plus roofline_fp64: algorithmic FP64 flops (970*N per IPM iteration x measured iterations) vs the DFMA peak
             measured in this run -- the bound that actually binds (state lives in registers)
  cpu_baseline  the oracle (C restatement of IPOPT's algorithm, OpenMP, all host cores) on a bounded sample -- a
             reported baseline, "IPOPT-restatement, not IPOPT" (CasADi/IPOPT are not installable here)
--impl reference times that CPU oracle alone (the reference's own CasADi/IPOPT path cannot run in this image).
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

B_PER_GPU = 65536
HORIZON = 30
TIME_STEP = 0.1
SEED = 1000
CPU_SAMPLE = 65536                    # the CPU arm solves the whole batch once (~1 s on 16 cores = 15-20 CPU-seconds)
ALG_BYTES_PER_SOLVE = 1288.0          # cold start, N=30: 48 B in + (5N+3)*8 + 16 B out
FLOPS_PER_ITER = 970.0 * HORIZON      # F_iter(N, O=0) = 970 N


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Index of the next sample (nvidia-smi is started before the warm-up so that it is already streaming)."""
        return len(self.rows)

    def stop(self, i0=0, i1=None):
        """Stats over the samples taken between two marks (the timed region); if the region was shorter than one sampling
        period, the samples next to it are used and `window` says so."""
        if self.proc is not None:
            self.proc.terminate()
        rows, window = self.rows[i0:i1], "timed region"
        if not rows:
            rows, window = self.rows[max(0, i0 - 2):(i1 or 0) + 2], "timed region shorter than the sampling period: adjacent samples"
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": window}


def cpu_oracle_rate(batch, nthreads=0, sample=CPU_SAMPLE):
    from oracle import oracle as ok
    ok.build()
    nthreads = nthreads or os.cpu_count()      # explicit: torchrun exports OMP_NUM_THREADS=1
    cfg = ok.OracleConfig(N=HORIZON, T=TIME_STEP, linsolve="riccati")
    x, g = batch["x_cur"][:sample], batch["goal"][:sample]
    ok.solve(cfg, x[:64], g[:64], nthreads=nthreads)  # warm the thread pool
    t0 = time.perf_counter()
    r = ok.solve(cfg, x, g, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample / dt, dt, r


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU path for this metric.  CasADi/IPOPT cannot be installed offline, so this is the
    oracle port (oracle/kmpc_oracle.c, OpenMP over all host cores) on a bounded sample of the same workload."""
    if rank != 0:
        return
    from kiss_mpc_b200.synthetic import make_batch
    batch = make_batch(B_PER_GPU, seed=SEED)
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_oracle_rate(batch, sample=2048)
    rates, times = [], []
    for _ in range(args.steps):
        rate, dt, _ = cpu_oracle_rate(batch)
        rates.append(rate); times.append(dt)
    v = CPU_SAMPLE * len(times) / sum(times)
    sample = f"{CPU_SAMPLE} of the {B_PER_GPU} instances per step, OpenMP one instance per thread"
    print(json.dumps({
        "impl": "reference", "metric": "mpc_solves_per_sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{B_PER_GPU}-instance batch per GPU, N={HORIZON}, T={TIME_STEP}, box bounds only, cold start, seed {SEED}",
                   "note": "CPU arm: IPOPT-restatement (oracle port), not IPOPT -- casadi==3.7.1 is not installable offline"},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
    from kiss_mpc_b200.synthetic import make_batch

    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU: kiss_mpc_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL may print its version banner on stdout; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    B, N = B_PER_GPU, HORIZON
    batch = make_batch(B, seed=SEED + rank)          # every rank owns its own slice of the global batch
    planner = BatchedMotionPlanner(PlannerConfig(N=N, T=TIME_STEP), max_batch=B, device=local_rank)
    x = torch.tensor(batch["x_cur"], device=dev); g = torch.tensor(batch["goal"], device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = None
    for _ in range(max(args.warmup, 3)):
        res = planner.solve(x, g)
    torch.cuda.synchronize(dev)
    fp64_peak = planner.measure_fp64_peak() if rank == 0 else 0.0

    # ---- device-resident throughput: CUDA events around each solve on its stream, L2 flushed between steps ----
    l0 = planner.stats()["launches"]
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    m0 = sampler.mark()
    w0 = time.perf_counter()
    for s, e in evs:
        flush.fill_(1)
        s.record()
        res = planner.solve(x, g)
        e.record()
    barrier()
    wall = time.perf_counter() - w0
    m1 = sampler.mark()
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    launches = planner.stats()["launches"] - l0
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    iters = res.iters.double()
    conv = (res.status == 0).double().mean()
    stat = torch.stack([iters.mean(), iters.max(), conv])
    if world > 1:
        dist.all_reduce(stat, op=dist.ReduceOp.SUM); stat /= world

    # ---- end to end through the public host API: NumPy in, NumPy out, H2D + D2H inside the timed region ----
    # (copy=False: the returned NumPy arrays are views of the planner's pinned result buffers -- the D2H copy lands there)
    xh, gh = batch["x_cur"], batch["goal"]
    planner.solve(xh, gh, copy=False)
    barrier()
    e0 = time.perf_counter()
    chk = 0.0
    for _ in range(args.steps):
        rh = planner.solve(xh, gh, copy=False)
        chk += float(rh.objective[0]) + int(rh.status[-1])      # the host reads the result of every step
    barrier()
    e2e_s = time.perf_counter() - e0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d = 2 * B * 3 * 8
    d2h = B * ((5 * N + 3) * 8 + 8 + 4 + 4)

    if rank == 0:
        hbm_peak, peak_src = peaks()
        kernel_s = ms_per_step * 1e-3
        ach_gbs = ALG_BYTES_PER_SOLVE * B / kernel_s / 1e9
        mean_it = float(stat[0].item())
        ach_tf = FLOPS_PER_ITER * mean_it * B / kernel_s / 1e12
        # CPU arm beside it: the whole batch at N = 1 GPU, a 4,096-instance parity sample otherwise
        cpu_n = CPU_SAMPLE if world == 1 else 4096
        cpu_rate, cpu_dt, cpu_res = cpu_oracle_rate(batch, sample=cpu_n)
        Ug = rh.controls[:cpu_n]
        conv_both = (rh.status[:cpu_n] == 0) & (cpu_res.status == 0)
        parity = {"status_equal": float((rh.status[:cpu_n] == cpu_res.status).mean()),
                  "max_abs_control_diff": float(np.abs(Ug - cpu_res.U)[conv_both].max()),
                  "max_rel_objective_diff": float((np.abs(rh.objective[:cpu_n] - cpu_res.obj) / np.abs(cpu_res.obj))[conv_both].max())}
        traffic = None
        import glob
        tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))   # newest round's ncu --set full capture
        if tps:
            traffic = json.load(open(tps[-1])).get("dram_bytes_per_launch")
        print(json.dumps({
            "metric": "mpc_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{B}-instance batch per GPU, N={N}, T={TIME_STEP}, box bounds only, cold start, seed {SEED}+rank",
                       "global_batch": world * B, "timing": "CUDA events per step on the solve stream, L2 flushed (256 MB fill) between steps, max over ranks",
                       "kernel": "kmpc_warp_kernel<SPL=1,FULL> (warp per instance + block-cooperative Riccati lane; per step 1 solver launch + the queue-order key kernel and radix sort)",
                       "e2e_api": "BatchedMotionPlanner.solve(numpy, numpy, copy=False) -> kmpc_solve_host (pinned staging, H2D + D2H inside)", "parallelism": f"batch slices x{world}, no collective in the solve"},
            "p50_us_per_solve_amortised": ms_per_step * 1e3 / B, "wall_s_timed_region": wall,
            "mean_ipm_iterations": mean_it, "max_ipm_iterations": float(stat[1].item()), "converged_fraction": float(stat[2].item()),
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_solve": ALG_BYTES_PER_SOLVE,
                         "note": "iterate stays in registers: compulsory HBM bytes are I/O only, so the HBM roofline is not the binding one"},
            "roofline_fp64": {"bound": "fp64_fma_pipe", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
                              "flops_per_iteration": FLOPS_PER_ITER, "peak_source": "DFMA micro-benchmark in this run (kmpc_measure_fp64_peak)"},
            "cpu_baseline": {"value": cpu_rate, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"first {cpu_n} instances of rank 0's batch, oracle (IPOPT-restatement, not IPOPT), OpenMP all cores, {cpu_dt:.2f} s"},
            "parity_vs_oracle_on_sample": parity,
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
