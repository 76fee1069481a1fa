#!/usr/bin/env python
"""Headline benchmark: batched MPC solves/sec at N=30 for ONE 65,536-instance batch on 1/2/4/8 B200 (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (the per-step NLP solve of mpc/optimizer.py:319-400) over the whole synthetic batch of
65,536 independent instances (seed 1000), cold start.  With N ranks the batch is split by contiguous slices
(`shard_range(65536, rank, N)`, STRONG scaling) and the results are gathered on rank 0 INSIDE the timed region: every rank's
solver kernel writes its finished instances straight into one buffer in rank 0's HBM over NVLink (CUDA IPC peer mapping,
kiss_mpc_b200.RankGather) -- no collective in the solve, none after it.
  value      65,536 / (mean over steps of the max-over-ranks device time of the step); inputs resident in each rank's HBM
  e2e        the same batch through the public host API at N GPUs: NumPy in -> ShardedMotionPlanner.solve (one process, one
             handle + stream per device, H2D + solve + results into one pinned buffer) -> NumPy out, wall clock on rank 0
  roofline   dominant kernel against the bound that binds, the FP64 FMA pipe: algorithmic flops (970 N per interior-point
             iteration x measured iterations) / (DFMA peak measured in this run); roofline_hbm: algorithmic I/O bytes
             (1,288 B per cold-start solve) / measured HBM peak -- not binding, the iterate lives in registers
  cpu_baseline  the oracle (C restatement of IPOPT's algorithm, OpenMP, all host cores) on the same batch -- a reported
             baseline, "IPOPT-restatement, not IPOPT" (CasADi/IPOPT are not installable here)
  weak       round 1's figure, kept for comparison: every rank solves its own 65,536 batch (seeds 1000 + rank)
--impl reference times that CPU oracle alone (the reference's own CasADi/IPOPT path cannot run in this image).
"""
import argparse
import faulthandler
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
faulthandler.enable()      # a native crash leaves a Python traceback on stderr

BATCH = 65536
HORIZON = 30
TIME_STEP = 0.1
SEED = 1000
CPU_SAMPLE = 65536                    # the CPU arm solves the whole batch once (~1 s on 16 cores = 15-20 CPU-seconds)
ALG_BYTES_PER_SOLVE = 1288.0          # cold start, N=30: 48 B in + (5N+3)*8 + 16 B out
FLOPS_PER_ITER = 970.0 * HORIZON      # F_iter(N, O=0) = 970 N
WORKLOAD = f"one {BATCH}-instance batch, N={HORIZON}, T={TIME_STEP}, box bounds only, cold start, seed {SEED}, sharded by contiguous slices over the ranks"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    h = hashlib.sha256()
    for f in ("kmpc.cu", "kmpc_core.cuh", "kmpc_warp.cuh", "kmpc_warp_prims.cuh", "kmpc_order_prior.h"):
        h.update(open(os.path.join(ROOT, "kiss_mpc_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per launch of the solver kernel from an ncu --set full capture -- only if that capture was taken from the kernel
    sources as they are now (profiles/*_traffic.json carries their hash); a capture of an older kernel is not this run's traffic."""
    import glob
    sha = kernel_source_sha()
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        d = json.load(open(p))
        if d.get("kernel_source_sha") == sha:
            return d.get("dram_bytes_per_launch"), os.path.basename(p)
    return None, None


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


def cpu_oracle_rate(batch, nthreads=0, sample=CPU_SAMPLE, **cfg_kw):
    from oracle import oracle as ok
    ok.build()
    nthreads = nthreads or os.cpu_count()      # explicit: torchrun exports OMP_NUM_THREADS=1
    kw = dict(N=HORIZON, T=TIME_STEP, linsolve="riccati"); kw.update(cfg_kw)
    cfg = ok.OracleConfig(**kw)
    x, g = batch["x_cur"][:sample], batch["goal"][:sample]
    obs = None if batch.get("obs") is None else batch["obs"][:sample]
    ok.solve(cfg, x[:64], g[:64], obs=None if obs is None else obs[:64], nthreads=nthreads)  # warm the thread pool
    t0 = time.perf_counter()
    r = ok.solve(cfg, x, g, obs=obs, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample / dt, dt, r


def oracle_p50_us(x, g, reps=15, **cfg_kw):
    """Single-core latency of the oracle on one instance (what one reference solve would be compared with)."""
    from oracle import oracle as ok
    cfg = ok.OracleConfig(linsolve="riccati", **cfg_kw)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = ok.solve(cfg, x, g, nthreads=1); ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e6), int(r.iters[0])


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU path for this metric.  CasADi/IPOPT cannot be installed offline, so this is the
    oracle port (oracle/kmpc_oracle.c, OpenMP over all host cores) on the same batch.  Rank 0 alone works."""
    if rank != 0:
        return
    from kiss_mpc_b200.synthetic import make_batch
    batch = make_batch(BATCH, seed=SEED)
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_oracle_rate(batch, sample=2048)
    times = []
    for _ in range(args.steps):
        _, dt, _ = cpu_oracle_rate(batch)
        times.append(dt)
    v = CPU_SAMPLE * len(times) / sum(times)
    sample = f"all {CPU_SAMPLE} instances of the batch per step, OpenMP one instance per thread"
    print(json.dumps({
        "impl": "reference", "metric": "mpc_solves_per_sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "CPU arm: IPOPT-restatement (oracle port), not IPOPT -- casadi==3.7.1 is not installable offline"},
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def side_configs(torch, dev, ok):
    """One-repetition numbers of BASELINE's other configurations with their parity against the oracle (N = 1 GPU only)."""
    from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
    from kiss_mpc_b200.synthetic import make_batch
    out = {}

    def timed(fn, reps=3):
        best, r = 1e30, None
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); r = fn(); e.record(); torch.cuda.synchronize(dev)
            best = min(best, s.elapsed_time(e))
        return best, r

    def parity(r, ref, n):
        st = r.status[:n].cpu().numpy(); U = r.controls[:n].cpu().numpy(); obj = r.objective[:n].cpu().numpy()
        conv = (st == 0) & (ref.status == 0)
        return {"sample": n, "status_equal": float((st == ref.status).mean()),
                "max_abs_control_diff": float(np.abs(U - ref.U)[conv].max()) if conv.any() else None,
                "max_rel_objective_diff": float((np.abs(obj - ref.obj) / np.abs(ref.obj))[conv].max()) if conv.any() else None}

    for name, B, N, O, seed, nref in (("cfg2_4096_N30", 4096, 30, 0, 1002, 4096), ("cfg3_65536_N50", 65536, 50, 0, 1003, 2048),
                                      ("cfg4_4096_N30_O10", 4096, 30, 10, 1004, 1024)):
        b = make_batch(B, seed=seed, O=O)
        pl = BatchedMotionPlanner(PlannerConfig(N=N, O_max=O), max_batch=B, device=dev.index)
        x, g = torch.tensor(b["x_cur"], device=dev), torch.tensor(b["goal"], device=dev)
        ob = None if not O else torch.tensor(b["obs"], device=dev)
        fn = lambda: pl.solve(x, g, obstacles=ob, obstacle_radius=0.3, inflation_radius=0.5 if O else 0.0)
        fn()
        ms, r = timed(fn)
        _, _, ref = cpu_oracle_rate(b, sample=nref, N=N, O=O)
        out[name] = {"ms": ms, "solves_per_s": B / ms * 1e3, "mean_iters": float(r.iters.float().mean()), "parity_vs_oracle": parity(r, ref, nref)}
        pl.close()
    # cfg 5: closed loop, 16,384 agents x 200 steps, warm-started, agents stop at their goal (environment.py:31-33)
    B, steps = 16384, 200
    b = make_batch(B, seed=1005)
    pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B, device=dev.index)
    g = torch.tensor(b["goal"], device=dev)
    best = None
    for _ in range(2):
        x = torch.tensor(b["x_cur"], device=dev)
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); X, U, app, it, st = pl.closed_loop(x, g, steps, log_applied=False, goal_radius=0.5); e.record(); torch.cuda.synchronize(dev)
        ms = s.elapsed_time(e)
        if best is None or ms < best[0]:
            best = (ms, int((st != 1000).sum()), float(it[st != 1000].float().mean()), float((st[st != 1000] == 0).float().mean()))
    out["cfg5_closed_loop_16384x200_stop_at_goal"] = {"ms": best[0], "solves": best[1], "solves_per_s": best[1] / best[0] * 1e3,
                                                       "mean_iters": best[2], "converged_fraction": best[3],
                                                       "note": "parity of this loop: tests/test_parity_gpu.py::test_closed_loop_matches_stepwise_oracle"}
    pl.close()
    return out


def latency_b1(ok):
    """p50 latency of ONE solve through the drop-in MotionPlanner (NumPy in/out, B = 1): BASELINE cfg 1 (N=30, T=0.1) and the ROS
    node's configuration (N=7, T=0.8, +-0.3 bounds, ros2interface.py:28-38), with the oracle's single-core p50 beside them."""
    from kiss_mpc_b200 import MotionPlanner
    from kiss_mpc_b200.synthetic import cfg1_instance
    out = {}
    for name, N, T, vb, wb in (("N30_T0.1", 30, 0.1, (-0.2, 0.5), (-0.5, 0.5)), ("N7_T0.8_ros", 7, 0.8, (-0.3, 0.3), (-0.3, 0.3))):
        x, g = cfg1_instance()
        mp = MotionPlanner(time_step=T, horizon=N, on_failure="ignore")
        X0 = np.tile(x[0], (N + 1, 1)).T; U0 = np.zeros((2, N))
        kw = dict(current_state=x[0], goal_state=g[0], states_matrix=X0, controls_matrix=U0, linear_velocity_bounds=vb, angular_velocity_bounds=wb)
        for _ in range(5):
            mp.solve(**kw)
        ts = []
        for _ in range(60):
            t0 = time.perf_counter(); mp.solve(**kw); ts.append(time.perf_counter() - t0)
        cpu_us, cpu_it = oracle_p50_us(x, g, N=N, T=T, v_bounds=vb, w_bounds=wb)
        out[name] = {"gpu_p50_us": float(np.median(ts) * 1e6), "gpu_iters": mp.last_iterations, "status": mp.last_status,
                     "oracle_1core_p50_us": cpu_us, "oracle_iters": cpu_it}
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, RankGather, ShardedMotionPlanner, shard_range
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
            host_group = dist.new_group(backend="gloo")     # host-side waits that leave no collective kernel spinning on a GPU
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    B, N = BATCH, HORIZON
    steps, warm = args.steps, max(args.warmup, 3)
    batch = make_batch(B, seed=SEED)                   # the ONE batch; rank r owns the slice shard_range(B, r, world)
    lo, hi = shard_range(B, rank, world)
    planner = BatchedMotionPlanner(PlannerConfig(N=N, T=TIME_STEP), max_batch=B, device=local_rank)
    x = torch.tensor(batch["x_cur"][lo:hi], device=dev); g = torch.tensor(batch["goal"][lo:hi], device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- result gather: straight into rank 0's memory from inside the solver kernel (IPC peer mapping); NCCL gather otherwise ----
    transport = os.environ.get("KMPC_BENCH_GATHER", "ipc")
    gather = None
    if world > 1:
        okf = torch.ones(1, device=dev)
        try:
            gather = RankGather(planner, B, transport=transport)
        except Exception as e:      # no peer mapping on this box: every rank must take the same fall-back
            sys.stderr.write(f"rank {rank}: RankGather({transport}) failed: {e}\n")
            okf.zero_()
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if okf.item() == 0:
            if gather is not None:
                gather.close()
            transport = "nccl"
            gather = RankGather(planner, B, transport="nccl")

    def step():
        if gather is not None:
            gather.solve(x, g)
            return None
        return planner.solve(x, g)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = None
    for _ in range(warm):
        res = step()
    barrier()
    fp64_peak = planner.measure_fp64_peak() if rank == 0 else 0.0

    # ---- device-resident throughput: per step  L2 flush -> barrier -> [event] solve (+ gather) [event]; max over ranks per step ----
    l0 = planner.stats()["launches"]
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    m0 = sampler.mark()
    w0 = time.perf_counter()
    for s, e in evs:
        flush.fill_(1)
        if world > 1:
            dist.barrier()
        s.record()
        res = step()
        e.record()
    barrier()
    wall = time.perf_counter() - w0
    m1 = sampler.mark()
    launches = planner.stats()["launches"] - l0
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    t = torch.tensor([s.elapsed_time(e) for s, e in evs], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.mean().item())
    value = B / (ms_per_step * 1e-3)
    if gather is not None:
        res = gather.result()          # rank 0: the whole batch, gathered; other ranks: None

    # ---- weak scaling (round 1's definition), kept as an extra: every rank its own 65,536 batch ----
    weak = None
    if world > 1:
        wb = make_batch(B, seed=SEED + rank)
        xw, gw = torch.tensor(wb["x_cur"], device=dev), torch.tensor(wb["goal"], device=dev)
        planner.solve(xw, gw)
        wevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(steps, 5))]
        barrier()
        for s, e in wevs:
            flush.fill_(1); dist.barrier(); s.record(); planner.solve(xw, gw); e.record()
        barrier()
        tw = torch.tensor([s.elapsed_time(e) for s, e in wevs], device=dev, dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak = {"value": world * B / (float(tw.mean().item()) * 1e-3), "unit": "solves/s", "ms_per_step": float(tw.mean().item()),
                "workload": f"{B}-instance batch PER GPU, seeds {SEED}+rank (round 1's weak-scaling figure)"}
        del xw, gw

    # ---- end to end through the public host API at `world` GPUs: rank 0 drives all devices through ShardedMotionPlanner ----
    e2e_value, rh = None, None
    h2d = 2 * B * 3 * 8
    d2h = B * ((5 * N + 3) * 8 + 8 + 4 + 4)
    if world > 1:
        barrier()
        dist.barrier(group=host_group)
    if rank == 0:
        sp = ShardedMotionPlanner(PlannerConfig(N=N, T=TIME_STEP), max_batch=B, devices=list(range(world)))
        xh, gh = batch["x_cur"], batch["goal"]
        for _ in range(2):
            sp.solve(xh, gh)
        e0 = time.perf_counter()
        chk = 0.0
        for _ in range(steps):
            rh = sp.solve(xh, gh)
            chk += float(rh.objective[0]) + int(rh.status[-1])      # the host reads the result of every step
        e2e_s = time.perf_counter() - e0
        e2e_value = B * steps / e2e_s
    if world > 1:
        dist.barrier(group=host_group)     # the other ranks wait on the host: their GPUs are free for rank 0's per-device handles

    if rank == 0:
        from oracle import oracle as ok
        ok.build()
        hbm_peak, peak_src = peaks()
        kernel_s = ms_per_step * 1e-3
        iters = res.iters.double()
        mean_it, max_it = float(iters.mean().item()), float(iters.max().item())
        conv_frac = float((res.status == 0).double().mean().item())
        # the dominant kernel is timed on its own (1 GPU: the whole step is that kernel + the queue-order kernels, < 1 %)
        ach_gbs = ALG_BYTES_PER_SOLVE * B / kernel_s / 1e9
        ach_tf = FLOPS_PER_ITER * mean_it * B / kernel_s / 1e12 / world      # per GPU
        cpu_rate, cpu_dt, cpu_res = cpu_oracle_rate(batch, sample=CPU_SAMPLE)
        Ug, sg, og = res.controls.cpu().numpy(), res.status.cpu().numpy(), res.objective.cpu().numpy()
        conv_both = (sg == 0) & (cpu_res.status == 0)
        du = np.abs(Ug - cpu_res.U).max(axis=(1, 2))
        parity = {"sample": CPU_SAMPLE, "status_equal": float((sg == cpu_res.status).mean()),
                  "iterations_equal": float((res.iters.cpu().numpy() == cpu_res.iters).mean()),
                  "max_abs_control_diff": float(du[conv_both].max()),
                  "instances_above_1e-9": int((du[conv_both] > 1e-9).sum()),
                  "max_rel_objective_diff": float((np.abs(og - cpu_res.obj) / np.abs(cpu_res.obj))[conv_both].max()),
                  "host_path_equals_device_path": bool(np.array_equal(rh.controls, Ug) and np.array_equal(rh.status, sg))}
        traffic, traffic_src = measured_traffic()
        extra = {}
        if world == 1:
            extra["p50_us_b1"] = latency_b1(ok)
            extra["configs"] = side_configs(torch, dev, ok)
        print(json.dumps({
            "metric": "mpc_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B, "per_rank": -(-B // world),
                       "timing": "per step: 256 MB L2 flush, barrier, CUDA events around solve (+ gather) on the solve stream; max over ranks per step, mean over steps",
                       "gather": "none (1 GPU)" if world == 1 else (f"inside the timed region, transport {transport}: " +
                                 ("every rank's solver kernel writes into one buffer in rank 0's HBM over NVLink (CUDA IPC)" if transport == "ipc" else "one torch.distributed.gather of a packed byte tensor (NCCL)")),
                       "kernel": "kmpc_warp_kernel<SPL=1,FULL> (warp per instance + block-cooperative Riccati lane; per step 1 solver launch + the queue-order key kernel and radix sort)",
                       "e2e_api": "ShardedMotionPlanner.solve(numpy, numpy) -> kmpc_solve_host_into per device (pinned staging, H2D, solve, results written into one pinned buffer) -> numpy views",
                       "parallelism": f"batch slices x{world}, no collective in the solve"},
            "p50_us_per_solve_amortised": ms_per_step * 1e3 / B, "wall_s_timed_region": wall,
            "mean_ipm_iterations": mean_it, "max_ipm_iterations": max_it, "converged_fraction": conv_frac,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
                         "traffic": traffic, "traffic_source": traffic_src, "flops_per_iteration": FLOPS_PER_ITER, "per": "GPU",
                         "peak_source": "DFMA micro-benchmark in this run (kmpc_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                         "note": "the binding roof: the iterate lives in registers / shared memory, HBM carries I/O only (roofline_hbm)"},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs / world, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / world / hbm_peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_solve": ALG_BYTES_PER_SOLVE},
            "cpu_baseline": {"value": cpu_rate, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"all {CPU_SAMPLE} instances of the batch, oracle (IPOPT-restatement, not IPOPT), OpenMP all cores, {cpu_dt:.2f} s"},
            "parity_vs_oracle": parity,
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "weak": weak, "gpu_launches": int(launches), "clocks": clocks, **extra,
        }))
        rh = None
        sp.close()          # (the pinned result buffer `rh` viewed)
    if gather is not None:
        barrier()
        gather.close()
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
