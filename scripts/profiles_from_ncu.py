#!/usr/bin/env python
"""Turn the ncu artefacts of one round into the tracked summaries under profiles/.
  python scripts/profiles_from_ncu.py <tag> <launches.csv> <full.ncu-rep> "<bench command>"
Writes profiles/<tag>_bench_launches.md (every kernel launch of the bench command with its duration and the solver kernel's
share), profiles/<tag>_kernel_ncu_details.txt (ncu --page details of the full capture) and profiles/<tag>_traffic.json
(dram__bytes_read/write of that capture)."""
import csv, hashlib, io, json, os, subprocess, sys
tag, launches, rep, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(launches)) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
iK, iB, iG, iS, iM, iV, iU = (hdr.index(k) for k in ("Kernel Name", "Block Size", "Grid Size", "Stream", "Metric Name", "Metric Value", "Metric Unit"))
L = []
for r in rows[hi + 1:]:
    if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    unit = r[iU]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    L.append((r[iK], r[iB], r[iG], r[iS], ms))
tot = sum(x[4] for x in L)
solver = sum(x[4] for x in L if "kmpc_warp_kernel" in x[0])
with open(os.path.join(ROOT, "profiles", f"{tag}_bench_launches.md"), "w") as f:
    f.write(f"# {tag}: every kernel launch of `{cmd}` (1 x B200)\n\n")
    f.write("Captured with `ncu --metrics gpu__time_duration.sum --clock-control none --csv` after the same command exited 0 without ncu.\n")
    f.write("Per-launch times under ncu are serialised/cold-cache: compare shares, not absolutes.\n\n")
    f.write(f"`kmpc_warp_kernel` share of all device time in the process: {solver / tot * 100:.1f} % ({solver:.1f} ms of {tot:.1f} ms); the rest is the "
            "FP64-peak micro-benchmark (`kmpc_dfma_kernel`), the 256 MB L2-flush fills and torch bookkeeping kernels of bench.py.\n\n")
    f.write("| # | kernel | block | grid | stream | ms |\n|---|---|---|---|---|---|\n")
    for i, (k, b, g, s, ms) in enumerate(L):
        f.write(f"| {i} | `{k[:90]}` | {b} | {g} | {s} | {ms:.3f} |\n")
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"{tag}_kernel_ncu_details.txt"), "w").write(det)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rr[0], rr[2]))
units = dict(zip(rr[0], rr[1]))
def byt(k):
    v = float(d[k].replace(",", "")); u = units[k].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
rd, wr = byt("dram__bytes_read.sum"), byt("dram__bytes_write.sum")
_h = hashlib.sha256()   # same hash as bench.py kernel_source_sha(): the capture belongs to these kernel sources and no others
for f in ("kmpc.cu", "kmpc_core.cuh", "kmpc_warp.cuh", "kmpc_warp_prims.cuh", "kmpc_order_prior.h"):
    _h.update(open(os.path.join(ROOT, "kiss_mpc_b200", "csrc", f), "rb").read())
json.dump({"kernel": d.get("Kernel Name", "kmpc_warp_kernel"), "kernel_source_sha": _h.hexdigest()[:16], "workload": "65536 instances, N=30, cold start (scripts/one_solve.py 65536 30 1)",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 65536 * 1288,
           "duration_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(units["gpu__time_duration.sum"], 1e-6),
           "fp64_pipe_pct_of_peak_active": float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]),
           "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
           "dram_throughput_pct": float(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]) if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in d else None,
           "tensor_pipe_pct": float(d.get("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", 0) or 0),
           "registers_per_thread": int(float(d["launch__registers_per_thread"])),
           "source": f"ncu --set full --clock-control none, profiles/{tag}_kernel_ncu_details.txt"},
          open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print("wrote profiles for", tag, "solver share %.1f%%" % (solver / tot * 100), "dram bytes", rd + wr)
