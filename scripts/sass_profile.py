#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counters with nvdisasm's inline line info and aggregate them by source
function / call site.  Usage:
    python scripts/sass_profile.py gpurun_out/prof.ncu-rep [kernel-substring] [--lines]
Needs ncu, cuobjdump and nvdisasm (CUDA toolkit) and the libkmpc.so the report was taken from (same build).
Output: executed warp instructions and stall samples per (outermost kmpc_warp.cuh line inside w_worker) and per
innermost function file:line bucket."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_with_lines(so, kern_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
    out, cur_chain, in_k = [], [], False
    pend = []
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            in_k = kern_sub in ln
            continue
        if not in_k:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pend.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pend:
                cur_chain = pend
                pend = []
            out.append((int(m.group(1), 16), m.group(2).strip(), list(cur_chain)))
    return out


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "kmpc_warp_kernelILi1"
    so = os.path.abspath(os.environ.get("KMPC_LIB", os.path.join(ROOT, "kiss_mpc_b200", "libkmpc.so")))
    sass = sass_with_lines(so, kern)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(int(r[0], 16), r[1].strip(), int(r[iI]), int(r[iS])) for r in rows[2:] if len(r) > iI]
    base = data[0][0]
    if len(data) != len(sass):
        print(f"warning: {len(data)} profiled instructions vs {len(sass)} disassembled (different build?)")
    by_off = {o: (t, ch) for o, t, ch in sass}
    tot_i = sum(d[2] for d in data)
    tot_s = sum(d[3] for d in data)
    site = collections.Counter(); site_s = collections.Counter(); site_n = collections.Counter()
    inner = collections.Counter(); inner_s = collections.Counter()
    for addr, txt, n, smp in data:
        t, ch = by_off.get(addr - base, (None, []))
        # outermost kmpc_warp.cuh frame = the call site inside w_worker
        top = None
        for f, l in ch:
            if f == "kmpc_warp.cuh":
                top = l
        key = f"w_worker:{top}" if top else (f"{ch[-1][0]}:{ch[-1][1]}" if ch else "?")
        site[key] += n; site_s[key] += smp; site_n[key] += 1
        ik = f"{ch[0][0]}:{ch[0][1]}" if ch else "?"
        inner[ik] += n; inner_s[ik] += smp
    print(f"total warp instructions {tot_i:.4g}, samples {tot_s}")
    print("-- by call site in w_worker (outermost kmpc_warp.cuh line): %instr  %samples  #sass")
    for k, v in site.most_common(40):
        print(f"{k:28s} {v / tot_i * 100:6.2f}% {site_s[k] / max(1, tot_s) * 100:6.2f}%  {site_n[k]}")
    if "--lines" in sys.argv:
        print("-- by innermost file:line")
        for k, v in inner.most_common(60):
            print(f"{k:34s} {v / tot_i * 100:6.2f}% {inner_s[k] / max(1, tot_s) * 100:6.2f}%")


if __name__ == "__main__":
    main()
