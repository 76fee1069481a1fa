#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), launch list and one full ncu capture of the solver kernel.
# usage (under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_gputests.log
tail -3 gpurun_out/${tag}_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
python scripts/one_solve.py 65536 30 3 > gpurun_out/${tag}_one_solve.log 2>&1
python scripts/b1_latency.py 30 >> gpurun_out/${tag}_one_solve.log 2>&1
python scripts/b1_latency.py 7 >> gpurun_out/${tag}_one_solve.log 2>&1
cat gpurun_out/${tag}_one_solve.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmpc_warp_kernel -c 1 -o gpurun_out/${tag}_full -f \
  python scripts/one_solve.py 65536 30 1 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out
# obstacle batches: warp kernel vs restoration finisher (per-launch durations under ncu)
for args in "65536 10 0" "4096 10 1" "65536 10 1"; do
  echo "== diag_resto $args" >> gpurun_out/${tag}_resto_split.txt
  python scripts/diag_resto.py $args >> gpurun_out/${tag}_resto_split.txt 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/diag_resto.py $args 2>/dev/null | grep -E "kmpc_(warp|finish)" | awk -F'","' '{print $5, $NF}' | tail -2 >> gpurun_out/${tag}_resto_split.txt
done
python scripts/bench_configs.py 3 > gpurun_out/${tag}_all_configs.json 2> gpurun_out/${tag}_all_configs.err
python scripts/strong_slices.py > gpurun_out/${tag}_strong_slices.json 2>/dev/null
python scripts/b1_breakdown.py 30 0.1 > gpurun_out/${tag}_b1_breakdown.txt 2>&1; python scripts/b1_breakdown.py 7 0.8 >> gpurun_out/${tag}_b1_breakdown.txt 2>&1
cat gpurun_out/${tag}_resto_split.txt gpurun_out/${tag}_b1_breakdown.txt
