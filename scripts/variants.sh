#!/bin/bash
# build + time launch-bound variants of the warp kernel on the GPU box:  "warps_per_block min_blocks"
for v in "4 2" "8 1" "2 4" "6 2" "4 3"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart shared -DKMPC_WARPS_PER_BLOCK=$1 -DKMPC_WARP_MINB=$2 -o kiss_mpc_b200/libkmpc.so kiss_mpc_b200/csrc/kmpc.cu
  echo "== WPB=$1 MINB=$2"; python scripts/one_solve.py 65536 30 2 | tail -1; python scripts/one_solve.py 65536 50 1 | tail -1
done
