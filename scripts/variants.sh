#!/bin/bash
# build + time launch-bound variants of the warp kernel on the GPU box
for mb in 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart shared -DKMPC_WARP_MINB=$mb -o kiss_mpc_b200/libkmpc.so kiss_mpc_b200/csrc/kmpc.cu
  echo "== MINB=$mb"; python scripts/one_solve.py 65536 30 2 | tail -1; python scripts/one_solve.py 65536 50 1 | tail -1
done
