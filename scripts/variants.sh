#!/bin/bash
# Build launch-shape variants of the warp kernel HERE (nvcc cross-compiles), time them on the GPU box with
#   KMPC_LIB=gpurun_variants/libkmpc_<tag>.so python scripts/one_solve.py 65536 30 3
# usage: scripts/variants.sh "WPB1 MINB1 WPB2 MINB2" ...
mkdir -p gpurun_variants
for v in "$@"; do
  set -- $v
  tag="$1_$2_$3_$4"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart shared \
    -DKMPC_WPB1=$1 -DKMPC_MINB1=$2 -DKMPC_WPB2=$3 -DKMPC_MINB2=$4 -o gpurun_variants/libkmpc_$tag.so kiss_mpc_b200/csrc/kmpc.cu &
done
wait
ls -la gpurun_variants
