// kmpc_warp_prims.cuh -- the handful of warp-collective primitives the warp-per-instance solver uses.
// Device build: CUDA shuffle / vote intrinsics.  Host build (tests/host_emul only): a cooperative 32-fibre emulator
// (tests/host_emul/simt.h) supplies the same functions so the identical solver source runs on a GPU-less machine.
#pragma once

#ifdef __CUDACC__
#define KMPC_W __device__ __forceinline__
#define KMPC_WN __device__
namespace kmpc {
KMPC_W int w_lane() { return (int)(threadIdx.x & 31u); }
KMPC_W int w_warp() { return (int)(threadIdx.x >> 5); }   // warp index inside the block
KMPC_W int w_warps() { return (int)(blockDim.x >> 5); }  // warps per block
KMPC_W int w_block() { return (int)blockIdx.x; }
KMPC_W double w_down(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
KMPC_W double w_up(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
KMPC_W double w_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
KMPC_W double w_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
KMPC_W int w_bcast_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
KMPC_W bool w_all(bool p) { return __all_sync(0xffffffffu, p); }
KMPC_W bool w_any(bool p) { return __any_sync(0xffffffffu, p); }
KMPC_W void w_sync() { __syncwarp(); }
KMPC_W unsigned w_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
KMPC_W void w_reconverge(unsigned mask) { __syncwarp(mask); }  // re-join the lanes of `mask` after a divergent region
KMPC_W void w_block_sync() { __syncthreads(); }
KMPC_W bool w_block_any(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }
KMPC_W unsigned w_smem_or(unsigned *p, unsigned v) { return atomicOr(p, v); }   // shared-memory word
KMPC_W int w_block_warps_with(bool p) { return __syncthreads_count((p && (threadIdx.x & 31u) == 0) ? 1 : 0); }   // block barrier + number of warps whose (warp-uniform) p holds
#ifndef KMPC_SERIAL_WARP_B
#define KMPC_SERIAL_WARP_B 2
#endif
KMPC_W int w_serial_warp(int W) { return (blockIdx.x >= (gridDim.x + 1) / 2 && W > KMPC_SERIAL_WARP_B) ? KMPC_SERIAL_WARP_B : 0; }
KMPC_W int w_fetch(int *queue) {  // next instance index for this warp
    int b = 0;
    if ((threadIdx.x & 31u) == 0) b = atomicAdd(queue, 1);
    return __shfl_sync(0xffffffffu, b, 0);
}
KMPC_W void w_count_trips(unsigned long long *total, int trips) { if (total) atomicAdd(total, (unsigned long long)trips); }
KMPC_W int w_take_slot(int *counter) { return atomicAdd(counter, 1); }   // global-memory counter
// max / min over the warp of doubles whose sign bit is clear (non-negative numbers, +inf, NaN with a clear sign bit): their
// bit patterns order like unsigned integers, so two 32-bit warp reductions (CREDUX) replace a 5-level shuffle butterfly.
// A NaN (pattern above +inf) wins the max, i.e. it propagates, exactly like w_maxabs_nan.
KMPC_W double w_max_nn(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
KMPC_W double w_min_nn(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
}  // namespace kmpc
#else
#define KMPC_W inline
#define KMPC_WN
#include "simt.h"  // tests/host_emul/simt.h (include path set by the test build only)
#endif

namespace kmpc {
KMPC_W int w_popc(unsigned m) {
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}
// position of the j-th (0-based) CLEAR bit of m among bits 0 .. W-1; W if there is none
KMPC_W int w_nth_clear(unsigned m, int W, int j) {
    for (int i = 0; i < W; ++i) if (!((m >> i) & 1u) && j-- == 0) return i;
    return W;
}
// butterfly reductions: every lane ends with the same bits
KMPC_W double w_sum(double v) { for (int m = 16; m > 0; m >>= 1) v += w_xor(v, m); return v; }
KMPC_W double w_min(double v) { for (int m = 16; m > 0; m >>= 1) { const double o = w_xor(v, m); v = o < v ? o : v; } return v; }
KMPC_W double w_max(double v) { for (int m = 16; m > 0; m >>= 1) { const double o = w_xor(v, m); v = o > v ? o : v; } return v; }
// max of |.| that propagates NaN (mirrors maxabs_nan of the thread solver)
KMPC_W double w_maxabs_nan(double v) {
    for (int m = 16; m > 0; m >>= 1) { const double o = w_xor(v, m); v = (o > v || o != o) ? o : v; }
    return v;
}
}  // namespace kmpc
