// Micro-benchmark: cycles of the serial Riccati phase (w_serial of kmpc_warp.cuh) for one warp with LANES active lanes,
// alone on an SM or next to NBG background warps running dependent FP64 chains.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../kiss_mpc_b200/csrc/kmpc_warp.cuh"
using namespace kmpc;
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__global__ void bench(Cfg c, int lanes, int reps, long long *cyc, double *sink, int mode) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int COOP = WLay<1>::COOP, NSTG = 32;
    for (int i = threadIdx.x; i < 16 * COOP; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    // plausible stage blocks: Q = 2, R = 1, small dynamics terms
    if (wid == 0 && lane < 16)
        for (int s = 0; s <= c.N; ++s) {
            double *q = sm + lane * COOP + s;
            q[C_A13 * NSTG] = -0.01; q[C_A23 * NSTG] = 0.02; q[C_B11 * NSTG] = 0.08; q[C_B21 * NSTG] = 0.06;
            q[C_Q00 * NSTG] = 2; q[C_Q11 * NSTG] = 2; q[C_Q22 * NSTG] = 1; q[C_DV * NSTG] = 1; q[C_DW * NSTG] = 1; q[C_HTV * NSTG] = 0.01;
            q[C_Q0 * NSTG] = 0.3; q[C_Q1 * NSTG] = -0.2; q[C_Q2 * NSTG] = 0.1; q[C_QV * NSTG] = 0.05; q[C_QW * NSTG] = 0.02;
            q[C_E0 * NSTG] = 0.001; q[C_E1 * NSTG] = 0.002; q[C_E2 * NSTG] = 0.0;
        }
    __syncthreads();
    double d0[3] = {0.1, 0.2, 0.3};
    if (wid == 0) {
        long long t0 = clock64();
        bool ok = true;
        for (int r = 0; r < reps; ++r) {
            if (lane < lanes) {
                // re-arm the inputs the sweep overwrote (cheap relative to the sweep; same for every variant)
                for (int s = 0; s <= c.N; ++s) {
                    double *q = sm + lane * COOP + s;
                    q[C_Q00 * NSTG] = 2; q[C_Q11 * NSTG] = 2; q[C_Q22 * NSTG] = 1; q[C_DV * NSTG] = 1; q[C_DW * NSTG] = 1; q[C_HTV * NSTG] = 0.01;
                    q[C_Q0 * NSTG] = 0.3; q[C_Q1 * NSTG] = -0.2; q[C_Q2 * NSTG] = 0.1; q[C_QV * NSTG] = 0.05; q[C_QW * NSTG] = 0.02;
                }
            }
            __syncwarp();
            long long a = clock64();
            if (lane < lanes) {
                if (mode == 0) ok = w_serial<false>(c, sm + lane * COOP, NSTG, d0) && ok;
            }
            __syncwarp();
            long long b = clock64();
            t0 += 0; if (lane == 0) cyc[r] = b - a;
        }
        if (lane == 0) sink[0] = ok ? sm[C_DX0 * NSTG + 5] + sm[C_PV0 * NSTG + 3] : -1.0;
    } else {
        // background: dependent FP64 chain (ILP 1) until warp 0 is done (fixed iteration count)
        double x = threadIdx.x;
        for (int i = 0; i < reps * 6000; ++i) x = fma(x, 0.999, 1e-9);
        sink[threadIdx.x] = x;
    }
}
// the whole sweep (w_serial) and the forward roll-out alone (w_serial_fwd), timed on the same warp
__global__ void bench3(Cfg c, int lanes, int reps, long long *cyc, double *sink) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    constexpr int COOP = WLay<1>::COOP, NSTG = 32;
    for (int i = threadIdx.x; i < 16 * COOP; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    double d0[3] = {0.1, 0.2, 0.3};
    bool ok = true;
    for (int r = 0; r < reps; ++r) {
        if (lane < 16)
            for (int s = 0; s <= c.N; ++s) {
                double *q = sm + lane * COOP + s;
                q[C_A13 * NSTG] = -0.01; q[C_A23 * NSTG] = 0.02; q[C_B11 * NSTG] = 0.08; q[C_B21 * NSTG] = 0.06;
                q[C_Q00 * NSTG] = 2; q[C_Q11 * NSTG] = 2; q[C_Q22 * NSTG] = 1; q[C_DV * NSTG] = 1; q[C_DW * NSTG] = 1; q[C_HTV * NSTG] = 0.01;
                q[C_Q0 * NSTG] = 0.3; q[C_Q1 * NSTG] = -0.2; q[C_Q2 * NSTG] = 0.1; q[C_QV * NSTG] = 0.05; q[C_QW * NSTG] = 0.02;
                q[C_E0 * NSTG] = 0.001; q[C_E1 * NSTG] = 0.002; q[C_E2 * NSTG] = 0.0;
            }
        __syncwarp();
        const long long t0 = clk();
        if (lane < lanes) ok = w_serial<false>(c, sm + lane * COOP, NSTG, d0) && ok;
        __syncwarp();
        const long long t1 = clk();
        const long long t2 = t1;
        if (lane < lanes) w_serial_fwd(c, sm + lane * COOP, NSTG, d0);
        __syncwarp();
        const long long t3 = clk();
        if (lane == 0) { cyc[r] = t1 - t0; cyc[32 + r] = t2 - t1; cyc[64 + r] = t3 - t2; }
    }
    if (lane == 0) sink[0] = ok ? sm[C_DX0 * NSTG + 5] + sm[C_PV0 * NSTG + 3] : -1.0;
}
int main() {
    {
        Cfg c; memset(&c, 0, sizeof c); c.N = 30; c.T = 0.1;
        long long *cyc; double *sink; cudaMalloc(&cyc, 8 * 128); cudaMalloc(&sink, 8 * 2048);
        const size_t smem = 16 * WLay<1>::COOP * 8;
        cudaFuncSetAttribute(bench3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        long long h[96];
        bench3<<<1, 32, smem>>>(c, 16, 8, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("one warp, 16 lanes: whole sweep %lld cycles, forward roll-out alone %lld cycles [%s]\n", h[5], h[64 + 5], cudaGetErrorString(e));
    }

    Cfg c; memset(&c, 0, sizeof c); c.N = 30; c.T = 0.1;
    long long *cyc; double *sink; cudaMalloc(&cyc, 8 * 64); cudaMalloc(&sink, 8 * 2048);
    const size_t smem = 16 * WLay<1>::COOP * 8;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode : {0}) for (int nbg : {0, 15}) for (int lanes : {16}) {
        const int reps = 8; long long h[8];
        bench<<<1, 32 * (1 + nbg), smem>>>(c, lanes, reps, cyc, sink, mode);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double sk; cudaMemcpy(&sk, sink, 8, cudaMemcpyDeviceToHost); printf("check %.12g  ", sk);
        printf("mode %d (0 full, 1 backward, 2 matrix part) background warps %2d lanes %d: %lld cycles per sweep (%.0f per stage)  [%s]\n", mode, nbg, lanes, h[5], h[5] / 31.0, cudaGetErrorString(e));
    }
    return 0;
}
