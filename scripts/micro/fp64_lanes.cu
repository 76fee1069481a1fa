// Micro-benchmark: cost of FP64 FMA warp instructions as a function of the number of ACTIVE lanes (one warp, 8 independent
// chains per lane so that issue, not latency, limits).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_lanes.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int iters, int active, double a, double b, long long *cyc) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    if ((int)(threadIdx.x & 31) < active) {
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
        }
    }
    __syncwarp();
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int it = 2000;
    for (int warps : {1, 4}) for (int active : {1, 4, 8, 12, 16, 24, 32}) {
        k<<<1, 32 * warps>>>(out, it, active, 0.999, 1e-9, cyc); cudaDeviceSynchronize();
        k<<<1, 32 * warps>>>(out, it, active, 0.999, 1e-9, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %d (one per scheduler) active lanes %2d : %.2f cycles per DFMA warp-instruction\n", warps, active, (double)h / (it * 64.0));
    }
    return 0;
}
