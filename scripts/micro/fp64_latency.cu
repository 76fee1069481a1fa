// Micro-benchmark: dependent-issue latency of FP64 instructions on one warp, and throughput vs. warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, int iters, double a, double b, long long *cyc) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_shfl(double *out, int iters, long long *cyc) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_rcp(double *out, int iters, long long *cyc) {
    double x = threadIdx.x + 1.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_lds(double *out, int iters, long long *cyc) {
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (double)((i * 7 + 1) & 1023);
    __syncthreads();
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x = s[(int)x & 1023];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const int it = 1000;
#define RUN(K, name, blocks, tpb, ...)                                                            \
    K<<<blocks, tpb>>>(__VA_ARGS__); cudaDeviceSynchronize(); K<<<blocks, tpb>>>(__VA_ARGS__);     \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                       \
    printf("%-28s blocks/SM-ish %4d tpb %4d : %.2f cycles per op-step\n", name, blocks, tpb, (double)h / (it * 16.0));
    RUN(chain<1>, "DFMA dep chain ILP1 1 warp", 1, 32, out, it, 0.999, 1e-9, cyc)
    RUN(chain<2>, "DFMA ILP2 1 warp", 1, 32, out, it, 0.999, 1e-9, cyc)
    RUN(chain<4>, "DFMA ILP4 1 warp", 1, 32, out, it, 0.999, 1e-9, cyc)
    RUN(chain<8>, "DFMA ILP8 1 warp", 1, 32, out, it, 0.999, 1e-9, cyc)
    RUN(chain<1>, "DFMA ILP1 4 warps (1/sched)", 1, 128, out, it, 0.999, 1e-9, cyc)
    RUN(chain<1>, "DFMA ILP1 8 warps (2/sched)", 1, 256, out, it, 0.999, 1e-9, cyc)
    RUN(chain<1>, "DFMA ILP1 16 warps (4/sched)", 1, 512, out, it, 0.999, 1e-9, cyc)
    RUN(chain<1>, "DFMA ILP1 32 warps (8/sched)", 1, 1024, out, it, 0.999, 1e-9, cyc)
    RUN(chain<2>, "DFMA ILP2 16 warps (4/sched)", 1, 512, out, it, 0.999, 1e-9, cyc)
    RUN(chain_shfl, "SHFL.64+DADD dep chain", 1, 32, out, it, cyc)
    RUN(chain_rcp, "MUFU.RCP64H dep chain", 1, 32, out, it, cyc)
    RUN(chain_lds, "LDS.64 dep chain (+cvt)", 1, 32, out, it, cyc)
    return 0;
}
