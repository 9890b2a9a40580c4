// Microbenchmark: FP64 MMA (mma.m8n8k4) fed from shared memory, as in the staged top product.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: RBxGW pattern (active warps < RB, each 4 A + 16 B LDS + 16 MMA per iteration)
// mode 1: 16 warps, each 4 A + 4 B LDS + 4 MMA (two chains)
template <int MODE, bool SYNC>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* cyc, int iters, int RB) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += 512) sm[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[4][2] = {};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (SYNC) __syncthreads();
        const double* st = sm + (it & 3) * 1024;
        if (MODE == 0) {
            if (warp < RB) {
                const double* pa = st + warp * 4 * 32 + lane;
                const double* pb = st + 640 + (lane & 3) * 8 + (lane >> 2);
                double av[4];
#pragma unroll
                for (int kl = 0; kl < 4; ++kl) av[kl] = pa[kl * 32];
#pragma unroll
                for (int kl = 0; kl < 4; ++kl) {
                    double bv[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) bv[g] = pb[((g * 4 + kl) * 32) & 1023];
#pragma unroll
                    for (int g = 0; g < 4; ++g) dmma884(acc[g][0], acc[g][1], av[kl], bv[g]);
                }
            }
        } else {
            const double* pa = st + (warp & 3) * 4 * 32 + lane;
            const double* pb = st + 512 + (warp >> 2) * 4 * 32 + (lane & 3) * 8 + (lane >> 2);
            double av[4], bv[4];
#pragma unroll
            for (int kl = 0; kl < 4; ++kl) { av[kl] = pa[kl * 32]; bv[kl] = pb[kl * 32]; }
#pragma unroll
            for (int kl = 0; kl < 4; ++kl) dmma884(acc[kl & 1][0], acc[kl & 1][1], av[kl], bv[kl]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int g = 0; g < 4; ++g) s += acc[g][0] + acc[g][1];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    long long h[148];
    auto run = [&](const char* name, auto kern, int RB, int dm) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        kern<<<148, 512, 65536>>>(out, cyc, iters, RB);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-44s %7.1f cycles/iteration, %5.1f cycles per MMA of the busiest warp (%s)\n", name, (double)h[0] / iters,
               (double)h[0] / iters / dm, cudaGetErrorString(cudaGetLastError()));
    };
    run("5 warps x 16 MMA, barrier per iteration", k<0, true>, 5, 16);
    run("5 warps x 16 MMA, no barrier", k<0, false>, 5, 16);
    run("4 warps x 16 MMA, no barrier", k<0, false>, 4, 16);
    run("8 warps x 16 MMA, no barrier", k<0, false>, 8, 16);
    run("16 warps x 16 MMA, no barrier", k<0, false>, 16, 16);
    run("1 warp x 16 MMA, no barrier", k<0, false>, 1, 16);
    run("16 warps x 4 MMA, barrier per iteration", k<1, true>, 16, 4);
    run("16 warps x 4 MMA, no barrier", k<1, false>, 16, 4);
    return 0;
}
