// Microbenchmark: the cp.async staging ring of the top product without the MMAs (what does the data supply cost?)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
template <int S, bool COMPUTE>
__global__ void __launch_bounds__(512, 1) k(const double* src, double* out, long long* cyc, int iters, int pieces, size_t task_stride) {
    extern __shared__ double sm[];
    const double* base = src + (size_t)(blockIdx.x % 18) * task_stride;
    const int stage_doubles = pieces * 2;
    auto issue = [&](int kb) {
        if (kb < iters) {
            double* dst = sm + (size_t)(kb % S) * stage_doubles + threadIdx.x * 2;
            const double* s0 = base + (size_t)kb * stage_doubles + threadIdx.x * 2;
            if ((int)threadIdx.x < pieces) cp_async16(dst, s0);
            if ((int)threadIdx.x + 512 < pieces) cp_async16(dst + 1024, s0 + 1024);
        }
        asm volatile("cp.async.commit_group;");
    };
    __syncthreads();
    long long t0 = clock64();
    for (int k2 = 0; k2 < S - 1; ++k2) issue(k2);
    double acc = 0;
    for (int kb = 0; kb < iters; ++kb) {
        if (S == 4) asm volatile("cp.async.wait_group 2;");
        else if (S == 3) asm volatile("cp.async.wait_group 1;");
        else if (S == 8) asm volatile("cp.async.wait_group 6;");
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        issue(kb + S - 1);
        if (COMPUTE) acc += sm[(size_t)(kb % S) * stage_doubles + threadIdx.x % stage_doubles];
    }
    asm volatile("cp.async.wait_group 0;");
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 512 + threadIdx.x] = acc;
}
int main() {
    double *src, *out; long long* cyc;
    const size_t n = (size_t)64 << 20;     // 512 MB of doubles? no: 64M doubles = 512 MB
    cudaMalloc(&src, n * 8); cudaMemset(src, 0, n * 8);
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    long long h[148];
    const int iters = 56;
    auto run = [&](const char* name, auto kern, int pieces, int reps) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        const size_t task_stride = (size_t)iters * pieces * 2;
        for (int r = 0; r < reps; ++r) kern<<<144, 512, 100 * 1024>>>(src, out, cyc, iters, pieces, task_stride);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-46s pieces %4d (%5.1f KB/stage): %7.1f cycles/iteration (%s)\n", name, pieces, pieces * 16 / 1024.0,
               (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
    };
    run("S=4, L2-resident source (18 distinct tasks)", k<4, true>, 432, 3);
    run("S=4", k<4, true>, 576, 3);
    run("S=3", k<3, true>, 576, 3);
    run("S=8", k<8, true>, 432, 3);
    run("S=8", k<8, true>, 216, 3);
    run("S=4, no smem read", k<4, false>, 432, 3);
    run("S=2", k<2, true>, 432, 3);
    return 0;
}
