// Microbenchmark: larger FP64 MMA shapes on sm_100a (m16n8k4 / m16n8k8 / m16n8k16) versus m8n8k4.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* cyc, int iters, const double* in) {
    const int lane = threadIdx.x & 31;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = in[lane + 32 * i];
    for (int i = 0; i < 4; ++i) b[i] = in[256 + lane + 32 * i];
    double c0[4][4] = {};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (MODE == 0) { double (&cc)[2] = *reinterpret_cast<double (*)[2]>(&c0[ch][0]); mma884(cc, a[ch], b[ch & 3]); }
            if (MODE == 1) mma16816(c0[ch], a, b);
            if (MODE == 2) { double aa[2] = {a[ch], a[ch + 4]}; mma1684(c0[ch], aa, b[ch & 3]); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c0[i][j];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}
int main() {
    double *out, *in; long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    long long h[148]; const int iters = 2000;
    auto run = [&](const char* name, auto kern, double fma) {
        kern<<<148, 512>>>(out, cyc, iters, in); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double cyc_per = (double)h[0] / iters / 4;
        printf("%-16s %7.1f cycles per MMA per warp (16 warps/SM, 4 chains) -> %6.1f FMA/clk/SM (%s)\n", name, cyc_per, 16 * fma / cyc_per, cudaGetErrorString(cudaGetLastError()));
    };
    run("m8n8k4", k<0>, 256.0);
    run("m16n8k16", k<1>, 2048.0);
    run("m16n8k4", k<2>, 512.0);
    return 0;
}
