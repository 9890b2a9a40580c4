// Micro-benchmarks that size the resident kernel's inner loops on the real part (latencies in SM cycles).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma_chain(long long* out, double* sink, int n) {
    double c0 = threadIdx.x, c1 = 1.0, a = 1e-9 * threadIdx.x, b = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) dmma884(c0, c1, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    sink[threadIdx.x] = c0 + c1;
}
__global__ void k_dmma_indep(long long* out, double* sink, int n) {
    double c[8][2]; for (int k = 0; k < 8; ++k) { c[k][0] = k; c[k][1] = threadIdx.x; }
    double a = 1e-9 * threadIdx.x, b = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    double s = 0; for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    sink[threadIdx.x + blockIdx.x * blockDim.x] = s;
}
__global__ void k_dfma_chain(long long* out, double* sink, int n) {
    double c = threadIdx.x, a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) c = fma(c, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    sink[threadIdx.x] = c;
}
__global__ void k_lds_dmma_chain(long long* out, double* sink, int n) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double c0 = 0, c1 = 0, a = 1e-9;
    int code = threadIdx.x * 7 % 4096;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { double b = sm[code]; dmma884(c0, c1, a, b); code = (code + 33) & 4095; }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    sink[threadIdx.x] = c0 + c1;
}
__global__ void k_ldg_chase(long long* out, const int* next, int n) {
    int p = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = __ldg(next + p);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0); out[1] = p; }
}
__global__ void k_prefetch(long long* out, const int* next, const int* chain, int mode) {
    // mode 0: cold chase; 1: prefetch.global.L1 of the chain first; 2: prefetch.global.L2; 3: plain loads first (warm)
    int acc = 0;
    for (int i = 0; i < 64; ++i) {
        const int* p = next + chain[i];
        if (mode == 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
        if (mode == 2) asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
        if (mode == 3) acc += __ldg(p);
    }
    long long t0 = clock64();
    while (clock64() - t0 < 20000) {}
    int p = chain[0] + (acc & 0);
    long long t1 = clock64();
    for (int i = 0; i < 64; ++i) p = __ldg(next + p);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t2 - t1; out[1] = p; }
}
__global__ void k_sync(long long* out, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
    long long *d, h[2]; double* sink; int* next;
    cudaMalloc(&d, 16); cudaMalloc(&sink, 1 << 20);
    const int N = 1 << 22;   // 16 MB pointer-chase array (L2 resident), stride large
    int* hn = new int[N]; for (int i = 0; i < N; ++i) hn[i] = (int)(((long long)i * 4099 + 12345) % N);
    cudaMalloc(&next, N * 4); cudaMemcpy(next, hn, N * 4, cudaMemcpyHostToDevice);
    int n = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        k_dmma_chain<<<1, 32>>>(d, sink, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DMMA m8n8k4 chain, 1 warp: %.1f cycles per DMMA\n", (double)h[0] / n);
        k_dmma_indep<<<1, 32>>>(d, sink, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("8 independent DMMA chains, 1 warp: %.1f cycles per DMMA\n", (double)h[0] / n / 8);
        k_dmma_indep<<<148, 512>>>(d, sink, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("8 independent DMMA chains, 16 warps/SM: %.1f cycles per DMMA per warp (=> %.1f FMA/clk/SM)\n", (double)h[0] / n / 8, 256.0 * 16 * n * 8 / h[0]);
        k_dfma_chain<<<1, 32>>>(d, sink, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DFMA chain: %.1f cycles per DFMA\n", (double)h[0] / n);
        k_lds_dmma_chain<<<1, 32>>>(d, sink, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("LDS + dependent DMMA chain: %.1f cycles per step\n", (double)h[0] / n);
        k_ldg_chase<<<1, 32>>>(d, next, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent __ldg chase (L2): %.1f cycles per load\n", (double)h[0] / n);
        for (int mode = 0; mode < 4; ++mode) {
            int hc[64]; int p = 777 + 100000 * (mode + 4 * rep);
            for (int i = 0; i < 64; ++i) { hc[i] = p; p = hn[p]; }
            int* dc; cudaMalloc(&dc, 256); cudaMemcpy(dc, hc, 256, cudaMemcpyHostToDevice);
            k_prefetch<<<1, 1>>>(d, next, dc, mode); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const char* nm[4] = {"nothing (cold)", "prefetch.global.L1", "prefetch.global.L2", "plain loads (warm L1/L2)"};
            if (rep) printf("dependent load latency after %s: %.1f cycles\n", nm[mode], h[0] / 64.0);
            cudaFree(dc);
        }
        k_sync<<<1, 512>>>(d, n); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("__syncthreads, 512 threads: %.1f cycles\n", (double)h[0] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
