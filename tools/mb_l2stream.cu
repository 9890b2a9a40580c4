// Microbenchmark: how fast can ONE SM stream a slab that sits in L2 (1.2 MB per SM, the state of one cfg2 item), with
// the access forms the step kernel could use? 1 block of 512 threads per SM (200 KB of dynamic shared memory keep it
// alone), every block its own slab, all SMs at once.
//   ldg256 x U : per-thread 256-bit loads, U independent loads in flight per thread (the junction pass has 2-3)
//   bulk       : cp.async.bulk (TMA, 1-D) of 16 KB pieces into a ring of shared-memory buffers, one issuing thread
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_l2stream mb_l2stream.cu && ./mb_l2stream
#include <cstdio>
#include <cuda_runtime.h>

struct double4v { double2 lo, hi; };
__device__ __forceinline__ double4v ldg256(const double* p) {
    double4v v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.lo.x), "=d"(v.lo.y), "=d"(v.hi.x), "=d"(v.hi.y) : "l"(p));
    return v;
}

template <int U>
__global__ void __launch_bounds__(512, 1) k_ldg(const double* base, size_t slab_doubles, int reps, long long* cyc, double* sink) {
    extern __shared__ double smem[];
    const double* p = base + (size_t)blockIdx.x * slab_doubles;
    const int n = (int)(slab_doubles / 4);           // 32-byte pieces
    double acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < n; i += 512 * U) {
            double4v v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (i + u * 512 < n) v[u] = ldg256(p + (size_t)(i + u * 512) * 4);
#pragma unroll
            for (int u = 0; u < U; ++u) if (i + u * 512 < n) acc += v[u].lo.x + v[u].hi.y;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 1.2345e300) sink[0] = acc + smem[0];
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 ::"r"(b), "r"(parity) : "memory");
}

// ring of S buffers of PIECE bytes; thread 0 keeps S - 1 pieces in flight; all threads read each piece once from smem
template <int S, int PIECE>
__global__ void __launch_bounds__(512, 1) k_bulk(const double* base, size_t slab_doubles, int reps, long long* cyc, double* sink) {
    extern __shared__ __align__(128) double smem[];
    __shared__ unsigned long long bar[S];
    const char* p = reinterpret_cast<const char*>(base + (size_t)blockIdx.x * slab_doubles);
    const int n = (int)(slab_doubles * 8 / PIECE) * reps;       // pieces in all (the slab is walked reps times)
    const int per = (int)(slab_doubles * 8 / PIECE);
    if (threadIdx.x == 0) for (int s = 0; s < S; ++s) mbar_init(&bar[s], 1);
    __syncthreads();
    double acc = 0;
    const long long t0 = clock64();
    if (threadIdx.x == 0) for (int i = 0; i < S - 1 && i < n; ++i) bulk_load(reinterpret_cast<char*>(smem) + (size_t)i * PIECE, p + (size_t)(i % per) * PIECE, PIECE, &bar[i]);
    for (int i = 0; i < n; ++i) {
        const int s = i % S;
        mbar_wait(&bar[s], (i / S) & 1);
        const double2* q = reinterpret_cast<const double2*>(reinterpret_cast<char*>(smem) + (size_t)s * PIECE);
        for (int e = threadIdx.x; e < PIECE / 16; e += 512) { const double2 v = q[e]; acc += v.x + v.y; }
        __syncthreads();                               // buffer (i - 1) % S is free for piece i + S - 1
        const int j = i + S - 1;
        if (threadIdx.x == 0 && j < n) bulk_load(reinterpret_cast<char*>(smem) + (size_t)(j % S) * PIECE, p + (size_t)(j % per) * PIECE, PIECE, &bar[j % S]);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 1.2345e300) sink[0] = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t slab = 1200 * 1024 / 8 / 2048 * 2048;      // doubles per SM (~1.2 MB), a multiple of 16 KB
    double* buf; long long* cyc; double* sink;
    cudaMalloc(&buf, slab * 8 * sms); cudaMemset(buf, 0, slab * 8 * sms);
    cudaMalloc(&cyc, sms * sizeof(long long)); cudaMalloc(&sink, 8);
    const int smem = 200 * 1024, reps = 8;
    long long h[256];
    auto report = [&](const char* name, int blocks) {
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; long long mx = 0;
        for (int b = 0; b < blocks; ++b) { avg += h[b]; if (h[b] > mx) mx = h[b]; }
        avg /= blocks;
        printf("%-28s %3d SMs: %6.1f B/clk/SM (avg), %6.1f (slowest)   %s\n", name, blocks, slab * 8.0 * reps / avg, slab * 8.0 * reps / mx,
               cudaGetErrorString(cudaGetLastError()));
    };
    for (int blocks : {sms, 18}) {
        for (int pass = 0; pass < 2; ++pass) {       // first pass warms L2
#define RUN_LDG(U) cudaFuncSetAttribute(k_ldg<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_ldg<U><<<blocks, 512, smem>>>(buf, slab, reps, cyc, sink); if (pass) report("ldg256 x " #U, blocks);
            RUN_LDG(1) RUN_LDG(2) RUN_LDG(4) RUN_LDG(8)
#define RUN_BULK(S, P) cudaFuncSetAttribute(k_bulk<S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_bulk<S, P><<<blocks, 512, smem>>>(buf, slab, reps, cyc, sink); if (pass) report("bulk ring " #S " x " #P, blocks);
            RUN_BULK(2, 16384) RUN_BULK(4, 16384) RUN_BULK(8, 16384) RUN_BULK(4, 32768) RUN_BULK(4, 4096)
        }
    }
    return 0;
}
