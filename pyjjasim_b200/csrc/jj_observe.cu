// libjjstep.so - observables computed on the device from the stored theta planes, and the zero-velocity
// restart, for the annealing caller of the stepping loop (reference: time_evolution.py:1142-1191):
//   vortex configuration  n = -A round(theta / 2 pi)                        (reference: time_evolution.py:734-755)
//   vortex mobility sums  sum_f sum_t |n_f(t+1) - n_f(t)| per problem        (reference: time_evolution.py:1128-1133)
//   restart at rest       theta(-2) := theta(-1)                            (reference: time_evolution.py:1169-1171)
// Everything here is integer arithmetic on rounded phases: results are exact, not approximate.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "jj_host.h"

using namespace jj;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return JJ_ECUDA;                                                                       \
        }                                                                                          \
    } while (0)

#define REQUIRE(cond, code, msg)                                                                   \
    do {                                                                                           \
        if (!(cond)) { h->err = (msg); return (code); }                                            \
    } while (0)

namespace {

constexpr int LANES = 32;     // problems per block (one 256-byte row segment per junction and warp)
constexpr int ROWS = 8;       // faces in flight per block

struct FaceView {
    int Nj, Nf, Wp;
    const int* face_ptr; const int* face_junc; const signed char* face_sign;
};

// vorticity of face f for problem w in one theta plane; the division is a true division and the rounding is
// to nearest-even, like np.round(theta / (2.0 * np.pi))
__device__ __forceinline__ int vorticity(const FaceView& c, const double* __restrict__ plane, int p0, int p1, int w) {
    double acc = 0.0;
    for (int p = p0; p < p1; ++p) {
        const double th = plane[(size_t)c.face_junc[p] * c.Wp + w];
        acc -= (double)c.face_sign[p] * phase_zone(th);
    }
    return (int)acc;
}

// dst_row: row of each (permuted) face in the output, or null for the permuted order itself - callers that want the
// circuit's own face numbering get it from the device instead of permuting hundreds of MB on the host
__global__ void __launch_bounds__(LANES * ROWS) k_vortex_configuration(const FaceView c, const double* __restrict__ plane,
                                                                        int* __restrict__ out, const int* __restrict__ dst_row) {
    const int w = blockIdx.x * LANES + threadIdx.x;
    const int f = blockIdx.y * ROWS + threadIdx.y;
    if (w >= c.Wp || f >= c.Nf) return;
    const int row = dst_row ? dst_row[f] : f;
    out[(size_t)row * c.Wp + w] = vorticity(c, plane, c.face_ptr[f], c.face_ptr[f + 1], w);
}

__global__ void k_permute_rows_i32(int Nf, int Wp, const int* __restrict__ src, int* __restrict__ dst, const int* __restrict__ dst_row) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Nf * Wp) return;
    const int f = (int)(idx / Wp), w = (int)(idx % Wp);
    dst[(size_t)dst_row[f] * Wp + w] = src[idx];
}

// out[w] += sum over this block's faces and over consecutive planes of |n(t+1) - n(t)|
// The kernel is bound by the latency of its gathers (one 256-byte row segment per junction, plane and warp): a face's
// junction list is read once, and the loads of all its junctions in two planes are issued before any is used.
constexpr int MOB_K = 8;      // junctions per face on the fast path (longer faces take the general loop)

template <int KU>
__device__ __forceinline__ int vorticity_k(const double* __restrict__ col, int Wp, const int (&j)[MOB_K], unsigned neg, int k_used) {
    double th[KU];
#pragma unroll
    for (int k = 0; k < KU; ++k) th[k] = k < k_used ? __ldg(col + (size_t)j[k] * Wp) : 0.0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        const double z = phase_zone(th[k]);            // (0 for the unused slots)
        acc += ((neg >> k) & 1u) ? z : -z;
    }
    return (int)acc;
}

__global__ void __launch_bounds__(LANES * ROWS) k_vortex_mobility(const FaceView c, const double* __restrict__ planes,
                                                                   long long n_planes, unsigned long long* __restrict__ out) {
    // block shape (bx, by), bx * by = 256: bx problems (up to 256: the warps of a block then read one contiguous
    // 2 KB row of a plane together, which HBM serves better than 256-byte pieces at different times) x by faces
    __shared__ unsigned long long part[LANES * ROWS];
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t plane = (size_t)c.Nj * c.Wp;
    unsigned long long acc = 0;
    if (w < c.Wp) {
        for (int f = blockIdx.y * blockDim.y + threadIdx.y; f < c.Nf; f += gridDim.y * blockDim.y) {
            const int p0 = c.face_ptr[f], p1 = c.face_ptr[f + 1];
            if (p1 == p0) continue;                       // (an empty row of the cycle matrix has no vorticity)
            if (p1 - p0 <= MOB_K) {
                int j[MOB_K]; unsigned neg = 0u;          // junction rows and the bits of the negative cycle-matrix entries
                const int ku = p1 - p0;
#pragma unroll
                for (int k = 0; k < MOB_K; ++k) {
                    const int p = min(p0 + k, p1 - 1);
                    j[k] = c.face_junc[p];
                    if (c.face_sign[p] < 0) neg |= 1u << k;
                }
                const double* col = planes + w;
                auto n_of = [&](const double* pl) { return ku <= 4 ? vorticity_k<4>(pl, c.Wp, j, neg, ku) : vorticity_k<MOB_K>(pl, c.Wp, j, neg, ku); };
                int prev = n_of(col);
                long long t = 1;
                for (; t + 1 < n_planes; t += 2) {        // two planes per round: twice the loads in flight
                    const double* pa = col + (size_t)t * plane;
                    const int ca = n_of(pa), cb = n_of(pa + plane);
                    acc += (unsigned long long)(abs(ca - prev) + abs(cb - ca));
                    prev = cb;
                }
                if (t < n_planes) acc += (unsigned long long)abs(n_of(col + (size_t)t * plane) - prev);
                continue;
            }
            int prev = vorticity(c, planes, p0, p1, w);
            for (long long t = 1; t < n_planes; ++t) {
                const int cur = vorticity(c, planes + (size_t)t * plane, p0, p1, w);
                acc += (unsigned long long)abs(cur - prev);
                prev = cur;
            }
        }
    }
    part[threadIdx.y * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && w < c.Wp) {
        unsigned long long s = 0;
        for (int r = 0; r < (int)blockDim.y; ++r) s += part[r * blockDim.x + threadIdx.x];
        if (s) atomicAdd(out + w, s);       // integer sum: independent of the order of arrival
    }
}

// launch shape of k_vortex_mobility: enough blocks to fill the machine, problem strips x face slices
static void mobility_shape(int Wp, int Nf, dim3& grid, dim3& block) {
    const int bx = std::min(LANES * ROWS, (Wp + 31) / 32 * 32), by = LANES * ROWS / bx;
    const int strips = (Wp + bx - 1) / bx;
    int slices = (Nf + by - 1) / by;
    slices = std::max(1, std::min(slices, (148 * 8 + strips - 1) / strips));
    grid = dim3(strips, slices); block = dim3(bx, by);
}

// The same sums from PHASE ZONE planes (jj_anneal on the subdomain engine): zones[t][j][w] is the low byte of
// round(theta / 2 pi). A thread takes four problems (one 32-bit word per junction) and works bytewise modulo 256:
// n_t = -A zones_t and d = n_t - n_(t-1) as signed bytes - exact while a face's vorticity moves by less than 128 per step.
__device__ __forceinline__ unsigned zone_vorticity4(const unsigned char* __restrict__ col, int Wp, const int (&j)[MOB_K], unsigned neg, int k_used) {
    unsigned z[MOB_K];
#pragma unroll
    for (int k = 0; k < MOB_K; ++k) z[k] = k < k_used ? __ldg(reinterpret_cast<const unsigned*>(col + (size_t)j[k] * Wp)) : 0u;
    unsigned n = 0u;
#pragma unroll
    for (int k = 0; k < MOB_K; ++k) n = ((neg >> k) & 1u) ? __vadd4(n, z[k]) : __vsub4(n, z[k]);
    return n;
}

__global__ void __launch_bounds__(LANES * ROWS) k_zone_mobility(const FaceView c, const unsigned char* __restrict__ zones,
                                                                 long long n_planes, unsigned long long* __restrict__ out) {
    __shared__ unsigned part[LANES * ROWS][4];
    const int w = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const size_t plane = (size_t)c.Nj * c.Wp;
    unsigned acc[4] = {0u, 0u, 0u, 0u};
    if (w < c.Wp) {
        for (int f = blockIdx.y * blockDim.y + threadIdx.y; f < c.Nf; f += gridDim.y * blockDim.y) {
            const int p0 = c.face_ptr[f], p1 = c.face_ptr[f + 1];
            if (p1 - p0 <= MOB_K) {
                int j[MOB_K]; unsigned neg = 0u;          // junction rows and the bits of the negative cycle-matrix entries
                const int ku = p1 - p0;
#pragma unroll
                for (int k = 0; k < MOB_K; ++k) {
                    const int p = min(p0 + k, max(p1 - 1, p0));
                    j[k] = ku > 0 ? c.face_junc[p] : 0;
                    if (ku > 0 && c.face_sign[p] < 0) neg |= 1u << k;
                }
                const unsigned char* col = zones + w;
                unsigned prev = zone_vorticity4(col, c.Wp, j, neg, ku);
                for (long long t = 1; t < n_planes; ++t) {
                    const unsigned cur = zone_vorticity4(col + (size_t)t * plane, c.Wp, j, neg, ku);
                    const unsigned d = __vabs4(__vsub4(cur, prev));
                    acc[0] += d & 0xffu; acc[1] += (d >> 8) & 0xffu; acc[2] += (d >> 16) & 0xffu; acc[3] += d >> 24;
                    prev = cur;
                }
            } else {
                unsigned prev = 0u;
                for (long long t = 0; t < n_planes; ++t) {
                    const unsigned char* col = zones + (size_t)t * plane + w;
                    unsigned cur = 0u;
                    for (int p = p0; p < p1; ++p) {
                        const unsigned z = __ldg(reinterpret_cast<const unsigned*>(col + (size_t)c.face_junc[p] * c.Wp));
                        cur = c.face_sign[p] < 0 ? __vadd4(cur, z) : __vsub4(cur, z);
                    }
                    if (t > 0) {
                        const unsigned d = __vabs4(__vsub4(cur, prev));
                        acc[0] += d & 0xffu; acc[1] += (d >> 8) & 0xffu; acc[2] += (d >> 16) & 0xffu; acc[3] += d >> 24;
                    }
                    prev = cur;
                }
            }
        }
    }
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) part[tid][e] = acc[e];
    __syncthreads();
    if (threadIdx.y == 0 && w < c.Wp) {
        for (int e = 0; e < 4; ++e) {
            unsigned long long s = 0;
            for (int r = 0; r < (int)blockDim.y; ++r) s += part[r * blockDim.x + threadIdx.x][e];
            if (s && w + e < c.Wp) atomicAdd(out + w + e, s);
        }
    }
}

// launch shape of k_zone_mobility: bx threads x 4 problems across, by faces down
static void zone_mobility_shape(int Wp, int Nf, dim3& grid, dim3& block) {
    const int words = (Wp + 3) / 4;
    const int bx = std::min(LANES * ROWS, (words + 31) / 32 * 32), by = LANES * ROWS / bx;
    const int strips = (words + bx - 1) / bx;
    int slices = (Nf + by - 1) / by;
    slices = std::max(1, std::min(slices, (148 * 8 + strips - 1) / strips));
    grid = dim3(strips, slices); block = dim3(bx, by);
}

// nsum[f][w] += n_f(theta): one observation of the streaming engine (the subdomain engine accumulates inside its step
// kernel, jj_subdomain.cu: vortex_pass)
__global__ void __launch_bounds__(LANES * ROWS) k_vortex_accumulate(const FaceView c, const double* __restrict__ plane,
                                                                     int* __restrict__ nsum) {
    const int w = blockIdx.x * LANES + threadIdx.x;
    const int f = blockIdx.y * ROWS + threadIdx.y;
    if (w >= c.Wp || f >= c.Nf) return;
    nsum[(size_t)f * c.Wp + w] += vorticity(c, plane, c.face_ptr[f], c.face_ptr[f + 1], w);
}

// ---- the annealing schedule's per-interval bookkeeping on the device (reference: time_evolution.py:1128-1140, 1164-1174)
// noise amplitudes of the next interval: sqrt(T) - or zero for everybody once every temperature is numerically zero,
// which is when the reference stops drawing noise (np.allclose(T, 0), time_evolution.py:512). One block.
__global__ void __launch_bounds__(256) k_anneal_amp(int W, int Wp, const double* __restrict__ T, double* __restrict__ amp) {
    int hot = 0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) hot |= fabs(T[w]) > 1.0e-8;
    hot = __syncthreads_or(hot);
    for (int w = threadIdx.x; w < Wp; w += blockDim.x) amp[w] = (hot && w < W) ? sqrt(T[w]) : 0.0;
}

// the temperature rule on the exact integer mobility sums of the interval just run: mobility = sum / norm, compared with
// this interval's target; T *= 1/T_factor above it, T *= T_factor otherwise (the reference's expression
// (m > u) * (1 / f) + (m <= u) * f evaluates to exactly one of the two factors); the new temperatures are the
// interval's row of the temperature profiles
__global__ void k_anneal_rule(int W, const unsigned long long* __restrict__ sums, double norm, double upper, double f,
                              double inv_f, double* __restrict__ T, double* __restrict__ profile) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const double m = (double)(long long)sums[w] / norm;
    const double factor = (m > upper ? inv_f : 0.0) + (m <= upper ? f : 0.0);
    const double t = T[w] * factor;
    T[w] = t;
    profile[w] = t;
}

// device scratch kept in the handle: an annealing schedule calls jj_vortex_mobility once per interval, and a
// cudaMalloc / cudaFree pair per call costs more than the kernel
int scratch(JJHandle* h, size_t bytes, void** out) {
    if (bytes > h->scratch_cap) {
        cudaStreamSynchronize(h->stream);
        dev_free(h, h->scratch, h->scratch_cap);
        h->scratch = nullptr; h->scratch_cap = 0;
        int rc = dev_alloc(h, &h->scratch, bytes);
        if (rc) return rc;
        h->scratch_cap = bytes;
    }
    *out = h->scratch;
    return JJ_OK;
}

FaceView view(const JJHandle* h) {
    FaceView c;
    c.Nj = h->cir.Nj; c.Nf = h->cir.Nf; c.Wp = h->Wp;
    c.face_ptr = h->cir.face_ptr; c.face_junc = h->cir.face_junc; c.face_sign = h->cir.face_sign;
    return c;
}

}  // namespace

namespace jj {

void observe_free(JJHandle* h) {
    dev_free(h, h->obs_nsum, h->obs_n_bytes);
    dev_free(h, h->obs_th_first, h->obs_th_bytes); dev_free(h, h->obs_th_last, h->obs_th_bytes);
    h->obs_nsum = nullptr; h->obs_th_first = h->obs_th_last = nullptr; h->obs_n_bytes = h->obs_th_bytes = 0;
    h->obs_interval = 0; h->obs_first = 0; h->obs_count = 0;
}

// observation off, accumulators kept: a cached engine that observes again (repeated compute() calls) does not pay
// cudaFree + cudaMalloc of its mark planes (2 GB on cfg3: 0.7 s per call)
void observe_off(JJHandle* h) { h->obs_interval = 0; h->obs_first = 0; h->obs_count = 0; }

bool observed_step(const JJHandle* h, long long step) {
    return h->obs_interval > 0 && step >= h->obs_first && (step - h->obs_first) % h->obs_interval == 0;
}

long long observations_in(const JJHandle* h, long long i0, long long n) {
    if (h->obs_interval <= 0 || n <= 0) return 0;
    const long long last = i0 + n - 1;
    if (last < h->obs_first) return 0;
    // m ranges over first + m * interval in [max(i0, first), last]
    const long long lo = i0 <= h->obs_first ? 0 : (i0 - h->obs_first + h->obs_interval - 1) / h->obs_interval;
    const long long hi = (last - h->obs_first) / h->obs_interval;
    return hi >= lo ? hi - lo + 1 : 0;
}

int observe_streaming(JJHandle* h, long long step, const double* theta) {
    const FaceView c = view(h);
    if (c.Nf > 0) {
        dim3 grid((c.Wp + LANES - 1) / LANES, (c.Nf + ROWS - 1) / ROWS), block(LANES, ROWS);
        k_vortex_accumulate<<<grid, block, 0, h->stream>>>(c, theta, h->obs_nsum);
        h->launches++;
    }
    if (step == h->obs_first) CK(cudaMemcpyAsync(h->obs_th_first, theta, h->obs_th_bytes, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->obs_th_last, theta, h->obs_th_bytes, cudaMemcpyDeviceToDevice, h->stream));
    h->obs_count++;
    return JJ_OK;
}

}  // namespace jj

extern "C" {

int jj_observe_begin(JJHandle* h, int64_t first_step, int32_t interval) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "observe_begin: problem not set");
    REQUIRE(interval >= 0, JJ_EINVAL, "observe_begin: negative interval");
    CK(cudaStreamSynchronize(h->stream));
    if (interval == 0) { observe_off(h); return JJ_OK; }
    REQUIRE(first_step >= h->steps_done, JJ_EINVAL, "observe_begin: first_step lies before the steps already run");
    const size_t nb = (size_t)std::max(h->cir.Nf, 1) * h->Wp * sizeof(int), tb = (size_t)h->cir.Nj * h->Wp * sizeof(double);
    if (nb != h->obs_n_bytes || tb != h->obs_th_bytes) {
        observe_free(h);
        int rc;
        if ((rc = dev_alloc(h, (void**)&h->obs_nsum, nb))) return rc;
        h->obs_n_bytes = nb;
        if ((rc = dev_alloc(h, (void**)&h->obs_th_first, tb))) return rc;
        h->obs_th_bytes = tb;
        if ((rc = dev_alloc(h, (void**)&h->obs_th_last, tb))) { h->obs_th_bytes = 0; return rc; }
    }
    CK(cudaMemsetAsync(h->obs_nsum, 0, nb, h->stream));
    CK(cudaMemsetAsync(h->obs_th_first, 0, tb, h->stream));
    CK(cudaMemsetAsync(h->obs_th_last, 0, tb, h->stream));
    h->obs_first = first_step; h->obs_interval = interval; h->obs_count = 0;
    return JJ_OK;
}

int jj_observe_fetch(JJHandle* h, int64_t* count, int32_t* nsum, double* theta_first, double* theta_latest,
                     const int32_t* face_order) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && h->obs_interval > 0, JJ_ESTATE, "observe_fetch: no observation in progress");
    if (count) *count = h->obs_count;
    const size_t wi = (size_t)h->W * sizeof(int), wpi = (size_t)h->Wp * sizeof(int);
    const size_t wd = (size_t)h->W * sizeof(double), wpd = (size_t)h->Wp * sizeof(double);
    if (nsum && h->cir.Nf > 0) {
        const int* src = h->obs_nsum;
        if (face_order) {
            int* buf = nullptr;
            const size_t pe = (size_t)h->cir.Nf * h->Wp;
            int rc = scratch(h, (pe + (size_t)h->cir.Nf) * sizeof(int), (void**)&buf);
            if (rc) return rc;
            CK(cudaMemcpyAsync(buf + pe, face_order, (size_t)h->cir.Nf * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            k_permute_rows_i32<<<(unsigned)((pe + 255) / 256), 256, 0, h->stream>>>(h->cir.Nf, h->Wp, h->obs_nsum, buf, buf + pe);
            h->launches++;
            src = buf;
        }
        CK(cudaMemcpy2DAsync(nsum, wi, src, wpi, wi, h->cir.Nf, cudaMemcpyDeviceToHost, h->stream));
    }
    if (theta_first) CK(cudaMemcpy2DAsync(theta_first, wd, h->obs_th_first, wpd, wd, h->cir.Nj, cudaMemcpyDeviceToHost, h->stream));
    if (theta_latest) CK(cudaMemcpy2DAsync(theta_latest, wd, h->obs_th_last, wpd, wd, h->cir.Nj, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JJ_OK;
}

int jj_vortex_configurations(JJHandle* h, int64_t plane0, int64_t n_planes, int32_t* dst, const int32_t* face_order) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "vortex_configurations: problem not set");
    REQUIRE(plane0 >= 0 && n_planes >= 0 && plane0 + n_planes <= h->n_th_planes, JJ_EINVAL,
            "vortex_configurations: theta plane range out of bounds");
    if (h->cir.Nf == 0 || n_planes == 0) return JJ_OK;
    const FaceView c = view(h);
    // two scratch planes: the kernel of plane p + 1 runs while plane p is copied out (+ the output order of the faces)
    int* buf = nullptr;
    const size_t pe = (size_t)c.Nf * c.Wp;
    int rc = scratch(h, (2 * pe + (size_t)c.Nf) * sizeof(int), (void**)&buf);
    if (rc) return rc;
    int* order_d = nullptr;
    if (face_order) {
        order_d = buf + 2 * pe;
        CK(cudaMemcpyAsync(order_d, face_order, (size_t)c.Nf * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    dim3 grid((c.Wp + LANES - 1) / LANES, (c.Nf + ROWS - 1) / ROWS), block(LANES, ROWS);
    const size_t wi = (size_t)h->W * sizeof(int), wpi = (size_t)c.Wp * sizeof(int);
    for (int64_t p = 0; p < n_planes; ++p) {
        int* b = buf + (size_t)(p & 1) * pe;
        k_vortex_configuration<<<grid, block, 0, h->stream>>>(c, h->th_out + (size_t)(plane0 + p) * c.Nj * c.Wp, b, order_d);
        h->launches++;
        CK(cudaMemcpy2DAsync(dst + (size_t)p * c.Nf * h->W, wi, b, wpi, wi, c.Nf, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return JJ_OK;
}

int jj_restart_at_rest(JJHandle* h) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && h->have_state, JJ_ESTATE, "restart_at_rest: problem/state not set");
    REQUIRE(!h->thetas, JJ_EINVAL, "restart_at_rest: not available with dense voltage sources");
    const size_t bytes = (size_t)h->cir.Nj * h->Wp * sizeof(double);
    CK(cudaMemcpyAsync(h->th2, h->th1, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return JJ_OK;
}

int jj_adopt_state_at_rest(JJHandle* h, JJHandle* from) {
    REQUIRE(from && from != h, JJ_EINVAL, "adopt_state_at_rest: needs another handle");
    CK(cudaSetDevice(h->device));
    REQUIRE(from->device == h->device, JJ_EINVAL, "adopt_state_at_rest: the handles live on different devices");
    REQUIRE(h->have_problem && from->have_problem && from->have_state, JJ_ESTATE, "adopt_state_at_rest: problem/state not set");
    REQUIRE(h->cir.Nj == from->cir.Nj && h->W == from->W && h->Wp == from->Wp, JJ_EINVAL,
            "adopt_state_at_rest: the handles differ in junction or problem count");
    REQUIRE(!h->thetas && !from->thetas, JJ_EINVAL, "adopt_state_at_rest: not available with dense voltage sources");
    CK(cudaStreamSynchronize(from->stream));          // (the state of `from` is final)
    const size_t bytes = (size_t)h->cir.Nj * h->Wp * sizeof(double);
    CK(cudaMemcpyAsync(h->th1, from->th1, bytes, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->th2, from->th1, bytes, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_state = true;
    return JJ_OK;
}

int jj_vortex_configuration(JJHandle* h, int64_t plane, int32_t* dst) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "vortex_configuration: problem not set");
    REQUIRE(plane >= -1 && plane < h->n_th_planes, JJ_EINVAL, "vortex_configuration: theta plane out of range");
    if (h->cir.Nf == 0) return JJ_OK;
    const FaceView c = view(h);
    const double* src = plane < 0 ? h->th1 : h->th_out + (size_t)plane * c.Nj * c.Wp;   // -1: the current state
    int* buf = nullptr;
    const size_t bytes = (size_t)c.Nf * c.Wp * sizeof(int);
    int rc = scratch(h, bytes, (void**)&buf);
    if (rc) return rc;
    dim3 grid((c.Wp + LANES - 1) / LANES, (c.Nf + ROWS - 1) / ROWS), block(LANES, ROWS);
    k_vortex_configuration<<<grid, block, 0, h->stream>>>(c, src, buf, nullptr);
    h->launches++;
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)h->W * sizeof(int), buf, (size_t)c.Wp * sizeof(int),
                                      (size_t)h->W * sizeof(int), c.Nf, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { h->err = std::string("vortex_configuration: ") + cudaGetErrorString(e); return JJ_ECUDA; }
    return JJ_OK;
}

int jj_vortex_mobility(JJHandle* h, int64_t plane0, int64_t n_planes, int64_t* dst) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "vortex_mobility: problem not set");
    REQUIRE(plane0 >= 0 && n_planes >= 0 && plane0 + n_planes <= h->n_th_planes, JJ_EINVAL,
            "vortex_mobility: theta plane range out of bounds");
    if (h->cir.Nf == 0 || n_planes < 2) { memset(dst, 0, (size_t)h->W * sizeof(int64_t)); return JJ_OK; }
    const FaceView c = view(h);
    unsigned long long* buf = nullptr;
    const size_t bytes = (size_t)c.Wp * sizeof(unsigned long long);
    int rc = scratch(h, bytes, (void**)&buf);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(buf, 0, bytes, h->stream);
    dim3 grid, block;
    mobility_shape(c.Wp, c.Nf, grid, block);
    k_vortex_mobility<<<grid, block, 0, h->stream>>>(c, h->th_out + (size_t)plane0 * c.Nj * c.Wp, n_planes, buf);
    h->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, buf, (size_t)h->W * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { h->err = std::string("vortex_mobility: ") + cudaGetErrorString(e); return JJ_ECUDA; }
    return JJ_OK;
}

int jj_anneal(JJHandle* h, int64_t first_interval, int32_t n_intervals, int32_t steps, const double* upper,
              double T_factor, double inv_T_factor, double norm, double* T, double* profiles, double* device_ms) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && h->have_state, JJ_ESTATE, "anneal: problem/state not set");
    REQUIRE(n_intervals >= 0 && steps >= 1 && upper && T && profiles, JJ_EINVAL, "anneal: bad arguments");
    REQUIRE(h->n_th_planes >= steps, JJ_ESTATE, "anneal: jj_alloc_outputs must provide one theta plane per interval step");
    const Source& ts = h->src[JJ_SRC_T].dev;
    REQUIRE(ts.kind == KIND_RANK1 && ts.is_static && h->src[JJ_SRC_T].table_buf, JJ_ESTATE,
            "anneal: the temperature must be declared as a static base x amplitude input with an uploaded amplitude row");
    REQUIRE(!h->thetas, JJ_EINVAL, "anneal: not available with dense voltage sources");
    if (device_ms) *device_ms = 0.0;
    if (n_intervals == 0) return JJ_OK;
    const int W = h->W, Wp = h->Wp;
    const FaceView c = view(h);
    double *Td = nullptr, *prof = nullptr;
    unsigned long long* sums = nullptr;
    const size_t tb = (size_t)Wp * sizeof(double), pb = (size_t)n_intervals * W * sizeof(double);
    int rc;
    if ((rc = dev_alloc(h, (void**)&Td, tb))) return rc;
    if ((rc = dev_alloc(h, (void**)&prof, pb))) { dev_free(h, Td, tb); return rc; }
    std::vector<long long> planes(steps);
    for (int k = 0; k < steps; ++k) planes[k] = k;
    // On the subdomain engine the interval's planes hold PHASE ZONES - the low byte of round(theta / 2 pi) per junction
    // and problem, written by the annealing variant of the step kernel into the theta plane buffer (an eighth of it) -
    // and k_zone_mobility takes the mobility from them: 4 bytes instead of 32 per junction and four problems leave the
    // step kernel, and the mobility kernel gathers bytes that are still in L2. JJ_ANNEAL_ZONES=0 (and the streaming
    // engine) store the phases themselves and run k_vortex_mobility on them.
    std::string why;
    const bool sub = h->engine_req == JJ_ENGINE_SUBDOMAIN || (h->engine_req == JJ_ENGINE_AUTO && subdomain_supported(h, why));
    const char* ze = getenv("JJ_ANNEAL_ZONES");
    bool zones = sub && !(ze && atoi(ze) == 0) && c.Nf > 0 && steps >= 2;
    if (zones && !subdomain_prepared(h) && (rc = subdomain_prepare(h))) { dev_free(h, Td, tb); dev_free(h, prof, pb); return rc; }
    zones = zones && subdomain_stores_zones(h);
    cudaError_t e = cudaMemsetAsync(Td, 0, tb, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Td, T, (size_t)W * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    dim3 mgrid, mblock, zgrid, zblock;
    mobility_shape(Wp, c.Nf, mgrid, mblock);
    zone_mobility_shape(Wp, c.Nf, zgrid, zblock);
    // the whole schedule is enqueued back to back; one wait, one timing, one non-finite check at the end
    if ((rc = scratch(h, (size_t)Wp * sizeof(unsigned long long), (void**)&sums))) { dev_free(h, Td, tb); dev_free(h, prof, pb); return rc; }
    if (e == cudaSuccess) e = cudaEventRecord(h->ev0, h->stream);
    rc = JJ_OK;
    for (int i = 0; i < n_intervals && rc == JJ_OK && e == cudaSuccess; ++i) {
        k_anneal_amp<<<1, 256, 0, h->stream>>>(W, Wp, Td, h->src[JJ_SRC_T].table_buf);
        h->launches++;
        // zero-velocity restart of every interval but the first (time_evolution.py:1169-1171): the subdomain engine's
        // first iteration reads theta(-2) from theta(-1) (it writes both back at the end of the run), the streaming
        // engine gets a device copy
        const bool restart = first_interval + i > 0;
        if (restart && !sub)
            e = cudaMemcpyAsync(h->th2, h->th1, (size_t)c.Nj * Wp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) break;
        h->zone8 = zones ? reinterpret_cast<unsigned char*>(h->th_out) : nullptr;
        h->start_at_rest = restart && sub;
        rc = run_enqueue(h, (first_interval + i) * (long long)steps, steps, planes.data(), nullptr);
        h->zone8 = nullptr; h->start_at_rest = false;
        if (rc) break;
        e = cudaMemsetAsync(sums, 0, (size_t)Wp * sizeof(unsigned long long), h->stream);
        if (zones) {
            k_zone_mobility<<<zgrid, zblock, 0, h->stream>>>(c, reinterpret_cast<const unsigned char*>(h->th_out), steps, sums);
            h->launches++;
        } else if (c.Nf > 0 && steps >= 2) {
            k_vortex_mobility<<<mgrid, mblock, 0, h->stream>>>(c, h->th_out, steps, sums);
            h->launches++;
        }
        k_anneal_rule<<<(W + 255) / 256, 256, 0, h->stream>>>(W, sums, norm, upper[i], T_factor, inv_T_factor, Td, prof + (size_t)i * W);
        h->launches++;
    }
    if (rc == JJ_OK && e == cudaSuccess) e = cudaEventRecord(h->ev1, h->stream);
    if (rc == JJ_OK && e == cudaSuccess) e = cudaMemcpyAsync(T, Td, (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (rc == JJ_OK && e == cudaSuccess) e = cudaMemcpyAsync(profiles, prof, pb, cudaMemcpyDeviceToHost, h->stream);
    int flag = 0;
    if (rc == JJ_OK && e == cudaSuccess) e = cudaMemcpyAsync(&flag, h->flag_d, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    dev_free(h, Td, tb); dev_free(h, prof, pb);
    if (rc) return rc;
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) { h->err = std::string("anneal: ") + cudaGetErrorString(e); return JJ_ECUDA; }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    if (device_ms) *device_ms = ms;
    if (flag) {
        h->non_finite = 1;
        h->err = "non-finite phase encountered during time evolution";
        return JJ_ENONFINITE;
    }
    return JJ_OK;
}

int jj_host_alloc(int device, uint64_t bytes, void** out) {
    if (!out) return JJ_EINVAL;
    *out = nullptr;
    if (bytes == 0) return JJ_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JJ_ECUDA;      // never create a context on another GPU by accident
    return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable) == cudaSuccess ? JJ_OK : JJ_ENOMEM;
}

int jj_host_free(void* p) {
    if (!p) return JJ_OK;
    return cudaFreeHost(p) == cudaSuccess ? JJ_OK : JJ_ECUDA;
}

}  // extern "C"
