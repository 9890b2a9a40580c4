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
        acc -= (double)c.face_sign[p] * rint(th / 6.283185307179586);
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
__global__ void __launch_bounds__(LANES * ROWS) k_vortex_mobility(const FaceView c, const double* __restrict__ planes,
                                                                   long long n_planes, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long part[ROWS][LANES];
    const int w = blockIdx.x * LANES + threadIdx.x;
    const size_t plane = (size_t)c.Nj * c.Wp;
    unsigned long long acc = 0;
    if (w < c.Wp) {
        for (int f = blockIdx.y * ROWS + threadIdx.y; f < c.Nf; f += gridDim.y * ROWS) {
            const int p0 = c.face_ptr[f], p1 = c.face_ptr[f + 1];
            int prev = vorticity(c, planes, p0, p1, w);
            for (long long t = 1; t < n_planes; ++t) {
                const int cur = vorticity(c, planes + (size_t)t * plane, p0, p1, w);
                acc += (unsigned long long)abs(cur - prev);
                prev = cur;
            }
        }
    }
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && w < c.Wp) {
        unsigned long long s = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) s += part[r][threadIdx.x];
        if (s) atomicAdd(out + w, s);       // integer sum: independent of the order of arrival
    }
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
    if (interval == 0) { observe_free(h); return JJ_OK; }
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
    // enough blocks to fill the machine: problem strips x face slices
    const int strips = (c.Wp + LANES - 1) / LANES;
    int slices = (c.Nf + ROWS - 1) / ROWS;
    const int want = (148 * 8 + strips - 1) / strips;
    if (slices > want) slices = want;
    dim3 grid(strips, slices), block(LANES, ROWS);
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
    cudaError_t e = cudaMemsetAsync(Td, 0, tb, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Td, T, (size_t)W * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    const int strips = (Wp + LANES - 1) / LANES;
    int slices = (c.Nf + ROWS - 1) / ROWS;
    slices = std::max(1, std::min(slices, (148 * 8 + strips - 1) / strips));
    // the whole schedule is enqueued back to back; one wait, one timing, one non-finite check at the end
    if ((rc = scratch(h, (size_t)Wp * sizeof(unsigned long long), (void**)&sums))) { dev_free(h, Td, tb); dev_free(h, prof, pb); return rc; }
    if (e == cudaSuccess) e = cudaEventRecord(h->ev0, h->stream);
    rc = JJ_OK;
    for (int i = 0; i < n_intervals && rc == JJ_OK && e == cudaSuccess; ++i) {
        k_anneal_amp<<<1, 256, 0, h->stream>>>(W, Wp, Td, h->src[JJ_SRC_T].table_buf);
        h->launches++;
        if (first_interval + i > 0)      // zero-velocity restart of every interval but the first (time_evolution.py:1169-1171)
            e = cudaMemcpyAsync(h->th2, h->th1, (size_t)c.Nj * Wp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
        if (e != cudaSuccess) break;
        rc = run_enqueue(h, (first_interval + i) * (long long)steps, steps, planes.data(), nullptr);
        if (rc) break;
        e = cudaMemsetAsync(sums, 0, (size_t)Wp * sizeof(unsigned long long), h->stream);
        if (c.Nf > 0 && steps >= 2) {
            k_vortex_mobility<<<dim3(strips, slices), dim3(LANES, ROWS), 0, h->stream>>>(c, h->th_out, steps, sums);
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
