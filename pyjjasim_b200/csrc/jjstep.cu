// libjjstep.so - C ABI (include/jjstep.h) and the STREAMING step engine for sm_100a.
//
// Streaming engine: all state lives in HBM as problem-minor (rows x Wp) float64 arrays
// (Wp = W rounded up to 4); one kernel per phase of a time step:
//   k_step   fused "finish step n-1 / start step n" per junction  (reference: time_evolution.py:533-558,570-580)
//   k_face   b = A (x/c0 - theta_s) - 2 pi f per face              (reference: time_evolution.py:560-569)
//   k_solve  one launch per level of the compiled solve program    (reference: time_evolution.py:506,562-569)
// It handles every input form (dense per-step tables, dense voltage sources) and any circuit size; the subdomain
// engine (jj_subdomain.cu) is the fast path for everything else.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "jj_host.h"

using namespace jj;

static std::string g_create_err;

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

namespace jj {
int dev_alloc(JJHandle* h, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return JJ_OK;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        return JJ_ENOMEM;
    }
    h->device_bytes += (long long)bytes;
    return JJ_OK;
}
void dev_free(JJHandle* h, void* p, size_t bytes) {
    if (p) { cudaFree(p); h->device_bytes -= (long long)bytes; }
}
}  // namespace jj

template <typename T>
static int upload(JJHandle* h, T** dst, const T* src, size_t n) {
    int rc = dev_alloc(h, (void**)dst, n * sizeof(T));
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return JJ_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct StepArgs {
    int Nj, Nf, Wp, G;                // G = Wp / 4 problem groups
    const int* junc_face; const signed char* junc_sign;
    const double *Ic, *c0, *c1, *c2;
    Cpr cpr;
    Source Is, Vs, T;
    const double* noise; long long noise_i0; int noise_K;   // injected standard normals [K][Nj][Wp]
    unsigned long long seed; long long group_offset;
    double dt;
    const double* J;                  // [Nf][Wp] solve output (permuted faces)
    double* x;                        // [Nj][Wp]
    const double* th_in1;             // theta_{n-1}  (only when !do_post)
    const double* th_in2;             // theta_{n-2}
    double* th_out;                   // where theta_{n-1} is written when do_post
    double* thetas;                   // dense voltage-source phase, or null
    double* snap_th; double* snap_I;  // snapshot planes or null
    int* flag;
    long long n;                      // boundary index: finishes step n-1, starts step n
    int do_post, do_pre;
};

template <bool DEFAULT_CPR>
__global__ void __launch_bounds__(256) k_step(const StepArgs a) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.Nj * a.G) return;
    int j = (int)(idx / a.G);
    int g = (int)(idx % a.G);
    int w = 4 * g;
    size_t off = (size_t)j * a.Wp + w;
    double c0 = a.c0[j];
    double th1[4], th2[4];
    if (a.do_post) {
        double y[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int f = a.junc_face[2 * j + k];
            if (f >= 0) {
                double s = (double)a.junc_sign[2 * j + k];
                const double2* Jp = reinterpret_cast<const double2*>(a.J + (size_t)f * a.Wp + w);
                double2 j0 = Jp[0], j1 = Jp[1];
                y[0] += s * j0.x; y[1] += s * j0.y; y[2] += s * j1.x; y[3] += s * j1.y;
            }
        }
        const double2* xp = reinterpret_cast<const double2*>(a.x + off);
        double2 x0 = xp[0], x1 = xp[1];
        double xv[4] = {x0.x, x0.y, x1.x, x1.y};
        const double2* tp = reinterpret_cast<const double2*>(a.th_in2 + off);
        double2 t0 = tp[0], t1 = tp[1];
        th2[0] = t0.x; th2[1] = t0.y; th2[2] = t1.x; th2[3] = t1.y;
        bool bad = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            th1[q] = (y[q] - xv[q]) / c0;
            bad |= !isfinite(th1[q]);
        }
        if (bad) atomicOr(a.flag, 1);
        double2* op = reinterpret_cast<double2*>(a.th_out + off);
        op[0] = make_double2(th1[0], th1[1]);
        op[1] = make_double2(th1[2], th1[3]);
        if (a.snap_th) {
            double2* sp = reinterpret_cast<double2*>(a.snap_th + off);
            sp[0] = make_double2(th1[0], th1[1]);
            sp[1] = make_double2(th1[2], th1[3]);
        }
        if (a.snap_I) {
            double is[4];
            source_eval4(a.Is, a.n - 1, j, a.Nj, a.Wp, w, is);
            double2* sp = reinterpret_cast<double2*>(a.snap_I + off);
            sp[0] = make_double2(y[0] + is[0], y[1] + is[1]);
            sp[1] = make_double2(y[2] + is[2], y[3] + is[3]);
        }
        if (a.thetas && a.Vs.kind == KIND_DENSE) {
            double vs[4];
            source_eval4(a.Vs, a.n - 1, j, a.Nj, a.Wp, w, vs);
            double2* sp = reinterpret_cast<double2*>(a.thetas + off);
            double2 s0 = sp[0], s1 = sp[1];
            sp[0] = make_double2(s0.x + vs[0] * a.dt, s0.y + vs[1] * a.dt);
            sp[1] = make_double2(s1.x + vs[2] * a.dt, s1.y + vs[3] * a.dt);
        }
    } else {
        const double2* p1 = reinterpret_cast<const double2*>(a.th_in1 + off);
        const double2* p2 = reinterpret_cast<const double2*>(a.th_in2 + off);
        double2 u0 = p1[0], u1 = p1[1], v0 = p2[0], v1 = p2[1];
        th1[0] = u0.x; th1[1] = u0.y; th1[2] = u1.x; th1[3] = u1.y;
        th2[0] = v0.x; th2[1] = v0.y; th2[2] = v1.x; th2[3] = v1.y;
    }
    if (!a.do_pre) return;
    double Ic = a.Ic[j], c1 = a.c1[j], c2 = a.c2[j];
    double fl[4] = {0, 0, 0, 0};
    if (a.T.kind != KIND_ZERO) {
        double z[4], amp[4];
        if (a.noise_K > 0) {
            const double2* zp = reinterpret_cast<const double2*>(
                a.noise + ((size_t)(a.n - a.noise_i0) * a.Nj + j) * a.Wp + w);
            double2 z0 = zp[0], z1 = zp[1];
            z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
        } else {
            normal4(a.seed, j, a.group_offset + g, a.n, z);
        }
        source_eval4(a.T, a.n, j, a.Nj, a.Wp, w, amp);
#pragma unroll
        for (int q = 0; q < 4; ++q) fl[q] = amp[q] * z[q];
    }
    double is[4];
    source_eval4(a.Is, a.n, j, a.Nj, a.Wp, w, is);
    double xn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double X = Ic * cpr_eval<DEFAULT_CPR>(a.cpr, 2.0 * th1[q] - th2[q]) + c1 * th1[q] + c2 * th2[q];
        xn[q] = (fl[q] - is[q]) + X;
    }
    double2* xo = reinterpret_cast<double2*>(a.x + off);
    xo[0] = make_double2(xn[0], xn[1]);
    xo[1] = make_double2(xn[2], xn[3]);
}

struct FaceArgs {
    int Nj, Nf, Wp, G;
    const int* face_ptr; const int* face_junc; const signed char* face_sign;
    const double* c0;
    const double* x; const double* thetas;
    Source Vs, F;
    double* b;
    long long n;
};

__global__ void __launch_bounds__(256) k_face(const FaceArgs a) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.Nf * a.G) return;
    int f = (int)(idx / a.G);
    int w = 4 * (int)(idx % a.G);
    double acc[4] = {0, 0, 0, 0};
    for (int p = a.face_ptr[f]; p < a.face_ptr[f + 1]; ++p) {
        int j = a.face_junc[p];
        double s = (double)a.face_sign[p];
        double c0 = a.c0[j];
        const double2* xp = reinterpret_cast<const double2*>(a.x + (size_t)j * a.Wp + w);
        double2 x0 = xp[0], x1 = xp[1];
        double u[4] = {x0.x / c0, x0.y / c0, x1.x / c0, x1.y / c0};
        if (a.Vs.kind == KIND_RANK1) {
            double ts[4];
            source_eval4(a.Vs, a.n, j, a.Nj, a.Wp, w, ts);
#pragma unroll
            for (int q = 0; q < 4; ++q) u[q] -= ts[q];
        } else if (a.Vs.kind == KIND_DENSE) {
            const double2* tp = reinterpret_cast<const double2*>(a.thetas + (size_t)j * a.Wp + w);
            double2 t0 = tp[0], t1 = tp[1];
            u[0] -= t0.x; u[1] -= t0.y; u[2] -= t1.x; u[3] -= t1.y;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += s * u[q];
    }
    if (a.F.kind != KIND_ZERO) {
        double fv[4];
        source_eval4(a.F, a.n, f, a.Nf, a.Wp, w, fv);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] -= 6.283185307179586 * fv[q];
    }
    double2* bp = reinterpret_cast<double2*>(a.b + (size_t)f * a.Wp + w);
    bp[0] = make_double2(acc[0], acc[1]);
    bp[1] = make_double2(acc[2], acc[3]);
}

// One level of the solve program; a block = one group of tiles x 32 problems, a warp = 8 rows of a tile.
__global__ void __launch_bounds__(128) k_solve_level(const SweepView s, int group0, double* __restrict__ v, int Wp) {
    int g = group0 + blockIdx.x;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int w = blockIdx.y * 32 + lane;
    bool active = w < Wp;
    for (int t = s.group_ptr[g]; t < s.group_ptr[g + 1]; ++t) {
        int row0 = s.tile_row0[t], nrows = s.tile_nrows[t], m = s.tile_lpr[t], nsteps = s.tile_nsteps[t];
        int flags = s.tile_flags[t];
        const int* cols = s.cols + s.tile_col_off[t];
        const double* vals = s.vals + s.tile_val_off[t];
        int rbase = warp * 8;
        int nrp = 32 / m;
        double acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.0;
        if (rbase < nrows) {
            int ncols = nsteps * m;
            int mshift = 31 - __clz(m);
            for (int c = 0; c < ncols; ++c) {
                int col = cols[c];
                double sv = active ? v[(size_t)col * Wp + w] : 0.0;
                const double* vp = vals + (size_t)(c >> mshift) * 32 + (c & (m - 1));
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (rbase + r < nrp) acc[r] = fma(vp[(rbase + r) * m], sv, acc[r]);
            }
            if ((flags & 1) && active) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (rbase + r < nrows) acc[r] += v[(size_t)(row0 + rbase + r) * Wp + w];
            }
        }
        __syncthreads();
        if (rbase < nrows && active) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (rbase + r < nrows) v[(size_t)(row0 + rbase + r) * Wp + w] = acc[r];
        }
        __syncthreads();
    }
}

__global__ void k_debug_noise(int Nj, int Wp, int G, unsigned long long seed, long long group_offset, long long step,
                              double* out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Nj * G) return;
    int j = (int)(idx / G), g = (int)(idx % G);
    double z[4];
    normal4(seed, j, group_offset + g, step, z);
    for (int q = 0; q < 4; ++q) out[(size_t)j * Wp + 4 * g + q] = z[q];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static SweepView view_of(const SweepDev& s) {
    SweepView v;
    v.group_ptr = s.group_ptr_d; v.tile_row0 = s.tile_row0; v.tile_nrows = s.tile_nrows; v.tile_lpr = s.tile_lpr;
    v.tile_nsteps = s.tile_nsteps; v.tile_flags = s.tile_flags; v.tile_col_off = s.tile_col_off;
    v.tile_val_off = s.tile_val_off; v.cols = s.cols; v.vals = s.vals;
    return v;
}

static void free_sweep(JJHandle* h, SweepDev& s) {
    dev_free(h, s.group_ptr_d, (s.n_groups + 1) * sizeof(int));
    dev_free(h, s.tile_row0, s.n_tiles * sizeof(int)); dev_free(h, s.tile_nrows, s.n_tiles * sizeof(int));
    dev_free(h, s.tile_lpr, s.n_tiles * sizeof(int)); dev_free(h, s.tile_nsteps, s.n_tiles * sizeof(int));
    dev_free(h, s.tile_flags, s.n_tiles * sizeof(int));
    dev_free(h, s.tile_col_off, s.n_tiles * sizeof(long long)); dev_free(h, s.tile_val_off, s.n_tiles * sizeof(long long));
    dev_free(h, s.cols, s.n_cols * sizeof(int)); dev_free(h, s.vals, s.n_vals * sizeof(double));
    s = SweepDev();
}

static int upload_sweep(JJHandle* h, SweepDev& d, const JJSweep* s) {
    free_sweep(h, d);
    d.n_levels = s->n_levels; d.n_tiles = s->n_tiles; d.stage_rows = s->stage_rows;
    d.level_ptr.assign(s->level_ptr, s->level_ptr + s->n_levels + 1);
    d.n_groups = d.level_ptr.back();
    d.group_ptr.assign(s->group_ptr, s->group_ptr + d.n_groups + 1);
    REQUIRE(d.group_ptr.back() == s->n_tiles, JJ_EINVAL, "solve program: group_ptr does not cover all tiles");
    d.tile_row0_h.assign(s->tile_row0, s->tile_row0 + s->n_tiles);
    d.tile_nrows_h.assign(s->tile_nrows, s->tile_nrows + s->n_tiles);
    d.tile_flags_h.assign(s->tile_flags, s->tile_flags + s->n_tiles);
    d.n_cols = s->n_cols; d.n_vals = s->n_vals;
    int rc;
    if ((rc = upload(h, &d.group_ptr_d, d.group_ptr.data(), d.group_ptr.size()))) return rc;
    if ((rc = upload(h, &d.tile_row0, s->tile_row0, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.tile_nrows, s->tile_nrows, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.tile_lpr, s->tile_lpr, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.tile_nsteps, s->tile_nsteps, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.tile_flags, s->tile_flags, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.tile_col_off, (const long long*)s->tile_col_off, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.tile_val_off, (const long long*)s->tile_val_off, s->n_tiles))) return rc;
    if ((rc = upload(h, &d.cols, s->cols, (size_t)s->n_cols))) return rc;
    if ((rc = upload(h, &d.vals, s->vals, (size_t)s->n_vals))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return JJ_OK;
}

static int run_sweep(JJHandle* h, const SweepDev& s, double* v) {
    SweepView sv = view_of(s);
    for (int l = 0; l < s.n_levels; ++l) {
        int g0 = s.level_ptr[l], g1 = s.level_ptr[l + 1];
        if (g1 == g0) continue;
        dim3 grid(g1 - g0, (h->Wp + 31) / 32);
        k_solve_level<<<grid, 128, 0, h->stream>>>(sv, g0, v, h->Wp);
        h->launches++;
    }
    CK(cudaGetLastError());
    return JJ_OK;
}

extern "C" {

int jj_sm_count(int device) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    return sms;
}

int jj_create(int device, JJHandle** out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                       "); the time-evolution engine has no CPU fallback";
        return JJ_ECUDA;
    }
    if (device < 0 || device >= n) { g_create_err = "invalid device ordinal"; return JJ_EINVAL; }
    JJHandle* h = new JJHandle();
    h->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreate(&h->stream)) != cudaSuccess ||
        (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess ||
        (e = cudaMalloc((void**)&h->flag_d, sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(h->flag_d, 0, sizeof(int))) != cudaSuccess) {
        g_create_err = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
        delete h;
        return JJ_ECUDA;
    }
    *out = h;
    return JJ_OK;
}

const char* jj_last_error(const JJHandle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

static void free_source(JJHandle* h, SourceHost& s) {
    h->src_gen++;          // whatever the subdomain engine gathered from this input's base is stale
    dev_free(h, s.base_buf, s.N * sizeof(double));
    dev_free(h, s.table_buf, s.table_cap);
    s = SourceHost();
}

static void free_problem(JJHandle* h) {
    size_t nj = (size_t)h->cir.Nj * h->Wp * sizeof(double), nf = (size_t)h->cir.Nf * h->Wp * sizeof(double);
    dev_free(h, h->th1, nj); dev_free(h, h->th2, nj); dev_free(h, h->x, nj); dev_free(h, h->thetas, nj);
    dev_free(h, h->v, nf);
    h->th1 = h->th2 = h->x = h->thetas = h->v = nullptr;
    for (int i = 0; i < 4; ++i) free_source(h, h->src[i]);
    dev_free(h, h->noise_buf, h->noise_cap); h->noise_buf = nullptr; h->noise_cap = 0; h->noise_K = 0;
    dev_free(h, h->th_out, (size_t)h->th_cap_planes * nj); dev_free(h, h->I_out, (size_t)h->I_cap_planes * nj);
    h->th_out = h->I_out = nullptr; h->n_th_planes = h->n_I_planes = 0; h->th_cap_planes = h->I_cap_planes = 0;
    subdomain_free_problem(h);
    observe_free(h);
    h->have_problem = h->have_state = false;
}

static void free_circuit(JJHandle* h) {
    CircuitDev& c = h->cir;
    cudaFree(c.face_ptr); cudaFree(c.face_junc); cudaFree(c.face_sign); cudaFree(c.junc_face); cudaFree(c.junc_sign);
    cudaFree(c.Ic); cudaFree(c.c0); cudaFree(c.c1); cudaFree(c.c2);
    c = CircuitDev();
    h->have_circuit = false;
}

void jj_destroy(JJHandle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_problem(h);
    subdomain_drop_plan(h);
    free_sweep(h, h->fwd); free_sweep(h, h->bwd);
    free_circuit(h);
    cudaFree(h->flag_d);
    cudaFree(h->scratch);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
}

int jj_set_circuit(JJHandle* h, const JJCircuit* c) {
    CK(cudaSetDevice(h->device));
    REQUIRE(c && c->Nj > 0 && c->Nf >= 0, JJ_EINVAL, "circuit: bad sizes");
    REQUIRE(c->cpr_harmonics >= 1 && c->cpr_harmonics <= 16, JJ_EINVAL, "circuit: cpr_harmonics must be 1..16");
    if (h->have_problem) free_problem(h);
    subdomain_drop_plan(h);
    free_circuit(h);
    CircuitDev& d = h->cir;
    d.Nj = c->Nj; d.Nf = c->Nf;
    int nnz = c->face_ptr[c->Nf];
    int rc;
    if ((rc = upload(h, &d.face_ptr, c->face_ptr, (size_t)c->Nf + 1))) return rc;
    if ((rc = upload(h, &d.face_junc, c->face_junc, (size_t)nnz))) return rc;
    if ((rc = upload(h, &d.face_sign, (const signed char*)c->face_sign, (size_t)nnz))) return rc;
    if ((rc = upload(h, &d.junc_face, c->junc_face, (size_t)2 * c->Nj))) return rc;
    if ((rc = upload(h, &d.junc_sign, (const signed char*)c->junc_sign, (size_t)2 * c->Nj))) return rc;
    if ((rc = upload(h, &d.Ic, c->Ic, (size_t)c->Nj))) return rc;
    if ((rc = upload(h, &d.c0, c->c0, (size_t)c->Nj))) return rc;
    if ((rc = upload(h, &d.c1, c->c1, (size_t)c->Nj))) return rc;
    if ((rc = upload(h, &d.c2, c->c2, (size_t)c->Nj))) return rc;
    d.max_face_len = 0;
    for (int f = 0; f < c->Nf; ++f) d.max_face_len = std::max(d.max_face_len, c->face_ptr[f + 1] - c->face_ptr[f]);
    d.cpr.M = c->cpr_harmonics;
    for (int m = 0; m <= 16; ++m) {
        d.cpr.a[m] = m <= c->cpr_harmonics ? c->cpr_a[m] : 0.0;
        d.cpr.b[m] = m <= c->cpr_harmonics ? c->cpr_b[m] : 0.0;
    }
    d.default_cpr = (c->cpr_harmonics == 1 && c->cpr_a[0] == 0.0 && c->cpr_a[1] == 0.0 && c->cpr_b[1] == 1.0);
    CK(cudaStreamSynchronize(h->stream));
    h->have_circuit = true;
    return JJ_OK;
}

int jj_set_solver(JJHandle* h, const JJSweep* fwd, const JJSweep* bwd) {
    CK(cudaSetDevice(h->device));
    REQUIRE(fwd && bwd, JJ_EINVAL, "solver: null sweep");
    CK(cudaStreamSynchronize(h->stream));      // the program may be replaced between runs of a live problem
    int rc;
    if ((rc = upload_sweep(h, h->fwd, fwd))) return rc;
    if ((rc = upload_sweep(h, h->bwd, bwd))) return rc;
    h->have_solver = true;
    return JJ_OK;
}

int jj_set_problem(JJHandle* h, int32_t W, double dt, uint64_t seed, int64_t problem_offset, int32_t engine) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_circuit && h->have_solver, JJ_ESTATE, "set_problem: circuit and solver must be set first");
    REQUIRE(W > 0 && dt > 0, JJ_EINVAL, "set_problem: W and dt must be positive");
    REQUIRE(problem_offset % 4 == 0, JJ_EINVAL, "set_problem: problem_offset must be a multiple of 4");
    REQUIRE(engine == JJ_ENGINE_AUTO || engine == JJ_ENGINE_STREAMING || engine == JJ_ENGINE_SUBDOMAIN, JJ_EINVAL, "set_problem: unknown engine");
    // same problem count as before (annealing loops, repeated compute() calls on a cached engine): keep the state
    // arrays of all engines and only reset them; cudaMalloc/cudaFree of ~10 large arrays costs tens of milliseconds
    const bool same = h->have_problem && h->W == W && h->th1 && h->th2 && h->x && !h->thetas;
    size_t nj = (size_t)h->cir.Nj * ((W + 3) / 4 * 4) * sizeof(double), nf = (size_t)h->cir.Nf * ((W + 3) / 4 * 4) * sizeof(double);
    int rc;
    if (same) {
        for (int i = 0; i < 4; ++i) free_source(h, h->src[i]);
        dev_free(h, h->noise_buf, h->noise_cap); h->noise_buf = nullptr; h->noise_cap = 0; h->noise_K = 0;
        // the output planes stay allocated (jj_alloc_outputs reuses them when they are large enough): cudaFree and
        // cudaMalloc of ~100 MB cost tens of milliseconds per compute() call
        h->n_th_planes = h->n_I_planes = 0;
        observe_off(h);           // (same problem size: the accumulators fit the next observation)
    } else {
        free_problem(h);
    }
    h->W = W; h->Wp = (W + 3) / 4 * 4; h->dt = dt; h->seed = seed; h->problem_offset = problem_offset;
    h->engine_req = engine;
    if (!same) {
        if ((rc = dev_alloc(h, (void**)&h->th1, nj))) return rc;
        if ((rc = dev_alloc(h, (void**)&h->th2, nj))) return rc;
        if ((rc = dev_alloc(h, (void**)&h->x, nj))) return rc;
        if ((rc = dev_alloc(h, (void**)&h->v, nf))) return rc;
    }
    CK(cudaMemsetAsync(h->th1, 0, nj, h->stream));
    CK(cudaMemsetAsync(h->th2, 0, nj, h->stream));
    CK(cudaMemsetAsync(h->x, 0, nj, h->stream));
    if (nf) CK(cudaMemsetAsync(h->v, 0, nf, h->stream));
    CK(cudaMemsetAsync(h->flag_d, 0, sizeof(int), h->stream));
    for (int i = 0; i < 4; ++i) { h->src[i] = SourceHost(); h->src[i].dev.kind = KIND_ZERO; }
    h->steps_done = 0; h->launches = 0; h->non_finite = 0; h->last_ms = 0.0;
    h->engine = JJ_ENGINE_STREAMING;
    h->have_problem = true; h->have_state = true;   // zero initial conditions are valid
    return JJ_OK;
}

int jj_set_engine(JJHandle* h, int32_t engine) {
    REQUIRE(h->have_problem, JJ_ESTATE, "set_engine: problem not set");
    REQUIRE(engine == JJ_ENGINE_AUTO || engine == JJ_ENGINE_STREAMING || engine == JJ_ENGINE_SUBDOMAIN, JJ_EINVAL,
            "set_engine: unknown engine");
    h->engine_req = engine;
    return JJ_OK;
}

static int h2d_padded(JJHandle* h, double* dst, const double* src, size_t rows) {
    // (rows, W) host -> (rows, Wp) device; the pad columns keep their previous (zero) content
    CK(cudaMemcpy2DAsync(dst, (size_t)h->Wp * sizeof(double), src, (size_t)h->W * sizeof(double),
                         (size_t)h->W * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream));
    return JJ_OK;
}
static int d2h_padded(JJHandle* h, double* dst, const double* src, size_t rows) {
    CK(cudaMemcpy2DAsync(dst, (size_t)h->W * sizeof(double), src, (size_t)h->Wp * sizeof(double),
                         (size_t)h->W * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream));
    return JJ_OK;
}

int jj_set_state(JJHandle* h, const double* t1, const double* t2) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "set_state: problem not set");
    int rc;
    if ((rc = h2d_padded(h, h->th1, t1, h->cir.Nj))) return rc;
    if ((rc = h2d_padded(h, h->th2, t2, h->cir.Nj))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->have_state = true;
    return JJ_OK;
}

int jj_get_state(JJHandle* h, double* t1, double* t2) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "get_state: problem not set");
    int rc;     // both engines keep theta(-1), theta(-2) in the canonical arrays between runs
    if (t1 && (rc = d2h_padded(h, t1, h->th1, h->cir.Nj))) return rc;       // either pointer may be NULL
    if (t2 && (rc = d2h_padded(h, t2, h->th2, h->cir.Nj))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return JJ_OK;
}

int jj_set_source(JJHandle* h, int32_t which, int32_t kind, int32_t is_static, const double* base) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "set_source: problem not set");
    REQUIRE(which >= 0 && which < 4 && kind >= 0 && kind <= 2, JJ_EINVAL, "set_source: bad arguments");
    SourceHost& s = h->src[which];
    free_source(h, s);
    s.N = (which == JJ_SRC_F) ? h->cir.Nf : h->cir.Nj;
    s.dev.kind = kind; s.dev.is_static = is_static ? 1 : 0; s.dev.i0 = 0; s.dev.K = 0;
    if (kind == JJ_KIND_RANK1) {
        REQUIRE(base != nullptr, JJ_EINVAL, "set_source: RANK1 needs a base vector");
        int rc = upload(h, &s.base_buf, base, (size_t)s.N);
        if (rc) return rc;
        CK(cudaStreamSynchronize(h->stream));
        s.dev.base = s.base_buf;
    }
    if (which == JJ_SRC_VS && kind == JJ_KIND_DENSE && !h->thetas) {
        size_t nj = (size_t)h->cir.Nj * h->Wp * sizeof(double);
        int rc = dev_alloc(h, (void**)&h->thetas, nj);
        if (rc) return rc;
        CK(cudaMemsetAsync(h->thetas, 0, nj, h->stream));
    }
    return JJ_OK;
}

int jj_upload_source(JJHandle* h, int32_t which, int64_t i0, int32_t K, const double* table) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && which >= 0 && which < 4, JJ_ESTATE, "upload_source: bad state/arguments");
    SourceHost& s = h->src[which];
    REQUIRE(s.dev.kind != JJ_KIND_ZERO && K > 0 && table, JJ_EINVAL, "upload_source: source is ZERO or empty table");
    size_t rows = (s.dev.kind == JJ_KIND_RANK1) ? (size_t)K : (size_t)K * s.N;
    size_t bytes = rows * h->Wp * sizeof(double);
    if (bytes > s.table_cap) {
        CK(cudaStreamSynchronize(h->stream));
        dev_free(h, s.table_buf, s.table_cap);
        s.table_buf = nullptr; s.table_cap = 0;
        int rc = dev_alloc(h, (void**)&s.table_buf, bytes);
        if (rc) return rc;
        s.table_cap = bytes;
    }
    if (h->Wp != h->W) CK(cudaMemsetAsync(s.table_buf, 0, bytes, h->stream));
    int rc = h2d_padded(h, s.table_buf, table, rows);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    s.dev.table = s.table_buf; s.dev.i0 = i0; s.dev.K = K;
    return JJ_OK;
}

int jj_upload_noise(JJHandle* h, int64_t i0, int32_t K, const double* Z) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "upload_noise: problem not set");
    if (K <= 0) { h->noise_K = 0; return JJ_OK; }
    size_t rows = (size_t)K * h->cir.Nj, bytes = rows * h->Wp * sizeof(double);
    if (bytes > h->noise_cap) {
        CK(cudaStreamSynchronize(h->stream));
        dev_free(h, h->noise_buf, h->noise_cap);
        h->noise_buf = nullptr; h->noise_cap = 0;
        int rc = dev_alloc(h, (void**)&h->noise_buf, bytes);
        if (rc) return rc;
        h->noise_cap = bytes;
    }
    if (h->Wp != h->W) CK(cudaMemsetAsync(h->noise_buf, 0, bytes, h->stream));
    int rc = h2d_padded(h, h->noise_buf, Z, rows);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->noise_i0 = i0; h->noise_K = K;
    return JJ_OK;
}

int jj_alloc_outputs(JJHandle* h, int64_t n_th, int64_t n_I) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && n_th >= 0 && n_I >= 0, JJ_ESTATE, "alloc_outputs: problem not set");
    size_t nj = (size_t)h->cir.Nj * h->Wp * sizeof(double);
    int rc;
    if (n_th > h->th_cap_planes || n_I > h->I_cap_planes) CK(cudaStreamSynchronize(h->stream));
    if (n_th > h->th_cap_planes) {
        dev_free(h, h->th_out, (size_t)h->th_cap_planes * nj);
        h->th_out = nullptr; h->th_cap_planes = 0; h->n_th_planes = 0;
        if ((rc = dev_alloc(h, (void**)&h->th_out, (size_t)n_th * nj))) return rc;
        h->th_cap_planes = n_th;
    }
    h->n_th_planes = n_th;
    if (n_I > h->I_cap_planes) {
        dev_free(h, h->I_out, (size_t)h->I_cap_planes * nj);
        h->I_out = nullptr; h->I_cap_planes = 0; h->n_I_planes = 0;
        if ((rc = dev_alloc(h, (void**)&h->I_out, (size_t)n_I * nj))) return rc;
        h->I_cap_planes = n_I;
    }
    h->n_I_planes = n_I;
    return JJ_OK;
}

static int check_sources(JJHandle* h, long long i0, int n) {
    for (int i = 0; i < 4; ++i) {
        const Source& s = h->src[i].dev;
        if (s.kind == KIND_ZERO) continue;
        REQUIRE(s.table != nullptr, JJ_ESTATE, "run: a non-zero source has no table uploaded");
        // the post half of the boundary kernel at i0 + n still reads Is / Vs of step i0 + n - 1
        if (!s.is_static)
            REQUIRE(s.i0 <= i0 && i0 + n <= s.i0 + s.K, JJ_ESTATE, "run: source table does not cover the requested steps");
    }
    if (h->noise_K > 0)
        REQUIRE(h->noise_i0 <= i0 && i0 + n <= h->noise_i0 + h->noise_K, JJ_ESTATE,
                "run: injected noise does not cover the requested steps");
    return JJ_OK;
}

static int streaming_run(JJHandle* h, long long i0, int n, const long long* th_plane, const long long* I_plane) {
    const CircuitDev& c = h->cir;
    int Wp = h->Wp, G = Wp / 4;
    size_t plane = (size_t)c.Nj * Wp;
    StepArgs a;
    a.Nj = c.Nj; a.Nf = c.Nf; a.Wp = Wp; a.G = G;
    a.junc_face = c.junc_face; a.junc_sign = c.junc_sign;
    a.Ic = c.Ic; a.c0 = c.c0; a.c1 = c.c1; a.c2 = c.c2; a.cpr = c.cpr;
    a.Is = h->src[JJ_SRC_IS].dev; a.Vs = h->src[JJ_SRC_VS].dev; a.T = h->src[JJ_SRC_T].dev;
    a.noise = h->noise_buf; a.noise_i0 = h->noise_i0; a.noise_K = h->noise_K;
    a.seed = h->seed; a.group_offset = h->problem_offset / 4; a.dt = h->dt;
    a.J = h->v; a.x = h->x; a.thetas = h->thetas; a.flag = h->flag_d;
    FaceArgs fa;
    fa.Nj = c.Nj; fa.Nf = c.Nf; fa.Wp = Wp; fa.G = G;
    fa.face_ptr = c.face_ptr; fa.face_junc = c.face_junc; fa.face_sign = c.face_sign; fa.c0 = c.c0;
    fa.x = h->x; fa.thetas = h->thetas; fa.Vs = h->src[JJ_SRC_VS].dev; fa.F = h->src[JJ_SRC_F].dev; fa.b = h->v;
    long long tot = (long long)c.Nj * G, ftot = (long long)c.Nf * G;
    int sblocks = (int)((tot + 255) / 256), fblocks = (int)((ftot + 255) / 256);
    for (long long k = 0; k <= n; ++k) {
        long long nn = i0 + k;
        a.n = nn; a.do_post = k > 0; a.do_pre = k < n;
        a.snap_th = a.snap_I = nullptr;
        if (k > 0) {
            if (th_plane && th_plane[k - 1] >= 0) {
                REQUIRE(th_plane[k - 1] < h->n_th_planes, JJ_EINVAL, "run: theta plane index out of range");
                a.snap_th = h->th_out + (size_t)th_plane[k - 1] * plane;
            }
            if (I_plane && I_plane[k - 1] >= 0) {
                REQUIRE(I_plane[k - 1] < h->n_I_planes, JJ_EINVAL, "run: current plane index out of range");
                a.snap_I = h->I_out + (size_t)I_plane[k - 1] * plane;
            }
        }
        if (k == 0) { a.th_in1 = h->th1; a.th_in2 = h->th2; a.th_out = nullptr; }
        else if (k < n) { a.th_in1 = nullptr; a.th_in2 = h->th1; a.th_out = h->th1; }
        else { a.th_in1 = nullptr; a.th_in2 = h->th1; a.th_out = h->th2; }
        if (c.default_cpr) k_step<true><<<sblocks, 256, 0, h->stream>>>(a);
        else k_step<false><<<sblocks, 256, 0, h->stream>>>(a);
        h->launches++;
        if (k > 0 && observed_step(h, nn - 1)) {
            int rc = observe_streaming(h, nn - 1, a.th_out);
            if (rc) return rc;
        }
        if (k == n) break;
        if (c.Nf > 0) {
            fa.n = nn;
            k_face<<<fblocks, 256, 0, h->stream>>>(fa);
            h->launches++;
            int rc;
            if ((rc = run_sweep(h, h->fwd, h->v))) return rc;
            if ((rc = run_sweep(h, h->bwd, h->v))) return rc;
        }
    }
    CK(cudaGetLastError());
    // the final post-only kernel left theta_{last-1} in the th1 buffer and wrote theta_last to the th2 buffer
    std::swap(h->th1, h->th2);
    return JJ_OK;
}

}  // extern "C" (run_enqueue has C++ linkage)

// Enqueue steps [i0, i0 + n) on the handle's stream without waiting for them (jj_run adds the timing events, the wait
// and the non-finite check; jj_anneal enqueues a whole schedule of intervals back to back and waits once).
int jj::run_enqueue(JJHandle* h, long long i0, int n, const long long* th_plane, const long long* I_plane) {
    int rc = check_sources(h, i0, n);
    if (rc) return rc;
    // engine choice
    int want = h->engine_req;
    std::string why;
    if (want == JJ_ENGINE_AUTO)
        want = (!h->thetas && subdomain_supported(h, why)) ? JJ_ENGINE_SUBDOMAIN : JJ_ENGINE_STREAMING;
    if (want == JJ_ENGINE_SUBDOMAIN) {
        if (!subdomain_supported(h, why)) { h->err = "subdomain engine not applicable: " + why; return JJ_EINVAL; }
        if (!subdomain_prepared(h)) { if ((rc = subdomain_prepare(h))) return rc; }
    }
    if (want == JJ_ENGINE_SUBDOMAIN && h->thetas) {
        h->err = "the subdomain engine does not support dense voltage sources";
        return JJ_EINVAL;
    }
    h->engine = want;
    if (want == JJ_ENGINE_SUBDOMAIN) rc = subdomain_run(h, i0, n, th_plane, I_plane);
    else {
        REQUIRE(h->cir.Nf == 0 || h->fwd.n_levels > 0, JJ_ESTATE,
                "run: the streaming engine needs a solve program (jj_set_solver was given empty sweeps)");
        rc = streaming_run(h, i0, n, th_plane, I_plane);
    }
    if (rc) return rc;
    if (want == JJ_ENGINE_SUBDOMAIN) h->obs_count += observations_in(h, i0, n);
    h->steps_done += n;
    return JJ_OK;
}

extern "C" {

int jj_run(JJHandle* h, int64_t i0, int32_t n, const int64_t* th_plane, const int64_t* I_plane) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && h->have_state, JJ_ESTATE, "run: problem/state not set");
    REQUIRE(n >= 0, JJ_EINVAL, "run: negative step count");
    if (n == 0) return JJ_OK;
    CK(cudaEventRecord(h->ev0, h->stream));
    int rc = run_enqueue(h, i0, n, (const long long*)th_plane, (const long long*)I_plane);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    int flag = 0;
    CK(cudaMemcpy(&flag, h->flag_d, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        h->non_finite = 1;
        h->err = "non-finite phase encountered during time evolution";
        return JJ_ENONFINITE;
    }
    return JJ_OK;
}

static int fetch_planes(JJHandle* h, const double* base, long long have, long long p0, long long np, double* dst) {
    REQUIRE(p0 >= 0 && np >= 0 && p0 + np <= have, JJ_EINVAL, "fetch: plane range out of bounds");
    if (np == 0) return JJ_OK;
    size_t plane = (size_t)h->cir.Nj * h->Wp;
    int rc = d2h_padded(h, dst, base + (size_t)p0 * plane, (size_t)np * h->cir.Nj);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return JJ_OK;
}

int jj_fetch_theta(JJHandle* h, int64_t p0, int64_t np, double* dst) {
    CK(cudaSetDevice(h->device));
    return fetch_planes(h, h->th_out, h->n_th_planes, p0, np, dst);
}
int jj_fetch_current(JJHandle* h, int64_t p0, int64_t np, double* dst) {
    CK(cudaSetDevice(h->device));
    return fetch_planes(h, h->I_out, h->n_I_planes, p0, np, dst);
}

int jj_debug_noise(JJHandle* h, int64_t step, double* dst) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem, JJ_ESTATE, "debug_noise: problem not set");
    double* tmp = nullptr;
    size_t bytes = (size_t)h->cir.Nj * h->Wp * sizeof(double);
    int rc = dev_alloc(h, (void**)&tmp, bytes);
    if (rc) return rc;
    long long tot = (long long)h->cir.Nj * (h->Wp / 4);
    k_debug_noise<<<(int)((tot + 255) / 256), 256, 0, h->stream>>>(h->cir.Nj, h->Wp, h->Wp / 4, h->seed,
                                                                 h->problem_offset / 4, step, tmp);
    h->launches++;
    rc = d2h_padded(h, dst, tmp, h->cir.Nj);
    cudaStreamSynchronize(h->stream);
    dev_free(h, tmp, bytes);
    return rc;
}

int jj_debug_solve(JJHandle* h, const double* b, double* J) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && h->cir.Nf > 0, JJ_ESTATE, "debug_solve: problem not set");
    int rc;
    if ((rc = h2d_padded(h, h->v, b, h->cir.Nf))) return rc;
    if ((rc = run_sweep(h, h->fwd, h->v))) return rc;
    if ((rc = run_sweep(h, h->bwd, h->v))) return rc;
    if ((rc = d2h_padded(h, J, h->v, h->cir.Nf))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return JJ_OK;
}

int jj_set_subdomain_plan(JJHandle* h, const JJSubdomainPlan* plan) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_circuit, JJ_ESTATE, "set_subdomain_plan: the circuit must be set first");
    if (h->have_problem) free_problem(h);
    return subdomain_set_plan(h, plan);
}

int jj_debug_subdomain_solve(JJHandle* h, const double* b, double* J) {
    CK(cudaSetDevice(h->device));
    REQUIRE(h->have_problem && h->cir.Nf > 0, JJ_ESTATE, "debug_subdomain_solve: problem not set");
    double* tmp = nullptr;
    size_t bytes = (size_t)h->cir.Nf * h->Wp * sizeof(double);
    int rc = dev_alloc(h, (void**)&tmp, bytes);
    if (rc) return rc;
    CK(cudaMemsetAsync(tmp, 0, bytes, h->stream));
    CK(cudaMemsetAsync(h->v, 0, bytes, h->stream));
    if ((rc = h2d_padded(h, h->v, b, h->cir.Nf)) == 0 && (rc = subdomain_debug_solve(h, h->v, tmp)) == 0)
        rc = d2h_padded(h, J, tmp, h->cir.Nf);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    dev_free(h, tmp, bytes);
    if (rc == 0 && e != cudaSuccess) { h->err = std::string("debug_subdomain_solve: ") + cudaGetErrorString(e); rc = JJ_ECUDA; }
    return rc;
}

int jj_stats(JJHandle* h, JJStats* out) {
    memset(out, 0, sizeof(*out));
    out->engine = h->engine;
    out->steps_done = h->steps_done;
    out->kernel_launches = h->launches;
    out->step_ms = h->last_ms;
    out->device_bytes = h->device_bytes;
    out->non_finite = h->non_finite;
    out->cluster_size = 1; out->tile_problems = h->Wp;
    if (h->engine == JJ_ENGINE_SUBDOMAIN) { int c, w; subdomain_get_config(h, &c, &w); out->cluster_size = c; out->tile_problems = w; }
    return JJ_OK;
}

}  // extern "C"
