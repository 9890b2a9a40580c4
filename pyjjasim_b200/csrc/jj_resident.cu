// RESIDENT step engine for sm_100a: one persistent kernel integrates a whole run.
//
// A thread-block cluster of C blocks owns a tile of WT problems for all time steps of a jj_run call.
// The elimination tree of the cycle-space system is cut at depth log2(C): block r keeps the
// right-hand-side rows of subtree r, plus replicas of the separators above the cut, in its shared
// memory (rows x WT float64, problem-minor). Per time step, with no global synchronisation at all:
//
//   junction pass   theta_n = (A^T J - x)/c0 from J in shared memory; x' = noise - Is + Ic cpr(..) + c1.. + c2..
//                   (reference: time_evolution.py:533-558, 570-580); theta, x stream through HBM once
//   face pass       b = A (x'/c0 - theta_s) - 2 pi f assembled into shared memory   (reference: :560-569)
//   forward sweep   local levels of the compiled solve program (warp per tile, lanes split rows x column
//                   groups, warp-shuffle reduction), then the replicated separators: partial sums per
//                   block, all-reduce over distributed shared memory (reduce-scatter + broadcast)
//   backward sweep  replicated separators redundantly, then local levels      (reference: :506, :562-569)
//
// Nothing but theta and x (and snapshots at stored steps) touches HBM: the right-hand side, the
// intermediate z and the cycle currents J never leave shared memory.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "jj_host.h"

namespace cg = cooperative_groups;
using namespace jj;

namespace {

constexpr int NT = 512;           // threads per block
constexpr int NWARPS = NT / 32;
constexpr int MAXC = 8;

struct RankProg {
    const int* wt_ptr;            // [n_levels * NWARPS + 1] -> tiles
    const int* ws_ptr;            // [n_levels * NWARPS + 1] -> stream steps
    const int2* thdr;             // packed tile headers
    const unsigned char* stream;  // [n_steps][320]
    int n_levels, n_tiles;
};

// The level cursors and tile headers of this rank's program are copied into shared memory once per
// kernel (they are the same for every tile of problems and every time step).
struct ProgSmem {
    const int* wt_ptr; const int* ws_ptr; const int2* thdr; const int4* ops;
};

struct ResArgs {
    // plan
    int C, n_rows, stage_rows, ar_rows, n_ops;
    const int4* ops;
    RankProg prog[MAXC];
    const int* junc_ptr; const int* junc_orig; const int2* junc_row; const char2* junc_sign;
    const int* face_ptr; const int* face_junc; const signed char* face_sign; const int* face_fidx;
    int face_K; const int* face_ell_j; const double* face_ell_c;   // [C][n_rows][K] device junction / sign/c0
    // circuit
    int Nj, Nf;
    const double *P0;              // [Nj'][4] = Ic, 1/c0, c1, c2 in device junction order
    const double *P1;              // [Nj'][4] = Is base, noise base, Vs base, 0
    Cpr cpr;
    // problem
    int Wp, n_tiles;
    double dt;
    unsigned long long seed; long long group_offset;
    Source Is, Vs, T, F;
    const double* noise; long long noise_i0; int noise_K;
    // state
    double* rth; double* rx;       // [tile][Nj][WT]
    double* th1; double* th2;      // canonical [Nj][Wp]: theta_{i0-1}, theta_{i0-2} in; theta_last, theta_{last-1} out
    // run
    long long i0; int n;
    const long long* th_plane; const long long* I_plane;
    double* snap_th; double* snap_I;   // canonical planes [plane][Nj][Wp]
    int* flag;
    // debug solve
    const double* dbg_b; double* dbg_J;   // canonical [Nf][Wp], permuted faces
    long long* dbg_prof;                  // per-block cycle counters [block][2 + n_ops] (JJ_RES_PROF), or null
    int dbg_skip;                         // timing experiments only (JJ_RES_DEBUG): 1 skip sweeps, 2 skip junction pass, 4 skip face pass
};

struct ResidentState {
    int C = 1, WT = 8, n_rows = 0, stage_rows = 0, ar_rows = 0, n_ops = 0;
    int4* ops = nullptr;
    RankProg prog[MAXC]{};
    std::vector<void*> allocs;
    std::vector<size_t> alloc_bytes;
    int *junc_ptr = nullptr, *junc_orig = nullptr; int2* junc_row = nullptr; char2* junc_sign = nullptr;
    int *face_ptr = nullptr, *face_junc = nullptr; signed char* face_sign = nullptr; int* face_fidx = nullptr;
    double *P0 = nullptr, *P1 = nullptr;
    int face_K = 0; int* face_ell_j = nullptr; double* face_ell_c = nullptr;
    // per problem
    double *rth = nullptr, *rx = nullptr; size_t state_bytes = 0;
    long long *plane_d = nullptr; size_t plane_cap = 0;
    int n_tiles = 0;
    size_t smem_bytes = 0, aux_ints = 0;
    int max_clusters = 0;
    bool prepared = false;
};

// ------------------------------------------------------------------------------------------------
// Shared-memory vector: row r holds WT float64 (problem-minor) as WT/2 16-byte chunks. Rows are 64 B
// (WT=8) or 32 B (WT=4) apart, so rows r and r+2 (r+4) would start in the same bank; lanes of a
// quarter-warp typically read the same chunk of different rows. The chunk index is therefore XOR-ed with
// a few row bits, which spreads eight consecutive rows over all eight 16-byte bank groups.
// ------------------------------------------------------------------------------------------------
template <int WT>
__device__ __forceinline__ int swz(int row) {
    return WT == 8 ? ((row >> 1) & 3) : ((row >> 2) & 1);
}
template <int WT>
__device__ __forceinline__ double2* chunk_ptr(double* v, int row, int chunk) {
    return reinterpret_cast<double2*>(v + (size_t)row * WT) + (chunk ^ swz<WT>(row));
}
template <int WT>
__device__ __forceinline__ const double2* chunk_ptr(const double* v, int row, int chunk) {
    return reinterpret_cast<const double2*>(v + (size_t)row * WT) + (chunk ^ swz<WT>(row));
}

constexpr int STEP_BYTES = 320;
#ifndef JJ_RING
#define JJ_RING 8
#endif
constexpr int RING = JJ_RING;     // stream steps kept in flight per warp

// Per-warp read position in the factor stream: tile range, step range and a register ring of RING
// prefetched steps (values + rows). The ring for the NEXT level is loaded before the barrier that ends the
// current one, so the L2 latency of the stream overlaps the barrier and the FMAs of earlier steps.
struct Cursor {
    int t0, t1, s, s_end;
    double ra[RING]; unsigned rc[RING];
};

__device__ __forceinline__ void cursor_open(Cursor& cu, const RankProg& p, const ProgSmem& ps, int level) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = level * NWARPS + warp;
    cu.t0 = ps.wt_ptr[idx]; cu.t1 = ps.wt_ptr[idx + 1];
    cu.s = ps.ws_ptr[idx]; cu.s_end = ps.ws_ptr[idx + 1];
    // ring slot of stream step s is s % RING, so a slot is refilled in place and never moved while in flight
    const int first = cu.s - (cu.s & (RING - 1));
#pragma unroll
    for (int k = 0; k < RING; ++k) {
        cu.ra[k] = 0.0; cu.rc[k] = 0;
        const int sk = first + k + ((first + k < cu.s) ? RING : 0);
        {   // unconditional: the stream buffer is padded by 2*RING steps, so reading past this warp's range is safe
            const unsigned char* rec = p.stream + (size_t)sk * STEP_BYTES;
            cu.ra[k] = __ldg(reinterpret_cast<const double*>(rec) + lane);
            cu.rc[k] = __ldg(reinterpret_cast<const unsigned short*>(rec + 256) + lane);
        }
    }
}

// C(8 rows x 8 problems) += A(8 x 4) . B(4 x 8) on the FP64 tensor core: lane = row*4 + kk holds A[row][kk],
// lane = n*4 + kk holds B[kk][n], lane = row*4 + c holds C[row][2c], C[row][2c+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One level of the solve program: every warp streams through its own tiles (8 output rows each).
template <int WT>
__device__ void exec_level(const RankProg& p, const ProgSmem& ps, Cursor& cu, int staged, int next_level,
                           double* __restrict__ v, double* __restrict__ stage) {
    static_assert(WT == 8, "the tensor-core sweep is written for 8 problems per tile");
    const int lane = threadIdx.x & 31;
    const int t0 = cu.t0, t1 = cu.t1;
    int s = cu.s;
    const unsigned char* base = p.stream;
    double ra[RING]; unsigned rc[RING];
#pragma unroll
    for (int k = 0; k < RING; ++k) { ra[k] = cu.ra[k]; rc[k] = cu.rc[k]; }
    const int r = lane >> 2, kk = lane & 3;
    for (int t = t0; t < t1; ++t) {
        const int2 hd = ps.thdr[t];
        const int row0 = hd.x & 0xffff, nrows = ((hd.x >> 16) & 7) + 1, flags = (hd.x >> 19) & 3;
        const int nsteps = hd.y & 0xffff, stage_off = (hd.y >> 16) & 0xffff;
        const int row = row0 + r;
        double2* self = chunk_ptr<WT>(v, row, kk);
        double c0 = 0.0, c1 = 0.0;
        if ((flags & 1) && r < nrows) { const double2 sv = *self; c0 = sv.x; c1 = sv.y; }
        // Ring slot K always holds a stream step congruent to K (mod RING): the MMA consumes it and the refill
        // for step s + RING is issued right behind it into the SAME registers. Loads are volatile inline PTX so
        // the compiler neither renames the slot (a copy-back would wait for the load every step) nor predicates
        // it through a temporary; the stream buffer is padded, so the refill is unconditional.
#define JJ_STEP(K)                                                                                     \
    if ((K) >= p_ && j < nsteps) {                                                                     \
        const double b_ = v[rc[K]];                                                                    \
        dmma884(c0, c1, ra[K], b_);                                                                    \
        const unsigned char* rec = base + (size_t)(s + RING) * STEP_BYTES;                             \
        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(ra[K]) : "l"(reinterpret_cast<const double*>(rec) + lane)); \
        asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(rc[K]) : "l"(reinterpret_cast<const unsigned short*>(rec + 256) + lane)); \
        ++s; ++j;                                                                                      \
    }
        for (int j = 0; j < nsteps;) {
            const int p_ = s & (RING - 1);
            JJ_STEP(0) JJ_STEP(1) JJ_STEP(2) JJ_STEP(3)
#if JJ_RING == 8
            JJ_STEP(4) JJ_STEP(5) JJ_STEP(6) JJ_STEP(7)
#endif
        }
#undef JJ_STEP
        __syncwarp();            // every lane has read its operands before rows of this block are overwritten
        if (r < nrows) {
            if (flags & 2) {
                // staged: the row keeps the physical layout of its destination and carries the destination index
                double2* srow = reinterpret_cast<double2*>(stage + (size_t)(stage_off + r) * (WT + 2));
                srow[kk ^ swz<WT>(row)] = make_double2(c0, c1);
                if (kk == 0) reinterpret_cast<int*>(srow + WT / 2)[0] = row;
            } else {
                *self = make_double2(c0, c1);
            }
        }
        __syncwarp();
    }
    if (next_level >= 0) cursor_open(cu, p, ps, next_level);
    __syncthreads();
    if (staged > 0) {
        // copy the staged rows (each carries its destination row) back into the vector
        for (int e = threadIdx.x; e < staged * (WT / 2); e += NT) {
            const int r = e / (WT / 2), q = e % (WT / 2);
            const double2* srow = reinterpret_cast<const double2*>(stage + (size_t)r * (WT + 2));
            const int dst_row = reinterpret_cast<const int*>(srow + WT / 2)[0];
            reinterpret_cast<double2*>(v + (size_t)dst_row * WT)[q] = srow[q];
        }
        __syncthreads();
    }
}

// all-reduce of rows [lo, hi) of every block's vector: reduce-scatter into mailboxes, sum in rank order,
// broadcast the sums back into every replica. Deterministic: the same block adds the partials in the same order.
// Rows are copied in their physical (swizzled) layout; row indices are identical on all ranks.
template <int WT>
__device__ void allreduce_rows(cg::cluster_group& cluster, int C, int rank, int lo, int hi, double* v, double* mbox) {
    const int n = hi - lo;
    int ch = (n + C - 1) / C;
    ch = (ch + 7) & ~7;           // chunk starts keep the row swizzle phase (multiple of 8 rows)
    for (int q = 0; q < C; ++q) {
        const int r0 = q * ch, r1 = min(n, r0 + ch);
        if (r1 <= r0) continue;
        double* dst = cluster.map_shared_rank(mbox, q) + (size_t)rank * ch * WT;
        const double* src = v + (size_t)(lo + r0) * WT;
        for (int e = threadIdx.x; e < (r1 - r0) * WT; e += NT) dst[e] = src[e];
    }
    cluster.sync();
    {
        const int r0 = rank * ch, r1 = min(n, r0 + ch);
        const int cnt = max(0, r1 - r0) * WT;
        for (int e = threadIdx.x; e < cnt; e += NT) {
            double sum = 0.0;
            for (int r = 0; r < C; ++r) sum += mbox[(size_t)r * ch * WT + e];
            for (int q = 0; q < C; ++q) cluster.map_shared_rank(v, q)[(size_t)(lo + r0) * WT + e] = sum;
        }
    }
    cluster.sync();
}

template <int WT, bool CL>
__device__ void run_ops(const ResArgs& a, const ProgSmem& ps, cg::cluster_group& cluster, int rank, double* v,
                        double* stage, double* mbox) {
    const RankProg& p = a.prog[rank];
    Cursor cu;
    bool open = false;
    long long tprev = a.dbg_prof ? clock64() : 0;
    for (int o = 0; o < a.n_ops; ++o) {
        if (a.dbg_prof && o > 0 && threadIdx.x == 0) {
            const long long tn = clock64();
            a.dbg_prof[(size_t)blockIdx.x * (2 + a.n_ops) + 2 + (o - 1)] += tn - tprev;
            tprev = tn;
        }
        const int4 op = ps.ops[o];
        if (op.x == 0) {
            if (!open) cursor_open(cu, p, ps, op.y);
            int next = -1;                       // the next level op, looking past an all-reduce
            for (int o2 = o + 1; o2 < a.n_ops && o2 <= o + 2; ++o2) {
                const int4 nx = ps.ops[o2];
                if (nx.x == 0) { next = nx.y; break; }
            }
            exec_level<WT>(p, ps, cu, op.z, next, v, stage);
            open = next >= 0;
        } else if (CL) {
            allreduce_rows<WT>(cluster, a.C, rank, op.y, op.z, v, mbox);
        }
    }
    if (a.dbg_prof && threadIdx.x == 0)
        a.dbg_prof[(size_t)blockIdx.x * (2 + a.n_ops) + 2 + (a.n_ops - 1)] += clock64() - tprev;
}

// Per-step amplitudes of the rank-one inputs for the 8 problems of this tile, cached in shared memory
// (double-buffered by step parity so that Is of the previous step is still there for current snapshots).
struct AmpCache {
    double T[2][8], Is[2][8], Vs[2][8], F[2][8];
};

__device__ __forceinline__ void amp_fill(const ResArgs& a, AmpCache* ac, int tile, long long n) {
    const int e = threadIdx.x & 7, which = threadIdx.x >> 3;
    if (threadIdx.x >= 32) return;
    const int w = tile * 8 + e;
    const Source& s = which == 0 ? a.T : which == 1 ? a.Is : which == 2 ? a.Vs : a.F;
    double val = 0.0;
    if (s.kind == KIND_RANK1 && w < a.Wp) val = __ldg(s.table + source_row(s, n) * a.Wp + w);
    double* dst = which == 0 ? ac->T[n & 1] : which == 1 ? ac->Is[n & 1] : which == 2 ? ac->Vs[n & 1] : ac->F[n & 1];
    dst[e] = val;
}

// State and constants of one junction item (one junction x 4 problems), loaded one iteration ahead.
struct JIn {
    double2 x0, x1, t0, t1, pIc, pc, pb;
    int2 rows;
    int sg, jo;
};

template <int WT>
__device__ __forceinline__ void jin_load(const ResArgs& a, int jlo, int tile, int idx, JIn& in) {
    constexpr int G = WT / 4;
    const int jp = jlo + idx / G;
    const size_t sidx = ((size_t)tile * a.Nj + jlo) * WT + (size_t)idx * 4;
    const double2* xp = reinterpret_cast<const double2*>(a.rx + sidx);
    const double2* tp = reinterpret_cast<const double2*>(a.rth + sidx);
    in.x0 = xp[0]; in.x1 = xp[1]; in.t0 = tp[0]; in.t1 = tp[1];
    in.pIc = __ldg(reinterpret_cast<const double2*>(a.P0) + 2 * jp);
    in.pc = __ldg(reinterpret_cast<const double2*>(a.P0) + 2 * jp + 1);
    in.pb = __ldg(reinterpret_cast<const double2*>(a.P1) + 2 * jp);
    in.rows = __ldg(a.junc_row + jp);
    const char2 sg = a.junc_sign[jp];
    in.sg = (int)sg.x * 4 + (int)sg.y;       // signs are -1, 0, +1
    in.jo = __ldg(a.junc_orig + jp);
}

// theta_n = (A^T J - x)/c0, snapshots, x' for the next step  (reference: time_evolution.py:533-558, 570-580)
template <int WT, bool DEF>
__device__ __forceinline__ void junction_item(const ResArgs& a, const AmpCache* ac, int tile, long long n, int idx,
                                              int jlo, bool do_pre, const JIn& in, const double* __restrict__ v,
                                              double* snap_th, double* snap_I) {
    constexpr int G = WT / 4;
    const int q = (idx % G) * 4;
    const int w = tile * WT + q;
    if (w >= a.Wp) return;
    const size_t sidx = ((size_t)tile * a.Nj + jlo) * WT + (size_t)idx * 4;
    double y[4] = {0, 0, 0, 0};
    if (in.rows.x >= 0) {
        const double2 j0 = *chunk_ptr<WT>(v, in.rows.x, q >> 1), j1 = *chunk_ptr<WT>(v, in.rows.x, (q >> 1) + 1);
        const double s = (double)((in.sg + 5) / 4 - 1);
        y[0] = s * j0.x; y[1] = s * j0.y; y[2] = s * j1.x; y[3] = s * j1.y;
    }
    if (in.rows.y >= 0) {
        const double2 j0 = *chunk_ptr<WT>(v, in.rows.y, q >> 1), j1 = *chunk_ptr<WT>(v, in.rows.y, (q >> 1) + 1);
        const double s = (double)((in.sg + 5) % 4 - 1);
        y[0] = fma(s, j0.x, y[0]); y[1] = fma(s, j0.y, y[1]); y[2] = fma(s, j1.x, y[2]); y[3] = fma(s, j1.y, y[3]);
    }
    const double ic0 = in.pIc.y;
    double th1[4] = {(y[0] - in.x0.x) * ic0, (y[1] - in.x0.y) * ic0, (y[2] - in.x1.x) * ic0, (y[3] - in.x1.y) * ic0};
    const double th2[4] = {in.t0.x, in.t0.y, in.t1.x, in.t1.y};
    // one finiteness test for the four phases (a NaN or Inf in any of them poisons the sum)
    if (!(fabs((th1[0] + th1[1]) + (th1[2] + th1[3])) < 1.0e300) && !a.dbg_skip) atomicOr(a.flag, 1);
    if (snap_th || snap_I || !do_pre) {
        const size_t cidx = (size_t)in.jo * a.Wp + w;
        if (snap_th) {
            double2* sp = reinterpret_cast<double2*>(snap_th + cidx);
            sp[0] = make_double2(th1[0], th1[1]); sp[1] = make_double2(th1[2], th1[3]);
        }
        if (snap_I) {
            const double* am = ac->Is[(n - 1) & 1] + q;
            double2* sp = reinterpret_cast<double2*>(snap_I + cidx);
            sp[0] = make_double2(y[0] + in.pb.x * am[0], y[1] + in.pb.x * am[1]);
            sp[1] = make_double2(y[2] + in.pb.x * am[2], y[3] + in.pb.x * am[3]);
        }
        if (!do_pre) {
            // end of the run: hand theta_last and theta_{last-1} back in the canonical layout
            double2* o1 = reinterpret_cast<double2*>(a.th1 + cidx);
            double2* o2 = reinterpret_cast<double2*>(a.th2 + cidx);
            o1[0] = make_double2(th1[0], th1[1]); o1[1] = make_double2(th1[2], th1[3]);
            o2[0] = in.t0; o2[1] = in.t1;
            return;
        }
    }
    double2* op = reinterpret_cast<double2*>(a.rth + sidx);
    op[0] = make_double2(th1[0], th1[1]); op[1] = make_double2(th1[2], th1[3]);
    const double Ic = in.pIc.x, c1 = in.pc.x, c2 = in.pc.y;
    double fl[4] = {0, 0, 0, 0};
    if (a.T.kind != KIND_ZERO && !(a.dbg_skip & 8)) {
        double z[4];
        if (a.noise_K > 0) {
            const double2* zp = reinterpret_cast<const double2*>(a.noise + ((size_t)(n - a.noise_i0) * a.Nj + in.jo) * a.Wp + w);
            const double2 z0 = zp[0], z1 = zp[1];
            z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
        } else {
            normal4(a.seed, in.jo, a.group_offset + (w >> 2), n, z);
        }
        const double* am = ac->T[n & 1] + q;
#pragma unroll
        for (int k = 0; k < 4; ++k) fl[k] = (in.pb.y * am[k]) * z[k];
    }
    const double* ia = ac->Is[n & 1] + q;
    double xn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double arg = 2.0 * th1[k] - th2[k];
        const double X = Ic * ((a.dbg_skip & 16) ? arg : cpr_eval<DEF>(a.cpr, arg)) + c1 * th1[k] + c2 * th2[k];
        xn[k] = (fl[k] - in.pb.x * ia[k]) + X;
    }
    double2* xo = reinterpret_cast<double2*>(a.rx + sidx);
    xo[0] = make_double2(xn[0], xn[1]); xo[1] = make_double2(xn[2], xn[3]);
}

template <int WT, bool DEF>
__device__ void junction_pass(const ResArgs& a, const AmpCache* ac, int rank, int tile, long long n, bool do_post,
                              bool do_pre, const double* __restrict__ v) {
    constexpr int G = WT / 4;
    const int jlo = a.junc_ptr[rank], jhi = a.junc_ptr[rank + 1];
    const int total = (jhi - jlo) * G;
    if (!do_post) {
        // first boundary of a run: theta(n-1), theta(n-2) come from the canonical arrays; no solve result yet
        for (int idx = threadIdx.x; idx < total; idx += NT) {
            const int jp = jlo + idx / G, q = (idx % G) * 4, w = tile * WT + q;
            if (w >= a.Wp) continue;
            const int jo = __ldg(a.junc_orig + jp);
            const size_t sidx = ((size_t)tile * a.Nj + jlo) * WT + (size_t)idx * 4;
            const size_t cidx = (size_t)jo * a.Wp + w;
            const double2* p1 = reinterpret_cast<const double2*>(a.th1 + cidx);
            const double2* p2 = reinterpret_cast<const double2*>(a.th2 + cidx);
            const double2 u0 = p1[0], u1 = p1[1], v0 = p2[0], v1 = p2[1];
            double2* op = reinterpret_cast<double2*>(a.rth + sidx);
            op[0] = u0; op[1] = u1;
            const double th1[4] = {u0.x, u0.y, u1.x, u1.y}, th2[4] = {v0.x, v0.y, v1.x, v1.y};
            const double2 pIc = __ldg(reinterpret_cast<const double2*>(a.P0) + 2 * jp);
            const double2 pc = __ldg(reinterpret_cast<const double2*>(a.P0) + 2 * jp + 1);
            const double2 pb = __ldg(reinterpret_cast<const double2*>(a.P1) + 2 * jp);
            double fl[4] = {0, 0, 0, 0};
            if (a.T.kind != KIND_ZERO) {
                double z[4];
                if (a.noise_K > 0) {
                    const double2* zp = reinterpret_cast<const double2*>(a.noise + ((size_t)(n - a.noise_i0) * a.Nj + jo) * a.Wp + w);
                    const double2 z0 = zp[0], z1 = zp[1];
                    z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
                } else {
                    normal4(a.seed, jo, a.group_offset + (w >> 2), n, z);
                }
                const double* am = ac->T[n & 1] + q;
                for (int k = 0; k < 4; ++k) fl[k] = (pb.y * am[k]) * z[k];
            }
            const double* ia = ac->Is[n & 1] + q;
            double xn[4];
            for (int k = 0; k < 4; ++k) {
                const double X = pIc.x * cpr_eval<DEF>(a.cpr, 2.0 * th1[k] - th2[k]) + pc.x * th1[k] + pc.y * th2[k];
                xn[k] = (fl[k] - pb.x * ia[k]) + X;
            }
            double2* xo = reinterpret_cast<double2*>(a.rx + sidx);
            xo[0] = make_double2(xn[0], xn[1]); xo[1] = make_double2(xn[2], xn[3]);
        }
        return;
    }
    double* snap_th = nullptr; double* snap_I = nullptr;
    {
        const long long k = n - 1 - a.i0;
        const long long pt = a.th_plane ? a.th_plane[k] : -1, pi = a.I_plane ? a.I_plane[k] : -1;
        if (pt >= 0) snap_th = a.snap_th + (size_t)pt * a.Nj * a.Wp;
        if (pi >= 0) snap_I = a.snap_I + (size_t)pi * a.Nj * a.Wp;
    }
    // software pipeline: the loads of the next item are in flight while the current one is computed
    JIn cur, nxt;
    int idx = threadIdx.x;
    if (idx < total) jin_load<WT>(a, jlo, tile, idx, cur);
    for (; idx < total; idx += NT) {
        const bool more = idx + NT < total;
        if (more) jin_load<WT>(a, jlo, tile, idx + NT, nxt);
        if (idx + 3 * NT < total) {      // and the state three iterations ahead is pulled into L2
            const size_t pf = ((size_t)tile * a.Nj + jlo) * WT + (size_t)(idx + 3 * NT) * 4;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rx + pf));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rth + pf));
        }
        junction_item<WT, DEF>(a, ac, tile, n, idx, jlo, do_pre, cur, v, snap_th, snap_I);
        if (more) cur = nxt;
    }
}

// b = A (x'/c0 - theta_s) - 2 pi f into shared memory (reference: time_evolution.py:560-569). Every row has a
// fixed-width list of (device junction, +-1/c0) pairs so all loads of a thread are independent.
template <int WT>
__device__ void face_pass(const ResArgs& a, const AmpCache* ac, int rank, int tile, long long n, double* __restrict__ v) {
    constexpr int G = WT / 4;
    const int K = a.face_K;
    const int* fj = a.face_ell_j + (size_t)rank * a.n_rows * K;
    const double* fc = a.face_ell_c + (size_t)rank * a.n_rows * K;
    const int* fidx = a.face_fidx + (size_t)rank * a.n_rows;
    const int total = a.n_rows * G;
    for (int idx = threadIdx.x; idx < total; idx += NT) {
        const int row = idx / G;
        const int q = (idx % G) * 4;
        const int w = tile * WT + q;
        double acc[4] = {0, 0, 0, 0};
        if (w < a.Wp) {
            const size_t tbase = (size_t)tile * a.Nj * WT + q;
            for (int k0 = 0; k0 < K; k0 += 4) {
                int jp[4]; double cf[4]; double2 xa[4], xb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    jp[k] = (k0 + k < K) ? __ldg(fj + (size_t)row * K + k0 + k) : -1;
                    cf[k] = (k0 + k < K) ? __ldg(fc + (size_t)row * K + k0 + k) : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (jp[k] >= 0) {
                        const double2* xp = reinterpret_cast<const double2*>(a.rx + tbase + (size_t)jp[k] * WT);
                        xa[k] = __ldcg(xp); xb[k] = __ldcg(xp + 1);
                    } else {
                        xa[k] = make_double2(0, 0); xb[k] = make_double2(0, 0);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[0] = fma(cf[k], xa[k].x, acc[0]); acc[1] = fma(cf[k], xa[k].y, acc[1]);
                    acc[2] = fma(cf[k], xb[k].x, acc[2]); acc[3] = fma(cf[k], xb[k].y, acc[3]);
                }
                if (a.Vs.kind == KIND_RANK1) {
                    const double* cum = ac->Vs[n & 1] + q;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (jp[k] >= 0) {
                            // coefficient = sign / c0: recover sign * Vs base from the per-junction records
                            const double ic0 = __ldg(a.P0 + 4 * (size_t)jp[k] + 1), vb = __ldg(a.P1 + 4 * (size_t)jp[k] + 2);
                            const double sv = (cf[k] / ic0) * vb;
                            for (int e = 0; e < 4; ++e) acc[e] -= sv * cum[e];
                        }
                    }
                }
            }
            const int g = __ldg(fidx + row);
            if (g >= 0 && a.F.kind == KIND_RANK1) {
                const double* am = ac->F[n & 1] + q;
                const double b = __ldg(a.F.base + g);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] -= 6.283185307179586 * (b * am[k]);
            }
        }
        *chunk_ptr<WT>(v, row, q >> 1) = make_double2(acc[0], acc[1]);
        *chunk_ptr<WT>(v, row, (q >> 1) + 1) = make_double2(acc[2], acc[3]);
    }
}

template <int WT, bool DEF, bool CL>
__global__ void __launch_bounds__(NT, 1) k_resident(const ResArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* v = smem;
    double* stage = v + (size_t)a.n_rows * WT;
    double* mbox = stage + (size_t)a.stage_rows * (WT + 2);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = CL ? (int)cluster.block_rank() : 0;
    ProgSmem ps;
    AmpCache* ac;
    {
        int* aux = reinterpret_cast<int*>(mbox + (size_t)(a.ar_rows + 8 * a.C) * WT);
        const RankProg& p = a.prog[rank];
        const int np = p.n_levels * NWARPS + 1;
        int* wt = aux; int* ws = aux + np; int2* th = reinterpret_cast<int2*>(aux + 2 * np);
        for (int e = threadIdx.x; e < np; e += NT) { wt[e] = p.wt_ptr[e]; ws[e] = p.ws_ptr[e]; }
        for (int e = threadIdx.x; e < p.n_tiles; e += NT) th[e] = p.thdr[e];
        int4* ops_s = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(th + p.n_tiles) + 15) & ~(uintptr_t)15);
        for (int e = threadIdx.x; e < a.n_ops; e += NT) ops_s[e] = a.ops[(size_t)rank * a.n_ops + e];
        ps.wt_ptr = wt; ps.ws_ptr = ws; ps.thdr = th; ps.ops = ops_s;
        ac = reinterpret_cast<AmpCache*>(ops_s + a.n_ops);
        __syncthreads();
    }
    const int cluster_id = blockIdx.x / a.C;
    const int n_clusters = gridDim.x / a.C;
    for (int tile = cluster_id; tile < a.n_tiles; tile += n_clusters) {
        if (a.dbg_b) {
            // debug: one solve of the compiled program on this tile's columns
            constexpr int G = WT / 4;
            const int* fidx = a.face_fidx + (size_t)rank * a.n_rows;
            for (int idx = threadIdx.x; idx < a.n_rows * G; idx += NT) {
                int row = idx / G, q = (idx % G) * 4, w = tile * WT + q, g = fidx[row];
                double t[4];
                for (int k = 0; k < 4; ++k) t[k] = (g >= 0 && w < a.Wp) ? a.dbg_b[(size_t)g * a.Wp + w + k] : 0.0;
                *chunk_ptr<WT>(v, row, q >> 1) = make_double2(t[0], t[1]);
                *chunk_ptr<WT>(v, row, (q >> 1) + 1) = make_double2(t[2], t[3]);
            }
            __syncthreads();
            run_ops<WT, CL>(a, ps, cluster, rank, v, stage, mbox);
            for (int idx = threadIdx.x; idx < a.n_rows * G; idx += NT) {
                int row = idx / G, q = (idx % G) * 4, w = tile * WT + q, g = fidx[row];
                if (g >= 0 && w < a.Wp) {
                    double2 t0 = *chunk_ptr<WT>(v, row, q >> 1), t1 = *chunk_ptr<WT>(v, row, (q >> 1) + 1);
                    double* o = a.dbg_J + (size_t)g * a.Wp + w;
                    o[0] = t0.x; o[1] = t0.y; o[2] = t1.x; o[3] = t1.y;
                }
            }
            __syncthreads();
            continue;
        }
        amp_fill(a, ac, tile, a.i0);
        __syncthreads();
        for (long long k = 0; k <= a.n; ++k) {
            const long long n = a.i0 + k;
            long long tq = a.dbg_prof ? clock64() : 0;
            if (!(a.dbg_skip & 2) || k == 0 || k == a.n) junction_pass<WT, DEF>(a, ac, rank, tile, n, k > 0, k < a.n, v);
            if (k == a.n) break;
            __syncthreads();
            if (a.dbg_prof && threadIdx.x == 0) { const long long tn = clock64(); a.dbg_prof[(size_t)blockIdx.x * (2 + a.n_ops)] += tn - tq; tq = tn; }
            if (!(a.dbg_skip & 4)) face_pass<WT>(a, ac, rank, tile, n, v);
            if (k + 1 < a.n) amp_fill(a, ac, tile, n + 1);     // other parity: read after the barriers of the sweeps
            __syncthreads();
            if (a.dbg_prof && threadIdx.x == 0) a.dbg_prof[(size_t)blockIdx.x * (2 + a.n_ops) + 1] += clock64() - tq;
            if (!(a.dbg_skip & 1)) run_ops<WT, CL>(a, ps, cluster, rank, v, stage, mbox);
        }
        __syncthreads();
    }
    if (CL) cluster.sync();      // keep shared memory alive until every peer is done with it
}

typedef void (*KernelPtr)(const ResArgs);

KernelPtr pick_kernel(int WT, bool def, bool cl) {
    (void)WT;      // 8 problems per tile: the N of the FP64 MMA
    if (def) return cl ? k_resident<8, true, true> : k_resident<8, true, false>;
    return cl ? k_resident<8, false, true> : k_resident<8, false, false>;
}

// per-junction constants in device junction order (one coalesced 32-byte record instead of gathers by original index)
__global__ void k_gather_params(int n, const int* orig, const double* Ic, const double* c0, const double* c1,
                                const double* c2, const double* isb, const double* tb, const double* vsb,
                                double* P0, double* P1) {
    int jp = blockIdx.x * blockDim.x + threadIdx.x;
    if (jp >= n) return;
    int jo = orig[jp];
    P0[4 * jp + 0] = Ic[jo]; P0[4 * jp + 1] = 1.0 / c0[jo]; P0[4 * jp + 2] = c1[jo]; P0[4 * jp + 3] = c2[jo];
    P1[4 * jp + 0] = isb ? isb[jo] : 0.0; P1[4 * jp + 1] = tb ? tb[jo] : 0.0; P1[4 * jp + 2] = vsb ? vsb[jo] : 0.0;
    P1[4 * jp + 3] = 0.0;
}

#define RCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            h->err = std::string("resident: ") + #call + ": " + cudaGetErrorString(e_);             \
            return JJ_ECUDA;                                                                        \
        }                                                                                           \
    } while (0)

template <typename T>
int up(JJHandle* h, ResidentState* st, T** dst, const T* src, size_t n) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    int rc = dev_alloc(h, &p, bytes);
    if (rc) return rc;
    st->allocs.push_back(p); st->alloc_bytes.push_back(bytes);
    if (n) {
        cudaError_t e = cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { h->err = std::string("resident upload: ") + cudaGetErrorString(e); return JJ_ECUDA; }
    }
    *dst = (T*)p;
    return JJ_OK;
}

}  // namespace

namespace jj {

void resident_free_problem(JJHandle* h) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    if (!st) return;
    dev_free(h, st->rth, st->state_bytes); dev_free(h, st->rx, st->state_bytes);
    dev_free(h, st->plane_d, st->plane_cap);
    st->rth = st->rx = nullptr; st->plane_d = nullptr; st->plane_cap = 0; st->state_bytes = 0;
    st->prepared = false;
}

void resident_free(JJHandle* h) { resident_free_problem(h); h->resident = nullptr; }

void resident_drop_plan(JJHandle* h) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    if (!st) return;
    resident_free_problem(h);
    for (size_t i = 0; i < st->allocs.size(); ++i) dev_free(h, st->allocs[i], st->alloc_bytes[i]);
    delete st;
    h->resident_plan = nullptr;
    h->resident = nullptr;
}

int resident_set_plan(JJHandle* h, const JJResidentPlan* pl) {
    resident_drop_plan(h);
    if (!pl) return JJ_OK;
    if (!(pl->C == 1 || pl->C == 2 || pl->C == 4 || pl->C == 8) || pl->tile_problems != 8) {
        h->err = "resident plan: cluster size must be 1/2/4/8 and tile_problems 8";
        return JJ_EINVAL;
    }
    ResidentState* st = new ResidentState();
    h->resident_plan = st;
    st->C = pl->C; st->WT = pl->tile_problems; st->n_rows = pl->n_rows; st->stage_rows = pl->stage_rows;
    st->ar_rows = pl->allreduce_rows; st->n_ops = pl->n_ops;
    const int Nj = h->cir.Nj;
    int rc;
    if ((rc = up(h, st, (int**)&st->ops, pl->ops, (size_t)pl->C * pl->n_ops * 4))) return rc;
    for (int r = 0; r < pl->C; ++r) {
        const JJRankStream& ps = pl->prog[r];
        if (ps.n_warps != NWARPS) { h->err = "resident plan: program packed for a different warp count"; return JJ_EINVAL; }
        int *wt, *ws; int* th; unsigned char* sb;
        size_t np = (size_t)ps.n_levels * ps.n_warps + 1;
        if ((rc = up(h, st, &wt, ps.wt_ptr, np))) return rc;
        if ((rc = up(h, st, &ws, ps.ws_ptr, np))) return rc;
        if ((rc = up(h, st, &th, ps.thdr, (size_t)ps.n_tiles * 2))) return rc;
        {   // stream + 2*RING zero steps of padding (the kernel prefetches RING steps ahead unconditionally)
            size_t nbytes = (size_t)ps.n_steps * STEP_BYTES, pad = (size_t)2 * RING * STEP_BYTES;
            void* pbuf = nullptr;
            if ((rc = dev_alloc(h, &pbuf, nbytes + pad))) return rc;
            st->allocs.push_back(pbuf); st->alloc_bytes.push_back(nbytes + pad);
            RCK(cudaMemsetAsync(pbuf, 0, nbytes + pad, h->stream));
            if (nbytes) RCK(cudaMemcpyAsync(pbuf, ps.stream, nbytes, cudaMemcpyHostToDevice, h->stream));
            sb = (unsigned char*)pbuf;
        }
        st->prog[r].wt_ptr = wt; st->prog[r].ws_ptr = ws; st->prog[r].thdr = (const int2*)th; st->prog[r].stream = sb;
        st->prog[r].n_levels = ps.n_levels; st->prog[r].n_tiles = ps.n_tiles;
        st->aux_ints = std::max<size_t>(st->aux_ints, 2 * np + 2 + 2 * (size_t)ps.n_tiles + 4 + 4 * (size_t)pl->n_ops + sizeof(AmpCache) / sizeof(int) + 16);
    }
    if ((rc = up(h, st, &st->junc_ptr, pl->junc_ptr, (size_t)pl->C + 1))) return rc;
    if ((rc = up(h, st, &st->junc_orig, pl->junc_orig, (size_t)Nj))) return rc;
    if ((rc = up(h, st, (int**)&st->junc_row, pl->junc_row, (size_t)Nj * 2))) return rc;
    if ((rc = up(h, st, (signed char**)&st->junc_sign, (const signed char*)pl->junc_sign, (size_t)Nj * 2))) return rc;
    size_t nfp = (size_t)pl->C * (pl->n_rows + 1);
    if ((rc = up(h, st, &st->face_ptr, pl->face_ptr, nfp))) return rc;
    int nent = 0;
    for (size_t i = 0; i < nfp; ++i) nent = std::max(nent, pl->face_ptr[i]);
    if ((rc = up(h, st, &st->face_junc, pl->face_junc, (size_t)nent))) return rc;
    if ((rc = up(h, st, &st->face_sign, (const signed char*)pl->face_sign, (size_t)nent))) return rc;
    if ((rc = up(h, st, &st->face_fidx, pl->face_fidx, (size_t)pl->C * pl->n_rows))) return rc;
    {   // fixed-width (ELL) form of the per-rank face lists with the coefficient sign / c0 folded in
        int K = 1;
        for (int r = 0; r < pl->C; ++r)
            for (int row = 0; row < pl->n_rows; ++row) {
                const int* fp = pl->face_ptr + (size_t)r * (pl->n_rows + 1) + row;
                K = std::max(K, fp[1] - fp[0]);
            }
        K = (K + 3) / 4 * 4;
        std::vector<int> ej((size_t)pl->C * pl->n_rows * K, -1);
        std::vector<double> ec((size_t)pl->C * pl->n_rows * K, 0.0);
        std::vector<double> c0h(Nj);
        RCK(cudaMemcpy(c0h.data(), h->cir.c0, (size_t)Nj * sizeof(double), cudaMemcpyDeviceToHost));
        for (int r = 0; r < pl->C; ++r)
            for (int row = 0; row < pl->n_rows; ++row) {
                const int* fp = pl->face_ptr + (size_t)r * (pl->n_rows + 1) + row;
                for (int p = fp[0]; p < fp[1]; ++p) {
                    size_t o = ((size_t)r * pl->n_rows + row) * K + (p - fp[0]);
                    ej[o] = pl->face_junc[p];
                    ec[o] = (double)pl->face_sign[p] * (1.0 / c0h[pl->junc_orig[pl->face_junc[p]]]);
                }
            }
        st->face_K = K;
        if ((rc = up(h, st, &st->face_ell_j, ej.data(), ej.size()))) return rc;
        if ((rc = up(h, st, &st->face_ell_c, ec.data(), ec.size()))) return rc;
    }
    for (double** pp : {&st->P0, &st->P1}) {
        void* p = nullptr;
        if ((rc = dev_alloc(h, &p, (size_t)Nj * 4 * sizeof(double)))) return rc;
        st->allocs.push_back(p); st->alloc_bytes.push_back((size_t)Nj * 4 * sizeof(double));
        *pp = (double*)p;
    }
    RCK(cudaStreamSynchronize(h->stream));
    st->smem_bytes = (((size_t)st->n_rows + st->ar_rows + 8 * st->C) * st->WT + (size_t)st->stage_rows * (st->WT + 2)) * sizeof(double)
                     + st->aux_ints * sizeof(int);
    return JJ_OK;
}

int resident_supported(JJHandle* h, std::string& why) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    if (!st) { why = "no resident plan was provided (problem too large for shared memory?)"; return 0; }
    for (int i = 0; i < 4; ++i)
        if (h->src[i].dev.kind == KIND_DENSE) { why = "a per-step input is dense (not base x amplitude)"; return 0; }
    if (h->cir.Nf == 0) { why = "circuit has no faces"; return 0; }
    return 1;
}

static int fill_args(JJHandle* h, ResidentState* st, ResArgs& a) {
    memset(&a, 0, sizeof(a));
    a.C = st->C; a.n_rows = st->n_rows; a.stage_rows = st->stage_rows; a.ar_rows = st->ar_rows; a.n_ops = st->n_ops;
    a.ops = st->ops;
    for (int r = 0; r < st->C; ++r) a.prog[r] = st->prog[r];
    a.junc_ptr = st->junc_ptr; a.junc_orig = st->junc_orig; a.junc_row = st->junc_row; a.junc_sign = st->junc_sign;
    a.face_ptr = st->face_ptr; a.face_junc = st->face_junc; a.face_sign = st->face_sign; a.face_fidx = st->face_fidx;
    a.face_K = st->face_K; a.face_ell_j = st->face_ell_j; a.face_ell_c = st->face_ell_c;
    a.Nj = h->cir.Nj; a.Nf = h->cir.Nf; a.P0 = st->P0; a.P1 = st->P1;
    a.cpr = h->cir.cpr;
    a.Wp = h->Wp; a.n_tiles = st->n_tiles; a.dt = h->dt; a.seed = h->seed; a.group_offset = h->problem_offset / 4;
    a.Is = h->src[JJ_SRC_IS].dev; a.Vs = h->src[JJ_SRC_VS].dev; a.T = h->src[JJ_SRC_T].dev; a.F = h->src[JJ_SRC_F].dev;
    a.noise = h->noise_buf; a.noise_i0 = h->noise_i0; a.noise_K = h->noise_K;
    a.rth = st->rth; a.rx = st->rx; a.th1 = h->th1; a.th2 = h->th2;
    a.snap_th = h->th_out; a.snap_I = h->I_out; a.flag = h->flag_d;
    const char* dbg = getenv("JJ_RES_DEBUG");
    a.dbg_skip = dbg ? atoi(dbg) : 0;
    return JJ_OK;
}

static int launch(JJHandle* h, ResidentState* st, const ResArgs& a) {
    KernelPtr k = pick_kernel(st->WT, h->cir.default_cpr, st->C > 1);
    {
        const Source &is = h->src[JJ_SRC_IS].dev, &t = h->src[JJ_SRC_T].dev, &vs = h->src[JJ_SRC_VS].dev;
        k_gather_params<<<(h->cir.Nj + 255) / 256, 256, 0, h->stream>>>(
            h->cir.Nj, st->junc_orig, h->cir.Ic, h->cir.c0, h->cir.c1, h->cir.c2,
            is.kind == KIND_RANK1 ? is.base : nullptr, t.kind == KIND_RANK1 ? t.base : nullptr,
            vs.kind == KIND_RANK1 ? vs.base : nullptr, st->P0, st->P1);
        h->launches++;
    }
    RCK(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem_bytes));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = st->smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = st->C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (st->max_clusters == 0) {
        cfg.gridDim = dim3(st->C, 1, 1);
        int nc = 0;
        RCK(cudaOccupancyMaxActiveClusters(&nc, (const void*)k, &cfg));
        if (nc <= 0) { h->err = "resident: kernel does not fit on the device (shared memory / cluster size)"; return JJ_EINVAL; }
        st->max_clusters = nc;
    }
    int ncl = std::min(st->n_tiles, st->max_clusters);
    cfg.gridDim = dim3(ncl * st->C, 1, 1);
    RCK(cudaLaunchKernelEx(&cfg, k, a));
    h->launches++;
    return JJ_OK;
}

int resident_prepare(JJHandle* h) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    if (!st) { h->err = "resident engine: no plan"; return JJ_ESTATE; }
    resident_free_problem(h);
    st->n_tiles = (h->Wp + st->WT - 1) / st->WT;
    st->state_bytes = (size_t)st->n_tiles * h->cir.Nj * st->WT * sizeof(double);
    int rc;
    if ((rc = dev_alloc(h, (void**)&st->rth, st->state_bytes))) return rc;
    if ((rc = dev_alloc(h, (void**)&st->rx, st->state_bytes))) return rc;
    RCK(cudaMemsetAsync(st->rth, 0, st->state_bytes, h->stream));
    RCK(cudaMemsetAsync(st->rx, 0, st->state_bytes, h->stream));
    st->prepared = true;
    h->resident = st;
    return JJ_OK;
}

int resident_run(JJHandle* h, long long i0, int n, const long long* th_plane, const long long* I_plane) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    if (!st || !st->prepared) { h->err = "resident engine not prepared"; return JJ_ESTATE; }
    size_t need = (size_t)2 * n * sizeof(long long);
    if (need > st->plane_cap) {
        RCK(cudaStreamSynchronize(h->stream));
        dev_free(h, st->plane_d, st->plane_cap);
        st->plane_d = nullptr; st->plane_cap = 0;
        int rc = dev_alloc(h, (void**)&st->plane_d, need);
        if (rc) return rc;
        st->plane_cap = need;
    }
    std::vector<long long> pl((size_t)2 * n, -1);
    for (int k = 0; k < n; ++k) {
        if (th_plane) pl[k] = th_plane[k];
        if (I_plane) pl[n + k] = I_plane[k];
        if (pl[k] >= h->n_th_planes || pl[n + k] >= h->n_I_planes) { h->err = "run: plane index out of range"; return JJ_EINVAL; }
    }
    RCK(cudaMemcpyAsync(st->plane_d, pl.data(), need, cudaMemcpyHostToDevice, h->stream));
    RCK(cudaStreamSynchronize(h->stream));     // pl goes out of scope
    ResArgs a;
    fill_args(h, st, a);
    a.i0 = i0; a.n = n; a.th_plane = st->plane_d; a.I_plane = st->plane_d + n;
    if (!getenv("JJ_RES_PROF")) return launch(h, st, a);
    // debugging aid: per-phase cycle counts of every block, printed as averages per time step
    int ncl = std::min(st->n_tiles, st->max_clusters ? st->max_clusters : st->n_tiles);
    size_t nb = (size_t)std::max(ncl, st->n_tiles) * st->C, slots = 2 + st->n_ops;
    long long* prof = nullptr;
    RCK(cudaMalloc((void**)&prof, nb * slots * sizeof(long long)));
    RCK(cudaMemset(prof, 0, nb * slots * sizeof(long long)));
    a.dbg_prof = prof;
    int rc = launch(h, st, a);
    if (rc) return rc;
    RCK(cudaStreamSynchronize(h->stream));
    std::vector<long long> hp(nb * slots);
    RCK(cudaMemcpy(hp.data(), prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(prof);
    std::vector<int> ops((size_t)st->C * st->n_ops * 4);
    RCK(cudaMemcpy(ops.data(), st->ops, ops.size() * sizeof(int), cudaMemcpyDeviceToHost));
    fprintf(stderr, "JJ_RES_PROF: cycles per time step, averaged over blocks (n=%d steps, C=%d)\n", n, st->C);
    double tot = 0;
    for (size_t sl = 0; sl < slots; ++sl) {
        double sum = 0, mx = 0; int cnt = 0;
        for (size_t b = 0; b < nb; ++b) { double v = (double)hp[b * slots + sl] / n; if (v > 0) { sum += v; ++cnt; } mx = std::max(mx, v); }
        double avg = cnt ? sum / cnt : 0; tot += avg;
        if (sl == 0) fprintf(stderr, "  junction pass   avg %9.0f max %9.0f\n", avg, mx);
        else if (sl == 1) fprintf(stderr, "  face pass       avg %9.0f max %9.0f\n", avg, mx);
        else fprintf(stderr, "  op %2zu kind %d arg %4d avg %9.0f max %9.0f\n", sl - 2, ops[(sl - 2) * 4], ops[(sl - 2) * 4 + 1], avg, mx);
    }
    fprintf(stderr, "  total avg cycles per time step %.0f\n", tot);
    return JJ_OK;
}

int resident_debug_solve(JJHandle* h, const double* b_d, double* J_d) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    if (!st) { h->err = "resident engine: no plan"; return JJ_ESTATE; }
    if (!st->prepared) { int rc = resident_prepare(h); if (rc) return rc; }
    ResArgs a;
    fill_args(h, st, a);
    a.dbg_b = b_d; a.dbg_J = J_d;
    return launch(h, st, a);
}

void resident_get_config(JJHandle* h, int* C, int* WT) {
    ResidentState* st = (ResidentState*)h->resident_plan;
    *C = st ? st->C : 1; *WT = st ? st->WT : 0;
}

int resident_set_state(JJHandle*, const double*, const double*) { return JJ_OK; }   // reads the canonical arrays
int resident_get_state(JJHandle* h, double*, double*) { h->err = "internal: canonical state is authoritative"; return JJ_ESTATE; }

}  // namespace jj
