// SUBDOMAIN step engine for sm_100a: one persistent cooperative kernel integrates a whole jj_run call.
//
// Host plan: pyjjasim_b200/subdomain.py. The elimination tree of the cycle-space system is cut into P
// mutually uncoupled subdomains plus the separators above the cut (the TOP rows). A thread block owns a
// (subdomain, chunk of PC = 8*NG problems) item; everything local to it runs out of shared memory:
//
//   backward sweep  J_loc = D^-T (z_loc - L[top, loc]^T J_top) level by level   (reference: time_evolution.py:562-569)
//   junction pass   theta_n = (A^T J - x)/c0; x' = noise - Is + Ic cpr(2 theta_n - theta_{n-1}) + c1.. + c2..  (:533-558, :570-580)
//   face pass       b = A (x'/c0 - theta_s) - 2 pi f for local faces, partial sums for top faces           (:560-569)
//   forward sweep   z_loc = D^-1 (b_loc - L z), then the subdomain's contribution to the top right-hand side
//
// followed, once per time step and for all problems at once, by
//
//   top assembly    r_top = sum of the subdomain contributions - 2 pi f_top
//   upper forward   the separators between the subdomains and the top of the tree, two phases per tree depth,
//                   every phase a set of independent gathered dense FP64 tensor-core products (upper_phase)
//   top product     J_tt = S_tt^-1 r_tt for the last few tree levels, one dense product spread over all SMs
//   upper backward  the same separators top-down
//
// with a grid barrier between stages (circuits whose separators all fit the dense product skip the upper phases). The local sweeps stream the factor from L2 as
// mma.m8n8k4 fragments (one A fragment feeds NG MMAs, one per group of 8 problems); theta and x stream
// through HBM once per step; b, z, J of the local rows never leave shared memory when every block has a
// single item.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "jj_host.h"

using namespace jj;

// kernel arguments: shared by the host unit and the per-chunk-width kernel units (see the end of the device part)
namespace jj {

struct SubProgDev {
    const int* wt_ptr; const int* ws_ptr; const int2* thdr; const int* lstaged;
    const unsigned char* stream;
    int n_levels, n_bwd, n_tiles, pad_;
};

struct SubArgs {
    // plan
    int P, n_rows, n_loc_max, stage_rows, n_top, n_up_pad, n_slots;
    int tt0, n_tt, n_tt_pad;        // dense top of the top: rows [tt0, tt0 + n_tt) of the top numbering
    int max_np, max_levels, max_tiles;
    // upper program (see JJSubdomainPlan): phases of tasks out[rows] = V . X[cols] over the r / z / J planes
    int up_RB, up_KB, n_up_fwd, n_up_bwd;
    const int* up_phase_ptr; const int* up_phase_split; const int4* up_task; const long long* up_task_aoff;
    unsigned* up_ctr;               // [phases] monotonic work counters of the block tasks (zeroed per launch)
    const int* up_cols; const double* up_A;
    double* U;                      // [4 planes: r, z, J, scratch][chunk][n_up_pad][PC]; rtop = plane 0, jtop = plane 2
    const SubProgDev* prog;
    const int *n_loc, *n_halo, *hptr, *halo_top, *tptr, *tslot, *top_face;
    const int4* tslot4;             // [n_top] the (at most four) slots of a top row, -1 padded; null when a row has more
    const double* topF;             // [n_top] flux base of each top row (gathered per launch), null without rank-one flux
    const double* rowF;             // [P][n_rows] flux base of each local row, 0 for halo rows (gathered per launch), null without rank-one flux
    const double* SinvP;
    const int* junc_ptr; const int* junc_orig; const int2* junc_row; const char2* junc_sign;
    int face_K; const int* face_ell_j; const double* face_ell_c; const int* face_fidx;
    // circuit
    int Nj, Nf;
    const double *P0, *P1;          // [Nj'][4] = Ic, 1/c0, c1, c2 | Is base, noise base, Vs base, 0 (device junction order)
    const double* jrec;             // [Nj'][8] junction records, see junction_item
    Cpr cpr;
    // problem
    int Wp, n_chunks;
    double dt;
    unsigned long long seed; long long group_offset;
    Source Is, Vs, T, F;
    const double* noise; long long noise_i0; int noise_K;
    // state
    double *rth, *rx;               // [chunk][Nj'][PC]
    double *th1, *th2;              // canonical [Nj][Wp]
    const double* th2_in;           // theta(-2) the run starts from: th2, or th1 for a start at rest (no copy needed)
    double *zloc, *ctop, *rtop, *jtop;
    unsigned* bar;
    // run
    long long i0; int n;
    const long long* th_plane; const long long* I_plane;
    double *snap_th, *snap_I;
    int* flag;
    // running observables (jj_observe_begin): steps obs_first + m * obs_interval add their vortex configuration to
    // obs_nsum [Nf][Wp] (permuted faces) and leave their phases in obs_th_last (the first one also in obs_th_first)
    long long obs_first; int obs_interval;
    int* obs_nsum; double *obs_th_first, *obs_th_last;
    // vortex mobility of an annealing interval (jj_anneal): the steps with a theta plane number p >= 0 store the low byte
    // of round(theta / 2 pi) to zone8[p][Nj][Wp] (original junction order) instead of theta; null = off
    unsigned char* zone8;
    const double* dbg_b; double* dbg_J;   // debug solve: canonical [Nf][Wp], permuted faces
    long long* prof;                       // optional per-block cycle counters [block][8]
    int stagger;                           // cycles by which consecutive problem chunks start apart (chunk-local top phase only)
    int dbg;                               // JJ_SUB_DEBUG (read once per plan): 4 no chunk-local top phase, 64 direct top
                                           // product; the physics-altering timing experiments (128, 256) exist only in
                                           // builds with -DJJ_EXPERIMENTS
};

}  // namespace jj

namespace {

// threads per block: 512 (one block per SM) or, compiled with -DJJ_SUB_NT=256, half blocks of 8 warps that run two to
// an SM on different (subdomain, chunk) items: the latency-bound phases of one overlap the pipe-bound phases of the other
#ifndef JJ_SUB_NT
#define JJ_SUB_NT 512
#endif
constexpr int NT = JJ_SUB_NT;
constexpr int BLOCKS_PER_SM = NT == 256 ? 2 : 1;
constexpr size_t BAR_COUNTERS = 8192;   // grid barrier counter + one counter per problem chunk, 128 bytes apart
constexpr size_t BAR_BYTES = BAR_COUNTERS + 128;     // ... + the 64-bit work counter of the item scheduler
constexpr int NWARPS = NT / 32;
// The annealing schedule's kernels (build.sh: units jj_subdomain_m_ng*, -DJJ_SUB_MOB=1) store the PHASE ZONES of the
// requested steps (one byte per junction and problem, SubArgs::zone8) instead of the phases; the code is compiled out of
// the ordinary step kernels, whose register allocation at the 128-register cap does not tolerate passengers
// (profiles/r02_experiments.md).
#ifndef JJ_SUB_MOB
#define JJ_SUB_MOB 0
#endif
constexpr bool MOB = JJ_SUB_MOB != 0;
constexpr int RING = 4;            // stream steps per ring block; every tile is padded to a multiple of it
constexpr int STEP_BYTES = 320;
constexpr double TWO_PI = 6.283185307179586;
constexpr int PROF_UP = 48;             // upper phases with their own counters
constexpr int PROF_STAMPS = 8;          // time stamps inside the first warp task of block 0 (latency chain of a small task)
constexpr int PROF_SLOTS = 8 + 48 + 2 * PROF_UP + PROF_STAMPS;     // phase counters + per-level counters of the sweeps + (work, wait) per upper phase (JJ_SUB_PROF)

struct SubState {
    int threads = 512;      // block size the plan was packed for (16 or 8 warps)
    int P = 1, NG = 4, PC = 32, n_rows = 0, n_loc_max = 0, stage_rows = 0, n_top = 0, n_up_pad = 0, n_slots = 0;
    int tt0 = 0, n_tt = 0, n_tt_pad = 0, up_RB = 4, up_KB = 8, n_up_fwd = 0, n_up_bwd = 0;
    int* up_phase_ptr = nullptr; int* up_phase_split = nullptr; int4* up_task = nullptr; long long* up_task_aoff = nullptr;
    unsigned* up_ctr = nullptr;
    int* up_cols = nullptr; double* up_A = nullptr;
    double* U = nullptr; size_t u_bytes = 0;
    int stagger = -1;       // JJ_SUB_STAGGER (cycles per chunk; -1: derived from the measured step time)
    int dbg = 0; bool prof = false, no_tslot4 = false, no_l2_window = false, l2_limit_set = false; int grid_env = 0;
    int max_np = 0, max_levels = 0, max_tiles = 0, face_K = 0;
    SubProgDev* prog = nullptr;
    int *n_loc = nullptr, *n_halo = nullptr, *hptr = nullptr, *halo_top = nullptr, *tptr = nullptr, *tslot = nullptr, *top_face = nullptr;
    double* SinvP = nullptr;
    int4* tslot4 = nullptr; double* topF = nullptr; double* rowF = nullptr;
    int *junc_ptr = nullptr, *junc_orig = nullptr; int2* junc_row = nullptr; char2* junc_sign = nullptr;
    int* face_ell_j = nullptr; double* face_ell_c = nullptr; int* face_fidx = nullptr;
    double *P0 = nullptr, *P1 = nullptr, *jrec = nullptr;
    std::vector<void*> allocs; std::vector<size_t> alloc_bytes;
    // per problem
    double *rth = nullptr, *rx = nullptr, *zloc = nullptr, *ctop = nullptr, *rtop = nullptr, *jtop = nullptr;
    size_t state_bytes = 0, z_bytes = 0, c_bytes = 0;
    unsigned* bar = nullptr;
    long long* plane_d = nullptr; size_t plane_cap = 0; std::vector<long long> plane_h;   // device plane table and what it holds
    int n_chunks = 0, grid = 0;
    size_t smem_bytes = 0;
    bool prepared = false;
    unsigned long long gather_gen = 0;     // JJHandle::src_gen the gathered constants (jrec, P0, P1, topF, rowF) were made for
};

// ------------------------------------------------------------------------------------------------
// shared-memory vector: row r holds PC float64 as PC/4 chunks of 32 bytes; chunk c of row r sits at chunk
// position c ^ (r & 3) (c ^ (r & 1) when a row has only two chunks). One MMA B fragment gathers, per half
// warp, the same logical chunk of four rows: consecutive rows land in four different bank octets.
// ------------------------------------------------------------------------------------------------
template <int NG>
__device__ __forceinline__ int velem(int row, int q) {
    constexpr int CM = NG >= 2 ? 3 : 1;
    return row * (8 * NG) + (((q >> 2) ^ (row & CM)) << 2) + (q & 3);
}

// L2 residency hints: the factor streams, the packed Schur inverse and the junction records are re-read every
// time step (evict_last), while the state streams through once per step (evict_first) and would otherwise
// push them out of the 126 MB L2 between two uses.
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 256-bit global accesses (sm_100: LDG/STG.E.256): four problems of one junction in one request
struct double4v { double2 lo, hi; };
__device__ __forceinline__ double4v ldg256_hint(const double* p, unsigned long long pol) {
    double4v v;
    asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(v.lo.x), "=d"(v.lo.y), "=d"(v.hi.x), "=d"(v.hi.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double4v ldg256(const double* p) {
    double4v v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.lo.x), "=d"(v.lo.y), "=d"(v.hi.x), "=d"(v.hi.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg256_hint(double* p, double2 lo, double2 hi, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(lo.x), "d"(lo.y), "d"(hi.x), "d"(hi.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg256(double* p, double2 lo, double2 hi) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(lo.x), "d"(lo.y), "d"(hi.x), "d"(hi.y) : "memory");
}

struct Cursor {
    int t0, t1, s;
    const unsigned char* pA;      // + lane*8: A fragment of the first step of the current ring block
    const unsigned char* pC;      // + 256 + lane*2: its element codes
    double ra[RING]; unsigned rc[RING];
};

struct ProgSmem {
    const int* wt; const int* ws; const int* lstaged; const int2* thdr;
    const unsigned char* stream;
    int n_levels, n_bwd;
};

__device__ __forceinline__ void cursor_open(Cursor& cu, const ProgSmem& ps, int level) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = level * NWARPS + warp;
    const unsigned long long pol = policy_evict_last();
    cu.t0 = ps.wt[2 * idx]; cu.t1 = ps.wt[2 * idx + 1];
    cu.s = ps.ws[idx];
    // every tile is a whole number of ring blocks, so a warp's stream position is always block aligned; the
    // stream buffer is padded by 2*RING steps
    const unsigned char* blk = ps.stream + (size_t)cu.s * STEP_BYTES;
    cu.pA = blk + lane * 8;
    cu.pC = blk + 256 + lane * 2;
#pragma unroll
    for (int k = 0; k < RING; ++k) {
        const int off = k * STEP_BYTES;
        asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(cu.ra[k]) : "l"(cu.pA + off), "l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=r"(cu.rc[k]) : "l"(cu.pC + off), "l"(pol));
    }
}

// C(8 rows x 8 problems) += A(8 x 4) . B(4 x 8) on the FP64 tensor core: lane = row*4 + kk holds A[row][kk],
// lane = n*4 + kk holds B[kk][n], lane = row*4 + c holds C[row][2c], C[row][2c+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double b;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b) : "r"(addr));
    return b;
}

// One tile for NGT of the NG problem groups: load the running sums (TILE_SELF) or start from zero, walk the
// tile's stream, write the 8-row result (in place or into the staging rows). A tile is a whole number of ring
// blocks of RING steps (padded with zero steps by the host), so the step loop has no per-step control flow: ring
// slot K holds step K of the current block and is refilled in place, right after use, with step K of the next
// block. All B fragments of a step are fetched before its MMAs issue. NGT is a template parameter so that none
// of the per-group work is predicated.
template <int NG, int NGT>
__device__ __forceinline__ void tile_run(int nsteps, int flags, bool in_rows, int celem0, int row, int stage_off,
                                         int kk, const unsigned char*& pA, const unsigned char*& pC,
                                         double (&ra)[RING], unsigned (&rc)[RING], unsigned vb, unsigned gx,
                                         unsigned long long pol, double* __restrict__ v, double* __restrict__ stage) {
    constexpr int PC = 8 * NG;
    double acc[NGT][2];
#pragma unroll
    for (int g = 0; g < NGT; ++g) { acc[g][0] = 0.0; acc[g][1] = 0.0; }
    if ((flags & 1) && in_rows) {
#pragma unroll
        for (int g = 0; g < NGT; ++g) {
            const double2 sv = *reinterpret_cast<const double2*>(v + (celem0 ^ (g << 3)));
            acc[g][0] = sv.x; acc[g][1] = sv.y;
        }
    }
    for (int j = 0; j < nsteps; j += RING) {
#pragma unroll
        for (int K = 0; K < RING; ++K) {
            const unsigned a0 = vb + ((rc[K] ^ gx) << 3);
            double b_[NGT];
#pragma unroll
            for (int g = 0; g < NGT; ++g) b_[g] = lds_f64(a0 ^ (unsigned)(g << 6));
#pragma unroll
            for (int g = 0; g < NGT; ++g) dmma884(acc[g][0], acc[g][1], ra[K], b_[g]);
            asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(ra[K]) : "l"(pA + (RING + K) * STEP_BYTES), "l"(pol));
            asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=r"(rc[K]) : "l"(pC + (RING + K) * STEP_BYTES), "l"(pol));
        }
        pA += RING * STEP_BYTES; pC += RING * STEP_BYTES;
    }
    __syncwarp();                // every lane has read its operands before rows of this block are overwritten
    if (in_rows) {
        if (flags & 2) {
            // staged: the row keeps the physical layout of its destination and carries the destination index
            double* srow = stage + (size_t)stage_off * (PC + 2) - (size_t)row * PC;
#pragma unroll
            for (int g = 0; g < NGT; ++g) *reinterpret_cast<double2*>(srow + (celem0 ^ (g << 3))) = make_double2(acc[g][0], acc[g][1]);
            if (kk == 0) reinterpret_cast<int*>(stage + (size_t)stage_off * (PC + 2) + PC)[0] = row;
        } else {
#pragma unroll
            for (int g = 0; g < NGT; ++g) *reinterpret_cast<double2*>(v + (celem0 ^ (g << 3))) = make_double2(acc[g][0], acc[g][1]);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ int bcast0(int x) { return __shfl_sync(0xffffffffu, x, 0); }

// One level of a subdomain's sweep program: every warp walks its own stream of 8-row tiles. Everything that
// steers control flow is broadcast from lane 0, so the compiler knows the warp is converged at every MMA.
template <int NG>
__device__ void exec_level(const ProgSmem& ps, Cursor& cu, int level, int next_level, double* __restrict__ v,
                           double* __restrict__ stage) {
    constexpr int PC = 8 * NG;
    const int lane = threadIdx.x & 31;
    const unsigned vb = (unsigned)__cvta_generic_to_shared(v);      // 1 KB aligned: group g of an element is addr ^ (g << 6)
    const unsigned long long pol = policy_evict_last();
    int s = bcast0(cu.s);
    const int t0 = bcast0(cu.t0), t1 = bcast0(cu.t1);
    const unsigned char* pA = cu.pA;
    const unsigned char* pC = cu.pC;
    double ra[RING]; unsigned rc[RING];
#pragma unroll
    for (int k = 0; k < RING; ++k) { ra[k] = cu.ra[k]; rc[k] = cu.rc[k]; }
    const int r = lane >> 2, kk = lane & 3;
    int2 hdn = t0 < t1 ? ps.thdr[t0] : make_int2(0, 0);
    for (int t = t0; t < t1; ++t) {
        const int hx = bcast0(hdn.x), hy = bcast0(hdn.y);
        if (t + 1 < t1) hdn = ps.thdr[t + 1];        // header of the next tile: its latency hides behind this tile
        const int row0 = hx & 0xffff, nrows = ((hx >> 16) & 7) + 1, flags = (hx >> 19) & 3;
        const int g0 = (hx >> 21) & 15, ng = ((hx >> 25) & 15) + 1;
        const int nsteps = hy & 0xffff, stage_off = (hy >> 16) & 0x7fff;
        const int row = row0 + r;
        // this lane's two C values of the tile's first group g0; group g0 + g: ^ (g << 3)
        const int celem0 = velem<NG>(row, 2 * kk) ^ (g0 << 3);
        const unsigned gx = (unsigned)g0 << 3;
        const bool in_rows = r < nrows;
        if (ng == NG) tile_run<NG, NG>(nsteps, flags, in_rows, celem0, row, stage_off + r, kk, pA, pC, ra, rc, vb, gx, pol, v, stage);
        else if (NG >= 4 && ng == NG / 2)
            tile_run<NG, (NG >= 4 ? NG / 2 : 1)>(nsteps, flags, in_rows, celem0, row, stage_off + r, kk, pA, pC, ra, rc, vb, gx, pol, v, stage);
        else if (NG >= 8 && ng == NG / 4)
            tile_run<NG, (NG >= 8 ? NG / 4 : 1)>(nsteps, flags, in_rows, celem0, row, stage_off + r, kk, pA, pC, ra, rc, vb, gx, pol, v, stage);
        else tile_run<NG, 1>(nsteps, flags, in_rows, celem0, row, stage_off + r, kk, pA, pC, ra, rc, vb, gx, pol, v, stage);
        s += nsteps;
    }
    if (next_level >= 0) {
        // the streams are laid out warp-major inside a sweep: normally the next level simply continues this
        // warp's stream and the ring stays as it is
        const int idx = next_level * NWARPS + (threadIdx.x >> 5);
        if (bcast0(ps.ws[idx]) == s) {
            cu.t0 = ps.wt[2 * idx]; cu.t1 = ps.wt[2 * idx + 1]; cu.s = s; cu.pA = pA; cu.pC = pC;
#pragma unroll
            for (int k = 0; k < RING; ++k) { cu.ra[k] = ra[k]; cu.rc[k] = rc[k]; }
        } else {
            cursor_open(cu, ps, next_level);
        }
    }
    __syncthreads();
    const int staged = ps.lstaged[level];
    if (staged > 0) {
        for (int e = threadIdx.x; e < staged * (PC / 2); e += NT) {
            const int rr = e / (PC / 2), q = e % (PC / 2);
            const double* srow = stage + (size_t)rr * (PC + 2);
            const int dst_row = reinterpret_cast<const int*>(srow + PC)[0];
            reinterpret_cast<double2*>(v + (size_t)dst_row * PC)[q] = reinterpret_cast<const double2*>(srow)[q];
        }
        __syncthreads();
    }
}

template <int NG>
__device__ void run_levels(const ProgSmem& ps, int l0, int l1, double* v, double* stage, long long* prof) {
    if (l0 >= l1) return;
    Cursor cu;
    cursor_open(cu, ps, l0);
    long long tq = prof ? clock64() : 0;
    for (int l = l0; l < l1; ++l) {
        exec_level<NG>(ps, cu, l, l + 1 < l1 ? l + 1 : -1, v, stage);
        if (prof && threadIdx.x == 0 && l < 48) { const long long tn = clock64(); prof[(size_t)blockIdx.x * PROF_SLOTS + 8 + l] += tn - tq; tq = tn; }
    }
}

// Per-step amplitudes of the rank-one inputs for the problems of a chunk (both step parities are filled
// per item: the current snapshot of step n-1 needs Is of step n-1).
template <int PC>
struct AmpCache {
    double T[2][PC], Is[2][PC], Vs[2][PC], F[2][PC];
};

template <int PC>
__device__ __forceinline__ void amp_fill(const SubArgs& a, AmpCache<PC>* ac, int c, long long n) {
    for (int t = threadIdx.x; t < 4 * PC; t += NT) {
        const int e = t % PC, which = t / PC;
        const int w = c * PC + e;
        const Source& s = which == 0 ? a.T : which == 1 ? a.Is : which == 2 ? a.Vs : a.F;
        double val = 0.0;
        if (s.kind == KIND_RANK1 && w < a.Wp && n >= a.i0 && n < a.i0 + a.n) val = __ldg(s.table + source_row(s, n) * a.Wp + w);
        double* dst = which == 0 ? ac->T[n & 1] : which == 1 ? ac->Is[n & 1] : which == 2 ? ac->Vs[n & 1] : ac->F[n & 1];
        dst[e] = val;
    }
}

// x' = (noise - Is) + Ic cpr(2 theta_n - theta_{n-1}) + c1 theta_n + c2 theta_{n-1} for four problems
// (reference: time_evolution.py:533-558); called with the coefficients divided by c0 it returns x'/c0
template <bool DEF>
__device__ __forceinline__ void next_x(const SubArgs& a, double Ic, double c1, double c2, double isb, double nb,
                                       const double* ampT, const double* ampIs, int jo, int w, long long n,
                                       const double th1[4], const double th2[4], double xn[4]) {
    double fl[4] = {0, 0, 0, 0};
#ifdef JJ_EXPERIMENTS
    if (a.dbg & 128) {      // timing experiment (JJ_SUB_DEBUG=128): linear CPR, no noise -> the pass is loads + stores + a few FMAs
#pragma unroll
        for (int k = 0; k < 4; ++k) xn[k] = (Ic * (2.0 * th1[k] - th2[k]) + c1 * th1[k] + c2 * th2[k]) - isb * ampIs[k];
        return;
    }
#endif
    if (a.T.kind != KIND_ZERO) {
        double z[4];
        if (a.noise_K > 0) {
            const double2* zp = reinterpret_cast<const double2*>(a.noise + ((size_t)(n - a.noise_i0) * a.Nj + jo) * a.Wp + w);
            const double2 z0 = zp[0], z1 = zp[1];
            z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
        } else {
            normal4(a.seed, jo, a.group_offset + (w >> 2), n, z);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) fl[k] = (nb * ampT[k]) * z[k];
    }
    double arg[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) arg[k] = 2.0 * th1[k] - th2[k];
    cpr_eval4<DEF>(a.cpr, arg, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double X = Ic * g[k] + c1 * th1[k] + c2 * th2[k];
        xn[k] = (fl[k] - isb * ampIs[k]) + X;
    }
}

// theta_n = (A^T J - x)/c0, snapshots, x' for the next step  (reference: time_evolution.py:533-558, 570-580).
// One thread = one junction x four problems. The 64-byte junction record holds Ic, 1/c0, c1, c2, the Is and
// noise bases, the two shared-memory rows of its faces (row 0 with sign 0 when absent) and the original index.
template <int NG, bool DEF>
__device__ __forceinline__ void junction_item(const SubArgs& a, const AmpCache<8 * NG>* ac, int c, long long n, int idx,
                                              int jlo, bool do_pre, const double* __restrict__ v,
                                              int plane_th, int plane_I, const double2 x0, const double2 x1,
                                              const double2 t0, const double2 t1, const double2 rIc, const double2 rc,
                                              const double2 rb, const int4 ri) {
    constexpr int PC = 8 * NG, G = PC / 4;
    const int q = (idx % G) * 4;
    const int w = c * PC + q;
    if (w >= a.Wp) return;
    const size_t sidx = ((size_t)c * a.Nj + jlo) * PC + (size_t)idx * 4;
    const unsigned long long pol_first = policy_evict_first();
    // ri = row0, row1, original junction, signs
    const double s0 = (double)((ri.w & 3) - 1), s1 = (double)(((ri.w >> 2) & 3) - 1);
    const double2* ja = reinterpret_cast<const double2*>(v + velem<NG>(ri.x, q));
    const double2* jb = reinterpret_cast<const double2*>(v + velem<NG>(ri.y, q));
    const double2 a0 = ja[0], a1 = ja[1], b0 = jb[0], b1 = jb[1];
    const double y[4] = {fma(s1, b0.x, s0 * a0.x), fma(s1, b0.y, s0 * a0.y), fma(s1, b1.x, s0 * a1.x), fma(s1, b1.y, s0 * a1.y)};
    // the x stream holds x'/c0 (what the face pass sums): theta_n = y/c0 - x'/c0
    const double ic0 = rIc.y;
    const double th1[4] = {fma(y[0], ic0, -x0.x), fma(y[1], ic0, -x0.y), fma(y[2], ic0, -x1.x), fma(y[3], ic0, -x1.y)};
    const double th2[4] = {t0.x, t0.y, t1.x, t1.y};
    // one finiteness test for the four phases (a NaN or Inf in any of them poisons the sum)
    if (!(fabs((th1[0] + th1[1]) + (th1[2] + th1[3])) < 1.0e300)) atomicOr(a.flag, 1);
    // (the snapshot planes travel as two plane numbers, -1 = none, not as two pointers: two registers less in the loop)
    if (plane_th >= 0 || plane_I >= 0 || !do_pre) {
        const size_t cidx = (size_t)ri.z * a.Wp + w;
        if (MOB && a.zone8 && plane_th >= 0) {
            // the vortex configuration n = -A round(theta / 2 pi) moves by a few units per step at most: its changes are
            // exact from the zones modulo 256 (k_zone_mobility), 4 bytes per thread instead of 32
            const unsigned z = ((unsigned)(int)phase_zone(th1[0]) & 0xffu) | (((unsigned)(int)phase_zone(th1[1]) & 0xffu) << 8) |
                               (((unsigned)(int)phase_zone(th1[2]) & 0xffu) << 16) | (((unsigned)(int)phase_zone(th1[3]) & 0xffu) << 24);
            *reinterpret_cast<unsigned*>(a.zone8 + (size_t)plane_th * a.Nj * a.Wp + cidx) = z;
        } else if (plane_th >= 0) {
            double2* sp = reinterpret_cast<double2*>(a.snap_th + (size_t)plane_th * a.Nj * a.Wp + cidx);
            sp[0] = make_double2(th1[0], th1[1]); sp[1] = make_double2(th1[2], th1[3]);
        }
        if (plane_I >= 0) {
            const double* am = ac->Is[(n - 1) & 1] + q;
            const double isb = __ldg(a.P1 + 4 * (size_t)(jlo + idx / G));       // (the record holds Is base / c0)
            double2* sp = reinterpret_cast<double2*>(a.snap_I + (size_t)plane_I * a.Nj * a.Wp + cidx);
            sp[0] = make_double2(y[0] + isb * am[0], y[1] + isb * am[1]);
            sp[1] = make_double2(y[2] + isb * am[2], y[3] + isb * am[3]);
        }
        if (!do_pre) {
            // end of the run: hand theta_last and theta_{last-1} back in the canonical layout
            double2* o1 = reinterpret_cast<double2*>(a.th1 + cidx);
            double2* o2 = reinterpret_cast<double2*>(a.th2 + cidx);
            o1[0] = make_double2(th1[0], th1[1]); o1[1] = make_double2(th1[2], th1[3]);
            o2[0] = t0; o2[1] = t1;
            // (the state slab is only read again by the vortex pass of an observed last step)
            stg256_hint(a.rth + sidx, make_double2(th1[0], th1[1]), make_double2(th1[2], th1[3]), pol_first);
            return;
        }
    }
    double xn[4];
#ifdef JJ_EXPERIMENTS       // timing experiment (JJ_SUB_DEBUG=256): no state traffic; never part of a release build
    if (!(a.dbg & 256)) stg256_hint(a.rth + sidx, make_double2(th1[0], th1[1]), make_double2(th1[2], th1[3]), pol_first);
    next_x<DEF>(a, rIc.x, rc.x, rc.y, rb.x, rb.y, ac->T[n & 1] + q, ac->Is[n & 1] + q, ri.z, w, n, th1, th2, xn);
    if (!(a.dbg & 256) || xn[0] == 12345.678) stg256(a.rx + sidx, make_double2(xn[0], xn[1]), make_double2(xn[2], xn[3]));
#else
    stg256_hint(a.rth + sidx, make_double2(th1[0], th1[1]), make_double2(th1[2], th1[3]), pol_first);
    next_x<DEF>(a, rIc.x, rc.x, rc.y, rb.x, rb.y, ac->T[n & 1] + q, ac->Is[n & 1] + q, ri.z, w, n, th1, th2, xn);
    stg256(a.rx + sidx, make_double2(xn[0], xn[1]), make_double2(xn[2], xn[3]));
#endif
}

template <int NG, bool DEF>
__device__ void junction_pass(const SubArgs& a, const AmpCache<8 * NG>* ac, int s, int c, long long n, bool do_post,
                              bool do_pre, const double* __restrict__ v) {
    constexpr int PC = 8 * NG, G = PC / 4;
    const int jlo = a.junc_ptr[s], jhi = a.junc_ptr[s + 1];
    const int total = (jhi - jlo) * G;
    if (!do_post) {
        // first boundary of a run: theta(n-1), theta(n-2) come from the canonical arrays; no solve result yet
        for (int idx = threadIdx.x; idx < total; idx += NT) {
            const int jp = jlo + idx / G, q = (idx % G) * 4, w = c * PC + q;
            if (w >= a.Wp) continue;
            const double2* rec = reinterpret_cast<const double2*>(a.jrec + 8 * (size_t)jp);
            const double2 rIc = __ldg(rec), rc = __ldg(rec + 1), rb = __ldg(rec + 2);
            const int jo = __ldg(reinterpret_cast<const int4*>(rec + 3)).z;
            const size_t sidx = ((size_t)c * a.Nj + jlo) * PC + (size_t)idx * 4;
            const size_t cidx = (size_t)jo * a.Wp + w;
            const double2* p1 = reinterpret_cast<const double2*>(a.th1 + cidx);
            const double2* p2 = reinterpret_cast<const double2*>(a.th2_in + cidx);
            const double2 u0 = p1[0], u1 = p1[1], v0 = p2[0], v1 = p2[1];
            double2* op = reinterpret_cast<double2*>(a.rth + sidx);
            op[0] = u0; op[1] = u1;
            const double th1[4] = {u0.x, u0.y, u1.x, u1.y}, th2[4] = {v0.x, v0.y, v1.x, v1.y};
            double xn[4];
            next_x<DEF>(a, rIc.x, rc.x, rc.y, rb.x, rb.y, ac->T[n & 1] + q, ac->Is[n & 1] + q, jo, w, n, th1, th2, xn);
            double2* xo = reinterpret_cast<double2*>(a.rx + sidx);
            xo[0] = make_double2(xn[0], xn[1]); xo[1] = make_double2(xn[2], xn[3]);
        }
        return;
    }
    int plane_th = -1, plane_I = -1;
    {
        const long long k = n - 1 - a.i0;
        if (a.th_plane) plane_th = (int)a.th_plane[k];
        if (a.I_plane) plane_I = (int)a.I_plane[k];
    }
    const size_t sbase = ((size_t)c * a.Nj + jlo) * PC;
    // software pipeline: the state of the next item is in flight (registers) while this one is computed, and the
    // state three iterations ahead is pulled into L2
    const unsigned long long pol_first = policy_evict_first();
    double2 x0, x1, t0, t1, rIc, rc, rb;
    int4 ri = make_int4(0, 0, 0, 0);
    x0 = x1 = t0 = t1 = rIc = rc = rb = make_double2(0.0, 0.0);
    if ((int)threadIdx.x < total) {
        const size_t si = sbase + (size_t)threadIdx.x * 4;
        const double4v xv = ldg256_hint(a.rx + si, pol_first), tv = ldg256_hint(a.rth + si, pol_first);
        x0 = xv.lo; x1 = xv.hi; t0 = tv.lo; t1 = tv.hi;
        const double2* rec = reinterpret_cast<const double2*>(a.jrec + 8 * (size_t)(jlo + threadIdx.x / G));
        rIc = __ldg(rec); rc = __ldg(rec + 1); rb = __ldg(rec + 2); ri = __ldg(reinterpret_cast<const int4*>(rec + 3));
    }
    for (int idx = threadIdx.x; idx < total; idx += NT) {
        if (idx + 3 * NT < total) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rx + sbase + (size_t)(idx + 3 * NT) * 4));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rth + sbase + (size_t)(idx + 3 * NT) * 4));
        }
        double2 nx0, nx1, nt0, nt1, nIc, nc, nb;
        int4 ni = make_int4(0, 0, 0, 0);
        nx0 = nx1 = nt0 = nt1 = nIc = nc = nb = make_double2(0.0, 0.0);
#ifdef JJ_EXPERIMENTS
        if (idx + NT < total && !(a.dbg & 256)) {
#else
        if (idx + NT < total) {
#endif
            const size_t si = sbase + (size_t)(idx + NT) * 4;
            const double4v xv = ldg256_hint(a.rx + si, pol_first), tv = ldg256_hint(a.rth + si, pol_first);
            nx0 = xv.lo; nx1 = xv.hi; nt0 = tv.lo; nt1 = tv.hi;
            const double2* rec = reinterpret_cast<const double2*>(a.jrec + 8 * (size_t)(jlo + (idx + NT) / G));
            nIc = __ldg(rec); nc = __ldg(rec + 1); nb = __ldg(rec + 2); ni = __ldg(reinterpret_cast<const int4*>(rec + 3));
        }
        junction_item<NG, DEF>(a, ac, c, n, idx, jlo, do_pre, v, plane_th, plane_I, x0, x1, t0, t1, rIc, rc, rb, ri);
        x0 = nx0; x1 = nx1; t0 = nt0; t1 = nt1; rIc = nIc; rc = nc; rb = nb; ri = ni;
    }
}

// b = A (x'/c0 - theta_s) - 2 pi f into shared memory (reference: time_evolution.py:560-569). Every row has a
// fixed-width list of signed junction entries (2 * device junction + (sign < 0), -1 = absent) and the x stream holds
// x'/c0, so a row is a signed sum of gathered values; halo rows hold the partial sum over the junctions this subdomain
// owns (their flux term is added by the top assembly).
// The pass is a chain of dependent L2 round trips (row list -> gathers), not bandwidth (a lone SM streams L2 at
// 80 B/clk, tools/mb_l2stream.cu; the pass moves 30): no coefficient loads travel with the gathers, and the flux base
// of a row comes from a per-row table gathered once per launch (rowF) instead of two dependent loads behind them.
// (Fetching the row lists one chunk ahead brought the pass from 24 k to 16 k cycles per time step on cfg2, but its 12
// extra registers pushed loop invariants of the junction pass into local memory: +11 k cycles there. Not kept.)
template <int NG>
__device__ void face_pass(const SubArgs& a, const AmpCache<8 * NG>* ac, int s, int c, long long n, int rows_used,
                          double* __restrict__ v) {
    constexpr int PC = 8 * NG, G = PC / 4;
    constexpr int U = 2;                  // rows per thread and iteration: 8 independent 32-byte gathers in flight
    const int K = a.face_K;
    const int* fj = a.face_ell_j + (size_t)s * a.n_rows * K;
    const double* rowF = a.rowF ? a.rowF + (size_t)s * a.n_rows : nullptr;
    const int total = rows_used * G;
    const bool has_vs = a.Vs.kind == KIND_RANK1;
    const double* xbase = a.rx + (size_t)c * a.Nj * PC;
    for (int idx0 = threadIdx.x; idx0 < total; idx0 += U * NT) {
        double acc[U][4];
        double fv[U];
        int row[U], q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = idx0 + u * NT;
            row[u] = idx < total ? idx / G : 0;
            q[u] = (idx % G) * 4;
            acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.0;
            fv[u] = rowF ? __ldg(rowF + row[u]) : 0.0;        // (independent of the gathers: in flight with them)
        }
        for (int k0 = 0; k0 < K; k0 += 4) {
            int je[U][4]; double2 xa[U][4], xb[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int4 j4 = __ldg(reinterpret_cast<const int4*>(fj + (size_t)row[u] * K + k0));
                je[u][0] = j4.x; je[u][1] = j4.y; je[u][2] = j4.z; je[u][3] = j4.w;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // absent entries (-1) read junction 0 and are not added; x'/c0 was written by this block (plain loads)
                    const double4v xv = ldg256(xbase + q[u] + (size_t)(max(je[u][k], 0) >> 1) * PC);
                    xa[u][k] = xv.lo; xb[u][k] = xv.hi;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (je[u][k] < 0) continue;
                    if (je[u][k] & 1) {
                        acc[u][0] -= xa[u][k].x; acc[u][1] -= xa[u][k].y; acc[u][2] -= xb[u][k].x; acc[u][3] -= xb[u][k].y;
                    } else {
                        acc[u][0] += xa[u][k].x; acc[u][1] += xa[u][k].y; acc[u][2] += xb[u][k].x; acc[u][3] += xb[u][k].y;
                    }
                }
                if (has_vs) {
                    const double* cum = ac->Vs[n & 1] + q[u];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (je[u][k] >= 0) {
                            // - sign * (Vs base) * (running sum of the Vs amplitude)
                            const double vb = __ldg(a.P1 + 4 * (size_t)(je[u][k] >> 1) + 2);
                            const double sv = (je[u][k] & 1) ? -vb : vb;
                            for (int e = 0; e < 4; ++e) acc[u][e] -= sv * cum[e];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (idx0 + u * NT >= total) continue;
            if (c * PC + q[u] < a.Wp) {
                const double* am = ac->F[n & 1] + q[u];            // (zeros unless the flux is rank one; fv is 0 for halo rows)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[u][k] -= TWO_PI * (fv[u] * am[k]);
            } else {
                acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.0;
            }
            double2* dst = reinterpret_cast<double2*>(v + velem<NG>(row[u], q[u]));
            dst[0] = make_double2(acc[u][0], acc[u][1]);
            dst[1] = make_double2(acc[u][2], acc[u][3]);
        }
    }
}

// One OBSERVATION (running observables, jj_observe_begin): add n = -A round(theta / 2 pi) of the step just finished
// into the per-(face, problem) sums and leave its phases in the mark planes, straight from the state slab this block
// has just written - no theta plane is stored. Same row lists as the face pass (the coefficient's sign is the cycle
// matrix entry); local rows belong to this block alone (plain read-modify-write, 16 bytes per thread), halo rows
// receive one partial sum per subdomain that owns junctions of the face (integer atomics: exact in any order). True
// division and round-to-nearest-even, like np.round(theta / (2 pi)) (reference: time_evolution.py:734-755).
// (not inlined, and called with plain values instead of the argument block: its registers must not weigh on the step
// kernel, which runs it only at observed steps)
struct VortexView {
    const int* fj; const int* fidx; const int* ht; const int* top_face;
    const double* th; const double* jrec;
    int* nsum; double* th_last; double* th_first;       // th_first: null except at the first observation
    int K, rows, nl, jlo, jhi, c, Wp;
};

template <int NG>
__device__ __noinline__ void vortex_pass(const VortexView o) {
    constexpr int PC = 8 * NG, G = PC / 4;
    for (int idx = threadIdx.x; idx < o.rows * G; idx += NT) {
        const int row = idx / G, q = (idx % G) * 4, w = o.c * PC + q;
        if (w >= o.Wp) continue;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k0 = 0; k0 < o.K; k0 += 4) {
            const int4 j4 = __ldg(reinterpret_cast<const int4*>(o.fj + (size_t)row * o.K + k0));
            const int je[4] = {j4.x, j4.y, j4.z, j4.w};
            double4v t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = ldg256(o.th + (size_t)(max(je[k], 0) >> 1) * PC + q);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (je[k] < 0) continue;
                const double sg = (je[k] & 1) ? -1.0 : 1.0;
                acc[0] -= sg * phase_zone(t[k].lo.x); acc[1] -= sg * phase_zone(t[k].lo.y);
                acc[2] -= sg * phase_zone(t[k].hi.x); acc[3] -= sg * phase_zone(t[k].hi.y);
            }
        }
        if (row < o.nl) {
            int4* dst = reinterpret_cast<int4*>(o.nsum + (size_t)__ldg(o.fidx + row) * o.Wp + w);
            int4 v4 = *dst;
            v4.x += (int)acc[0]; v4.y += (int)acc[1]; v4.z += (int)acc[2]; v4.w += (int)acc[3];
            *dst = v4;
        } else {
            int* dst = o.nsum + (size_t)__ldg(o.top_face + __ldg(o.ht + (row - o.nl))) * o.Wp + w;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if ((int)acc[e] != 0) atomicAdd(dst + e, (int)acc[e]);
        }
    }
    // phase marks of the junctions this subdomain owns, canonical [Nj][Wp] by original junction index
    for (int idx = threadIdx.x; idx < (o.jhi - o.jlo) * G; idx += NT) {
        const int jp = o.jlo + idx / G, q = (idx % G) * 4, w = o.c * PC + q;
        if (w >= o.Wp) continue;
        const int jo = __ldg(reinterpret_cast<const int4*>(o.jrec + 8 * (size_t)jp + 6)).z;
        // (plain 16-byte accesses: ptxas 12.9 narrows a 256-bit load/store pair that only copies to its first 8 bytes)
        const double2* src = reinterpret_cast<const double2*>(o.th + (size_t)jp * PC + q);
        const double2 lo = src[0], hi = src[1];
        double2* dl = reinterpret_cast<double2*>(o.th_last + (size_t)jo * o.Wp + w);
        dl[0] = lo; dl[1] = hi;
        if (o.th_first) {
            double2* df = reinterpret_cast<double2*>(o.th_first + (size_t)jo * o.Wp + w);
            df[0] = lo; df[1] = hi;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// grid barrier: monotonic arrival counter (zeroed by the host before the launch); the kernel is launched
// cooperatively, so all blocks are resident.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned& target, unsigned members) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += members;
        // arrive with release semantics (cumulative over the writes of the whole block, which the __syncthreads above
        // ordered before this thread) without waiting for the atomic's round trip; the acquire load of the poll is
        // all the consumer side needs (cross-block data is read with ld.cg / cp.async.cg / after this point only)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) { group_barrier(ctr, target, gridDim.x); }

// r_top = sum of the subdomain contributions - 2 pi f_top (+ debug right-hand side), logical [chunk][row][PC]
template <int NG>
__device__ void top_assemble(const SubArgs& a, long long n, int c0, int nc, int worker, int n_workers) {
    // chunks [c0, c0 + nc) are assembled by n_workers thread blocks; this block is number `worker` of them
    constexpr int PC = 8 * NG, G = PC / 4;
    const long long total = (long long)nc * a.n_top * G;
    for (long long idx = (long long)worker * NT + threadIdx.x; idx < total; idx += (long long)n_workers * NT) {
        const int q = (int)(idx % G) * 4;
        const long long t = idx / G;
        const int k = (int)(t % a.n_top), c = c0 + (int)(t / a.n_top);
        const int w = c * PC + q;
        double acc[4] = {0, 0, 0, 0};
        if (a.tslot4) {
            // one 16-byte load names all slots of the row: two dependent memory levels instead of three; the slots
            // are added in the same (ascending) order as the list walk below
            const int4 s4 = __ldg(a.tslot4 + k);
            const int sl[4] = {s4.x, s4.y, s4.z, s4.w};
            double2 u0[4], u1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double2* p = reinterpret_cast<const double2*>(a.ctop + ((size_t)c * a.n_slots + max(sl[e], 0)) * PC + q);
                u0[e] = __ldcg(p); u1[e] = __ldcg(p + 1);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (sl[e] >= 0) { acc[0] += u0[e].x; acc[1] += u0[e].y; acc[2] += u1[e].x; acc[3] += u1[e].y; }
        } else {
            for (int sl = a.tptr[k]; sl < a.tptr[k + 1]; ++sl) {
                const double2* p = reinterpret_cast<const double2*>(a.ctop + ((size_t)c * a.n_slots + a.tslot[sl]) * PC + q);
                const double2 u0 = __ldcg(p), u1 = __ldcg(p + 1);
                acc[0] += u0.x; acc[1] += u0.y; acc[2] += u1.x; acc[3] += u1.y;
            }
        }
        if (w < a.Wp) {
            const int g = a.top_face[k];
            if (a.dbg_b) {
                for (int e = 0; e < 4; ++e) acc[e] += a.dbg_b[(size_t)g * a.Wp + w + e];
            } else if (a.F.kind == KIND_RANK1) {
                const double b = a.topF ? __ldg(a.topF + k) : __ldg(a.F.base + g);
                const double* am = a.F.table + source_row(a.F, n) * a.Wp + w;
                for (int e = 0; e < 4; ++e) acc[e] -= TWO_PI * (b * __ldg(am + e));
            }
        }
        double2* dst = reinterpret_cast<double2*>(a.rtop + ((size_t)c * a.n_up_pad + k) * PC + q);
        dst[0] = make_double2(acc[0], acc[1]);
        dst[1] = make_double2(acc[2], acc[3]);
    }
}

// J_tt = S_tt^-1 r_tt, direct version (no shared-memory staging; used when the staging rows are too few).
// Task = (chunk, 8-row tile, problem group), one per warp and round; consecutive tasks differ in the group
// first, so the warps of a block share A row tiles and B fragments through L1.
template <int NG>
__device__ void top_product_direct(const SubArgs& a) {
    constexpr int PC = 8 * NG;
    const int RT = (a.n_tt + 7) / 8, KS = a.n_tt_pad / 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tasks = a.n_chunks * RT * NG;
    for (int task = blockIdx.x * NWARPS + warp; task < n_tasks; task += gridDim.x * NWARPS) {
        const int g = task % NG, rt = (task / NG) % RT, c = task / (NG * RT);
        const double* A = a.SinvP + ((size_t)rt * KS) * 32 + lane;
        const double* B = a.rtop + ((size_t)c * a.n_up_pad + a.tt0 + (lane & 3)) * PC + 8 * g + (lane >> 2);
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        for (int ks0 = 0; ks0 < KS; ks0 += 8) {
            double ac[8], bc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { ac[j] = __ldg(A + (size_t)(ks0 + j) * 32); bc[j] = __ldcg(B + (size_t)(ks0 + j) * 4 * PC); }
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                dmma884(c00, c01, ac[j], bc[j]);
                dmma884(c10, c11, ac[j + 1], bc[j + 1]);
            }
        }
        const int row = 8 * rt + (lane >> 2);
        double2* dst = reinterpret_cast<double2*>(a.jtop + ((size_t)c * a.n_up_pad + a.tt0 + row) * PC + 8 * g + 2 * (lane & 3));
        *dst = make_double2(c00 + c10, c01 + c11);
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}

// ---- bulk copies through the TMA unit (cp.async.bulk, 1-D): one thread moves a whole contiguous block between global
// and shared memory; the other threads only wait on an mbarrier (loads) or carry on (stores). Used for the solve
// vector of the local rows, which passes through global memory between the forward and the backward sweep whenever a
// block works through several (subdomain, chunk) items per time step.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// global -> shared, completion counted in bytes on the mbarrier (bytes: multiple of 16, below 2^20)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// J_tt = S_tt^-1 r_tt (dense top of the top) with the A fragments (the packed inverse) read straight from L2 into
// registers: a fragment is used by exactly one warp, so staging it in shared memory only costs shared-memory bandwidth.
// Only the B fragments (r_tt, shared by all row tiles of the block) go through a cp.async ring in the staging rows
// (idle between the sweeps); a stage holds KB = KQ * KM k-steps; warp = (row tile, K slot): the K range is split over
// otherwise idle warps, the KQ partial sums of a row tile meet in shared memory and are added in a fixed order. The A
// fragments of the next stage are in flight while this stage is multiplied.
template <int NG, int KM>
__device__ void top_product_areg(const SubArgs& a, double* buf, int RB, int KQ, int S, int c0, int nc, int worker,
                                 int n_workers) {
    constexpr int PC = 8 * NG;
    constexpr int GW = NG < 4 ? NG : 4;
    constexpr int GP = (NG + 3) / 4;
    const int KB = KQ * KM;
    const int RT = (a.n_tt + 7) / 8, RTP = a.n_tt_pad / 8, KS = a.n_tt_pad / 4;
    const int NB = (RT + RB - 1) / RB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rtl = warp / KQ, kq = warp % KQ;
    const int stage_doubles = GW * KB * 32;
    const int pieces = stage_doubles / 2;            // <= NT: one 16-byte piece per thread and stage
    const int n_tasks = nc * NB * GP;
    const int nkb = KS / KB;
    const unsigned long long pol = policy_evict_last();
    for (int task = worker; task < n_tasks; task += n_workers) {
        // chunk-minor task order: blocks that run at the same time multiply the SAME row block of the inverse by
        // different problem chunks, so its A fragments are read from HBM once and hit in L2 for the other chunks
        // (large circuits: the packed inverse is hundreds of MB, far more than the L2)
        const int c = c0 + task % nc, nb = (task / nc) % NB, gp = task / (nc * NB);
        const int rt = nb * RB + rtl;
        const bool active = rtl < RB && rt < RT;
        // B fragments are staged as [kk][n ^ 4*(kk>>1)]: a half warp (n = 0..3 or 4..7, kk = 0..3) then reads 16
        // doubles that fall into 16 different 8-byte banks
        const int i = threadIdx.x;
        const bool have = i < pieces;
        const int frag = i / 16, piece = i % 16, gsel = frag / KB, kl = frag % KB;
        const int gg = min(4 * gp + gsel, NG - 1), kk = piece >> 2, n2 = (piece & 3) * 2;
        const double* sp = a.rtop + ((size_t)c * a.n_up_pad + a.tt0 + 4 * kl + kk) * PC + 8 * gg + n2;
        const size_t sstep = (size_t)KB * 4 * PC;
        const int dofs = frag * 32 + kk * 8 + (n2 ^ ((kk >> 1) << 2));
        int issued = 0, islot = 0;
        auto issue = [&]() {
            if (issued < nkb) {
                if (have) cp_async16(buf + (size_t)islot * stage_doubles + dofs, sp);
                sp += sstep;
                ++issued;
                if (++islot == S) islot = 0;
            }
            asm volatile("cp.async.commit_group;");
        };
        const double* ap = a.SinvP + ((size_t)min(rt, RTP - 1) * KS + kq) * 32 + lane;
        double an[KM];
#pragma unroll
        for (int m = 0; m < KM; ++m) {
            asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(an[m]) : "l"(ap + (size_t)m * KQ * 32), "l"(pol));
        }
        __syncthreads();                           // the ring is free (previous task / phase done)
        for (int k = 0; k < S - 1; ++k) issue();
        double acc[GW][2];
#pragma unroll
        for (int g = 0; g < GW; ++g) { acc[g][0] = 0.0; acc[g][1] = 0.0; }
        const int b_off = kq * 32 + (lane & 3) * 8 + ((lane >> 2) ^ (((lane & 3) >> 1) << 2));   // group g: + g*KB*32
        const double* st = buf;
        int cslot = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            if (S >= 4) asm volatile("cp.async.wait_group 2;");
            else if (S == 3) asm volatile("cp.async.wait_group 1;");
            else asm volatile("cp.async.wait_group 0;");
            __syncthreads();
            issue();
            double av[KM];
#pragma unroll
            for (int m = 0; m < KM; ++m) av[m] = an[m];
            ap += (size_t)KB * 32;
            if (kb + 1 < nkb) {
#pragma unroll
                for (int m = 0; m < KM; ++m)
                    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(an[m]) : "l"(ap + (size_t)m * KQ * 32), "l"(pol));
            }
            if (active) {
#pragma unroll
                for (int m = 0; m < KM; ++m) {
                    double bv[GW];
#pragma unroll
                    for (int g = 0; g < GW; ++g) bv[g] = st[b_off + m * KQ * 32 + g * KB * 32];
#pragma unroll
                    for (int g = 0; g < GW; ++g) dmma884(acc[g][0], acc[g][1], av[m], bv[g]);
                }
            }
            st += stage_doubles;
            if (++cslot == S) { cslot = 0; st = buf; }
        }
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();                           // everyone is done with the ring: reuse it for the partial sums
        if (active) {
#pragma unroll
            for (int g = 0; g < GW; ++g)
                *reinterpret_cast<double2*>(buf + ((size_t)(warp * GW + g) * 32 + lane) * 2) = make_double2(acc[g][0], acc[g][1]);
        }
        __syncthreads();
        if (active && kq == 0) {
            const int row = 8 * rt + (lane >> 2);
#pragma unroll
            for (int g = 0; g < GW; ++g) {
                const int gg2 = 4 * gp + g;
                if (gg2 < NG) {
                    double2 sum = make_double2(0.0, 0.0);
                    for (int k2 = 0; k2 < KQ; ++k2) {
                        const double2 part = *reinterpret_cast<const double2*>(buf + ((size_t)((warp + k2) * GW + g) * 32 + lane) * 2);
                        sum.x += part.x; sum.y += part.y;
                    }
                    double2* dst = reinterpret_cast<double2*>(a.jtop + ((size_t)c * a.n_up_pad + a.tt0 + row) * PC + 8 * gg2 + 2 * (lane & 3));
                    *dst = sum;
                }
            }
        }
    }
    __syncthreads();
}

// A task small enough for one warp (at most 4 tiles, a few hundred A fragments): no shared memory, no block barrier.
// The B fragments come straight from L2 (ld.global.cg: the rows were written by other blocks in earlier phases), two
// row tiles per pass; the warps of a block work on different (task, chunk) pairs at the same time, which hides the
// dependent loads (column code -> row address -> fragment) that dominate tasks of this size.
template <int NG>
__device__ void upper_warp_task(const SubArgs& a, int t, int c, int half, long long* stamp) {
    constexpr int PC = 8 * NG;
    const int lane = threadIdx.x & 31, n = lane >> 2, kk = lane & 3;
    const long long ts0 = stamp ? clock64() : 0;
    const int4 hd = __ldg(a.up_task + t);
    const int tiles = (hd.y + 7) >> 3, nk = hd.z;
    if (stamp && nk > 0 && lane == 0) stamp[0] += clock64() - ts0;          // header arrived
    const size_t plane_elems = (size_t)a.n_chunks * a.n_up_pad * PC;
    const int* cols = a.up_cols + hd.w;
    const double* ubase = a.U + (size_t)c * a.n_up_pad * PC + n;
    const double* abase = a.up_A + __ldg(a.up_task_aoff + t) + lane;
    double* obase = a.U + (size_t)(hd.x >> 28) * plane_elems + ((size_t)c * a.n_up_pad + (hd.x & 0xfffffff)) * PC;
    // pull the item's A fragments and column codes into L2 now (one memory latency for all of them): the loop below
    // is a chain of dependent loads, and everything outside the r / z / J planes was evicted by the state streams
    {
        const int t0p = min(2 * half, tiles - 1);
        const char* pa = reinterpret_cast<const char*>(abase - lane + (size_t)t0p * nk * 32);
        const int a_bytes = min(2, tiles - t0p) * nk * 256;
        for (int o = lane * 128; o < a_bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + o));
        const char* pc = reinterpret_cast<const char*>(cols);
        for (int o = lane * 128; o < nk * 16; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
    }
    // a warp takes ONE pair of row tiles (half = 0: tiles 0, 1; half = 1: tiles 2, 3); the two halves of a task run in
    // different warps at the same time (in an in-place phase a half may read rows the other half is rewriting, but
    // only with coefficients that are exactly zero)
    {
        const int t0 = 2 * half;
        if (t0 >= tiles) return;
        const bool two = t0 + 1 < tiles;
        const double* a0 = abase + (size_t)t0 * nk * 32;
        const double* a1 = abase + (size_t)(two ? t0 + 1 : t0) * nk * 32;
        double acc0[NG][2], acc1[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) { acc0[g][0] = acc0[g][1] = acc1[g][0] = acc1[g][1] = 0.0; }
        // the column codes of 8 k-steps arrive in one coalesced load (nk is a multiple of 16) and are handed to the
        // lanes by shuffles; the codes of the next 8 k-steps are in flight meanwhile. Four k-steps of fragments are
        // fetched before their MMAs issue: the task is a chain of dependent latencies, not of arithmetic.
        int code_next = __ldg(cols + lane);
        if (stamp && code_next != 0x7fffffff && lane == 0) stamp[1] += clock64() - ts0;     // first column codes arrived
        for (int k0 = 0; k0 < nk; k0 += 8) {
            const int code32 = code_next;
            if (k0 + 8 < nk) code_next = __ldg(cols + 4 * (k0 + 8) + lane);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double av0[4], av1[4], bv[4][NG];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int k = k0 + 4 * h + m;
                    const int code = __shfl_sync(0xffffffffu, code32, 4 * (4 * h + m) + kk);
                    const double* bp = ubase + (size_t)(code >> 28) * plane_elems + (size_t)(code & 0xfffffff) * PC;
                    av0[m] = __ldg(a0 + (size_t)k * 32);
                    av1[m] = __ldg(a1 + (size_t)k * 32);
#pragma unroll
                    for (int g = 0; g < NG; ++g) bv[m][g] = __ldcg(bp + 8 * g);
                }
#pragma unroll
                for (int m = 0; m < 4; ++m) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) { dmma884(acc0[g][0], acc0[g][1], av0[m], bv[m][g]); dmma884(acc1[g][0], acc1[g][1], av1[m], bv[m][g]); }
                }
            }
        }
        if (stamp && acc0[0][0] != 1.2345e300 && lane == 0) stamp[2] += clock64() - ts0;  // all products done
        __syncwarp();            // all fragments of this pass are read before rows of the task are rewritten (in-place phases)
        const int row = 8 * t0 + n;
        if (row < hd.y) {
#pragma unroll
            for (int g = 0; g < NG; ++g) *reinterpret_cast<double2*>(obase + (size_t)row * PC + 8 * g + 2 * kk) = make_double2(acc0[g][0], acc0[g][1]);
        }
        if (two && row + 8 < hd.y) {
#pragma unroll
            for (int g = 0; g < NG; ++g) *reinterpret_cast<double2*>(obase + (size_t)(row + 8) * PC + 8 * g + 2 * kk) = make_double2(acc1[g][0], acc1[g][1]);
        }
    }
}

// One phase of the upper program (the separators between the subdomains and the dense top of the top): every task is
// a gathered dense product  out[rows] = V . X[cols]  over the r / z / J planes of the separator rows, for up to 16
// 8-row tiles and one problem chunk. Tasks of a phase are independent (a grid barrier ends the phase).
//
// BLOCK tasks [phase_ptr, phase_split): the X rows named by the task's column list (row codes plane << 28 | row) are
// gathered into shared memory in PANELS (cp.async, two buffers: the next panel lands while this one is multiplied),
// in the swizzled row layout of the local sweeps; warp = (row tile, K slot): a task with T tiles splits the k-steps of
// a panel over KQ = 16 / pow2ceil(T) warps per tile, the A fragments come straight from global memory (each is used
// by exactly one warp), and the KQ partial sums of a tile meet in shared memory in a fixed order. (task, chunk) pairs
// are handed out by a work counter in decreasing order of cost (the host sorts them), chunk-minor: the blocks that run
// at the same time apply the SAME A fragments to different problem chunks, so these come from HBM once and hit in
// L2 afterwards. Which block runs a pair does not change its arithmetic: results do not depend on the schedule.
// WARP tasks [phase_split, phase_ptr next): see upper_warp_task; dealt to the warps of the grid round-robin.
// buf: the whole dynamic shared memory in front of the amplitude cache (the local vector is idle: with upper phases z
// always goes through global memory), buf_rows: rows of PC float64 per panel buffer (a multiple of 256).
template <int NG>
__device__ void upper_phase(const SubArgs& a, double* buf, int buf_rows, int phase, long long step) {
    constexpr int PC = 8 * NG;
    constexpr int PPR = PC / 2;                        // 16-byte pieces per row
    __shared__ int s_pair;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tph0 = a.prof ? clock64() : 0;
    const int t_lo = __ldg(a.up_phase_ptr + phase), t_split = __ldg(a.up_phase_split + phase), t_hi = __ldg(a.up_phase_ptr + phase + 1);
    const int n_pairs = (t_split - t_lo) * a.n_chunks;
    const size_t plane_elems = (size_t)a.n_chunks * a.n_up_pad * PC;
    const unsigned long long pol = policy_evict_last();
    const int KP = buf_rows / 4;                       // k-steps per panel (a multiple of 16)
    const int n = lane >> 2, kk = lane & 3;
    // every block draws pairs until it draws one past the end: n_pairs + gridDim.x draws per time step
    const unsigned base = (unsigned)(step * (long long)(n_pairs + (int)gridDim.x));
    unsigned drawn = 0;
    if (n_pairs > 0 && threadIdx.x == 0) drawn = atomicAdd(a.up_ctr + phase, 1u) - base;
    while (n_pairs > 0) {
        if (threadIdx.x == 0) s_pair = (int)drawn;
        __syncthreads();                               // the pair is known; the buffers are free (previous task done)
        const int p = s_pair;
        if (p >= n_pairs) break;
        if (threadIdx.x == 0) drawn = atomicAdd(a.up_ctr + phase, 1u) - base;       // the next pair, while this one runs
        const int c = p % a.n_chunks, t = t_lo + p / a.n_chunks;
        const int4 hd = __ldg(a.up_task + t);          // out code, rows, k-steps, first column
        const int tiles = (hd.y + 7) >> 3, nk = hd.z;
        int rbt = NG >= 8 ? 2 : 1;                     // (64 problems per chunk: panels of 128 rows, at most 8 K slots)
        while (rbt < tiles) rbt <<= 1;
        const int KQ = NWARPS / rbt;
        const int rtl = warp / KQ, kq = warp % KQ;
        const bool active = rtl < tiles;
        const int n_panels = (nk + KP - 1) / KP;
        const int* cols = a.up_cols + hd.w;
        const double* ubase = a.U + (size_t)c * a.n_up_pad * PC;
        auto issue_panel = [&](int pp) {
            const int r0 = pp * buf_rows, nr = min(nk * 4 - r0, buf_rows);
            double* dst = buf + (size_t)(pp & 1) * buf_rows * PC;
            // a thread copies at most 8 pieces of a panel (buf_rows * PPR <= 8 * NT): all its row codes are fetched first
            // (independent loads, one memory latency), then the copies are issued
            constexpr int MAXP = 8;
            int code[MAXP];
#pragma unroll
            for (int i = 0; i < MAXP; ++i) {
                const int e = threadIdx.x + i * NT;
                code[i] = e < nr * PPR ? __ldg(cols + r0 + e / PPR) : 0;
            }
#pragma unroll
            for (int i = 0; i < MAXP; ++i) {
                const int e = threadIdx.x + i * NT;
                if (e < nr * PPR) {
                    const int r = e / PPR, q = (e % PPR) * 2;
                    cp_async16(dst + velem<NG>(r, q), ubase + (size_t)(code[i] >> 28) * plane_elems + (size_t)(code[i] & 0xfffffff) * PC + q);
                }
            }
            asm volatile("cp.async.commit_group;");
        };
        issue_panel(0);
        double acc[NG][2];
#pragma unroll
        for (int g = 0; g < NG; ++g) { acc[g][0] = 0.0; acc[g][1] = 0.0; }
        // this warp's A fragments: k-steps kq, kq + KQ, ... of row tile rtl, through a register ring that stays UPR
        // k-steps ahead across panel boundaries (nk is a multiple of 64 and a panel holds a multiple of 64 k-steps, so
        // a warp's share of a panel is a whole number of ring turns; reading past the task's end is harmless: the
        // array is padded and the values are never used)
        constexpr int UPR = 4;
        const double* ap = a.up_A + __ldg(a.up_task_aoff + t) + ((size_t)min(rtl, tiles - 1) * nk + kq) * 32 + lane;
        const size_t astep = (size_t)KQ * 32;
        double ra[UPR];
#pragma unroll
        for (int u = 0; u < UPR; ++u)
            asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(ra[u]) : "l"(ap + (size_t)u * astep), "l"(pol));
        for (int pp = 0; pp < n_panels; ++pp) {
            if (pp + 1 < n_panels) { issue_panel(pp + 1); asm volatile("cp.async.wait_group 1;"); }
            else asm volatile("cp.async.wait_group 0;");
            __syncthreads();                           // panel pp has landed for everyone
            if (active) {
                const int own = min(nk - pp * KP, KP) / KQ;          // k-steps of this warp in the panel
                const unsigned vb = (unsigned)__cvta_generic_to_shared(buf + (size_t)(pp & 1) * buf_rows * PC);
                // B fragment of local k-step j: rows 4 j + kk (row & 3 == kk), problems 8 g + n
                unsigned baddr = vb + (unsigned)(velem<NG>(4 * kq + kk, n) * 8);
                const unsigned bstep = (unsigned)(4 * KQ * PC * 8);
                for (int j = 0; j < own; j += UPR) {
#pragma unroll
                    for (int u = 0; u < UPR; ++u) {
                        double bv[NG];
#pragma unroll
                        for (int g = 0; g < NG; ++g) bv[g] = lds_f64(baddr ^ (unsigned)(g << 6));
#pragma unroll
                        for (int g = 0; g < NG; ++g) dmma884(acc[g][0], acc[g][1], ra[u], bv[g]);
                        asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(ra[u]) : "l"(ap + (size_t)(UPR + u) * astep), "l"(pol));
                        baddr += bstep;
                    }
                    ap += UPR * astep;
                }
            }
            __syncthreads();                           // everyone is done with buffer pp & 1
        }
        // partial sums of the K slots meet in shared memory (buffer 0 is free), fixed order
        if (active && KQ > 1) {
#pragma unroll
            for (int g = 0; g < NG; ++g)
                *reinterpret_cast<double2*>(buf + ((size_t)(warp * NG + g) * 32 + lane) * 2) = make_double2(acc[g][0], acc[g][1]);
        }
        if (KQ > 1) __syncthreads();
        const int row = 8 * rtl + n;
        if (active && kq == 0 && row < hd.y) {
            double* obase = a.U + (size_t)(hd.x >> 28) * plane_elems + ((size_t)c * a.n_up_pad + (hd.x & 0xfffffff) + row) * PC;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                double2 sum = make_double2(acc[g][0], acc[g][1]);
                for (int k2 = 1; k2 < KQ; ++k2) {
                    const double2 part = *reinterpret_cast<const double2*>(buf + ((size_t)((warp + k2) * NG + g) * 32 + lane) * 2);
                    sum.x += part.x; sum.y += part.y;
                }
                *reinterpret_cast<double2*>(obase + 8 * g + 2 * kk) = sum;
            }
        }
    }
    // warp tasks: (task, chunk, pair of row tiles) items; consecutive items go to different BLOCKS (all SMs busy even
    // when there are fewer items than warps)
    const long long n_witems = (long long)(t_hi - t_split) * a.n_chunks * 2;
    for (long long p = (long long)warp * gridDim.x + blockIdx.x; p < n_witems; p += (long long)gridDim.x * NWARPS) {
        const long long q = p >> 1;
        long long* stamp = (a.prof && blockIdx.x == 0 && warp == 0 && phase == 0) ? a.prof + 8 + 48 + 2 * PROF_UP : nullptr;
        upper_warp_task<NG>(a, t_split + (int)(q / a.n_chunks), (int)(q % a.n_chunks), (int)(p & 1), stamp);
        if (stamp && lane == 0) stamp[3] += clock64() - tph0;
    }
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0 && phase == 0) a.prof[8 + 48 + 2 * PROF_UP + 4] += clock64() - tph0;
    __syncthreads();
}

template <int NG>
__device__ __forceinline__ void load_prog(const SubArgs& a, int s, ProgSmem& ps, int* aux) {
    const SubProgDev p = a.prog[s];
    int* wt = aux; int* ws = aux + 2 * a.max_np; int* ls = ws + a.max_np;
    int2* th = reinterpret_cast<int2*>(ls + a.max_levels + (a.max_levels & 1));
    const int np = p.n_levels * NWARPS;
    for (int e = threadIdx.x; e < 2 * np; e += NT) wt[e] = p.wt_ptr[e];
    for (int e = threadIdx.x; e < np; e += NT) ws[e] = p.ws_ptr[e];
    for (int e = threadIdx.x; e < p.n_levels; e += NT) ls[e] = p.lstaged[e];
    for (int e = threadIdx.x; e < p.n_tiles; e += NT) th[e] = p.thdr[e];
    ps.wt = wt; ps.ws = ws; ps.lstaged = ls; ps.thdr = th; ps.stream = p.stream;
    ps.n_levels = p.n_levels; ps.n_bwd = p.n_bwd;
}

// copy between the shared-memory vector (swizzled rows) and logical [row][PC] global arrays
template <int NG>
__device__ __forceinline__ void rows_to_global(const double* v, int row0, int nrows, double* dst) {
    constexpr int PC = 8 * NG;
    for (int e = threadIdx.x; e < nrows * (PC / 2); e += NT) {
        const int r = e / (PC / 2), q = (e % (PC / 2)) * 2;
        reinterpret_cast<double2*>(dst + (size_t)r * PC)[q >> 1] = *reinterpret_cast<const double2*>(v + velem<NG>(row0 + r, q));
    }
}

// UPPER = false is the LEAN kernel: at most one (subdomain, chunk) item per block and no upper phases, z stays in shared
// memory; UPPER = true is the general one (several items per block handed out by the work counter, z through global
// memory by bulk copies, upper phases - possibly none).
// UPPER: the plan has upper phases (compiled out otherwise: their registers and code must not weigh on the kernel of
// circuits whose separators all fit the dense top product, whose time step is a tenth as long)
template <int NG, bool DEF, bool UPPER>
__global__ void __launch_bounds__(NT, BLOCKS_PER_SM) k_subdomain(const SubArgs a) {
    constexpr int PC = 8 * NG;
    extern __shared__ __align__(1024) double smem[];
    double* v = smem;
    double* stage = v + (size_t)a.n_rows * PC;
    AmpCache<PC>* ac = reinterpret_cast<AmpCache<PC>*>(stage + (size_t)a.stage_rows * (PC + 2));
    int* aux = reinterpret_cast<int*>(ac + 1);
    ProgSmem ps;
    long long* const prof_p = a.prof;       // in-kernel phase counters (JJ_SUB_PROF), null in normal runs
    const int n_items = a.P * a.n_chunks;
    // More items than blocks: a block takes a contiguous range of the subdomain-major item list, so its consecutive items
    // are chunks of the SAME subdomain (its program stays loaded, its factor stream is hot in L2, and the neighbouring
    // blocks stream the same factor at the same time). Otherwise block b owns item b = (chunk b / P, subdomain b % P).
    const bool multi = UPPER && n_items > (int)gridDim.x;
    const int it_lo = multi ? (int)((long long)blockIdx.x * n_items / gridDim.x) : (int)blockIdx.x;
    const int it_hi = multi ? (int)((long long)(blockIdx.x + 1) * n_items / gridDim.x) : min(n_items, (int)blockIdx.x + 1);
    const int n_up = UPPER ? a.n_up_fwd + a.n_up_bwd : 0;
    // every block has at most one item: z stays in shared memory - unless upper phases run between the sweeps, which
    // gather their operands into the whole shared memory (two panel buffers of up_rows rows)
    const bool keep_z = !UPPER || (n_items <= (int)gridDim.x && n_up == 0);
    constexpr int UP_GRAN = NG >= 8 ? 128 : 256;     // panel rows: whole ring turns for every K split (upper_phase)
    const int up_rows = (int)((((size_t)a.n_rows * PC + (size_t)a.stage_rows * (PC + 2)) / (2 * PC)) / UP_GRAN * UP_GRAN);
    unsigned bar_target = 0;
    int cur_s = -1;
    __shared__ unsigned long long s_zbar;      // mbarrier of the bulk loads of z
    __shared__ int s_item;
    if (UPPER && threadIdx.x == 0) mbar_init(&s_zbar, 1);
    unsigned zphase = 0;
    __syncthreads();
    // More items than blocks: items are handed out by a work counter, one draw per item and time step (subdomain-major
    // order: the blocks that run at the same time work on chunks of the same few subdomains, whose factor streams stay
    // hot in L2). A static split leaves every block waiting for the one with the most expensive items at every time
    // step (boundary subdomains are cheaper, 2048 items do not divide by 148 blocks); with the counter the spread is
    // at most one item. Which block runs an item does not change its arithmetic. Needs the grid barriers of the top
    // phase between the time steps (the counter only ever grows: step k draws from k * (n_items + gridDim.x) on).
    unsigned long long* item_ctr = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(a.bar) + BAR_COUNTERS);
    const bool dyn = UPPER && n_items > (int)gridDim.x && a.n_top > 0 && !a.dbg_b && !(a.dbg & 8);
    // dense top product (top_product_areg): RB row tiles per block so that one round of blocks covers it, the K range
    // of a stage split over KQ = 16 / RB warps (KQ must divide the number of k-steps), KM k-steps per warp and stage,
    // a ring of S stages in the staging rows. Without upper phases and with exactly one item per block the whole top
    // phase is CHUNK-LOCAL: the top rows of a problem chunk are assembled and multiplied by the P blocks of that
    // chunk, and the barriers only join those blocks.
    int top_rb = 0, ar_kq = 1, top_ar = 0, ar_s = 0;
    bool chunk_local = false;
    if (a.n_tt > 0) {
        const int stage_bytes = a.stage_rows * (PC + 2) * 8;
        const int GWr = NG < 4 ? NG : 4, GPr = (NG + 3) / 4;
        const int RT = (a.n_tt + 7) / 8, KS = a.n_tt_pad / 4;
        auto pick = [&](int rb, int& kq, int& km, int& ss) {
            kq = max(1, NWARPS / rb);
            while (kq > 1 && KS % kq != 0) --kq;
            km = 0;
            for (int m = 8; m >= 1 && km == 0; m >>= 1) {
                const int sb = GWr * kq * m * 256;
                if (KS % (kq * m) == 0 && sb / 16 <= NT && 3 * sb <= stage_bytes && NWARPS * GWr * 64 * 8 <= stage_bytes) {
                    km = m; ss = min(4, stage_bytes / sb);
                }
            }
        };
        if (n_up == 0 && n_items == (int)gridDim.x && a.n_chunks <= 60 && !(a.dbg & 4)) {
            const int rb = min(NWARPS, (RT * GPr + a.P - 1) / a.P);
            int kq, km, ss = 0;
            pick(rb, kq, km, ss);
            if (kq >= (NWARPS >= 16 ? 2 : 1) && km > 0) { chunk_local = true; top_rb = rb; ar_kq = kq; top_ar = km; ar_s = ss; }
        }
        if (!chunk_local) {
            const int per_chunk = max(1, (int)gridDim.x / (a.n_chunks * GPr));
            top_rb = min(NWARPS, (RT + per_chunk - 1) / per_chunk);
            if (!(a.dbg & 64)) pick(top_rb, ar_kq, top_ar, ar_s);
        }
    }
    auto dense_top = [&](int c0, int nc, int worker, int n_workers) {
        if (top_ar == 8) top_product_areg<NG, 8>(a, stage, top_rb, ar_kq, ar_s, c0, nc, worker, n_workers);
        else if (top_ar == 4) top_product_areg<NG, 4>(a, stage, top_rb, ar_kq, ar_s, c0, nc, worker, n_workers);
        else if (top_ar == 2) top_product_areg<NG, 2>(a, stage, top_rb, ar_kq, ar_s, c0, nc, worker, n_workers);
        else if (top_ar == 1) top_product_areg<NG, 1>(a, stage, top_rb, ar_kq, ar_s, c0, nc, worker, n_workers);
        else top_product_direct<NG>(a);
    };
    long long up_step = 0;               // time steps done: the work counters of the upper phases only ever grow
    auto up_phase_timed = [&](int ph) {
        if (!UPPER) return;
        const bool pr = prof_p && threadIdx.x == 0 && ph < PROF_UP;
        const long long t0 = pr ? clock64() : 0;
        upper_phase<NG>(a, smem, up_rows, ph, up_step);
        const long long t1 = pr ? clock64() : 0;
        grid_barrier(a.bar, bar_target);
        if (pr) {
            prof_p[(size_t)blockIdx.x * PROF_SLOTS + 56 + 2 * ph] += t1 - t0;
            prof_p[(size_t)blockIdx.x * PROF_SLOTS + 57 + 2 * ph] += clock64() - t1;
        }
    };
    // the solve of the separator rows for all chunks, between the forward and the backward local sweeps
    auto top_phase_grid = [&](long long n, long long& tq) {
        grid_barrier(a.bar, bar_target);
        if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 4] += tn - tq; tq = tn; }
        top_assemble<NG>(a, n, 0, a.n_chunks, blockIdx.x, gridDim.x);
        grid_barrier(a.bar, bar_target);
        if (UPPER) for (int ph = 0; ph < a.n_up_fwd; ++ph) up_phase_timed(ph);
        if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 5] += tn - tq; tq = tn; }
        if (a.n_tt > 0) {
            dense_top(0, a.n_chunks, blockIdx.x, gridDim.x);
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 6] += tn - tq; tq = tn; }
            grid_barrier(a.bar, bar_target);
        }
        if (UPPER) for (int ph = a.n_up_fwd; ph < n_up; ++ph) up_phase_timed(ph);
        ++up_step;
        if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 7] += tn - tq; tq = tn; }
    };
    if (UPPER && a.dbg_b) {
        // ---- debug: one solve J = S^-1 b through the plan (general kernel only: the host launches that one for it)
        for (int item = it_lo; item < it_hi; ++item) {
            const int s = multi ? item / a.n_chunks : item % a.P, c = multi ? item % a.n_chunks : item / a.P;
            __syncthreads();
            load_prog<NG>(a, s, ps, aux);
            const int nl = a.n_loc[s], nh = a.n_halo[s];
            const int* fidx = a.face_fidx + (size_t)s * a.n_rows;
            for (int e = threadIdx.x; e < (nl + nh) * PC; e += NT) {
                const int row = e / PC, q = e % PC, w = c * PC + q;
                double val = 0.0;
                if (row < nl && w < a.Wp) val = a.dbg_b[(size_t)fidx[row] * a.Wp + w];
                v[velem<NG>(row, q)] = val;
            }
            __syncthreads();
            run_levels<NG>(ps, ps.n_bwd, ps.n_levels, v, stage, prof_p);
            rows_to_global<NG>(v, nl, nh, a.ctop + ((size_t)c * a.n_slots + a.hptr[s]) * PC);
            rows_to_global<NG>(v, 0, nl, a.zloc + ((size_t)item * a.n_loc_max) * PC);
        }
        { long long tq = 0; top_phase_grid(0, tq); }
        for (int item = it_lo; item < it_hi; ++item) {
            const int s = multi ? item / a.n_chunks : item % a.P, c = multi ? item % a.n_chunks : item / a.P;
            __syncthreads();
            load_prog<NG>(a, s, ps, aux);
            const int nl = a.n_loc[s], nh = a.n_halo[s];
            const int* fidx = a.face_fidx + (size_t)s * a.n_rows;
            const int* ht = a.halo_top + a.hptr[s];
            for (int e = threadIdx.x; e < (nl + nh) * PC; e += NT) {
                const int row = e / PC, q = e % PC;
                double val;
                if (row < nl) val = __ldcg(a.zloc + ((size_t)item * a.n_loc_max + row) * PC + q);
                else val = __ldcg(a.jtop + ((size_t)c * a.n_up_pad + ht[row - nl]) * PC + q);
                v[velem<NG>(row, q)] = val;
            }
            __syncthreads();
            run_levels<NG>(ps, 0, ps.n_bwd, v, stage, prof_p);
            for (int e = threadIdx.x; e < nl * PC; e += NT) {
                const int row = e / PC, q = e % PC, w = c * PC + q;
                if (w < a.Wp) a.dbg_J[(size_t)fidx[row] * a.Wp + w] = v[velem<NG>(row, q)];
            }
        }
        for (long long e = (long long)blockIdx.x * NT + threadIdx.x; e < (long long)a.n_top * a.n_chunks * PC; e += (long long)gridDim.x * NT) {
            const int q = (int)(e % PC); const long long t = e / PC;
            const int k = (int)(t % a.n_top), c = (int)(t / a.n_top), w = c * PC + q;
            if (w < a.Wp) a.dbg_J[(size_t)a.top_face[k] * a.Wp + w] = __ldcg(a.jtop + ((size_t)c * a.n_up_pad + k) * PC + q);
        }
        return;
    }

    if (chunk_local && a.stagger > 0) {
        // Problem chunks are independent pipelines with the same period. Started together they stay in lockstep: every
        // SM streams the state through HBM during the same third of the time step (the junction pass then runs at
        // the HBM limit) and leaves it idle for the rest. Started apart, the chunks are in different phases at any
        // moment: the HBM load is even over the time step and the pass is no longer bandwidth-bound.
        const long long t0 = clock64(), wait = (long long)(blockIdx.x / a.P) * a.stagger;
        while (clock64() - t0 < wait) __nanosleep(256);
    }
    for (int k = 0; k <= a.n; ++k) {          // (a 32-bit counter: it stays live through every loop of the time step)
        const long long n = a.i0 + k;
        long long tq = prof_p ? clock64() : 0;
        const unsigned long long draw_base = (unsigned long long)k * (unsigned long long)(n_items + (int)gridDim.x);
        unsigned long long drawn = 0;
        if (dyn && threadIdx.x == 0) drawn = atomicAdd(item_ctr, 1ull) - draw_base;
        int item = it_lo - 1;
        auto next_item = [&]() -> bool {
            if (!dyn) return ++item < it_hi;
            if (threadIdx.x == 0) s_item = drawn < (unsigned long long)n_items ? (int)drawn : n_items;
            __syncthreads();                 // (every item body has block barriers: s_item was read by all before this)
            item = s_item;
            if (item >= n_items) return false;
            if (threadIdx.x == 0) drawn = atomicAdd(item_ctr, 1ull) - draw_base;    // the next draw, while this item runs
            return true;
        };
        while (next_item()) {
            const int s = multi ? item / a.n_chunks : item % a.P, c = multi ? item % a.n_chunks : item / a.P;
            if (s != cur_s) { __syncthreads(); load_prog<NG>(a, s, ps, aux); cur_s = s; }
            const int nl = a.n_loc[s], nh = a.n_halo[s];
            amp_fill<PC>(a, ac, c, n - 1);
            amp_fill<PC>(a, ac, c, n);
            if (k > 0) {
                // z of the local rows (one bulk copy, in flight while the halo rows are gathered) and J_top of the halo
                // rows, then the backward sweep
                if (UPPER && !keep_z && nl > 0 && threadIdx.x == 0) {
                    fence_proxy_async();     // the vector was last touched through the generic proxy (barrier at the item's end)
                    bulk_load(v, a.zloc + ((size_t)item * a.n_loc_max) * PC, (unsigned)(nl * PC * sizeof(double)), &s_zbar);
                }
                const int* ht = a.halo_top + a.hptr[s];
                for (int e = threadIdx.x; e < nh * (PC / 2); e += NT) {
                    const int r = e / (PC / 2), q = (e % (PC / 2)) * 2;
                    const double2 val = __ldcg(reinterpret_cast<const double2*>(a.jtop + ((size_t)c * a.n_up_pad + ht[r]) * PC + q));
                    *reinterpret_cast<double2*>(v + velem<NG>(nl + r, q)) = val;
                }
                if (UPPER && !keep_z && nl > 0) { mbar_wait(&s_zbar, zphase); zphase ^= 1u; }
                __syncthreads();
                run_levels<NG>(ps, 0, ps.n_bwd, v, stage, prof_p);
            } else {
                __syncthreads();
            }
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 0] += tn - tq; tq = tn; }
            junction_pass<NG, DEF>(a, ac, s, c, n, k > 0, k < a.n, v);
            if (k > 0 && a.obs_interval > 0 && n - 1 >= a.obs_first && (n - 1 - a.obs_first) % a.obs_interval == 0) {
                __syncthreads();                 // the phases of step n - 1 are in the state slab
                VortexView o;
                o.fj = a.face_ell_j + (size_t)s * a.n_rows * a.face_K;
                o.fidx = a.face_fidx + (size_t)s * a.n_rows; o.ht = a.halo_top + a.hptr[s]; o.top_face = a.top_face;
                o.th = a.rth + (size_t)c * a.Nj * PC; o.jrec = a.jrec;
                o.nsum = a.obs_nsum; o.th_last = a.obs_th_last; o.th_first = (n - 1 == a.obs_first) ? a.obs_th_first : nullptr;
                o.K = a.face_K; o.rows = nl + nh; o.nl = nl; o.jlo = a.junc_ptr[s]; o.jhi = a.junc_ptr[s + 1]; o.c = c; o.Wp = a.Wp;
                vortex_pass<NG>(o);
            }
            if (k == a.n) { __syncthreads(); continue; }
            __syncthreads();
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 1] += tn - tq; tq = tn; }
            face_pass<NG>(a, ac, s, c, n, nl + nh, v);
            __syncthreads();
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 2] += tn - tq; tq = tn; }
            run_levels<NG>(ps, ps.n_bwd, ps.n_levels, v, stage, prof_p);
            if (UPPER && !keep_z && nl > 0 && threadIdx.x == 0) {
                // z of the local rows leaves as one bulk copy while the block writes out the halo contributions
                fence_proxy_async();         // the sweep's shared-memory writes (all before its closing barrier) -> async proxy
                bulk_store(a.zloc + ((size_t)item * a.n_loc_max) * PC, v, (unsigned)(nl * PC * sizeof(double)));
            }
            rows_to_global<NG>(v, nl, nh, a.ctop + ((size_t)c * a.n_slots + a.hptr[s]) * PC);
            if (UPPER && !keep_z && nl > 0 && threadIdx.x == 0) {
                bulk_store_wait();           // complete (not only read): the next reader is another block, after a grid barrier
                fence_proxy_async();
            }
            __syncthreads();     // the vector is reused by the next item / the next step
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 3] += tn - tq; tq = tn; }
        }
        if (k == a.n) break;
        if (a.n_top > 0 && chunk_local) {
            // every block owns one (subdomain, chunk) item and the top rows of a chunk are assembled and multiplied
            // by the P blocks of that chunk: the three barriers of a time step only join those P blocks (one counter
            // per chunk, each on its own 128-byte line), and the chunks drift apart freely
            const int c = blockIdx.x / a.P, s = blockIdx.x % a.P;
            unsigned* ctr = a.bar + 32 * (1 + c);
            group_barrier(ctr, bar_target, a.P);
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 4] += tn - tq; tq = tn; }
            top_assemble<NG>(a, n, c, 1, s, a.P);
            group_barrier(ctr, bar_target, a.P);
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 5] += tn - tq; tq = tn; }
            dense_top(c, 1, s, a.P);
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 6] += tn - tq; tq = tn; }
            group_barrier(ctr, bar_target, a.P);
            if (prof_p && threadIdx.x == 0) { const long long tn = clock64(); prof_p[(size_t)blockIdx.x * PROF_SLOTS + 7] += tn - tq; tq = tn; }
        } else if (a.n_top > 0) {
            top_phase_grid(n, tq);
        }
    }
}

// flux base of the top rows, gathered once per launch (the assembly then needs no dependent index load for it)
__global__ void k_sub_gather_top(int n_top, const int* top_face, const double* fbase, double* topF) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_top) topF[k] = fbase[top_face[k]];
}

// flux base of the local rows of every subdomain (0 for halo rows), gathered once per launch
__global__ void k_sub_gather_rows(long long n, const int* fidx, const double* fbase, double* rowF) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { const int g = fidx[k]; rowF[k] = g >= 0 ? fbase[g] : 0.0; }
}

typedef void (*KernelPtr)(const SubArgs);

}  // namespace

// The step kernel is compiled once per chunk width in its own translation unit (build.sh compiles this file with
// -DJJ_SUB_NG=1|2|4|8 in parallel; without the macro only the host side below is compiled).
namespace jj {
KernelPtr subdomain_kernel_ng1(bool def, bool upper);
KernelPtr subdomain_kernel_ng2(bool def, bool upper);
KernelPtr subdomain_kernel_ng4(bool def, bool upper);
KernelPtr subdomain_kernel_ng8(bool def, bool upper);
// the annealing schedule's variant (full blocks): phase zones instead of phases in the stored planes
KernelPtr subdomain_kernel_m_ng1(bool def, bool upper);
KernelPtr subdomain_kernel_m_ng2(bool def, bool upper);
KernelPtr subdomain_kernel_m_ng4(bool def, bool upper);
KernelPtr subdomain_kernel_m_ng8(bool def, bool upper);
// half blocks (256 threads, two per SM); no upper program
KernelPtr subdomain_kernel_h_ng1(bool def, bool upper);
KernelPtr subdomain_kernel_h_ng2(bool def, bool upper);
KernelPtr subdomain_kernel_h_ng4(bool def, bool upper);
}

#ifdef JJ_SUB_NG
#define JJ_SUB_CAT2(a, b) a##b
#define JJ_SUB_CAT(a, b) JJ_SUB_CAT2(a, b)
namespace jj {
#if JJ_SUB_NT == 256
KernelPtr JJ_SUB_CAT(subdomain_kernel_h_ng, JJ_SUB_NG)(bool def, bool upper) {
    if (upper) return nullptr;           // the upper phases are written for 16 warps
    return def ? k_subdomain<JJ_SUB_NG, true, false> : k_subdomain<JJ_SUB_NG, false, false>;
}
#elif JJ_SUB_MOB
KernelPtr JJ_SUB_CAT(subdomain_kernel_m_ng, JJ_SUB_NG)(bool def, bool upper) {
    if (upper) return def ? k_subdomain<JJ_SUB_NG, true, true> : k_subdomain<JJ_SUB_NG, false, true>;
    return def ? k_subdomain<JJ_SUB_NG, true, false> : k_subdomain<JJ_SUB_NG, false, false>;
}
#else
KernelPtr JJ_SUB_CAT(subdomain_kernel_ng, JJ_SUB_NG)(bool def, bool upper) {
    if (upper) return def ? k_subdomain<JJ_SUB_NG, true, true> : k_subdomain<JJ_SUB_NG, false, true>;
    return def ? k_subdomain<JJ_SUB_NG, true, false> : k_subdomain<JJ_SUB_NG, false, false>;
}
#endif
}
#else

namespace {

KernelPtr pick_kernel(int threads, int NG, bool def, bool upper, bool zones) {
    if (zones) {
        if (threads != NT) return nullptr;
        switch (NG) {
            case 1: return subdomain_kernel_m_ng1(def, upper);
            case 2: return subdomain_kernel_m_ng2(def, upper);
            case 4: return subdomain_kernel_m_ng4(def, upper);
            default: return subdomain_kernel_m_ng8(def, upper);
        }
    }
    if (threads == 256) {
        switch (NG) {
            case 1: return subdomain_kernel_h_ng1(def, upper);
            case 2: return subdomain_kernel_h_ng2(def, upper);
            case 4: return subdomain_kernel_h_ng4(def, upper);
            default: return nullptr;
        }
    }
    switch (NG) {
        case 1: return subdomain_kernel_ng1(def, upper);
        case 2: return subdomain_kernel_ng2(def, upper);
        case 4: return subdomain_kernel_ng4(def, upper);
        default: return subdomain_kernel_ng8(def, upper);
    }
}

// per-junction constants in device junction order (coalesced records instead of gathers by original index)
__global__ void k_sub_gather_params(int n, const int* orig, const double* Ic, const double* c0, const double* c1,
                                    const double* c2, const double* isb, const double* tb, const double* vsb,
                                    const int2* rows, const char2* sign, double* P0, double* P1, double* jrec) {
    int jp = blockIdx.x * blockDim.x + threadIdx.x;
    if (jp >= n) return;
    int jo = orig[jp];
    P0[4 * jp + 0] = Ic[jo]; P0[4 * jp + 1] = 1.0 / c0[jo]; P0[4 * jp + 2] = c1[jo]; P0[4 * jp + 3] = c2[jo];
    P1[4 * jp + 0] = isb ? isb[jo] : 0.0; P1[4 * jp + 1] = tb ? tb[jo] : 0.0; P1[4 * jp + 2] = vsb ? vsb[jo] : 0.0;
    P1[4 * jp + 3] = 0.0;
    // the junction record holds the coefficients of x'/c0 (the x stream is kept divided by c0: it is what the face
    // pass sums and what theta_n = y/c0 - x'/c0 subtracts), so 1/c0 itself is only needed for y
    double* r = jrec + 8 * (size_t)jp;
    const double ic0 = 1.0 / c0[jo];
    r[0] = Ic[jo] * ic0; r[1] = ic0; r[2] = c1[jo] * ic0; r[3] = c2[jo] * ic0;
    r[4] = (isb ? isb[jo] : 0.0) * ic0; r[5] = (tb ? tb[jo] : 0.0) * ic0;
    const int2 rw = rows[jp];
    const char2 sg = sign[jp];
    const int s0 = rw.x >= 0 ? (int)sg.x : 0, s1 = rw.y >= 0 ? (int)sg.y : 0;
    int4 ri = make_int4(rw.x >= 0 ? rw.x : 0, rw.y >= 0 ? rw.y : 0, jo, (s0 + 1) | ((s1 + 1) << 2));
    *reinterpret_cast<int4*>(r + 6) = ri;
}

#define SCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            h->err = std::string("subdomain: ") + #call + ": " + cudaGetErrorString(e_);            \
            return JJ_ECUDA;                                                                        \
        }                                                                                           \
    } while (0)

template <typename T>
int up(JJHandle* h, SubState* st, T** dst, const T* src, size_t n, size_t pad_bytes = 0) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T) + pad_bytes;
    int rc = dev_alloc(h, &p, bytes);
    if (rc) return rc;
    st->allocs.push_back(p); st->alloc_bytes.push_back(bytes);
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, h->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { h->err = std::string("subdomain upload: ") + cudaGetErrorString(e); return JJ_ECUDA; }
    *dst = (T*)p;
    return JJ_OK;
}

size_t smem_for(const SubState* st) {
    const int PC = st->PC;
    size_t amp = (size_t)4 * 2 * PC * sizeof(double);
    size_t aux = ((size_t)3 * st->max_np + st->max_levels + 2 + 2 * (size_t)st->max_tiles + 8) * sizeof(int);
    return ((size_t)st->n_rows * PC + (size_t)st->stage_rows * (PC + 2)) * sizeof(double) + amp + aux;
}

}  // namespace

namespace jj {

void subdomain_free_problem(JJHandle* h) {
    SubState* st = (SubState*)h->subdomain_plan;
    if (!st) return;
    dev_free(h, st->rth, st->state_bytes); dev_free(h, st->rx, st->state_bytes);
    dev_free(h, st->zloc, st->z_bytes); dev_free(h, st->ctop, st->c_bytes);
    dev_free(h, st->U, st->u_bytes);
    dev_free(h, st->bar, BAR_BYTES);
    dev_free(h, st->plane_d, st->plane_cap);
    st->rth = st->rx = st->zloc = st->ctop = st->rtop = st->jtop = st->U = nullptr; st->bar = nullptr;
    st->plane_d = nullptr; st->plane_cap = 0; st->plane_h.clear(); st->state_bytes = st->z_bytes = st->c_bytes = st->u_bytes = 0;
    st->prepared = false;
}

void subdomain_drop_plan(JJHandle* h) {
    SubState* st = (SubState*)h->subdomain_plan;
    if (!st) return;
    subdomain_free_problem(h);
    for (size_t i = 0; i < st->allocs.size(); ++i) dev_free(h, st->allocs[i], st->alloc_bytes[i]);
    delete st;
    h->subdomain_plan = nullptr;
}

int subdomain_set_plan(JJHandle* h, const JJSubdomainPlan* pl) {
    subdomain_drop_plan(h);
    if (!pl) return JJ_OK;
    if (!(pl->NG == 1 || pl->NG == 2 || pl->NG == 4 || pl->NG == 8) || pl->P < 1 || pl->n_tt_pad % 32 != 0 ||
        pl->n_up_pad % 32 != 0 || pl->n_up_pad <= pl->n_top || pl->tt0 + pl->n_tt != pl->n_top ||
        pl->tt0 + pl->n_tt_pad > pl->n_up_pad || pl->n_up_pad >= (1 << 28)) {
        h->err = "subdomain plan: NG must be 1/2/4/8, P >= 1, n_tt_pad and n_up_pad multiples of 32 with "
                 "tt0 + n_tt == n_top < n_up_pad and tt0 + n_tt_pad <= n_up_pad";
        return JJ_EINVAL;
    }
    if (pl->n_up_fwd + pl->n_up_bwd > 0 && !(pl->up_RB == 16 && pl->up_KB == 16)) {
        h->err = "subdomain plan: upper program must be packed for up_RB == 16 row tiles per task, up_KB == 16";
        return JJ_EINVAL;
    }
    SubState* st = new SubState();
    h->subdomain_plan = st;
    st->P = pl->P; st->NG = pl->NG; st->PC = 8 * pl->NG;
    st->n_rows = pl->n_rows; st->n_loc_max = pl->n_loc_max; st->stage_rows = pl->stage_rows;
    st->n_top = pl->n_top; st->n_up_pad = pl->n_up_pad; st->n_slots = pl->n_slots; st->face_K = pl->face_K;
    st->tt0 = pl->tt0; st->n_tt = pl->n_tt; st->n_tt_pad = pl->n_tt_pad;
    st->up_RB = pl->up_RB; st->up_KB = pl->up_KB; st->n_up_fwd = pl->n_up_fwd; st->n_up_bwd = pl->n_up_bwd;
    {   // switches read once per plan (never on the launch path)
        const char* e;
        st->dbg = (e = getenv("JJ_SUB_DEBUG")) ? atoi(e) : 0;
        st->prof = getenv("JJ_SUB_PROF") != nullptr;
        st->no_tslot4 = getenv("JJ_SUB_NO_TSLOT4") != nullptr;
        st->no_l2_window = getenv("JJ_SUB_NO_L2_WINDOW") != nullptr;
        st->grid_env = (e = getenv("JJ_SUB_GRID")) ? atoi(e) : 0;
        st->stagger = (e = getenv("JJ_SUB_STAGGER")) ? atoi(e) : -1;
    }
    if ((size_t)st->n_rows * st->PC > 65536) { h->err = "subdomain plan: shared-memory vector exceeds the 16-bit element codes"; return JJ_EINVAL; }
    const int Nj = h->cir.Nj, P = pl->P;
    int rc;
    std::vector<SubProgDev> progs(P);
    std::map<const void*, int> seen;          // host stream pointer -> first subdomain with that program
    for (int s = 0; s < P; ++s) {
        const JJSubProgram& ps = pl->prog[s];
        if (s == 0) st->threads = 32 * ps.n_warps;
        if ((ps.n_warps != 16 && ps.n_warps != 8) || 32 * ps.n_warps != st->threads ||
            (ps.n_warps == 8 && (pl->n_up_fwd + pl->n_up_bwd > 0 || pl->NG == 8))) {
            h->err = "subdomain plan: programs must all be packed for 16 warps, or for 8 warps (half blocks: no upper program, NG <= 4)";
            return JJ_EINVAL;
        }
        // congruent subdomains (translated copies on a regular lattice) come with the SAME host arrays: one device
        // copy serves them all, and the shared factor stream stays in L2
        if (ps.n_steps > 0) {
            auto it = seen.find((const void*)ps.stream);
            if (it != seen.end() && pl->prog[it->second].wt_ptr == ps.wt_ptr && pl->prog[it->second].thdr == ps.thdr &&
                pl->prog[it->second].n_steps == ps.n_steps) {
                progs[s] = progs[it->second];
                progs[s].n_bwd = ps.n_bwd;
                continue;
            }
            seen[(const void*)ps.stream] = s;
        }
        const size_t np = (size_t)ps.n_levels * ps.n_warps;
        int *wt, *ws, *th, *ls; unsigned char* sb;
        if ((rc = up(h, st, &wt, ps.wt_ptr, 2 * np))) return rc;
        if ((rc = up(h, st, &ws, ps.ws_ptr, np))) return rc;
        if ((rc = up(h, st, &th, ps.thdr, (size_t)ps.n_tiles * 2))) return rc;
        if ((rc = up(h, st, &ls, ps.lstaged, (size_t)ps.n_levels))) return rc;
        // stream + 2*RING zero steps of padding (the kernel prefetches RING steps ahead unconditionally)
        if ((rc = up(h, st, &sb, ps.stream, (size_t)ps.n_steps * STEP_BYTES, (size_t)2 * RING * STEP_BYTES))) return rc;
        progs[s].wt_ptr = wt; progs[s].ws_ptr = ws; progs[s].thdr = (const int2*)th; progs[s].lstaged = ls; progs[s].stream = sb;
        progs[s].n_levels = ps.n_levels; progs[s].n_bwd = ps.n_bwd; progs[s].n_tiles = ps.n_tiles; progs[s].pad_ = 0;
        st->max_np = std::max(st->max_np, (int)np);
        st->max_levels = std::max(st->max_levels, ps.n_levels);
        st->max_tiles = std::max(st->max_tiles, ps.n_tiles);
    }
    if ((rc = up(h, st, &st->prog, progs.data(), (size_t)P))) return rc;
    if ((rc = up(h, st, &st->n_loc, pl->n_loc, (size_t)P))) return rc;
    if ((rc = up(h, st, &st->n_halo, pl->n_halo, (size_t)P))) return rc;
    if ((rc = up(h, st, &st->hptr, pl->hptr, (size_t)P + 1))) return rc;
    if ((rc = up(h, st, &st->halo_top, pl->halo_top, (size_t)pl->n_slots))) return rc;
    if ((rc = up(h, st, &st->tptr, pl->tptr, (size_t)pl->n_top + 1))) return rc;
    if ((rc = up(h, st, &st->tslot, pl->tslot, (size_t)pl->n_slots))) return rc;
    if ((rc = up(h, st, &st->top_face, pl->top_face, (size_t)pl->n_top))) return rc;
    {
        // fixed-width slot lists of the top rows (a row of a planar circuit is shared by at most a few subdomains)
        std::vector<int4> t4((size_t)std::max(pl->n_top, 1), make_int4(-1, -1, -1, -1));
        bool fits = true;
        for (int k = 0; k < pl->n_top && fits; ++k) {
            const int cnt = pl->tptr[k + 1] - pl->tptr[k];
            if (cnt > 4) { fits = false; break; }
            int v4[4] = {-1, -1, -1, -1};
            for (int e = 0; e < cnt; ++e) v4[e] = pl->tslot[pl->tptr[k] + e];
            t4[k] = make_int4(v4[0], v4[1], v4[2], v4[3]);
        }
        st->tslot4 = nullptr;
        if (fits && pl->n_top > 0) { if ((rc = up(h, st, &st->tslot4, t4.data(), (size_t)pl->n_top))) return rc; }
        void* p = nullptr;
        const size_t bytes = (size_t)std::max(pl->n_top, 1) * sizeof(double);
        if ((rc = dev_alloc(h, &p, bytes))) return rc;
        st->allocs.push_back(p); st->alloc_bytes.push_back(bytes);
        st->topF = (double*)p;
        void* p2 = nullptr;
        const size_t bytes2 = (size_t)std::max(pl->P * pl->n_rows, 1) * sizeof(double);
        if ((rc = dev_alloc(h, &p2, bytes2))) return rc;
        st->allocs.push_back(p2); st->alloc_bytes.push_back(bytes2);
        st->rowF = (double*)p2;
    }
    if ((rc = up(h, st, &st->SinvP, pl->Sinv_packed, (size_t)pl->n_tt_pad * pl->n_tt_pad))) return rc;
    {
        const int n_ph = pl->n_up_fwd + pl->n_up_bwd;
        if (n_ph > 0) {
            if (pl->up_phase_ptr[n_ph] != pl->n_up_tasks) { h->err = "subdomain plan: up_phase_ptr does not cover the tasks"; return JJ_EINVAL; }
            for (int t = 0; t < pl->n_up_tasks; ++t) {
                const int32_t* hd = pl->up_task + 4 * (size_t)t;
                const int64_t tiles = (hd[1] + 7) / 8;
                const bool block_task = t < pl->up_phase_split[std::upper_bound(pl->up_phase_ptr, pl->up_phase_ptr + n_ph + 1, t) - pl->up_phase_ptr - 1];
                if (hd[1] < 1 || hd[1] > 8 * pl->up_RB || hd[2] < pl->up_KB || hd[2] % (block_task ? 64 : pl->up_KB) != 0 || hd[3] < 0 ||
                    (!block_task && hd[1] > 32) ||
                    (int64_t)hd[3] + 4 * (int64_t)hd[2] > pl->n_up_cols ||
                    pl->up_task_aoff[t] < 0 || pl->up_task_aoff[t] + tiles * hd[2] * 32 > pl->n_up_vals) {
                    h->err = "subdomain plan: malformed task in the upper program"; return JJ_EINVAL;
                }
            }
        }
        if ((rc = up(h, st, &st->up_phase_ptr, pl->up_phase_ptr, (size_t)n_ph + 1))) return rc;
        for (int ph = 0; ph < n_ph; ++ph)
            if (pl->up_phase_split[ph] < pl->up_phase_ptr[ph] || pl->up_phase_split[ph] > pl->up_phase_ptr[ph + 1]) {
                h->err = "subdomain plan: up_phase_split outside its phase"; return JJ_EINVAL;
            }
        if ((rc = up(h, st, &st->up_phase_split, pl->up_phase_split, (size_t)n_ph + 1))) return rc;
        if ((rc = up(h, st, &st->up_ctr, (const unsigned*)nullptr, 0, ((size_t)n_ph + 1) * sizeof(unsigned)))) return rc;
        if ((rc = up(h, st, (int**)&st->up_task, pl->up_task, (size_t)pl->n_up_tasks * 4))) return rc;
        if ((rc = up(h, st, (long long**)&st->up_task_aoff, (const long long*)pl->up_task_aoff, (size_t)pl->n_up_tasks))) return rc;
        if ((rc = up(h, st, &st->up_cols, pl->up_cols, (size_t)pl->n_up_cols))) return rc;
        if ((rc = up(h, st, &st->up_A, pl->up_A, (size_t)pl->n_up_vals, (size_t)8 * 16 * 32 * sizeof(double)))) return rc;
    }
    if ((rc = up(h, st, &st->junc_ptr, pl->junc_ptr, (size_t)P + 1))) return rc;
    if ((rc = up(h, st, &st->junc_orig, pl->junc_orig, (size_t)Nj))) return rc;
    if ((rc = up(h, st, (int**)&st->junc_row, pl->junc_row, (size_t)Nj * 2))) return rc;
    if ((rc = up(h, st, (signed char**)&st->junc_sign, (const signed char*)pl->junc_sign, (size_t)Nj * 2))) return rc;
    const size_t nell = (size_t)P * pl->n_rows * pl->face_K;
    {
        // device form of the row lists: 2 * junction + (coefficient < 0), -1 = absent (the magnitude 1/c0 of the plan's
        // coefficients is applied by the junction pass, which stores x'/c0)
        std::vector<int> packed(std::max<size_t>(nell, 1), -1);
        for (size_t e = 0; e < nell; ++e)
            if (pl->face_ell_j[e] >= 0) packed[e] = 2 * pl->face_ell_j[e] + (pl->face_ell_c[e] < 0.0 ? 1 : 0);
        if ((rc = up(h, st, &st->face_ell_j, packed.data(), nell))) return rc;
        SCK(cudaStreamSynchronize(h->stream));       // `packed` goes out of scope
    }
    if ((rc = up(h, st, &st->face_fidx, pl->face_fidx, (size_t)P * pl->n_rows))) return rc;
    for (double** pp : {&st->P0, &st->P1, &st->jrec}) {
        void* p = nullptr;
        const size_t bytes = (size_t)Nj * (pp == &st->jrec ? 8 : 4) * sizeof(double);
        if ((rc = dev_alloc(h, &p, bytes))) return rc;
        st->allocs.push_back(p); st->alloc_bytes.push_back(bytes);
        *pp = (double*)p;
    }
    SCK(cudaStreamSynchronize(h->stream));
    st->smem_bytes = smem_for(st);
    // 228 KB per SM, 1 KB reserved per resident block (+ the kernel's own static shared memory)
    const size_t smem_cap = st->threads == 256 ? (size_t)112 * 1024 : (size_t)226 * 1024;
    if (st->smem_bytes > smem_cap) { h->err = "subdomain plan: shared memory per block exceeds 226 KB (112 KB for half blocks)"; return JJ_EINVAL; }
    return JJ_OK;
}

int subdomain_supported(JJHandle* h, std::string& why) {
    SubState* st = (SubState*)h->subdomain_plan;
    if (!st) { why = "no subdomain plan was provided"; return 0; }
    for (int i = 0; i < 4; ++i)
        if (h->src[i].dev.kind == KIND_DENSE) { why = "a per-step input is dense (not base x amplitude)"; return 0; }
    if (h->cir.Nf == 0) { why = "circuit has no faces"; return 0; }
    return 1;
}

static void fill_args(JJHandle* h, SubState* st, SubArgs& a) {
    memset(&a, 0, sizeof(a));
    a.P = st->P; a.n_rows = st->n_rows; a.n_loc_max = st->n_loc_max; a.stage_rows = st->stage_rows;
    a.n_top = st->n_top; a.n_up_pad = st->n_up_pad; a.n_slots = st->n_slots;
    a.tt0 = st->tt0; a.n_tt = st->n_tt; a.n_tt_pad = st->n_tt_pad;
    a.up_RB = st->up_RB; a.up_KB = st->up_KB; a.n_up_fwd = st->n_up_fwd; a.n_up_bwd = st->n_up_bwd;
    a.up_phase_ptr = st->up_phase_ptr; a.up_phase_split = st->up_phase_split; a.up_task = st->up_task; a.up_task_aoff = st->up_task_aoff;
    a.up_ctr = st->up_ctr;
    a.up_cols = st->up_cols; a.up_A = st->up_A; a.U = st->U;
    a.max_np = st->max_np; a.max_levels = st->max_levels; a.max_tiles = st->max_tiles;
    a.prog = st->prog; a.n_loc = st->n_loc; a.n_halo = st->n_halo; a.hptr = st->hptr; a.halo_top = st->halo_top;
    a.tptr = st->tptr; a.tslot = st->tslot; a.top_face = st->top_face; a.SinvP = st->SinvP;
    a.tslot4 = st->no_tslot4 ? nullptr : st->tslot4;
    a.topF = (h->src[JJ_SRC_F].dev.kind == KIND_RANK1 && st->n_top > 0) ? st->topF : nullptr;
    a.rowF = h->src[JJ_SRC_F].dev.kind == KIND_RANK1 ? st->rowF : nullptr;
    a.junc_ptr = st->junc_ptr; a.junc_orig = st->junc_orig; a.junc_row = st->junc_row; a.junc_sign = st->junc_sign;
    a.face_K = st->face_K; a.face_ell_j = st->face_ell_j; a.face_ell_c = st->face_ell_c; a.face_fidx = st->face_fidx;
    a.Nj = h->cir.Nj; a.Nf = h->cir.Nf; a.P0 = st->P0; a.P1 = st->P1; a.jrec = st->jrec; a.cpr = h->cir.cpr;
    a.Wp = h->Wp; a.n_chunks = st->n_chunks; a.dt = h->dt; a.seed = h->seed; a.group_offset = h->problem_offset / 4;
    a.Is = h->src[JJ_SRC_IS].dev; a.Vs = h->src[JJ_SRC_VS].dev; a.T = h->src[JJ_SRC_T].dev; a.F = h->src[JJ_SRC_F].dev;
    a.noise = h->noise_buf; a.noise_i0 = h->noise_i0; a.noise_K = h->noise_K;
    a.rth = st->rth; a.rx = st->rx; a.th1 = h->th1; a.th2 = h->th2; a.th2_in = h->start_at_rest ? h->th1 : h->th2;
    a.zloc = st->zloc; a.ctop = st->ctop; a.rtop = st->rtop; a.jtop = st->jtop; a.bar = st->bar;
    a.snap_th = h->th_out; a.snap_I = h->I_out; a.flag = h->flag_d;
    a.obs_first = h->obs_first; a.obs_interval = h->obs_interval;
    a.obs_nsum = h->obs_nsum; a.obs_th_first = h->obs_th_first; a.obs_th_last = h->obs_th_last;
    a.zone8 = h->zone8;
    a.dbg = st->dbg;
    a.stagger = st->stagger > 0 ? st->stagger : 0;
}

static int launch(JJHandle* h, SubState* st, SubArgs& a) {
    // the lean kernel when every block has at most one item and there are no upper phases, else the general one
    // (both are limited to one block per SM by their shared memory - two for half blocks -, so the capacity is known)
    int sms = 0;
    SCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    int cap = sms * (st->threads == 256 ? 2 : 1);
    if (st->grid_env > 0) cap = std::min(cap, st->grid_env);
    const bool general = st->n_up_fwd + st->n_up_bwd > 0 || st->P * st->n_chunks > cap || a.dbg_b != nullptr;
    KernelPtr k = pick_kernel(st->threads, st->NG, h->cir.default_cpr, general, a.zone8 != nullptr);
    if (!k) { h->err = "subdomain: no kernel for this block size / chunk width / item count"; return JJ_EINVAL; }
    // per-junction records and per-row flux bases in device order: gathered again only after an input was redeclared
    // (an annealing schedule launches the kernel hundreds of times on the same inputs)
    const bool regather = st->gather_gen != h->src_gen;
    st->gather_gen = h->src_gen;
    if (regather) {
        const Source &is = h->src[JJ_SRC_IS].dev, &t = h->src[JJ_SRC_T].dev, &vs = h->src[JJ_SRC_VS].dev;
        k_sub_gather_params<<<(h->cir.Nj + 255) / 256, 256, 0, h->stream>>>(
            h->cir.Nj, st->junc_orig, h->cir.Ic, h->cir.c0, h->cir.c1, h->cir.c2,
            is.kind == KIND_RANK1 ? is.base : nullptr, t.kind == KIND_RANK1 ? t.base : nullptr,
            vs.kind == KIND_RANK1 ? vs.base : nullptr, st->junc_row, st->junc_sign, st->P0, st->P1, st->jrec);
        h->launches++;
    }
    if (regather && a.topF) {
        k_sub_gather_top<<<(st->n_top + 255) / 256, 256, 0, h->stream>>>(st->n_top, st->top_face, h->src[JJ_SRC_F].dev.base, st->topF);
        h->launches++;
    }
    if (regather && a.rowF) {
        const long long nr = (long long)st->P * st->n_rows;
        k_sub_gather_rows<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, st->face_fidx, h->src[JJ_SRC_F].dev.base, st->rowF);
        h->launches++;
    }
    SCK(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem_bytes));
    if (st->grid == 0) {
        int per_sm = 0;
        SCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, st->threads, st->smem_bytes));
        if (per_sm <= 0) { h->err = "subdomain: kernel does not fit on the device"; return JJ_EINVAL; }
        st->grid = sms * std::min(per_sm, st->threads == 256 ? 2 : 1);
    }
    // the staged top product sizes its row blocks to the grid: give it up to 18 blocks per chunk and group pass
    const int top_tasks = st->n_chunks * ((st->NG + 3) / 4) * std::min(18, (st->n_tt + 7) / 8);
    int want = std::max(st->P * st->n_chunks, top_tasks);
    if (st->n_up_fwd + st->n_up_bwd > 0) want = st->grid;      // the upper phases spread over every SM
    int grid = std::min(st->grid, std::max(1, want));
    if (st->grid_env > 0) grid = std::min(st->grid, st->grid_env);
    SCK(cudaMemsetAsync(st->bar, 0, BAR_BYTES, h->stream));
    if (st->up_ctr) SCK(cudaMemsetAsync(st->up_ctr, 0, ((size_t)st->n_up_fwd + st->n_up_bwd + 1) * sizeof(unsigned), h->stream));
    if (st->n_tt > 0 && !st->no_l2_window) {
        // keep the packed Schur inverse resident in L2: it is re-read by every block once per time step while
        // the state (hundreds of MB per step) streams through the same cache. The limit is a per-device setting.
        if (!st->l2_limit_set) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)32 << 20); st->l2_limit_set = true; }
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = (void*)st->SinvP;
        attr.accessPolicyWindow.num_bytes = (size_t)st->n_tt_pad * st->n_tt_pad * sizeof(double);
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    void* params[] = {(void*)&a};
    SCK(cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(st->threads), params, st->smem_bytes, h->stream));
    h->launches++;
    return JJ_OK;
}

int subdomain_prepare(JJHandle* h) {
    SubState* st = (SubState*)h->subdomain_plan;
    if (!st) { h->err = "subdomain engine: no plan"; return JJ_ESTATE; }
    subdomain_free_problem(h);
    const int PC = st->PC;
    st->n_chunks = (h->Wp + PC - 1) / PC;
    st->state_bytes = (size_t)st->n_chunks * h->cir.Nj * PC * sizeof(double);
    st->z_bytes = (size_t)st->n_chunks * st->P * std::max(st->n_loc_max, 1) * PC * sizeof(double);
    st->c_bytes = (size_t)st->n_chunks * std::max(st->n_slots, 1) * PC * sizeof(double);
    st->u_bytes = (size_t)4 * st->n_chunks * std::max(st->n_up_pad, 32) * PC * sizeof(double);     // planes r, z, J, scratch
    int rc;
    if ((rc = dev_alloc(h, (void**)&st->rth, st->state_bytes))) return rc;
    if ((rc = dev_alloc(h, (void**)&st->rx, st->state_bytes))) return rc;
    if ((rc = dev_alloc(h, (void**)&st->zloc, st->z_bytes))) return rc;
    if ((rc = dev_alloc(h, (void**)&st->ctop, st->c_bytes))) return rc;
    if ((rc = dev_alloc(h, (void**)&st->U, st->u_bytes))) return rc;
    st->rtop = st->U; st->jtop = st->U + st->u_bytes / sizeof(double) / 4 * 2;
    if ((rc = dev_alloc(h, (void**)&st->bar, BAR_BYTES))) return rc;
    SCK(cudaMemsetAsync(st->rth, 0, st->state_bytes, h->stream));
    SCK(cudaMemsetAsync(st->rx, 0, st->state_bytes, h->stream));
    SCK(cudaMemsetAsync(st->zloc, 0, st->z_bytes, h->stream));
    SCK(cudaMemsetAsync(st->ctop, 0, st->c_bytes, h->stream));
    SCK(cudaMemsetAsync(st->U, 0, st->u_bytes, h->stream));
    st->prepared = true;
    st->gather_gen = 0;
    return JJ_OK;
}

bool subdomain_prepared(JJHandle* h) {
    SubState* st = (SubState*)h->subdomain_plan;
    return st && st->prepared;
}

int subdomain_run(JJHandle* h, long long i0, int n, const long long* th_plane, const long long* I_plane) {
    SubState* st = (SubState*)h->subdomain_plan;
    if (!st || !st->prepared) { h->err = "subdomain engine not prepared"; return JJ_ESTATE; }
    size_t need = (size_t)2 * n * sizeof(long long);
    if (need > st->plane_cap) {
        SCK(cudaStreamSynchronize(h->stream));
        dev_free(h, st->plane_d, st->plane_cap);
        st->plane_d = nullptr; st->plane_cap = 0; st->plane_h.clear();
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
    if (pl != st->plane_h) {
        // (an annealing schedule asks for the same planes every interval: uploaded once. The host copy lives in the
        // plan state, so nothing has to wait for the transfer.)
        st->plane_h.swap(pl);
        SCK(cudaMemcpyAsync(st->plane_d, st->plane_h.data(), need, cudaMemcpyHostToDevice, h->stream));
    }
    SubArgs a;
    fill_args(h, st, a);
    a.i0 = i0; a.n = n; a.th_plane = st->plane_d; a.I_plane = st->plane_d + n;
    if (!st->prof) return launch(h, st, a);
    // debugging aid: per-phase cycle counts of every block, printed as averages per time step
    const size_t nb = 1024;
    long long* prof = nullptr;
    SCK(cudaMalloc((void**)&prof, nb * PROF_SLOTS * sizeof(long long)));
    SCK(cudaMemset(prof, 0, nb * PROF_SLOTS * sizeof(long long)));
    a.prof = prof;
    int rc = launch(h, st, a);
    if (rc) { cudaFree(prof); return rc; }
    SCK(cudaStreamSynchronize(h->stream));
    std::vector<long long> hp(nb * PROF_SLOTS);
    SCK(cudaMemcpy(hp.data(), prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(prof);
    static const char* names[8] = {"bwd sweep", "junction pass", "face pass", "fwd sweep + store", "barrier 1 (wait)", "top assemble + barrier 2",
                                   "top product", "barrier 3 (wait)"};
    fprintf(stderr, "JJ_SUB_PROF: cycles per time step (n=%d steps, P=%d, PC=%d, chunks=%d)\n", n, st->P, st->PC, st->n_chunks);
    double tot = 0;
    for (int sl = 0; sl < PROF_SLOTS; ++sl) {
        double sum = 0, mx = 0; int cnt = 0;
        for (size_t b = 0; b < nb; ++b) { double v = (double)hp[b * PROF_SLOTS + sl] / n; if (v > 0) { sum += v; ++cnt; } mx = std::max(mx, v); }
        double avg = cnt ? sum / cnt : 0;
        if (sl < 8) { tot += avg; fprintf(stderr, "  %-26s avg %9.0f max %9.0f (blocks %d)\n", names[sl], avg, mx, cnt); }
        else if (sl < 56) { if (cnt) fprintf(stderr, "    sweep level %2d           avg %9.0f max %9.0f\n", sl - 8, avg, mx); }
        else if (sl < 56 + 2 * PROF_UP) { if (cnt) fprintf(stderr, "    upper phase %2d %s        avg %9.0f max %9.0f (blocks %d)\n", (sl - 56) / 2, (sl & 1) ? "wait" : "work", avg, mx, cnt); }
        else if (cnt) fprintf(stderr, "    stamp %d (phase 0, block 0, warp 0: header / codes / products / item end / before final sync)  %9.0f\n", sl - 56 - 2 * PROF_UP, avg);
    }
    fprintf(stderr, "  total avg cycles per time step %.0f\n", tot);
    // local work (sweeps + junction + face pass) per subdomain, averaged over its chunks: what the dissection balances
    fprintf(stderr, "  local cycles per subdomain:");
    for (int s = 0; s < std::min(st->P, 64); ++s) {
        double sum = 0; int cnt = 0;
        for (size_t b = s; b < nb && b < (size_t)st->P * st->n_chunks; b += st->P) {
            double v = 0; for (int sl = 0; sl < 4; ++sl) v += (double)hp[b * PROF_SLOTS + sl] / n;
            if (v > 0) { sum += v; ++cnt; }
        }
        fprintf(stderr, " %.0f", cnt ? sum / cnt : 0.0);
    }
    fprintf(stderr, "\n");
    return JJ_OK;
}

int subdomain_debug_solve(JJHandle* h, const double* b_d, double* J_d) {
    SubState* st = (SubState*)h->subdomain_plan;
    if (!st) { h->err = "subdomain engine: no plan"; return JJ_ESTATE; }
    if (!st->prepared) { int rc = subdomain_prepare(h); if (rc) return rc; }
    SubArgs a;
    fill_args(h, st, a);
    a.dbg_b = b_d; a.dbg_J = J_d;
    return launch(h, st, a);
}

// whether runs of this plan can store phase zones instead of phases (jj_anneal)
bool subdomain_stores_zones(JJHandle* h) {
    SubState* st = (SubState*)h->subdomain_plan;
    return st && st->prepared && st->threads == NT;
}

void subdomain_get_config(JJHandle* h, int* P, int* PC) {
    SubState* st = (SubState*)h->subdomain_plan;
    *P = st ? st->P : 1; *PC = st ? st->PC : 0;
}

}  // namespace jj

#endif  // JJ_SUB_NG
