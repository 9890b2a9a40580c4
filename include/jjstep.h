/*
 * jjstep.h - C ABI of the B200 time-evolution engine (libjjstep.so).
 *
 * Drop-in boundary for ONE hot path of pyjjasim: the implicit RCSJ stepping loop
 *   time_evolution_core(problem, th_store_mask, I_store_mask) -> (th_out, I_out)
 *   (reference: /root/reference/time_evolution.py:461-582)
 * The reference has no FFI layer (it is pure Python); the seam is the Python call
 * TimeEvolutionProblem.compute() -> time_evolution() -> time_evolution_core()
 * (reference: time_evolution.py:303-307, :422-458, :461). Each entry point below cites the part of
 * that function it replaces. Plain pointers and sizes only; every host pointer is BORROWED for the
 * duration of the call and copied to the device inside it; outputs are copied into caller-allocated
 * buffers. All functions return 0 on success and a negative JJ_E* code on failure;
 * jj_last_error() returns a message. There is NO CPU fallback: without a CUDA device jj_create fails.
 *
 * Face indices crossing this interface are in the PERMUTED numbering chosen by the host-side
 * nested dissection (the permutation never needs to be undone on the device because all stored
 * outputs are per junction).
 *
 * Layouts: host arrays named (Nj, W) are C-ordered with the problem index minor, as in the reference.
 */
#ifndef JJSTEP_H
#define JJSTEP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct JJHandle JJHandle;

enum { JJ_OK = 0, JJ_ECUDA = -1, JJ_EINVAL = -2, JJ_ESTATE = -3, JJ_ENOMEM = -4, JJ_ENONFINITE = -5 };

/* which per-step input a source call refers to (reference: time_evolution.py:524-531) */
enum { JJ_SRC_IS = 0, JJ_SRC_F = 1, JJ_SRC_VS = 2, JJ_SRC_T = 3 };
/* device forms of an input (see pyjjasim_b200/sources.py) */
enum { JJ_KIND_ZERO = 0, JJ_KIND_RANK1 = 1, JJ_KIND_DENSE = 2 };
/* step engines */
enum { JJ_ENGINE_AUTO = 0,      /* SUBDOMAIN when a plan was set and the inputs are rank one, else STREAMING */
       JJ_ENGINE_STREAMING = 1, /* problem-minor (Nj, W) arrays in HBM, one kernel per phase: any input form
                                   (dense per-step tables, dense voltage sources); needs a real jj_set_solver program */
       /* 2 was the cluster-resident engine of round 1 (retired: the subdomain engine covers its shapes) */
       JJ_ENGINE_SUBDOMAIN = 3  /* persistent cooperative kernel: a thread block owns (subdomain of the elimination
                                   tree, chunk of problems) items; the separators above the subdomains are swept level
                                   by level as gathered dense FP64 tensor-core products over all SMs, and the last few
                                   tree levels by one dense inverse */ };

/* One sweep (forward or backward) of the compiled solve program, see pyjjasim_b200/factor.py.
 * Replaces the two SuperLU triangular sweeps per step (reference: time_evolution.py:506, :560-569). */
typedef struct {
    int32_t n_levels;
    const int32_t *level_ptr;    /* [n_levels+1] -> groups */
    const int32_t *group_ptr;    /* [n_groups+1] -> tiles; tiles of a group are processed in order */
    int32_t n_tiles;
    const int32_t *tile_row0;    /* first output row (permuted face index) */
    const int32_t *tile_nrows;   /* 1..32 */
    const int32_t *tile_lpr;     /* lanes per row; (32 / lpr) >= nrows */
    const int32_t *tile_nsteps;  /* columns = nsteps * lpr */
    const int32_t *tile_flags;   /* bit0: add src[row]; bit1: staged (block spans several tiles) */
    const int64_t *tile_col_off; /* into cols */
    const int64_t *tile_val_off; /* into vals */
    int64_t n_cols;  const int32_t *cols;
    int64_t n_vals;  const double  *vals;   /* [tile][step][32 lanes] */
    int32_t stage_rows;
    const int32_t *tile_stage_off; /* staged tiles: first staging row */
} JJSweep;

/* Subdomain engine: sweep program of one subdomain: the 8-row tiles of every (level, warp) pair packed as one
 * contiguous stream; levels [0, n_bwd) are the backward sweep, [n_bwd, n_levels) the forward sweep
 * (pyjjasim_b200/subdomain.py). */
typedef struct {
    int32_t n_levels, n_bwd, n_warps, n_tiles;
    const int32_t *wt_ptr;       /* [n_levels*n_warps][2] first / past-the-last tile of (level, warp) */
    const int32_t *ws_ptr;       /* [n_levels*n_warps] first stream step of (level, warp); tiles and steps are laid
                                    out warp-major inside a sweep, so a warp's stream continues across levels */
    const int32_t *thdr;         /* [n_tiles][2]: row0 | (nrows-1)<<16 | flags<<19 | g0<<21 | (ng-1)<<25 ; nsteps | stage_off<<16 */
    const int32_t *lstaged;      /* [n_levels] staged rows of the level */
    int64_t n_steps;
    const uint8_t *stream;       /* [n_steps][320]: 32 float64 (A fragment of an 8x4 block of the factor) then 32 uint16
                                    (shared-memory element codes of the B fragment; rows of 8*NG float64) */
} JJSubProgram;

/* Subdomain-engine plan (pyjjasim_b200/subdomain.py: subdomain_plan). Replaces the two SuperLU sweeps per step
 * (reference: time_evolution.py:506, :560-569). The elimination tree of the nested dissection is cut in three:
 *   - P mutually uncoupled SUBDOMAINS (leaves of the cut), swept out of shared memory by one thread block each;
 *   - the UPPER separators, rows [0, tt0) of the top numbering: swept depth by depth in global memory (three planes
 *     r, z, J of n_up_pad rows per problem chunk) as gathered dense products, two phases per tree depth
 *     (t = r - L[B, below] z ; z = inv(L[B,B]) t, and transposed for the backward sweep);
 *   - the TOP OF THE TOP, rows [tt0, tt0 + n_tt): the last few tree levels, one dense inverse of their Schur complement. */
typedef struct {
    int32_t P, NG;                   /* subdomains; problems per chunk = 8*NG (NG = 1, 2, 4, 8) */
    int32_t n_rows, n_loc_max, stage_rows;
    int32_t n_top, n_up_pad, n_slots;/* separator rows above the subdomains; row stride of the r / z / J planes */
    int32_t tt0, n_tt, n_tt_pad;     /* dense top of the top (n_tt_pad: n_tt rounded up to 32; tt0 + n_tt_pad <= n_up_pad) */
    const int32_t *n_loc, *n_halo;   /* [P] local rows / halo rows (coupled top rows) of each subdomain */
    const int32_t *hptr;             /* [P+1] first contribution slot of each subdomain */
    const int32_t *halo_top;         /* [n_slots] top row of each slot */
    const int32_t *tptr, *tslot;     /* [n_top+1], [n_slots]: slots that sum into each top row */
    const int32_t *top_face;         /* [n_top] permuted face index of each top row */
    const double  *Sinv_packed;      /* [n_tt_pad/8][n_tt_pad/4][8][4] MMA A fragments of the inverse Schur complement */
    /* upper program: phases [0, n_up_fwd) run after the top assembly, phases [n_up_fwd, n_up_fwd + n_up_bwd) after the
     * dense product; a grid barrier separates phases. A task computes out[rows] = V . X[cols] for up_RB 8-row tiles. */
    int32_t up_RB, up_KB;            /* most 8-row tiles of a task (16); the k-steps nk of a task are a multiple of up_KB (16) */
    int32_t n_up_fwd, n_up_bwd, n_up_tasks;
    const int32_t *up_phase_ptr;     /* [n_up_fwd + n_up_bwd + 1] first task of each phase */
    const int32_t *up_phase_split;   /* [n_up_fwd + n_up_bwd + 1] tasks [ptr, split) of a phase are run by whole thread blocks
                                        (sorted by decreasing cost), tasks [split, next ptr) by single warps (<= 4 tiles) */
    const int32_t *up_task;          /* [n_up_tasks][4]: out code, rows (1..8*up_RB), nk, first column (index into up_cols) */
    const int64_t *up_task_aoff;     /* [n_up_tasks] offset of the task's A fragments in up_A (float64 units) */
    int64_t n_up_cols; const int32_t *up_cols;   /* row codes: plane << 28 | row, plane 0 = r, 1 = z, 2 = J, 3 = scratch (partial
                                                    products of column-split tasks, summed by a phase of their own) */
    int64_t n_up_vals; const double *up_A;       /* per task [row tile][nk][32]: lane = row*4 + kk holds V[row][4k + kk] */
    const JJSubProgram *prog;        /* [P] */
    const int32_t *junc_ptr;         /* [P+1] subdomain s owns device junctions junc_ptr[s]:junc_ptr[s+1] */
    const int32_t *junc_orig;        /* [Nj] */
    const int32_t *junc_row;         /* [Nj*2] shared-memory rows of its faces on the owner, -1 none */
    const int8_t  *junc_sign;        /* [Nj*2] */
    int32_t face_K;                  /* fixed width of the face lists (multiple of 4) */
    const int32_t *face_ell_j;       /* [P][n_rows][face_K] device junction, -1 pad */
    const double  *face_ell_c;       /* [P][n_rows][face_K] sign / c0 */
    const int32_t *face_fidx;        /* [P][n_rows] permuted face index of local rows, -1 for halo rows */
} JJSubdomainPlan;

/* Circuit constants, per junction, precomputed on the host in float64 exactly as the reference does
 * (reference: time_evolution.py:470-478): Rv = 1/(dt R), Cv = C/dt^2, c0 = Rv+Cv, c1 = -Rv-2Cv, c2 = Cv. */
typedef struct {
    int32_t Nj, Nf;
    const int32_t *face_ptr;    /* [Nf+1] CSR rows of the cycle matrix A, permuted face order */
    const int32_t *face_junc;   /* junction index per entry */
    const int8_t  *face_sign;   /* +1 / -1 */
    const int32_t *junc_face;   /* [Nj*2] the (at most two) faces of each junction, -1 if none */
    const int8_t  *junc_sign;   /* [Nj*2] */
    const double *Ic, *c0, *c1, *c2;   /* [Nj] */
    int32_t cpr_harmonics;      /* M >= 1; Icp = Ic * sum_{m<=M} a[m] cos(m th) + b[m] sin(m th) */
    const double *cpr_a, *cpr_b; /* [M+1]; DefaultCPR is M=1, a={0,0}, b={0,1} (reference: static_problem.py:77-85) */
} JJCircuit;

/* multiprocessor count of a CUDA device (the host sizes the subdomain plan to it); <= 0 on failure */
int  jj_sm_count(int device);
/* create / destroy an engine bound to one CUDA device */
int  jj_create(int device, JJHandle **out);
void jj_destroy(JJHandle *h);
const char *jj_last_error(const JJHandle *h);   /* h may be NULL: error of the last failed jj_create */

/* setup (reference: time_evolution.py:466-519) */
int jj_set_circuit(JJHandle *h, const JJCircuit *c);
/* the streaming engine's solve program; sweeps with n_levels == 0 are accepted when only the subdomain engine runs */
int jj_set_solver(JJHandle *h, const JJSweep *fwd, const JJSweep *bwd);
/* optional: enables JJ_ENGINE_SUBDOMAIN. Must follow jj_set_circuit/jj_set_solver. plan == NULL removes it. */
int jj_set_subdomain_plan(JJHandle *h, const JJSubdomainPlan *plan);
/* W problems, time step dt, Philox seed, index of this shard's first problem in the global batch
 * (keeps noise identical however the batch is sharded over GPUs; must be a multiple of 4) */
int jj_set_problem(JJHandle *h, int32_t W, double dt, uint64_t seed, int64_t problem_offset, int32_t engine);
/* change the step engine of the current problem between two jj_run calls; the state is handed over on the device
 * (the host does this when a callable input stops being of the form base x amplitude and must be uploaded densely) */
int jj_set_engine(JJHandle *h, int32_t engine);
/* theta(-1), theta(-2): (Nj, W) host arrays (reference: time_evolution.py:480,490; config_at_minus_1/2) */
int jj_set_state(JJHandle *h, const double *theta_m1, const double *theta_m2);
int jj_get_state(JJHandle *h, double *theta_m1, double *theta_m2);   /* either output may be NULL */

/* per-step inputs (reference: time_evolution.py:342-356, :524-531).
 * RANK1: value(e,w,i) = base[e] * amp[i - i0][w];  DENSE: value = table[i - i0][e][w].
 * is_static != 0: a single row/plane valid for all steps. For JJ_SRC_T the host passes
 * base = sqrt(2 * Tbase * Rv) and amp = sqrt(Tamp) (RANK1) so that the device multiplies once;
 * for JJ_SRC_VS (RANK1) amp is the running sum of Vs*dt BEFORE step i (reference: :579-580). */
int jj_set_source(JJHandle *h, int32_t which, int32_t kind, int32_t is_static, const double *base /* [N] or NULL */);
int jj_upload_source(JJHandle *h, int32_t which, int64_t i0, int32_t K, const double *table /* (K,W) or (K,N,W) */);
/* standard-normal draws to use instead of the Philox generator for steps [i0, i0+K): (K, Nj, W); K = 0 disables */
int jj_upload_noise(JJHandle *h, int64_t i0, int32_t K, const double *Z);

/* output planes kept on the device: theta and current snapshots, (plane, Nj, W) */
int jj_alloc_outputs(JJHandle *h, int64_t n_theta_planes, int64_t n_current_planes);
/* Run steps [i0, i0+n) (reference: time_evolution.py:523-580). th_plane[k] / I_plane[k] >= 0: store
 * theta / current of step i0+k into that plane; -1: do not store. */
int jj_run(JJHandle *h, int64_t i0, int32_t n, const int64_t *th_plane, const int64_t *I_plane);
int jj_fetch_theta(JJHandle *h, int64_t plane0, int64_t n_planes, double *dst /* (n_planes, Nj, W) */);
int jj_fetch_current(JJHandle *h, int64_t plane0, int64_t n_planes, double *dst);
/* device generator check: the standard normal draws the device would use at one step, (Nj, W) */
int jj_debug_noise(JJHandle *h, int64_t step, double *dst);
/* one solve J = S^-1 b through the compiled program, (Nf, W) host arrays in permuted face order */
int jj_debug_solve(JJHandle *h, const double *b, double *J);
/* the same through the subdomain engine's cooperative kernel (requires a subdomain plan) */
int jj_debug_subdomain_solve(JJHandle *h, const double *b, double *J);

/* ---- the annealing caller of the path (reference: time_evolution.py:1070-1191, AnnealingProblem.compute) ----
 * The reference re-enters compute() every `interval_steps` steps with theta(-2) := theta(-1), pulls every theta
 * plane to the host, derives the vortex configurations there and adapts the temperatures. These three entry points
 * keep the state and the stored planes on the device between intervals. */
/* zero-velocity restart: theta(-2) := theta(-1) on the device (reference: time_evolution.py:1169-1171) */
int jj_restart_at_rest(JJHandle *h);
/* the closing runs of the schedule use another time step, i.e. another handle (its own factor): theta(-1) = theta(-2)
 * := theta(-1) of `from`, device to device (same device, same junction and problem counts)
 * (reference: time_evolution.py:1176-1183 hands the phases over through the host) */
int jj_adopt_state_at_rest(JJHandle *h, JJHandle *from);
/* n = -A round(theta / 2 pi) of stored theta plane `plane` (-1: the current state theta(-1)); dst is (Nf, W) int32,
 * faces in the PERMUTED order (reference: time_evolution.py:734-755, get_vortex_configuration) */
int jj_vortex_configuration(JJHandle *h, int64_t plane, int32_t *dst);
/* dst[w] = sum over faces and over consecutive planes p0 <= p < p0+n-1 of |n(p+1) - n(p)|: the numerator of the
 * reference's vortex mobility (reference: time_evolution.py:1128-1133, get_vortex_mobility); exact integers */
int jj_vortex_mobility(JJHandle *h, int64_t plane0, int64_t n_planes, int64_t *dst /* [W] */);

/* The whole temperature schedule without host round trips: for interval i = first_interval .. + n_intervals - 1
 *   noise amplitude := sqrt(T) (zero for all once every T is numerically zero, time_evolution.py:512); theta(-2) := theta(-1)
 *   (not before interval 0); `steps` time steps; the interval's exact integer mobility sums;
 *   T *= (sum / norm > upper[i]) ? inv_T_factor : T_factor; profiles[i] := T        (reference: time_evolution.py:1128-1140)
 * The temperature must have been declared jj_set_source(JJ_SRC_T, JJ_KIND_RANK1, static) with one amplitude row uploaded,
 * and jj_alloc_outputs must provide `steps` theta planes: they are SCRATCH of the schedule (the streaming engine keeps the
 * interval's phases there; the subdomain engine keeps one byte per junction and problem, round(theta / 2 pi) mod 256, from
 * which the mobility sums are exact) - their contents are unspecified afterwards. T: (W,) in/out; profiles:
 * (n_intervals, W) out; the arithmetic is the reference's on the same integers, so the schedule is bit-identical to one
 * evaluated on the host. */
int jj_anneal(JJHandle *h, int64_t first_interval, int32_t n_intervals, int32_t steps, const double *upper,
              double T_factor, double inv_T_factor, double norm, double *T, double *profiles, double *device_ms);
/* all stored theta planes [plane0, plane0 + n_planes) at once: dst is (n_planes, Nf, W) int32, permuted faces. With it the
 * vortex configurations of the stored steps leave the device as integers and the theta planes never have to
 * (reference: time_evolution.py:734-755 evaluated on host copies of theta) */
int jj_vortex_configurations(JJHandle *h, int64_t plane0, int64_t n_planes, int32_t *dst,
                             const int32_t *face_order /* [Nf] output row of each permuted face, or NULL: permuted order */);

/* ---- running observables: accumulated on the device WHILE stepping, no theta plane stored or copied ----
 * Steps first_step + m * interval (m = 0, 1, ...) are OBSERVATIONS. At every observation the step kernel itself adds
 * the vortex configuration n = -A round(theta / 2 pi) of that step into a per-(face, problem) integer sum and keeps
 * theta of the first and of the latest observation, from which the host forms the quantities the reference's users
 * average over stored planes (reference: time_evolution.py:734-755 and :422-458 over all stored steps):
 *   time-averaged vortex configuration = nsum / count                     (exact integers)
 *   DC junction voltage = (theta_latest - theta_first) / ((count - 1) interval dt)   (the stencil sum telescopes)
 * The accumulators belong to the problem: they survive chunked jj_run calls and are cleared by jj_observe_begin
 * (interval = 0 switches observation off). first_step must not lie before the steps already run. */
int jj_observe_begin(JJHandle *h, int64_t first_step, int32_t interval);
/* count = observations so far; nsum (Nf, W) int32; theta_first / theta_latest (Nj, W); any pointer may be NULL */
int jj_observe_fetch(JJHandle *h, int64_t *count, int32_t *nsum, double *theta_first, double *theta_latest,
                     const int32_t *face_order /* as for jj_vortex_configurations */);

/* Page-locked host memory for result planes: jj_fetch_* into such a buffer is a single DMA at PCIe speed instead of
 * a staged copy into pageable memory. The Python wrapper pools these blocks and hands them out as numpy arrays. */
int jj_host_alloc(int device, uint64_t bytes, void **out);   /* device: whose context pins the block (portable) */
int jj_host_free(void *p);

typedef struct {
    int32_t engine;             /* engine actually used */
    int32_t cluster_size, tile_problems;
    int64_t steps_done;
    int64_t kernel_launches;    /* launches of OUR kernels since jj_set_problem */
    double  step_ms;            /* device time of the last jj_run (CUDA events on the run stream) */
    int64_t device_bytes;       /* device memory held */
    int32_t non_finite;         /* 1 if a non-finite phase was seen */
} JJStats;
int jj_stats(JJHandle *h, JJStats *out);

#ifdef __cplusplus
}
#endif
#endif
