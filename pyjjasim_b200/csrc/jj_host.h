// Host-side engine state behind the opaque JJHandle (see include/jjstep.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/jjstep.h"
#include "jj_device.cuh"

namespace jj {

struct SweepDev {
    int n_levels = 0, n_tiles = 0, n_groups = 0, stage_rows = 0;
    std::vector<int> level_ptr, group_ptr;      // host copies for launch configuration
    std::vector<int> tile_row0_h, tile_nrows_h, tile_flags_h;
    int *group_ptr_d = nullptr;
    int *tile_row0 = nullptr, *tile_nrows = nullptr, *tile_lpr = nullptr, *tile_nsteps = nullptr, *tile_flags = nullptr;
    long long *tile_col_off = nullptr, *tile_val_off = nullptr;
    int *cols = nullptr;
    double *vals = nullptr;
    long long n_cols = 0, n_vals = 0;
};

// kernel-side view of a sweep
struct SweepView {
    const int *group_ptr, *tile_row0, *tile_nrows, *tile_lpr, *tile_nsteps, *tile_flags;
    const long long *tile_col_off, *tile_val_off;
    const int *cols;
    const double *vals;
};

struct CircuitDev {
    int Nj = 0, Nf = 0;
    int *face_ptr = nullptr, *face_junc = nullptr;
    signed char *face_sign = nullptr;
    int *junc_face = nullptr;
    signed char *junc_sign = nullptr;
    double *Ic = nullptr, *c0 = nullptr, *c1 = nullptr, *c2 = nullptr;
    Cpr cpr;
    bool default_cpr = true;
    int max_face_len = 0;
};

struct SourceHost {
    Source dev{};            // device view (pointers are device pointers)
    double *base_buf = nullptr, *table_buf = nullptr;
    size_t table_cap = 0;    // bytes
    int N = 0;
};

}  // namespace jj

struct JJHandle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    jj::CircuitDev cir;
    jj::SweepDev fwd, bwd;
    bool have_circuit = false, have_solver = false, have_problem = false, have_state = false;
    // problem
    int W = 0, Wp = 0;
    double dt = 0.0;
    uint64_t seed = 0;
    long long problem_offset = 0;
    int engine_req = 0, engine = 0;
    // state (streaming engine): problem-minor [Nj][Wp]
    double *th1 = nullptr, *th2 = nullptr, *x = nullptr, *thetas = nullptr, *v = nullptr;
    jj::SourceHost src[4];
    // injected noise
    double *noise_buf = nullptr; size_t noise_cap = 0; long long noise_i0 = 0; int noise_K = 0;
    // outputs
    double *th_out = nullptr, *I_out = nullptr; long long n_th_planes = 0, n_I_planes = 0;
    long long th_cap_planes = 0, I_cap_planes = 0;   // allocated planes (kept across jj_set_problem calls of equal W)
    int *flag_d = nullptr;
    void *scratch = nullptr; size_t scratch_cap = 0;     // grow-only device scratch of the observable kernels (jj_observe.cu)
    // running observables (jj_observe_begin): observations at steps obs_first + m * obs_interval
    long long obs_first = 0, obs_count = 0; int obs_interval = 0;
    int *obs_nsum = nullptr;                                  // [Nf][Wp] permuted faces
    double *obs_th_first = nullptr, *obs_th_last = nullptr;   // canonical [Nj][Wp]
    size_t obs_n_bytes = 0, obs_th_bytes = 0;
    // annealing (jj_anneal): the subdomain engine stores phase zones (one byte each) instead of phases, [planes][Nj][Wp]
    unsigned char* zone8 = nullptr;
    unsigned long long src_gen = 1;   // bumped by jj_set_source: the subdomain engine re-gathers its per-junction / per-row constants
    bool start_at_rest = false;       // the next subdomain run reads theta(-2) from theta(-1) (jj_anneal: no restart copy)
    // subdomain engine (see jj_subdomain.cu)
    void *subdomain_plan = nullptr;
    // stats
    long long steps_done = 0, launches = 0, device_bytes = 0;
    double last_ms = 0.0;
    int non_finite = 0;
};

namespace jj {
// implemented in jj_subdomain.cu
int subdomain_supported(JJHandle* h, std::string& why_not);
int subdomain_prepare(JJHandle* h);
bool subdomain_prepared(JJHandle* h);
bool subdomain_stores_zones(JJHandle* h);
int subdomain_run(JJHandle* h, long long i0, int n, const long long* th_plane, const long long* I_plane);
void subdomain_free_problem(JJHandle* h);
int subdomain_set_plan(JJHandle* h, const JJSubdomainPlan* plan);
void subdomain_drop_plan(JJHandle* h);
int subdomain_debug_solve(JJHandle* h, const double* b_d, double* J_d);
void subdomain_get_config(JJHandle* h, int* P, int* PC);
// implemented in jjstep.cu: enqueue steps on the stream, no wait (see jj_run)
int run_enqueue(JJHandle* h, long long i0, int n, const long long* th_plane, const long long* I_plane);
// implemented in jj_observe.cu
void observe_free(JJHandle* h);
void observe_off(JJHandle* h);
bool observed_step(const JJHandle* h, long long step);
long long observations_in(const JJHandle* h, long long i0, long long n);      // observations among steps [i0, i0 + n)
int observe_streaming(JJHandle* h, long long step, const double* theta);      // streaming engine: one observation
int dev_alloc(JJHandle* h, void** p, size_t bytes);
void dev_free(JJHandle* h, void* p, size_t bytes);
}
