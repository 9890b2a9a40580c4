// Device helpers shared by the streaming and the resident step engines (sm_100a).
//
// Arithmetic contract (reference: /root/reference/time_evolution.py:533-580), per junction j, problem w:
//   theta_n   = (y - x) / c0                                   y = (A^T J)[j]
//   X         = Ic * cpr(2 theta_n - theta_{n-1}) + c1 theta_n + c2 theta_{n-1}
//   x'        = (noise - Is) + X                               noise = sqrt(2 T Rv) * Z,  Z ~ N(0,1)
//   b[f]      = sum_j A[f,j] (x'_j / c0_j - theta_s_j) - 2 pi f_f
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jj {

enum { KIND_ZERO = 0, KIND_RANK1 = 1, KIND_DENSE = 2 };

// One per-step input on the device. `table` rows are indexed by (step - i0), or 0 when is_static.
struct Source {
    int kind;
    int is_static;
    long long i0;
    int K;
    const double* base;    // [N]           (RANK1)
    const double* table;   // [K][ld]       (RANK1: ld = Wp)   or [K][N][ld] (DENSE)
};

__device__ __forceinline__ long long source_row(const Source& s, long long step) {
    return s.is_static ? 0 : (step - s.i0);
}

// value(e, w..w+3, step) for four consecutive problems
__device__ __forceinline__ void source_eval4(const Source& s, long long step, int e, int N, int ld, int w,
                                             double out[4]) {
    if (s.kind == KIND_ZERO) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
    long long r = source_row(s, step);
    if (s.kind == KIND_RANK1) {
        double b = s.base[e];
        const double2* t = reinterpret_cast<const double2*>(s.table + r * ld + w);
        double2 a0 = t[0], a1 = t[1];
        out[0] = b * a0.x; out[1] = b * a0.y; out[2] = b * a1.x; out[3] = b * a1.y;
    } else {
        const double2* t = reinterpret_cast<const double2*>(s.table + (r * N + e) * (long long)ld + w);
        double2 a0 = t[0], a1 = t[1];
        out[0] = a0.x; out[1] = a0.y; out[2] = a1.x; out[3] = a1.y;
    }
}

// ---------------------------------------------------------------------------------------------
// Counter-based noise: Philox4x32-10 (Salmon et al., SC'11), key = seed, counter = (junction,
// problem group = w/4, step lo, step hi). The four 32-bit outputs make two Box-Muller pairs, i.e. the
// four standard normals of problems 4g..4g+3 at this junction and step. Single precision
// transcendentals are enough for thermal noise (relative error 1e-6 on a random number) and keep the
// FP64 pipe free for the phase update. Replaces np.random.randn (reference: time_evolution.py:533-538).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    // u in (0, 1]: all 32 bits are kept where it matters (small a is exact in float -> 6.7 sigma tail);
    // angle in [0, 2pi)
    float u = fminf((float(a) + 0.5f) * 2.3283064365386963e-10f, 1.0f);
    float v = float(b >> 8) * (1.0f / 16777216.0f);
    // u >= 2^-33 is a normal float and -2 ln u lies in [0, 46]: the flush-to-zero approximations need none of the
    // denormal / special-case fix-ups that sqrtf and __logf carry (one MUFU each instead of ~12 instructions)
    float l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));     // -2 ln 2 * log2 u
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// Double-precision variant (build with -DJJ_NOISE_F64): the same uniforms through log / sqrt / sincospi in float64. Kept
// as a compile-time alternative so that its cost and its effect on the statistics can be measured
// (profiles/r02_noise_precision.md); the shipped build uses the single-precision transform above.
__device__ __forceinline__ void box_muller_f64(uint32_t a, uint32_t b, double& z0, double& z1) {
    const double u = ((double)a + 0.5) * 2.3283064365386963e-10;      // (0, 1)
    const double v = ((double)b + 0.5) * 2.3283064365386963e-10;
    const double r = sqrt(-2.0 * log(u));
    double s, c;
    sincospi(2.0 * v, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// Phase zone of a junction: np.round(theta / (2.0 * np.pi)) - a true division, rounded to nearest-even. The product with
// the reciprocal differs from the quotient by less than 2 ulp, i.e. by less than 1e-9 for |theta / 2 pi| < 1e6, so it
// rounds to the same integer unless it lies within 1e-6 of a tie; only then (and for huge or non-finite phases) the
// division itself is evaluated. One multiply + round instead of a ~25-instruction FP64 division per junction.
__device__ __forceinline__ double phase_zone(double th) {
    const double t = th * 0.15915494309189535, q = rint(t);
    if (fabs(t - q) < 0.499999 && fabs(t) < 1.0e6) return q;
    return rint(th / 6.283185307179586);
}

__device__ __forceinline__ void normal4(uint64_t seed, int junction, long long group, long long step, double z[4]) {
    uint32_t o[4];
    philox4x32_10((uint32_t)junction, (uint32_t)group, (uint32_t)step, (uint32_t)((unsigned long long)step >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
#ifdef JJ_NOISE_F64
    box_muller_f64(o[0], o[1], z[0], z[1]);
    box_muller_f64(o[2], o[3], z[2], z[3]);
#else
    float a, b, c, d;
    box_muller(o[0], o[1], a, b);
    box_muller(o[2], o[3], c, d);
    z[0] = a; z[1] = b; z[2] = c; z[3] = d;
#endif
}

// ---------------------------------------------------------------------------------------------
// Current-phase relation Ic * g(theta), g a trigonometric polynomial (DefaultCPR: g = sin).
// (reference: static_problem.py:34-85; custom relations are fitted on the host, current_phase_relation.py)
// ---------------------------------------------------------------------------------------------
struct Cpr {
    int M;
    double a[17], b[17];
};

// sin / cos for |x| < 1e8 with ~20 FP64 instructions (the library sin() spends 2-3x that on its general
// argument reduction): k = rint(x/pi), r = x - k*pi with a two-term FMA reduction (error ~1e-16 + k*1e-33),
// odd/even Taylor polynomials on [-pi/2, pi/2] (truncation < 2e-18), sign (-1)^k. Checked against sinl():
// max abs error 2.2e-16 for |x| up to 1e8. Phases reach 1e3-1e4 rad in long runs; anything
// larger than 1e8 takes the library path.
__device__ __forceinline__ void fast_sincos(double x, double* sn, double* cs, bool want_cos) {
    if (!(fabs(x) < 1.0e8)) {
        if (want_cos) sincos(x, sn, cs); else *sn = sin(x);
        return;
    }
    const double k = rint(x * 0.31830988618379067154);
    double r = fma(-k, 3.141592653589793116, x);
    r = fma(-k, 1.2246467991473532072e-16, r);
    const double r2 = r * r;
    double p = -1.9572941063391261e-20;
    p = fma(p, r2, 8.2206352466243295e-18);
    p = fma(p, r2, -2.8114572543455206e-15);
    p = fma(p, r2, 7.6471637318198164e-13);
    p = fma(p, r2, -1.6059043836821613e-10);
    p = fma(p, r2, 2.5052108385441720e-08);
    p = fma(p, r2, -2.7557319223985893e-06);
    p = fma(p, r2, 1.9841269841269841e-04);
    p = fma(p, r2, -8.3333333333333332e-03);
    p = fma(p, r2, 1.6666666666666666e-01);
    double s = fma(-r * r2, p, r);
    const bool odd = ((long long)k) & 1;
    *sn = odd ? -s : s;
    if (want_cos) {
        double q = 4.1103176233121648e-19;           //  1/20!
        q = fma(q, r2, -1.5619206968586225e-16);     // -1/18!
        q = fma(q, r2, 4.7794773323873853e-14);      //  1/16!
        q = fma(q, r2, -1.1470745597729725e-11);     // -1/14!
        q = fma(q, r2, 2.0876756987868100e-09);      //  1/12!
        q = fma(q, r2, -2.7557319223985888e-07);     // -1/10!
        q = fma(q, r2, 2.4801587301587302e-05);      //  1/8!
        q = fma(q, r2, -1.3888888888888889e-03);     // -1/6!
        q = fma(q, r2, 4.1666666666666664e-02);      //  1/4!
        q = fma(q, r2, -0.5);
        double c = fma(q, r2, 1.0);
        *cs = odd ? -c : c;
    }
}

// Four sines at once, branch-free on the common path: the Horner steps of the four arguments interleave
// (independent FP64 chains hide the ~9-cycle DFMA latency) and the coefficients come from the constant bank
// as direct instruction operands instead of 64-bit immediates rebuilt per use.
static __constant__ double c_sin_poly[10] = {
    -1.9572941063391261e-20, 8.2206352466243295e-18, -2.8114572543455206e-15, 7.6471637318198164e-13,
    -1.6059043836821613e-10, 2.5052108385441720e-08, -2.7557319223985893e-06, 1.9841269841269841e-04,
    -8.3333333333333332e-03, 1.6666666666666666e-01};
static __constant__ double c_sin_red[3] = {0.31830988618379067154, 3.141592653589793116, 1.2246467991473532072e-16};

__device__ __forceinline__ void fast_sin4(const double x[4], double out[4]) {
    const bool ok = (fabs(x[0]) < 1.0e8) & (fabs(x[1]) < 1.0e8) & (fabs(x[2]) < 1.0e8) & (fabs(x[3]) < 1.0e8);
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = sin(x[i]);
        return;
    }
    double k[4], r[4], r2[4], p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k[i] = rint(x[i] * c_sin_red[0]);
        r[i] = fma(-k[i], c_sin_red[1], x[i]);
        r[i] = fma(-k[i], c_sin_red[2], r[i]);
        r2[i] = r[i] * r[i];
        p[i] = c_sin_poly[0];
    }
#pragma unroll
    for (int m = 1; m < 10; ++m) {
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r2[i], c_sin_poly[m]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double sv = fma(-r[i] * r2[i], p[i], r[i]);
        // sign (-1)^k: flip the sign bit when k is odd (k is integral and |k| < 2^31 here)
        const int ki = __double2int_rn(k[i]);
        out[i] = __hiloint2double(__double2hiint(sv) ^ (ki << 31), __double2loint(sv));
    }
}

template <bool DEFAULT>
__device__ __forceinline__ void cpr_eval4(const Cpr& c, const double th[4], double out[4]);

template <bool DEFAULT>
__device__ __forceinline__ double cpr_eval(const Cpr& c, double th) {
    double s, co;
    if (DEFAULT) { fast_sincos(th, &s, &co, false); return s; }
    fast_sincos(th, &s, &co, true);
    double acc = c.a[0] + c.a[1] * co + c.b[1] * s;
    double cm = co, sm = s;
    for (int m = 2; m <= c.M; ++m) {
        double cn = cm * co - sm * s;
        double sn = sm * co + cm * s;
        cm = cn; sm = sn;
        acc += c.a[m] * cm + c.b[m] * sm;
    }
    return acc;
}

template <bool DEFAULT>
__device__ __forceinline__ void cpr_eval4(const Cpr& c, const double th[4], double out[4]) {
    if (DEFAULT) { fast_sin4(th, out); return; }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = cpr_eval<false>(c, th[i]);
}

}  // namespace jj
