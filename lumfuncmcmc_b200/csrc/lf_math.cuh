// lf_math.cuh -- FP64 building blocks of the walker x source loop, written for the sm_100a FP64 pipe.
//
// The hot loop is FP64-FMA-pipe bound (64 lanes/clk/SM), so every routine here is counted in FP64-pipe
// instructions; integer / MUFU / LDS work rides in the issue slots the FP64 pipe leaves free.
//   * reciprocal square root and reciprocal: one MUFU seed (rsqrt/rcp.approx.ftz.f64, ~2^-22) + one
//     third-order correction (error ~2^-63): 5 resp. 3 FP64 instructions, no IEEE div/sqrt sequences.
//   * exp: 64-entry table of 2^(j/64) (shared memory, replicated so a half-warp never bank-conflicts)
//     + degree-5 polynomial on |r| <= ln2/128.
//   * log: 256-entry table of (1/c_j, log c_j) (replicated x8) + degree-5 log1p on |eps| <= 2^-9.
// All of them are accurate to ~1e-16 absolute/relative over the ranges the classifier admits to the fast
// path (see k_prologue_* in lf_engine.cu); anything outside goes to the literal kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lfm {

constexpr double LN10 = 2.302585092994045684017991454684364208;
constexpr double LN2 = 0.693147180559945309417232121458176568;
constexpr double LOG2E = 1.442695040888963407359924681001892137;
constexpr double LOG2_10 = 3.321928094887362347870319429489390176;
constexpr double LNLN10 = 0.83403244524795579980321304785753909551;   // ln(ln 10)
constexpr double SQARCSEC = 42545170296.15221;                      // (180/pi*3600)^2, VmaxLumFunc.py:43
constexpr double MAGIC52 = 6755399441055744.0;                      // 1.5 * 2^52
constexpr double FOURPI = 12.566370614359172;                       // 4.0*np.pi
constexpr double MPC_CM_REF = 3.086e24;                             // the reference's Mpc->cm constant

constexpr int EXP_TAB_N = 64, EXP_TAB_REP = 16;   // doubles:  64*16*8  =  8 KB
constexpr int LOG_TAB_N = 256, LOG_TAB_REP = 8;   // double2: 256*8*16  = 32 KB

// polynomial / reduction constants live in the constant bank: DFMA reads them as c[3][off] operands, which keeps
// the inner loop free of the IMAD.MOV/UMOV pairs ptxas otherwise emits to materialise 64-bit immediates
__constant__ double KC[16] = {
    64.0 * LOG2E,          // 0
    -LN2 / 64.0,           // 1
    1.0 / 120.0,           // 2
    1.0 / 24.0,            // 3
    1.0 / 6.0,             // 4
    LN2,                   // 5
    0.2,                   // 6
    1.0 / 3.0,             // 7
    4503601774854144.0,    // 8   2^52 + 2^31
    MAGIC52,               // 9
    -0x1.62e42fefa38p-7,   // 10  -(ln2/64) high part
    -0x1.ef35793c7673p-51, // 11  -(ln2/64) low part
    LN10,                  // 12
    0.0, 0.0, 0.0};

struct Tables {                 // device-global master copies (filled by the host at lf_create)
    double exp2_frac[EXP_TAB_N];        // 2^(j/64)
    double2 log_tab[LOG_TAB_N];         // (invc_j, -log(invc_j)),  c_j = 1 + (j+0.5)/256
};

__device__ __forceinline__ double rsqrt_seed(double y) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    return r;
}
__device__ __forceinline__ double rcp_seed(double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    return r;
}

// 1/d for normal positive d, 3 FP64 instructions + 1 MUFU
__device__ __forceinline__ double rcp_fast(double d) {
    double r0 = rcp_seed(d);
    double e = fma(-d, r0, 1.0);
    double p = fma(e, e, e);
    return fma(r0, p, r0);
}

// cooperative fill of the replicated shared-memory tables
__device__ __forceinline__ void load_tables(const Tables* __restrict__ t, double* s_exp, double2* s_log) {
    for (int i = threadIdx.x; i < EXP_TAB_N * EXP_TAB_REP; i += blockDim.x) s_exp[i] = t->exp2_frac[i / EXP_TAB_REP];
    for (int i = threadIdx.x; i < LOG_TAB_N * LOG_TAB_REP; i += blockDim.x) s_log[i] = t->log_tab[i / LOG_TAB_REP];
}

// exp(x) - 1 pieces for x in (-2e7, 0]:  returns Ts = 2^K * 2^(j/64) and q = expm1(r), exp(x) = Ts*(1+q)
// 1 reduction FMA suffices here because only ABSOLUTE accuracy of exp(x) <= 1 is needed.
// FP64 instructions: 3 (t, kf, r) + 5 (q)
__device__ __forceinline__ void exp_neg_parts(double x, const double* s_exp, int rep, double& Ts, double& q) {
    double t = fma(x, KC[0], KC[9]);
    int k = __double2loint(t);
    double kf = t - KC[9];
    double r = fma(kf, KC[1], x);
    double T = s_exp[(k & 63) * EXP_TAB_REP + rep];
    int K = max(k >> 6, -1000);                 // exp(x) < 2^-1000 is 0 against 1; keeps the exponent field valid
    Ts = __hiloint2double(__double2hiint(T) + (K << 20), __double2loint(T));
    double p = fma(r, KC[2], KC[3]);
    p = fma(r, p, KC[4]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    q = p * r;
}

// full-range exp(x) with RELATIVE accuracy for x in [-708, 709]; below -708 returns 0.
// FP64 instructions: 4 (t, kf, r hi, r lo) + 5 (q) + 1 (Ts*q+Ts)
__device__ __forceinline__ double exp_full(double x, const double* s_exp, int rep) {
    bool under = x < -708.0;
    x = under ? -708.0 : x;
    double t = fma(x, KC[0], KC[9]);
    int k = __double2loint(t);
    double kf = t - KC[9];
    double r = fma(kf, KC[10], x);
    r = fma(kf, KC[11], r);
    double T = s_exp[(k & 63) * EXP_TAB_REP + rep];
    double Ts = __hiloint2double(__double2hiint(T) + ((k >> 6) << 20), __double2loint(T));
    double p = fma(r, KC[2], KC[3]);
    p = fma(r, p, KC[4]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    double q = p * r;
    double e = fma(Ts, q, Ts);
    return under ? 0.0 : e;
}

// log(v) for normal positive v. FP64 instructions: eps 1 + Ed 1 + base 1 + Horner 4 + final 1 = 8
__device__ __forceinline__ double log_fast(double v, const double2* s_log, int rep) {
    int hi = __double2hiint(v), lo = __double2loint(v);
    int j = (hi >> 12) & 0xff;
    int E = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    double2 tb = s_log[j * LOG_TAB_REP + rep];
    double eps = fma(m, tb.x, -1.0);
    double Ed = __hiloint2double(0x43300000, E ^ 0x80000000) - KC[8];   // int -> double
    double base = fma(Ed, KC[5], tb.y);
    double a = fma(eps, KC[6], -0.25);
    a = fma(eps, a, KC[7]);
    a = fma(eps, a, -0.5);
    a = fma(eps, a, 1.0);
    return fma(eps, a, base);
}

// t = ln(modified Fleming completeness) for one (walker, flux) pair -- the walker x source term.
//   n  = alpha*log10(f/F50) = fma(alpha, g, aF)            g = log10 f,  aF = -alpha*log10(F50)
//   fc = 1/2 (1 + n/sqrt(1+n^2))                           VmaxLumFunc.py:118-120
//   t  = ln(fc) / (1 - exp(-f/ftau))                       VmaxLumFunc.py:124-126, 141   (cinv = -1/ftau)
// MODIFIED=false: plain Fleming curve (fcmin falsy, VmaxLumFunc.py:121-122): t = ln(fc).
// FP64-pipe instruction count (MODIFIED): 9 + 8 + 11 + 3 = 31, +1 for the caller's accumulate.
template <bool MODIFIED>
__device__ __forceinline__ void fleming_log_parts(double g, double f, double alpha, double aF, double cinv,
                                                  const double* s_exp, const double2* s_log, int rep16, int rep8,
                                                  double& lg, double& rdec) {
    double n = fma(alpha, g, aF);
    double y = fma(n, n, 1.0);
    double r0 = rsqrt_seed(y);
    double h = y * r0;
    double e = fma(-h, r0, 1.0);
    double p = fma(0.375, e, 0.5);
    double pe = p * e;
    double nr = n * r0;
    double q = fma(nr, pe, nr);
    double fc = fma(0.5, q, 0.5);
    lg = log_fast(fc, s_log, rep8);
    if (MODIFIED) {
        double x = f * cinv;                                  // <= 0; |x| < 2e7 guaranteed by the classifier
        double Ts, qx;
        exp_neg_parts(x, s_exp, rep16, Ts, qx);
        double omT = 1.0 - Ts;
        double dec = fma(-Ts, qx, omT);                       // 1 - exp(x)
        rdec = rcp_fast(dec);
    } else {
        rdec = 1.0;
    }
}

}  // namespace lfm
