// lf_math.cuh -- FP64 building blocks of the walker x source loop, written for the sm_100a FP64 pipe.
//
// The hot loop is FP64-pipe bound (64 lanes/clk/SM; ncu: math-pipe-throttle is the dominant stall), so every
// routine here is counted in FP64-pipe instructions, and in how many of them read THREE distinct register operands
// (measured on B200: a DFMA with three register sources sustains 76 % of the rate of one with an immediate /
// constant-bank operand).  Integer / MUFU / LDS work rides in the issue slots the FP64 pipe leaves free.
//   * reciprocal square root and reciprocal: one MUFU seed (rsqrt/rcp.approx.ftz.f64, ~2^-22) + one
//     third-order correction (error ~2^-63): 5 resp. 3 FP64 instructions, no IEEE div/sqrt sequences.
//   * exp: 256-entry table of 2^(j/256) in shared memory + degree-4 polynomial on |r| <= ln2/512.
//   * log: one table lookup indexed by the top 20 bits of the argument (exponent AND 8 mantissa bits, covering
//     [2^-12, 1]) returning (1/c, ln c) with the exponent folded in, + degree-5 log1p on |eps| <= 2^-9.  No
//     exponent extraction, no int->double conversion.
// All of them are accurate to ~1e-16 absolute/relative over the ranges the classifier admits to the fast
// path (see k_prologue in lf_engine.cu); anything outside goes to the literal kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lfm {

constexpr double LN10 = 2.302585092994045684017991454684364208;
constexpr double LN2 = 0.693147180559945309417232121458176568;
constexpr double LOG2E = 1.442695040888963407359924681001892137;
constexpr double LNLN10 = 0.83403244524795579980321304785753909551;   // ln(ln 10)
constexpr double SQARCSEC = 42545170296.15221;                      // (180/pi*3600)^2, VmaxLumFunc.py:43
constexpr double MAGIC52 = 6755399441055744.0;                      // 1.5 * 2^52
constexpr double FOURPI = 12.566370614359172;                       // 4.0*np.pi
constexpr double MPC_CM_REF = 3.086e24;                             // the reference's Mpc->cm constant

#ifndef LF_EXP_REP
#define LF_EXP_REP 16
#endif
#ifndef LF_LOG_REP
#define LF_LOG_REP 2
#endif
constexpr int EXP_TAB_BITS = 8, EXP_TAB_N = 1 << EXP_TAB_BITS, EXP_TAB_REP = LF_EXP_REP;   // 256*16*8 B = 32 KB, replicated x16: a half-warp never bank-conflicts
constexpr int LOG_OCTAVES = 12, LOG_MANT_BITS = 8;
constexpr int LOG_TAB_N = LOG_OCTAVES * (1 << LOG_MANT_BITS) + 1, LOG_TAB_REP = LF_LOG_REP;  // 3073*2*16 B = 96 KB, two replicas (even / odd lanes)
constexpr int LOG_TAB_BASE = (1023 - LOG_OCTAVES) << LOG_MANT_BITS;                        // index of 2^-12 in (hi >> 12)
constexpr double LOG_ARG_MIN = 0x1p-12;

// polynomial / reduction constants live in the constant bank: DFMA reads them as c[3][off] operands (two register
// sources instead of three, and no IMAD.MOV/UMOV pairs to materialise 64-bit immediates)
__constant__ double KC[16] = {
    256.0 * LOG2E,          // 0
    -LN2 / 256.0,           // 1
    1.0 / 24.0,             // 2
    1.0 / 6.0,              // 3
    0.2,                    // 4
    1.0 / 3.0,              // 5
    MAGIC52,                // 6
    -0x1.62e42fee00000p-9,  // 7   -(ln2/256) high part (32 significant bits)
    -0x1.a39ef35793c76p-41, // 8   -(ln2/256) low part
    LN10,                   // 9
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0};

struct Tables {                 // device-global master copies (filled by the host at lf_create)
    double exp2_frac[EXP_TAB_N];        // 2^(j/256)
    double2 log_tab[LOG_TAB_N + 1];     // bin b <-> argument bits (hi >> 12) == LOG_TAB_BASE + b:
                                        //   (1/c_b, ln c_b), c_b = bin centre incl. its power of two; last: (1, 0)
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

// cooperative fill of the (replicated) shared-memory tables
__device__ __forceinline__ void load_tables(const Tables* __restrict__ t, double* s_exp, double2* s_log) {
    for (int i = threadIdx.x; i < EXP_TAB_N * EXP_TAB_REP; i += blockDim.x) s_exp[i] = t->exp2_frac[i / EXP_TAB_REP];
    for (int i = threadIdx.x; i < LOG_TAB_N * LOG_TAB_REP; i += blockDim.x) s_log[i] = t->log_tab[i / LOG_TAB_REP];
}

// 1 - exp(x) for x in (-2e7, 0]; ABSOLUTE accuracy ~1e-16 (single-FMA range reduction suffices because exp(x) <= 1).
// FP64 instructions: t, kf, r (3) + Horner (4) + 1 = 8
__device__ __forceinline__ double one_minus_exp_neg(double x, const double* s_exp, int rep) {
    double t = fma(x, KC[0], KC[6]);
    int k = __double2loint(t);
    double kf = t - KC[6];
    double r = fma(kf, KC[1], x);                 // |r| <= ln2/512
    double T = s_exp[(k & (EXP_TAB_N - 1)) * EXP_TAB_REP + rep];
    int K = max(k >> EXP_TAB_BITS, -1000);       // exp(x) < 2^-1000 is 0 against 1; keeps the exponent field valid
    double Ts = __hiloint2double(__double2hiint(T) + (K << 20), __double2loint(T));
    double p = fma(r, KC[2], KC[3]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);                           // exp(r)
    return fma(-Ts, p, 1.0);
}

// full-range exp(x) with RELATIVE accuracy for x in [-708, 709]; below -708 returns 0.
// FP64 instructions: compare 1 + t, kf, r hi, r lo (4) + Horner (4) + scale (1) = 10
__device__ __forceinline__ double exp_full(double x, const double* s_exp, int rep) {
    bool under = x < -708.0;
    x = under ? -708.0 : x;
    double t = fma(x, KC[0], KC[6]);
    int k = __double2loint(t);
    double kf = t - KC[6];
    double r = fma(kf, KC[7], x);
    r = fma(kf, KC[8], r);
    double T = s_exp[(k & (EXP_TAB_N - 1)) * EXP_TAB_REP + rep];
    double Ts = __hiloint2double(__double2hiint(T) + ((k >> EXP_TAB_BITS) << 20), __double2loint(T));
    double p = fma(r, KC[2], KC[3]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = p * r;                                    // expm1(r)
    double e = fma(Ts, p, Ts);
    return under ? 0.0 : e;
}

// log(v) for v in [2^-12, 1].  FP64 instructions: eps 1 + Horner 4 + final 1 = 6
__device__ __forceinline__ double log_unit(double v, const double2* s_log, int rep) {
    int b = max((__double2hiint(v) >> (20 - LOG_MANT_BITS)) - LOG_TAB_BASE, 0);
    double2 tb = s_log[b * LOG_TAB_REP + rep];
    double eps = fma(v, tb.x, -1.0);              // v / c_b - 1, |eps| <= 2^-9
    double a = fma(eps, KC[4], -0.25);
    a = fma(eps, a, KC[5]);
    a = fma(eps, a, -0.5);
    a = fma(eps, a, 1.0);
    return fma(eps, a, tb.y);
}

// t = ln(modified Fleming completeness) for one (walker, flux) pair -- the walker x source term.
//   n  = alpha*log10(f/F50) = fma(alpha, g, aF)            g = log10 f,  aF = -alpha*log10(F50)
//   fc = 1/2 (1 + n/sqrt(1+n^2))                           VmaxLumFunc.py:118-120
//   t  = ln(fc) / (1 - exp(-f/ftau))                       VmaxLumFunc.py:124-126, 141   (cinv = -1/ftau)
// MODIFIED=false: plain Fleming curve (fcmin falsy, VmaxLumFunc.py:121-122): t = ln(fc).
// FP64-pipe instruction count (MODIFIED): 9 + 6 + 9 + 3 = 27, +1 for the caller's accumulate.
template <bool MODIFIED>
__device__ __forceinline__ void fleming_log_parts(double g, double f, double alpha, double aF, double cinv,
                                                  const double* s_exp, const double2* s_log, int repe, int repl,
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
    lg = log_unit(fc, s_log, repl);
    if (MODIFIED) {
        double x = f * cinv;                                  // <= 0; |x| < 2e7 guaranteed by the classifier
        rdec = rcp_fast(one_minus_exp_neg(x, s_exp, repe));
    } else {
        rdec = 1.0;
    }
}

}  // namespace lfm
