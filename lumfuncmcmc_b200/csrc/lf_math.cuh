// lf_math.cuh -- FP64 building blocks of the walker x source loop, written for the sm_100a FP64 pipe.
//
// The hot loop is FP64-pipe bound (64 lanes/clk/SM; ncu: math-pipe-throttle is the dominant stall), so every
// routine here is counted in FP64-pipe instructions, and in how many of them read THREE distinct register operands
// (measured on B200: a DFMA with three register sources sustains 76 % of the rate of one with an immediate /
// constant-bank operand).  Integer / MUFU / LDS work rides in the issue slots the FP64 pipe leaves free.
//   * reciprocal square root and reciprocal: one MUFU seed (rsqrt/rcp.approx.ftz.f64, ~2^-20) + a correction folded into the
//     value that needs it (third-order for the rsqrt, one Newton step for the reciprocal), no IEEE div/sqrt sequences; the
//     rsqrt seed is built in its argument's register pair (rsqrt_seed_donor: no low-word zeroing instruction).
//   * exp: base-2 range reduction done by ONE fma against 1.5*2^44 (the product f*c2 is never rounded on its own),
//     256-entry table of 2^(j/256) in shared memory + minimax degree-3 polynomial on |r| <= 2^-9 (1.8e-14).
//   * log: one table lookup indexed by the top 20 bits of the argument (exponent AND 8 mantissa bits, covering
//     [2^-12, 1]) returning (1/c, ln c) with the exponent folded in, + minimax degree-3 log1p on |eps| <= 2^-9
//     (4.6e-13 absolute, zero mean; the constant term of the fit is folded into the table).  No exponent
//     extraction, no int->double conversion.
// Error budget (tools/math/fit_coeffs.py, tests/test_engine_gpu.py): the north-star tolerance is 1e-10 RELATIVE on
// lnprob, whose terms are O(10) each, so a per-term absolute error of 1e-12 leaves three orders of magnitude of
// margin; the polynomial degrees below are chosen for ~5e-13 per term instead of the 1e-16 of a libm-grade routine
// (22 FP64-pipe instructions per term instead of 28).  Anything outside the validated argument ranges goes to the
// literal kernels (see k_prologue in lf_engine.cu).
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
#ifndef LF_EXP_BIG
#define LF_EXP_BIG 1
#endif
constexpr int EXP_TAB_BITS = 8, EXP_TAB_N = 1 << EXP_TAB_BITS, EXP_TAB_REP = LF_EXP_REP;   // small table: 256*16*8 B = 32 KB, replicated x16 (a half-warp never bank-conflicts)
// LF_EXP_BIG: one unreplicated table of 2^(k/256) for k = EXPB_KMIN .. 0 (80 KB): the decay factor 2^(x2), x2 <= 0, is a
// single look-up with the integer k clamped at -40*256 (2^-40 = 9e-13 against 1, times |ln fc| < 0.01 for a source that
// bright) -- no index masking, no exponent arithmetic; the full-range exp of the FREE model's quadrature uses its top
// 256 entries.  The Z / FIXED models have no log table and keep the small replicated table in its place (their hot
// loop is the full-range exp, where bank conflicts of an unreplicated table cost 20 %).
constexpr int EXPB_KMIN = -40 * (1 << EXP_TAB_BITS), EXPB_N = -EXPB_KMIN + 1;
constexpr int EXP_SMEM_DOUBLES = LF_EXP_BIG ? ((EXPB_N + 1) & ~1) : EXP_TAB_N * EXP_TAB_REP;
constexpr bool EXP_BIG = LF_EXP_BIG != 0;
// Z / FIXED models (no log table; their hot loop IS the full-range 2^x): the replicated table in a layout that makes the
// look-up three integer instructions instead of six.  Row j is 256 B = 32 doubles, one column per lane (a half-warp never
// bank-conflicts), and holds 2^(j/256) with its HIGH WORD LOWERED by j << 12.  With k = 256 K + j, the low mantissa word of
// the magic-number sum:
//     byte offset of the entry = (j << 8) | (lane << 3) = PRMT(k, lane8)      (byte 0 of k moved into byte 1)
//     high word of 2^K T_j     = hi'(j) + (k << 12)     = one LEA             (j << 12 cancels the pre-compensation)
// instead of shift + mask-or for the address and shift + clamp + multiply-add for the exponent.
constexpr int EXPR_ROW_DOUBLES = 32, EXPR_SMEM_DOUBLES = EXP_TAB_N * EXPR_ROW_DOUBLES;   // 64 KB
constexpr int LOG_OCTAVES = 12, LOG_MANT_BITS = 8;
constexpr int LOG_TAB_N = LOG_OCTAVES * (1 << LOG_MANT_BITS) + 1, LOG_TAB_REP = LF_LOG_REP;  // 3073*2*16 B = 96 KB, two replicas (even / odd lanes)
constexpr int LOG_TAB_BASE = (1023 - LOG_OCTAVES) << LOG_MANT_BITS;                        // index of 2^-12 in (hi >> 12)
constexpr double LOG_ARG_MIN = 0x1p-12;

// Polynomial / reduction constants.  On sm_100a a constant reaches the FP64 pipe as an immediate (only when its low 32 bits
// are zero), through a uniform register, or from a vector register; ptxas keeps only some __constant__ values in uniform
// registers, and one it parks in a vector register turns "fma(r, p, C)" into a DFMA with three distinct register sources
// (76 % rate; 6 of 23 per term in math v4).  So the polynomials are arranged to need as few non-immediate constants as possible:
//   * 2^r = C0 (1 + c1 r + c2 r^2 + c3 r^3): C0 is folded into the table entries (EXP_TAB_SCALE, applied where the shared-memory
//     tables are filled / on the host for the big table), the last Horner step adds the immediate 1.0;
//   * log1p(eps) = c0 + eps + c2 eps^2 + c3 eps^3 with the linear coefficient FIXED at 1 in the minimax fit (same 4.55e-13
//     as the free fit, tools/math/fit_coeffs.py) and c0 folded into the ln c column of the log table;
//   * the cubic coefficients are rounded to their high words (immediates): moves the cubic term by <= 1.2e-7 relative, i.e.
//     <= 5e-17 in 2^r and 3e-16 in log1p on |r|, |eps| <= 2^-9.
// What is left (c1, c2 of 2^r, c2 of log1p, log2(e)) is written as literals: ptxas materialises a literal in a UNIFORM register
// (a pair of UMOVs, hoisted or ~0.4 per term) and feeds it to the DFMA as a UR operand, whereas values read from a
// __constant__ array ended up in vector registers (LDC.64 inside the loop) for half of the constants in math v4 and for all
// four of them once only four were left.  The magic number has a zero low word: an immediate as well.
constexpr double MAGIC44 = 26388279066624.0;                       // 1.5 * 2^44: ulp 2^-8, low mantissa bits = round(256 x)
// minimax fits on [-2^-9, 2^-9] (tools/math/fit_coeffs.py)
constexpr double LOG1P_C0 = 4.554593420756767562287e-13;           // folded into the ln c column of the log table
constexpr double LOG1P_C2 = -0.5000009544211419596383;
constexpr double LOG1P_C3_HI = 0x1.55557p-2;                       // 0.3333337151419489602819 rounded to its high word
constexpr double EXP2_C0 = 0.99999999999998250473;                 // = EXP_TAB_SCALE; the free fit is C0 (1, c1, c2, c3):
constexpr double EXP2N_C1 = 0.6931471805599477377071, EXP2N_C2 = 0.2402265436493579639682;
constexpr double EXP2N_C3_HI = 0x1.c6b09p-5;                       // 0.0555041162935780710385 rounded to its high word
constexpr double EXP_TAB_SCALE = EXP2_C0;
struct __align__(16) Tables {    // device-global master copies (filled by the host at lf_create)
    double exp2_frac[EXP_TAB_N];        // 2^(j/256)
    double exp2_big[EXPB_N + 1];        // EXP_TAB_SCALE * 2^(k/256), k = EXPB_KMIN .. 0
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

// The same seeds with the LOW WORD TAKEN FROM A DONOR instead of zeroed.  MUFU.RSQ64H / RCP64H write only the high word of
// the result; PTX semantics make the low word zero, which costs one IMAD.MOV per seed in the hot loop.  The corrections that
// follow only need a seed within ~2^-20 of the true value, so any low word will do: when the donor is a double that dies
// here, the seed is built in its register pair with no extra instruction.  Deterministic (the donor is a program value).
#ifndef LF_SEED_DONOR
#define LF_SEED_DONOR 1
#endif
__device__ __forceinline__ double rsqrt_seed_donor(double y, double donor) {
#if LF_SEED_DONOR
    double r;
    asm("{\n\t.reg .b32 dl, dh, sl, sh;\n\t.reg .f64 s;\n\tmov.b64 {dl, dh}, %2;\n\trsqrt.approx.ftz.f64 s, %1;\n\t"
        "mov.b64 {sl, sh}, s;\n\tmov.b64 %0, {dl, sh};\n\t}" : "=d"(r) : "d"(y), "d"(donor));
    return r;
#else
    return rsqrt_seed(y);
#endif
}
__device__ __forceinline__ double rcp_seed_donor(double y, double donor) {
#if LF_SEED_DONOR
    double r;
    asm("{\n\t.reg .b32 dl, dh, sl, sh;\n\t.reg .f64 s;\n\tmov.b64 {dl, dh}, %2;\n\trcp.approx.ftz.f64 s, %1;\n\t"
        "mov.b64 {sl, sh}, s;\n\tmov.b64 %0, {dl, sh};\n\t}" : "=d"(r) : "d"(y), "d"(donor));
    return r;
#else
    return rcp_seed(y);
#endif
}

// 1/d for normal positive d: MUFU seed (2^-22) + one Newton step (relative error ~6e-14): 2 FP64 instructions + 1 MUFU
__device__ __forceinline__ double rcp_fast(double d) {
    double r0 = rcp_seed(d);
    double e = fma(-d, r0, 1.0);
    return fma(r0, e, r0);
}

// cooperative fill of the (replicated) shared-memory tables
// models without a log table: only the small replicated exp table
__device__ __forceinline__ void load_exp_replicated(const Tables* __restrict__ t, double* s_exp_rep) {
    for (int i = threadIdx.x; i < EXPR_SMEM_DOUBLES; i += blockDim.x) {
        const int j = i / EXPR_ROW_DOUBLES;
        const double T = t->exp2_frac[j] * EXP_TAB_SCALE;
        s_exp_rep[i] = __hiloint2double(__double2hiint(T) - (j << 12), __double2loint(T));
    }
}
__device__ __forceinline__ void load_tables(const Tables* __restrict__ t, double* s_exp, double2* s_log) {
#if LF_EXP_BIG
    {   // 16-byte copies (the table has EXPB_N + 1 = an even number of doubles; both sides are 16-byte aligned)
        const double2* __restrict__ src = reinterpret_cast<const double2*>(t->exp2_big);
        double2* dst = reinterpret_cast<double2*>(s_exp);
        for (int i = threadIdx.x; i < (EXPB_N + 1) / 2; i += blockDim.x) dst[i] = src[i];
    }
#else
    for (int i = threadIdx.x; i < EXP_TAB_N * EXP_TAB_REP; i += blockDim.x) s_exp[i] = t->exp2_frac[i / EXP_TAB_REP] * EXP_TAB_SCALE;
#endif
    for (int i = threadIdx.x; i < LOG_TAB_N * LOG_TAB_REP; i += blockDim.x) s_log[i] = t->log_tab[i / LOG_TAB_REP];
}

// 2^(x2) split as 2^K * T[j] * p(r): returns p(r) ~ 2^r and the scaled table entry Ts = 2^K T[j];  x2 = a * b is
// formed inside the two fmas only (no separately rounded product).  Needs |x2| < 8.3e6.  FP64 instructions: 6.
template <bool BIG>
__device__ __forceinline__ void exp2_parts(double a, double b, const double* s_exp, int rep, int kmin, double& Ts, double& p) {
    double t = fma(a, b, MAGIC44);
    int k = __double2loint(t);                    // round(256 x2)
    double kf = t - MAGIC44;
    double r = fma(a, b, -kf);                    // |r| <= 2^-9, exact up to one rounding
    double T;
    int K = max(k >> EXP_TAB_BITS, kmin);         // keeps the exponent field valid
    if (BIG) {
        T = s_exp[(k & (EXP_TAB_N - 1)) + (EXPB_N - 1 - EXP_TAB_N)];         // 2^((j - 256)/256), j = k mod 256
        K += 1;
    } else {
        T = s_exp[(k & (EXP_TAB_N - 1)) * EXP_TAB_REP + rep];
    }
    Ts = __hiloint2double(__double2hiint(T) + (int)((unsigned)K << 20), __double2loint(T));
    p = fma(r, EXP2N_C3_HI, EXP2N_C2);
    p = fma(r, p, EXP2N_C1);
    p = fma(r, p, 1.0);                           // the table entry carries EXP_TAB_SCALE
}

// 1 - 2^(f * c2) for f * c2 in (-8.3e6, 0]; ABSOLUTE accuracy ~2e-14.  FP64 instructions: 7
__device__ __forceinline__ double one_minus_exp2(double f, double c2, const double* s_exp, int rep) {
    double Ts, p;
#if LF_EXP_BIG
    double t = fma(f, c2, MAGIC44);
    int k = max(__double2loint(t), EXPB_KMIN);    // round(256 x2), clamped: 2^-40 is 0 against 1 at the budget of this routine
    double kf = t - MAGIC44;
    double r = fma(f, c2, -kf);
    Ts = s_exp[k - EXPB_KMIN];
    p = fma(r, EXP2N_C3_HI, EXP2N_C2);
    p = fma(r, p, EXP2N_C1);
    p = fma(r, p, 1.0);                           // the table entry carries EXP_TAB_SCALE
#else
    exp2_parts<false>(f, c2, s_exp, rep, -1000, Ts, p);  // 2^x2 < 2^-1000 is 0 against 1
#endif
    return fma(-Ts, p, 1.0);
}

// 2^(x2) with RELATIVE accuracy ~2e-14 for x2 in [-1020, 1020] (the callers guarantee the range).  FP64 instructions: 7
template <bool BIG>
__device__ __forceinline__ double exp2_full(double x2, const double* s_exp, int rep) {
    double Ts, p;
    exp2_parts<BIG>(x2, 1.0, s_exp, rep, -1022, Ts, p);
    return Ts * p;
}

// exp(x) for x in [-707, 707]; below -707 returns 0.  FP64 instructions: 1 + 1 + 7
template <bool BIG>
__device__ __forceinline__ double exp_full(double x, const double* s_exp, int rep) {
    bool under = x < -707.0;
    double Ts, p;
    exp2_parts<BIG>(under ? -707.0 : x, LOG2E, s_exp, rep, -1022, Ts, p);
    return under ? 0.0 : Ts * p;
}

// The same split on the pre-compensated replicated table (Z / FIXED models): lane8 = 8 * lane, s_rep = the table.
// CLAMP: keep the exponent field valid below 2^-1022 (one more integer instruction); callers whose argument is known to
// stay inside (-1021, 1023) leave it off.  Same values as exp2_parts<false> bit for bit (same T_j, K, r and polynomial).
template <bool CLAMP>
__device__ __forceinline__ void exp2r_parts(double a, double b, const double* s_rep, unsigned lane8, double& Ts, double& p) {
    double t = fma(a, b, MAGIC44);
    int k = __double2loint(t);                    // round(256 x2)
    double kf = t - MAGIC44;
    double r = fma(a, b, -kf);
    if (CLAMP) k = max(k, -1022 * EXP_TAB_N);
    const unsigned off = __byte_perm((unsigned)k, lane8, 0x7704);           // (j << 8) | lane8
    const double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(s_rep) + off);
    Ts = __hiloint2double(__double2hiint(T) + (int)((unsigned)k << 12), __double2loint(T));
    p = fma(r, EXP2N_C3_HI, EXP2N_C2);
    p = fma(r, p, EXP2N_C1);
    p = fma(r, p, 1.0);                           // the table entry carries EXP_TAB_SCALE
}
template <bool CLAMP>
__device__ __forceinline__ double exp2r_full(double x2, const double* s_rep, unsigned lane8) {
    double Ts, p;
    exp2r_parts<CLAMP>(x2, 1.0, s_rep, lane8, Ts, p);
    return Ts * p;
}
// exp(x) for x <= 707; below -707 returns 0
__device__ __forceinline__ double expr_full(double x, const double* s_rep, unsigned lane8) {
    bool under = x < -707.0;
    double Ts, p;
    exp2r_parts<false>(under ? -707.0 : x, LOG2E, s_rep, lane8, Ts, p);
    return under ? 0.0 : Ts * p;
}

// log(v) for v in [2^-12, 1], absolute accuracy 4.6e-13 (zero-mean).  FP64 instructions: 4
__device__ __forceinline__ double log_unit(double v, const double2* s_log, int rep) {
    int b = max((__double2hiint(v) >> (20 - LOG_MANT_BITS)) - LOG_TAB_BASE, 0);
    double2 tb = s_log[b * LOG_TAB_REP + rep];
    double eps = fma(v, tb.x, -1.0);              // v / c_b - 1, |eps| <= 2^-9
    double a = fma(eps, LOG1P_C3_HI, LOG1P_C2);
    a = fma(eps, a, 1.0);
    return fma(eps, a, tb.y);                     // tb.y = ln c_b + LOG1P_C0
}

// t = ln(modified Fleming completeness) for one (walker, flux) pair -- the walker x source term.
//   n  = alpha*log10(f/F50) = fma(alpha, g, aF)            g = log10 f,  aF = -alpha*log10(F50)
//   fc = 1/2 (1 + n/sqrt(1+n^2))                           VmaxLumFunc.py:118-120
//   t  = ln(fc) / (1 - exp(-f/ftau))                       VmaxLumFunc.py:124-126, 141   (c2 = -log2(e)/ftau)
// MODIFIED=false: plain Fleming curve (fcmin falsy, VmaxLumFunc.py:121-122): t = ln(fc).
// FP64-pipe instruction count (MODIFIED): 8 + 4 + 7 + 2 = 21, +1 for the caller's accumulate.
template <bool MODIFIED>
__device__ __forceinline__ void fleming_log_parts(double g, double f, double alpha, double aF, double c2,
                                                  const double* s_exp, const double2* s_log, int repe, int repl,
                                                  double& lg, double& rdec) {
    double n = fma(alpha, g, aF);
    double y = fma(n, n, 1.0);
    double r0 = rsqrt_seed_donor(y, y);           // y dies here
    double nr = n * r0;
    double e = fma(-nr, nr, fma(-r0, r0, 1.0));   // 1 - y r0^2 with y = n^2 + 1
    double q = fma(fma(0.1875, e, 0.25), e, 0.5); // 1/2 (1 + e/2 + 3 e^2/8)
    double fc = fma(nr, q, 0.5);
    lg = log_unit(fc, s_log, repl);
    if (MODIFIED) {
        rdec = rcp_fast(one_minus_exp2(f, c2, s_exp, repe));   // f * c2 <= 0, > -8e6 guaranteed by the classifier
    } else {
        rdec = 1.0;
    }
}

// NT independent (walker, source) terms evaluated in lock-step and added to acc[0..NT-1].  The source is written phase
// by phase so that the NT independent FP64 chains sit next to each other in the instruction stream: on B200 the FP64
// pipe issues one warp instruction per 2 cycles only while consecutive DFMAs come from the SAME warp (3 cycles when the
// scheduler has to alternate between warps, tools/microbench/fp64_mix.cu), so a warp must offer >= 4 independent DFMAs
// at every point of the chain (DFMA latency 8 cycles).
// Chain i evaluates the source (ux[i], uy[i]) for the walker constants (al[i], af[i], cc[i]); the arrays are register
// names, so chains that share a source or a walker cost no extra registers.
template <int NT>
__device__ __forceinline__ void fleming_terms_v(const double (&ux)[NT], const double (&uy)[NT], const double (&al)[NT],
                                                const double (&af)[NT], const double (&cc)[NT], const double* s_exp,
                                                const double2* s_log, int repe, int repl, double* acc) {
    double n[NT], y[NT], r0[NT], e[NT], q[NT], fc[NT], lg[NT], dec[NT];
    double t[NT], r[NT], Ts[NT], p[NT];
    double2 tb[NT];
    int k[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) n[i] = fma(al[i], ux[i], af[i]);
#pragma unroll
    for (int i = 0; i < NT; ++i) t[i] = fma(uy[i], cc[i], MAGIC44);
#pragma unroll
    for (int i = 0; i < NT; ++i) y[i] = fma(n[i], n[i], 1.0);
#pragma unroll
    for (int i = 0; i < NT; ++i) r0[i] = rsqrt_seed_donor(y[i], y[i]);    // y dies here: e = 1 - r0^2 - (n r0)^2 below
    // exp branch while the MUFUs are in flight
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        k[i] = __double2loint(t[i]);
        t[i] = t[i] - MAGIC44;
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
#if LF_EXP_BIG
        Ts[i] = s_exp[max(k[i], EXPB_KMIN) - EXPB_KMIN];
#else
        double T = s_exp[(k[i] & (EXP_TAB_N - 1)) * EXP_TAB_REP + repe];
        int K = max(k[i] >> EXP_TAB_BITS, -1000);
        Ts[i] = __hiloint2double(__double2hiint(T) + (int)((unsigned)K << 20), __double2loint(T));
#endif
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) r[i] = fma(uy[i], cc[i], -t[i]);
#pragma unroll
    for (int i = 0; i < NT; ++i) p[i] = fma(r[i], EXP2N_C3_HI, EXP2N_C2);
    // rsqrt correction: e = 1 - y r0^2 with y = n^2 + 1 written as 1 - r0^2 - (n r0)^2 (same three instructions, y not needed)
#pragma unroll
    for (int i = 0; i < NT; ++i) n[i] = n[i] * r0[i];
#pragma unroll
    for (int i = 0; i < NT; ++i) e[i] = fma(-r0[i], r0[i], 1.0);
#pragma unroll
    for (int i = 0; i < NT; ++i) e[i] = fma(-n[i], n[i], e[i]);
#pragma unroll
    // fc = 1/2 + 1/2 n rsqrt(y) with rsqrt(y) = r0 (1 + e/2 + 3 e^2/8): = 1/2 + (n r0) (1/2 + e/4 + 3 e^2/16), three instructions
    for (int i = 0; i < NT; ++i) y[i] = fma(0.1875, e[i], 0.25);
#pragma unroll
    for (int i = 0; i < NT; ++i) p[i] = fma(r[i], p[i], EXP2N_C1);
#pragma unroll
    for (int i = 0; i < NT; ++i) q[i] = fma(y[i], e[i], 0.5);
#pragma unroll
    for (int i = 0; i < NT; ++i) p[i] = fma(r[i], p[i], 1.0);
#pragma unroll
    for (int i = 0; i < NT; ++i) fc[i] = fma(n[i], q[i], 0.5);
#pragma unroll
    for (int i = 0; i < NT; ++i) dec[i] = fma(-Ts[i], p[i], 1.0);
#pragma unroll
    // (a donor low word here would save one more IMAD.MOV per term.  An arbitrary donor costs accuracy -- the single Newton
    //  step below squares the seed error: per-term bound 1.6e-12 -> 4.3e-12, tools/math/term_accuracy.cpp.  kf = k / 256 as
    //  the donor, whose low word is zero for |k| < 2^20, is bit-identical and was measured: 9.3 instead of 10.5 non-FP64
    //  instructions per term, and 1.7 % SLOWER in the product kernel (6.43 vs 6.54e11 terms/s) -- not taken.)
    for (int i = 0; i < NT; ++i) r0[i] = rcp_seed(dec[i]);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        // byte offset of table row b = (hi >> 12) - LOG_TAB_BASE: ((hi >> 12) * 32) = (hi >> 7) & ~31, the base folds into
        // the load's immediate, the replica offset is OR-ed in (LOP3 co-issues with the FP64 pipe, IMAD does not)
        static_assert(LOG_TAB_REP == 2 && LOG_MANT_BITS == 8, "index arithmetic below assumes 32-byte rows");
        const unsigned off = ((unsigned)(__double2hiint(fc[i]) >> 7) & ~31u) | (unsigned)(repl << 4);
        tb[i] = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(s_log) + off - (LOG_TAB_BASE << 5));
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) e[i] = fma(-dec[i], r0[i], 1.0);
#pragma unroll
    for (int i = 0; i < NT; ++i) fc[i] = fma(fc[i], tb[i].x, -1.0);          // eps
#pragma unroll
    for (int i = 0; i < NT; ++i) r0[i] = fma(r0[i], e[i], r0[i]);           // 1 / dec
#pragma unroll
    for (int i = 0; i < NT; ++i) lg[i] = fma(fc[i], LOG1P_C3_HI, LOG1P_C2);
#pragma unroll
    for (int i = 0; i < NT; ++i) lg[i] = fma(fc[i], lg[i], 1.0);
#pragma unroll
    for (int i = 0; i < NT; ++i) lg[i] = fma(fc[i], lg[i], tb[i].y);
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i] = fma(lg[i], r0[i], acc[i]);
}

// NT sources of one walker
template <int NT>
__device__ __forceinline__ void fleming_terms(const double2* u, double alpha, double aF, double c2, const double* s_exp,
                                              const double2* s_log, int repe, int repl, double* acc) {
    double ux[NT], uy[NT], al[NT], af[NT], cc[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { ux[i] = u[i].x; uy[i] = u[i].y; al[i] = alpha; af[i] = aF; cc[i] = c2; }
    fleming_terms_v<NT>(ux, uy, al, af, cc, s_exp, s_log, repe, repl, acc);
}

// ---- libm-grade (~2e-16) routines for the streaming 1/V_eff kernel: its per-source weights are compared at 1e-13 and
// ---- its shared memory holds histograms, so these use two small UNREPLICATED tables (6 KB): the [1/2, 1) octave of
// ---- the log table and the 2^(j/256) table, with the exponent handled arithmetically and Taylor polynomials ----
constexpr int STREAM_LOG_N = 1 << LOG_MANT_BITS;
constexpr double LN2_HI = 0.693147180369123816490, LN2_LO = 1.90821492927058770002e-10;
// coefficients in the constant bank (c[3][..] operands: no UMOV pairs to materialise 64-bit immediates)
static __constant__ double KS[16] = {
    0.2, 1.0 / 3.0,                                  // 0-1   log1p (the other coefficients are immediates)
    4503599627370496.0 + 1022.0,                     // 2     2^52 + exponent bias of the [1/2, 1) mantissa
    LN2_HI, LN2_LO,                                  // 3-4
    256.0 * LOG2E, MAGIC52, -LN2_HI / 256.0, -LN2_LO / 256.0,   // 5-8   exp range reduction
    1.0 / 24.0, 1.0 / 6.0,                           // 9-10  expm1
    0.0,
    0.43429448190325182765,                          // 12    log10(e)
    -256.0 * LOG2E,                                  // 13    exp(-x) range reduction
    0.0, 0.0};

__device__ __forceinline__ void load_stream_tables(const Tables* __restrict__ t, double* s_exp, double2* s_logm) {
    // 2^(j/256) with the high word lowered by j << 12: hi(2^K T_j) = hi'(j) + (k << 12) for k = 256 K + j -- one multiply-add,
    // no masking of j out of the shifted k (exp_stream / exp_stream_signed)
    for (int i = threadIdx.x; i < EXP_TAB_N; i += blockDim.x) {
        const double T = t->exp2_frac[i];
        s_exp[i] = __hiloint2double(__double2hiint(T) - (i << 12), __double2loint(T));
    }
    for (int i = threadIdx.x; i < STREAM_LOG_N; i += blockDim.x) {
        double2 e = t->log_tab[(LOG_OCTAVES - 1) * STREAM_LOG_N + i];     // octave [1/2, 1): (1/c, ln c + LOG1P_C0)
        s_logm[i] = make_double2(e.x, e.y - LOG1P_C0);
    }
}

// ln(v) for positive normal v: |error| <= 2.1e-16 max(|ln v|, 1) (tools/math/stream_accuracy.cpp).  10 FP64-pipe
// instructions; the exponent becomes a double by pairing it with the high word of 2^52 (no I2F)
__device__ __forceinline__ double log_stream(double v, const double2* s_logm) {
    const int hi = __double2hiint(v);
    int mh;                                                                 // (hi & 0xfffff) | 0x3fe00000 as ONE LOP3 (constants in registers)
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(mh) : "r"(hi), "r"(0x000fffff), "r"(0x3fe00000));
    const double m = __hiloint2double(mh, __double2loint(v));               // v = 2^e m, m in [1/2, 1)
    const double2 tb = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(s_logm) +
                                                         (((unsigned)hi >> (20 - LOG_MANT_BITS - 4)) & ((STREAM_LOG_N - 1) << 4)));
    const double eps = fma(m, tb.x, -1.0);                                  // |eps| <= 2^-9
    double p = fma(eps, KS[0], -0.25);
    p = fma(eps, p, KS[1]);
    p = fma(eps, p, -0.5);
    p = fma(eps * eps, p, eps);                                             // log1p(eps), truncation eps^6/6 < 1e-17
    const double ef = __hiloint2double(0x43300000, hi >> 20) - KS[2];       // e = (hi >> 20) - 1022
    return fma(ef, KS[3], fma(ef, KS[4], tb.y + p));
}

// exp(x) for x in [-700, 700] (callers guarantee the range): relative error <= 2.4e-16.  9 FP64-pipe instructions
__device__ __forceinline__ double exp_stream(double x, const double* s_exp) {
    const double t = fma(x, KS[5], KS[6]);
    const int k = __double2loint(t);
    const double kf = t - KS[6];
    double r = fma(kf, KS[7], x);
    r = fma(kf, KS[8], r);                                                  // |r| <= ln2/512
    const double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(s_exp) + (((unsigned)k << 3) & ((EXP_TAB_N - 1) << 3)));
    const double Ts = __hiloint2double(__double2hiint(T) + (int)((unsigned)k << (20 - EXP_TAB_BITS)), __double2loint(T));   // pre-compensated table
    double p = fma(r, KS[9], KS[10]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = p * r;                                                              // expm1(r), truncation r^5/120 < 4e-17
    return fma(Ts, p, Ts);
}

// exp(+x) (NEG = false) or exp(-x) (NEG = true) with the sign folded into the constants (no negation instruction) and the
// power of two clamped at 2^-1000 instead of the argument: for arguments below -693 the result is a finite ~1e-301 (0
// against 1 for the decay factor 1 - exp(-x), the only caller that can get there), so no compare-and-select is needed.
// Same arithmetic and accuracy as exp_stream on its range.
// CLAMP = false: the caller guarantees |x| < 700 or discards the result (the table index is masked either way).
template <bool NEG, bool CLAMP = true>
__device__ __forceinline__ double exp_stream_signed(double x, const double* s_exp) {
    const double t = fma(x, NEG ? KS[13] : KS[5], KS[6]);
    const int k = CLAMP ? max(__double2loint(t), -1000 * EXP_TAB_N) : __double2loint(t);
    const double kf = t - KS[6];
    double r = NEG ? fma(kf, KS[7], -x) : fma(kf, KS[7], x);
    r = fma(kf, KS[8], r);                                                  // |r| <= ln2/512
    const double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(s_exp) + (((unsigned)k << 3) & ((EXP_TAB_N - 1) << 3)));
    const double Ts = __hiloint2double(__double2hiint(T) + (int)((unsigned)k << (20 - EXP_TAB_BITS)), __double2loint(T));   // pre-compensated table
    double p = fma(r, KS[9], KS[10]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = p * r;
    return fma(Ts, p, Ts);
}

// 1/d to ~1e-16 for normal positive d: MUFU seed + two Newton steps
__device__ __forceinline__ double rcp_stream(double d) {
    double r = rcp_seed(d);
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

}  // namespace lfm
