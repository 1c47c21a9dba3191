// Host replica of one walker x source term of the free-completeness kernel (lf_math.cuh: fleming_log_parts<true> with
// log_unit, one_minus_exp2 and rcp_fast -- the same operations in the same order, fma for fma) against long double.
// The two MUFU seeds are emulated as the exact value times (1 + delta) with delta drawn uniformly inside the error
// bounds measured on B200 (tools/microbench/mufu64_accuracy.cu: |1 - y r0^2| <= 1.86e-6, |1 - d r0| <= 9.9e-7), so
// the maxima printed here hold for any seed inside those bounds; the rsqrt seed's low word is a donor's
// (rsqrt_seed_donor: no zeroing instruction), which adds up to 2^-20, one-sided.
//     g++ -O2 -o term_accuracy term_accuracy.cpp && ./term_accuracy [samples]
// tests/test_math_replica.py runs it with 2e6 samples.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
static const double MAGIC44 = 26388279066624.0, LOG2E = 1.442695040888963407359924681001892137;
// as lf_math.cuh ("math v6"): log1p fit with the linear coefficient fixed at 1, 2^r normalised by its constant term (which the
// table entries carry), cubic coefficients rounded to their high words (DFMA immediates)
static const double LOG1P_C0 = 4.554593420756767562287e-13, LOG1P_C2 = -0.5000009544211419596383, LOG1P_C3_HI = 0x1.55557p-2;
static const double EXP2_C0 = 0.99999999999998250473, EXP2N_C1 = 0.6931471805599477377071, EXP2N_C2 = 0.2402265436493579639682,
                    EXP2N_C3_HI = 0x1.c6b09p-5;
static const int LOG_OCTAVES = 12, M = 256, LOG_TAB_BASE = (1023 - LOG_OCTAVES) << 8, EXPB_KMIN = -40 * 256;
static std::vector<double> tab_invc, tab_lnc, exp_big;
static inline int hi(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u >> 32); }
static inline int lo(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u & 0xffffffffu); }
static std::mt19937_64 rng(7);
static double unit() { return std::uniform_real_distribution<double>(-1.0, 1.0)(rng); }

static double term_fast(double g, double f, double alpha, double aF, double c2) {
    const double n = fma(alpha, g, aF);
    const double y = fma(n, n, 1.0);
    // rsqrt.approx.ftz.f64; its low word is a donor's (rsqrt_seed_donor): up to 2^-20 more, one-sided
    const double r0 = (double)(1.0L / sqrtl((long double)y)) * (1.0 + 0.93e-6 * unit()) * (1.0 + 9.54e-7 * fabs(unit()));
    const double nr = n * r0;                        // as lf_math.cuh: e = 1 - r0^2 - (n r0)^2, fc = 1/2 + nr (1/2 + e/4 + 3 e^2/16)
    const double e = fma(-nr, nr, fma(-r0, r0, 1.0));
    const double q = fma(fma(0.1875, e, 0.25), e, 0.5);
    const double fc = fma(nr, q, 0.5);
    // log_unit
    int b = (hi(fc) >> 12) - LOG_TAB_BASE;
    if (b < 0) b = 0;
    const double eps = fma(fc, tab_invc[b], -1.0);
    double a = fma(eps, LOG1P_C3_HI, LOG1P_C2);
    a = fma(eps, a, 1.0);
    const double lg = fma(eps, a, tab_lnc[b]);
    // one_minus_exp2 (one look-up in the big table, k clamped at -40 * 256 entries)
    const double t = fma(f, c2, MAGIC44);
    int k = lo(t);
    if (k < EXPB_KMIN) k = EXPB_KMIN;
    const double kf = t - MAGIC44;
    const double r = fma(f, c2, -kf);
    const double Ts = exp_big[k - EXPB_KMIN];        // EXP2_C0 * 2^(k / 256)
    double pp = fma(r, EXP2N_C3_HI, EXP2N_C2);
    pp = fma(r, pp, EXP2N_C1);
    pp = fma(r, pp, 1.0);
    const double dec = fma(-Ts, pp, 1.0);
    // rcp_fast
    const double s0 = (double)(1.0L / (long double)dec) * (1.0 + 9.9e-7 * unit());         // rcp.approx.ftz.f64 (low word zero)
    const double ee = fma(-dec, s0, 1.0);
    const double rdec = fma(s0, ee, s0);
    return lg * rdec;
}

int main(int argc, char** argv) {
    const long nsamp = argc > 1 ? atol(argv[1]) : 20000000L;
    tab_invc.resize(LOG_OCTAVES * M + 2); tab_lnc.resize(LOG_OCTAVES * M + 2); exp_big.resize(-EXPB_KMIN + 2);
    for (int i = 0; i <= -EXPB_KMIN; ++i) exp_big[i] = (double)((long double)EXP2_C0 * exp2l((long double)(EXPB_KMIN + i) / 256));
    for (int b = 0; b < LOG_OCTAVES * M; ++b) {
        const int E = -LOG_OCTAVES + b / M, j = b % M;
        const long double cm = 1.0L + ((long double)j + 0.5L) / M;
        const double invc = ldexp((double)(1.0L / cm), -E);
        tab_invc[b] = invc;
        tab_lnc[b] = (double)(-logl((long double)invc) + (long double)LOG1P_C0);
    }
    tab_invc[LOG_OCTAVES * M] = 1.0; tab_lnc[LOG_OCTAVES * M] = 0.0;
    std::uniform_real_distribution<double> ualpha(1.0, 7.0), ulogr(-1.3, 4.0), uF(-17.0, -16.0);
    double worst_abs = 0.0, worst_rel = 0.0, worst_t = 0.0, worst_typ = 0.0;
    for (long i = 0; i < nsamp; ++i) {
        const double alpha = ualpha(rng), logF50 = uF(rng);
        const double g = logF50 + ulogr(rng);                       // log10 f: f / F50 in [0.05, 1e4]
        const double f = pow(10.0, g);
        const double aF = -alpha * logF50;
        const double ftau = pow(10.0, logF50 - sqrt(0.64 / 0.36) / alpha);   // fcmin = 0.1 (VmaxLumFunc.py:164-167)
        const double c2 = -LOG2E / ftau;
        const long double n = (long double)alpha * g + aF;
        if (n < -30.0L) continue;                                   // fc below the log table: literal class in the engine
        const long double fc = 0.5L * (1.0L + n / sqrtl(1.0L + n * n));
        const long double ref = logl(fc) / (1.0L - exp2l((long double)f * c2));
        const double got = term_fast(g, f, alpha, aF, c2);
        const double ea = fabs((double)(got - ref)), er = ea / fmax(fabs((double)ref), 1.0);
        if (ea > worst_abs) { worst_abs = ea; worst_t = (double)ref; }
        if (er > worst_rel) worst_rel = er;
        // well-conditioned terms: fc = (1 + n / sqrt(1 + n^2)) / 2 carries eps / fc in ANY double evaluation (NumPy's too)
        if (fc > 0.01L && (double)f / ftau > 0.1 && ea > worst_typ) worst_typ = ea;
    }
    printf("term: max abs err %.3e (at t = %.3f), max err / max(|t|, 1) %.3e, max abs err for fc > 0.01 and f / ftau > 0.1: %.3e\n",
           worst_abs, worst_t, worst_rel, worst_typ);
    return 0;
}
