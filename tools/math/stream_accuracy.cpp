// Host replica of the streaming log / exp routines (lf_math.cuh: log_stream, exp_stream -- the same operations in the
// same order, fma for fma) to measure their error against long double.
//     g++ -O2 -o stream_accuracy stream_accuracy.cpp && ./stream_accuracy [samples]
// Prints the three maxima on one line; tests/test_math_replica.py runs it with 2e6 samples.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <random>
static double tabx[256], taby[256], e2[256];
static const double LN2_HI = 0.693147180369123816490, LN2_LO = 1.90821492927058770002e-10, LOG2E = 1.4426950408889634074;
static const double MAGIC52 = 6755399441055744.0;
static inline int hi(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u >> 32); }
static inline int lo(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u & 0xffffffffu); }
static inline double mk(int h, int l) { uint64_t u = ((uint64_t)(uint32_t)h << 32) | (uint32_t)l; double v; memcpy(&v, &u, 8); return v; }
static double log_stream(double v) {
    const int h = hi(v);
    const double m = mk((h & 0x000fffff) | 0x3fe00000, lo(v));
    const int idx = (h >> 12) & 255;
    const double eps = fma(m, tabx[idx], -1.0);
    double p = fma(eps, 0.2, -0.25);
    p = fma(eps, p, 1.0 / 3.0);
    p = fma(eps, p, -0.5);
    p = fma(eps * eps, p, eps);
    const double ef = mk(0x43300000, h >> 20) - (4503599627370496.0 + 1022.0);
    return fma(ef, LN2_HI, fma(ef, LN2_LO, taby[idx] + p));
}
static double exp_stream(double x) {
    const double t = fma(x, 256.0 * LOG2E, MAGIC52);
    const int k = lo(t);
    const double kf = t - MAGIC52;
    double r = fma(kf, -LN2_HI / 256.0, x);
    r = fma(kf, -LN2_LO / 256.0, r);
    const double T = e2[k & 255];
    const double Ts = mk(hi(T) + ((k >> 8) << 20), lo(T));
    double p = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = p * r;
    return fma(Ts, p, Ts);
}
int main(int argc, char** argv) {
    const long nsamp = argc > 1 ? atol(argv[1]) : 20000000L;
    for (int j = 0; j < 256; ++j) {
        e2[j] = (double)exp2l((long double)j / 256);
        long double cm = 1.0L + ((long double)j + 0.5L) / 256;
        double invc = ldexp((double)(1.0L / cm), 1);
        tabx[j] = invc; taby[j] = (double)(-logl((long double)invc));
    }
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> ue(-60.0, 20.0), ux(-700.0, 700.0);
    double worst_abs = 0, worst_rel = 0, worst_e = 0;
    for (long i = 0; i < nsamp; ++i) {
        double v = exp2(ue(g));
        long double ref = logl((long double)v);
        double got = log_stream(v);
        double ea = fabs((double)(got - ref));
        double er = ea / fmax(fabs((double)ref), 1.0);
        if (ea > worst_abs) worst_abs = ea;
        if (er > worst_rel) worst_rel = er;
        double x = ux(g);
        long double re = expl((long double)x);
        double ge = exp_stream(x);
        double ee = fabs((double)((ge - re) / re));
        if (ee > worst_e) worst_e = ee;
    }
    printf("log: max abs err %.3e, max err / max(|ln|, 1) %.3e; exp: max rel err %.3e\n", worst_abs, worst_rel, worst_e);
    return 0;
}
