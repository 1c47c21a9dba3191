// Is the walker x source loop of k_main<false, FREE> held back by shared-memory bank conflicts, or by issue slots?
//
// The loop body of the product kernel (fleming_terms_v<4> from lf_math.cuh: two sources x two walkers per lane, 22 FP64 +
// ~10.7 other instructions per term) is run here on an L1-resident block of synthetic sources, 12 warps per SM with the
// product's 217 KB of tables, in two set-ups that execute the IDENTICAL instruction stream:
//   distinct : every lane carries its own walker constants, as in a real ensemble -> per-lane table indices -> the bank
//              conflicts ncu reports for the product kernel (39 % of the shared wavefronts)
//   uniform  : all lanes of a warp carry the same constants -> every table look-up is a broadcast, ZERO bank conflicts
// plus a register-only DFMA loop for the peak.  If "uniform" is not faster than "distinct", the conflicts cost nothing and
// the loop sits at the issue-slot bound of its instruction mix: cycles per term per scheduler = 2 x N_fp64 (the FP64 issue
// path takes a warp instruction every 2 cycles) + ~1 per other instruction that cannot hide in a DFMA's second cycle.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lumfuncmcmc_b200/csrc -o term_issue_bound tools/microbench/term_issue_bound.cu
#include <cmath>
#include <cstdio>
#include <vector>

#include "lf_math.cuh"
using namespace lfm;

static void fill_tables(Tables& t) {          // as lf_engine.cu
    for (int j = 0; j < EXP_TAB_N; ++j) t.exp2_frac[j] = (double)exp2l((long double)j / EXP_TAB_N);
    for (int i = 0; i < EXPB_N; ++i) t.exp2_big[i] = (double)((long double)EXP_TAB_SCALE * exp2l((long double)(EXPB_KMIN + i) / EXP_TAB_N));
    t.exp2_big[EXPB_N] = EXP_TAB_SCALE;
    const int M = 1 << LOG_MANT_BITS;
    for (int b = 0; b < LOG_OCTAVES * M; ++b) {
        int E = -LOG_OCTAVES + b / M, j = b % M;
        long double cm = 1.0L + ((long double)j + 0.5L) / M;
        double invc = ldexp((double)(1.0L / cm), -E);
        t.log_tab[b].x = invc;
        t.log_tab[b].y = (double)(-logl((long double)invc) + (long double)LOG1P_C0);
    }
    t.log_tab[LOG_OCTAVES * M].x = 1.0; t.log_tab[LOG_OCTAVES * M].y = 0.0;
    t.log_tab[LOG_OCTAVES * M + 1] = t.log_tab[LOG_OCTAVES * M];
}

constexpr int WARPS = 12;
constexpr size_t SMEM = sizeof(double2) * LOG_TAB_N * LOG_TAB_REP + sizeof(double) * EXP_SMEM_DOUBLES;

// wc[lane-slot][0..5] = alphaA, aFA, c2A, alphaB, aFB, c2B
__global__ void __launch_bounds__(32 * WARPS, 1) k_terms(const Tables* __restrict__ tables, const double2* __restrict__ src, int nsrc,
                                                        int reps, const double* __restrict__ wc, int uniform, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* s_log = reinterpret_cast<double2*>(smem);
    double* s_exp = reinterpret_cast<double*>(s_log + LOG_TAB_N * LOG_TAB_REP);
    load_tables(tables, s_exp, s_log);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int rep16 = lane & (EXP_TAB_REP - 1), rep8 = lane & (LOG_TAB_REP - 1);
    const double* w = wc + 6 * (uniform ? 0 : (blockIdx.x * blockDim.x + threadIdx.x) % 1024);
    const double al[4] = {w[0], w[3], w[0], w[3]}, af[4] = {w[1], w[4], w[1], w[4]}, cc[4] = {w[2], w[5], w[2], w[5]};
    double accv[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < reps; ++r) {
        const double2* __restrict__ ps = src + (threadIdx.x >> 5) * 64;               // warp-uniform address: broadcast loads
        double2 A0, A1, A2, A3, B0, B1, B2, B3;
        auto two = [&](const double2& s0, const double2& s1) {
            const double ux[4] = {s0.x, s0.x, s1.x, s1.x}, uy[4] = {s0.y, s0.y, s1.y, s1.y};
            fleming_terms_v<4>(ux, uy, al, af, cc, s_exp, s_log, rep16, rep8, accv);
        };
        const int cnt = nsrc - 64;
        int j = 0;
        A0 = __ldg(ps); A1 = __ldg(ps + 1); A2 = __ldg(ps + 2); A3 = __ldg(ps + 3);
        for (; j + 8 <= cnt; j += 8) {                                              // the product loop, verbatim
            B0 = __ldg(ps + j + 4); B1 = __ldg(ps + j + 5); B2 = __ldg(ps + j + 6); B3 = __ldg(ps + j + 7);
            two(A0, A1); two(A2, A3);
            if (j + 12 <= cnt) { A0 = __ldg(ps + j + 8); A1 = __ldg(ps + j + 9); A2 = __ldg(ps + j + 10); A3 = __ldg(ps + j + 11); }
            two(B0, B1); two(B2, B3);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (accv[0] + accv[2]) + (accv[1] + accv[3]);
}

__global__ void __launch_bounds__(256) k_dfma(int iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1.0e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) sink[0] = s;
}

int main() {
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    Tables* h_t = new Tables;
    fill_tables(*h_t);
    Tables* d_t;
    cudaMalloc(&d_t, sizeof(Tables));
    cudaMemcpy(d_t, h_t, sizeof(Tables), cudaMemcpyHostToDevice);
    // synthetic catalogue block: log10 flux in [-16.9, -14.5], flux = 10^g; walker constants of a converged ensemble
    const int nsrc = 4096 + 64;
    std::vector<double2> src(nsrc + 1024);                     // warps start at offsets of up to 11 * 64 sources
    unsigned long long lcg = 12345;
    auto rnd = [&]() { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(lcg >> 11) / 9007199254740992.0; };
    for (auto& s : src) { double g = -16.9 + 2.4 * rnd(); s = make_double2(g, fmin(pow(10.0, g), 3.84e-15)); }
    std::vector<double> wc(6 * 1024);
    for (int i = 0; i < 1024; ++i)
        for (int h = 0; h < 2; ++h) {
            double alpha = 3.5 + 0.2 * (rnd() - 0.5), F50 = (2.7 + 0.3 * (rnd() - 0.5)) * 1e-17;
            double ftau = F50 * pow(10.0, -sqrt(fabs(0.64 / 0.36)) / alpha);
            wc[6 * i + 3 * h] = alpha; wc[6 * i + 3 * h + 1] = -alpha * log10(F50); wc[6 * i + 3 * h + 2] = -LOG2E / ftau;
        }
    double2* d_src; double *d_wc, *d_out, *d_sink;
    cudaMalloc(&d_src, sizeof(double2) * src.size());
    cudaMalloc(&d_wc, sizeof(double) * wc.size());
    cudaMalloc(&d_out, sizeof(double) * sms * 32 * WARPS);
    cudaMalloc(&d_sink, 8);
    cudaMemcpy(d_src, src.data(), sizeof(double2) * src.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_wc, wc.data(), sizeof(double) * wc.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_terms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // DFMA peak
    k_dfma<<<sms * 8, 256>>>(64, 1.0, d_sink);
    cudaEventRecord(e0);
    k_dfma<<<sms * 8, 256>>>(100000, 1.0, d_sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_s = (double)sms * 8 * 256 * 100000.0 * 64.0 / (ms * 1e-3);
    const double cyc_per_dfma = (double)sms * 4 * 32 * (khz * 1e3) / dfma_per_s;      // issue cycles per warp-DFMA per scheduler
    printf("DFMA peak: %.3e thread-instr/s = %.2f cycles per warp instruction per scheduler at %.0f MHz\n", dfma_per_s, cyc_per_dfma, khz / 1e3);
    const int reps = 400;
    const double terms_per_warp = (double)reps * ((nsrc - 64) / 8 * 8) * 2.0;          // two walkers per lane
    for (int uniform = 0; uniform < 2; ++uniform) {
        k_terms<<<sms, 32 * WARPS, SMEM>>>(d_t, d_src, nsrc, 4, d_wc, uniform, d_out);
        cudaDeviceSynchronize();
        float best = 1e9f;
        for (int t = 0; t < 3; ++t) {
            cudaEventRecord(e0);
            k_terms<<<sms, 32 * WARPS, SMEM>>>(d_t, d_src, nsrc, reps, d_wc, uniform, d_out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            best = fminf(best, ms);
        }
        const double terms_per_s = terms_per_warp * 32.0 * WARPS * sms / (best * 1e-3);
        const double cyc_per_term = (double)sms * 4 * 32 * (khz * 1e3) / terms_per_s;  // per warp-term per scheduler
        const double NF = 22.0, NO = 10.7;                                             // per term, from the SASS of this loop
        printf("%-8s walker constants: %.3e terms/s  = %.1f cycles per term per scheduler; %.0f FP64 at the measured %.2f = %.1f -> "
               "%.3f of the DFMA peak; the other %.1f cycles = %.2f per non-FP64 instruction (%.1f per term)\n",
               uniform ? "uniform" : "distinct", terms_per_s, cyc_per_term, NF, cyc_per_dfma, NF * cyc_per_dfma,
               terms_per_s * NF / dfma_per_s, cyc_per_term - NF * cyc_per_dfma, (cyc_per_term - NF * cyc_per_dfma) / NO, NO);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
