// lf_engine.cu -- B200 (sm_100a) likelihood engine behind include/lf_engine.h.
//
// Work decomposition (all three models): one THREAD per WALKER, one WARP = 32 walkers sweeping a contiguous
// slab of sources (or quadrature points).  Source data are warp-uniform broadcast loads (every lane reads the
// same 16 B), walker constants live in registers, so the loop is pure FP64-pipe arithmetic; there is no
// cross-lane reduction at all -- each lane owns its walker's partial sum and writes partial[slab][walker],
// which k_finish adds up in a fixed order (deterministic, no floating-point atomics).
//
// Per call: k_prologue (unpack theta, prior gate, walker constants, fast/literal classification)
//        -> k_main<fast> + k_main<literal> (source sums and quadrature, one grid of warp work items each)
//        -> k_finish (fixed-order reduction, sufficient statistics, -inf semantics).
//
// "fast" kernels use the hoisted log-space form with the lf_math.cuh routines; they are only
// used for walkers for which the prologue PROVES that no term of the reference's product-then-log can
// underflow or leave the validated argument ranges.  Every other walker is evaluated by the "literal" kernels,
// which follow the reference's order of operations with IEEE/libdevice arithmetic so that -inf / denormal
// behaviour is reproduced, not imitated (lumfuncmcmc.py:370).
#include "../../include/lf_engine.h"
#include "lf_math.cuh"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

using namespace lfm;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}
#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                            \
    } while (0)

extern "C" const char* lf_last_error(void) { return g_err.c_str(); }
extern "C" const char* lf_version(void) { return "lfengine 0.1 (sm_100a)"; }
extern "C" int lf_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

// ------------------------------------------------------------------------------------------------
// device-side data layout
// ------------------------------------------------------------------------------------------------
// walker-parameter slots: wp[slot * Wcap + w]
enum {
    P_ALPHA = 0,   // completeness slope alpha_c
    P_TENML = 1,   // 10^-L*
    P_C0 = 2,      // ln ln10 + phi* ln10 - L* c1
    P_C1 = 3,      // (alpha_s + 1) ln10
    P_LSTAR = 4,
    P_PHISTAR = 5,
    P_SCHAL = 6,
    P_LNPART0 = 7,  // source-sum part that collapses to sufficient statistics (fast class)
    // z model: quadratic coefficients
    P_AL = 8, P_BL = 9, P_CL = 10, P_AP = 11, P_BP = 12, P_CP = 13,
    P_FIELD0 = 16,  // + 4*k + {0: aF = -alpha log10 F50, 1: c2 = -log2(e)/ftau, 2: F50 (cgs), 3: ftau}
    P_NSLOTS = P_FIELD0 + 4 * LF_MAX_FIELDS
};

enum { CLS_NONE = 0, CLS_FAST = 1, CLS_LIT = 2 };

struct FieldStats {          // per-field sufficient statistics and ranges of the resident sources
    double n, sum_lum, sum_L, sum_lnom, sum_z, sum_z2;
    double lum_min, lum_max, g_min, f_min, lnom_min, z_min, z_max;
    double grid_g_min, grid_f_min;     // same ranges over the field's quadrature points (FREE)
    double ln_om0;                      // ln(int(Omega_0)/sqarcsec)                        (FREE)
    double om0_over_sq;                 // int(Omega_0)/sqarcsec                            (FREE)
};

// 16-byte aligned so a point is fetched with LDG.128s
struct __align__(16) QuadPointFree { double g, f, x, Lx, wt, ftrue; };   // log10 flux, flux, logL, 10^logL, trapezoid*volume*area weight
struct __align__(16) QuadPoint { double x, Lx, wt, pad; };             // FIXED / Z (weight carries integ_part)

struct KArgs {
    int model, K, S, fix_sch_al, fixed_prior_ok, force_literal, modified, prior_gate;
    int ndim;
    double fcmin, fcA2;                // fcA2 = |a/(1-a)|, a = (2 fcmin - 1)^2        (VmaxLumFunc.py:164-165)
    double sch_al;
    double Lstar_lims[2], phistar_lims[2], sch_al_lims[2], Flim_lims[2], alpha_lims[2];
    double z1, z2, z3;
    long long field_ind[LF_MAX_FIELDS + 1];
    FieldStats fs[LF_MAX_FIELDS];
    double lum_max_all;
    double fcap;                       // cap of the flux copy used in the decay argument (see k_derive_free)
    // resident arrays
    const double2* src2;               // FREE: (log10 flux, flux)   Z: (lum, z)
    const double2* csrc;               // compressed catalogue, two entries per pseudo-source: (xi = log10 f, min(10^xi, fcap)), (weight, -); or NULL
    long long M; long long cfield_ind[LF_MAX_FIELDS + 1];
    const float2* src2f;               // LF_PREC_F32 copy: FREE (log10 f + 17, f * 1e17)   Z: (lum - 42, z - z2)
    int precision;                     // LF_PREC_F64 | LF_PREC_F32 (arithmetic of the walker x source loop only)
    const double* lum;
    const double* flux;
    const double* z;
    const double* om_arr;
    const double* zarr;                // Z: quadrature redshifts (column i <-> zarr[i])
    const QuadPointFree* qpf;
    const QuadPoint* qp;
    long long N, NQ;                   // sources, quadrature points (K*S*S)
    // per-call
    const double* thetas;
    double* out;
    long long W, Wcap;
    double* wp;
    double* colA;                      // Z: per (column, walker) ln-amplitude  [K? no: S][Wcap]
    double* colB;                      // Z: per (column, walker) 10^-L*(z_col)
    int* cls_count;                    // [0..2] walkers per class, [3] / [4] work-item counters of k_main<fast/literal>,
                                       // [5] / [6] walkers per class in the quadrature lists
    int* list_fast;
    int* list_lit;
    int* list_fastq;                   // walkers of each class whose quadrature THIS rank integrates (w % nshare == share)
    int* list_litq;
    double* partial;                   // [rows][Wcap]
    int n_src_slabs, n_quad_slabs;
    int share, nshare;
    const Tables* tables;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ bool in_box(double v, const double* lims) { return (v >= lims[0]) && (v <= lims[1]); }
__device__ __forceinline__ bool in_box_strict(double v, const double* lims) { return (v > lims[0]) && (v < lims[1]); }

// literal modified-Fleming value, reference operation order (VmaxLumFunc.py:118-126, 141, 164-167)
__device__ __forceinline__ double fleming_literal(double f, double F50, double alpha, double ftau, bool modified) {
    double num = alpha * log10(f / F50);
    double den = sqrt(1.0 + num * num);
    double fc = 0.5 * (1.0 + num / den);
    if (!modified) return fc;
    double dec = 1.0 - exp(-f / ftau);
    return pow(fc, 1.0 / dec);
}

// literal Schechter value (lumfuncmcmc.py:44)
__device__ __forceinline__ double schechter_literal(double logL, double sch_al, double Lstar, double phistar) {
    double dex = logL - Lstar;
    return LN10 * pow(10.0, phistar) * pow(10.0, dex * (sch_al + 1.0)) * exp(-pow(10.0, dex));
}

// getQuadCoef, reference operation order (lumfuncmcmc_z.py:40-42)
__device__ __forceinline__ void quad_coef(double y1, double y2, double y3, double z1, double z2, double z3,
                                          double& a, double& b, double& c) {
    a = ((y3 - y1) + (y2 - y1) * (z1 - z3) / (z2 - z1)) /
        (z3 * z3 - z1 * z1 + (z2 * z2 - z1 * z1) * (z1 - z3) / (z2 - z1));
    b = (y2 - y1 - a * (z2 * z2 - z1 * z1)) / (z2 - z1);
    c = y1 - a * z1 * z1 - b * z1;
}

// ---- FP32 mode of the walker x source loop: MUFU (SFU) transcendentals, FP32 FMA pipe, chunked accumulation ----
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// log2 of the modified Fleming completeness: 8 FP32-pipe instructions + 4 MUFU per (walker, source) term.
//   gs = log10 f + 17, fs = f * 1e17, aFs = -alpha * log10(F50 * 1e17), c2 = -log2(e) / (ftau * 1e17)
template <bool MODIFIED>
__device__ __forceinline__ float fleming_log2_f32(float gs, float fs, float alpha, float aFs, float c2) {
    float n = fmaf(alpha, gs, aFs);
    float y = fmaf(n, n, 1.0f);
    float q = n * mufu_rsq(y);
    float fc = fmaf(0.5f, q, 0.5f);
    float l2 = mufu_lg2(fc);
    if (!MODIFIED) return l2;
    float dec = 1.0f - mufu_ex2(fs * c2);
    return l2 * mufu_rcp(dec);
}

// ------------------------------------------------------------------------------------------------
// prologue: one thread per walker
// ------------------------------------------------------------------------------------------------
// Classification thresholds: a walker takes the fast kernels only if every term of the reference's
// product-then-log is provably >= exp(-700) (normal range, 8 above the denormal boundary) and every fast-math
// argument stays inside its validated range.
#define LB_SAFE (-700.0)
#define N_MIN_SAFE (-28.0)      /* fc >= 3.2e-4 > 2^-12, the lower end of the log table */
#define X_MIN_SAFE (1.0e-4)

// Thread (x, y) = (walker lane, field): the per-field work (a dozen libdevice pow/log/exp calls each) runs in parallel
// and thread y == 0 combines the fields in index order, so the sums are the same as a sequential loop over fields.
#define PRO_WALKERS 32
__global__ void k_prologue(KArgs a) {
    __shared__ double s_part[LF_MAX_FIELDS][PRO_WALKERS], s_lb[LF_MAX_FIELDS][PRO_WALKERS];
    __shared__ int s_rok[LF_MAX_FIELDS][PRO_WALKERS];
    const int wl = threadIdx.x, k = threadIdx.y;
    const long long w = blockIdx.x * (long long)PRO_WALKERS + wl;
    const bool live = w < a.W;
    const double* th = a.thetas + (live ? w : 0) * a.ndim;
    double* wp = a.wp + (live ? w : 0);
    const long long WS = a.Wcap;
    const double NINF = neg_inf();
    bool ok = a.fixed_prior_ok != 0;
    const bool gate = a.prior_gate != 0;      // lnlike() (no prior) vs lnprob()
    bool rejected = false;
    double part = 0.0, lb = 1.0e300;          // this field's share of the sufficient-statistics sum, and of the lower bound
    int rok = 1;
    const FieldStats& s = a.fs[k];

    if (a.model == LF_MODEL_Z) {
        double L1 = th[0], L2 = th[1], L3 = th[2], p1 = th[3], p2 = th[4], p3 = th[5];
        double sal = a.fix_sch_al ? a.sch_al : th[6];
        // lumfuncmcmc_z.py:350-358: alpha inclusive (only when sampled), L and phi strict
        if (!a.fix_sch_al) ok = ok && in_box(sal, a.sch_al_lims);
        ok = ok && in_box_strict(L1, a.Lstar_lims) && in_box_strict(L2, a.Lstar_lims) && in_box_strict(L3, a.Lstar_lims);
        ok = ok && in_box_strict(p1, a.phistar_lims) && in_box_strict(p2, a.phistar_lims) && in_box_strict(p3, a.phistar_lims);
        rejected = gate && !ok;
        double aL, bL, cL, aP, bP, cP;
        quad_coef(L1, L2, L3, a.z1, a.z2, a.z3, aL, bL, cL);
        quad_coef(p1, p2, p3, a.z1, a.z2, a.z3, aP, bP, cP);
        double c1 = (sal + 1.0) * LN10;
        if (k == 0 && live && !rejected) {
            wp[P_AL * WS] = aL; wp[P_BL * WS] = bL; wp[P_CL * WS] = cL;
            wp[P_AP * WS] = aP; wp[P_BP * WS] = bP; wp[P_CP * WS] = cP;
            wp[P_C1 * WS] = c1; wp[P_SCHAL * WS] = sal;
        }
        if (s.n != 0.0) {
            // sum_i [ ln ln10 + ln10 phi*(z_i) + c1 (lum_i - L*(z_i)) + ln Om_i ]
            double sphi = aP * s.sum_z2 + bP * s.sum_z + cP * s.n;
            double sL = aL * s.sum_z2 + bL * s.sum_z + cL * s.n;
            part = s.n * LNLN10 + LN10 * sphi + c1 * (s.sum_lum - sL) + s.sum_lnom;
            // ranges of the quadratics over [z_min, z_max]
            double Llo = fmin(fma(fma(aL, s.z_min, bL), s.z_min, cL), fma(fma(aL, s.z_max, bL), s.z_max, cL));
            double Lhi = fmax(fma(fma(aL, s.z_min, bL), s.z_min, cL), fma(fma(aL, s.z_max, bL), s.z_max, cL));
            double Plo = fmin(fma(fma(aP, s.z_min, bP), s.z_min, cP), fma(fma(aP, s.z_max, bP), s.z_max, cP));
            if (aL != 0.0) { double zv = -bL / (2.0 * aL); if (zv > s.z_min && zv < s.z_max) { double v = fma(fma(aL, zv, bL), zv, cL); Llo = fmin(Llo, v); Lhi = fmax(Lhi, v); } }
            if (aP != 0.0) { double zv = -bP / (2.0 * aP); if (zv > s.z_min && zv < s.z_max) { double v = fma(fma(aP, zv, bP), zv, cP); Plo = fmin(Plo, v); } }
            double dlo = s.lum_min - Lhi, dhi = s.lum_max - Llo;
            double emax = pow(10.0, dhi);
            lb = LNLN10 + LN10 * Plo + fmin(c1 * dlo, c1 * dhi) - emax + s.lnom_min;
            if (!(emax < 690.0)) lb = -1.0e300;
            if (!(dlo > -40.0)) lb = -1.0e300;
        }
    } else {
        const int K = a.K;
        double Lstar = th[0], phistar = th[1];
        int p = 2;
        double sal = a.sch_al;
        if (!a.fix_sch_al) sal = th[p++];
        // lumfuncmcmc.py:347-354: inclusive boxes on every parameter; parameters not in theta were checked on
        // the host (fixed_prior_ok)
        ok = ok && in_box(Lstar, a.Lstar_lims) && in_box(phistar, a.phistar_lims);
        if (!a.fix_sch_al) ok = ok && in_box(sal, a.sch_al_lims);
        double alpha_c = 0.0;
        if (a.model == LF_MODEL_FREE) {
            for (int kk = 0; kk < K; ++kk) ok = ok && in_box(th[p + kk], a.Flim_lims);
            alpha_c = th[p + K];
            ok = ok && in_box(alpha_c, a.alpha_lims);
        }
        rejected = gate && !ok;
        // certain underflow: exp(-10^(lum_max - L*)) == 0 makes Phi == 0 for the brightest source
        if (!rejected && a.N > 0 && exp(-pow(10.0, a.lum_max_all - Lstar)) == 0.0) rejected = true;
        double tenmL = pow(10.0, -Lstar);
        double c1 = (sal + 1.0) * LN10;
        double c0 = LNLN10 + phistar * LN10 - Lstar * c1;
        if (k == 0 && live && !rejected) {
            wp[P_TENML * WS] = tenmL; wp[P_C0 * WS] = c0; wp[P_C1 * WS] = c1;
            wp[P_LSTAR * WS] = Lstar; wp[P_PHISTAR * WS] = phistar; wp[P_SCHAL * WS] = sal;
            wp[P_ALPHA * WS] = alpha_c;
        }
        double tmin = 0.0, lnom = 0.0;
        if (a.model == LF_MODEL_FREE) {
            double b = -1.0 * sqrt(a.fcA2 * pow(alpha_c, -2.0));   // inverse_fleming, VmaxLumFunc.py:164-165
            if (!(alpha_c > 0.0)) rok = 0;
            double F50 = 1.0e-17 * th[p + k];
            double lgF = log10(F50);
            double ftau = F50 * pow(10.0, b);
            if (live && !rejected) {
                wp[(P_FIELD0 + 4 * k + 0) * WS] = -alpha_c * lgF;
                wp[(P_FIELD0 + 4 * k + 1) * WS] = -LOG2E / ftau;
                wp[(P_FIELD0 + 4 * k + 2) * WS] = F50;
                wp[(P_FIELD0 + 4 * k + 3) * WS] = ftau;
            }
            // faintest flux the fast math will see in this field: sources and quadrature points
            double gmin = fmin(s.n > 0.0 ? s.g_min : 1.0e300, s.grid_g_min);
            double fmn = fmin(s.n > 0.0 ? s.f_min : 1.0e300, s.grid_f_min);
            if (!(alpha_c * (gmin - lgF) > (a.precision == LF_PREC_F32 ? -12.0 : N_MIN_SAFE))) rok = 0;
            if (a.modified && !(fmn / ftau > X_MIN_SAFE)) rok = 0;
            if (a.modified && !(a.fcap / ftau < 5.0e6)) rok = 0;      // exp range reduction stays in int32
            if (s.n > 0.0) tmin = log(fleming_literal(s.f_min, F50, alpha_c, ftau, a.modified != 0));
            lnom = s.ln_om0;
        }
        if (s.n != 0.0) {
            // Schechter exponent S(l) = c0 + c1 l - 10^(l - L*) is concave in l: minimum at an end of the range
            double s_lo = fmin(c0 + c1 * s.lum_min - pow(10.0, s.lum_min - Lstar),
                               c0 + c1 * s.lum_max - pow(10.0, s.lum_max - Lstar));
            if (a.model == LF_MODEL_FREE) {
                lb = s_lo + lnom + tmin;
                part = s.n * (c0 + lnom) + c1 * s.sum_lum - tenmL * s.sum_L;
            } else {
                lb = s_lo + s.lnom_min;
                part = s.n * c0 + c1 * s.sum_lum - tenmL * s.sum_L + s.sum_lnom;
            }
        }
    }
    s_part[k][wl] = part; s_lb[k][wl] = lb; s_rok[k][wl] = rok;
    __syncthreads();
    if (k != 0 || !live) return;
    if (rejected) { a.out[w] = NINF; atomicAdd(&a.cls_count[CLS_NONE], 1); return; }
    double lnpart0 = 0.0, lbm = 1.0e300;
    bool range_ok = true;
    for (int kk = 0; kk < a.K; ++kk) {
        if (a.fs[kk].n != 0.0) { lnpart0 += s_part[kk][wl]; lbm = fmin(lbm, s_lb[kk][wl]); }
        range_ok = range_ok && s_rok[kk][wl] != 0;
    }
    wp[P_LNPART0 * WS] = lnpart0;
    int cls = CLS_FAST;
    if (!range_ok || !(lbm > LB_SAFE)) cls = CLS_LIT;
    if (a.force_literal) cls = CLS_LIT;
    int pos = atomicAdd(&a.cls_count[cls], 1);
    (cls == CLS_FAST ? a.list_fast : a.list_lit)[pos] = (int)w;
    if (a.nshare <= 1 || (w % a.nshare) == a.share) {          // a function of w only: every rank agrees on who integrates w
        int posq = atomicAdd(&a.cls_count[cls == CLS_FAST ? 5 : 6], 1);
        (cls == CLS_FAST ? a.list_fastq : a.list_litq)[posq] = (int)w;
    }
}

// Z model: per (column i, walker) constants of the quadrature integrand
//   colA = ln ln10 + ln10 phi*(z_i) - c1 L*(z_i),   colB = 10^-L*(z_i)
__global__ void k_zcolumns(KArgs a) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int nf = a.cls_count[CLS_FAST];
    if (idx >= (long long)a.S * nf) return;
    int i = (int)(idx / nf);
    long long w = a.list_fast[idx % nf];
    const long long WS = a.Wcap;
    const double* wp = a.wp + w;
    double z = a.zarr[i];
    double Ls = wp[P_AL * WS] * z * z + wp[P_BL * WS] * z + wp[P_CL * WS];     // lumfuncmcmc_z.py:65-66 order
    double Ps = wp[P_AP * WS] * z * z + wp[P_BP * WS] * z + wp[P_CP * WS];
    double c1 = wp[P_C1 * WS];
    a.colA[(long long)i * WS + w] = LNLN10 + LN10 * Ps - c1 * Ls;
    a.colB[(long long)i * WS + w] = pow(10.0, -Ls);
}

// ------------------------------------------------------------------------------------------------
// main kernels: grid of warp work items = (walker group of 32) x (slab of sources | slab of quadrature points)
// ------------------------------------------------------------------------------------------------
// one 12-warp block per SM (3 warps per scheduler, up to 168 registers per thread): the shared-memory tables
// (130 KB) are filled once per SM.  Measured sweep of (sources in lock-step, warps): profiles/README.md
#ifndef LF_WARPS_PER_BLOCK
#define LF_WARPS_PER_BLOCK 12
#endif
#ifndef LF_MIN_BLOCKS
#define LF_MIN_BLOCKS 1
#endif
#ifndef LF_ILP
#define LF_ILP 4
#endif
#define WARPS_PER_BLOCK LF_WARPS_PER_BLOCK
#define BLOCK_THREADS (32 * WARPS_PER_BLOCK)
// per-warp staging buffer of the quadrature loop: QSTAGE points of 48 B, filled with coalesced loads
#define QSTAGE 64
// Per-model launch shape of the fast kernels.  FREE: 12 warps (up to 168 registers), shared memory = log table + exp table +
// staging.  Z / FIXED: lighter loops (<= 128 registers) -> 16 warps; no log table, the small replicated exp table + staging.
__host__ __device__ constexpr int main_warps(int model) { return model == LF_MODEL_FREE ? WARPS_PER_BLOCK : 16; }
__host__ __device__ constexpr size_t main_table_bytes(int model) {
    return model == LF_MODEL_FREE ? sizeof(double2) * LOG_TAB_N * LOG_TAB_REP + sizeof(double) * EXP_SMEM_DOUBLES
                                  : sizeof(double) * EXP_TAB_N * EXP_TAB_REP;
}
__host__ __device__ constexpr size_t main_smem_bytes(int model) {
    return main_table_bytes(model) + (size_t)main_warps(model) * QSTAGE * 3 * sizeof(double2);
}

__device__ __forceinline__ int field_of(const KArgs& a, long long i) {
    int k = 0;
    while (k + 1 < a.K && i >= a.field_ind[k + 1]) ++k;
    return k;
}

// One instantiation per (class, model): each model's loop gets its own register allocation and instruction schedule.
template <bool LITERAL, int MODEL>
__global__ void __launch_bounds__(32 * main_warps(MODEL), LF_MIN_BLOCKS) k_main(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_tables[];       // fast kernels only (main_smem_bytes(MODEL))
    double2* s_log = reinterpret_cast<double2*>(smem_tables);          // FREE: log table; Z / FIXED: the replicated exp table
    double* s_exp = reinterpret_cast<double*>(s_log + LOG_TAB_N * LOG_TAB_REP);     // FREE only
    double2* s_stage = reinterpret_cast<double2*>(smem_tables + main_table_bytes(MODEL)) + (threadIdx.x >> 5) * (QSTAGE * 3);
    const int cls = LITERAL ? CLS_LIT : CLS_FAST;
    const int count_src = a.cls_count[cls], count_quad = a.cls_count[LITERAL ? 6 : 5];
    const int n_wg = (count_src + 31) >> 5, n_wgq = (count_quad + 31) >> 5;
    if (n_wg == 0) return;
    const long long n_src_items = (long long)n_wg * a.n_src_slabs;
    const long long n_items = n_src_items + (long long)n_wgq * a.n_quad_slabs;
    if (!LITERAL) {
        // FREE: log table + (big) exp table; Z / FIXED: no log table, the small replicated exp table takes its place
        if (MODEL == LF_MODEL_FREE) load_tables(a.tables, s_exp, s_log);
        else load_exp_replicated(a.tables, reinterpret_cast<double*>(s_log));
        __syncthreads();
    }
    // persistent warps: every warp pulls (walker group, slab) items from a global counter until none are left.
    // Items are small (tens per warp slot), so SMs finish within one item of each other (no wave tail), and each
    // item owns its own partial[] row, so the result does not depend on which warp ran it.
    const int lane = threadIdx.x & 31;
    const int rep16 = lane & (EXP_TAB_REP - 1), rep8 = lane & (LOG_TAB_REP - 1);   // table replica of this lane
    const double* s_exp_rep = reinterpret_cast<const double*>(s_log);               // Z / FIXED models only
    const long long WS = a.Wcap;
    int* counter = a.cls_count + (LITERAL ? 4 : 3);
  for (;;) {
    long long item = 0;
    if (lane == 0) item = atomicAdd(counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    // source items first (all walkers of the class), then quadrature items (the walkers this rank integrates)
    const bool is_src = item < n_src_items;
    const long long it = is_src ? item : item - n_src_items;
    const int groups = is_src ? n_wg : n_wgq;
    const int wg = (int)(it % groups);
    const int row = (int)(it / groups) + (is_src ? 0 : a.n_src_slabs);
    const int count = is_src ? count_src : count_quad;
    const int* list = is_src ? (LITERAL ? a.list_lit : a.list_fast) : (LITERAL ? a.list_litq : a.list_fastq);
    const int slot = wg * 32 + lane;
    const bool active = slot < count;
    const long long w = list[active ? slot : count - 1];       // inactive lanes shadow a valid walker
    const double* wp = a.wp + w;
    double acc0 = 0.0, acc1 = 0.0;

    if (row < a.n_src_slabs) {
        // ---------------- source slab ----------------
        long long i0 = (a.N * row) / a.n_src_slabs, i1 = (a.N * (row + 1)) / a.n_src_slabs;
        if (MODEL == LF_MODEL_FREE && !LITERAL && a.csrc != nullptr) {
            // compressed catalogue: sum_m w_m t(xi_m) over this slab of pseudo-sources (see lumfuncmcmc_b200/compress.py)
            const double alpha = wp[P_ALPHA * WS];
            long long m0 = (a.M * row) / a.n_src_slabs, m1 = (a.M * (row + 1)) / a.n_src_slabs;
            int k = 0;
            while (k + 1 < a.K && m0 >= a.cfield_ind[k + 1]) ++k;
            while (m0 < m1) {
                const long long seg_end = a.cfield_ind[k + 1] < m1 ? a.cfield_ind[k + 1] : m1;
                const double aF = wp[(P_FIELD0 + 4 * k + 0) * WS], c2 = wp[(P_FIELD0 + 4 * k + 1) * WS];
                long long m = m0;
                for (; m + 1 < seg_end; m += 2) {
                    const double2 u0 = __ldg(&a.csrc[2 * m]), u1 = __ldg(&a.csrc[2 * m + 2]);
                    const double w0 = __ldg(&a.csrc[2 * m + 1]).x, w1 = __ldg(&a.csrc[2 * m + 3]).x;
                    double lg0, rd0, lg1, rd1;
                    if (a.modified) {
                        fleming_log_parts<true>(u0.x, u0.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg0, rd0);
                        fleming_log_parts<true>(u1.x, u1.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg1, rd1);
                    } else {
                        fleming_log_parts<false>(u0.x, u0.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg0, rd0);
                        fleming_log_parts<false>(u1.x, u1.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg1, rd1);
                    }
                    acc0 = fma(w0 * lg0, rd0, acc0);
                    acc1 = fma(w1 * lg1, rd1, acc1);
                }
                if (m < seg_end) {
                    const double2 u0 = __ldg(&a.csrc[2 * m]);
                    const double w0 = __ldg(&a.csrc[2 * m + 1]).x;
                    double lg0, rd0;
                    if (a.modified) fleming_log_parts<true>(u0.x, u0.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg0, rd0);
                    else fleming_log_parts<false>(u0.x, u0.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg0, rd0);
                    acc0 = fma(w0 * lg0, rd0, acc0);
                }
                m0 = seg_end;
                ++k;
            }
        } else if (MODEL == LF_MODEL_FREE) {
            const double alpha = wp[P_ALPHA * WS];
            int k = field_of(a, i0);
            while (i0 < i1) {
                long long seg_end = a.field_ind[k + 1] < i1 ? a.field_ind[k + 1] : i1;
                if (!LITERAL && a.precision == LF_PREC_F32) {
                    // FP32 mode: MUFU transcendentals; partial sums in FP32 over 64-source chunks, flushed to FP64
                    const double aFd = wp[(P_FIELD0 + 4 * k + 0) * WS];           // -alpha*log10(F50)
                    const float af = (float)alpha;
                    const float aFs = (float)(aFd - 17.0 * alpha);                 // -alpha*log10(F50*1e17)
                    const float c2 = (float)(wp[(P_FIELD0 + 4 * k + 1) * WS] * 1.0e-17);
                    const float2* __restrict__ pf = a.src2f + i0;
                    const int cnt = (int)(seg_end - i0);
                    double sum = 0.0;
                    for (int j0 = 0; j0 < cnt; j0 += 64) {
                        const int j1 = min(j0 + 64, cnt);
                        float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
                        int j = j0;
                        if (a.modified) {
#ifndef LF_F32_ILP
#define LF_F32_ILP 16
#endif
                            for (; j + LF_F32_ILP <= j1; j += LF_F32_ILP) {
                                float2 u[LF_F32_ILP];
                                float tt[LF_F32_ILP];
#pragma unroll
                                for (int t = 0; t < LF_F32_ILP; ++t) u[t] = __ldg(pf + j + t);
#pragma unroll
                                for (int t = 0; t < LF_F32_ILP; ++t) tt[t] = fleming_log2_f32<true>(u[t].x, u[t].y, af, aFs, c2);
#pragma unroll
                                for (int t = 0; t < LF_F32_ILP; t += 4) { f0 += tt[t]; f1 += tt[t + 1]; f2 += tt[t + 2]; f3 += tt[t + 3]; }
                            }
                            for (; j < j1; ++j) { float2 u0 = __ldg(pf + j); f0 += fleming_log2_f32<true>(u0.x, u0.y, af, aFs, c2); }
                        } else {
                            for (; j < j1; ++j) { float2 u0 = __ldg(pf + j); f0 += fleming_log2_f32<false>(u0.x, u0.y, af, aFs, c2); }
                        }
                        sum += (double)((f0 + f1) + (f2 + f3));
                    }
                    acc0 = fma(sum, LN2, acc0);
                } else if (!LITERAL) {
                    const double aF = wp[(P_FIELD0 + 4 * k + 0) * WS], c2 = wp[(P_FIELD0 + 4 * k + 1) * WS];
                    long long i = i0;
                    if (a.modified) {
                        // LF_ILP sources per group, evaluated in lock-step (LF_ILP independent FP64 chains per thread);
                        // groups are double-buffered: the next group's broadcast loads are issued before the current
                        // group's arithmetic.  32-bit trip counter.
                        const double2* __restrict__ ps = a.src2 + i;
                        const int cnt = (int)(seg_end - i);
                        constexpr int NT = LF_ILP;
                        double2 A[NT], B[NT];
                        double accv[NT];
#pragma unroll
                        for (int t = 0; t < NT; ++t) accv[t] = 0.0;
                        int j = 0;
                        if (cnt >= NT) {
#pragma unroll
                            for (int t = 0; t < NT; ++t) A[t] = __ldg(ps + t);
                        }
                        for (; j + 2 * NT <= cnt; j += 2 * NT) {
#pragma unroll
                            for (int t = 0; t < NT; ++t) B[t] = __ldg(ps + j + NT + t);
                            fleming_terms<NT>(A, alpha, aF, c2, s_exp, s_log, rep16, rep8, accv);
                            if (j + 3 * NT <= cnt) {
#pragma unroll
                                for (int t = 0; t < NT; ++t) A[t] = __ldg(ps + j + 2 * NT + t);
                            }
                            fleming_terms<NT>(B, alpha, aF, c2, s_exp, s_log, rep16, rep8, accv);
                        }
                        if (j + NT <= cnt) {
                            fleming_terms<NT>(A, alpha, aF, c2, s_exp, s_log, rep16, rep8, accv);
                            j += NT;
                        }
#pragma unroll
                        for (int t = 0; t < NT; t += 2) { acc0 += accv[t]; acc1 += accv[t + 1]; }
                        for (; j < cnt; ++j) {
                            double2 s0 = __ldg(ps + j);
                            double lg0, rd0;
                            fleming_log_parts<true>(s0.x, s0.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg0, rd0);
                            acc0 = fma(lg0, rd0, acc0);
                        }
                    } else {
                        for (; i < seg_end; ++i) {
                            double2 s0 = __ldg(&a.src2[i]);
                            double lg0, rd0;
                            fleming_log_parts<false>(s0.x, s0.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg0, rd0);
                            acc0 += lg0;
                        }
                    }
                } else {
                    // reference order: log( Phi_i * (int(Omega_0)/sqarcsec * fleming(f_i)) )  (lumfuncmcmc.py:370)
                    const double F50 = wp[(P_FIELD0 + 4 * k + 2) * WS], ftau = wp[(P_FIELD0 + 4 * k + 3) * WS];
                    const double Lstar = wp[P_LSTAR * WS], phistar = wp[P_PHISTAR * WS], sal = wp[P_SCHAL * WS];
                    const double om = a.fs[k].om0_over_sq;
                    for (long long i = i0; i < seg_end; ++i) {
                        double phi = schechter_literal(__ldg(&a.lum[i]), sal, Lstar, phistar);
                        double Om = om * fleming_literal(__ldg(&a.flux[i]), F50, alpha, ftau, a.modified != 0);
                        acc0 += log(phi * Om);
                    }
                }
                i0 = seg_end;
                ++k;
            }
        } else if (MODEL == LF_MODEL_FIXED) {
            // fast class: the whole source sum is in P_LNPART0 (sufficient statistics); literal class sums terms
            if (LITERAL) {
                const double Lstar = wp[P_LSTAR * WS], phistar = wp[P_PHISTAR * WS], sal = wp[P_SCHAL * WS];
                for (long long i = i0; i < i1; ++i)                                   // lumfuncmcmc.py:388
                    acc0 += log(schechter_literal(__ldg(&a.lum[i]), sal, Lstar, phistar) * __ldg(&a.om_arr[i]));
            }
        } else {
            const double aL = wp[P_AL * WS], bL = wp[P_BL * WS], cL = wp[P_CL * WS];
            if (!LITERAL && a.precision == LF_PREC_F32) {
                // FP32 mode: L*(z) re-centred on the middle pivot and on 42 so the float polynomial keeps ~1e-7 dex
                const double z2 = a.z2;
                const float q2 = (float)aL, q1 = (float)(bL + 2.0 * aL * z2);
                const float q0 = (float)((fma(fma(aL, z2, bL), z2, cL) - 42.0));
                const float L2T = 3.3219280948873623f;                              // log2(10)
                const float2* __restrict__ pf = a.src2f + i0;
                const long long cnt = i1 - i0;
                double sum = 0.0;
                for (long long j0 = 0; j0 < cnt; j0 += 64) {
                    const long long j1 = j0 + 64 < cnt ? j0 + 64 : cnt;
                    float f0 = 0.f, f1 = 0.f;
                    long long j = j0;
                    for (; j + 8 <= j1; j += 8) {                   // eight independent MUFU chains per thread
                        float2 u[8];
                        float e[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) u[t] = __ldg(pf + j + t);
#pragma unroll
                        for (int t = 0; t < 8; ++t) e[t] = mufu_ex2((u[t].x - fmaf(fmaf(q2, u[t].y, q1), u[t].y, q0)) * L2T);
#pragma unroll
                        for (int t = 0; t < 8; t += 2) { f0 += e[t]; f1 += e[t + 1]; }
                    }
                    for (; j < j1; ++j) { float2 u0 = __ldg(pf + j); f0 += mufu_ex2((u0.x - fmaf(fmaf(q2, u0.y, q1), u0.y, q0)) * L2T); }
                    sum += (double)(f0 + f1);
                }
                acc0 -= sum;
            } else if (!LITERAL) {
                // only sum_i 10^(lum_i - L*(z_i)) needs the walker x source loop; the rest is in P_LNPART0.
                // Base 2 throughout: 2^(log2(10) lum_i - P2(z_i)), 11 FP64 instructions per term
                const double L2T = 3.32192809488736234787;                              // log2(10)
                const double a2 = aL * L2T, b2 = bL * L2T, c2 = cL * L2T;
                // four sources in lock-step (four independent FP64 chains per thread), the next four prefetched
                const double2* __restrict__ ps = a.src2 + i0;
                const int cnt = (int)(i1 - i0);
                double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
                double2 A0, A1, A2, A3;
                int j = 0;
                if (cnt >= 4) { A0 = __ldg(ps); A1 = __ldg(ps + 1); A2 = __ldg(ps + 2); A3 = __ldg(ps + 3); }
                for (; j + 4 <= cnt; j += 4) {
                    const double2 s0 = A0, s1 = A1, s2 = A2, s3 = A3;
                    if (j + 8 <= cnt) { A0 = __ldg(ps + j + 4); A1 = __ldg(ps + j + 5); A2 = __ldg(ps + j + 6); A3 = __ldg(ps + j + 7); }
                    const double d0 = fma(L2T, s0.x, -fma(fma(a2, s0.y, b2), s0.y, c2));
                    const double d1 = fma(L2T, s1.x, -fma(fma(a2, s1.y, b2), s1.y, c2));
                    const double d2 = fma(L2T, s2.x, -fma(fma(a2, s2.y, b2), s2.y, c2));
                    const double d3 = fma(L2T, s3.x, -fma(fma(a2, s3.y, b2), s3.y, c2));
                    e0 += exp2_full<false>(d0, s_exp_rep, rep16);
                    e1 += exp2_full<false>(d1, s_exp_rep, rep16);
                    e2 += exp2_full<false>(d2, s_exp_rep, rep16);
                    e3 += exp2_full<false>(d3, s_exp_rep, rep16);
                }
                for (; j < cnt; ++j) {
                    const double2 s0 = __ldg(ps + j);
                    e0 += exp2_full<false>(fma(L2T, s0.x, -fma(fma(a2, s0.y, b2), s0.y, c2)), s_exp_rep, rep16);
                }
                acc0 -= (e0 + e1) + (e2 + e3);
            } else {
                const double aP = wp[P_AP * WS], bP = wp[P_BP * WS], cP = wp[P_CP * WS], sal = wp[P_SCHAL * WS];
                for (long long i = i0; i < i1; ++i) {                                 // lumfuncmcmc_z.py:371
                    double z = __ldg(&a.z[i]);
                    double ps = aP * z * z + bP * z + cP, Ls = aL * z * z + bL * z + cL;
                    acc0 += log(schechter_literal(__ldg(&a.lum[i]), sal, Ls, ps) * __ldg(&a.om_arr[i]));
                }
            }
        }
    } else {
        // ---------------- quadrature slab: contributes the integral (k_finish subtracts it) ----------------
        {
            const int qrow = row - a.n_src_slabs;
            long long q0 = (a.NQ * qrow) / a.n_quad_slabs, q1 = (a.NQ * (qrow + 1)) / a.n_quad_slabs;
            const long long SS = (long long)a.S * a.S;
            if (MODEL == LF_MODEL_FREE) {
                const double alpha = wp[P_ALPHA * WS];
                const double c0 = wp[P_C0 * WS], c1 = wp[P_C1 * WS], tenmL = wp[P_TENML * WS];
                const double Lstar = wp[P_LSTAR * WS], phistar = wp[P_PHISTAR * WS], sal = wp[P_SCHAL * WS];
                while (q0 < q1) {
                    int k = (int)(q0 / SS);
                    long long seg_end = (k + 1) * SS < q1 ? (k + 1) * SS : q1;
                    if (!LITERAL) {
                        const double aF = wp[(P_FIELD0 + 4 * k + 0) * WS], c2 = wp[(P_FIELD0 + 4 * k + 1) * WS];
                        // one quadrature point: weight * exp(Schechter exponent + ln completeness)
                        auto point = [&](const double2* pt, double& acc) {
                            double2 gf = pt[0], xl = pt[1];
                            double wt = pt[2].x;
                            double lg, rd;
                            if (a.modified) fleming_log_parts<true>(gf.x, gf.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg, rd);
                            else fleming_log_parts<false>(gf.x, gf.y, alpha, aF, c2, s_exp, s_log, rep16, rep8, lg, rd);
                            double arg = fma(c1, xl.x, c0);
                            arg = fma(-xl.y, tenmL, arg);
                            arg = fma(lg, rd, arg);
                            acc = fma(wt, exp_full<EXP_BIG>(arg, s_exp, rep16), acc);
                        };
                        // the warp stages QSTAGE points at a time in shared memory with coalesced 16-byte loads (all of
                        // them in flight at once), then every lane reads the points back as broadcasts
                        for (long long q = q0; q < seg_end; q += QSTAGE) {
                            const int cnt = (int)(seg_end - q < QSTAGE ? seg_end - q : QSTAGE);
                            const double2* __restrict__ src = reinterpret_cast<const double2*>(a.qpf + q);
                            __syncwarp();
                            for (int t = lane; t < cnt * 3; t += 32) s_stage[t] = __ldg(src + t);
                            __syncwarp();
                            int j = 0;
                            if (a.modified) {
                                constexpr int NT = 4;
                                for (; j + NT <= cnt; j += NT) {
                                    double2 u[NT];
                                    double arg[NT], wt[NT];
#pragma unroll
                                    for (int t = 0; t < NT; ++t) {
                                        u[t] = s_stage[(j + t) * 3];
                                        double2 xl = s_stage[(j + t) * 3 + 1];
                                        wt[t] = s_stage[(j + t) * 3 + 2].x;
                                        arg[t] = fma(-xl.y, tenmL, fma(c1, xl.x, c0));
                                    }
                                    fleming_terms<NT>(u, alpha, aF, c2, s_exp, s_log, rep16, rep8, arg);   // arg += ln completeness
#pragma unroll
                                    for (int t = 0; t < NT; t += 2) {
                                        acc0 = fma(wt[t], exp_full<EXP_BIG>(arg[t], s_exp, rep16), acc0);
                                        acc1 = fma(wt[t + 1], exp_full<EXP_BIG>(arg[t + 1], s_exp, rep16), acc1);
                                    }
                                }
                            }
                            for (; j < cnt; ++j) point(s_stage + j * 3, acc0);
                        }
                    } else {
                        const double F50 = wp[(P_FIELD0 + 4 * k + 2) * WS], ftau = wp[(P_FIELD0 + 4 * k + 3) * WS];
                        for (long long q = q0; q < seg_end; ++q) {                    // lumfuncmcmc.py:375-376
                            const QuadPointFree* pt = &a.qpf[q];
                            double y = schechter_literal(__ldg(&pt->x), sal, Lstar, phistar) *
                                       fleming_literal(__ldg(&pt->ftrue), F50, alpha, ftau, a.modified != 0);
                            acc0 = fma(__ldg(&pt->wt), y, acc0);
                        }
                    }
                    q0 = seg_end;
                }
            } else {
                // FIXED / Z fast path: sum_q wt_q exp(c1 x_q - Lx_q cB + cA) over [qa, qb); the warp stages QSTAGE points
                // (32 B each) in shared memory with coalesced loads and evaluates them four at a time
                const double c1 = wp[P_C1 * WS];
                auto quad_segment = [&](long long qa, long long qb, double cA, double cB) {
                    for (long long q = qa; q < qb; q += QSTAGE) {
                        const int cnt = (int)(qb - q < QSTAGE ? qb - q : QSTAGE);
                        const double2* __restrict__ src = reinterpret_cast<const double2*>(a.qp + q);
                        __syncwarp();
                        for (int t = lane; t < cnt * 2; t += 32) s_stage[t] = __ldg(src + t);
                        __syncwarp();
                        int j = 0;
                        for (; j + 4 <= cnt; j += 4) {
                            double arg[4], wt[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const double2 xl = s_stage[(j + t) * 2];
                                wt[t] = s_stage[(j + t) * 2 + 1].x;
                                arg[t] = fma(-xl.y, cB, fma(c1, xl.x, cA));
                            }
#pragma unroll
                            for (int t = 0; t < 4; t += 2) {
                                acc0 = fma(wt[t], exp_full<false>(arg[t], s_exp_rep, rep16), acc0);
                                acc1 = fma(wt[t + 1], exp_full<false>(arg[t + 1], s_exp_rep, rep16), acc1);
                            }
                        }
                        for (; j < cnt; ++j) {
                            const double2 xl = s_stage[j * 2];
                            acc0 = fma(s_stage[j * 2 + 1].x, exp_full<false>(fma(-xl.y, cB, fma(c1, xl.x, cA)), s_exp_rep, rep16), acc0);
                        }
                    }
                };
                if (MODEL == LF_MODEL_FIXED) {
                    if (!LITERAL) {
                        quad_segment(q0, q1, wp[P_C0 * WS], wp[P_TENML * WS]);
                    } else {
                        const double Lstar = wp[P_LSTAR * WS], phistar = wp[P_PHISTAR * WS], sal = wp[P_SCHAL * WS];
                        for (long long q = q0; q < q1; ++q) {                         // lumfuncmcmc.py:391
                            const QuadPoint* pt = &a.qp[q];
                            acc0 = fma(__ldg(&pt->wt), schechter_literal(__ldg(&pt->x), sal, Lstar, phistar), acc0);
                        }
                    }
                } else {
                    // Z: points are stored column-major: q = (k*S + i)*S + j  (column i <-> zarr_i)
                    const double sal = wp[P_SCHAL * WS];
                    const double aL = wp[P_AL * WS], bL = wp[P_BL * WS], cL = wp[P_CL * WS];
                    const double aP = wp[P_AP * WS], bP = wp[P_BP * WS], cP = wp[P_CP * WS];
                    while (q0 < q1) {
                        long long col = q0 / a.S;                    // global column index k*S + i
                        int i = (int)(col % a.S);
                        long long seg_end = (col + 1) * a.S < q1 ? (col + 1) * a.S : q1;
                        if (!LITERAL) {
                            quad_segment(q0, seg_end, a.colA[(long long)i * WS + w], a.colB[(long long)i * WS + w]);
                        } else {
                            double z = __ldg(&a.zarr[i]);
                            double ps = aP * z * z + bP * z + cP, Ls = aL * z * z + bL * z + cL;
                            for (long long q = q0; q < seg_end; ++q) {                // lumfuncmcmc_z.py:374
                                const QuadPoint* pt = &a.qp[q];
                                acc0 = fma(__ldg(&pt->wt), schechter_literal(__ldg(&pt->x), sal, Ls, ps), acc0);
                            }
                        }
                        q0 = seg_end;
                    }
                }
            }
        }
    }
    if (active) a.partial[(long long)row * WS + w] = acc0 + acc1;
  }
}

// ------------------------------------------------------------------------------------------------
// finish: fixed-order reduction over slabs; lnprob = lnpart - fullint
// ------------------------------------------------------------------------------------------------
#define FIN_GROUPS 32
__global__ void __launch_bounds__(32 * FIN_GROUPS) k_finish(KArgs a) {
    // one block per 32 walkers; FIN_GROUPS row-groups read the slab partials in parallel (coalesced along walkers),
    // each in ascending row order, and are combined in a fixed order: deterministic, no atomics
    __shared__ double s_src[FIN_GROUPS][32], s_quad[FIN_GROUPS][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int nf = a.cls_count[CLS_FAST], nl = a.cls_count[CLS_LIT];
    const long long idx = (long long)blockIdx.x * 32 + lane;
    if ((long long)blockIdx.x * 32 >= nf + nl) return;
    const bool valid = idx < nf + nl;
    const bool fast = idx < nf;
    long long w = 0;
    if (valid) w = fast ? a.list_fast[idx] : a.list_lit[idx - nf];
    const long long WS = a.Wcap;
    double lnpart = 0.0, fullint = 0.0;
    if (valid) {
        const double* col = a.partial + w;
#pragma unroll 4
        for (int r = g; r < a.n_src_slabs; r += FIN_GROUPS) lnpart += col[(long long)r * WS];
#pragma unroll 4
        for (int r = a.n_src_slabs + g; r < a.n_src_slabs + a.n_quad_slabs; r += FIN_GROUPS)
            fullint += col[(long long)r * WS];
    }
    s_src[g][lane] = lnpart;
    s_quad[g][lane] = fullint;
    __syncthreads();
    if (g != 0 || !valid) return;
    lnpart = 0.0; fullint = 0.0;
    for (int k = 0; k < FIN_GROUPS; ++k) { lnpart += s_src[k][lane]; fullint += s_quad[k][lane]; }
    if (fast) lnpart += a.wp[P_LNPART0 * WS + w];
    if (a.nshare > 1 && (w % a.nshare) != a.share) fullint = 0.0;
    double v = lnpart - fullint;
    if (v != v) v = neg_inf();            // the engine never returns NaN
    a.out[w] = v;
}

// ------------------------------------------------------------------------------------------------
// set-up kernels: derived per-source arrays and per-field statistics
// ------------------------------------------------------------------------------------------------
__global__ void k_derive_free(long long n, const double* __restrict__ lum, const double* __restrict__ flux,
                              double2* __restrict__ src2, double* __restrict__ Lsrc, double fcap, float2* __restrict__ src2f) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f = flux[i];
    if (src2f) src2f[i] = make_float2((float)(log10(f) + 17.0), (float)(fmin(f, fcap) * 1.0e17));
    // the decay argument f/ftau only matters below ~46 (exp(-46) = 1e-20 against 1) and ftau < F50 <= prior
    // maximum, so the flux copy used for it is capped at 64 x that maximum: bit-identical results, bounded range
    src2[i] = make_double2(log10(f), fmin(f, fcap));
    Lsrc[i] = pow(10.0, lum[i]);
}
__global__ void k_derive_z(long long n, const double* __restrict__ lum, const double* __restrict__ z,
                           double2* __restrict__ src2, float2* __restrict__ src2f, double zref) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (src2f) src2f[i] = make_float2((float)(lum[i] - 42.0), (float)(z[i] - zref));
    src2[i] = make_double2(lum[i], z[i]);
}
__global__ void k_pow10(long long n, const double* __restrict__ lum, double* __restrict__ Lsrc) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lsrc[i] = pow(10.0, lum[i]);
}
__global__ void k_log(long long n, const double* __restrict__ x, double* __restrict__ y) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = log(x[i]);
}

// block partials of (sum, min, max) over a range; host adds the partials in long double
#define STAT_THREADS 256
__global__ void k_stats(const double* __restrict__ x, long long i0, long long i1, double* __restrict__ out3) {
    __shared__ double ss[STAT_THREADS], smin[STAT_THREADS], smax[STAT_THREADS];
    double s = 0.0, c = 0.0, mn = 1.0e300, mx = -1.0e300;
    for (long long i = i0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        double y = v - c, t = s + y;           // Kahan
        c = (t - s) - y;
        s = t;
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    ss[threadIdx.x] = s; smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int o = STAT_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            ss[threadIdx.x] += ss[threadIdx.x + o];
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + o]);
            smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out3[blockIdx.x * 3 + 0] = ss[0]; out3[blockIdx.x * 3 + 1] = smin[0]; out3[blockIdx.x * 3 + 2] = smax[0]; }
}
__global__ void k_square(long long n, const double* __restrict__ x, double* __restrict__ y) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] * x[i];
}
__global__ void k_take(long long n, const double2* __restrict__ s, int comp, double* __restrict__ y) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = comp ? s[i].y : s[i].x;
}

// ------------------------------------------------------------------------------------------------
// 1/V_eff weights + binned luminosity function  (HBM-bound streaming pass)
// ------------------------------------------------------------------------------------------------
#define VEFF_MAX_BINS 1024
struct VeffArgs {
    long long n;
    const double* flux; const double* lum; const double* vol; const unsigned char* valid;
    double* phi;
    int K; long long field_ind[LF_MAX_FIELDS + 1]; double F50[LF_MAX_FIELDS]; double ftau[LF_MAX_FIELDS];
    double invF50[LF_MAX_FIELDS]; double inv_ftau[LF_MAX_FIELDS];
    double alpha, pref, vol_int, inv_pref_vol; int modified;
    const Tables* tables;
    const double* edges; int nbins;
    unsigned long long* counts; double* sumphi;     // [gridDim.x][nbins] block partials
    const int* mult;                                 // bootstrap multiplicities (NULL: original sample)
    short* bin;                                      // per-source bin index (-1: none), written by MODE 0/2, read by MODE 1
};

__device__ __forceinline__ int bin_of(double L, const double* e, int nb) {
    // half-open bins [e_j, e_{j+1}), exact comparisons against the caller's edges (VmaxLumFunc.py:346-348)
    if (!(L >= e[0]) || !(L < e[nb])) return -1;
    int j = (int)((L - e[0]) / (e[nb] - e[0]) * nb);
    j = j < 0 ? 0 : (j > nb - 1 ? nb - 1 : j);
    while (j > 0 && L < e[j]) --j;
    while (j < nb - 1 && L >= e[j + 1]) ++j;
    return j;
}

// same search on an edge table replicated x16 in shared memory (e[j * 16 + col]: a half-warp never bank-conflicts),
// candidate bin from a precomputed scale instead of a division; the comparisons against the caller's exact edges decide
__device__ __forceinline__ int bin_of_rep(double L, const double* e, int col, int nb, double e0, double enb, double scale) {
    if (!(L >= e0) || !(L < enb)) return -1;
    int j = (int)((L - e0) * scale);
    j = j < 0 ? 0 : (j > nb - 1 ? nb - 1 : j);
    while (j > 0 && L < e[j * 16 + col]) --j;
    while (j < nb - 1 && L >= e[(j + 1) * 16 + col]) ++j;
    return j;
}

// MODE 0: compute phi from the completeness and bin; MODE 1: bootstrap replicate (multiplicities) on resident
// lum/phi; MODE 2: bin caller-provided (resident) phi
template <int MODE>
__global__ void __launch_bounds__(256) k_veff(VeffArgs a) {
    constexpr bool BOOT = MODE == 1;
    extern __shared__ unsigned char smem_raw[];
    double* s_edges = reinterpret_cast<double*>(smem_raw);                 // nbins+1
    double* s_sum = s_edges + (a.nbins + 1);                               // 8 warps x nbins
    unsigned long long* s_cnt = reinterpret_cast<unsigned long long*>(s_sum + 8 * a.nbins);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i <= a.nbins; i += blockDim.x) s_edges[i] = a.edges[i];
    for (int i = threadIdx.x; i < 8 * a.nbins; i += blockDim.x) { s_sum[i] = 0.0; s_cnt[i] = 0ULL; }
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        double phi;
        unsigned long long m = 1ULL;
        if (BOOT) {
            m = (unsigned long long)a.mult[i];
            if (m == 0ULL) continue;
            phi = a.phi[i];
        } else if (MODE == 2) {
            phi = a.phi[i];
        } else {
            int k = 0;
            while (k + 1 < a.K && i >= a.field_ind[k + 1]) ++k;
            double comp = fleming_literal(a.flux[i], a.F50[k], a.alpha, a.ftau[k], a.modified != 0);
            double vol = a.vol ? a.vol[i] : a.vol_int;
            bool ok = a.valid ? (a.valid[i] != 0) : true;
            phi = ok ? 1.0 / (a.pref * comp * vol) : 0.0;       // lumfuncmcmc.py:524, VmaxLumFunc.py:256-257
            a.phi[i] = phi;
        }
        int j = bin_of(a.lum[i], s_edges, a.nbins);
        if (j >= 0) {
            atomicAdd(&s_sum[warp * a.nbins + j], BOOT ? phi * (double)m : phi);
            atomicAdd(&s_cnt[warp * a.nbins + j], m);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < a.nbins; j += blockDim.x) {
        double s = 0.0; unsigned long long c = 0ULL;
        for (int wv = 0; wv < 8; ++wv) { s += s_sum[wv * a.nbins + j]; c += s_cnt[wv * a.nbins + j]; }
        a.sumphi[(long long)blockIdx.x * a.nbins + j] = s;
        a.counts[(long long)blockIdx.x * a.nbins + j] = c;
    }
}

// ---- streaming version with private histogram columns (no atomics, deterministic) ----
// Each warp owns VP_COLS = 16 columns per bin of the block's shared-memory histogram, s_sum[warp][bin][col] (f64) and
// s_cnt[warp][bin][col] (u32); lanes l and l + 16 share column l and update it in two turns separated by __syncwarp,
// so an update is a plain read-modify-write of a word nobody else touches in that turn (no atomics, no races, one
// bank per column).  At the end each warp folds its columns with a fixed shuffle tree and the block adds the warps
// in order: the result does not depend on scheduling.  Shared memory: 8 warps x nbins x 16 x 12 B (76.8 KB at the
// reference's nbins = 50 -> two blocks = 16 warps per SM); larger histograms fall back to k_veff (atomics).
// The per-source completeness is evaluated as exp(-ln(fc)/dec) with the ~2e-16 routines of lf_math.cuh (two 2-4 KB
// tables) instead of libdevice pow/log10/exp/sqrt and five IEEE divisions: ~75 FP64-pipe instructions per source, so
// the pass stays close to its HBM time (24 B per source).
#define VP_WARPS 8
#define VP_COLS 16
#define VP_UNROLL 4
static const size_t VP_SMEM_MAX = 200 * 1024;

__device__ __noinline__ double inv_fleming_literal(double f, double F50, double alpha, double ftau, bool modified) {
    return 1.0 / fleming_literal(f, F50, alpha, ftau, modified);
}

// 1 / fleming(f): VmaxLumFunc.py:118-126, 141.  Sources outside the range where the fast evaluation is accurate to
// ~1e-15 (fc < 1e-6, decay argument < 1e-6, |ln comp| > 690) take the literal libdevice route.
__device__ __forceinline__ double inv_fleming_stream(double f, double F50, double invF50, double alpha_log10e, double alpha,
                                                     double ftau, double inv_ftau, bool modified, const double* s_exp,
                                                     const double2* s_logm) {
    const double num = alpha_log10e * log_stream(f * invF50, s_logm);       // alpha * log10(f / F50)
    const double y = fma(num, num, 1.0);
    double r0 = rsqrt_seed(y);
    const double e = fma(-(y * r0), r0, 1.0);
    const double pe = fma(0.375, e, 0.5) * e;
    const double nr = num * r0;
    const double fc = fma(0.5, fma(nr, pe, nr), 0.5);
    const double x = f * inv_ftau;
    double t = -log_stream(fc > 1.0e-300 ? fc : 1.0e-300, s_logm);
    if (modified) t *= rcp_stream(1.0 - exp_stream(x < 690.0 ? -x : -690.0, s_exp));
    if (!(fc > 1.0e-6) || (modified && !(x > 1.0e-6)) || !(t < 690.0)) return inv_fleming_literal(f, F50, alpha, ftau, modified);
    return exp_stream(t, s_exp);
}

template <int MODE>
__global__ void __launch_bounds__(32 * VP_WARPS) k_veff_priv(VeffArgs a) {
    constexpr bool BOOT = MODE == 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = a.nbins;
    double* s_edges = reinterpret_cast<double*>(smem_raw);                  // [nbins + 1][16] replicated
    double* s_sum = s_edges + (nb + 1) * 16;                                // [VP_WARPS][nbins][VP_COLS]
    unsigned* s_cnt = reinterpret_cast<unsigned*>(s_sum + VP_WARPS * nb * VP_COLS);
    double2* s_logm = reinterpret_cast<double2*>(s_cnt + VP_WARPS * nb * VP_COLS);   // MODE 0 only
    double* s_exp = reinterpret_cast<double*>(s_logm + STREAM_LOG_N);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (nb + 1) * 16; i += blockDim.x) s_edges[i] = a.edges[i >> 4];
    for (int i = threadIdx.x; i < VP_WARPS * nb * VP_COLS; i += blockDim.x) { s_sum[i] = 0.0; s_cnt[i] = 0u; }
    if (MODE == 0) load_stream_tables(a.tables, s_exp, s_logm);
    const double e0 = a.edges[0], enb = a.edges[nb], scale = (double)nb / (enb - e0);
    __syncthreads();
    double* my_sum = s_sum + warp * nb * VP_COLS + (lane & (VP_COLS - 1));
    unsigned* my_cnt = s_cnt + warp * nb * VP_COLS + (lane & (VP_COLS - 1));
    const int turn = lane >> 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double alpha_log10e = a.alpha * KS[12];
    int k = 0;                                  // sources are field-sorted: the field index only moves forward
    // whole warps iterate together (the trip count is computed from the warp's first lane) so that __syncwarp is legal
    for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < a.n; i0 += stride * VP_UNROLL) {
        const long long i = i0 + lane;
        double lum[VP_UNROLL], phi[VP_UNROLL], flux[VP_UNROLL], vol[VP_UNROLL];
        unsigned m[VP_UNROLL];
        int jb[VP_UNROLL];
        bool ok[VP_UNROLL];
#pragma unroll
        for (int u = 0; u < VP_UNROLL; ++u) {                                // all loads of the trip first
            const long long ii = i + u * stride;
            const bool in = ii < a.n;
            m[u] = in ? 1u : 0u;
            ok[u] = in;
            phi[u] = 0.0;
            if (BOOT) {                                                      // replicate: resident bin index, weight, multiplicity
                jb[u] = in ? (int)__ldcs(a.bin + ii) : -1;
                m[u] = in ? (unsigned)__ldcs(a.mult + ii) : 0u;
                phi[u] = in ? __ldcs(a.phi + ii) : 0.0;
                continue;
            }
            lum[u] = in ? __ldcs(a.lum + ii) : -1.0e300;                     // below every edge: lands in no bin
            if (MODE == 2) {
                phi[u] = in ? __ldcs(a.phi + ii) : 0.0;
            } else {
                flux[u] = in ? __ldcs(a.flux + ii) : 1.0;
                vol[u] = (in && a.vol) ? __ldcs(a.vol + ii) : a.vol_int;
                if (in && a.valid) ok[u] = a.valid[ii] != 0;
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < VP_UNROLL; ++u) {
                const long long ii = i + u * stride;
                while (k + 1 < a.K && ii >= a.field_ind[k + 1]) ++k;
                const double icomp = inv_fleming_stream(flux[u], a.F50[k], a.invF50[k], alpha_log10e, a.alpha, a.ftau[k],
                                                        a.inv_ftau[k], a.modified != 0, s_exp, s_logm);
                const double ipv = a.vol ? 1.0 / (a.pref * vol[u]) : a.inv_pref_vol;
                phi[u] = ok[u] ? icomp * ipv : 0.0;                          // lumfuncmcmc.py:524, VmaxLumFunc.py:256-257
                if (ii < a.n) __stcs(a.phi + ii, phi[u]);
            }
        }
        if (!BOOT) {
#pragma unroll
            for (int u = 0; u < VP_UNROLL; ++u) {
                const long long ii = i + u * stride;
                jb[u] = bin_of_rep(lum[u], s_edges, lane & 15, nb, e0, enb, scale);
                if (ii < a.n) a.bin[ii] = (short)jb[u];                      // kept resident for the bootstrap replicates
            }
        }
#pragma unroll
        for (int tn = 0; tn < 2; ++tn) {
            if (turn == tn) {
#pragma unroll
                for (int u = 0; u < VP_UNROLL; ++u)
                    if (jb[u] >= 0 && m[u] != 0u) {
                        my_sum[jb[u] * VP_COLS] += BOOT ? phi[u] * (double)m[u] : phi[u];
                        my_cnt[jb[u] * VP_COLS] += m[u];
                    }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // fold: warp w handles bins w, w + VP_WARPS, ...; lane l reads column l % 16 of warp-slices l / 16, l / 16 + 2, ...
    for (int jb = warp; jb < nb; jb += VP_WARPS) {
        double s = 0.0;
        unsigned long long c = 0ULL;
        for (int wv = lane >> 4; wv < VP_WARPS; wv += 2) {
            s += s_sum[(wv * nb + jb) * VP_COLS + (lane & (VP_COLS - 1))];
            c += s_cnt[(wv * nb + jb) * VP_COLS + (lane & (VP_COLS - 1))];
        }
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            a.sumphi[(long long)blockIdx.x * nb + jb] = s;
            a.counts[(long long)blockIdx.x * nb + jb] = c;
        }
    }
}

// one warp per bin: lanes stride over the block partials, fixed shuffle tree (deterministic)
__global__ void k_veff_reduce(int nblocks, int nbins, const unsigned long long* __restrict__ counts,
                              const double* __restrict__ sumphi, long long* __restrict__ out_c, double* __restrict__ out_s) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= nbins) return;
    double s = 0.0; unsigned long long c = 0ULL;
    for (int b = lane; b < nblocks; b += 32) { s += sumphi[(long long)b * nbins + j]; c += counts[(long long)b * nbins + j]; }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { out_c[j] = (long long)c; out_s[j] = s; }
}

// ------------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmark: 8 independent FMA chains per thread, registers only
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* sink) {
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

// MUFU (SFU) micro-benchmark: 8 independent ex2 chains per thread
__global__ void __launch_bounds__(256) k_mufu_peak(int iters, float seed, float* sink) {
    float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + 0.4f, a5 = a0 + 0.5f, a6 = a0 + 0.6f, a7 = a0 + 0.7f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = mufu_ex2(a0); a1 = mufu_ex2(a1); a2 = mufu_ex2(a2); a3 = mufu_ex2(a3);
            a4 = mufu_ex2(a4); a5 = mufu_ex2(a5); a6 = mufu_ex2(a6); a7 = mufu_ex2(a7);
            a0 -= 1.0f; a1 -= 1.0f; a2 -= 1.0f; a3 -= 1.0f; a4 -= 1.0f; a5 -= 1.0f; a6 -= 1.0f; a7 -= 1.0f;
        }
    }
    float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678f) sink[0] = s;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// peer-memory exchange (definitions used by lf_ctx; kernels further down)
#define PEER_MAX 16
#define PEER_CHUNK 256                 // walkers per block = per flag
struct PeerArgs {
    int rank, world;
    long long wcap;                    // capacity of one slot (doubles)
    int nchunk_cap;
    double* data[PEER_MAX];            // rank r's receive buffer: data[r][parity][sender][wcap]
    unsigned* flags[PEER_MAX];         // flags[r][parity][sender][nchunk_cap]
    const unsigned* seq;               // device counter of this rank: sequence number of the current exchange (starts at 1)
    int* timed_out;
};

struct lf_ctx {
    lf_config cfg;
    int device = 0, sm_count = 148;
    int occ_fast = 2, occ_lit = 2;     // resident blocks per SM of the persistent main kernels
    int ndim = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    KArgs ka;
    // resident
    long long N = 0, NQ = 0;
    double* d_lum = nullptr; double* d_flux = nullptr; double* d_z = nullptr; double* d_om = nullptr;
    double* d_Lsrc = nullptr; double2* d_src2 = nullptr; float2* d_src2f = nullptr; double2* d_csrc = nullptr;
    QuadPointFree* d_qpf = nullptr; QuadPoint* d_qp = nullptr; double* d_zarr = nullptr;
    Tables* d_tables = nullptr;
    bool have_sources = false, have_grid = false;
    // per-call scratch (grown on demand)
    long long Wcap = 0; int rows_cap = 0;
    double* d_wp = nullptr; double* d_colA = nullptr; double* d_colB = nullptr; double* d_partial = nullptr;
    int* d_cls = nullptr; int* d_list_fast = nullptr; int* d_list_lit = nullptr; int* d_list_fastq = nullptr; int* d_list_litq = nullptr;
    double* d_thetas = nullptr; double* d_out = nullptr;
    double* h_thetas = nullptr; double* h_out = nullptr;           // pinned staging
    long long launches = 0; double last_ms = 0.0, sampler_ms = 0.0;
    int h_cls[3] = {0, 0, 0};
    // Veff residency
    long long vN = 0; double* v_lum = nullptr; double* v_phi = nullptr; double* v_edges = nullptr; int v_nbins = 0;
    unsigned long long* v_counts = nullptr; double* v_sums = nullptr; long long* v_outc = nullptr; double* v_outs = nullptr;
    int* v_mult = nullptr; short* v_bin = nullptr; int v_blocks = 0;
    // peer exchange
    unsigned char* peer_base = nullptr; unsigned* peer_seq = nullptr; int* peer_timeout = nullptr; int* peer_timeout_h = nullptr;
    size_t peer_data_bytes = 0; bool peer_connected = false; void* peer_opened[PEER_MAX] = {};
    PeerArgs peer;
};

static int ndim_of(const lf_config& c) {
    int free_al = c.fix_sch_al ? 0 : 1;
    if (c.model == LF_MODEL_FREE) return 2 + free_al + c.nfields + 1;
    if (c.model == LF_MODEL_FIXED) return 2 + free_al;
    return 6 + free_al;
}

extern "C" int lf_ndim(const lf_ctx* ctx) { return ctx ? ctx->ndim : -1; }

static void fill_tables(Tables& t) {
    for (int j = 0; j < EXP_TAB_N; ++j) t.exp2_frac[j] = (double)exp2l((long double)j / EXP_TAB_N);
    for (int i = 0; i < EXPB_N; ++i) t.exp2_big[i] = (double)exp2l((long double)(EXPB_KMIN + i) / EXP_TAB_N);
    t.exp2_big[EXPB_N] = 1.0;
    const int M = 1 << LOG_MANT_BITS;
    for (int b = 0; b < LOG_OCTAVES * M; ++b) {
        int E = -LOG_OCTAVES + b / M, j = b % M;
        long double cm = 1.0L + ((long double)j + 0.5L) / M;          // bin centre of the mantissa
        double invc = ldexp((double)(1.0L / cm), -E);                    // 1 / (cm * 2^E), power-of-two scaling is exact
        t.log_tab[b].x = invc;
        t.log_tab[b].y = (double)(-logl((long double)invc) + (long double)LOG1P_C0);   // consistent with the rounded 1/c; + fit constant
    }
    t.log_tab[LOG_OCTAVES * M].x = 1.0;  t.log_tab[LOG_OCTAVES * M].y = 0.0;       // argument exactly 1
    t.log_tab[LOG_OCTAVES * M + 1] = t.log_tab[LOG_OCTAVES * M];
}

extern "C" int lf_create(lf_ctx** out, const lf_config* cfg) {
    if (!out || !cfg) return fail("lf_create: null argument");
    if (cfg->nfields < 1 || cfg->nfields > LF_MAX_FIELDS) return fail("lf_create: nfields out of range");
    if (cfg->model < 0 || cfg->model > 2) return fail("lf_create: unknown model");
    if (cfg->precision != LF_PREC_F64 && cfg->precision != LF_PREC_F32) return fail("lf_create: unknown precision");
    if (cfg->size_ln < 2) return fail("lf_create: size_ln must be >= 2");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("lf_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("lf_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    lf_ctx* c = new lf_ctx();
    c->cfg = *cfg;
    c->device = cfg->device;
    c->ndim = ndim_of(*cfg);
    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    Tables t;
    fill_tables(t);
    CK(cudaMalloc(&c->d_tables, sizeof(Tables)));
    CK(cudaMemcpy(c->d_tables, &t, sizeof(Tables), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_cls, 8 * sizeof(int)));
    CK(cudaFuncSetAttribute(k_main<false, LF_MODEL_FREE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(LF_MODEL_FREE)));
    CK(cudaFuncSetAttribute(k_main<false, LF_MODEL_FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(LF_MODEL_FIXED)));
    CK(cudaFuncSetAttribute(k_main<false, LF_MODEL_Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(LF_MODEL_Z)));
    {
        const int atom_smem = (int)(sizeof(double) * (VEFF_MAX_BINS + 1) + (sizeof(double) + sizeof(unsigned long long)) * 8 * VEFF_MAX_BINS);
        CK(cudaFuncSetAttribute(k_veff<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, atom_smem));
        CK(cudaFuncSetAttribute(k_veff<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, atom_smem));
        CK(cudaFuncSetAttribute(k_veff<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, atom_smem));
    }
    CK(cudaFuncSetAttribute(k_veff_priv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VP_SMEM_MAX));
    CK(cudaFuncSetAttribute(k_veff_priv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VP_SMEM_MAX));
    CK(cudaFuncSetAttribute(k_veff_priv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VP_SMEM_MAX));
    if (cfg->model == LF_MODEL_FREE) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_fast, k_main<false, LF_MODEL_FREE>, 32 * main_warps(LF_MODEL_FREE), main_smem_bytes(LF_MODEL_FREE)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_lit, k_main<true, LF_MODEL_FREE>, 32 * main_warps(LF_MODEL_FREE), 0));
    } else if (cfg->model == LF_MODEL_FIXED) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_fast, k_main<false, LF_MODEL_FIXED>, 32 * main_warps(LF_MODEL_FIXED), main_smem_bytes(LF_MODEL_FIXED)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_lit, k_main<true, LF_MODEL_FIXED>, 32 * main_warps(LF_MODEL_FIXED), 0));
    } else {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_fast, k_main<false, LF_MODEL_Z>, 32 * main_warps(LF_MODEL_Z), main_smem_bytes(LF_MODEL_Z)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_lit, k_main<true, LF_MODEL_Z>, 32 * main_warps(LF_MODEL_Z), 0));
    }
    if (c->occ_fast < 1) c->occ_fast = 1;
    if (c->occ_lit < 1) c->occ_lit = 1;
    memset(&c->ka, 0, sizeof(KArgs));
    KArgs& a = c->ka;
    a.model = cfg->model; a.K = cfg->nfields; a.S = cfg->size_ln; a.fix_sch_al = cfg->fix_sch_al;
    a.fixed_prior_ok = cfg->fixed_prior_ok; a.force_literal = cfg->force_literal;
    a.modified = (cfg->fcmin != 0.0) ? 1 : 0;
    a.ndim = c->ndim; a.fcmin = cfg->fcmin; a.precision = cfg->precision;
    {
        double aa = (2.0 * cfg->fcmin - 1.0) * (2.0 * cfg->fcmin - 1.0);
        a.fcA2 = fabs(aa / (1.0 - aa));
    }
    a.sch_al = cfg->sch_al;
    for (int i = 0; i < 2; ++i) {
        a.Lstar_lims[i] = cfg->Lstar_lims[i]; a.phistar_lims[i] = cfg->phistar_lims[i];
        a.sch_al_lims[i] = cfg->sch_al_lims[i]; a.Flim_lims[i] = cfg->Flim_lims[i]; a.alpha_lims[i] = cfg->alpha_lims[i];
    }
    a.z1 = cfg->z_pivots[0]; a.z2 = cfg->z_pivots[1]; a.z3 = cfg->z_pivots[2];
    a.share = 0; a.nshare = 1; a.prior_gate = 1;
    a.fcap = 64.0e-17 * cfg->Flim_lims[1];
    a.tables = c->d_tables;
    *out = c;
    return 0;
}

struct DevBufs {                       // frees whatever was allocated when it goes out of scope
    std::vector<void*> p;
    ~DevBufs() { for (void* q : p) cudaFree(q); }
    template <typename T> cudaError_t alloc(T** out, size_t bytes) {
        cudaError_t e = cudaMalloc(out, bytes ? bytes : 8);
        if (e == cudaSuccess) p.push_back(*out);
        return e;
    }
};

template <typename T>
static void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

extern "C" void lf_destroy(lf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    dfree(c->d_lum); dfree(c->d_flux); dfree(c->d_z); dfree(c->d_om); dfree(c->d_Lsrc); dfree(c->d_src2); dfree(c->d_src2f); dfree(c->d_csrc);
    dfree(c->d_qpf); dfree(c->d_qp); dfree(c->d_zarr); dfree(c->d_tables);
    dfree(c->d_wp); dfree(c->d_colA); dfree(c->d_colB); dfree(c->d_partial);
    dfree(c->d_cls); dfree(c->d_list_fast); dfree(c->d_list_lit); dfree(c->d_list_fastq); dfree(c->d_list_litq); dfree(c->d_thetas); dfree(c->d_out);
    dfree(c->v_lum); dfree(c->v_phi); dfree(c->v_edges); dfree(c->v_counts); dfree(c->v_sums);
    dfree(c->v_outc); dfree(c->v_outs); dfree(c->v_mult); dfree(c->v_bin);
    for (int r = 0; r < PEER_MAX; ++r) if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peer_opened[r]);
    dfree(c->peer_base); dfree(c->peer_seq);
    if (c->peer_timeout_h) cudaFreeHost(c->peer_timeout_h);
    if (c->h_thetas) cudaFreeHost(c->h_thetas);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// sum / min / max of a device range, deterministic
static int range_stats(lf_ctx* c, const double* d_x, long long i0, long long i1, double* d_tmp, std::vector<double>& h_tmp,
                       double& sum, double& mn, double& mx) {
    sum = 0.0; mn = 1.0e300; mx = -1.0e300;
    if (i1 <= i0) return 0;
    int nb = (int)std::min<long long>(1024, (i1 - i0 + STAT_THREADS - 1) / STAT_THREADS);
    k_stats<<<nb, STAT_THREADS, 0, c->stream>>>(d_x, i0, i1, d_tmp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_tmp.data(), d_tmp, sizeof(double) * 3 * nb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    long double s = 0.0L;
    for (int b = 0; b < nb; ++b) {
        s += h_tmp[3 * b];
        mn = std::min(mn, h_tmp[3 * b + 1]);
        mx = std::max(mx, h_tmp[3 * b + 2]);
    }
    sum = (double)s;
    return 0;
}

extern "C" int lf_set_sources(lf_ctx* c, int64_t n, const double* lum, const double* flux, const double* z,
                              const double* om_arr, const int64_t* field_ind, const int64_t* omega0_int) {
    if (!c) return fail("lf_set_sources: null context");
    if (n < 0 || !field_ind) return fail("lf_set_sources: bad arguments");
    const int K = c->cfg.nfields, model = c->cfg.model;
    if (field_ind[0] != 0 || field_ind[K] != n) return fail("lf_set_sources: field_ind must run from 0 to n");
    for (int k = 0; k < K; ++k)
        if (field_ind[k + 1] < field_ind[k]) return fail("lf_set_sources: field_ind must be non-decreasing");
    if (n > 0 && !lum) return fail("lf_set_sources: lum is required");
    if (model == LF_MODEL_FREE && n > 0 && (!flux || !omega0_int)) return fail("lf_set_sources: FREE model needs flux and omega0_int");
    if (model != LF_MODEL_FREE && n > 0 && !om_arr) return fail("lf_set_sources: FIXED/Z models need om_arr");
    if (model == LF_MODEL_Z && n > 0 && !z) return fail("lf_set_sources: Z model needs z");
    CK(cudaSetDevice(c->device));
    dfree(c->d_lum); dfree(c->d_flux); dfree(c->d_z); dfree(c->d_om); dfree(c->d_Lsrc); dfree(c->d_src2); dfree(c->d_src2f);
    c->N = n;
    KArgs& a = c->ka;
    a.N = n;
    dfree(c->d_csrc); a.csrc = nullptr; a.M = 0;              // a new catalogue invalidates its compressed form
    for (int k = 0; k <= K; ++k) a.field_ind[k] = field_ind[k];
    for (int k = 0; k < K; ++k) {
        FieldStats& s = a.fs[k];
        double gq = s.grid_g_min, fq = s.grid_f_min;            // keep grid ranges if the grid came first
        memset(&s, 0, sizeof(FieldStats));
        s.grid_g_min = c->have_grid ? gq : 1.0e300;
        s.grid_f_min = c->have_grid ? fq : 1.0e300;
        s.n = (double)(field_ind[k + 1] - field_ind[k]);
        if (model == LF_MODEL_FREE && omega0_int) {
            s.om0_over_sq = (double)omega0_int[k] / SQARCSEC;
            s.ln_om0 = log(s.om0_over_sq);
        }
    }
    a.lum_max_all = -1.0e300;
    const size_t nb = sizeof(double) * (size_t)std::max<long long>(n, 1);
    CK(cudaMalloc(&c->d_lum, nb));
    CK(cudaMalloc(&c->d_Lsrc, nb));
    double* d_tmp = nullptr; double* d_scratch = nullptr;
    CK(cudaMalloc(&d_tmp, sizeof(double) * 3 * 1024));
    CK(cudaMalloc(&d_scratch, nb));
    std::vector<double> h_tmp(3 * 1024);
    const int T = 256;
    const unsigned G = (unsigned)((n + T - 1) / T);
    if (n > 0) {
        CK(cudaMemcpyAsync(c->d_lum, lum, nb, cudaMemcpyHostToDevice, c->stream));
        if (model == LF_MODEL_FREE) {
            CK(cudaMalloc(&c->d_flux, nb));
            CK(cudaMalloc(&c->d_src2, sizeof(double2) * (size_t)n));
            if (c->cfg.precision == LF_PREC_F32) CK(cudaMalloc(&c->d_src2f, sizeof(float2) * (size_t)n));
            CK(cudaMemcpyAsync(c->d_flux, flux, nb, cudaMemcpyHostToDevice, c->stream));
            k_derive_free<<<G, T, 0, c->stream>>>(n, c->d_lum, c->d_flux, c->d_src2, c->d_Lsrc, c->ka.fcap, c->d_src2f);
        } else {
            CK(cudaMalloc(&c->d_om, nb));
            CK(cudaMemcpyAsync(c->d_om, om_arr, nb, cudaMemcpyHostToDevice, c->stream));
            k_pow10<<<G, T, 0, c->stream>>>(n, c->d_lum, c->d_Lsrc);
            if (model == LF_MODEL_Z) {
                CK(cudaMalloc(&c->d_z, nb));
                CK(cudaMalloc(&c->d_src2, sizeof(double2) * (size_t)n));
                if (c->cfg.precision == LF_PREC_F32) CK(cudaMalloc(&c->d_src2f, sizeof(float2) * (size_t)n));
                CK(cudaMemcpyAsync(c->d_z, z, nb, cudaMemcpyHostToDevice, c->stream));
                k_derive_z<<<G, T, 0, c->stream>>>(n, c->d_lum, c->d_z, c->d_src2, c->d_src2f, c->cfg.z_pivots[1]);
            }
        }
        CK(cudaGetLastError());
        bool took = false;
        for (int k = 0; k < K; ++k) {
            FieldStats& s = a.fs[k];
            long long i0 = field_ind[k], i1 = field_ind[k + 1];
            if (i1 <= i0) continue;
            double sm, mn, mx;
            if (range_stats(c, c->d_lum, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
            s.sum_lum = sm; s.lum_min = mn; s.lum_max = mx;
            a.lum_max_all = std::max(a.lum_max_all, mx);
            if (range_stats(c, c->d_Lsrc, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
            s.sum_L = sm;
            if (model == LF_MODEL_FREE) {
                if (!took) { k_take<<<G, T, 0, c->stream>>>(n, c->d_src2, 0, d_scratch); took = true; }
                if (range_stats(c, d_scratch, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
                s.g_min = mn;
                if (range_stats(c, c->d_flux, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
                s.f_min = mn;
            } else {
                k_log<<<G, T, 0, c->stream>>>(n, c->d_om, d_scratch);
                if (range_stats(c, d_scratch, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
                s.sum_lnom = sm; s.lnom_min = mn;
                if (!(mn == mn) || !(sm == sm)) { s.lnom_min = -1.0e300; s.sum_lnom = 0.0; }   // Om_arr <= 0: literal only
                if (model == LF_MODEL_Z) {
                    if (range_stats(c, c->d_z, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
                    s.sum_z = sm; s.z_min = mn; s.z_max = mx;
                    k_square<<<G, T, 0, c->stream>>>(n, c->d_z, d_scratch);
                    if (range_stats(c, d_scratch, i0, i1, d_tmp, h_tmp, sm, mn, mx)) return 1;
                    s.sum_z2 = sm;
                }
            }
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_tmp);
    cudaFree(d_scratch);
    a.src2 = c->d_src2; a.src2f = c->d_src2f; a.lum = c->d_lum; a.flux = c->d_flux; a.z = c->d_z; a.om_arr = c->d_om;
    c->have_sources = true;
    return 0;
}

extern "C" int lf_set_grid(lf_ctx* c, const double* logL, const double* zarr, const double* DL_zarr,
                           const double* volume_part, const double* integ_part, const double* omega0) {
    if (!c) return fail("lf_set_grid: null context");
    const int K = c->cfg.nfields, S = c->cfg.size_ln, model = c->cfg.model;
    if (!logL || !zarr) return fail("lf_set_grid: logL and zarr are required");
    if (model == LF_MODEL_FREE && (!DL_zarr || !volume_part || !omega0)) return fail("lf_set_grid: FREE model needs DL_zarr, volume_part, omega0");
    if (model != LF_MODEL_FREE && !integ_part) return fail("lf_set_grid: FIXED/Z models need integ_part");
    CK(cudaSetDevice(c->device));
    dfree(c->d_qpf); dfree(c->d_qp); dfree(c->d_zarr);
    const long long SS = (long long)S * S, NQ = SS * K;
    c->NQ = NQ;
    KArgs& a = c->ka;
    a.NQ = NQ;
    // trapezoid weights: trapz(trapz(y, logL, axis=0), zarr) = sum_ji wl_ji wz_i y_ji     (lumfuncmcmc.py:377)
    std::vector<double> wz(S);
    for (int i = 0; i < S; ++i) {
        double wgt = 0.0;
        if (i + 1 < S) wgt += zarr[i + 1] - zarr[i];
        if (i > 0) wgt += zarr[i] - zarr[i - 1];
        wz[i] = 0.5 * wgt;
    }
    auto at = [&](const double* arr, int k, int j, int i) { return arr[((long long)k * S + j) * S + i]; };
    // points are stored column-major: q = (k*S + i)*S + j
    if (model == LF_MODEL_FREE) {
        std::vector<QuadPointFree> pts((size_t)NQ);
        for (int k = 0; k < K; ++k) {
            double gmin = 1.0e300, fmin_ = 1.0e300;
            for (int i = 0; i < S; ++i) {
                double dl = MPC_CM_REF * DL_zarr[i];
                double den = FOURPI * (dl * dl);
                for (int j = 0; j < S; ++j) {
                    double x = at(logL, k, j, i);
                    double wl = 0.0;
                    if (j + 1 < S) wl += at(logL, k, j + 1, i) - x;
                    if (j > 0) wl += x - at(logL, k, j - 1, i);
                    wl *= 0.5;
                    QuadPointFree& p = pts[((size_t)k * S + i) * S + j];
                    p.x = x;
                    p.Lx = pow(10.0, x);
                    p.f = p.Lx / den;                                   // lumfuncmcmc.py:69-70
                    p.g = log10(p.f);
                    p.ftrue = p.f;
                    p.f = std::min(p.f, a.fcap);                        // decay-argument copy, see k_derive_free
                    p.wt = wl * wz[i] * volume_part[i] * (omega0[k] / SQARCSEC);
                    gmin = std::min(gmin, p.g);
                    fmin_ = std::min(fmin_, p.ftrue);
                }
            }
            a.fs[k].grid_g_min = gmin;
            a.fs[k].grid_f_min = fmin_;
        }
        CK(cudaMalloc(&c->d_qpf, sizeof(QuadPointFree) * (size_t)NQ));
        CK(cudaMemcpy(c->d_qpf, pts.data(), sizeof(QuadPointFree) * (size_t)NQ, cudaMemcpyHostToDevice));
    } else {
        std::vector<QuadPoint> pts((size_t)NQ);
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < S; ++i)
                for (int j = 0; j < S; ++j) {
                    double x = at(logL, k, j, i);
                    double wl = 0.0;
                    if (j + 1 < S) wl += at(logL, k, j + 1, i) - x;
                    if (j > 0) wl += x - at(logL, k, j - 1, i);
                    wl *= 0.5;
                    QuadPoint& p = pts[((size_t)k * S + i) * S + j];
                    p.x = x;
                    p.Lx = pow(10.0, x);
                    p.wt = wl * wz[i] * at(integ_part, k, j, i);
                    p.pad = 0.0;
                }
        CK(cudaMalloc(&c->d_qp, sizeof(QuadPoint) * (size_t)NQ));
        CK(cudaMemcpy(c->d_qp, pts.data(), sizeof(QuadPoint) * (size_t)NQ, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&c->d_zarr, sizeof(double) * S));
    CK(cudaMemcpy(c->d_zarr, zarr, sizeof(double) * S, cudaMemcpyHostToDevice));
    a.qpf = c->d_qpf; a.qp = c->d_qp; a.zarr = c->d_zarr;
    c->have_grid = true;
    return 0;
}

__global__ void k_derive_compressed(long long m, const double* __restrict__ xi, const double* __restrict__ w, double fcap,
                                    double2* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double g = xi[i];
    out[2 * i] = make_double2(g, fmin(pow(10.0, g), fcap));
    out[2 * i + 1] = make_double2(w[i], 0.0);
}

extern "C" int lf_set_compressed_sources(lf_ctx* c, int64_t M, const double* xi, const double* w, const int64_t* cfield_ind) {
    if (!c) return fail("lf_set_compressed_sources: null context");
    if (c->cfg.model != LF_MODEL_FREE) return fail("lf_set_compressed_sources: only the free-completeness model has a compressed form");
    if (c->cfg.precision != LF_PREC_F64) return fail("lf_set_compressed_sources: FP64 only");
    if (!c->have_sources) return fail("lf_set_compressed_sources: call lf_set_sources first");
    CK(cudaSetDevice(c->device));
    dfree(c->d_csrc);
    c->ka.csrc = nullptr; c->ka.M = 0;
    if (M == 0) return 0;
    if (M < 0 || !xi || !w || !cfield_ind) return fail("lf_set_compressed_sources: bad arguments");
    const int K = c->cfg.nfields;
    if (cfield_ind[0] != 0 || cfield_ind[K] != M) return fail("lf_set_compressed_sources: cfield_ind must run from 0 to M");
    for (int k = 0; k < K; ++k) {
        if (cfield_ind[k + 1] < cfield_ind[k]) return fail("lf_set_compressed_sources: cfield_ind must be non-decreasing");
        if ((cfield_ind[k + 1] > cfield_ind[k]) != (c->ka.fs[k].n > 0.0)) return fail("lf_set_compressed_sources: a field has sources but no pseudo-sources (or the reverse)");
    }
    // every node must lie inside the range of fluxes the classification bounds were computed for
    for (int k = 0; k < K; ++k)
        for (int64_t m = cfield_ind[k]; m < cfield_ind[k + 1]; ++m)
            if (!(xi[m] >= c->ka.fs[k].g_min - 1.0e-9) || !(xi[m] == xi[m])) return fail("lf_set_compressed_sources: a node lies below the field's faintest source");
    double *d_xi = nullptr, *d_w = nullptr;
    DevBufs tmp;
    CK(tmp.alloc(&d_xi, sizeof(double) * M));
    CK(tmp.alloc(&d_w, sizeof(double) * M));
    CK(cudaMalloc(&c->d_csrc, sizeof(double2) * 2 * (size_t)M));
    CK(cudaMemcpy(d_xi, xi, sizeof(double) * M, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w, sizeof(double) * M, cudaMemcpyHostToDevice));
    k_derive_compressed<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(M, d_xi, d_w, c->ka.fcap, c->d_csrc);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->ka.csrc = c->d_csrc; c->ka.M = M;
    for (int k = 0; k <= K; ++k) c->ka.cfield_ind[k] = cfield_ind[k];
    return 0;
}

extern "C" int lf_set_prior_gate(lf_ctx* c, int32_t enabled) {
    if (!c) return fail("lf_set_prior_gate: null context");
    c->ka.prior_gate = enabled ? 1 : 0;
    return 0;
}

extern "C" int lf_set_quadrature_share(lf_ctx* c, int32_t share, int32_t nshare) {
    if (!c) return fail("lf_set_quadrature_share: null context");
    if (nshare < 1 || share < 0 || share >= nshare) return fail("lf_set_quadrature_share: need 0 <= share < nshare");
    c->ka.share = share;
    c->ka.nshare = nshare;
    return 0;
}

static int ensure_scratch(lf_ctx* c, long long W, int rows) {
    if (W > c->Wcap) {
        long long cap = std::max<long long>(64, W);
        cap = (cap + 31) / 32 * 32;
        dfree(c->d_wp); dfree(c->d_colA); dfree(c->d_colB); dfree(c->d_partial);
        dfree(c->d_list_fast); dfree(c->d_list_lit); dfree(c->d_list_fastq); dfree(c->d_list_litq); dfree(c->d_thetas); dfree(c->d_out);
        if (c->h_thetas) { cudaFreeHost(c->h_thetas); c->h_thetas = nullptr; }
        if (c->h_out) { cudaFreeHost(c->h_out); c->h_out = nullptr; }
        CK(cudaMalloc(&c->d_wp, sizeof(double) * P_NSLOTS * cap));
        if (c->cfg.model == LF_MODEL_Z) {
            CK(cudaMalloc(&c->d_colA, sizeof(double) * c->cfg.size_ln * cap));
            CK(cudaMalloc(&c->d_colB, sizeof(double) * c->cfg.size_ln * cap));
        }
        CK(cudaMalloc(&c->d_list_fast, sizeof(int) * cap));
        CK(cudaMalloc(&c->d_list_lit, sizeof(int) * cap));
        CK(cudaMalloc(&c->d_list_fastq, sizeof(int) * cap));
        CK(cudaMalloc(&c->d_list_litq, sizeof(int) * cap));
        CK(cudaMalloc(&c->d_thetas, sizeof(double) * c->ndim * cap));
        CK(cudaMalloc(&c->d_out, sizeof(double) * cap));
        CK(cudaMallocHost(&c->h_thetas, sizeof(double) * c->ndim * cap));
        CK(cudaMallocHost(&c->h_out, sizeof(double) * cap));
        c->Wcap = cap;
        c->rows_cap = 0;
    }
    if (rows > c->rows_cap) {
        dfree(c->d_partial);
        CK(cudaMalloc(&c->d_partial, sizeof(double) * (size_t)rows * c->Wcap));
        c->rows_cap = rows;
    }
    return 0;
}

// choose slab counts so that one class fills the machine with a few waves of warp items
static void plan_rows(const lf_ctx* c, long long W, int& n_src, int& n_quad) {
    const long long n_wg = (W + 31) / 32;
    const long long target_items = (long long)c->sm_count * 16 * 24;   // ~24-32 items per resident warp
    long long rows = std::min<long long>(4096, std::max<long long>(1, target_items / n_wg));
    const int model = c->cfg.model;
    // relative cost of a quadrature point vs a source term
    double src_cost = model == LF_MODEL_FREE ? 1.0 : (model == LF_MODEL_Z ? 0.5 : 0.0);
    double quad_cost = model == LF_MODEL_FREE ? 1.5 : 0.6;
    const long long n_eff = (c->d_csrc && model == LF_MODEL_FREE) ? c->ka.M : c->N;      // pseudo-sources when compressed
    double wsrc = src_cost * (double)n_eff, wq = quad_cost * (double)c->NQ;
    double tot = wsrc + wq;
    if (tot <= 0.0) { n_src = 1; n_quad = 1; return; }
    long long rs = (long long)llround((double)rows * wsrc / tot), rq = rows - rs;
    const long long min_per = 64;       // at least this many sources / points per work item
    rs = std::max<long long>(1, std::min<long long>(rs, std::max<long long>(1, n_eff / min_per)));
    rq = std::max<long long>(1, std::min<long long>(rq, std::max<long long>(1, c->NQ / min_per)));
    n_src = (int)rs;
    n_quad = (int)rq;
}

static int launch_pipeline(lf_ctx* c, const double* d_thetas, long long W, double* d_out, cudaStream_t st) {
    if (!c->have_sources || !c->have_grid) return fail("lf_lnprob: call lf_set_sources and lf_set_grid first");
    if (W <= 0) return 0;
    int n_src, n_quad;
    plan_rows(c, W, n_src, n_quad);
    if (ensure_scratch(c, W, n_src + n_quad)) return 1;
    KArgs a = c->ka;
    a.thetas = d_thetas; a.out = d_out; a.W = W; a.Wcap = c->Wcap;
    a.wp = c->d_wp; a.colA = c->d_colA; a.colB = c->d_colB;
    a.cls_count = c->d_cls; a.list_fast = c->d_list_fast; a.list_lit = c->d_list_lit;
    a.list_fastq = c->d_list_fastq; a.list_litq = c->d_list_litq;
    a.partial = c->d_partial; a.n_src_slabs = n_src; a.n_quad_slabs = n_quad;
    CK(cudaMemsetAsync(c->d_cls, 0, 8 * sizeof(int), st));
    k_prologue<<<(unsigned)((W + PRO_WALKERS - 1) / PRO_WALKERS), dim3(PRO_WALKERS, c->cfg.nfields), 0, st>>>(a);
    c->launches++;
    if (c->cfg.model == LF_MODEL_Z) {
        long long tot = (long long)c->cfg.size_ln * W;
        k_zcolumns<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a);
        c->launches++;
    }
    const long long n_wg = (W + 31) / 32;
    const long long items = n_wg * (n_src + n_quad);
    const int wpb = main_warps(c->cfg.model);
    const long long need = (items + wpb - 1) / wpb;      // never more blocks than items
    const unsigned bf = (unsigned)std::min<long long>(need, (long long)c->sm_count * c->occ_fast);
    const unsigned bl = (unsigned)std::min<long long>(need, (long long)c->sm_count * c->occ_lit);
    if (c->cfg.model == LF_MODEL_FREE) {
        k_main<false, LF_MODEL_FREE><<<bf, 32 * main_warps(LF_MODEL_FREE), main_smem_bytes(LF_MODEL_FREE), st>>>(a);
        k_main<true, LF_MODEL_FREE><<<bl, 32 * main_warps(LF_MODEL_FREE), 0, st>>>(a);
    } else if (c->cfg.model == LF_MODEL_FIXED) {
        k_main<false, LF_MODEL_FIXED><<<bf, 32 * main_warps(LF_MODEL_FIXED), main_smem_bytes(LF_MODEL_FIXED), st>>>(a);
        k_main<true, LF_MODEL_FIXED><<<bl, 32 * main_warps(LF_MODEL_FIXED), 0, st>>>(a);
    } else {
        k_main<false, LF_MODEL_Z><<<bf, 32 * main_warps(LF_MODEL_Z), main_smem_bytes(LF_MODEL_Z), st>>>(a);
        k_main<true, LF_MODEL_Z><<<bl, 32 * main_warps(LF_MODEL_Z), 0, st>>>(a);
    }
    k_finish<<<(unsigned)((W + 31) / 32), 32 * FIN_GROUPS, 0, st>>>(a);
    c->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lf_lnprob_batch_device(lf_ctx* c, const double* d_thetas, int64_t W, double* d_out, void* stream) {
    if (!c) return fail("lf_lnprob_batch_device: null context");
    if (W < 0 || (W > 0 && (!d_thetas || !d_out))) return fail("lf_lnprob_batch_device: bad arguments");
    CK(cudaSetDevice(c->device));
    return launch_pipeline(c, d_thetas, W, d_out, (cudaStream_t)stream);
}

extern "C" int lf_lnprob_batch(lf_ctx* c, const double* thetas, int64_t W, double* out) {
    if (!c) return fail("lf_lnprob_batch: null context");
    if (W < 0 || (W > 0 && (!thetas || !out))) return fail("lf_lnprob_batch: bad arguments");
    if (W == 0) return 0;
    CK(cudaSetDevice(c->device));
    if (!c->have_sources || !c->have_grid) return fail("lf_lnprob: call lf_set_sources and lf_set_grid first");
    int n_src, n_quad;
    plan_rows(c, W, n_src, n_quad);
    if (ensure_scratch(c, W, n_src + n_quad)) return 1;
    memcpy(c->h_thetas, thetas, sizeof(double) * c->ndim * W);
    CK(cudaMemcpyAsync(c->d_thetas, c->h_thetas, sizeof(double) * c->ndim * W, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    if (launch_pipeline(c, c->d_thetas, W, c->d_out, c->stream)) return 1;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(c->h_out, c->d_out, sizeof(double) * W, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->h_cls, c->d_cls, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_ms = ms;
    memcpy(out, c->h_out, sizeof(double) * W);
    return 0;
}

extern "C" int lf_last_call_info(lf_ctx* c, int64_t counts[3], int64_t* launches) {
    if (!c) return fail("lf_last_call_info: null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(c->h_cls, c->d_cls, 3 * sizeof(int), cudaMemcpyDeviceToHost));
    if (counts) for (int i = 0; i < 3; ++i) counts[i] = c->h_cls[i];
    if (launches) *launches = c->launches;
    return 0;
}

extern "C" int lf_last_kernel_ms(lf_ctx* c, double* ms) {
    if (!c || !ms) return fail("lf_last_kernel_ms: null argument");
    *ms = c->last_ms;
    return 0;
}

extern "C" int lf_fp64_peak(lf_ctx* c, int32_t iters, double* dfma_per_s, double* ms_out) {
    if (!c || !dfma_per_s) return fail("lf_fp64_peak: null argument");
    CK(cudaSetDevice(c->device));
    double* sink = nullptr;
    CK(cudaMalloc(&sink, sizeof(double)));
    const int blocks = c->sm_count * 8, threads = 256;
    k_fp64_peak<<<blocks, threads, 0, c->stream>>>(64, 1.0, sink);       // warm-up
    CK(cudaEventRecord(c->ev0, c->stream));
    k_fp64_peak<<<blocks, threads, 0, c->stream>>>(iters, 1.0, sink);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    cudaFree(sink);
    c->launches += 2;
    double n = (double)blocks * threads * (double)iters * 64.0;
    *dfma_per_s = n / (ms * 1.0e-3);
    if (ms_out) *ms_out = ms;
    return 0;
}

extern "C" int lf_mufu_peak(lf_ctx* c, int32_t iters, double* mufu_per_s, double* ms_out) {
    if (!c || !mufu_per_s) return fail("lf_mufu_peak: null argument");
    CK(cudaSetDevice(c->device));
    float* sink = nullptr;
    CK(cudaMalloc(&sink, sizeof(float)));
    const int blocks = c->sm_count * 8, threads = 256;
    k_mufu_peak<<<blocks, threads, 0, c->stream>>>(64, 0.25f, sink);       // warm-up
    CK(cudaEventRecord(c->ev0, c->stream));
    k_mufu_peak<<<blocks, threads, 0, c->stream>>>(iters, 0.25f, sink);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    cudaFree(sink);
    c->launches += 2;
    *mufu_per_s = (double)blocks * threads * (double)iters * 64.0 / (ms * 1.0e-3);
    if (ms_out) *ms_out = ms;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// peer-memory all-reduce of the per-walker partials (one process per GPU, NVLink / NVSwitch)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(PEER_CHUNK) k_allreduce_p2p(PeerArgs a, double* __restrict__ vec, long long W) {
    const unsigned seq = *a.seq;
    const int par = (int)(seq & 1u);
    const int chunk = blockIdx.x;
    const long long w = (long long)chunk * PEER_CHUNK + threadIdx.x;
    const size_t slot = ((size_t)par * a.world + a.rank) * (size_t)a.wcap;
    // 1. push my values into my slot of every rank's buffer (coalesced 8-byte stores over NVLink; own buffer included)
    if (w < W) {
        const double v = vec[w];
        for (int r = 0; r < a.world; ++r) a.data[r][slot + w] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag for this chunk on every rank
    if (threadIdx.x < a.world)
        st_release_sys(a.flags[threadIdx.x] + ((size_t)par * a.world + a.rank) * a.nchunk_cap + chunk, seq);
    // 3. wait for every sender's flag on my own buffer (bounded spin: a dead peer must not hang the GPU)
    if (threadIdx.x < a.world) {
        const unsigned* f = a.flags[a.rank] + ((size_t)par * a.world + threadIdx.x) * a.nchunk_cap + chunk;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - seq) < 0) {
            if (clock64() - t0 > 8000000000LL) { atomicExch(a.timed_out, 1); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    // 4. add the slots in rank order: the same sum, bit for bit, on every rank
    if (w < W) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) s += a.data[a.rank][((size_t)par * a.world + r) * (size_t)a.wcap + w];
        vec[w] = s;
    }
}
__global__ void k_seq_advance(unsigned* seq) { *seq += 1u; }

extern "C" int lf_peer_buffer_create(lf_ctx* c, int32_t rank, int32_t world, int64_t wcap, unsigned char handle_out[64]) {
    if (!c || !handle_out) return fail("lf_peer_buffer_create: null argument");
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || wcap < 1) return fail("lf_peer_buffer_create: bad rank / world / capacity");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CK(cudaSetDevice(c->device));
    if (c->peer_base) return fail("lf_peer_buffer_create: already created");
    const long long cap = (wcap + PEER_CHUNK - 1) / PEER_CHUNK * PEER_CHUNK;
    const int nchunk = (int)(cap / PEER_CHUNK);
    const size_t data_bytes = sizeof(double) * 2 * (size_t)world * cap;
    const size_t flag_bytes = sizeof(unsigned) * 2 * (size_t)world * nchunk;
    CK(cudaMalloc(&c->peer_base, data_bytes + flag_bytes));
    CK(cudaMemset(c->peer_base, 0, data_bytes + flag_bytes));
    CK(cudaMalloc(&c->peer_seq, sizeof(unsigned)));
    const unsigned one = 1u;
    CK(cudaMemcpy(c->peer_seq, &one, sizeof(unsigned), cudaMemcpyHostToDevice));
    // time-out flag in mapped pinned host memory: the kernel writes it (zero-copy) only when a wait expires, the host
    // reads it without a device round trip
    CK(cudaHostAlloc(&c->peer_timeout_h, sizeof(int), cudaHostAllocMapped));
    *c->peer_timeout_h = 0;
    CK(cudaHostGetDevicePointer(&c->peer_timeout, c->peer_timeout_h, 0));
    PeerArgs& p = c->peer;
    memset(&p, 0, sizeof(p));
    p.rank = rank; p.world = world; p.wcap = cap; p.nchunk_cap = nchunk;
    p.data[rank] = reinterpret_cast<double*>(c->peer_base);
    p.flags[rank] = reinterpret_cast<unsigned*>(c->peer_base + data_bytes);
    p.seq = c->peer_seq; p.timed_out = c->peer_timeout;
    c->peer_data_bytes = data_bytes;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->peer_base));
    memcpy(handle_out, &h, 64);
    c->peer_connected = (world == 1);
    return 0;
}

extern "C" int lf_peer_buffer_connect(lf_ctx* c, const unsigned char* handles) {
    if (!c || !handles) return fail("lf_peer_buffer_connect: null argument");
    if (!c->peer_base) return fail("lf_peer_buffer_connect: call lf_peer_buffer_create first");
    CK(cudaSetDevice(c->device));
    PeerArgs& p = c->peer;
    for (int r = 0; r < p.world; ++r) {
        if (r == p.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * r, 64);
        void* base = nullptr;
        CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[r] = base;
        p.data[r] = reinterpret_cast<double*>(base);
        p.flags[r] = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(base) + c->peer_data_bytes);
    }
    c->peer_connected = true;
    return 0;
}

extern "C" int lf_allreduce_device(lf_ctx* c, double* d_vec, int64_t W, void* stream) {
    if (!c || (W > 0 && !d_vec)) return fail("lf_allreduce_device: null argument");
    if (!c->peer_base || !c->peer_connected) return fail("lf_allreduce_device: peer buffers are not connected");
    if (W > c->peer.wcap) return fail("lf_allreduce_device: vector longer than the peer buffer capacity");
    if (W <= 0) return 0;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    k_allreduce_p2p<<<(unsigned)((W + PEER_CHUNK - 1) / PEER_CHUNK), PEER_CHUNK, 0, st>>>(c->peer, d_vec, W);
    k_seq_advance<<<1, 1, 0, st>>>(c->peer_seq);
    c->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lf_peer_status(lf_ctx* c, int32_t* timed_out) {
    if (!c || !timed_out) return fail("lf_peer_status: null argument");
    *timed_out = 0;
    if (!c->peer_timeout_h) return 0;
    *timed_out = *(volatile int*)c->peer_timeout_h;        // meaningful after the stream that ran the exchange was synchronised
    *c->peer_timeout_h = 0;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// set-up tables on the GPU: cosmology distances and NumPy-exact linear interpolation
// ------------------------------------------------------------------------------------------------
// E(z) with NumPy's order of operations and no fused multiply-adds (cosmology.py efunc)
__device__ __forceinline__ double efunc_np(const lf_cosmology& c, double z) {
    const double zp1 = __dadd_rn(1.0, z);
    double t = __dadd_rn(__dmul_rn(c.Or0, zp1), c.Om0);
    t = __dadd_rn(__dmul_rn(t, zp1), c.Ok0);
    t = __dadd_rn(__dmul_rn(__dmul_rn(zp1, zp1), t), c.Ode0);
    return sqrt(t);
}

__global__ void k_cosmo(lf_cosmology c, const double* __restrict__ cum, long long ncum, long long n,
                        const double* __restrict__ z, const double* __restrict__ glx, const double* __restrict__ glw,
                        double* __restrict__ DL, double* __restrict__ dV, int* __restrict__ bad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    const long long p = (long long)floor(__ddiv_rn(zi, c.panel));
    if (!(zi >= 0.0) || p >= ncum) { atomicExch(bad, 1); return; }
    const double lo = __dmul_rn((double)p, c.panel);
    const double half = __dmul_rn(0.5, __dsub_rn(zi, lo));
    double acc = 0.0;
    for (int q = 0; q < 8; ++q) {                                   // acc += w / E(lo + half * (1 + x)), in node order
        const double node = __dadd_rn(lo, __dmul_rn(half, __dadd_rn(1.0, glx[q])));
        acc = __dadd_rn(acc, __ddiv_rn(glw[q], efunc_np(c, node)));
    }
    const double dc = __dadd_rn(cum[p], __dmul_rn(half, acc));      // D_C / d_H
    double dm = dc;
    if (c.Ok0 > 0.0) { const double s = sqrt(c.Ok0); dm = __ddiv_rn(sinh(__dmul_rn(s, dc)), s); }
    else if (c.Ok0 < 0.0) { const double s = sqrt(-c.Ok0); dm = __ddiv_rn(sin(__dmul_rn(s, dc)), s); }
    const double dH = __ddiv_rn(299792.458, c.H0);
    dm = __dmul_rn(dH, dm);                                         // transverse comoving distance [Mpc]
    if (DL) DL[i] = __dmul_rn(__dadd_rn(1.0, zi), dm);
    if (dV) dV[i] = __ddiv_rn(__dmul_rn(__dmul_rn(dH, dm), dm), efunc_np(c, zi));
}

// numpy.interp, compiled_base.c arr_interp: j = last knot <= x (candidate from the mean spacing, exact comparisons decide)
__global__ void k_interp(long long nk, const double* __restrict__ xk, const double* __restrict__ yk, long long n,
                         const double* __restrict__ x, double* __restrict__ y, int* __restrict__ bad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xv = x[i];
    const double x0 = xk[0], x1 = xk[nk - 1];
    if (!(xv >= x0) || !(xv <= x1)) { atomicExch(bad, 1); y[i] = xv != xv ? xv : 0.0; return; }
    long long j = (long long)((xv - x0) / (x1 - x0) * (double)(nk - 1));
    j = j < 0 ? 0 : (j > nk - 1 ? nk - 1 : j);
    int steps = 0;
    while (j > 0 && xk[j] > xv && steps < 8) { --j; ++steps; }
    while (j < nk - 1 && xk[j + 1] <= xv && steps < 8) { ++j; ++steps; }
    if (steps >= 8) {                                               // knots far from uniform: plain binary search
        long long lo = 0, hi = nk;                                  // invariant: xk[lo] <= xv, (hi == nk or xk[hi] > xv)
        while (hi - lo > 1) { long long mid = (lo + hi) >> 1; if (xk[mid] <= xv) lo = mid; else hi = mid; }
        j = lo;
    }
    double r;
    if (j == nk - 1) r = yk[j];
    else if (xk[j] == xv) r = yk[j];
    else {
        const double slope = __ddiv_rn(__dsub_rn(yk[j + 1], yk[j]), __dsub_rn(xk[j + 1], xk[j]));
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(xv, xk[j])), yk[j]);
        if (r != r) {
            r = __dadd_rn(__dmul_rn(slope, __dsub_rn(xv, xk[j + 1])), yk[j + 1]);
            if (r != r && yk[j] == yk[j + 1]) r = yk[j];
        }
    }
    y[i] = r;
}


extern "C" int lf_cosmo_distances(int32_t device, const lf_cosmology* cosmo, const double* cum, int64_t ncum, int64_t n,
                                  const double* z, double* DL_Mpc, double* dVdz) {
    if (!cosmo || !cum || ncum < 1 || n < 0 || (n > 0 && !z)) return fail("lf_cosmo_distances: bad arguments");
    if (n == 0) return 0;
    if (!(cosmo->panel > 0.0) || !(cosmo->H0 > 0.0)) return fail("lf_cosmo_distances: need panel > 0 and H0 > 0");
    CK(cudaSetDevice(device));
    DevBufs bufs;
    double *d_cum = nullptr, *d_z = nullptr, *d_DL = nullptr, *d_dV = nullptr, *d_gl = nullptr;
    int* d_bad = nullptr;
    CK(bufs.alloc(&d_cum, sizeof(double) * ncum));
    CK(bufs.alloc(&d_z, sizeof(double) * n));
    CK(bufs.alloc(&d_gl, sizeof(double) * 16));
    CK(bufs.alloc(&d_bad, sizeof(int)));
    if (DL_Mpc) CK(bufs.alloc(&d_DL, sizeof(double) * n));
    if (dVdz) CK(bufs.alloc(&d_dV, sizeof(double) * n));
    CK(cudaMemcpy(d_gl, cosmo->gl_x, sizeof(double) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gl + 8, cosmo->gl_w, sizeof(double) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cum, cum, sizeof(double) * ncum, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_z, z, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bad, 0, sizeof(int)));
    k_cosmo<<<(unsigned)((n + 255) / 256), 256>>>(*cosmo, d_cum, ncum, n, d_z, d_gl, d_gl + 8, d_DL, d_dV, d_bad);
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return fail("lf_cosmo_distances: a redshift is negative, NaN or beyond the cumulative panel table");
    if (DL_Mpc) CK(cudaMemcpy(DL_Mpc, d_DL, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (dVdz) CK(cudaMemcpy(dVdz, d_dV, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int lf_interp_linear(int32_t device, int64_t nk, const double* xk, const double* yk, int64_t n, const double* x,
                                double* y) {
    if (nk < 2 || !xk || !yk || n < 0 || (n > 0 && (!x || !y))) return fail("lf_interp_linear: bad arguments");
    if (n == 0) return 0;
    CK(cudaSetDevice(device));
    DevBufs bufs;
    double *d_xk = nullptr, *d_yk = nullptr, *d_x = nullptr, *d_y = nullptr;
    int* d_bad = nullptr;
    CK(bufs.alloc(&d_xk, sizeof(double) * nk));
    CK(bufs.alloc(&d_yk, sizeof(double) * nk));
    CK(bufs.alloc(&d_x, sizeof(double) * n));
    CK(bufs.alloc(&d_y, sizeof(double) * n));
    CK(bufs.alloc(&d_bad, sizeof(int)));
    CK(cudaMemcpy(d_xk, xk, sizeof(double) * nk, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_yk, yk, sizeof(double) * nk, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bad, 0, sizeof(int)));
    k_interp<<<(unsigned)((n + 255) / 256), 256>>>(nk, d_xk, d_yk, n, d_x, d_y, d_bad);
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return fail("lf_interp_linear: a value in x_new is outside the interpolation range (or NaN)");
    CK(cudaMemcpy(y, d_y, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device-resident ensemble sampler (stretch move)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01(uint32_t x) { return __dmul_rn((double)x + 0.5, 2.3283064365386963e-10); }   // (x + 1/2) / 2^32

struct SamplerArgs {
    int W, half, ndim;
    double a;
    uint32_t k0, k1;
    const long long* step;      // device counter: index of the current ensemble update
    long long step0;            // value of *step at the first replay (chain rows are relative to it)
    double* pos; double* lp;    // [W][ndim], [W]
    double* prop; double* lpnew; double* lnz; double* lnu;     // [half][ndim], [half] x 3
    double* chain; double* lnp; long long* nacc;               // [nsteps][W][ndim], [nsteps][W], [W]  (chain / lnp may be NULL)
};

// proposals for the walkers of half h (h = 0: [0, half), h = 1: [half, W)) against the other half
__global__ void k_stretch_propose(SamplerArgs s, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.half) return;
    const int me = h * s.half + i;
    const long long step = *s.step;
    uint32_t r[4];
    philox4x32_10((uint32_t)me, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)h, s.k0, s.k1, r);
    // z ~ g(z) propto 1/sqrt(z) on [1/a, a]:  z = ((a - 1) u + 1)^2 / a      (rounded operation by operation: the host
    // replay in tests/ reproduces the chain bit for bit)
    const double t = __dadd_rn(__dmul_rn(s.a - 1.0, u01(r[0])), 1.0);
    const double z = __ddiv_rn(__dmul_rn(t, t), s.a);
    const int partner = (1 - h) * s.half + (int)(((unsigned long long)r[1] * (unsigned long long)s.half) >> 32);
    for (int d = 0; d < s.ndim; ++d) {
        const double pp = s.pos[(long long)partner * s.ndim + d], pm = s.pos[(long long)me * s.ndim + d];
        s.prop[(long long)i * s.ndim + d] = __dsub_rn(pp, __dmul_rn(__dsub_rn(pp, pm), z));
    }
    s.lnz[i] = (s.ndim - 1.0) * log(z);
    s.lnu[i] = log(u01(r[2]));
}

__global__ void k_stretch_accept(SamplerArgs s, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.half) return;
    const int me = h * s.half + i;
    const double lnratio = s.lnz[i] + s.lpnew[i] - s.lp[me];       // NaN (inf - inf) compares false: rejected
    const bool acc = s.lnu[i] < lnratio;
    if (acc) {
        for (int d = 0; d < s.ndim; ++d) s.pos[(long long)me * s.ndim + d] = s.prop[(long long)i * s.ndim + d];
        s.lp[me] = s.lpnew[i];
        s.nacc[me] += 1;
    }
    const long long row = *s.step - s.step0;
    if (s.chain)
        for (int d = 0; d < s.ndim; ++d) s.chain[(row * s.W + me) * s.ndim + d] = s.pos[(long long)me * s.ndim + d];
    if (s.lnp) s.lnp[row * s.W + me] = s.lp[me];
}
__global__ void k_step_advance(long long* step) { *step += 1; }

__global__ void k_boot_draw(long long n, uint32_t k0, uint32_t k1, uint32_t rep_lo, uint32_t rep_hi, int* __restrict__ mult) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;       // Philox call q yields draws 4q .. 4q+3
    if (4 * q >= n) return;
    uint32_t r[4];
    philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), rep_lo, rep_hi, k0, k1, r);
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (4 * q + t < n) atomicAdd(&mult[(long long)(((unsigned long long)r[t] * (unsigned long long)n) >> 32)], 1);
}

extern "C" int lf_sampler_run(lf_ctx* c, const double* pos0, int64_t W, int64_t nsteps, uint64_t seed, double a, int64_t step0,
                              double* chain, double* lnprob, int64_t* naccepted, double* pos_out, double* lnprob_out) {
    if (!c || !pos0) return fail("lf_sampler_run: null argument");
    if (W < 2 || (W & 1)) return fail("lf_sampler_run: the number of walkers must be even");
    if (nsteps < 0 || !(a > 1.0)) return fail("lf_sampler_run: need nsteps >= 0 and a > 1");
    if (!c->have_sources || !c->have_grid) return fail("lf_sampler_run: call lf_set_sources and lf_set_grid first");
    CK(cudaSetDevice(c->device));
    const int ndim = c->ndim, half = (int)(W / 2);
    cudaStream_t st = c->stream;
    double *d_pos = nullptr, *d_lp = nullptr, *d_prop = nullptr, *d_lpnew = nullptr, *d_lnz = nullptr, *d_lnu = nullptr;
    double *d_chain = nullptr, *d_lnp = nullptr;
    long long *d_nacc = nullptr, *d_step = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int rc = 1;
    auto cleanup = [&]() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        dfree(d_pos); dfree(d_lp); dfree(d_prop); dfree(d_lpnew); dfree(d_lnz); dfree(d_lnu);
        dfree(d_chain); dfree(d_lnp); dfree(d_nacc); dfree(d_step);
        return rc;
    };
#define SCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return cleanup();                                                                       \
        }                                                                                           \
    } while (0)
    SCK(cudaMalloc(&d_pos, sizeof(double) * W * ndim));
    SCK(cudaMalloc(&d_lp, sizeof(double) * W));
    SCK(cudaMalloc(&d_prop, sizeof(double) * half * ndim));
    SCK(cudaMalloc(&d_lpnew, sizeof(double) * half));
    SCK(cudaMalloc(&d_lnz, sizeof(double) * half));
    SCK(cudaMalloc(&d_lnu, sizeof(double) * half));
    SCK(cudaMalloc(&d_nacc, sizeof(long long) * W));
    SCK(cudaMalloc(&d_step, sizeof(long long)));
    if (chain && nsteps > 0) SCK(cudaMalloc(&d_chain, sizeof(double) * (size_t)nsteps * W * ndim));
    if (lnprob && nsteps > 0) SCK(cudaMalloc(&d_lnp, sizeof(double) * (size_t)nsteps * W));
    SCK(cudaMemcpyAsync(d_pos, pos0, sizeof(double) * W * ndim, cudaMemcpyHostToDevice, st));
    SCK(cudaMemsetAsync(d_nacc, 0, sizeof(long long) * W, st));
    const long long s0 = step0;
    SCK(cudaMemcpyAsync(d_step, &s0, sizeof(long long), cudaMemcpyHostToDevice, st));
    // log-posterior of the starting ensemble, and one un-captured half-ensemble call so that every scratch buffer
    // the captured pipeline needs already has its final size
    // source-sharded run over several GPUs: every rank draws the same Philox proposals and the per-walker partials are
    // summed by the peer-memory kernel inside the captured update (identical bits on every rank keep the chains equal)
    const bool exchange = c->peer_connected && c->peer.world > 1;
    if (exchange && W > c->peer.wcap) { fail("lf_sampler_run: more walkers than the peer buffers hold"); return cleanup(); }
    auto lnprob_all_ranks = [&](const double* th, long long nw, double* out) -> int {
        if (launch_pipeline(c, th, nw, out, st)) return 1;
        if (exchange) {
            k_allreduce_p2p<<<(unsigned)((nw + PEER_CHUNK - 1) / PEER_CHUNK), PEER_CHUNK, 0, st>>>(c->peer, out, nw);
            k_seq_advance<<<1, 1, 0, st>>>(c->peer_seq);
            c->launches += 2;
        }
        return 0;
    };
    if (lnprob_all_ranks(d_pos, W, d_lp)) return cleanup();
    if (lnprob_all_ranks(d_pos, half, d_lpnew)) return cleanup();
    SCK(cudaStreamSynchronize(st));
    SamplerArgs sa;
    sa.W = (int)W; sa.half = half; sa.ndim = ndim; sa.a = a;
    sa.k0 = (uint32_t)seed; sa.k1 = (uint32_t)(seed >> 32);
    sa.step = d_step; sa.step0 = step0;
    sa.pos = d_pos; sa.lp = d_lp; sa.prop = d_prop; sa.lpnew = d_lpnew; sa.lnz = d_lnz; sa.lnu = d_lnu;
    sa.chain = d_chain; sa.lnp = d_lnp; sa.nacc = d_nacc;
    const int T = 128, G = (half + T - 1) / T;
    const long long launches0 = c->launches;
    if (nsteps > 0) {
        SCK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        bool ok = true;
        for (int h = 0; h < 2 && ok; ++h) {
            k_stretch_propose<<<G, T, 0, st>>>(sa, h);
            ok = lnprob_all_ranks(d_prop, half, d_lpnew) == 0;
            k_stretch_accept<<<G, T, 0, st>>>(sa, h);
            c->launches += 2;
        }
        k_step_advance<<<1, 1, 0, st>>>(d_step);
        c->launches += 1;
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (!ok) return cleanup();
        SCK(ce);
        SCK(cudaGraphInstantiate(&exec, graph, 0));
        const long long per_step = c->launches - launches0;
        SCK(cudaEventRecord(c->ev0, st));
        for (int64_t t = 0; t < nsteps; ++t) SCK(cudaGraphLaunch(exec, st));
        SCK(cudaEventRecord(c->ev1, st));
        c->launches = launches0 + per_step * nsteps;
    }
    if (chain && nsteps > 0) SCK(cudaMemcpyAsync(chain, d_chain, sizeof(double) * (size_t)nsteps * W * ndim, cudaMemcpyDeviceToHost, st));
    if (lnprob && nsteps > 0) SCK(cudaMemcpyAsync(lnprob, d_lnp, sizeof(double) * (size_t)nsteps * W, cudaMemcpyDeviceToHost, st));
    if (naccepted) SCK(cudaMemcpyAsync(naccepted, d_nacc, sizeof(long long) * W, cudaMemcpyDeviceToHost, st));
    if (pos_out) SCK(cudaMemcpyAsync(pos_out, d_pos, sizeof(double) * W * ndim, cudaMemcpyDeviceToHost, st));
    if (lnprob_out) SCK(cudaMemcpyAsync(lnprob_out, d_lp, sizeof(double) * W, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    SCK(cudaGetLastError());
    if (nsteps > 0) { float ms = 0.f; SCK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->sampler_ms = ms; }
#undef SCK
    rc = 0;
    return cleanup();
}

extern "C" int lf_sampler_last_ms(lf_ctx* c, double* ms) {
    if (!c || !ms) return fail("lf_sampler_last_ms: null argument");
    *ms = c->sampler_ms;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Veff launch plan: lane-private histograms whenever they fit in shared memory, atomics otherwise
// ------------------------------------------------------------------------------------------------
struct VeffPlan { bool priv; int blocks, threads; size_t smem; };
static VeffPlan veff_plan(const lf_ctx* c, long long n, int nbins) {
    VeffPlan p;
    const size_t priv = sizeof(double) * (nbins + 1) * 16 + (size_t)VP_WARPS * nbins * VP_COLS * (sizeof(double) + sizeof(unsigned)) +
                        sizeof(double2) * STREAM_LOG_N + sizeof(double) * EXP_TAB_N;
    if (priv <= VP_SMEM_MAX) {
        int per_sm = (int)std::min<size_t>(8, (size_t)(227 * 1024) / (priv + 1024));
        per_sm = std::max(per_sm, 1);
        p.priv = true; p.threads = 32 * VP_WARPS; p.smem = priv;
        p.blocks = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * per_sm,
                                                                     (n + p.threads * VP_UNROLL - 1) / (p.threads * VP_UNROLL)));
    } else {
        p.priv = false; p.threads = 256;
        p.smem = sizeof(double) * (nbins + 1) + (sizeof(double) + sizeof(unsigned long long)) * 8 * (size_t)nbins;
        p.blocks = (int)std::min<long long>((long long)c->sm_count * 8, (n + 255) / 256);
    }
    return p;
}
template <int MODE>
static void veff_launch(const VeffPlan& p, const VeffArgs& a, cudaStream_t st) {
    if (p.priv) k_veff_priv<MODE><<<p.blocks, p.threads, p.smem, st>>>(a);
    else k_veff<MODE><<<p.blocks, p.threads, p.smem, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// Veff host entry points
// ------------------------------------------------------------------------------------------------
extern "C" int lf_veff_bin(lf_ctx* c, int64_t n, const double* flux, const double* lum, const int64_t* field_ind,
                           int32_t nfields, const double* flim, double alpha, double fcmin, double sum_omega,
                           double vol_int, const double* vol_per_source, const uint8_t* valid,
                           const double* edges, int32_t nbins, double* phi_out, int64_t* counts, double* sumphi) {
    if (!c) return fail("lf_veff_bin: null context");
    if (n <= 0 || !flux || !lum || !field_ind || !flim || !edges || !counts || !sumphi) return fail("lf_veff_bin: bad arguments");
    if (nfields < 1 || nfields > LF_MAX_FIELDS) return fail("lf_veff_bin: nfields out of range");
    if (nbins < 1 || nbins > VEFF_MAX_BINS) return fail("lf_veff_bin: nbins out of range");
    if (field_ind[0] != 0 || field_ind[nfields] != n) return fail("lf_veff_bin: field_ind must run from 0 to n");
    CK(cudaSetDevice(c->device));
    dfree(c->v_lum); dfree(c->v_phi); dfree(c->v_edges); dfree(c->v_counts); dfree(c->v_sums);
    dfree(c->v_outc); dfree(c->v_outs); dfree(c->v_mult); dfree(c->v_bin);
    const size_t nb = sizeof(double) * (size_t)n;
    double* d_flux = nullptr; double* d_vol = nullptr; unsigned char* d_valid = nullptr;
    CK(cudaMalloc(&d_flux, nb));
    CK(cudaMalloc(&c->v_lum, nb));
    CK(cudaMalloc(&c->v_phi, nb));
    CK(cudaMalloc(&c->v_edges, sizeof(double) * (nbins + 1)));
    const VeffPlan plan = veff_plan(c, n, nbins);
    const int blocks = plan.blocks;
    c->v_blocks = blocks; c->v_nbins = nbins; c->vN = n;
    CK(cudaMalloc(&c->v_counts, sizeof(unsigned long long) * (size_t)blocks * nbins));
    CK(cudaMalloc(&c->v_sums, sizeof(double) * (size_t)blocks * nbins));
    CK(cudaMalloc(&c->v_outc, sizeof(long long) * nbins));
    CK(cudaMalloc(&c->v_outs, sizeof(double) * nbins));
    CK(cudaMalloc(&c->v_bin, sizeof(short) * (size_t)n));
    CK(cudaMemcpyAsync(d_flux, flux, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v_lum, lum, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v_edges, edges, sizeof(double) * (nbins + 1), cudaMemcpyHostToDevice, c->stream));
    if (vol_per_source) {
        CK(cudaMalloc(&d_vol, nb));
        CK(cudaMemcpyAsync(d_vol, vol_per_source, nb, cudaMemcpyHostToDevice, c->stream));
    }
    if (valid) {
        CK(cudaMalloc(&d_valid, (size_t)n));
        CK(cudaMemcpyAsync(d_valid, valid, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    }
    VeffArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.flux = d_flux; a.lum = c->v_lum; a.vol = d_vol; a.valid = d_valid; a.phi = c->v_phi;
    a.K = nfields;
    for (int k = 0; k <= nfields; ++k) a.field_ind[k] = field_ind[k];
    const bool modified = fcmin != 0.0;
    double aa = (2.0 * fcmin - 1.0) * (2.0 * fcmin - 1.0);
    for (int k = 0; k < nfields; ++k) {
        a.F50[k] = 1.0e-17 * flim[k];
        // inverse_fleming, reference operation order (VmaxLumFunc.py:164-167)
        double b = -1.0 * pow(fabs(aa / (1.0 - aa)) * pow(alpha, -2.0), 0.5);
        a.ftau[k] = a.F50[k] * pow(10.0, b);
        a.invF50[k] = 1.0 / a.F50[k];
        a.inv_ftau[k] = 1.0 / a.ftau[k];
    }
    a.alpha = alpha; a.pref = sum_omega / SQARCSEC; a.vol_int = vol_int; a.modified = modified ? 1 : 0;
    a.inv_pref_vol = 1.0 / (a.pref * vol_int); a.tables = c->d_tables;
    a.edges = c->v_edges; a.nbins = nbins; a.counts = c->v_counts; a.sumphi = c->v_sums; a.mult = nullptr; a.bin = c->v_bin;
    CK(cudaEventRecord(c->ev0, c->stream));
    veff_launch<0>(plan, a, c->stream);
    k_veff_reduce<<<(nbins + 3) / 4, 128, 0, c->stream>>>(blocks, nbins, c->v_counts, c->v_sums, c->v_outc, c->v_outs);
    CK(cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, c->v_outc, sizeof(long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sumphi, c->v_outs, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
    if (phi_out) CK(cudaMemcpyAsync(phi_out, c->v_phi, nb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->last_ms = ms; }
    cudaFree(d_flux);
    if (d_vol) cudaFree(d_vol);
    if (d_valid) cudaFree(d_valid);
    return 0;
}

extern "C" int lf_bin_weights(lf_ctx* c, int64_t n, const double* lum, const double* phi, const double* edges,
                              int32_t nbins, int64_t* counts, double* sumphi) {
    if (!c) return fail("lf_bin_weights: null context");
    if (n <= 0 || !lum || !phi || !edges || !counts || !sumphi) return fail("lf_bin_weights: bad arguments");
    if (nbins < 1 || nbins > VEFF_MAX_BINS) return fail("lf_bin_weights: nbins out of range");
    CK(cudaSetDevice(c->device));
    dfree(c->v_lum); dfree(c->v_phi); dfree(c->v_edges); dfree(c->v_counts); dfree(c->v_sums);
    dfree(c->v_outc); dfree(c->v_outs); dfree(c->v_mult); dfree(c->v_bin);
    const size_t nb = sizeof(double) * (size_t)n;
    CK(cudaMalloc(&c->v_lum, nb));
    CK(cudaMalloc(&c->v_phi, nb));
    CK(cudaMalloc(&c->v_edges, sizeof(double) * (nbins + 1)));
    const VeffPlan plan = veff_plan(c, n, nbins);
    const int blocks = plan.blocks;
    c->v_blocks = blocks; c->v_nbins = nbins; c->vN = n;
    CK(cudaMalloc(&c->v_counts, sizeof(unsigned long long) * (size_t)blocks * nbins));
    CK(cudaMalloc(&c->v_sums, sizeof(double) * (size_t)blocks * nbins));
    CK(cudaMalloc(&c->v_outc, sizeof(long long) * nbins));
    CK(cudaMalloc(&c->v_outs, sizeof(double) * nbins));
    CK(cudaMalloc(&c->v_bin, sizeof(short) * (size_t)n));
    CK(cudaMemcpyAsync(c->v_lum, lum, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v_phi, phi, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v_edges, edges, sizeof(double) * (nbins + 1), cudaMemcpyHostToDevice, c->stream));
    VeffArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.lum = c->v_lum; a.phi = c->v_phi; a.edges = c->v_edges; a.nbins = nbins;
    a.counts = c->v_counts; a.sumphi = c->v_sums; a.bin = c->v_bin;
    CK(cudaEventRecord(c->ev0, c->stream));
    veff_launch<2>(plan, a, c->stream);
    k_veff_reduce<<<(nbins + 3) / 4, 128, 0, c->stream>>>(blocks, nbins, c->v_counts, c->v_sums, c->v_outc, c->v_outs);
    CK(cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, c->v_outc, sizeof(long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sumphi, c->v_outs, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (cudaEventQuery(c->ev1) == cudaSuccess) { float ms = 0.f; if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms; }
    return 0;
}

// multiplicities of one bootstrap replicate: n uniform draws with replacement, 4 per Philox call
__global__ void k_boot_draw(long long n, uint32_t k0, uint32_t k1, uint32_t rep_lo, uint32_t rep_hi, int* __restrict__ mult);

extern "C" int lf_boot_bin_device(lf_ctx* c, uint64_t seed, int64_t replicate, int64_t* counts, double* sumphi) {
    if (!c) return fail("lf_boot_bin_device: null context");
    if (!c->v_phi || c->vN <= 0) return fail("lf_boot_bin_device: call lf_veff_bin or lf_bin_weights first");
    if (!counts || !sumphi) return fail("lf_boot_bin_device: bad arguments");
    if (c->vN >= (1LL << 32)) return fail("lf_boot_bin_device: more than 2^32 sources");
    CK(cudaSetDevice(c->device));
    if (!c->v_mult) CK(cudaMalloc(&c->v_mult, sizeof(int) * (size_t)c->vN));
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(c->v_mult, 0, sizeof(int) * (size_t)c->vN, c->stream));
    const long long calls = (c->vN + 3) / 4;
    k_boot_draw<<<(unsigned)((calls + 255) / 256), 256, 0, c->stream>>>(c->vN, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                      (uint32_t)replicate, (uint32_t)((uint64_t)replicate >> 32), c->v_mult);
    VeffArgs a;
    memset(&a, 0, sizeof(a));
    a.n = c->vN; a.lum = c->v_lum; a.phi = c->v_phi; a.edges = c->v_edges; a.nbins = c->v_nbins;
    a.counts = c->v_counts; a.sumphi = c->v_sums; a.mult = c->v_mult; a.bin = c->v_bin;
    const int nbins = c->v_nbins, blocks = c->v_blocks;
    const VeffPlan plan = veff_plan(c, c->vN, nbins);
    veff_launch<1>(plan, a, c->stream);
    k_veff_reduce<<<(nbins + 3) / 4, 128, 0, c->stream>>>(blocks, nbins, c->v_counts, c->v_sums, c->v_outc, c->v_outs);
    CK(cudaEventRecord(c->ev1, c->stream));
    c->launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, c->v_outc, sizeof(long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sumphi, c->v_outs, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    { float ms = 0.f; if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms; }
    return 0;
}

extern "C" int lf_boot_bin(lf_ctx* c, const int32_t* mult, int64_t* counts, double* sumphi) {
    if (!c) return fail("lf_boot_bin: null context");
    if (!c->v_phi || c->vN <= 0) return fail("lf_boot_bin: call lf_veff_bin first");
    if (!mult || !counts || !sumphi) return fail("lf_boot_bin: bad arguments");
    CK(cudaSetDevice(c->device));
    if (!c->v_mult) CK(cudaMalloc(&c->v_mult, sizeof(int) * (size_t)c->vN));
    CK(cudaMemcpyAsync(c->v_mult, mult, sizeof(int) * (size_t)c->vN, cudaMemcpyHostToDevice, c->stream));
    VeffArgs a;
    memset(&a, 0, sizeof(a));
    a.n = c->vN; a.lum = c->v_lum; a.phi = c->v_phi; a.edges = c->v_edges; a.nbins = c->v_nbins;
    a.counts = c->v_counts; a.sumphi = c->v_sums; a.mult = c->v_mult; a.bin = c->v_bin;
    const int nbins = c->v_nbins, blocks = c->v_blocks;
    const VeffPlan plan = veff_plan(c, c->vN, nbins);
    CK(cudaEventRecord(c->ev0, c->stream));
    veff_launch<1>(plan, a, c->stream);
    k_veff_reduce<<<(nbins + 3) / 4, 128, 0, c->stream>>>(blocks, nbins, c->v_counts, c->v_sums, c->v_outc, c->v_outs);
    CK(cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, c->v_outc, sizeof(long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sumphi, c->v_outs, sizeof(double) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (cudaEventQuery(c->ev1) == cudaSuccess) { float ms = 0.f; if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms; }
    return 0;
}
