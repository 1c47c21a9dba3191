// lf_internal.cuh -- declarations shared by the translation units of liblfengine.so (not part of the C ABI).
//   lf_engine.cu   context, likelihood kernels (prologue / main / finish), lnprob entry points, pipe micro-benchmarks
//   lf_veff.cu     1/V_eff weights, binned LF, bootstrap replicates
//   lf_sampler.cu  device-resident ensemble sampler
//   lf_peer.cu     peer-memory all-reduce (multi-GPU exchange)
//   lf_setup.cu    set-up tables on the GPU (cosmology distances, NumPy-exact interpolation)
#pragma once
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
// error plumbing (defined in lf_engine.cu): 0 = OK, non-zero + lf_last_error() otherwise
// ------------------------------------------------------------------------------------------------
int fail(const std::string& msg);
#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                            \
    } while (0)

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
    double c_alpha_max;                // largest alpha_c the compressed catalogue is accurate for
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
// Philox4x32-10 counter RNG (device sampler, device-resampled bootstrap)
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


// ------------------------------------------------------------------------------------------------
// LambdaCDM distances, the arithmetic of lumfuncmcmc_b200/cosmology.py operation for operation (lf_setup.cu, lf_veff.cu)
// ------------------------------------------------------------------------------------------------
// E(z) with NumPy's order of operations and no fused multiply-adds (cosmology.py efunc)
__device__ __forceinline__ double efunc_np(const lf_cosmology& c, double z) {
    const double zp1 = __dadd_rn(1.0, z);
    double t = __dadd_rn(__dmul_rn(c.Or0, zp1), c.Om0);
    t = __dadd_rn(__dmul_rn(t, zp1), c.Ok0);
    t = __dadd_rn(__dmul_rn(__dmul_rn(zp1, zp1), t), c.Ode0);
    return sqrt(t);
}

// transverse comoving distance dm [Mpc] and D_C / d_H at redshift zi: cumulative 8-point Gauss-Legendre panels (cum[p] =
// int_0^{p * panel} dz / E) + one 8-point closure.  Returns false for a redshift outside the panel table.
__device__ __forceinline__ bool cosmo_dm(const lf_cosmology& c, const double* __restrict__ cum, long long ncum,
                                         const double* __restrict__ glx, const double* __restrict__ glw, double zi,
                                         double& dm, double& dc) {
    const long long p = (long long)floor(__ddiv_rn(zi, c.panel));
    if (!(zi >= 0.0) || p >= ncum) return false;
    const double lo = __dmul_rn((double)p, c.panel);
    const double half = __dmul_rn(0.5, __dsub_rn(zi, lo));
    double acc = 0.0;
    for (int q = 0; q < 8; ++q) {                                   // acc += w / E(lo + half * (1 + x)), in node order
        const double node = __dadd_rn(lo, __dmul_rn(half, __dadd_rn(1.0, glx[q])));
        acc = __dadd_rn(acc, __ddiv_rn(glw[q], efunc_np(c, node)));
    }
    dc = __dadd_rn(cum[p], __dmul_rn(half, acc));                   // D_C / d_H
    dm = dc;
    if (c.Ok0 > 0.0) { const double s = sqrt(c.Ok0); dm = __ddiv_rn(sinh(__dmul_rn(s, dc)), s); }
    else if (c.Ok0 < 0.0) { const double s = sqrt(-c.Ok0); dm = __ddiv_rn(sin(__dmul_rn(s, dc)), s); }
    dm = __dmul_rn(__ddiv_rn(299792.458, c.H0), dm);                // [Mpc]
    return true;
}

// j = last knot <= xv for increasing knots xk[nk] and xk[0] <= xv <= xk[nk - 1] (numpy.interp's segment): candidate from
// the mean spacing, exact comparisons decide; knots far from uniform fall back to a binary search
__device__ __forceinline__ long long knot_segment(long long nk, const double* __restrict__ xk, double xv) {
    const double x0 = xk[0], x1 = xk[nk - 1];
    long long j = (long long)((xv - x0) / (x1 - x0) * (double)(nk - 1));
    j = j < 0 ? 0 : (j > nk - 1 ? nk - 1 : j);
    int steps = 0;
    while (j > 0 && xk[j] > xv && steps < 8) { --j; ++steps; }
    while (j < nk - 1 && xk[j + 1] <= xv && steps < 8) { ++j; ++steps; }
    if (steps >= 8) {
        long long lo = 0, hi = nk;                                  // invariant: xk[lo] <= xv, (hi == nk or xk[hi] > xv)
        while (hi - lo > 1) { long long mid = (lo + hi) >> 1; if (xk[mid] <= xv) lo = mid; else hi = mid; }
        j = lo;
    }
    return j;
}

// ------------------------------------------------------------------------------------------------
// the context
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
    int* timed_out;                    // host-mapped, sticky: set when a wait expired
    long long spin_limit;              // clock64 ticks a wait may last
};
#define PEER_SPIN_CLOCKS_DEFAULT 60000000000LL   /* ~30 s at 1.97 GHz: first-call module loads, re-allocations and host pauses fit */

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
    // sample kept resident across V_eff calls (lf_veff_set_sample) and its per-source volumes (lf_veff_volumes)
    double* v_flux = nullptr; double* v_vol = nullptr; unsigned char* v_valid = nullptr;
    uint32_t* v_mt_vals = nullptr; long long v_mt_cap = 0;      // candidate values of one replicate (+ the count of words used)
    uint32_t* v_mt_state = nullptr;                // MT19937 key[624] + position (device bootstrap with NumPy's stream)
    double* v_u = nullptr;                         // log10(flux / VRES_F0), computed once per sample
    bool v_rows_valid = false; std::vector<double> v_edges_host;     // edges the resident rows / counts were computed for
    unsigned long long* v_rowcounts = nullptr; int v_rowcounts_n = 0; unsigned* v_ticket = nullptr;
    // volume table of lf_veff_set_volume_table: cosmology + panel integrals, knots of the dV/dz interpolant, its cumulative integral
    lf_cosmology v_cosmo; double* v_cum = nullptr; long long v_ncum = 0; double* v_gl = nullptr;
    double* v_zk = nullptr; double* v_dVk = nullptr; double* v_cumV = nullptr; long long v_nk = 0;
    bool v_have_sample = false, v_have_volumes = false; int v_K = 0; long long v_field_ind[LF_MAX_FIELDS + 1] = {};
    // ordering of the shared per-call scratch across streams (see scratch_acquire)
    cudaEvent_t ev_scratch = nullptr; cudaStream_t scratch_stream = nullptr; bool scratch_pending = false;
    // peer exchange
    unsigned char* peer_base = nullptr; unsigned* peer_seq = nullptr; int* peer_timeout = nullptr; int* peer_timeout_h = nullptr;
    bool walker_shard = false;         // lf_sampler_run over several ranks: shard walkers (all sources on every rank) instead of sources
    size_t peer_data_bytes = 0; bool peer_connected = false; void* peer_opened[PEER_MAX] = {};
    PeerArgs peer;
};

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
inline void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}


// ------------------------------------------------------------------------------------------------
// cross-unit host functions
// ------------------------------------------------------------------------------------------------
// lf_engine.cu: enqueue one batched log-posterior evaluation (prologue, main kernels, finish) on `st`
int launch_pipeline(lf_ctx* c, const double* d_thetas, long long W, double* d_out, cudaStream_t st);
// lf_engine.cu: the per-call scratch (walker constants, partial rows, class lists, V_eff residency) is shared by every entry
// point of a context.  A caller may drive lf_lnprob_batch_device on its own stream while the host API, the sampler and the
// V_eff calls use the context's stream: scratch_acquire makes `st` wait for the last user on another stream,
// scratch_release records where the next user has to wait.  (Skipped while `st` is being captured into a CUDA graph.)
int scratch_acquire(lf_ctx* c, cudaStream_t st);
int scratch_release(lf_ctx* c, cudaStream_t st);
// lf_veff.cu: opt the V_eff kernels into their dynamic shared-memory sizes (called once per context)
int veff_init(lf_ctx* c);
// lf_peer.cu: enqueue the in-place sum over ranks of d_vec[W] (two launches); release peer mappings at destroy
int peer_allreduce_launch(lf_ctx* c, double* d_vec, long long W, cudaStream_t st);
void peer_release(lf_ctx* c);
