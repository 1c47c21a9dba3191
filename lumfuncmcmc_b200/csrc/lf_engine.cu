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
#include "lf_internal.cuh"

#ifndef LF_PDL
#define LF_PDL 1          /* programmatic dependent launch of the fast main kernel (its table fill overlaps the prologue) */
#endif

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

extern "C" const char* lf_last_error(void) { return g_err.c_str(); }
extern "C" const char* lf_version(void) { return "lfengine 0.2 (sm_100a)"; }
extern "C" int lf_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
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
#if LF_PDL
    asm volatile("griddepcontrol.launch_dependents;");          // the next kernel's blocks may start their table fill now
#endif
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
            double emax = exp10(dhi);
            lb = LNLN10 + LN10 * Plo + fmin(c1 * dlo, c1 * dhi) - emax + s.lnom_min;
            if (!(emax < 690.0)) lb = -1.0e300;
            if (!(dlo > -40.0)) lb = -1.0e300;
            // the source loop evaluates 2^(log2(10) (lum - L*(z)) - log2(10) (42 - L*(z2))) without clamping the exponent field:
            // with the two bounds above the first part lies in (-133, 9.5); this keeps the second below 665
            if (!(fabs(42.0 - L2) < 200.0)) lb = -1.0e300;
            // compressed catalogue: |dL*/dz| over the field's redshift range must stay inside the bound it was built for
            if (a.csrc != nullptr && !(fmax(fabs(fma(2.0 * aL, s.z_min, bL)), fabs(fma(2.0 * aL, s.z_max, bL))) <= a.c_alpha_max)) lb = -1.0e300;
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
        // exp10 instead of the generic pow: same <= 1 ulp accuracy, a fraction of the latency (the prologue is a chain of
        // dependent transcendentals; it matters for small ensembles)
        if (!rejected && a.N > 0 && exp(-exp10(a.lum_max_all - Lstar)) == 0.0) rejected = true;
        double tenmL = exp10(-Lstar);
        double c1 = (sal + 1.0) * LN10;
        double c0 = LNLN10 + phistar * LN10 - Lstar * c1;
        if (k == 0 && live && !rejected) {
            wp[P_TENML * WS] = tenmL; wp[P_C0 * WS] = c0; wp[P_C1 * WS] = c1;
            wp[P_LSTAR * WS] = Lstar; wp[P_PHISTAR * WS] = phistar; wp[P_SCHAL * WS] = sal;
            wp[P_ALPHA * WS] = alpha_c;
        }
        double tmin = 0.0, lnom = 0.0;
        if (a.model == LF_MODEL_FREE) {
            double b = -1.0 * sqrt(a.fcA2 / (alpha_c * alpha_c));   // inverse_fleming, VmaxLumFunc.py:164-165
            if (!(alpha_c > 0.0)) rok = 0;
            if (a.csrc != nullptr && !(alpha_c <= a.c_alpha_max)) rok = 0;      // outside the compressed catalogue's error bound
            double F50 = 1.0e-17 * th[p + k];
            double lgF = log10(F50);
            double ftau = F50 * exp10(b);
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
            if (a.modified && !(a.fcap / ftau < 7.0e5)) rok = 0;      // round(2048 log2(e) f / ftau) of the exp range reduction stays in int32
            if (s.n > 0.0) tmin = log(fleming_literal(s.f_min, F50, alpha_c, ftau, a.modified != 0));
            lnom = s.ln_om0;
        }
        if (s.n != 0.0) {
            // Schechter exponent S(l) = c0 + c1 l - 10^(l - L*) is concave in l: minimum at an end of the range
            double s_lo = fmin(c0 + c1 * s.lum_min - exp10(s.lum_min - Lstar),
                               c0 + c1 * s.lum_max - exp10(s.lum_max - Lstar));
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
#if LF_PDL
    asm volatile("griddepcontrol.launch_dependents;");
#endif
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
    a.colB[(long long)i * WS + w] = exp10(-Ls);
}

// ------------------------------------------------------------------------------------------------
// main kernels: grid of warp work items = (walker group of 32) x (slab of sources | slab of quadrature points)
// ------------------------------------------------------------------------------------------------
// one 12-warp block per SM (3 warps per scheduler, up to 168 registers per thread): the shared-memory tables
// (217 KB) are filled once per SM.  Measured sweep of (sources in lock-step, warps): profiles/README.md
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
#ifndef LF_Z_CLAMP
#define LF_Z_CLAMP 0
#endif
#ifndef LF_ZF_WARPS
#define LF_ZF_WARPS 16
#endif
__host__ __device__ constexpr int main_warps(int model) { return model == LF_MODEL_FREE ? WARPS_PER_BLOCK : LF_ZF_WARPS; }
__host__ __device__ constexpr size_t main_table_bytes(int model) {
    return model == LF_MODEL_FREE ? sizeof(double2) * LOG_TAB_N * LOG_TAB_REP + sizeof(double) * EXP_SMEM_DOUBLES
                                  : sizeof(double) * EXPR_SMEM_DOUBLES;
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
// The literal kernels are chains of dependent libdevice calls (~1150 instructions per term, ncu: IPC 0.2 with 12 warps per SM):
// they want warps, not registers, so the free-completeness one is held to two resident blocks per SM.
template <bool LITERAL, int MODEL>
__global__ void __launch_bounds__(32 * main_warps(MODEL), (LITERAL && MODEL == LF_MODEL_FREE) ? 2 : LF_MIN_BLOCKS) k_main(KArgs a) {
    extern __shared__ __align__(16) unsigned char smem_tables[];       // fast kernels only (main_smem_bytes(MODEL))
    double2* s_log = reinterpret_cast<double2*>(smem_tables);          // FREE: log table; Z / FIXED: the replicated exp table
    double* s_exp = reinterpret_cast<double*>(s_log + LOG_TAB_N * LOG_TAB_REP);     // FREE only
    double2* s_stage = reinterpret_cast<double2*>(smem_tables + main_table_bytes(MODEL)) + (threadIdx.x >> 5) * (QSTAGE * 3);
    const int cls = LITERAL ? CLS_LIT : CLS_FAST;
    if (!LITERAL) {
        // FREE: log table + (big) exp table; Z / FIXED: no log table, the small replicated exp table takes its place.
        // The fast kernel is launched with programmatic stream serialisation: its blocks become resident while the
        // prologue is still running and fill their tables (which do not depend on it) in its shadow; everything the
        // prologue writes is read only after griddepcontrol.wait.
        if (MODEL == LF_MODEL_FREE) load_tables(a.tables, s_exp, s_log);
        else load_exp_replicated(a.tables, reinterpret_cast<double*>(s_log));
#if LF_PDL
        asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
        __syncthreads();
    }
    const int count_src = a.cls_count[cls], count_quad = a.cls_count[LITERAL ? 6 : 5];
    // FREE / Z fast kernels: a source item covers TWO walkers per lane (64 per warp), so every broadcast source load
    // (16 B x 32 lanes = 4 wavefronts of the load-store data path, the busiest unit of these loops) feeds two terms
    constexpr bool PAIR = (MODEL == LF_MODEL_Z || MODEL == LF_MODEL_FREE) && !LITERAL;
    // Literal source items take ONE walker each and spread the slab's sources over the 32 lanes (fixed shuffle tree at the
    // end): the literal class is small (a handful of walkers during burn-in, none afterwards), so one walker per lane would
    // leave most of every warp idle while each term costs ~25x a fast one.
    constexpr bool LANE_SRC = LITERAL;
    const int n_wg = LANE_SRC ? count_src : (PAIR ? (count_src + 63) >> 6 : (count_src + 31) >> 5);   // n_wg == 0: no source items
    const int n_wgq = (count_quad + 31) >> 5;
    const long long n_src_items = (long long)n_wg * a.n_src_slabs;
    const long long n_items = n_src_items + (long long)n_wgq * a.n_quad_slabs;
    // persistent warps: every warp pulls (walker group, slab) items from a global counter until none are left.
    // Items are small (tens per warp slot), so SMs finish within one item of each other (no wave tail), and each
    // item owns its own partial[] row, so the result does not depend on which warp ran it.
    const int lane = threadIdx.x & 31;
    const int rep16 = lane & (EXP_TAB_REP - 1), rep8 = lane & (LOG_TAB_REP - 1);   // table replica of this lane
    const double* s_exp_rep = reinterpret_cast<const double*>(s_log);               // Z / FIXED models only
    const unsigned lane8 = 8u * lane;                                                // this lane's column of that table
    const long long WS = a.Wcap;
    int* counter = a.cls_count + (LITERAL ? 4 : 3);
    // the first item of every warp is assigned statically (its global warp index), the following ones come from the
    // counter: no burst of a few thousand atomics on one address at kernel start
    const long long total_warps = (long long)gridDim.x * (blockDim.x >> 5);
    bool first = true;
  for (;;) {
    long long item = 0;
    if (first) {
        item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        first = false;
    } else {
        if (lane == 0) item = total_warps + atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
    }
    if (item >= n_items) break;
    // source items first (all walkers of the class), then quadrature items (the walkers this rank integrates)
    const bool is_src = item < n_src_items;
    const long long it = is_src ? item : item - n_src_items;
    const int groups = is_src ? n_wg : n_wgq;
    const int wg = (int)(it % groups);
    const int row = (int)(it / groups) + (is_src ? 0 : a.n_src_slabs);
    const int count = is_src ? count_src : count_quad;
    const int* list = is_src ? (LITERAL ? a.list_lit : a.list_fast) : (LITERAL ? a.list_litq : a.list_fastq);
    const int slot = (LANE_SRC && is_src) ? wg : (PAIR && is_src ? wg * 64 : wg * 32) + lane;
    const bool active = slot < count;
    const int lstep = (LANE_SRC && is_src) ? 32 : 1, loff = (LANE_SRC && is_src) ? lane : 0;   // literal source loops: lane-strided
    const long long w = list[active ? slot : count - 1];       // inactive lanes shadow a valid walker
    const double* wp = a.wp + w;
    const bool activeB = PAIR && is_src && slot + 32 < count;   // second walker of the lane (Z source items)
    const long long wB = PAIR ? list[activeB ? slot + 32 : count - 1] : w;
    double acc0 = 0.0, acc1 = 0.0, accB = 0.0;

    if (row < a.n_src_slabs) {
        // ---------------- source slab ----------------
        long long i0 = (a.N * row) / a.n_src_slabs, i1 = (a.N * (row + 1)) / a.n_src_slabs;
        if (MODEL == LF_MODEL_FREE && PAIR && a.precision != LF_PREC_F32 && a.csrc == nullptr && a.modified) {
            // FP64 walker x source loop, two walkers per lane: every broadcast source load (4 wavefronts of the
            // load-store data path for 16 B x 32 lanes) feeds two terms.  Chains: (s0, A), (s0, B), (s1, A), (s1, B).
            const double* wq = a.wp + wB;
            const double alphaA = wp[P_ALPHA * WS], alphaB = wq[P_ALPHA * WS];
            const double al[4] = {alphaA, alphaB, alphaA, alphaB};
            double accv[4] = {0.0, 0.0, 0.0, 0.0};
            int k = field_of(a, i0);
            while (i0 < i1) {
                const long long seg_end = a.field_ind[k + 1] < i1 ? a.field_ind[k + 1] : i1;
                const double aFA = wp[(P_FIELD0 + 4 * k + 0) * WS], c2A = wp[(P_FIELD0 + 4 * k + 1) * WS];
                const double aFB = wq[(P_FIELD0 + 4 * k + 0) * WS], c2B = wq[(P_FIELD0 + 4 * k + 1) * WS];
                const double af[4] = {aFA, aFB, aFA, aFB}, cc[4] = {c2A, c2B, c2A, c2B};
                auto two = [&](const double2& s0, const double2& s1) {
                    const double ux[4] = {s0.x, s0.x, s1.x, s1.x}, uy[4] = {s0.y, s0.y, s1.y, s1.y};
                    fleming_terms_v<4>(ux, uy, al, af, cc, s_exp, s_log, rep16, rep8, accv);
                };
                // groups of four sources are double-buffered: the next group's broadcast loads are issued before the
                // current group's arithmetic.  32-bit trip counter.
                const double2* __restrict__ ps = a.src2 + i0;
                const int cnt = (int)(seg_end - i0);
                double2 A0, A1, A2, A3, B0, B1, B2, B3;
                int j = 0;
                if (cnt >= 4) { A0 = __ldg(ps); A1 = __ldg(ps + 1); A2 = __ldg(ps + 2); A3 = __ldg(ps + 3); }
                for (; j + 8 <= cnt; j += 8) {
                    B0 = __ldg(ps + j + 4); B1 = __ldg(ps + j + 5); B2 = __ldg(ps + j + 6); B3 = __ldg(ps + j + 7);
                    two(A0, A1); two(A2, A3);
                    if (j + 12 <= cnt) { A0 = __ldg(ps + j + 8); A1 = __ldg(ps + j + 9); A2 = __ldg(ps + j + 10); A3 = __ldg(ps + j + 11); }
                    two(B0, B1); two(B2, B3);
                }
                if (j + 4 <= cnt) { two(A0, A1); two(A2, A3); j += 4; }
                for (; j < cnt; ++j) {
                    const double2 s0 = __ldg(ps + j);
                    double lg0, rd0;
                    fleming_log_parts<true>(s0.x, s0.y, alphaA, aFA, c2A, s_exp, s_log, rep16, rep8, lg0, rd0);
                    accv[0] = fma(lg0, rd0, accv[0]);
                    fleming_log_parts<true>(s0.x, s0.y, alphaB, aFB, c2B, s_exp, s_log, rep16, rep8, lg0, rd0);
                    accv[1] = fma(lg0, rd0, accv[1]);
                }
                i0 = seg_end;
                ++k;
            }
            acc0 = accv[0] + accv[2];
            accB = accv[1] + accv[3];
        } else if (MODEL == LF_MODEL_FREE) {
          // every other route of the free-completeness model: one walker of the lane after the other
          const long long i0s = i0;
          double& out0 = acc0;
          double& out1 = acc1;
          const int nh = PAIR && wg * 64 + 32 < count ? 2 : 1;               // warp-uniform
          for (int h = 0; h < nh; ++h) {
            const double* wp = a.wp + (h ? wB : w);
            long long i0 = i0s;
            double acc0 = 0.0, acc1 = 0.0;
            if (!LITERAL && a.csrc != nullptr) {
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
            } else {
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
                        for (long long i = i0 + loff; i < seg_end; i += lstep) {
                            double phi = schechter_literal(__ldg(&a.lum[i]), sal, Lstar, phistar);
                            double Om = om * fleming_literal(__ldg(&a.flux[i]), F50, alpha, ftau, a.modified != 0);
                            acc0 += log(phi * Om);
                        }
                    }
                    i0 = seg_end;
                    ++k;
                }
            }
            if (h == 0) { out0 = acc0; out1 = acc1; } else accB = acc0 + acc1;
          }
        } else if (MODEL == LF_MODEL_FIXED) {
            // fast class: the whole source sum is in P_LNPART0 (sufficient statistics); literal class sums terms
            if (LITERAL) {
                const double Lstar = wp[P_LSTAR * WS], phistar = wp[P_PHISTAR * WS], sal = wp[P_SCHAL * WS];
                for (long long i = i0 + loff; i < i1; i += lstep)                     // lumfuncmcmc.py:388
                    acc0 += log(schechter_literal(__ldg(&a.lum[i]), sal, Lstar, phistar) * __ldg(&a.om_arr[i]));
            }
        } else {
            const double aL = wp[P_AL * WS], bL = wp[P_BL * WS], cL = wp[P_CL * WS];
            if (!LITERAL && (a.precision == LF_PREC_F32 || a.csrc != nullptr)) {
              // FP32 mode / compressed catalogue: one walker of the pair after the other
              const int nh = wg * 64 + 32 < count ? 2 : 1;                   // warp-uniform
              for (int h = 0; h < nh; ++h) {
                const double* wq = h == 0 ? wp : a.wp + wB;
                const double aL = wq[P_AL * WS], bL = wq[P_BL * WS], cL = wq[P_CL * WS];
                double accH = 0.0;
                if (a.precision == LF_PREC_F32) {
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
                    accH -= sum;
                } else {
                    // compressed catalogue: sum_m v_m 2^(-P2(xi_m)) over this slab of pseudo-sources (compress_sources_z)
                    const double L2T = 3.32192809488736234787;
                    const double a2 = aL * L2T, b2 = bL * L2T, c2 = cL * L2T;
                    const long long m0 = (a.M * row) / a.n_src_slabs, m1 = (a.M * (row + 1)) / a.n_src_slabs;
                    for (long long m = m0; m < m1; ++m) {
                        const double xi = __ldg(&a.csrc[2 * m]).x, v = __ldg(&a.csrc[2 * m + 1]).x;
                        accH = fma(-v, exp2r_full<true>(-fma(fma(a2, xi, b2), xi, c2), s_exp_rep, lane8), accH);
                    }
                }
                if (h == 0) acc0 = accH; else accB = accH;
              }
            } else if (!LITERAL) {
                // only sum_i 10^(lum_i - L*(z_i)) needs the walker x source loop; the rest is in P_LNPART0.
                // Base 2 and centred: with u = z - z2 (middle pivot) L*(z) = aL u^2 + b' u + L*(z2), so
                // 10^(lum - L*) = 10^(42 - L*(z2)) 2^(x - (a2 u + b2) u) with x = log2(10) (lum - 42) stored per source:
                // 9 FP64 instructions per term, the factor is applied once per work item
                const double L2T = 3.32192809488736234787;                              // log2(10)
                const double zp = a.z2;
                const double* wq = a.wp + wB;                                            // the lane's second walker
                const double aLB = wq[P_AL * WS], bLB = wq[P_BL * WS], cLB = wq[P_CL * WS];
                const double a2 = aL * L2T, b2 = fma(2.0 * aL, zp, bL) * L2T;
                const double a2B = aLB * L2T, b2B = fma(2.0 * aLB, zp, bLB) * L2T;
                const double scale = exp2r_full<true>(L2T * (42.0 - fma(fma(aL, zp, bL), zp, cL)), s_exp_rep, lane8);
                const double scaleB = exp2r_full<true>(L2T * (42.0 - fma(fma(aLB, zp, bLB), zp, cLB)), s_exp_rep, lane8);
                // fast-class walkers only (list_fast): k_prologue bounds the exponent to (-800, 680), no clamp needed
                constexpr bool ZCLAMP = LF_Z_CLAMP != 0;
                // two sources x two walkers in lock-step (four independent FP64 chains per thread); groups of four sources
                // are double-buffered in two register sets (no copies): the next group's loads are issued before the
                // current group's arithmetic
                const double2* __restrict__ ps = a.src2 + i0;
                const int cnt = (int)(i1 - i0);
                double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
                auto two = [&](const double2& s0, const double2& s1) {
                    const double d0 = fma(-fma(a2, s0.y, b2), s0.y, s0.x);
                    const double d1 = fma(-fma(a2, s1.y, b2), s1.y, s1.x);
                    const double g0 = fma(-fma(a2B, s0.y, b2B), s0.y, s0.x);
                    const double g1 = fma(-fma(a2B, s1.y, b2B), s1.y, s1.x);
                    e0 += exp2r_full<ZCLAMP>(d0, s_exp_rep, lane8);
                    e1 += exp2r_full<ZCLAMP>(d1, s_exp_rep, lane8);
                    f0 += exp2r_full<ZCLAMP>(g0, s_exp_rep, lane8);
                    f1 += exp2r_full<ZCLAMP>(g1, s_exp_rep, lane8);
                };
                double2 A0, A1, A2, A3, B0, B1, B2, B3;
                int j = 0;
                if (cnt >= 4) { A0 = __ldg(ps); A1 = __ldg(ps + 1); A2 = __ldg(ps + 2); A3 = __ldg(ps + 3); }
                for (; j + 8 <= cnt; j += 8) {
                    B0 = __ldg(ps + j + 4); B1 = __ldg(ps + j + 5); B2 = __ldg(ps + j + 6); B3 = __ldg(ps + j + 7);
                    two(A0, A1); two(A2, A3);
                    if (j + 12 <= cnt) { A0 = __ldg(ps + j + 8); A1 = __ldg(ps + j + 9); A2 = __ldg(ps + j + 10); A3 = __ldg(ps + j + 11); }
                    two(B0, B1); two(B2, B3);
                }
                if (j + 4 <= cnt) { two(A0, A1); two(A2, A3); j += 4; }
                for (; j < cnt; ++j) {
                    const double2 s0 = __ldg(ps + j);
                    e0 += exp2r_full<ZCLAMP>(fma(-fma(a2, s0.y, b2), s0.y, s0.x), s_exp_rep, lane8);
                    f0 += exp2r_full<ZCLAMP>(fma(-fma(a2B, s0.y, b2B), s0.y, s0.x), s_exp_rep, lane8);
                }
                acc0 = -scale * (e0 + e1);
                accB = -scaleB * (f0 + f1);
            } else {
                const double aP = wp[P_AP * WS], bP = wp[P_BP * WS], cP = wp[P_CP * WS], sal = wp[P_SCHAL * WS];
                for (long long i = i0 + loff; i < i1; i += lstep) {                   // lumfuncmcmc_z.py:371
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
                                acc0 = fma(wt[t], expr_full(arg[t], s_exp_rep, lane8), acc0);
                                acc1 = fma(wt[t + 1], expr_full(arg[t + 1], s_exp_rep, lane8), acc1);
                            }
                        }
                        for (; j < cnt; ++j) {
                            const double2 xl = s_stage[j * 2];
                            acc0 = fma(s_stage[j * 2 + 1].x, expr_full(fma(-xl.y, cB, fma(c1, xl.x, cA)), s_exp_rep, lane8), acc0);
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
    if (LANE_SRC && is_src) {
        double v = acc0 + acc1;                                              // lanes hold disjoint sources of ONE walker
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) a.partial[(long long)row * WS + w] = v;
    } else {
        if (active) a.partial[(long long)row * WS + w] = acc0 + acc1;
        if (PAIR && activeB) a.partial[(long long)row * WS + wB] = accB;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// finish: fixed-order reduction over slabs; lnprob = lnpart - fullint
// ------------------------------------------------------------------------------------------------
#define FIN_WALKERS 8            /* walkers per block: 64-byte rows of partials per warp-load, W / 8 blocks (128 at W = 1024) */
#define FIN_WARPS 32
__global__ void __launch_bounds__(32 * FIN_WARPS) k_finish(KArgs a) {
    // one block per FIN_WALKERS walkers.  Lane = (walker, row sub-slot): a warp reads 4 consecutive rows x 8 walkers (four
    // 64-byte segments) per step, the 32 warps stride over the rows; every thread adds its rows in ascending order, the
    // partial sums are combined by a fixed shuffle tree and a fixed-order sum over the warps: deterministic, no atomics.
    // (Round 1 used 32 walkers x 32 row groups per block = W / 32 blocks: 24 us at 1763 rows x 1024 walkers, most of it the
    // serial chain of 55 dependent L2 round trips per thread on 32 of the 148 SMs.)
    __shared__ double s_src[FIN_WARPS][FIN_WALKERS], s_quad[FIN_WARPS][FIN_WALKERS];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int wsub = lane & (FIN_WALKERS - 1), rsub = lane / FIN_WALKERS;          // 8 walkers x 4 row sub-slots
    constexpr int RSTEP = FIN_WARPS * (32 / FIN_WALKERS);
    const int nf = a.cls_count[CLS_FAST], nl = a.cls_count[CLS_LIT];
    const long long idx = (long long)blockIdx.x * FIN_WALKERS + wsub;
    if ((long long)blockIdx.x * FIN_WALKERS >= nf + nl) return;
    const bool valid = idx < nf + nl;
    const bool fast = idx < nf;
    long long w = 0;
    if (valid) w = fast ? a.list_fast[idx] : a.list_lit[idx - nf];
    const long long WS = a.Wcap;
    double lnpart = 0.0, fullint = 0.0;
    if (valid) {
        const double* col = a.partial + w;
        const int r0 = g * (32 / FIN_WALKERS) + rsub;
#pragma unroll 4
        for (int r = r0; r < a.n_src_slabs; r += RSTEP) lnpart += col[(long long)r * WS];
#pragma unroll 4
        for (int r = a.n_src_slabs + r0; r < a.n_src_slabs + a.n_quad_slabs; r += RSTEP) fullint += col[(long long)r * WS];
    }
    for (int o = FIN_WALKERS; o < 32; o <<= 1) {                                    // over the row sub-slots of the warp
        lnpart += __shfl_xor_sync(0xffffffffu, lnpart, o);
        fullint += __shfl_xor_sync(0xffffffffu, fullint, o);
    }
    if (rsub == 0) { s_src[g][wsub] = lnpart; s_quad[g][wsub] = fullint; }
    __syncthreads();
    if (g != 0 || rsub != 0 || !valid) return;
    lnpart = 0.0; fullint = 0.0;
    for (int k = 0; k < FIN_WARPS; ++k) { lnpart += s_src[k][wsub]; fullint += s_quad[k][wsub]; }
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
    // fast FP64 loop: (log2(10) (lum - 42), z - z_pivot2): the walker's quadratic L*(z) is evaluated about the middle
    // pivot, its constant term leaves the loop as one factor per walker
    src2[i] = make_double2((lum[i] - 42.0) * 3.32192809488736234787, z[i] - zref);
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
static int ndim_of(const lf_config& c) {
    int free_al = c.fix_sch_al ? 0 : 1;
    if (c.model == LF_MODEL_FREE) return 2 + free_al + c.nfields + 1;
    if (c.model == LF_MODEL_FIXED) return 2 + free_al;
    return 6 + free_al;
}

extern "C" int lf_ndim(const lf_ctx* ctx) { return ctx ? ctx->ndim : -1; }

static void fill_tables(Tables& t) {
    for (int j = 0; j < EXP_TAB_N; ++j) t.exp2_frac[j] = (double)exp2l((long double)j / EXP_TAB_N);
    for (int i = 0; i < EXPB_N; ++i) t.exp2_big[i] = (double)((long double)EXP_TAB_SCALE * exp2l((long double)(EXPB_KMIN + i) / EXP_TAB_N));
    t.exp2_big[EXPB_N] = EXP_TAB_SCALE;
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
    CK(cudaEventCreateWithFlags(&c->ev_scratch, cudaEventDisableTiming));
    Tables t;
    fill_tables(t);
    CK(cudaMalloc(&c->d_tables, sizeof(Tables)));
    CK(cudaMemcpy(c->d_tables, &t, sizeof(Tables), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_cls, 8 * sizeof(int)));
    CK(cudaFuncSetAttribute(k_main<false, LF_MODEL_FREE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(LF_MODEL_FREE)));
    CK(cudaFuncSetAttribute(k_main<false, LF_MODEL_FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(LF_MODEL_FIXED)));
    CK(cudaFuncSetAttribute(k_main<false, LF_MODEL_Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)main_smem_bytes(LF_MODEL_Z)));
    if (veff_init(c)) return 1;
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

extern "C" void lf_destroy(lf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    dfree(c->d_lum); dfree(c->d_flux); dfree(c->d_z); dfree(c->d_om); dfree(c->d_Lsrc); dfree(c->d_src2); dfree(c->d_src2f); dfree(c->d_csrc);
    dfree(c->d_qpf); dfree(c->d_qp); dfree(c->d_zarr); dfree(c->d_tables);
    dfree(c->d_wp); dfree(c->d_colA); dfree(c->d_colB); dfree(c->d_partial);
    dfree(c->d_cls); dfree(c->d_list_fast); dfree(c->d_list_lit); dfree(c->d_list_fastq); dfree(c->d_list_litq); dfree(c->d_thetas); dfree(c->d_out);
    dfree(c->v_lum); dfree(c->v_phi); dfree(c->v_edges); dfree(c->v_counts); dfree(c->v_sums);
    dfree(c->v_outc); dfree(c->v_outs); dfree(c->v_mult); dfree(c->v_bin);
    dfree(c->v_flux); dfree(c->v_vol); dfree(c->v_valid); dfree(c->v_u); dfree(c->v_rowcounts); dfree(c->v_ticket); dfree(c->v_mt_state); dfree(c->v_mt_vals);
    dfree(c->v_cum); dfree(c->v_gl); dfree(c->v_zk); dfree(c->v_dVk); dfree(c->v_cumV);
    if (c->ev_scratch) cudaEventDestroy(c->ev_scratch);
    peer_release(c);
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
    DevBufs scratch;                               // released on every exit path
    CK(scratch.alloc(&d_tmp, sizeof(double) * 3 * 1024));
    CK(scratch.alloc(&d_scratch, nb));
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
        // The reference hands every field the SAME luminosity grid (lumfuncmcmc.py:232 appends one array object for all
        // fields), and the FIXED / Z integrand depends on the field only through integ_part: when the K
        // grids are bit-identical the K weight planes are summed once here and the walkers integrate S^2 points instead
        // of K S^2.  (sum_k trapz(trapz(phi integ_k)) == trapz(trapz(phi sum_k integ_k)) up to the order of additions.)
        bool merged = K > 1;
        for (int k = 1; k < K && merged; ++k) merged = memcmp(logL + (size_t)k * SS, logL, sizeof(double) * (size_t)SS) == 0;
        const int Kq = merged ? 1 : K;
        const long long NQe = SS * Kq;
        c->NQ = NQe; a.NQ = NQe;
        std::vector<QuadPoint> pts((size_t)NQe);
        for (int k = 0; k < Kq; ++k)
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
                    if (merged) {
                        long double sum = 0.0L;
                        for (int kk = 0; kk < K; ++kk) sum += (long double)at(integ_part, kk, j, i);
                        p.wt = wl * wz[i] * (double)sum;
                    } else {
                        p.wt = wl * wz[i] * at(integ_part, k, j, i);
                    }
                    p.pad = 0.0;
                }
        CK(cudaMalloc(&c->d_qp, sizeof(QuadPoint) * (size_t)NQe));
        CK(cudaMemcpy(c->d_qp, pts.data(), sizeof(QuadPoint) * (size_t)NQe, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&c->d_zarr, sizeof(double) * S));
    CK(cudaMemcpy(c->d_zarr, zarr, sizeof(double) * S, cudaMemcpyHostToDevice));
    a.qpf = c->d_qpf; a.qp = c->d_qp; a.zarr = c->d_zarr;
    c->have_grid = true;
    return 0;
}

__global__ void k_derive_compressed(long long m, const double* __restrict__ xi, const double* __restrict__ w, double fcap,
                                    double2* __restrict__ out, bool zmodel) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double g = xi[i];
    out[2 * i] = make_double2(g, zmodel ? 0.0 : fmin(pow(10.0, g), fcap));
    out[2 * i + 1] = make_double2(w[i], 0.0);
}

extern "C" int lf_set_compressed_sources(lf_ctx* c, int64_t M, const double* xi, const double* w, const int64_t* cfield_ind,
                                         double alpha_max) {
    if (!c) return fail("lf_set_compressed_sources: null context");
    if (c->cfg.model == LF_MODEL_FIXED) return fail("lf_set_compressed_sources: the fixed-completeness source sum is already O(1) per walker");
    if (c->cfg.precision != LF_PREC_F64) return fail("lf_set_compressed_sources: FP64 only");
    if (!c->have_sources) return fail("lf_set_compressed_sources: call lf_set_sources first");
    CK(cudaSetDevice(c->device));
    dfree(c->d_csrc);
    c->ka.csrc = nullptr; c->ka.M = 0;
    if (M == 0) return 0;
    if (M < 0 || !xi || !w || !cfield_ind || !(alpha_max > 0.0)) return fail("lf_set_compressed_sources: bad arguments");
    const int K = c->cfg.nfields;
    if (cfield_ind[0] != 0 || cfield_ind[K] != M) return fail("lf_set_compressed_sources: cfield_ind must run from 0 to M");
    for (int k = 0; k < K; ++k) {
        if (cfield_ind[k + 1] < cfield_ind[k]) return fail("lf_set_compressed_sources: cfield_ind must be non-decreasing");
        if ((cfield_ind[k + 1] > cfield_ind[k]) != (c->ka.fs[k].n > 0.0)) return fail("lf_set_compressed_sources: a field has sources but no pseudo-sources (or the reverse)");
    }
    // every node must lie inside the range of fluxes the classification bounds were computed for
    for (int k = 0; k < K; ++k)
        for (int64_t m = cfield_ind[k]; m < cfield_ind[k + 1]; ++m) {
            const bool inside = c->cfg.model == LF_MODEL_FREE ? xi[m] >= c->ka.fs[k].g_min - 1.0e-9
                                                              : (xi[m] >= c->ka.fs[k].z_min - 1.0e-9 && xi[m] <= c->ka.fs[k].z_max + 1.0e-9);
            if (!inside || !(xi[m] == xi[m])) return fail("lf_set_compressed_sources: a node lies outside the range of the field's sources");
        }
    double *d_xi = nullptr, *d_w = nullptr;
    DevBufs tmp;
    CK(tmp.alloc(&d_xi, sizeof(double) * M));
    CK(tmp.alloc(&d_w, sizeof(double) * M));
    CK(cudaMalloc(&c->d_csrc, sizeof(double2) * 2 * (size_t)M));
    CK(cudaMemcpy(d_xi, xi, sizeof(double) * M, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w, sizeof(double) * M, cudaMemcpyHostToDevice));
    k_derive_compressed<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(M, d_xi, d_w, c->ka.fcap, c->d_csrc, c->cfg.model == LF_MODEL_Z);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->ka.csrc = c->d_csrc; c->ka.M = M; c->ka.c_alpha_max = alpha_max;
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
    if ((W > c->Wcap || rows > c->rows_cap) && c->scratch_pending) {
        // the scratch is about to be re-allocated: whoever still uses it (on any stream) has to finish first
        CK(cudaStreamSynchronize(c->scratch_stream));
        c->scratch_pending = false;
    }
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
    static const int items_per_slot = []() { const char* e = getenv("LF_PLAN_ITEMS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 24; }();
    const long long target_items = (long long)c->sm_count * 16 * items_per_slot;   // ~24-32 items per resident warp (LF_PLAN_ITEMS: tuning)
    long long rows = std::min<long long>(items_per_slot > 24 ? 16384 : 4096, std::max<long long>(1, target_items / n_wg));
    const int model = c->cfg.model;
    // relative cost of a quadrature point vs a source term
    double src_cost = model == LF_MODEL_FREE ? 1.0 : (model == LF_MODEL_Z ? 0.5 : 0.0);
    double quad_cost = model == LF_MODEL_FREE ? 1.5 : 0.6;
    const long long n_eff = c->d_csrc ? c->ka.M : c->N;      // pseudo-sources when compressed
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

static bool stream_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

int scratch_acquire(lf_ctx* c, cudaStream_t st) {
    if (c->scratch_pending && c->scratch_stream != st && !stream_capturing(st))
        CK(cudaStreamWaitEvent(st, c->ev_scratch, 0));
    return 0;
}

int scratch_release(lf_ctx* c, cudaStream_t st) {
    if (stream_capturing(st)) return 0;            // graph replays run on the context's own stream, in order
    CK(cudaEventRecord(c->ev_scratch, st));
    c->scratch_stream = st; c->scratch_pending = true;
    return 0;
}

int launch_pipeline(lf_ctx* c, const double* d_thetas, long long W, double* d_out, cudaStream_t st) {
    if (!c->have_sources || !c->have_grid) return fail("lf_lnprob: call lf_set_sources and lf_set_grid first");
    if (W <= 0) return 0;
    int n_src, n_quad;
    plan_rows(c, W, n_src, n_quad);
    if (ensure_scratch(c, W, n_src + n_quad)) return 1;
    if (scratch_acquire(c, st)) return 1;
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
    // literal source items are one WALKER (not one group of 32) x one slab each: up to W n_src of them
    const long long need_lit = (W * n_src + n_wg * n_quad + wpb - 1) / wpb;
    const unsigned bl = (unsigned)std::min<long long>(need_lit, (long long)c->sm_count * c->occ_lit);
    {
        // fast kernel: launched programmatically dependent on the kernel before it (prologue, or k_zcolumns for the Z model)
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(bf); lc.blockDim = dim3(32 * main_warps(c->cfg.model)); lc.dynamicSmemBytes = main_smem_bytes(c->cfg.model);
        lc.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = LF_PDL;
        lc.attrs = attr; lc.numAttrs = 1;
        if (c->cfg.model == LF_MODEL_FREE) {
            CK(cudaLaunchKernelEx(&lc, k_main<false, LF_MODEL_FREE>, a));
            k_main<true, LF_MODEL_FREE><<<bl, 32 * main_warps(LF_MODEL_FREE), 0, st>>>(a);
        } else if (c->cfg.model == LF_MODEL_FIXED) {
            CK(cudaLaunchKernelEx(&lc, k_main<false, LF_MODEL_FIXED>, a));
            k_main<true, LF_MODEL_FIXED><<<bl, 32 * main_warps(LF_MODEL_FIXED), 0, st>>>(a);
        } else {
            CK(cudaLaunchKernelEx(&lc, k_main<false, LF_MODEL_Z>, a));
            k_main<true, LF_MODEL_Z><<<bl, 32 * main_warps(LF_MODEL_Z), 0, st>>>(a);
        }
    }
    k_finish<<<(unsigned)((W + FIN_WALKERS - 1) / FIN_WALKERS), 32 * FIN_WARPS, 0, st>>>(a);
    c->launches += 3;
    CK(cudaGetLastError());
    return scratch_release(c, st);
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

