/*
 * lf_engine.h -- C ABI of the B200 likelihood engine for LumFuncMCMC's hot path.
 *
 * One shared library (lumfuncmcmc_b200/csrc/liblfengine.so, sm_100a) exports exactly these symbols.  They are
 * what a binding inside the reference would call in place of its NumPy arithmetic; INTEGRATION.md shows the
 * ctypes stub.  Plain pointers and sizes only; no torch / Python types.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   lf_lnprob_batch            <- LumFuncMCMC.lnprob / lnprob_fix_comp   lumfuncmcmc.py:395-424
 *                                 (= set_parameters_from_list :320-337, lnprior :339-358,
 *                                    lnlike :360-378, lnlike_fix_comp :380-393), one row of `thetas` per call
 *                                 of the reference; LumFuncMCMCz.lnprob   lumfuncmcmc_z.py:378-392
 *                                 (= :332-341, :343-362, :364-376)
 *   lf_set_sources/lf_set_grid <- the arrays those methods read from `self`, produced once by
 *                                 setDLdVdz/setOmegaLz/setlnsimple/defineFlimOmArr  lumfuncmcmc.py:180-235, 283-288
 *   lf_veff_bin                <- LumFuncMCMC.VeffLF lumfuncmcmc.py:515-525 -> V.lumfunc VmaxLumFunc.py:215-257
 *                                 and the original-sample binning of V.getBootErrLog VmaxLumFunc.py:336-350
 *   lf_boot_bin                <- one bootstrap replicate of V.getBootErrLog VmaxLumFunc.py:352-359
 *
 * Conventions: every function returns 0 on success, non-zero on error; the message is available from
 * lf_last_error() (thread-local).  Numerical outcomes (-inf) are values, not errors; NaN is never returned by
 * lf_lnprob_*.  The caller owns all host buffers; lf_set_* copy to the device and the context owns device memory
 * until lf_destroy.  A context is single-caller (like the reference, whose lnprob mutates `self`).
 * There is no CPU fallback: without a CUDA device lf_create fails.
 */
#ifndef LF_ENGINE_H
#define LF_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LF_MAX_FIELDS 16

/* model kind: which reference likelihood the context evaluates */
enum { LF_MODEL_FREE = 0,   /* lnlike: completeness parameters sampled   (lumfuncmcmc.py:360-378) */
       LF_MODEL_FIXED = 1,  /* lnlike_fix_comp: tabulated Omega          (lumfuncmcmc.py:380-393) */
       LF_MODEL_Z = 2 };    /* LumFuncMCMCz.lnlike: evolving L*, phi*    (lumfuncmcmc_z.py:364-376) */

/* arithmetic of the walker x source loop (statistics, quadrature, prior and the literal kernels stay FP64) */
enum { LF_PREC_F64 = 0, LF_PREC_F32 = 1 };

typedef struct lf_ctx lf_ctx;

typedef struct lf_config {
    int32_t model;           /* LF_MODEL_*                                                            */
    int32_t precision;       /* LF_PREC_*                                                             */
    int32_t device;          /* CUDA device ordinal                                                   */
    int32_t nfields;         /* K <= LF_MAX_FIELDS                                                    */
    int32_t size_ln;         /* S: quadrature grid is S x S per field (101 free / 201 fixed, z)       */
    int32_t fix_sch_al;      /* 1: Schechter alpha is not in theta, use `sch_al` below                */
    int32_t fixed_prior_ok;  /* 1 if the parameters NOT in theta lie inside their prior boxes: the
                                reference's lnprior range-checks those too (lumfuncmcmc.py:347-354)   */
    int32_t force_literal;   /* testing: 1 = evaluate every walker with the literal (reference-order)
                                kernels instead of the fast kernels                                   */
    double fcmin;            /* modified-Fleming threshold; 0 selects the plain curve (VmaxLumFunc.py:121) */
    double sch_al;           /* value used when fix_sch_al                                            */
    double Lstar_lims[2], phistar_lims[2], sch_al_lims[2], Flim_lims[2], alpha_lims[2];
    double z_pivots[3];      /* LF_MODEL_Z: z1, z2, z3 (lumfuncmcmc_z.py:191)                         */
} lf_config;

/* LambdaCDM parameters + quadrature rule of the set-up / volume entry points (see lf_cosmo_distances) */
typedef struct lf_cosmology {
    double H0, Om0, Ode0, Or0, Ok0, panel;
    double gl_x[8], gl_w[8];     /* Gauss-Legendre nodes / weights on [-1, 1], as numpy.polynomial.legendre.leggauss(8) */
} lf_cosmology;

/* theta layout (row-major, `ndim` doubles per walker), as set_parameters_from_list:
 *   FREE : L*, phi*, [alpha_s], F50_0..F50_{K-1}, alpha_c      ndim = 2 + !fix_sch_al + K + 1
 *   FIXED: L*, phi*, [alpha_s]                                  ndim = 2 + !fix_sch_al
 *   Z    : L1, L2, L3, phi1, phi2, phi3, [alpha_s]              ndim = 6 + !fix_sch_al           */
int lf_ndim(const lf_ctx* ctx);

int lf_create(lf_ctx** out, const lf_config* cfg);
void lf_destroy(lf_ctx* ctx);

/* Per-source arrays (host pointers, n = field_ind[K]; sources sorted by field).
 *   lum      log10 L_i                                      all models
 *   flux     10^lum_i / (4 pi (3.086e24 DLf(z_i))^2)        FREE   (lumfuncmcmc.py:69-70; caller evaluates DLf once)
 *   z        redshift                                        Z
 *   om_arr   Om_arr_i (tabulated Omega, lumfuncmcmc.py:235)  FIXED, Z
 *   omega0_int  per-field area truncated to integer           FREE   (dtype=int copy, lumfuncmcmc.py:285)
 * Unused pointers may be NULL. */
int lf_set_sources(lf_ctx* ctx, int64_t n, const double* lum, const double* flux, const double* z,
                   const double* om_arr, const int64_t* field_ind, const int64_t* omega0_int);

/* Quadrature grid (host pointers).  logL[K][S][S] (row = luminosity index, column = redshift index),
 * zarr[S], DL_zarr[S] = DLf(zarr) in Mpc, volume_part[S], omega0[K] (float areas) for FREE;
 * integ_part[K][S][S] for FIXED / Z (DL_zarr, volume_part, omega0 may then be NULL). */
int lf_set_grid(lf_ctx* ctx, const double* logL, const double* zarr, const double* DL_zarr,
                const double* volume_part, const double* integ_part, const double* omega0);

/* Optional (LF_MODEL_FREE, LF_MODEL_Z): replace the fast kernels' walker x source loop by a weighted sum over M pseudo-sources,
 *   sum_i t(g_i)  ->  sum_m w[m] t(xi[m]),   t(g) = ln fc(alpha_c (g - log10 F50)) / (1 - exp(-10^g / f_tau)),
 * with (xi, w) built from the catalogue by lumfuncmcmc_b200/compress.py (piecewise Chebyshev interpolation of t in
 * g = log10 flux; truncation ~1e-15 per source for every alpha_c <= the alpha_max it was built for).  cfield_ind[K+1]
 * delimits the pseudo-sources of each field.  lf_set_sources must have been called with the real sources first: their
 * sufficient statistics, the classification bounds and the literal kernels keep using them.  Walkers with
 * alpha_c > alpha_max (possible only when the prior gate is off) are evaluated by the literal kernels on the real
 * sources.  LF_MODEL_Z: xi are redshift nodes, w[m] = sum_i 10^lum_i l_m(z_i) (compress_sources_z), the term is
 * sum_m w[m] 10^(-L*(xi[m])), and `alpha_max` is the largest |d(L_star)/dz| the weights are accurate for.  M = 0 switches back. */
int lf_set_compressed_sources(lf_ctx* ctx, int64_t M, const double* xi, const double* w, const int64_t* cfield_ind,
                              double alpha_max);

/* Which walkers' quadrature term this context subtracts: walkers w with w % nshare == share.
 * Default (0, 1) = all.  Multi-GPU source sharding: rank r calls (r, world) and the per-rank outputs are
 * summed (all-reduce) to give lnprob. */
int lf_set_quadrature_share(lf_ctx* ctx, int32_t share, int32_t nshare);

/* enabled = 1 (default): lf_lnprob_* returns -inf outside the prior box (lnprob).  enabled = 0: the prior is
 * not consulted, i.e. the call evaluates lnlike() at the given parameters (lumfuncmcmc.py:360-393). */
int lf_set_prior_gate(lf_ctx* ctx, int32_t enabled);

/* Batched log-posterior, HOST buffers: thetas[W][ndim] -> out[W].  Includes H2D, kernels, D2H, sync. */
int lf_lnprob_batch(lf_ctx* ctx, const double* thetas, int64_t W, double* out);

/* Same with DEVICE buffers on a caller-provided CUDA stream (cudaStream_t as void*); asynchronous. */
int lf_lnprob_batch_device(lf_ctx* ctx, const double* d_thetas, int64_t W, double* d_out, void* stream);

/* Per-walker evaluation class of the last call: counts[0] = rejected before any source is read (prior or
 * certain underflow), counts[1] = fast kernels, counts[2] = literal kernels.  `launches` = kernels launched. */
int lf_last_call_info(lf_ctx* ctx, int64_t counts[3], int64_t* launches);

/* 1/V_eff weights + binned LF over the original sample (host buffers).
 *   phi_i = [zmax_i > zmin] / (sum_omega/sqarcsec * fleming(flux_i; 1e-17*flim[field(i)], alpha, fcmin) * vol_i)
 *   vol_i = vol_int (shared) when vol_per_source == NULL, else vol_per_source[i]
 *   counts[j], sumphi[j] over the half-open bins [edges[j], edges[j+1]) in lum.
 * phi_out may be NULL.  The engine keeps lum/phi resident for lf_boot_bin. */
int lf_veff_bin(lf_ctx* ctx, int64_t n, const double* flux, const double* lum, const int64_t* field_ind,
                int32_t nfields, const double* flim, double alpha, double fcmin, double sum_omega,
                double vol_int, const double* vol_per_source, const uint8_t* valid_or_null,
                const double* edges, int32_t nbins, double* phi_out, int64_t* counts, double* sumphi);

/* The same pass for a sample that STAYS on the device between calls -- VeffLF runs after every fit and once more per
 * posterior summary with new completeness parameters on the same catalogue (lumfuncmcmc.py:541, :567, :650):
 *   lf_veff_set_sample     upload flux (cgs), lum and the field offsets once
 *   lf_veff_bin_resident   weights + binning on the resident sample: nothing per-source crosses PCIe unless phi_out is
 *                          given (phi stays resident: lf_veff_get_phi downloads it on demand, lf_bin_weights(NULL, NULL)
 *                          and lf_boot_bin* read it in place).  use_device_volumes = 1 takes the per-source volume and
 *                          validity lf_veff_volumes left on the device (vol_int is then ignored).
 *   lf_veff_get_phi        the resident per-source weights -> host */
int lf_veff_set_sample(lf_ctx* ctx, int64_t n, const double* flux, const double* lum, const int64_t* field_ind,
                       int32_t nfields);
int lf_veff_bin_resident(lf_ctx* ctx, const double* flim, double alpha, double fcmin, double sum_omega, double vol_int,
                         int32_t use_device_volumes, const double* edges, int32_t nbins, double* phi_out,
                         int64_t* counts, double* sumphi);
int lf_veff_get_phi(lf_ctx* ctx, double* phi_out);

/* min_comp_frac > 0 (lumfuncmcmc.py:521-524): every source of the resident sample integrates dV/dz only up to the
 * redshift where its luminosity drops to its field's minimum flux.  The reference finds that redshift with one fsolve
 * per source (V.getMaxz, VmaxLumFunc.py:739-753: 4 pi (D_L(z) cm)^2 fmin = 10^lum, xtol 1.5e-8) and the volume with one
 * QUADPACK call per source on the linear interpolant dVdzf (V.lumfunc, VmaxLumFunc.py:235-257, 1.5e-8).  Here:
 *   zmax_i  = min(zmax, root)   by Newton on the D_L of lf_cosmo_distances (same arithmetic, converged to ~1e-15)
 *   vol_i   = int_zmin^zmax_i dVdzf dz   EXACTLY (cumulative trapezoids of the piecewise-linear interpolant + partial segment)
 *   valid_i = zmax_i > zmin
 * so the two agree to the reference's own solver tolerances (tested at 1e-7 on weights, exact on counts).
 *   lf_veff_set_volume_table  cosmology (as lf_cosmo_distances) + the knots (zk, dVk) of dVdzf; kept until replaced
 *   lf_veff_volumes           fmin[K] = minimum flux per field (cgs); DL_zmin / DL_zmax = D_L [Mpc] at zmin / zmax;
 *                             results stay resident for lf_veff_bin_resident(use_device_volumes = 1); the three host
 *                             outputs (n each) may be NULL */
int lf_veff_set_volume_table(lf_ctx* ctx, const lf_cosmology* cosmo, const double* cum, int64_t ncum, int64_t nk,
                             const double* zk, const double* dVk);
int lf_veff_volumes(lf_ctx* ctx, double zmin, double zmax, double DL_zmin, double DL_zmax, const double* fmin,
                    double* zmax_out, double* vol_out, uint8_t* valid_out);

/* Bin caller-provided weights: counts[j], sumphi[j] of (lum_i, phi_i) over [edges[j], edges[j+1]); keeps lum/phi
 * resident for lf_boot_bin.  This is the original-sample pass of V.getBootErrLog when phi is an input
 * (VmaxLumFunc.py:345-350).  lum == NULL and phi == NULL: bin the sample and weights already resident (n is ignored). */
int lf_bin_weights(lf_ctx* ctx, int64_t n, const double* lum, const double* phi, const double* edges,
                   int32_t nbins, int64_t* counts, double* sumphi);

/* One bootstrap replicate on the resident sample: mult[i] = how many times source i was drawn.
 * sumphi[j] = sum_i mult[i] * phi_i over bin j; counts[j] = sum_i mult[i]. */
int lf_boot_bin(lf_ctx* ctx, const int32_t* mult, int64_t* counts, double* sumphi);

/* The same with the resampling done on the device: replicate `replicate` draws n indices from a Philox4x32-10 stream
 * keyed by `seed` (counter = draw index, replicate), accumulates the multiplicities with integer atomics and bins them.
 * Statistically equivalent to VmaxLumFunc.py:352-359, not NumPy's MT19937 stream (use lf_boot_bin with host-drawn
 * multiplicities for that).  Integer counts do not depend on the order of the atomics, so a replicate is reproducible. */
int lf_boot_bin_device(lf_ctx* ctx, uint64_t seed, int64_t replicate, int64_t* counts, double* sumphi);

/* The reference's resampling stream itself, generated on the device: VmaxLumFunc.py:353 draws np.random.randint(n, size=n)
 * from NumPy's legacy global MT19937 generator (per sample: 32-bit outputs masked with the smallest 2^k - 1 >= n - 1 until one
 * is <= n - 1).  lf_boot_mt_set_state takes that generator's state (np.random.get_state(): key[624], pos), lf_boot_bin_mt
 * performs ONE replicate -- n accepted draws -> multiplicities -> binning, all on the device -- and leaves the state where
 * NumPy's would be; lf_boot_mt_get_state returns it so the host generator can be re-synchronised (np.random.set_state).
 * Replicates are bit-identical to lf_boot_bin with host-drawn multiplicities, without n host draws per replicate. */
int lf_boot_mt_set_state(lf_ctx* ctx, const uint32_t* key, int32_t pos);
int lf_boot_mt_get_state(lf_ctx* ctx, uint32_t* key_out, int32_t* pos_out);
int lf_boot_bin_mt(lf_ctx* ctx, int64_t* counts, double* sumphi);

/* Register-only FP64 FMA micro-benchmark on the context's device: sustained DFMA thread-instructions / s. */
int lf_fp64_peak(lf_ctx* ctx, int32_t iters, double* dfma_per_s, double* ms);

/* Same for the MUFU (SFU) pipe: sustained ex2.approx.f32 thread-instructions / s (roofline of LF_PREC_F32). */
int lf_mufu_peak(lf_ctx* ctx, int32_t iters, double* mufu_per_s, double* ms);

/* ---- multi-GPU exchange over peer memory (one process per GPU on one NVLink/NVSwitch node) ----
 * The only exchange of the path is the sum over ranks of the W per-walker partial log-posteriors (lumfuncmcmc.py:370 is a
 * plain sum over sources, so source shards add up).  Instead of an NCCL call it can run as one kernel of this library:
 * every rank stores its vector into a slot of every peer's receive buffer (P2P stores), raises a flag, waits for the
 * peers' flags and adds the slots in rank order -- the result is identical on all ranks and independent of timing, the
 * kernel is capturable in a CUDA graph, and no host round trip or second library sits between k_finish and the sampler.
 *   lf_peer_buffer_create   allocate this rank's receive buffer for vectors of up to wcap doubles; returns its CUDA IPC handle
 *   lf_peer_buffer_connect  open the world handles (row r = rank r's handle, own row ignored)
 *   lf_allreduce_device     in-place SUM over ranks of d_vec[W] on `stream` (asynchronous)
 *   lf_peer_status          0, or non-zero if a wait for a peer has timed out.  The flag is STICKY: the exchange that
 *                           timed out and every later one return NaN in d_vec (never a stale partial sum),
 *                           lf_allreduce_device / lf_sampler_run fail, until lf_peer_reset acknowledges it
 *   lf_peer_set_timeout     bound of one wait in seconds (default ~30 s: first-call module loads and host pauses fit)
 * All ranks must create their buffers with the same world and wcap (the slot offsets are computed from them); the
 * Python binding all-gathers (rank, world, wcap) with the handles and refuses a mismatch.  An all-zero handle row in
 * lf_peer_buffer_connect means "no peer mapped for that rank" (used by the time-out test). */
int lf_peer_buffer_create(lf_ctx* ctx, int32_t rank, int32_t world, int64_t wcap, unsigned char handle_out[64]);
int lf_peer_buffer_connect(lf_ctx* ctx, const unsigned char* handles);
int lf_allreduce_device(lf_ctx* ctx, double* d_vec, int64_t W, void* stream);
int lf_peer_status(lf_ctx* ctx, int32_t* timed_out);
int lf_peer_reset(lf_ctx* ctx);
int lf_peer_set_timeout(lf_ctx* ctx, double seconds);

/* ---- set-up tables on the GPU (the step before the path; reference lumfuncmcmc.py:180-202, VmaxLumFunc.py:14-17) ----
 * Context-free: they take a device ordinal and host buffers.
 *
 * FLRW distances of LambdaCDM(H0, Om0, Ode0, radiation Or0, curvature Ok0 = 1 - Om0 - Ode0 - Or0) for n redshifts:
 * luminosity distance [Mpc] and differential comoving volume [Mpc^3 / sr]; either output may be NULL.  Same arithmetic
 * as lumfuncmcmc_b200/cosmology.py (cumulative 8-point Gauss-Legendre panels of width `panel` + one 8-point closure per
 * redshift), operation for operation, so the two agree to the last ulp of sin/sinh.  cum[ncum] is the host's
 * cumulative panel integral of dz/E (cum[p] = int_0^{p*panel}). */
int lf_cosmo_distances(int32_t device, const lf_cosmology* cosmo, const double* cum, int64_t ncum, int64_t n,
                       const double* z, double* DL_Mpc, double* dVdz);

/* y[i] = numpy.interp(x[i], xk, yk) for increasing knots xk[nk], BIT-IDENTICAL to NumPy's arithmetic
 * (slope * (x - xk[j]) + yk[j] with the slope formed per call, exact-knot and last-knot special cases); this is what
 * scipy.interpolate.interp1d(kind='linear') evaluates (lumfuncmcmc.py:196-197).  x outside [xk[0], xk[nk-1]] is an
 * error (interp1d raises). */
int lf_interp_linear(int32_t device, int64_t nk, const double* xk, const double* yk, int64_t n, const double* x, double* y);

/* Per-source tabulated Omega, the array lumfuncmcmc.py:235 builds with Omega(lum, z, DLf, Omega_0_arr, 1e-17 Flims_arr, alpha,
 * fcmin) (lumfuncmcmc.py:47-70 -> V.fleming): om_out[i] = omega0_int[k] / sqarcsec * fleming(10^lum_i / (4 pi (3.086e24 DLf(z_i))^2);
 * 1e-17 flim[k], alpha, fcmin), k = field of source i, DLf = numpy.interp through the knots (zk, DLk).  Reference order of
 * operations with libdevice pow / log10 / exp / sqrt: within ~1e-15 of NumPy's values (not bit-identical). */
int lf_omega_sources(int32_t device, int64_t n, const double* lum, const double* z, const int64_t* field_ind,
                     int32_t nfields, const int64_t* omega0_int, const double* flim, double alpha, double fcmin,
                     int64_t nk, const double* zk, const double* DLk, double* om_out);

/* Device-resident affine-invariant ensemble sampler: replaces emcee.EnsembleSampler(...).run_mcmc(pos, nsteps)
 * (lumfuncmcmc.py:489-491, lumfuncmcmc_z.py:444-446) for a context on one GPU.  Goodman & Weare stretch move with
 * scale `a` (emcee's default 2), fixed split of the ensemble into walkers [0, W/2) and [W/2, W) as in emcee 2.x,
 * Philox4x32-10 counter RNG keyed by `seed` (counter = walker, step, half), proposals / log-posterior / accept all on
 * the device, one CUDA graph per ensemble update replayed nsteps times.  W must be even.
 *   pos0[W][ndim]            initial positions (host)
 *   chain[nsteps][W][ndim]   positions after every update (host, may be NULL)
 *   lnprob[nsteps][W]        log-posterior of those positions (host, may be NULL)
 *   naccepted[W]             accepted proposals per walker (host, may be NULL)
 *   step0                    index of the first update (continuing a run: pass the number of updates already done)
 * Returns the final positions in pos_out[W][ndim] and lnprob_out[W] (host, may be NULL). */
int lf_sampler_run(lf_ctx* ctx, const double* pos0, int64_t W, int64_t nsteps, uint64_t seed, double a, int64_t step0,
                   double* chain, double* lnprob, int64_t* naccepted, double* pos_out, double* lnprob_out);

/* Several ranks, small catalogue (SURVEY.md 8e "shard walkers when the source count is small"): every rank holds ALL sources
 * (lf_set_sources with the whole catalogue, quadrature share (0, 1)) and, with walker sharding enabled and the peer buffers
 * connected, lf_sampler_run lets rank r evaluate walkers [nw r / world, nw (r + 1) / world) of every half-ensemble; the other
 * entries are exact zeros, so the rank-ordered sum of lf_allreduce_device is the all-gather of the slices.  Every rank ends
 * with the same chain. */
int lf_set_walker_sharding(lf_ctx* ctx, int32_t enabled);

/* Device time (ms) of the nsteps graph replays of the last lf_sampler_run call. */
int lf_sampler_last_ms(lf_ctx* ctx, double* ms);

/* Device time (ms, CUDA events on the engine's stream) of the kernels of the last lf_lnprob_batch call. */
int lf_last_kernel_ms(lf_ctx* ctx, double* ms);

/* Number of CUDA devices visible to the library (0 when there is none or the driver is missing). */
int lf_device_count(void);

const char* lf_last_error(void);
const char* lf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LF_ENGINE_H */
