// lf_setup.cu -- set-up tables on the GPU: LambdaCDM distances and NumPy-exact linear
// interpolation (reference lumfuncmcmc.py:180-202, VmaxLumFunc.py:14-17).
#include "lf_internal.cuh"

// ------------------------------------------------------------------------------------------------
// set-up tables on the GPU: cosmology distances and NumPy-exact linear interpolation
// ------------------------------------------------------------------------------------------------
__global__ void k_cosmo(lf_cosmology c, const double* __restrict__ cum, long long ncum, long long n,
                        const double* __restrict__ z, const double* __restrict__ glx, const double* __restrict__ glw,
                        double* __restrict__ DL, double* __restrict__ dV, int* __restrict__ bad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    double dm, dc;
    if (!cosmo_dm(c, cum, ncum, glx, glw, zi, dm, dc)) { atomicExch(bad, 1); return; }
    const double dH = __ddiv_rn(299792.458, c.H0);
    if (DL) DL[i] = __dmul_rn(__dadd_rn(1.0, zi), dm);
    if (dV) dV[i] = __ddiv_rn(__dmul_rn(__dmul_rn(dH, dm), dm), efunc_np(c, zi));
}

// numpy.interp (compiled_base.c arr_interp) for one value inside the knot range
__device__ __forceinline__ double interp_np(long long nk, const double* __restrict__ xk, const double* __restrict__ yk, double xv) {
    const long long j = knot_segment(nk, xk, xv);
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
    return r;
}

__global__ void k_interp(long long nk, const double* __restrict__ xk, const double* __restrict__ yk, long long n,
                         const double* __restrict__ x, double* __restrict__ y, int* __restrict__ bad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xv = x[i];
    if (!(xv >= xk[0]) || !(xv <= xk[nk - 1])) { atomicExch(bad, 1); y[i] = xv != xv ? xv : 0.0; return; }
    y[i] = interp_np(nk, xk, yk, xv);
}

// per-source tabulated Omega (reference lumfuncmcmc.py:235 -> Omega :47-70 -> V.fleming): reference order of operations,
// libdevice pow / log10 / exp / sqrt (<= 2 ulp from NumPy's)
struct OmArgs {
    long long n; const double* lum; const double* z; int K; long long field_ind[LF_MAX_FIELDS + 1];
    double om0[LF_MAX_FIELDS], F50[LF_MAX_FIELDS], ftau[LF_MAX_FIELDS]; double alpha; int modified;
    long long nk; const double* zk; const double* DLk; double* out; int* bad;
};
__global__ void k_omega_sources(OmArgs a) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int k = 0;
    while (k + 1 < a.K && i >= a.field_ind[k + 1]) ++k;
    const double zv = a.z[i];
    if (!(zv >= a.zk[0]) || !(zv <= a.zk[a.nk - 1])) { atomicExch(a.bad, 1); a.out[i] = 0.0; return; }
    const double DL = interp_np(a.nk, a.zk, a.DLk, zv);
    const double d = __dmul_rn(MPC_CM_REF, DL);
    const double flux = __ddiv_rn(pow(10.0, a.lum[i]), __dmul_rn(FOURPI, __dmul_rn(d, d)));      // lumfuncmcmc.py:69-70
    a.out[i] = __dmul_rn(a.om0[k], fleming_literal(flux, a.F50[k], a.alpha, a.ftau[k], a.modified != 0));
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


extern "C" int lf_omega_sources(int32_t device, int64_t n, const double* lum, const double* z, const int64_t* field_ind,
                                int32_t nfields, const int64_t* omega0_int, const double* flim, double alpha, double fcmin,
                                int64_t nk, const double* zk, const double* DLk, double* om_out) {
    if (n < 0 || (n > 0 && (!lum || !z || !om_out)) || !field_ind || !omega0_int || !flim || !zk || !DLk || nk < 2)
        return fail("lf_omega_sources: bad arguments");
    if (nfields < 1 || nfields > LF_MAX_FIELDS) return fail("lf_omega_sources: nfields out of range");
    if (field_ind[0] != 0 || field_ind[nfields] != n) return fail("lf_omega_sources: field_ind must run from 0 to n");
    if (n == 0) return 0;
    CK(cudaSetDevice(device));
    DevBufs bufs;
    double *d_lum = nullptr, *d_z = nullptr, *d_zk = nullptr, *d_DLk = nullptr, *d_out = nullptr;
    int* d_bad = nullptr;
    CK(bufs.alloc(&d_lum, sizeof(double) * n));
    CK(bufs.alloc(&d_z, sizeof(double) * n));
    CK(bufs.alloc(&d_out, sizeof(double) * n));
    CK(bufs.alloc(&d_zk, sizeof(double) * nk));
    CK(bufs.alloc(&d_DLk, sizeof(double) * nk));
    CK(bufs.alloc(&d_bad, sizeof(int)));
    CK(cudaMemcpy(d_lum, lum, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_z, z, sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_zk, zk, sizeof(double) * nk, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_DLk, DLk, sizeof(double) * nk, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bad, 0, sizeof(int)));
    OmArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.lum = d_lum; a.z = d_z; a.K = nfields; a.alpha = alpha; a.modified = fcmin != 0.0;
    const double aa = (2.0 * fcmin - 1.0) * (2.0 * fcmin - 1.0);
    for (int k = 0; k <= nfields; ++k) a.field_ind[k] = field_ind[k];
    for (int k = 0; k < nfields; ++k) {
        a.om0[k] = (double)omega0_int[k] / SQARCSEC;                                      // Omega_0 / V.sqarcsec, integer-typed areas (:285)
        a.F50[k] = 1.0e-17 * flim[k];
        a.ftau[k] = a.F50[k] * pow(10.0, -1.0 * pow(fabs(aa / (1.0 - aa)) * pow(alpha, -2.0), 0.5));   // VmaxLumFunc.py:164-167
    }
    a.nk = nk; a.zk = d_zk; a.DLk = d_DLk; a.out = d_out; a.bad = d_bad;
    k_omega_sources<<<(unsigned)((n + 255) / 256), 256>>>(a);
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return fail("lf_omega_sources: a redshift is outside the interpolation range (or NaN)");
    CK(cudaMemcpy(om_out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}
