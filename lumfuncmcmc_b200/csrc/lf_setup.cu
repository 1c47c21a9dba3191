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

// numpy.interp, compiled_base.c arr_interp: j = last knot <= x (candidate from the mean spacing, exact comparisons decide)
__global__ void k_interp(long long nk, const double* __restrict__ xk, const double* __restrict__ yk, long long n,
                         const double* __restrict__ x, double* __restrict__ y, int* __restrict__ bad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xv = x[i];
    const double x0 = xk[0], x1 = xk[nk - 1];
    if (!(xv >= x0) || !(xv <= x1)) { atomicExch(bad, 1); y[i] = xv != xv ? xv : 0.0; return; }
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

