// Accuracy of the 64-bit MUFU seeds on B200 (feeds the error budget of lf_math.cuh):
// max |1 - y r0^2| for r0 = rsqrt.approx.ftz.f64(y), y in [1, 1e3]; max |1 - d r0| for r0 = rcp.approx.ftz.f64(d), d in [1e-4, 1]
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int n, double* out) {
    double m1 = 0, m2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double y = exp(log(1.0e3) * (i + 0.37) / n);
        double r;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
        m1 = fmax(m1, fabs(fma(-y * r, r, 1.0)));
        double d = exp(log(1.0e-4) * (i + 0.61) / n);
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        m2 = fmax(m2, fabs(fma(-d, r, 1.0)));
    }
    // crude reduction
    atomicMax((unsigned long long*)&out[0], __double_as_longlong(m1));
    atomicMax((unsigned long long*)&out[1], __double_as_longlong(m2));
}
int main() {
    double* d;
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    k<<<592, 256>>>(200000000, d);
    double h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("rsqrt seed: max |1 - y r0^2| = %.3e (2^%.2f)   -> 3e^2/8 = %.2e, 5e^3/16 = %.2e\n", h[0], log2(h[0]), 0.375 * h[0] * h[0], 0.3125 * h[0] * h[0] * h[0]);
    printf("rcp seed:   max |1 - d r0|   = %.3e (2^%.2f)   -> e^2 = %.2e\n", h[1], log2(h[1]), h[1] * h[1]);
    return 0;
}
