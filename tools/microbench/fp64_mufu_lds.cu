// Do MUFU.RSQ64H / MUFU.RCP64H and shared-memory table lookups steal FP64 issue slots on B200?
// 8 independent DFMA chains per warp, 4 warps per SMSP; per 16 DFMAs: NM 64-bit MUFUs and NL conflict-free LDS.64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mufu_lds tools/microbench/fp64_mufu_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NM, int NL, int F32MUFU>
__global__ void k_mix(int iters, const double* in, double* sink) {
    __shared__ double tab[128 * 32];
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) tab[i] = in[i & 255];
    __syncthreads();
    constexpr int CH = 8;
    double a[CH], b[CH];
    double msum = 0.0;
    float fsum = 0.f;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < CH; ++i) { a[i] = in[threadIdx.x % 8] + i; b[i] = in[8 + i]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int j = 0; j < CH; ++j) a[j] = fma(a[j], b[j], 0.5);
            if (NM > 0) {
#pragma unroll
                for (int k = 0; k < NM; ++k) {
                    if (F32MUFU) {
                        float r;
                        asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)__double2hiint(a[(u * NM + k) % CH])));
                        fsum += r;
                    } else {
                        double r;
                        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[(u * NM + k) % CH]));
                        msum += r;                                  // one extra DADD per MUFU (counted below)
                    }
                }
            }
            if (NL > 0) {
#pragma unroll
                for (int k = 0; k < NL; ++k) {
                    int idx = (__double2loint(a[(u * NL + k) % CH]) >> 3) & 127;     // data-dependent row, lane-private column
                    msum += tab[idx * 32 + lane];
                }
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) a[j] = fma(a[j], b[j], 0.25);
        }
    }
    double s = msum + fsum;
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}

template <int NM, int NL, int F32MUFU>
void run(const double* in, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 16000, warps = 16;
    k_mix<NM, NL, F32MUFU><<<148, 32 * warps>>>(100, in, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mix<NM, NL, F32MUFU><<<148, 32 * warps>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // FP64-pipe instructions per warp per iteration: 2 x (16 DFMA + NM DADD (64-bit MUFU only) + NL DADD)
    double fp64 = (double)iters * 2 * (16 + (F32MUFU ? 0 : NM) + NL), cyc = ms * 1e-3 * 1.965e9;
    printf("per 16 DFMA: %d %s MUFU, %d LDS.64 : %5.2f cycles per FP64 instr per SMSP  (%.1f cycles per 16-DFMA group)\n", NM,
           F32MUFU ? "f32" : "f64", NL, cyc / (fp64 * warps / 4.0), cyc / ((double)iters * 2 * warps / 4.0));
}

int main() {
    double *sink, *in;
    cudaMalloc(&sink, 8);
    cudaMalloc(&in, 8 * 256);
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0, 0, 0>(in, sink);
    run<1, 0, 0>(in, sink);
    run<2, 0, 0>(in, sink);
    run<4, 0, 0>(in, sink);
    run<1, 0, 1>(in, sink);
    run<2, 0, 1>(in, sink);
    run<4, 0, 1>(in, sink);
    run<0, 1, 0>(in, sink);
    run<0, 2, 0>(in, sink);
    run<0, 4, 0>(in, sink);
    run<1, 1, 0>(in, sink);
    run<2, 2, 0>(in, sink);
    return 0;
}
