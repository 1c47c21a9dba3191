// Issue-model micro-benchmark for the FP64 pipe of B200 (sm_100a): how many cycles does a warp-level DFMA cost per
// SM sub-partition (SMSP) when integer / MUFU / LDS instructions are interleaved, as a function of warps per SMSP and
// independent chains per warp?  Decides the warps x ILP shape of k_main and the ceiling quoted in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix tools/microbench/fp64_mix.cu && ./fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

// CH independent DFMA chains, NI integer (IMAD) instructions and NL shared-memory loads per CH DFMAs
template <int CH, int NI, int NL>
__global__ void k_mix(int iters, const double* in, double* sink) {
    __shared__ double tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = in[i & 255];
    __syncthreads();
    double a[CH];
    unsigned x[NI > 0 ? NI : 1];
    double m = in[1], c = in[2];
    for (int i = 0; i < CH; ++i) a[i] = in[threadIdx.x % 8] + i;
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) x[i] = threadIdx.x * 7 + i;
    double ls = 0.0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                a[j] = fma(a[j], m, c);
                if (j < NI) x[j] = x[j] * 1664525u + 1013904223u;
                if (j < NL) ls += tab[(x[j % (NI > 0 ? NI : 1)] >> 8) & 1023];
            }
        }
    }
    double s = ls;
    for (int i = 0; i < CH; ++i) s += a[i];
    unsigned xs = 0;
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) xs ^= x[i];
    if (s == 12345.678 || xs == 0x12345u) sink[0] = s;
}

template <int CH, int NI, int NL>
void run(int warps, const double* in, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 8000;
    k_mix<CH, NI, NL><<<148, 32 * warps>>>(100, in, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mix<CH, NI, NL><<<148, 32 * warps>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double dfma_per_warp = (double)iters * 8 * CH, cyc = ms * 1e-3 * 1.965e9;
    printf("chains=%d int/dfma=%.2f lds/dfma=%.2f warps/SM=%2d : %5.2f cycles per DFMA per SMSP  (%.0f%% of 2.0)\n", CH,
           (double)NI / CH, (double)NL / CH, warps, cyc / (dfma_per_warp * warps / 4.0), 200.0 / (cyc / (dfma_per_warp * warps / 4.0)));
}

int main() {
    double *sink, *in;
    cudaMalloc(&sink, 8);
    cudaMalloc(&in, 8 * 256);
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int ws[5] = {4, 8, 16, 32, 64};
    for (int wi = 0; wi < 4; ++wi) {
        int w = ws[wi];
        printf("---- %d warps per SM (%d per SMSP)\n", w, w / 4);
        run<1, 0, 0>(w, in, sink);
        run<2, 0, 0>(w, in, sink);
        run<4, 0, 0>(w, in, sink);
        run<8, 0, 0>(w, in, sink);
        run<4, 2, 0>(w, in, sink);
        run<4, 4, 0>(w, in, sink);
        run<8, 2, 0>(w, in, sink);
        run<8, 4, 0>(w, in, sink);
        run<8, 8, 0>(w, in, sink);
        run<8, 4, 1>(w, in, sink);
        run<8, 4, 2>(w, in, sink);
        run<4, 2, 1>(w, in, sink);
    }
    return 0;
}
