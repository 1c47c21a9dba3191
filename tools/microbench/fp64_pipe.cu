// FP64-pipe micro-benchmarks behind DESIGN.md section 4.1 (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe tools/microbench/fp64_pipe.cu && ./fp64_pipe
// Findings recorded in profiles/README.md:
//   * DFMA/DMUL/DADD all sustain ~1.99 cycles per warp instruction per SMSP regardless of operand form
//     (register / constant bank / immediate): 1.86e13 instr/s per GPU.
//   * dependent DFMA latency 8.2 cycles; one warp per SMSP with 4-8 independent chains reaches 87-90 % of peak,
//     but FOUR warps per SMSP with 1 (2) chain(s) each only 66 % (79 %): ~3 cycles per instruction when consecutive
//     FP64 instructions come from different warps, 2 when they come from the same warp.
//   * MUFU.RSQ64H: 17 cycles dependent, 8 cycles per warp instruction per SMSP throughput.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dfma(int iters, const double* in, double* sink) {
    double a[CH];
    double m = in[1], c = in[2];
    for (int i = 0; i < CH; ++i) a[i] = in[threadIdx.x % 8] + i;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < CH; ++j) a[j] = fma(a[j], m, c);
    }
    double s = 0;
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}

template <int CH>
__global__ void k_rsq(int iters, const double* in, double* sink) {
    double a[CH];
    for (int i = 0; i < CH; ++i) a[i] = in[threadIdx.x % 8] + i;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                double r;
                asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[j]));
                a[j] = r;
            }
    }
    double s = 0;
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 12345.678) sink[0] = s;
}

template <typename K>
void run(const char* name, K k, int warps, int ch, const double* in, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 20000;
    k<<<148, 32 * warps>>>(100, in, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148, 32 * warps>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double per_warp = (double)iters * 16 * ch, cyc = ms * 1e-3 * 1.95e9;
    printf("%-24s warps/SM=%2d chains=%d : %6.2f cycles per instr per warp, %5.2f per instr per SMSP\n", name, warps, ch,
           cyc / per_warp, cyc / (per_warp * warps / 4.0));
}

int main() {
    double *sink, *in;
    cudaMalloc(&sink, 8);
    cudaMalloc(&in, 8 * 256);
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run("DFMA dependent", k_dfma<1>, 4, 1, in, sink);
    run("DFMA", k_dfma<2>, 4, 2, in, sink);
    run("DFMA", k_dfma<4>, 4, 4, in, sink);
    run("DFMA", k_dfma<8>, 4, 8, in, sink);
    run("DFMA", k_dfma<1>, 16, 1, in, sink);
    run("DFMA", k_dfma<2>, 16, 2, in, sink);
    run("DFMA", k_dfma<4>, 16, 4, in, sink);
    run("DFMA", k_dfma<8>, 32, 8, in, sink);
    run("MUFU.RSQ64H dependent", k_rsq<1>, 4, 1, in, sink);
    run("MUFU.RSQ64H", k_rsq<4>, 16, 4, in, sink);
    return 0;
}
