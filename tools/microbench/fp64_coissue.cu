// Which instruction types issue for free in the shadow of a DFMA on B200?  8 DFMA chains per warp, 4 warps per SMSP,
// R other-instructions per DFMA of one type (inline PTX so the compiler cannot fold them).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_coissue tools/microbench/fp64_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { T_NONE, T_IMAD, T_LOP, T_SHF, T_IADD, T_FFMA, T_FADD, T_MOV, T_IMNMX, T_SEL };

template <int TYPE, int R>
__global__ void k(int iters, const double* in, double* sink) {
    constexpr int CH = 8;
    double a[CH], b[CH];
    unsigned x[CH], y[CH];
    float f[CH];
    for (int i = 0; i < CH; ++i) {
        a[i] = in[threadIdx.x % 8] + i; b[i] = in[8 + i];
        x[i] = threadIdx.x * 3 + i; y[i] = threadIdx.x + 7 * i; f[i] = 0.5f + i;
    }
    unsigned m = threadIdx.x | 1;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                a[j] = fma(a[j], b[j], 0.5);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (TYPE == T_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(m), "r"(y[j]));
                    if (TYPE == T_LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(y[j]));
                    if (TYPE == T_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(x[j]) : "r"(y[j]));
                    if (TYPE == T_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y[j]));
                    if (TYPE == T_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[j]) : "f"(f[(j + 1) % CH]));
                    if (TYPE == T_FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(f[(j + 1) % CH]));
                    if (TYPE == T_IMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y[j]));
                    if (TYPE == T_SEL) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.b32 %0, %1, %0, p;}" : "+r"(x[j]) : "r"(y[j]));
                }
            }
        }
    }
    double s = 0;
    unsigned xs = 0;
    float fs = 0;
    for (int i = 0; i < CH; ++i) { s += a[i]; xs ^= x[i]; fs += f[i]; }
    if (s == 12345.678 || xs == 0x12345u || fs == 1.2345f) sink[0] = s;
}

template <int TYPE, int R>
void run(const char* name, const double* in, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 8000, warps = 16;
    k<TYPE, R><<<148, 32 * warps>>>(100, in, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<TYPE, R><<<148, 32 * warps>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double per_warp = (double)iters * 64, cyc = ms * 1e-3 * 1.965e9;
    printf("%-8s x%d per DFMA : %5.2f cycles per DFMA per SMSP\n", name, R, cyc / (per_warp * warps / 4.0));
}

int main() {
    double *sink, *in;
    cudaMalloc(&sink, 8);
    cudaMalloc(&in, 8 * 256);
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<T_NONE, 0>("none", in, sink);
    run<T_IMAD, 1>("IMAD", in, sink);
    run<T_IMAD, 2>("IMAD", in, sink);
    run<T_LOP, 1>("LOP3", in, sink);
    run<T_LOP, 2>("LOP3", in, sink);
    run<T_SHF, 1>("SHF", in, sink);
    run<T_SHF, 2>("SHF", in, sink);
    run<T_IADD, 1>("IADD", in, sink);
    run<T_IADD, 2>("IADD", in, sink);
    run<T_FFMA, 1>("FFMA", in, sink);
    run<T_FFMA, 2>("FFMA", in, sink);
    run<T_FADD, 1>("FADD", in, sink);
    run<T_IMNMX, 1>("IMNMX", in, sink);
    run<T_IMNMX, 2>("IMNMX", in, sink);
    run<T_SEL, 1>("SETP+SEL", in, sink);
    return 0;
}
