// Does the operand form of a DFMA change its issue cost on B200?  8 independent chains, 4 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands tools/microbench/fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double KC[4] = {0.999999, 1.0e-9, 0.5, 0.25};

// MODE 0: a = fma(a, m, c)        m, c shared registers
// MODE 1: a_j = fma(a_j, b_j, c_j) three distinct registers per chain
// MODE 2: a_j = fma(a_j, b_j, 0.5) two registers + immediate
// MODE 3: a_j = fma(a_j, b_j, KC[1]) two registers + constant bank
// MODE 4: a_j = a_j * b_j          DMUL, two registers
// MODE 5: a_j = fma(a_j, KC[0], KC[1])  one register (the compiler may materialise one constant)
// MODE 6: a_j = fma(b_j, c_j, a_j) three distinct, accumulator last
// MODE 7: a_j = fma(a_j, a_j, b_j)
template <int MODE, int WITH_INT>
__global__ void k_ops(int iters, const double* in, double* sink) {
    constexpr int CH = 8;
    double a[CH], b[CH], c[CH];
    unsigned x[CH];
    double m = in[1], cc = in[2];
    for (int i = 0; i < CH; ++i) {
        a[i] = in[threadIdx.x % 8] + i;
        b[i] = in[8 + i];
        c[i] = in[16 + i] - 0.999;
        x[i] = threadIdx.x + i;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (MODE == 0) a[j] = fma(a[j], m, cc);
                if (MODE == 1) a[j] = fma(a[j], b[j], c[j]);
                if (MODE == 2) a[j] = fma(a[j], b[j], 0.5);
                if (MODE == 3) a[j] = fma(a[j], b[j], KC[1]);
                if (MODE == 4) a[j] = a[j] * b[j];
                if (MODE == 5) a[j] = fma(a[j], KC[0], KC[1]);
                if (MODE == 6) a[j] = fma(b[j], c[j], a[j]);
                if (MODE == 7) a[j] = fma(a[j], a[j], b[j]);
                if (WITH_INT) x[j] = x[j] * 1664525u + 1013904223u;
            }
        }
    }
    double s = 0;
    unsigned xs = 0;
    for (int i = 0; i < CH; ++i) { s += a[i]; xs ^= x[i]; }
    if (s == 12345.678 || xs == 0x12345u) sink[0] = s;
}

template <int MODE, int WITH_INT>
void run(const char* name, const double* in, double* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 8000, warps = 16;
    k_ops<MODE, WITH_INT><<<148, 32 * warps>>>(100, in, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_ops<MODE, WITH_INT><<<148, 32 * warps>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double per_warp = (double)iters * 8 * 8, cyc = ms * 1e-3 * 1.965e9;
    printf("%-44s int=%d : %5.2f cycles per FP64 instr per SMSP\n", name, WITH_INT, cyc / (per_warp * warps / 4.0));
}

int main() {
    double *sink, *in;
    cudaMalloc(&sink, 8);
    cudaMalloc(&in, 8 * 256);
    double h[256];
    for (int i = 0; i < 256; ++i) h[i] = 0.999 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0, 0>("fma(a, m, c)  shared m, c", in, sink);
    run<0, 1>("fma(a, m, c)  shared m, c", in, sink);
    run<1, 0>("fma(a_j, b_j, c_j)  three distinct", in, sink);
    run<1, 1>("fma(a_j, b_j, c_j)  three distinct", in, sink);
    run<6, 0>("fma(b_j, c_j, a_j)  three distinct, acc last", in, sink);
    run<6, 1>("fma(b_j, c_j, a_j)  three distinct, acc last", in, sink);
    run<2, 0>("fma(a_j, b_j, imm)", in, sink);
    run<2, 1>("fma(a_j, b_j, imm)", in, sink);
    run<3, 0>("fma(a_j, b_j, c[3][..])", in, sink);
    run<3, 1>("fma(a_j, b_j, c[3][..])", in, sink);
    run<4, 0>("a_j * b_j", in, sink);
    run<4, 1>("a_j * b_j", in, sink);
    run<5, 0>("fma(a_j, const, const)", in, sink);
    run<5, 1>("fma(a_j, const, const)", in, sink);
    run<7, 0>("fma(a_j, a_j, b_j)", in, sink);
    run<7, 1>("fma(a_j, a_j, b_j)", in, sink);
    return 0;
}
