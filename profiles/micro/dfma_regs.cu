// Microbenchmark (round 2): DFMA cost vs number of distinct register operands, registers only (no loads, no filler).
// All operands but the accumulators are loop invariant; accumulators are read-modify-write (16 independent chains).
//   Q0 a[i] = fma(a[i], m, c)            accumulator in slot a, two invariant operands
//   Q1 a[i] = fma(t[i], d[i], a[i])      three distinct registers per instruction, nothing shared between neighbours
//   Q2 a[i] = fma(t0, d[i], a[i])        slot a shared by all
//   Q3 a[i] = fma(t0, d0, a[i])          slots a and b shared
//   Q4 a[i] = fma(t[i], d[i], a[i]); b[i] = fma(t[i], d[i], b[i])    pairs sharing two operands
//   Q5 a[i] = fma(t[i], d[i], a[i]); b[i] = fma(t[i], e[i], b[i])    pairs sharing slot a
//   Q6 a[i] = fma(t[i], d0, a[i])        slot b shared by all
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o dfma_regs dfma_regs.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int Q, int NT>
__global__ void __launch_bounds__(NT, 1) k(int iters, double *sink, const double *src) {
    double t[16], d[16], e[16], a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        t[i] = src[i] * 1e-9 * (1 + threadIdx.x); d[i] = src[16 + i] + threadIdx.x * 1e-6; e[i] = src[32 + i] - threadIdx.x * 1e-6;
        a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; b[i] = 0.5 + i * 1e-3;
    }
    const double m = 1.0000001, cc = 1e-9 * (threadIdx.x + 1);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (Q == 0) a[i] = fma(a[i], m, cc);
            if (Q == 1) a[i] = fma(t[i], d[i], a[i]);
            if (Q == 2) a[i] = fma(t[0], d[i], a[i]);
            if (Q == 3) a[i] = fma(t[0], d[0], a[i]);
            if (Q == 4 && i < 8) { a[i] = fma(t[i], d[i], a[i]); b[i] = fma(t[i], d[i], b[i]); }
            if (Q == 5 && i < 8) { a[i] = fma(t[i], d[i], a[i]); b[i] = fma(t[i], e[i], b[i]); }
            if (Q == 6) a[i] = fma(t[i], d[0], a[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i] + b[i] + t[i] + d[i] + e[i];
    if (s == 123.456) sink[0] = s;
}

template <int Q, int NT>
void run(const char *name, const double *src) {
    double *sink; cudaMalloc(&sink, 64);
    const int iters = 1 << 16, blocks = 148;
    k<Q, NT><<<blocks, NT>>>(64, sink, src);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<Q, NT><<<blocks, NT>>>(iters, sink, src);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double nf = 16;
    printf("%-62s NT=%d  %6.2f TFLOP/s  %.2f cycles/DFMA/scheduler\n", name, NT, 2.0 * nf * iters * blocks * NT / (ms * 1e-3) * 1e-12,
           ms * 1e-3 * 1.965e9 / ((double)iters * nf * (NT / 128.0)));
}
template <int NT> void all(const double *src) {
    run<0, NT>("Q0 fma(a, m, c)", src);
    run<1, NT>("Q1 fma(t[i], d[i], a[i]) three distinct", src);
    run<2, NT>("Q2 fma(t0, d[i], a[i]) slot a shared", src);
    run<3, NT>("Q3 fma(t0, d0, a[i]) slots a, b shared", src);
    run<4, NT>("Q4 pairs sharing two operands", src);
    run<5, NT>("Q5 pairs sharing slot a", src);
    run<6, NT>("Q6 fma(t[i], d0, a[i]) slot b shared", src);
}
int main() {
    double h[48]; for (int i = 0; i < 48; i++) h[i] = 1.0 + i * 1e-3;
    double *src; cudaMalloc(&src, sizeof h); cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
    all<384>(src); all<128>(src); return 0;
}
