// Microbenchmark: does the DFMA rate depend on where the operands come from?  16 independent chains per thread,
// 148*8 CTAs of 256 threads (same shape as sk_fp64_peak):
//   A  a[i] = fma(a[i], m, c)        two loop-invariant register operands (operand-reuse cache)
//   B  a[i] = fma(a[i], b[i], c[i])  three distinct register operands per instruction
//   C  a[i] = fma(a[i], b[i], c)     two distinct + one invariant
//   D  a[i] = fma(b[i], m, a[i])     accumulate form with a reused multiplier (what the nested kernels do)
//   E  a[i] = fma(b[i], cth[u + i], a[i])  multiplier from __constant__ memory at a warp-uniform runtime index
//   F  same with the multiplier read from shared memory (broadcast LDS) — what the leaf kernel did
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu && ./dfma_operands
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double cth[4096];

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, double *sink) {
    __shared__ double sth[4096];
    if (MODE == 5) { for (int i = threadIdx.x; i < 4096; i += 256) sth[i] = 1.0 + i * 1e-9; __syncthreads(); }
    double a[16], b[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; b[i] = 1.0 - i * 1e-9; c[i] = 1e-9 * (i + 1); }
    const double m = 1.0000001, cc = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) a[i] = fma(a[i], m, cc);
            if (MODE == 1) a[i] = fma(a[i], b[i], c[i]);
            if (MODE == 2) a[i] = fma(a[i], b[i], cc);
            if (MODE == 3) a[i] = fma(b[i], m, a[i]);
            if (MODE == 4) a[i] = fma(b[i], cth[((it * 16) & 2047) + i], a[i]);
            if (MODE == 5) a[i] = fma(b[i], sth[((it * 16) & 2047) + i], a[i]);
        }
        if (MODE >= 1) {      // keep b, c live and changing without FP64 work
#pragma unroll
            for (int i = 0; i < 16; i++) { b[i] = __longlong_as_double(__double_as_longlong(b[i]) ^ (it & 1)); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i] + b[i] + c[i];
    if (s == 123.456) sink[0] = s;
}

template <int MODE>
void run(const char *name) {
    double *sink; cudaMalloc(&sink, 64);
    const int iters = 8192, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(64, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(iters, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-45s %.2f TFLOP/s\n", name, 2.0 * 16 * iters * blocks * 256 / (ms * 1e-3) * 1e-12);
}
int main() {
    run<0>("A fma(a, m, c): invariant operands");
    run<1>("B fma(a, b[i], c[i]): three distinct");
    run<2>("C fma(a, b[i], c): two distinct");
    run<3>("D fma(b[i], m, a): accumulate, reused m");
    { double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + i * 1e-9; cudaMemcpyToSymbol(cth, h, sizeof h); }
    run<4>("E fma(b[i], const[u+i], a): constant bank");
    run<5>("F fma(b[i], smem[u+i], a): shared memory");
    return 0;
}
