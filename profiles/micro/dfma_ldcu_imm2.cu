// LDCU.128 with immediate constant addresses: separate the code-size effect (instruction cache) from the constant-footprint effect
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cth[8192];
template <int NT, int NP, int MASK>
__global__ void __launch_bounds__(NT, 1) k(int iters, double *sink) {
    double t[8], d[8], a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = 1e-9 * (i + 1 + threadIdx.x); d[i] = 1e-9 * (i + 3 + threadIdx.x); a0[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; a1[i] = 0.5 + i * 1e-3; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const double2 *p = reinterpret_cast<const double2 *>(cth);
#pragma unroll
        for (int j = 0; j < NP; j++) { const double2 v = p[j & MASK];
            a0[j & 7] = fma(t[j & 7], v.x, a0[j & 7]); a1[j & 7] = fma(d[j & 7], v.x, a1[j & 7]);
            a0[(j + 4) & 7] = fma(t[(j + 4) & 7], v.y, a0[(j + 4) & 7]); a1[(j + 4) & 7] = fma(d[(j + 4) & 7], v.y, a1[(j + 4) & 7]); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a0[i] + a1[i] + t[i] + d[i];
    if (s == 123.456) sink[0] = s;
}
template <int NT, int NP, int MASK>
void run() {
    double *sink; cudaMalloc(&sink, 64);
    const int iters = (1 << 21) / NP, blocks = 148;
    k<NT, NP, MASK><<<blocks, NT>>>(8, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NT, NP, MASK><<<blocks, NT>>>(iters, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double nf = 4.0 * NP;
    printf("NT=%d code=%4d KB const=%5.1f KB  %6.2f TFLOP/s  %.2f cycles/DFMA/scheduler\n", NT, NP * 5 * 16 / 1024, ((MASK < NP ? MASK + 1 : NP) * 16) / 1024.0,
           2.0 * nf * iters * blocks * NT / (ms * 1e-3) * 1e-12, ms * 1e-3 * 1.965e9 / ((double)iters * nf * (NT / 128.0)));
}
int main() {
    double h[8192]; for (int i = 0; i < 8192; i++) h[i] = 1.0 + i * 1e-9;
    cudaMemcpyToSymbol(cth, h, sizeof h);
    printf("code-size sweep, 2 KB of constants\n");
    run<384, 128, 127>(); run<384, 256, 127>(); run<384, 512, 127>(); run<384, 1024, 127>(); run<384, 1536, 127>(); run<384, 2048, 127>();
    printf("constant-footprint sweep, 80 KB of code\n");
    run<384, 1024, 63>(); run<384, 1024, 255>(); run<384, 1024, 511>(); run<384, 1024, 1023>();
    printf("both, unique addresses\n");
    run<384, 128, 4095>(); run<384, 256, 4095>(); run<384, 512, 4095>(); run<384, 2048, 4095>();
    printf("256 threads\n");
    run<256, 256, 127>(); run<256, 1024, 127>(); run<256, 1024, 1023>();
    return 0;
}
