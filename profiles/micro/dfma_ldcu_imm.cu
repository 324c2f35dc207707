// LDCU.128 with IMMEDIATE constant addresses (fully unrolled straight-line code): cost of (LDCU.128 + 4 DFMA with a UR operand)
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cth[4096];
template <int U, int NT, int NP>
__global__ void __launch_bounds__(NT, 1) k(int iters, double *sink) {
    __shared__ __align__(16) double sth[4096];
    for (int i = threadIdx.x; i < 4096; i += NT) sth[i] = cth[i];
    __syncthreads();
    double t[8], d[8], a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = 1e-9 * (i + 1 + threadIdx.x); d[i] = 1e-9 * (i + 3 + threadIdx.x); a0[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; a1[i] = 0.5 + i * 1e-3; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (U == 0) {   // shared memory, immediate offsets
            const double2 *p = reinterpret_cast<const double2 *>(sth);
#pragma unroll
            for (int j = 0; j < NP; j++) { const double2 v = p[j];
                a0[j & 7] = fma(t[j & 7], v.x, a0[j & 7]); a1[j & 7] = fma(d[j & 7], v.x, a1[j & 7]);
                a0[(j + 4) & 7] = fma(t[(j + 4) & 7], v.y, a0[(j + 4) & 7]); a1[(j + 4) & 7] = fma(d[(j + 4) & 7], v.y, a1[(j + 4) & 7]); }
        } else {        // constant bank, immediate offsets
            const double2 *p = reinterpret_cast<const double2 *>(cth);
#pragma unroll
            for (int j = 0; j < NP; j++) { const double2 v = p[j];
                a0[j & 7] = fma(t[j & 7], v.x, a0[j & 7]); a1[j & 7] = fma(d[j & 7], v.x, a1[j & 7]);
                a0[(j + 4) & 7] = fma(t[(j + 4) & 7], v.y, a0[(j + 4) & 7]); a1[(j + 4) & 7] = fma(d[(j + 4) & 7], v.y, a1[(j + 4) & 7]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a0[i] + a1[i] + t[i] + d[i];
    if (s == 123.456) sink[0] = s;
}
template <int U, int NT, int NP>
void run(const char *name) {
    double *sink; cudaMalloc(&sink, 64);
    const int iters = (1 << 22) / NP, blocks = 148;
    k<U, NT, NP><<<blocks, NT>>>(64, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<U, NT, NP><<<blocks, NT>>>(iters, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double nf = 4.0 * NP;
    printf("%-40s NT=%d NP=%d %6.2f TFLOP/s  %.2f cycles/DFMA/scheduler\n", name, NT, NP, 2.0 * nf * iters * blocks * NT / (ms * 1e-3) * 1e-12,
           ms * 1e-3 * 1.965e9 / ((double)iters * nf * (NT / 128.0)));
}
int main() {
    double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + i * 1e-9;
    cudaMemcpyToSymbol(cth, h, sizeof h);
    run<0, 384, 256>("LDS.128 imm pairs, 2 fma per coeff"); run<1, 384, 256>("LDCU.128 imm pairs, 2 fma per coeff");
    run<0, 384, 1024>("LDS.128 imm pairs, 2 fma per coeff"); run<1, 384, 1024>("LDCU.128 imm pairs, 2 fma per coeff");
    run<0, 256, 256>("LDS.128 imm pairs, 2 fma per coeff"); run<1, 256, 256>("LDCU.128 imm pairs, 2 fma per coeff");
    return 0;
}
