// Microbenchmark (round 2): coefficients from the constant bank through the uniform datapath (LDCU -> UR operand of DFMA):
// no register-file write for the load, no register-file read for the coefficient.
//   U0  LDS.128 broadcast (shared memory), two fmas per coefficient      -- what the leaf kernel does today
//   U1  LDCU.64, two fmas per coefficient (value, derivative)
//   U2  LDCU.128 (pairs), two fmas per coefficient
//   U3  LDCU.64, one fma per coefficient
//   U4  LDCU.128, one fma per coefficient
// Tables t[], d[] are per-thread registers (loop invariant), accumulators read-modify-write: 2 register reads per DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cth[4096];

template <int U, int NT>
__global__ void __launch_bounds__(NT, 1) k(int iters, double *sink) {
    __shared__ __align__(16) double sth[4096];
    for (int i = threadIdx.x; i < 4096; i += NT) sth[i] = cth[i];
    __syncthreads();
    double t[8], d[8], a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = 1e-9 * (i + 1 + threadIdx.x); d[i] = 1e-9 * (i + 3 + threadIdx.x); a0[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; a1[i] = 0.5 + i * 1e-3; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const int base = ((it * 16) & 2047);
        if (U == 0) {
            const double2 *p = reinterpret_cast<const double2 *>(sth + base);
#pragma unroll
            for (int j = 0; j < 8; j++) { const double2 v = p[j];
                a0[j] = fma(t[j], v.x, a0[j]); a1[j] = fma(d[j], v.x, a1[j]);
                a0[(j + 4) & 7] = fma(t[(j + 4) & 7], v.y, a0[(j + 4) & 7]); a1[(j + 4) & 7] = fma(d[(j + 4) & 7], v.y, a1[(j + 4) & 7]); }
        } else if (U == 1) {
#pragma unroll
            for (int j = 0; j < 16; j++) { const double v = cth[base + j]; a0[j & 7] = fma(t[j & 7], v, a0[j & 7]); a1[j & 7] = fma(d[j & 7], v, a1[j & 7]); }
        } else if (U == 2) {
            const double2 *p = reinterpret_cast<const double2 *>(cth + base);
#pragma unroll
            for (int j = 0; j < 8; j++) { const double2 v = p[j];
                a0[j] = fma(t[j], v.x, a0[j]); a1[j] = fma(d[j], v.x, a1[j]);
                a0[(j + 4) & 7] = fma(t[(j + 4) & 7], v.y, a0[(j + 4) & 7]); a1[(j + 4) & 7] = fma(d[(j + 4) & 7], v.y, a1[(j + 4) & 7]); }
        } else if (U == 3) {
#pragma unroll
            for (int j = 0; j < 32; j++) { const double v = cth[base + j]; if (j & 8) a1[j & 7] = fma(d[j & 7], v, a1[j & 7]); else a0[j & 7] = fma(t[j & 7], v, a0[j & 7]); }
        } else {
            const double2 *p = reinterpret_cast<const double2 *>(cth + base);
#pragma unroll
            for (int j = 0; j < 16; j++) { const double2 v = p[j];
                if (j & 4) { a1[j & 7] = fma(d[j & 7], v.x, a1[j & 7]); a1[(j + 4) & 7] = fma(d[(j + 4) & 7], v.y, a1[(j + 4) & 7]); }
                else { a0[j & 7] = fma(t[j & 7], v.x, a0[j & 7]); a0[(j + 4) & 7] = fma(t[(j + 4) & 7], v.y, a0[(j + 4) & 7]); } }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a0[i] + a1[i] + t[i] + d[i];
    if (s == 123.456) sink[0] = s;
}
template <int U, int NT>
void run(const char *name) {
    double *sink; cudaMalloc(&sink, 64);
    const int iters = 1 << 16, blocks = 148;
    k<U, NT><<<blocks, NT>>>(64, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<U, NT><<<blocks, NT>>>(iters, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double nf = 32;
    printf("%-50s NT=%d  %6.2f TFLOP/s  %.2f cycles/DFMA/scheduler\n", name, NT, 2.0 * nf * iters * blocks * NT / (ms * 1e-3) * 1e-12,
           ms * 1e-3 * 1.965e9 / ((double)iters * nf * (NT / 128.0)));
}
int main() {
    double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + i * 1e-9;
    cudaMemcpyToSymbol(cth, h, sizeof h);
    run<0, 384>("U0 LDS.128 pairs, 2 fma per coefficient"); run<1, 384>("U1 LDCU.64, 2 fma per coefficient"); run<2, 384>("U2 LDCU.128, 2 fma per coefficient");
    run<3, 384>("U3 LDCU.64, 1 fma per coefficient"); run<4, 384>("U4 LDCU.128, 1 fma per coefficient");
    run<0, 256>("U0 LDS.128 pairs, 2 fma per coefficient"); run<1, 256>("U1 LDCU.64, 2 fma per coefficient"); run<2, 256>("U2 LDCU.128, 2 fma per coefficient");
    run<3, 256>("U3 LDCU.64, 1 fma per coefficient"); run<4, 256>("U4 LDCU.128, 1 fma per coefficient");
    return 0;
}
