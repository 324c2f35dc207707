#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cth[4096];
struct P { int iters; double *sink; double th[3000]; };   // th at offset 16: 16-byte aligned
template <int MODE>
__global__ void __launch_bounds__(256) k(const __grid_constant__ P q) {
    double a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; b[i] = 1.0 - (i + threadIdx.x) * 1e-9; }
    for (int it = 0; it < q.iters; it++) {
        const int base = (it & 31) * 64;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 64; j++) a[j & 15] = fma(b[j & 15], cth[base + j], a[j & 15]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 64; j++) a[j & 15] = fma(b[j & 15], q.th[base + j], a[j & 15]);
        } else if (MODE == 2) {      // each coefficient feeds two fmas (value and derivative), like the gradient leaf
#pragma unroll
            for (int j = 0; j < 32; j++) { const double t = q.th[base + j]; a[j & 7] = fma(b[j & 7], t, a[j & 7]); a[8 + (j & 7)] = fma(b[8 + (j & 7)], t, a[8 + (j & 7)]); }
        } else {                     // same, coefficients fetched in pairs
            const double2 *t2 = reinterpret_cast<const double2 *>(q.th + base);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const double2 t = t2[j];
                a[j & 3] = fma(b[j & 3], t.x, a[j & 3]); a[4 + (j & 3)] = fma(b[4 + (j & 3)], t.x, a[4 + (j & 3)]);
                a[8 + (j & 3)] = fma(b[8 + (j & 3)], t.y, a[8 + (j & 3)]); a[12 + (j & 3)] = fma(b[12 + (j & 3)], t.y, a[12 + (j & 3)]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123.456) q.sink[0] = s;
}
template <int MODE> void run(const char *name) {
    static P p; cudaMalloc(&p.sink, 64); p.iters = 2048;
    for (int i = 0; i < 3000; i++) p.th[i] = 1.0 + i * 1e-9;
    cudaMemcpyToSymbol(cth, p.th, 3000 * 8);
    const int blocks = 148 * 8;
    p.iters = 16; k<MODE><<<blocks, 256>>>(p); p.iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(p); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %.2f TFLOP/s  (%s)\n", name, 2.0 * 64 * p.iters * blocks * 256 / (ms * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<0>("__constant__ window, 1 fma per coefficient"); run<1>("kernel parameter, 1 fma per coefficient"); run<2>("kernel parameter, 2 fma per coefficient"); run<3>("kernel parameter, pairs, 2 fma per coefficient"); return 0; }
