// Microbenchmark (round 2): what does a DFMA cost as a function of how many of its three 64-bit register operands are
// "fresh" (not held in the operand-reuse cache from the previous instruction of the same warp)?
// Shape of the real leaf kernel: per-thread Chebyshev tables t[], d[] in registers (loop invariant), coefficients theta
// from shared memory by broadcast LDS.128 (fresh registers every time), accumulators read-modify-write.
//   P0  a[i] = fma(a[i], m, c)                                  peak (two invariant operands)
//   P1  a[i] = fma(t[i], th[i], a[i])                           three fresh operands, no reuse possible
//   P2  a0[i] = fma(t[i], th[i], a0[i]); a1[i] = fma(d[i], th[i], a1[i])     pairs sharing theta (slot b)
//   P3  a0[i] = fma(t0, th[i], a0[i]) for all i, then a1[i] = fma(d0, th[i], a1[i])   runs of 8 sharing slot a
//   P4  level fold t=3: fma(dT, c0, a3), fma(T, c0, a0), fma(T, c1, a1), fma(T, c2, a2) with c from LDS
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o dfma_reuse dfma_reuse.cu ; ./dfma_reuse [threads]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int P, int NT>
__global__ void __launch_bounds__(NT, 1) k(int iters, double *sink) {
    __shared__ __align__(16) double sth[2048];
    for (int i = threadIdx.x; i < 2048; i += NT) sth[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double t[16], d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { t[i] = 1.0 - (i + threadIdx.x) * 1e-9; d[i] = 1e-9 * (i + 1 + threadIdx.x); }
    double a0[16], a1[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { a0[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; a1[i] = 0.5 + i * 1e-3; }
    const double m = 1.0000001, cc = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; it++) {
        const double2 *th2 = reinterpret_cast<const double2 *>(sth + ((it * 16) & 1023));
        if (P == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a0[i] = fma(a0[i], m, cc);
        } else {
            double th[16];
#pragma unroll
            for (int i = 0; i < 8; i++) { const double2 v = th2[i]; th[2 * i] = v.x; th[2 * i + 1] = v.y; }
            if (P == 1) {
#pragma unroll
                for (int i = 0; i < 16; i++) a0[i] = fma(t[i], th[i], a0[i]);
            } else if (P == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) { a0[i] = fma(t[i], th[i], a0[i]); a1[i] = fma(d[i], th[i], a1[i]); }
            } else if (P == 3) {
#pragma unroll
                for (int i = 0; i < 8; i++) a0[i] = fma(t[it & 1 ? 0 : 0], th[i], a0[i]);
#pragma unroll
                for (int i = 0; i < 8; i++) a1[i] = fma(d[0], th[i], a1[i]);
            } else if (P == 4) {
#pragma unroll
                for (int g = 0; g < 4; g++) {     // four folds of four fmas, different (T, dT) and accumulators
                    a1[g] = fma(d[g], th[4 * g], a1[g]);
                    a0[4 * g] = fma(t[g], th[4 * g], a0[4 * g]);
                    a0[4 * g + 1] = fma(t[g], th[4 * g + 1], a0[4 * g + 1]);
                    a0[4 * g + 2] = fma(t[g], th[4 * g + 2], a0[4 * g + 2]);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a0[i] + a1[i] + t[i] + d[i];
    if (s == 123.456) sink[0] = s;
}

template <int P, int NT>
void run(const char *name) {
    double *sink; cudaMalloc(&sink, 64);
    const int iters = 1 << 16, blocks = 148;
    k<P, NT><<<blocks, NT>>>(64, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<P, NT><<<blocks, NT>>>(iters, sink);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double nf = 16;
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * nf * (NT / 128.0));   // cycles per warp-DFMA per scheduler at 1965 MHz
    printf("%-70s NT=%d  %.2f TFLOP/s  %.2f cycles/DFMA/scheduler\n", name, NT, 2.0 * nf * iters * blocks * NT / (ms * 1e-3) * 1e-12, cyc);
}
template <int NT> void all() {
    run<0, NT>("P0 fma(a, m, c): peak");
    run<1, NT>("P1 fma(t[i], th[i], a[i]): three fresh");
    run<2, NT>("P2 pairs sharing theta: (3 fresh, 2 fresh)");
    run<3, NT>("P3 runs of 8 sharing T in slot a: 2 fresh");
    run<4, NT>("P4 level fold t=3: (3, 2, 2, 2)");
}
int main() { all<384>(); all<256>(); all<128>(); return 0; }
