// Microbenchmark: latency of dependent FP64 CUDA-core ops (DFMA) issued by "producer" warps while
// "consumer" warps keep the FP64 pipe saturated with DMMA.m8n8k4.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu && ./fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH>
__global__ void __launch_bounds__(1024, 1) mix(int cons_warps, int prod_warps, int dmma_iters, int chain_len, long long *cyc, double *sink) {
    const int warp = threadIdx.x >> 5;
    if (warp < cons_warps) {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
        double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
        for (int it = 0; it < dmma_iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
        if (s == 123.456) sink[0] = s;
    } else if (warp < cons_warps + prod_warps) {
        double v[NCH];
#pragma unroll
        for (int i = 0; i < NCH; i++) v[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6;
        const double m = 1.0000001, k = 1e-9;
        long long t0 = clock64();
        for (int it = 0; it < chain_len; it++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) v[i] = fma(v[i], m, k);
        }
        long long t1 = clock64();
        double s = 0;
#pragma unroll
        for (int i = 0; i < NCH; i++) s += v[i];
        if (s == 123.456) sink[0] = s;
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp - cons_warps] = t1 - t0;
    }
}

int main() {
    long long *cyc; double *sink;
    cudaMalloc(&cyc, 64 * 8); cudaMalloc(&sink, 64);
    long long h[64];
    const int chain_len = 4096;
    printf("cons_warps prod_warps NCH  cycles_per_dependent_step (warp0)  [DMMA time ms]\n");
    for (int cons : {0, 4, 8, 16}) for (int prod : {1, 4, 8}) for (int nch : {1, 4, 16}) {
        int dm = cons ? 200000 : 1;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int threads = (cons + prod) * 32;
        if (nch == 1) mix<1><<<148, threads>>>(cons, prod, dm, chain_len, cyc, sink);
        if (nch == 4) mix<4><<<148, threads>>>(cons, prod, dm, chain_len, cyc, sink);
        if (nch == 16) mix<16><<<148, threads>>>(cons, prod, dm, chain_len, cyc, sink);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, 64 * 8, cudaMemcpyDeviceToHost);
        double tf = cons ? 2.0 * 256 * 8 * dm * cons * 148 / (ms * 1e-3) * 1e-12 : 0;
        printf("%2d %2d %2d   %8.1f   [%.2f ms, DMMA %.1f TF]\n", cons, prod, nch, (double)h[0] / chain_len, ms, tf);
    }
    return 0;
}
