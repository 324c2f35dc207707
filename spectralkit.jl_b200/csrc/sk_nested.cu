// sm_100a kernels, part 2 — K2: Smolyak linear_combination, nested run-structured evaluation.
//
// What the reference does per point (smolyak_api.jl:89-110, smolyak_traversal.jl:373-379,
// derivatives.jl:502-511, generic_api.jl:114-128): build d tables of H jets, then for each of the
// K index tuples re-derive the tuple with a state machine, multiply d table entries per output
// component and accumulate θ_k·term.  K·(C·(d−1)+2C) flops, K steps of integer bookkeeping.
//
// What this kernel does instead (B200-first):
//   * The traversal is column-major, so θ is a sequence of RUNS: one run per distinct tail
//     (i_2..i_d), of length ℓ(remaining slack).  The plan turns the index
//     set into a run program once: (run length, number of upper levels that wrap afterwards).
//     The program and θ are staged in shared memory; every thread of the grid executes the same
//     program, so all control flow is warp-uniform and all θ/program reads are broadcasts.
//   * One point per thread, everything in registers.  The sum is evaluated dimension by
//     dimension (Horner-like): a run is a univariate Chebyshev sum in x_1; finished runs are
//     folded into level 2 with T_{i_2}(x_2), and so on.  Level t carries (value, ∂_1..∂_t), so
//     the gradient costs (t+1) fma per fold instead of C·d multiplies per term.
//   * T_1 = 1, T_1' = 0: the first child of every node is an assignment, not a multiply.
//   * Chebyshev values of level 1 for degrees ≤ 6 (blocks 0–2 of all three nested families)
//     are a per-point register table; higher degrees and all upper levels advance a three-term
//     recurrence in registers exactly when their index advances.  No per-point table ever
//     touches shared or local memory: the FP64 pipe is fed from registers and broadcast loads.
//   * All multiply-adds are explicit fma (the file is compiled with -fmad=false).
//
// Result: within the stated tolerance of the reference (different summation order, fma), never
// bit-identical; SK_MODE_EXACT selects the direct kernel in sk_kernels.cu instead.
#include <algorithm>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

struct NestedParams {
    const double *x;
    double *out;
    const double *theta;
    const uint32_t *runs;
    int64_t n, K, n_runs;
    int stage;                       // θ and the run program staged in shared memory
    sk_transform tr[SK_MAX_DIM];
};

constexpr int kLow = 5;              // register table of level 1: degrees 2..6

template <int D, bool GRAD>
struct Levels {
    double x0[D], x1[D], x2[D], x2d[D];      // coordinate, chain factor dx/dy, and their doubles
    double T[D], Tp[D], dT[D], dTp[D];       // running recurrence of levels 1..D-1 (index 0 unused)
    double av[D];                            // accumulated value per level
    double ag[D][D];                         // accumulated partials: ag[t][m], m <= t
    bool first[D];                           // level index is 1 (T = 1, T' = 0)
};

// fold the finished child (level t-1) into level t with the current T_{i_t}(x_t)
template <int t, int D, bool GRAD>
__device__ __forceinline__ void fold(Levels<D, GRAD> &s) {
    if (s.first[t]) {
        s.av[t] = s.av[t - 1];
        if (GRAD) {
#pragma unroll
            for (int m = 0; m < t; m++) s.ag[t][m] = s.ag[t - 1][m];
            s.ag[t][t] = 0.0;
        }
    } else {
        if (GRAD) {
            s.ag[t][t] = fma(s.dT[t], s.av[t - 1], s.ag[t][t]);
#pragma unroll
            for (int m = 0; m < t; m++) s.ag[t][m] = fma(s.T[t], s.ag[t - 1][m], s.ag[t][m]);
        }
        s.av[t] = fma(s.T[t], s.av[t - 1], s.av[t]);
    }
}

// index of level t advances by one: step its recurrence
template <int t, int D, bool GRAD>
__device__ __forceinline__ void advance(Levels<D, GRAD> &s) {
    if (s.first[t]) {            // index 1 -> 2: T_1(x) = x, derivative = chain factor
        s.Tp[t] = 1.0; s.T[t] = s.x0[t];
        if (GRAD) { s.dTp[t] = 0.0; s.dT[t] = s.x1[t]; }
        s.first[t] = false;
    } else {
        double Tn = fma(s.x2[t], s.T[t], -s.Tp[t]);
        if (GRAD) {
            double dTn = fma(s.x2[t], s.dT[t], fma(s.x2d[t], s.T[t], -s.dTp[t]));
            s.dTp[t] = s.dT[t]; s.dT[t] = dTn;
        }
        s.Tp[t] = s.T[t]; s.T[t] = Tn;
    }
}

// after a run: levels 1..c wrap (fold upwards, reset), level c+1 advances
template <int t, int D, bool GRAD>
__device__ __forceinline__ void carry(Levels<D, GRAD> &s, int c) {
    if constexpr (t < D) {
        if (c >= t) {
            if constexpr (t + 1 < D) fold<t + 1, D, GRAD>(s);
            s.first[t] = true;
            carry<t + 1, D, GRAD>(s, c);
        } else {
            advance<t, D, GRAD>(s);
        }
    }
}

template <int D, bool GRAD>
__global__ void __launch_bounds__(128) smolyak_nested_kernel(const __grid_constant__ NestedParams q) {
    extern __shared__ double smem[];
    const double *theta = q.theta;
    const uint32_t *runs = q.runs;
    if (q.stage) {
        double *s_theta = smem;
        uint32_t *s_runs = reinterpret_cast<uint32_t *>(smem + q.K);
        for (int64_t i = threadIdx.x; i < q.K; i += blockDim.x) s_theta[i] = q.theta[i];
        for (int64_t i = threadIdx.x; i < q.n_runs; i += blockDim.x) s_runs[i] = q.runs[i];
        __syncthreads();
        theta = s_theta;
        runs = s_runs;
    }
    constexpr int E = GRAD ? D + 1 : 1;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < q.n; p += (int64_t)gridDim.x * blockDim.x) {
        Levels<D, GRAD> s;
#pragma unroll
        for (int j = 0; j < D; j++) {
            double xj[GRAD ? 2 : 1];
            skd::to_pm1_jet<GRAD ? 2 : 1>(q.tr[j], q.x[p * D + j], xj);
            s.x0[j] = xj[0];
            s.x1[j] = GRAD ? xj[GRAD ? 1 : 0] : 0.0;
            s.x2[j] = 2.0 * s.x0[j];
            s.x2d[j] = 2.0 * s.x1[j];
            s.first[j] = true;
        }
        // level-1 register table: degrees 2..6 (indices 3..7)
        double lt[kLow], ldt[kLow];
        {
            const double X2 = s.x2[0], X2d = s.x2d[0];
            lt[0] = fma(X2, s.x0[0], -1.0);
            if (GRAD) ldt[0] = X2 * X2d;                           // 4·x0·x1
            double tp = s.x0[0], dtp = s.x1[0];
#pragma unroll
            for (int i = 1; i < kLow; i++) {
                lt[i] = fma(X2, lt[i - 1], -tp);
                if (GRAD) ldt[i] = fma(X2, ldt[i - 1], fma(X2d, lt[i - 1], -dtp));
                tp = lt[i - 1];
                if (GRAD) dtp = ldt[i - 1];
            }
        }
        const double *th = theta;
        for (int64_t r = 0; r < q.n_runs; r++) {
            const uint32_t w = runs[r];
            const int L = (int)(w & 0xFFFFFFu), c = (int)(w >> 24);
            // ---- level 1: univariate sum over the run
            double a0 = th[0], a1 = 0.0;
            if (L > 1) {                                   // every family has ℓ(1) = 3
                const double t1 = th[1], t2 = th[2];
                a0 = fma(t1, s.x0[0], a0);
                a0 = fma(t2, lt[0], a0);
                if (GRAD) { a1 = t1 * s.x1[0]; a1 = fma(t2, ldt[0], a1); }
                if (L > 3) {                               // ℓ(2) >= 5
                    const double t3 = th[3], t4 = th[4];
                    a0 = fma(t3, lt[1], a0);
                    a0 = fma(t4, lt[2], a0);
                    if (GRAD) { a1 = fma(t3, ldt[1], a1); a1 = fma(t4, ldt[2], a1); }
                    if (L > 5) {                           // >= 7
                        const double t5 = th[5], t6 = th[6];
                        a0 = fma(t5, lt[3], a0);
                        a0 = fma(t6, lt[4], a0);
                        if (GRAD) { a1 = fma(t5, ldt[3], a1); a1 = fma(t6, ldt[4], a1); }
                        if (L > 7) {                       // continue the recurrence past degree 6
                            double Tc = lt[4], Tq = lt[3], dTc = GRAD ? ldt[4] : 0.0, dTq = GRAD ? ldt[3] : 0.0;
                            for (int i = 7; i < L; i++) {
                                const double Tn = fma(s.x2[0], Tc, -Tq);
                                const double ti = th[i];
                                if (GRAD) {
                                    const double dTn = fma(s.x2[0], dTc, fma(s.x2d[0], Tc, -dTq));
                                    dTq = dTc; dTc = dTn;
                                    a1 = fma(ti, dTn, a1);
                                }
                                Tq = Tc; Tc = Tn;
                                a0 = fma(ti, Tn, a0);
                            }
                        }
                    }
                }
            }
            th += L;
            s.av[0] = a0;
            if (GRAD) s.ag[0][0] = a1;
            // ---- upper levels
            if constexpr (D > 1) {
                fold<1, D, GRAD>(s);
                carry<1, D, GRAD>(s, c);
            }
        }
        double *o = q.out + p * E;
        o[0] = s.av[D - 1];
        if (GRAD) {
#pragma unroll
            for (int m = 0; m < D; m++) o[1 + m] = s.ag[D - 1][m];
        }
    }
}

static int sm_count_n() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

constexpr int kMaxNestedDim = 10;

bool smolyak_nested_supported(const sk_plan &plan) {
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.fast_mode < 0) return false;
    return plan.basis.d >= 1 && plan.basis.d <= kMaxNestedDim;
}

template <int D, bool GRAD>
static cudaError_t launch_nested(const NestedParams &q, size_t smem, cudaStream_t stream) {
    auto kern = smolyak_nested_kernel<D, GRAD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t tiles = (q.n + 127) / 128;
    int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count_n() * per_sm);
    kern<<<grid, 128, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

template <bool GRAD>
static cudaError_t dispatch_nested(int d, const NestedParams &q, size_t smem, cudaStream_t stream) {
    switch (d) {
    case 1: return launch_nested<1, GRAD>(q, smem, stream);
    case 2: return launch_nested<2, GRAD>(q, smem, stream);
    case 3: return launch_nested<3, GRAD>(q, smem, stream);
    case 4: return launch_nested<4, GRAD>(q, smem, stream);
    case 5: return launch_nested<5, GRAD>(q, smem, stream);
    case 6: return launch_nested<6, GRAD>(q, smem, stream);
    case 7: return launch_nested<7, GRAD>(q, smem, stream);
    case 8: return launch_nested<8, GRAD>(q, smem, stream);
    case 9: return launch_nested<9, GRAD>(q, smem, stream);
    case 10: return launch_nested<10, GRAD>(q, smem, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t smolyak_nested_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n, double *out,
                                   cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    NestedParams q{};
    q.x = x; q.out = out; q.theta = theta; q.runs = plan.d_runs;
    q.n = n; q.K = plan.K; q.n_runs = plan.n_runs;
    const int d = plan.basis.d;
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    size_t smem = (size_t)plan.K * 8 + (size_t)plan.n_runs * 4;
    q.stage = smem <= 96 * 1024;
    if (!q.stage) smem = 0;
    return plan.fast_mode == 1 ? dispatch_nested<true>(d, q, smem, stream) : dispatch_nested<false>(d, q, smem, stream);
}

// Closed-form FP64 flop count per point of the kernel above (fma = 2, mul/add = 1), obtained by
// walking the same run program; transforms are charged 10 flops per coordinate like SURVEY.md §8(d).
double smolyak_nested_flops(const sk_plan &plan) {
    const int d = plan.basis.d;
    const bool g = plan.fast_mode == 1;
    double f = 10.0 * d + 2.0 * d;                       // transforms, 2x and 2x'
    f += g ? (2 + 1) + (kLow - 1) * (2 + 4) : kLow * 2;  // level-1 register table
    std::vector<char> first((size_t)d, 1);
    for (int64_t r = 0; r < plan.n_runs; r++) {
        const int64_t L = plan.set.run_len[(size_t)r];
        const int c = plan.set.run_carry[(size_t)r];
        if (L > 1) f += g ? 7 : 4;
        if (L > 3) f += g ? 8 : 4;
        if (L > 5) f += g ? 8 : 4;
        if (L > 7) f += (double)(L - 7) * (g ? 10 : 4);
        if (d > 1) {
            if (!first[1]) f += g ? 2.0 * (1 + 2) : 2.0;   // fold into level 2 (t = 1)
            for (int t = 1; t < d; t++) {
                if (c >= t) {
                    if (t + 1 < d && !first[(size_t)t + 1]) f += g ? 2.0 * (t + 1 + 2) : 2.0;
                    first[(size_t)t] = 1;
                } else {
                    if (!first[(size_t)t]) f += g ? 6 : 2;
                    first[(size_t)t] = 0;
                    break;
                }
            }
        }
    }
    return f;
}

}  // namespace skk
