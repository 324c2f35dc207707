// sm_100a kernels, part 3 — K2 (headline): Smolyak linear_combination with compile-time unrolled
// leaves.  Same nested (Horner-like) evaluation as sk_nested.cu, restructured so that the FP64
// pipe, not instruction issue, is the limit:
//
//   * The lowest LL (<= 4) dimensions of the index set are a "leaf": for a remaining block budget
//     r, the set {(i_1..i_LL) : Σ b_j <= r} is FIXED, so Leaf<GK, LL, r> is straight-line code
//     produced by template recursion over the block structure of the nested family GK
//     (smolyak_traversal.jl:69-84,112-116,146-148).  No index table, no state machine, no branch:
//     every fma has register operands and a θ read at a compile-time offset from the unit's base.
//   * Chebyshev values/derivatives of leaf dimensions for indices 2..7 (degrees 1..6: blocks 0-2 of
//     all three families) are per-point register tables; higher indices continue the three-term
//     recurrence in registers inside the unrolled code.
//   * Index 1 is T_0 = 1, T_0' = 0: the first child of every node is a register rename.
//   * Dimensions above LL are walked by a tiny uniform interpreter over a unit program
//     (one word per distinct tail (i_{LL+1}..i_d): budget r, levels that wrap afterwards); its
//     per-level state lives in shared memory, [field][thread], touched once per unit.
//   * θ (and the unit program) are staged in shared memory once per CTA; CTAs are persistent.
//
// Reference semantics reproduced (within tolerance; the summation order differs and fma is used):
// smolyak_api.jl:89-110, smolyak_traversal.jl:373-379, derivatives.jl:502-511, generic_api.jl:114-128.
#include <algorithm>
#include <cstdlib>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

// ---- block structure at compile time
template <int GK>
__host__ __device__ constexpr int ellc(int b) {
    if (GK == SK_GRID_ENDPOINT) return b == 0 ? 1 : (1 << b) + 1;
    if (GK == SK_GRID_INTERIOR) { int r = 1; for (int i = 0; i < b; i++) r *= 3; return r; }
    return (1 << (b + 1)) - 1;
}
template <int GK>
__host__ __device__ constexpr int blockof(int i) {
    int b = 0;
    while (ellc<GK>(b) < i) b++;
    return b;
}

constexpr int kTab = 6;      // register table: indices 2..7 (degrees 1..6)

template <bool GRAD>
struct DimTab {
    double t[kTab];
    double dt[GRAD ? kTab : 1];
    double x2, x2d;
};

template <bool GRAD, int T>
struct NC { static constexpr int value = GRAD ? T + 1 : 1; };

template <int GK, int T, int R, bool GRAD>
struct Leaf;

// indices I..ℓ(R) of dimension T
template <int GK, int T, int R, bool GRAD, int I>
struct LeafStep {
    static __device__ __forceinline__ void run(const DimTab<GRAD> *tabs, const double *&th, double (&a)[NC<GRAD, T>::value],
                                               double &Tc, double &Tq, double &dTc, double &dTq) {
        if constexpr (I <= ellc<GK>(R)) {
            constexpr int b = blockof<GK>(I);
            const DimTab<GRAD> &tb = tabs[T - 1];
            double Tv, dTv = 0.0;
            if constexpr (I <= kTab + 1) {
                Tv = tb.t[I - 2];
                if (GRAD) dTv = tb.dt[GRAD ? I - 2 : 0];
            } else {
                if constexpr (I == kTab + 2) {
                    Tc = tb.t[kTab - 1]; Tq = tb.t[kTab - 2];
                    if (GRAD) { dTc = tb.dt[GRAD ? kTab - 1 : 0]; dTq = tb.dt[GRAD ? kTab - 2 : 0]; }
                }
                Tv = fma(tb.x2, Tc, -Tq);
                if (GRAD) { dTv = fma(tb.x2, dTc, fma(tb.x2d, Tc, -dTq)); dTq = dTc; dTc = dTv; }
                Tq = Tc; Tc = Tv;
            }
            double c[NC<GRAD, T - 1>::value];
            Leaf<GK, T - 1, R - b, GRAD>::run(tabs, th, c);
            if (GRAD) {
                a[GRAD ? T : 0] = fma(dTv, c[0], a[GRAD ? T : 0]);
#pragma unroll
                for (int m = 1; m < T; m++) a[GRAD ? m : 0] = fma(Tv, c[GRAD ? m : 0], a[GRAD ? m : 0]);
            }
            a[0] = fma(Tv, c[0], a[0]);
            LeafStep<GK, T, R, GRAD, I + 1>::run(tabs, th, a, Tc, Tq, dTc, dTq);
        }
    }
};

template <int GK, int R, bool GRAD>
struct Leaf<GK, 0, R, GRAD> {
    static __device__ __forceinline__ void run(const DimTab<GRAD> *, const double *&th, double (&a)[1]) {
#ifdef SK_THETA_VOLATILE
        a[0] = *reinterpret_cast<const volatile double *>(th);
#else
        a[0] = *th;
#endif
        th += 1;
    }
};

template <int GK, int T, int R, bool GRAD>
struct Leaf {
    // a[0] = Σ θ·Π T, a[m] = ∂/∂y_m (m = 1..T) over {(i_1..i_T): Σ b_j <= R}; consumes θ in traversal order
    static __device__ __forceinline__ void run(const DimTab<GRAD> *tabs, const double *&th, double (&a)[NC<GRAD, T>::value]) {
#ifdef SK_FENCE_LEVEL
        if constexpr (T == SK_FENCE_LEVEL) __syncwarp();
#endif
        {   // index 1: T_0 = 1, derivative 0
            double c[NC<GRAD, T - 1>::value];
            Leaf<GK, T - 1, R, GRAD>::run(tabs, th, c);
            a[0] = c[0];
            if (GRAD) {
#pragma unroll
                for (int m = 1; m < T; m++) a[GRAD ? m : 0] = c[GRAD ? m : 0];
                a[GRAD ? T : 0] = 0.0;
            }
        }
        double Tc = 0, Tq = 0, dTc = 0, dTq = 0;
        LeafStep<GK, T, R, GRAD, 2>::run(tabs, th, a, Tc, Tq, dTc, dTq);
    }
};

// ---- kernel
struct LeafParams {
    const double *x;
    double *out;
    const double *theta;
    const uint32_t *units;        // r | carry << 8
    int64_t n, K, n_units;
    int D, nup, stage_theta;
    sk_transform tr[SK_MAX_DIM];
};

constexpr int kLeafThreads = 128;
enum { F_X0 = 0, F_X1, F_X2, F_X2D, F_T, F_TP, F_DT, F_DTP, F_ACC };   // upper-level fields, then ACC[0..D]

template <int GK, int LL, int RMAX, bool GRAD>
__global__ void __launch_bounds__(kLeafThreads) smolyak_leaf_kernel(const __grid_constant__ LeafParams q) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int D = q.D, NUP = q.nup;
    const int NF = F_ACC + D + 1;
    double *up = smem;                                                   // [NUP][NF][threads]
    uint32_t *s_units = reinterpret_cast<uint32_t *>(smem + (size_t)NUP * NF * kLeafThreads);
    double *s_theta = reinterpret_cast<double *>(s_units + ((q.n_units + 1) & ~(int64_t)1));
    for (int64_t i = tid; i < q.n_units; i += kLeafThreads) s_units[i] = q.units[i];
    const double *theta = q.theta;
    if (q.stage_theta) {
        for (int64_t i = tid; i < q.K; i += kLeafThreads) s_theta[i] = q.theta[i];
        theta = s_theta;
    }
    __syncthreads();
#define UP(u, f) up[((size_t)(u) * NF + (f)) * kLeafThreads + tid]
    constexpr int NCL = NC<GRAD, LL>::value;
    const int E = GRAD ? D + 1 : 1;

    for (int64_t p = (int64_t)blockIdx.x * kLeafThreads + tid; p < q.n; p += (int64_t)gridDim.x * kLeafThreads) {
        const double *xp = q.x + p * D;
        DimTab<GRAD> tabs[LL];
#pragma unroll
        for (int j = 0; j < LL; j++) {
            double xj[GRAD ? 2 : 1];
            skd::to_pm1_jet<GRAD ? 2 : 1>(q.tr[j], xp[j], xj);
            const double x0 = xj[0], x1 = GRAD ? xj[GRAD ? 1 : 0] : 0.0;
            DimTab<GRAD> &tb = tabs[j];
            tb.x2 = 2.0 * x0; tb.x2d = 2.0 * x1;
            tb.t[0] = x0;
            tb.t[1] = fma(tb.x2, x0, -1.0);
            if (GRAD) { tb.dt[0] = x1; tb.dt[GRAD ? 1 : 0] = tb.x2 * tb.x2d; }
#pragma unroll
            for (int i = 2; i < kTab; i++) {
                tb.t[i] = fma(tb.x2, tb.t[i - 1], -tb.t[i - 2]);
                if (GRAD) tb.dt[GRAD ? i : 0] = fma(tb.x2, tb.dt[GRAD ? i - 1 : 0], fma(tb.x2d, tb.t[i - 1], -tb.dt[GRAD ? i - 2 : 0]));
            }
        }
        for (int u = 0; u < NUP; u++) {
            double xj[GRAD ? 2 : 1];
            skd::to_pm1_jet<GRAD ? 2 : 1>(q.tr[LL + u], xp[LL + u], xj);
            UP(u, F_X0) = xj[0]; UP(u, F_X2) = 2.0 * xj[0];
            if (GRAD) { UP(u, F_X1) = xj[GRAD ? 1 : 0]; UP(u, F_X2D) = 2.0 * xj[GRAD ? 1 : 0]; }
        }
        uint32_t first = 0xFFFFFFFFu;                       // bit u: upper level u sits at index 1
        const double *th = theta;
        double a[NCL];
        for (int64_t un = 0; un < q.n_units; un++) {
            const uint32_t w = s_units[un];
            const int r = (int)(w & 0xFF), c = (int)(w >> 8);
            switch (r) {
            case 0: Leaf<GK, LL, 0, GRAD>::run(tabs, th, a); break;
            case 1: if constexpr (RMAX >= 1) Leaf<GK, LL, RMAX >= 1 ? 1 : 0, GRAD>::run(tabs, th, a); break;
            case 2: if constexpr (RMAX >= 2) Leaf<GK, LL, RMAX >= 2 ? 2 : 0, GRAD>::run(tabs, th, a); break;
            case 3: if constexpr (RMAX >= 3) Leaf<GK, LL, RMAX >= 3 ? 3 : 0, GRAD>::run(tabs, th, a); break;
            case 4: if constexpr (RMAX >= 4) Leaf<GK, LL, RMAX >= 4 ? 4 : 0, GRAD>::run(tabs, th, a); break;
            case 5: if constexpr (RMAX >= 5) Leaf<GK, LL, RMAX >= 5 ? 5 : 0, GRAD>::run(tabs, th, a); break;
            default: break;
            }
            if (NUP > 0) {
                // fold the leaf into upper level 0 (dimension LL+1)
                if (first & 1u) {
#pragma unroll
                    for (int m = 0; m < NCL; m++) UP(0, F_ACC + m) = a[m];
                    if (GRAD) UP(0, F_ACC + LL + 1) = 0.0;
                } else {
                    const double Tv = UP(0, F_T);
#pragma unroll
                    for (int m = 0; m < NCL; m++) UP(0, F_ACC + m) = fma(Tv, a[m], UP(0, F_ACC + m));
                    if (GRAD) UP(0, F_ACC + LL + 1) = fma(UP(0, F_DT), a[0], UP(0, F_ACC + LL + 1));
                }
                // levels 0..c-1 wrap (fold upwards), level c advances
                for (int u = 0; u < NUP; u++) {
                    if (u < c) {
                        if (u + 1 < NUP) {
                            const int nc = GRAD ? LL + u + 2 : 1;              // components of level u
                            if (first & (2u << u)) {
                                for (int m = 0; m < nc; m++) UP(u + 1, F_ACC + m) = UP(u, F_ACC + m);
                                if (GRAD) UP(u + 1, F_ACC + nc) = 0.0;
                            } else {
                                const double Tv = UP(u + 1, F_T);
                                for (int m = 0; m < nc; m++) UP(u + 1, F_ACC + m) = fma(Tv, UP(u, F_ACC + m), UP(u + 1, F_ACC + m));
                                if (GRAD) UP(u + 1, F_ACC + nc) = fma(UP(u + 1, F_DT), UP(u, F_ACC), UP(u + 1, F_ACC + nc));
                            }
                        }
                        first |= 1u << u;
                    } else {
                        if (first & (1u << u)) {                                 // index 1 -> 2
                            UP(u, F_TP) = 1.0; UP(u, F_T) = UP(u, F_X0);
                            if (GRAD) { UP(u, F_DTP) = 0.0; UP(u, F_DT) = UP(u, F_X1); }
                            first &= ~(1u << u);
                        } else {
                            const double X2 = UP(u, F_X2), Tc = UP(u, F_T), Tq = UP(u, F_TP);
                            UP(u, F_T) = fma(X2, Tc, -Tq); UP(u, F_TP) = Tc;
                            if (GRAD) {
                                const double dTc = UP(u, F_DT);
                                UP(u, F_DT) = fma(X2, dTc, fma(UP(u, F_X2D), Tc, -UP(u, F_DTP)));
                                UP(u, F_DTP) = dTc;
                            }
                        }
                        break;
                    }
                }
            }
        }
        double *o = q.out + p * E;
        if (NUP > 0) {
            for (int m = 0; m < E; m++) o[m] = UP(NUP - 1, F_ACC + m);
        } else {
#pragma unroll
            for (int m = 0; m < NCL; m++) o[m] = a[m];
        }
    }
#undef UP
}

// ---- host side: unit program and dispatch ------------------------------------------------------
constexpr int kLeafMaxLLBuilt = 4;
// number of unrolled leaf dimensions; SK_LEAF_LL=1..4 overrides for experiments (read once)
static int leaf_max_ll() {
    static int v = [] {
        const char *e = getenv("SK_LEAF_LL");
        int x = e ? atoi(e) : kLeafMaxLLBuilt;
        return x >= 1 && x <= kLeafMaxLLBuilt ? x : kLeafMaxLLBuilt;
    }();
    return v;
}
#define kLeafMaxLL leaf_max_ll()
static int leaf_rmax(int gk, int LL) {
    // largest budget with pre-built unrolled code (code size / compile time bound)
    if (gk == SK_GRID_INTERIOR) return LL >= 3 ? 3 : 4;
    return LL >= 4 ? 5 : 5;
}

bool smolyak_leaf_supported(const sk_plan &plan) {
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.fast_mode < 0) return false;
    const int d = plan.basis.d, B = plan.basis.B, M = plan.basis.M;
    if (M != B) return false;                       // block cap below the budget: generic kernel
    const int LL = std::min(d, kLeafMaxLL);
    if (B > leaf_rmax(plan.basis.grid_kind, LL)) return false;
    if (d - LL > 12) return false;
    return true;
}

// units: distinct tails (i_{LL+1}..i_d) in traversal order; word = r | carry << 8
void smolyak_leaf_units(const sk_plan &plan, std::vector<uint32_t> &units) {
    const skh::SmolyakSet &S = plan.set;
    const int d = S.d, LL = std::min(d, kLeafMaxLL), nup = d - LL;
    units.clear();
    if (nup == 0) { units.push_back((uint32_t)S.B); return; }
    std::vector<int> blk_of((size_t)S.H + 2, S.M);
    for (int b = S.M - 1; b >= 0; b--) for (int64_t i = 1; i <= S.ell[(size_t)b]; i++) blk_of[(size_t)i] = b;
    std::vector<int> idx((size_t)nup, 1), blk((size_t)nup, 0);
    int used = 0;
    for (;;) {
        const uint32_t r = (uint32_t)(S.B - used);
        int j = 0;
        for (; j < nup; j++) {
            int cand = idx[(size_t)j] + 1;
            if (cand <= S.H) {
                int nb = blk_of[(size_t)cand];
                if (used - blk[(size_t)j] + nb <= S.B) { used += nb - blk[(size_t)j]; idx[(size_t)j] = cand; blk[(size_t)j] = nb; break; }
            }
            used -= blk[(size_t)j]; idx[(size_t)j] = 1; blk[(size_t)j] = 0;
        }
        units.push_back(r | ((uint32_t)j << 8));      // j levels wrapped (j == nup: end)
        if (j == nup) break;
    }
}

// closed-form FP64 flop count per point of the leaf kernel (fma = 2, mul/add = 1; transforms 10/coordinate)
static double leaf_flops_rec(const skh::SmolyakSet &S, int T, int R, bool g, std::vector<std::vector<double>> &memo) {
    if (T == 0) return 0.0;
    double &m = memo[(size_t)T][(size_t)R];
    if (m >= 0) return m;
    double f = leaf_flops_rec(S, T - 1, R, g, memo);
    int b = 1;
    for (int64_t i = 2; i <= S.ell[(size_t)R]; i++) {
        while (S.ell[(size_t)b] < i) b++;
        if (i > kTab + 1) f += g ? 6 : 2;
        f += leaf_flops_rec(S, T - 1, R - b, g, memo);
        f += g ? 2.0 * (T + 1) : 2.0;
    }
    return m = f;
}
double smolyak_leaf_flops(const sk_plan &plan) {
    const skh::SmolyakSet &S = plan.set;
    const bool g = plan.fast_mode == 1;
    const int d = S.d, LL = std::min(d, kLeafMaxLL), nup = d - LL;
    std::vector<std::vector<double>> memo((size_t)LL + 1, std::vector<double>((size_t)S.B + 1, -1.0));
    std::vector<uint32_t> units;
    smolyak_leaf_units(plan, units);
    double f = 12.0 * d + LL * (g ? 3 + (kTab - 2) * 6 : (kTab - 1) * 2);
    std::vector<char> first((size_t)std::max(nup, 1), 1);
    for (uint32_t w : units) {
        const int r = (int)(w & 0xFF), c = (int)(w >> 8);
        f += leaf_flops_rec(S, LL, r, g, memo);
        if (nup == 0) continue;
        if (!first[0]) f += g ? 2.0 * (LL + 2) : 2.0;
        for (int u = 0; u < nup; u++) {
            if (u < c) {
                if (u + 1 < nup && !first[(size_t)u + 1]) f += g ? 2.0 * (LL + u + 3) : 2.0;
                first[(size_t)u] = 1;
            } else {
                if (!first[(size_t)u]) f += g ? 6 : 2;
                first[(size_t)u] = 0;
                break;
            }
        }
    }
    return f;
}

template <int GK, int LL, int RMAX, bool GRAD>
static cudaError_t launch_leaf(const LeafParams &q, size_t smem, cudaStream_t stream) {
    auto kern = smolyak_leaf_kernel<GK, LL, RMAX, GRAD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1, dev = 0, sms = 148;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLeafThreads, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (q.n + kLeafThreads - 1) / kLeafThreads;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * std::max(per_sm, 1));
    kern<<<grid, kLeafThreads, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

template <int GK, bool GRAD>
static cudaError_t dispatch_leaf_ll(int LL, const LeafParams &q, size_t smem, cudaStream_t stream) {
    constexpr int R12 = GK == SK_GRID_INTERIOR ? 4 : 5;
    constexpr int R34 = GK == SK_GRID_INTERIOR ? 3 : 5;
    switch (LL) {
    case 1: return launch_leaf<GK, 1, R12, GRAD>(q, smem, stream);
    case 2: return launch_leaf<GK, 2, R12, GRAD>(q, smem, stream);
    case 3: return launch_leaf<GK, 3, R34, GRAD>(q, smem, stream);
    case 4: return launch_leaf<GK, 4, R34, GRAD>(q, smem, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t smolyak_leaf_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n, double *out,
                                 cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int d = plan.basis.d, LL = std::min(d, kLeafMaxLL), nup = d - LL;
    LeafParams q{};
    q.x = x; q.out = out; q.theta = theta; q.units = plan.d_units;
    q.n = n; q.K = plan.K; q.n_units = plan.n_units; q.D = d; q.nup = nup;
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    const size_t up_bytes = (size_t)nup * (F_ACC + d + 1) * kLeafThreads * 8;
    const size_t unit_bytes = (size_t)((plan.n_units + 1) & ~(int64_t)1) * 4;
    size_t smem = up_bytes + unit_bytes;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    q.stage_theta = smem + (size_t)plan.K * 8 <= 100 * 1024;
    if (q.stage_theta) smem += (size_t)plan.K * 8;
    const bool g = plan.fast_mode == 1;
#ifdef SK_LEAF_EXPERIMENT
    return launch_leaf<SK_GRID_INTERIOR2, 4, SK_LEAF_EXPERIMENT, true>(q, smem, stream);
#else
    switch (plan.basis.grid_kind) {
    case SK_GRID_INTERIOR: return g ? dispatch_leaf_ll<SK_GRID_INTERIOR, true>(LL, q, smem, stream) : dispatch_leaf_ll<SK_GRID_INTERIOR, false>(LL, q, smem, stream);
    case SK_GRID_ENDPOINT: return g ? dispatch_leaf_ll<SK_GRID_ENDPOINT, true>(LL, q, smem, stream) : dispatch_leaf_ll<SK_GRID_ENDPOINT, false>(LL, q, smem, stream);
    case SK_GRID_INTERIOR2: return g ? dispatch_leaf_ll<SK_GRID_INTERIOR2, true>(LL, q, smem, stream) : dispatch_leaf_ll<SK_GRID_INTERIOR2, false>(LL, q, smem, stream);
    }
#endif
    return cudaErrorInvalidValue;
}

}  // namespace skk
