// sm_100a kernels, part 3d — plain values, TWO POINTS PER THREAD on the unrolled leaves (one coefficient vector).
//
// The values-only leaf kernel is bound by the register file, not the FP64 pipe: fma(T, θ, a) reads three fresh 64-bit
// operands (3 cycles on a pipe that needs 2) and every coefficient costs one more register-file cycle when its LDS writes it
// back (profiles/r2_leaf_register_file.md) — 46 % of the DFMA peak at cfg 4a.  With two points per thread a coefficient is
// loaded once for two fmas (half the LDS instructions and write-backs per fma), it is the SAME operand of two consecutive
// fmas, every thread carries two independent dependency chains, and the code is fetched once for two points: 2 700 -> 3 870
// Mpoints/s at cfg 4a (66 % of the DFMA peak), 4 983 -> 6 802 at d = 5 (profiles/r2_values_two_points.jsonl).  Tables, recurrences and accumulators exist per point; the template
// recursion, the coefficient layout (leaf_end) and the per-block loops are those of sk_leaf_impl.cuh.  One-leaf bases (d <= 6,
// block cap M = B).  Fast mode only (fma, nested summation order): within the stated tolerance of the reference.
#pragma once
#include "sk_leaf_impl.cuh"

namespace skk {

constexpr int kPP = 2;      // points per thread

template <int TREG_, int LOOPD_, int NT_>
struct PtsCtx {
    static constexpr int TREG = TREG_, LOOPD = LOOPD_, NT = NT_;
    double t[TREG_ > 0 ? TREG_ : 1][kTab][kPP];      // T_2..T_7 of the dimensions with register tables, per point
    double x2[TREG_ > 0 ? TREG_ : 1][kPP];
    const double *stab;                              // shared tables of the dimensions >= TREG: [dim][entry | 2x][point of the thread][thread]
};

template <int T, int E, class CX>
__device__ __forceinline__ void ptab_get(const CX &cx, double (&Tv)[kPP]) {
#pragma unroll
    for (int v = 0; v < kPP; v++) {
        if constexpr (T - 1 < CX::TREG) Tv[v] = cx.t[T - 1][E][v];
        else Tv[v] = cx.stab[(((T - 1 - CX::TREG) * (kTab + 1) + E) * kPP + v) * CX::NT];
    }
}
template <int T, class CX>
__device__ __forceinline__ void ptab_x2(const CX &cx, double (&x2)[kPP]) {
#pragma unroll
    for (int v = 0; v < kPP; v++) {
        if constexpr (T - 1 < CX::TREG) x2[v] = cx.x2[T - 1][v];
        else x2[v] = cx.stab[(((T - 1 - CX::TREG) * (kTab + 1) + kTab) * kPP + v) * CX::NT];
    }
}

template <int GK, int T, int R, class CX>
struct PLeaf;

template <int GK, int T, int R, class CX, int I>
struct PLeafStep {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[kPP], double (&Tc)[kPP], double (&Tq)[kPP]) {
        if constexpr (I <= ellc<GK>(R)) {
            constexpr int b = blockof<GK>(I);
            double Tv[kPP];
            if constexpr (I <= kTab + 1) {
                ptab_get<T, I - 2>(cx, Tv);
            } else {
                double x2[kPP];
                ptab_x2<T>(cx, x2);
                if constexpr (I == kTab + 2) { ptab_get<T, kTab - 1>(cx, Tc); ptab_get<T, kTab - 2>(cx, Tq); }
#pragma unroll
                for (int v = 0; v < kPP; v++) { Tv[v] = fma(x2[v], Tc[v], -Tq[v]); Tq[v] = Tc[v]; Tc[v] = Tv[v]; }
            }
            if constexpr (R - b == 0) {
                const double c0 = th_get(th, off);      // one coefficient, two fmas
                off += 1;
#pragma unroll
                for (int v = 0; v < kPP; v++) a[v] = fma(Tv[v], c0, a[v]);
            } else {
                double c[kPP];
                off = PLeaf<GK, T - 1, R - b, CX>::run(cx, th, off, c);
#pragma unroll
                for (int v = 0; v < kPP; v++) a[v] = fma(Tv[v], c[v], a[v]);
            }
            return PLeafStep<GK, T, R, CX, I + 1>::run(cx, th, off, a, Tc, Tq);
        } else {
            return off;
        }
    }
};

template <int GK, int T, int R, class CX, int B>
struct PLeafBlocks {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[kPP], double (&Tc)[kPP], double (&Tq)[kPP]) {
        if constexpr (B <= R) {
            constexpr int lo = ellc<GK>(B - 1) + 1, hi = ellc<GK>(B);
            constexpr int CS = R - B == 0 ? 1 : even_up(leaf_end<GK>(CX::LOOPD, T - 1, R - B > 0 ? R - B : 0, 0));
            const int start = R - B == 0 ? off : even_up(off);
            const double *blk = th.base + start;
            const double *tb = cx.stab + (size_t)(T - 1 - CX::TREG) * (kTab + 1) * kPP * CX::NT;
            double x2[kPP] = {};
            if constexpr (hi > kTab + 1) {
#pragma unroll
                for (int v = 0; v < kPP; v++) {
                    x2[v] = tb[(kTab * kPP + v) * CX::NT];
                    if constexpr (lo <= kTab + 2) { Tc[v] = tb[((kTab - 1) * kPP + v) * CX::NT]; Tq[v] = tb[((kTab - 2) * kPP + v) * CX::NT]; }
                }
            }
#pragma unroll 1
            for (int i = lo; i <= hi; i++) {
                double Tv[kPP];
#pragma unroll
                for (int v = 0; v < kPP; v++) {
                    if (hi <= kTab + 1 || i <= kTab + 1) Tv[v] = tb[((i - 2) * kPP + v) * CX::NT];
                    else { Tv[v] = fma(x2[v], Tc[v], -Tq[v]); Tq[v] = Tc[v]; Tc[v] = Tv[v]; }
                }
                if constexpr (R - B == 0) {
                    const double c0 = blk[i - lo];
#pragma unroll
                    for (int v = 0; v < kPP; v++) a[v] = fma(Tv[v], c0, a[v]);
                } else {
                    double c[kPP];
                    ThRd sub{blk + (i - lo) * CS, 0.0};
                    PLeaf<GK, T - 1, R - B, CX>::run(cx, sub, 0, c);
#pragma unroll
                    for (int v = 0; v < kPP; v++) a[v] = fma(Tv[v], c[v], a[v]);
                }
            }
            return PLeafBlocks<GK, T, R, CX, B + 1>::run(cx, th, even_up(start + (hi - lo + 1) * CS), a, Tc, Tq);
        } else {
            return off;
        }
    }
};

template <int GK, int R, class CX>
struct PLeaf<GK, 0, R, CX> {
    static __device__ __forceinline__ int run(const CX &, ThRd &th, int off, double (&a)[kPP]) {
        const double c0 = th_get(th, off);
#pragma unroll
        for (int v = 0; v < kPP; v++) a[v] = c0;
        return off + 1;
    }
};

template <int GK, int T, int R, class CX>
struct PLeaf {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[kPP]) {
        if constexpr (R == 0) {
            const double c0 = th_get(th, off);
#pragma unroll
            for (int v = 0; v < kPP; v++) a[v] = c0;
            return off + 1;
        } else {
            off = PLeaf<GK, T - 1, R, CX>::run(cx, th, off, a);      // index 1: T_0 = 1
            double Tc[kPP] = {}, Tq[kPP] = {};
            if constexpr (T < CX::LOOPD) return PLeafStep<GK, T, R, CX, 2>::run(cx, th, off, a, Tc, Tq);
            else return PLeafBlocks<GK, T, R, CX, 1>::run(cx, th, off, a, Tc, Tq);
        }
    }
};

struct PtsParams {
    const double *x;
    double *out;
    const double *theta;
    const uint32_t *thmap;        // slot of every coefficient (layout of leaf_end with this kernel's LOOPD)
    int64_t n, K, Kslots;
    int64_t theta_stride, out_stride;      // coefficient vector v of a K×nvec block: θ[k·nvec + v] → out[p·nvec + v]
    sk_transform tr[8];
};

constexpr size_t pts_smem_bytes(int NT, int LL, int TREG, int64_t Kslots) {
    return (size_t)(LL > TREG ? LL - TREG : 0) * (kTab + 1) * kPP * NT * 8 + (size_t)Kslots * 8 + 16;
}

template <int GK, int LL, int R, int TREG, int LOOPD, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) smolyak_pts_kernel(const __grid_constant__ PtsParams q) {
    static_assert(LOOPD > TREG, "looped dimensions read their table from shared memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    double *s_theta = reinterpret_cast<double *>(smem_raw);
    double *stab = reinterpret_cast<double *>(smem_raw + (((size_t)q.Kslots * 8 + 15) & ~(size_t)15)) + tid;
    for (int64_t i = tid; i < q.K; i += NT) s_theta[q.thmap[i]] = q.theta[i * q.theta_stride];
    __syncthreads();
    using CX = PtsCtx<TREG, LOOPD, NT>;
    for (int64_t p0 = (int64_t)blockIdx.x * NT * kPP; p0 < q.n; p0 += (int64_t)gridDim.x * NT * kPP) {
        CX cx;
        cx.stab = stab;
        int64_t pv[kPP];
#pragma unroll
        for (int v = 0; v < kPP; v++) pv[v] = min(p0 + tid + (int64_t)v * NT, q.n - 1);      // thread's points: NT apart (coalesced)
#pragma unroll
        for (int j = 0; j < LL; j++) {
#pragma unroll
            for (int v = 0; v < kPP; v++) {
                const double x0 = skd::to_pm1(q.tr[j], __ldcs(q.x + pv[v] * LL + j));
                const double x2 = 2.0 * x0;
                double t[kTab];
                t[0] = x0;
                t[1] = fma(x2, x0, -1.0);
#pragma unroll
                for (int i = 2; i < kTab; i++) t[i] = fma(x2, t[i - 1], -t[i - 2]);
                if (j < TREG) {
                    cx.x2[j < TREG ? j : 0][v] = x2;
#pragma unroll
                    for (int i = 0; i < kTab; i++) cx.t[j < TREG ? j : 0][i][v] = t[i];
                } else {
                    double *sd = stab + (size_t)(j - TREG) * (kTab + 1) * kPP * NT;
#pragma unroll
                    for (int i = 0; i < kTab; i++) sd[(i * kPP + v) * NT] = t[i];
                    sd[(kTab * kPP + v) * NT] = x2;
                }
            }
        }
        double a[kPP];
        ThRd th{s_theta, 0.0};
        PLeaf<GK, LL, R, CX>::run(cx, th, 0, a);
#pragma unroll
        for (int v = 0; v < kPP; v++) if (p0 + tid + (int64_t)v * NT < q.n) __stcs(q.out + (p0 + tid + (int64_t)v * NT) * q.out_stride, a[v]);
    }
}

template <int GK, int LL, int R, int TREG, int LOOPD, int NT, int MINB>
cudaError_t launch_pts(const PtsParams &q, cudaStream_t stream) {
    auto kern = smolyak_pts_kernel<GK, LL, R, TREG, LOOPD, NT, MINB>;
    const size_t smem = pts_smem_bytes(NT, LL, TREG, q.Kslots);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1, dev = 0, sms = 148;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (q.n + (int64_t)NT * kPP - 1) / ((int64_t)NT * kPP);
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * std::max(per_sm, 1));
    kern<<<grid, NT, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

// ---- which (family, depth, budget) are built, and their shapes
constexpr bool pts_built_c(int gk, int LL, int R) {
    if (LL < 3 || LL > 8 || R < 2) return false;      // small leaves are bandwidth-bound already
    if (LL > 6) return gk != SK_GRID_INTERIOR && R <= 3;      // seven and eight dimensions: budgets 2, 3 (799 / 1 121 terms)
    return gk == SK_GRID_INTERIOR ? R <= (LL <= 4 ? 4 : 3) : R <= (LL <= 4 ? 5 : 4);
}
// measured at InteriorGrid2 B = 4, 2²³ points (one point per thread: d = 6 2 700, d = 5 4 983 Mpoints/s): 256 threads with three
// register tables 3 770 / 6 802, two tables 3 870 / 6 341, 384 threads 3 487 / 6 229 (three tables), 3 179 / 6 221 (two), 320: 2 827 / 5 134
#ifndef SK_PTS_NT
#define SK_PTS_NT 256
#endif
constexpr int pts_treg_c(int LL) { return LL == 6 ? 2 : LL < 3 ? LL : 3; }
constexpr int pts_loopd_c(int, int, int) { return 99; }      // values-only code is short: straight-line (as the one-point kernel)
constexpr int pts_nt_c(int) { return SK_PTS_NT; }

template <int GK, int LL, int R>
cudaError_t launch_pts_built(const PtsParams &q, cudaStream_t stream) {
    if constexpr (pts_built_c(GK, LL, R))
        return launch_pts<GK, LL, R, pts_treg_c(LL), pts_loopd_c(GK, LL, R), pts_nt_c(LL), 1>(q, stream);
    else
        return cudaErrorNotSupported;
}
template <int GK, int LL>
cudaError_t dispatch_pts_r(int R, const PtsParams &q, cudaStream_t stream) {
    switch (R) {
    case 2: return launch_pts_built<GK, LL, 2>(q, stream);
    case 3: return launch_pts_built<GK, LL, 3>(q, stream);
    case 4: return launch_pts_built<GK, LL, 4>(q, stream);
    case 5: return launch_pts_built<GK, LL, 5>(q, stream);
    }
    return cudaErrorNotSupported;
}
template <int GK>
cudaError_t dispatch_pts(int LL, int R, const PtsParams &q, cudaStream_t stream) {
    switch (LL) {
    case 7: return dispatch_pts_r<GK, 7>(R, q, stream);
    case 8: return dispatch_pts_r<GK, 8>(R, q, stream);
    case 3: return dispatch_pts_r<GK, 3>(R, q, stream);
    case 4: return dispatch_pts_r<GK, 4>(R, q, stream);
    case 5: return dispatch_pts_r<GK, 5>(R, q, stream);
    case 6: return dispatch_pts_r<GK, 6>(R, q, stream);
    }
    return cudaErrorNotSupported;
}

}  // namespace skk
