// EXPERIMENT: value + gradient with TWO POINTS PER THREAD on the unrolled leaf (one-leaf bases), after the values-only result.
#pragma once
#include "sk_leaf_impl.cuh"

namespace skk {

#ifndef GPP
#define GPP 2
#endif
constexpr int kGP = GPP;

template <int TREG_, int LOOPD_, int NT_>
struct GPtsCtx {
    static constexpr int TREG = TREG_, LOOPD = LOOPD_, NT = NT_;
    double t[TREG_ > 0 ? TREG_ : 1][kTab][kGP], dt[TREG_ > 0 ? TREG_ : 1][kTab][kGP];
    double x2[TREG_ > 0 ? TREG_ : 1][kGP], x2d[TREG_ > 0 ? TREG_ : 1][kGP];
    const double2 *stab;      // [dim][entry | 2x][point of the thread][thread] (T, T')
};

template <int T, int E, class CX>
__device__ __forceinline__ void gtab_get(const CX &cx, double (&Tv)[kGP], double (&dTv)[kGP]) {
#pragma unroll
    for (int v = 0; v < kGP; v++) {
        if constexpr (T - 1 < CX::TREG) { Tv[v] = cx.t[T - 1][E][v]; dTv[v] = cx.dt[T - 1][E][v]; }
        else { const double2 u = cx.stab[(((T - 1 - CX::TREG) * (kTab + 1) + E) * kGP + v) * CX::NT]; Tv[v] = u.x; dTv[v] = u.y; }
    }
}
template <int T, class CX>
__device__ __forceinline__ void gtab_x2(const CX &cx, double (&x2)[kGP], double (&x2d)[kGP]) {
#pragma unroll
    for (int v = 0; v < kGP; v++) {
        if constexpr (T - 1 < CX::TREG) { x2[v] = cx.x2[T - 1][v]; x2d[v] = cx.x2d[T - 1][v]; }
        else { const double2 u = cx.stab[(((T - 1 - CX::TREG) * (kTab + 1) + kTab) * kGP + v) * CX::NT]; x2[v] = u.x; x2d[v] = u.y; }
    }
}

struct GRec { double Tc[kGP], Tq[kGP], dTc[kGP], dTq[kGP]; };

template <int T>
__device__ __forceinline__ void g_fold(double (&a)[T + 1][kGP], const double (&c)[T][kGP], const double (&Tv)[kGP], const double (&dTv)[kGP]) {
#pragma unroll
    for (int v = 0; v < kGP; v++) a[T][v] = fma(dTv[v], c[0][v], a[T][v]);
#pragma unroll
    for (int m = 1; m < T; m++)
#pragma unroll
        for (int v = 0; v < kGP; v++) a[m][v] = fma(Tv[v], c[m][v], a[m][v]);
#pragma unroll
    for (int v = 0; v < kGP; v++) a[0][v] = fma(Tv[v], c[0][v], a[0][v]);
}
template <int T>
__device__ __forceinline__ void g_fold_theta(double (&a)[T + 1][kGP], double c0, const double (&Tv)[kGP], const double (&dTv)[kGP]) {
#pragma unroll
    for (int v = 0; v < kGP; v++) a[T][v] = fma(dTv[v], c0, a[T][v]);
#pragma unroll
    for (int v = 0; v < kGP; v++) a[0][v] = fma(Tv[v], c0, a[0][v]);
}

template <int GK, int T, int R, class CX>
struct GLeaf;

template <int GK, int T, int R, class CX, int I>
struct GLeafStep {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[T + 1][kGP], GRec &rec) {
        if constexpr (I <= ellc<GK>(R)) {
            constexpr int b = blockof<GK>(I);
            double Tv[kGP], dTv[kGP];
            if constexpr (I <= kTab + 1) {
                gtab_get<T, I - 2>(cx, Tv, dTv);
            } else {
                double x2[kGP], x2d[kGP];
                gtab_x2<T>(cx, x2, x2d);
                if constexpr (I == kTab + 2) { gtab_get<T, kTab - 1>(cx, rec.Tc, rec.dTc); gtab_get<T, kTab - 2>(cx, rec.Tq, rec.dTq); }
#pragma unroll
                for (int v = 0; v < kGP; v++) {
                    Tv[v] = fma(x2[v], rec.Tc[v], -rec.Tq[v]);
                    dTv[v] = fma(x2[v], rec.dTc[v], fma(x2d[v], rec.Tc[v], -rec.dTq[v]));
                    rec.dTq[v] = rec.dTc[v]; rec.dTc[v] = dTv[v]; rec.Tq[v] = rec.Tc[v]; rec.Tc[v] = Tv[v];
                }
            }
            if constexpr (R - b == 0) {
                const double c0 = th_get(th, off);
                off += 1;
                g_fold_theta<T>(a, c0, Tv, dTv);
            } else {
                double c[T][kGP];
                off = GLeaf<GK, T - 1, R - b, CX>::run(cx, th, off, c);
                g_fold<T>(a, c, Tv, dTv);
            }
            return GLeafStep<GK, T, R, CX, I + 1>::run(cx, th, off, a, rec);
        } else {
            return off;
        }
    }
};

template <int GK, int T, int R, class CX, int B>
struct GLeafBlocks {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[T + 1][kGP], GRec &rec) {
        if constexpr (B <= R) {
            constexpr int lo = ellc<GK>(B - 1) + 1, hi = ellc<GK>(B);
            constexpr int CS = R - B == 0 ? 1 : even_up(leaf_end<GK>(CX::LOOPD, T - 1, R - B > 0 ? R - B : 0, 0));
            const int start = R - B == 0 ? off : even_up(off);
            const double *blk = th.base + start;
            const double2 *tb = cx.stab + (size_t)(T - 1 - CX::TREG) * (kTab + 1) * kGP * CX::NT;
            double x2[kGP] = {}, x2d[kGP] = {};
            if constexpr (hi > kTab + 1) {
#pragma unroll
                for (int v = 0; v < kGP; v++) {
                    const double2 u = tb[(kTab * kGP + v) * CX::NT];
                    x2[v] = u.x; x2d[v] = u.y;
                    if constexpr (lo <= kTab + 2) {
                        const double2 e1 = tb[((kTab - 1) * kGP + v) * CX::NT], e0 = tb[((kTab - 2) * kGP + v) * CX::NT];
                        rec.Tc[v] = e1.x; rec.dTc[v] = e1.y; rec.Tq[v] = e0.x; rec.dTq[v] = e0.y;
                    }
                }
            }
#pragma unroll 1
            for (int i = lo; i <= hi; i++) {
                double Tv[kGP], dTv[kGP];
#pragma unroll
                for (int v = 0; v < kGP; v++) {
                    if (hi <= kTab + 1 || i <= kTab + 1) { const double2 u = tb[((i - 2) * kGP + v) * CX::NT]; Tv[v] = u.x; dTv[v] = u.y; }
                    else {
                        Tv[v] = fma(x2[v], rec.Tc[v], -rec.Tq[v]);
                        dTv[v] = fma(x2[v], rec.dTc[v], fma(x2d[v], rec.Tc[v], -rec.dTq[v]));
                        rec.dTq[v] = rec.dTc[v]; rec.dTc[v] = dTv[v]; rec.Tq[v] = rec.Tc[v]; rec.Tc[v] = Tv[v];
                    }
                }
                if constexpr (R - B == 0) {
                    g_fold_theta<T>(a, blk[i - lo], Tv, dTv);
                } else {
                    double c[T][kGP];
                    ThRd sub{blk + (i - lo) * CS, 0.0};
                    GLeaf<GK, T - 1, R - B, CX>::run(cx, sub, 0, c);
                    g_fold<T>(a, c, Tv, dTv);
                }
            }
            return GLeafBlocks<GK, T, R, CX, B + 1>::run(cx, th, even_up(start + (hi - lo + 1) * CS), a, rec);
        } else {
            return off;
        }
    }
};

template <int GK, int R, class CX>
struct GLeaf<GK, 0, R, CX> {
    static __device__ __forceinline__ int run(const CX &, ThRd &th, int off, double (&a)[1][kGP]) {
        const double c0 = th_get(th, off);
#pragma unroll
        for (int v = 0; v < kGP; v++) a[0][v] = c0;
        return off + 1;
    }
};

template <int GK, int T, int R, class CX>
struct GLeaf {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[T + 1][kGP]) {
        if constexpr (R == 0) {
            const double c0 = th_get(th, off);
#pragma unroll
            for (int v = 0; v < kGP; v++) { a[0][v] = c0;
#pragma unroll
                for (int m = 1; m <= T; m++) a[m][v] = 0.0; }
            return off + 1;
        } else {
            {
                double c[T][kGP];
                off = GLeaf<GK, T - 1, R, CX>::run(cx, th, off, c);
#pragma unroll
                for (int v = 0; v < kGP; v++) {
#pragma unroll
                    for (int m = 0; m < T; m++) a[m][v] = c[m][v];
                    a[T][v] = 0.0;
                }
            }
            GRec rec{};
            if constexpr (T < CX::LOOPD) return GLeafStep<GK, T, R, CX, 2>::run(cx, th, off, a, rec);
            else return GLeafBlocks<GK, T, R, CX, 1>::run(cx, th, off, a, rec);
        }
    }
};

template <int GK, int LL, int R, int TREG, int LOOPD, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) smolyak_gpts_kernel(const __grid_constant__ LeafParams q) {
    static_assert(LOOPD > TREG, "looped dimensions read their table from shared memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    double *s_theta = reinterpret_cast<double *>(smem_raw);
    double2 *stab = reinterpret_cast<double2 *>(smem_raw + (((size_t)q.Kslots * 8 + 15) & ~(size_t)15)) + tid;
    for (int64_t i = tid; i < q.K; i += NT) s_theta[q.thmap[i]] = q.theta[i * q.theta_stride];
    __syncthreads();
    using CX = GPtsCtx<TREG, LOOPD, NT>;
    for (int64_t p0 = (int64_t)blockIdx.x * NT * kGP; p0 < q.n; p0 += (int64_t)gridDim.x * NT * kGP) {
        CX cx;
        cx.stab = stab;
        int64_t pv[kGP];
#pragma unroll
        for (int v = 0; v < kGP; v++) pv[v] = min(p0 + tid + (int64_t)v * NT, q.n - 1);
#pragma unroll
        for (int j = 0; j < LL; j++) {
#pragma unroll
            for (int v = 0; v < kGP; v++) {
                double xj[2];
                skd::to_pm1_jet<2>(q.tr[j], __ldcs(q.x + pv[v] * LL + j), xj);
                const double x0 = xj[0], x1 = xj[1], x2 = 2.0 * x0, x2d = 2.0 * x1;
                double t[kTab], dt[kTab];
                t[0] = x0; dt[0] = x1;
                t[1] = fma(x2, x0, -1.0); dt[1] = x2 * x2d;
#pragma unroll
                for (int i = 2; i < kTab; i++) { t[i] = fma(x2, t[i - 1], -t[i - 2]); dt[i] = fma(x2, dt[i - 1], fma(x2d, t[i - 1], -dt[i - 2])); }
                if (j < TREG) {
                    cx.x2[j < TREG ? j : 0][v] = x2; cx.x2d[j < TREG ? j : 0][v] = x2d;
#pragma unroll
                    for (int i = 0; i < kTab; i++) { cx.t[j < TREG ? j : 0][i][v] = t[i]; cx.dt[j < TREG ? j : 0][i][v] = dt[i]; }
                } else {
                    double2 *sd = stab + (size_t)(j - TREG) * (kTab + 1) * kGP * NT;
#pragma unroll
                    for (int i = 0; i < kTab; i++) sd[(i * kGP + v) * NT] = make_double2(t[i], dt[i]);
                    sd[(kTab * kGP + v) * NT] = make_double2(x2, x2d);
                }
            }
        }
        double a[LL + 1][kGP];
        ThRd th{s_theta, 0.0};
        GLeaf<GK, LL, R, CX>::run(cx, th, 0, a);
#pragma unroll
        for (int v = 0; v < kGP; v++) {
            const int64_t p = p0 + tid + (int64_t)v * NT;
            if (p < q.n) {
                double *o = q.out + p * q.out_stride;
#pragma unroll
                for (int m = 0; m <= LL; m++) __stcs(o + m, a[m][v]);
            }
        }
    }
}

template <int GK, int LL, int R, int TREG, int LOOPD, int NT, int MINB>
cudaError_t launch_gpts(const LeafParams &q, cudaStream_t stream) {
    auto kern = smolyak_gpts_kernel<GK, LL, R, TREG, LOOPD, NT, MINB>;
    const size_t smem = (size_t)(LL - TREG) * (kTab + 1) * kGP * NT * 16 + (size_t)q.Kslots * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1, dev = 0, sms = 148;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (q.n + (int64_t)NT * kGP - 1) / ((int64_t)NT * kGP);
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * std::max(per_sm, 1));
    kern<<<grid, NT, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

}  // namespace skk
