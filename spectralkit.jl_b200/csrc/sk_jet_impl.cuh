// sm_100a kernels, part 3b — second-order jets on the unrolled leaves: Smolyak linear_combination for ∂ specifications
// whose components all have total order <= 2 (Hessian entries, Hessian diagonals, single cross partials such as ∂(1,1):
// src/derivatives.jl:336-421, test/test_smolyak.jl:39, test/test_derivatives.jl:107-112).
//
// Same nested evaluation and the same straight-line template recursion over the block structure as sk_leaf_impl.cuh
// (whose coefficient layout, θ pair reads and per-block runtime loops are reused), but level T carries a second-order jet
// in the coordinates 1..T instead of (value, gradient):
//     MODE 2: value, ∂_m, ∂²_m                (2T + 1 components)   — requests without cross partials
//     MODE 3: value, ∂_m, ∂_m∂_n (m <= n)     ((T + 1)(T + 2)/2)    — any request of total order <= 2
// and the kernel stores the requested components only (column map in the parameters).  One fold of a child c into its
// parent a at index i of dimension T, with (T_i, T_i', T_i'') of that coordinate:
//     a[pm(j)] += T_i·c[j] for every child component; a[∂_T] += T_i'·c[0]; a[∂_m∂_T] += T_i'·c[∂_m] (MODE 3);
//     a[∂²_T] += T_i''·c[0]
// Index 1 (T_0 = 1) is a rename, a child without budget is the single coefficient θ (three fmas).  The basis must be one
// leaf (d <= 6, block cap M = B); everything else stays with smolyak_general_kernel.
// Fast mode only (fma, nested summation order): within the stated tolerance of the reference; exact mode uses the direct
// kernel.
#pragma once
#include "sk_leaf_impl.cuh"

namespace skk {

template <int MODE, int T>
struct JNC { static constexpr int value = T == 0 ? 1 : MODE == 2 ? 2 * T + 1 : (T + 1) * (T + 2) / 2; };
// component index of a child (level T-1) component j in its parent (level T): the first partials shift the second ones by one
template <int T>
__host__ __device__ constexpr int jet_pm(int j) { return j <= T - 1 ? j : j + 1; }
// second partial ∂_m∂_n (1 <= m <= n <= T) at level T, MODE 3
__host__ __device__ constexpr int jet_sec3(int T, int m, int n) { return T + n * (n - 1) / 2 + m; }

struct JetTab {
    double t[kTab], dt[kTab], ddt[kTab];
    double x2, x2d, x2dd;
};

template <int MODE_, int TREG_, int LOOPD_, int NT_>
struct JetCtx {
    static constexpr int MODE = MODE_, TREG = TREG_, LOOPD = LOOPD_, NT = NT_;
    JetTab tabs[TREG_ > 0 ? TREG_ : 1];
    const double2 *stab2;       // shared tables of the dimensions >= TREG: (T, T') per entry, offset by the thread id
    const double *stab1;        // T''
};

template <int T, int E, class CX>
__device__ __forceinline__ void jtab_get(const CX &cx, double &Tv, double &dTv, double &ddTv) {
    if constexpr (T - 1 < CX::TREG) {
        Tv = cx.tabs[T - 1].t[E]; dTv = cx.tabs[T - 1].dt[E]; ddTv = cx.tabs[T - 1].ddt[E];
    } else {
        const double2 v = cx.stab2[((T - 1 - CX::TREG) * (kTab + 1) + E) * CX::NT];
        Tv = v.x; dTv = v.y;
        ddTv = cx.stab1[((T - 1 - CX::TREG) * (kTab + 1) + E) * CX::NT];
    }
}
template <int T, class CX>
__device__ __forceinline__ void jtab_x2(const CX &cx, double &x2, double &x2d, double &x2dd) {
    if constexpr (T - 1 < CX::TREG) { x2 = cx.tabs[T - 1].x2; x2d = cx.tabs[T - 1].x2d; x2dd = cx.tabs[T - 1].x2dd; }
    else {
        const double2 v = cx.stab2[((T - 1 - CX::TREG) * (kTab + 1) + kTab) * CX::NT];
        x2 = v.x; x2d = v.y;
        x2dd = cx.stab1[((T - 1 - CX::TREG) * (kTab + 1) + kTab) * CX::NT];
    }
}

// three-term recurrence on (T, T', T'') with x2 = 2x, x2d = 2x', x2dd = 2x'' (raw derivatives in the caller's coordinate):
// T_{n+1}'' = 2x·T_n'' + 4x'·T_n' + 2x''·T_n − T_{n−1}''   (chebyshev.jl:65-70 through derivatives.jl:125-134)
struct JetRec { double Tc, Tq, dTc, dTq, ddTc, ddTq; };
__device__ __forceinline__ void jet_rec_step(JetRec &r, double x2, double x2d, double x2dd, double &Tv, double &dTv, double &ddTv) {
    Tv = fma(x2, r.Tc, -r.Tq);
    dTv = fma(x2, r.dTc, fma(x2d, r.Tc, -r.dTq));
    ddTv = fma(x2, r.ddTc, fma(2.0 * x2d, r.dTc, fma(x2dd, r.Tc, -r.ddTq)));
    r.Tq = r.Tc; r.Tc = Tv; r.dTq = r.dTc; r.dTc = dTv; r.ddTq = r.ddTc; r.ddTc = ddTv;
}

// fold a child jet (level T-1) into its parent at one index of dimension T
template <int MODE, int T>
__device__ __forceinline__ void jet_fold(double (&a)[JNC<MODE, T>::value], const double (&c)[JNC<MODE, T - 1>::value], double Tv, double dTv,
                                         double ddTv) {
    constexpr int NCC = JNC<MODE, T - 1>::value, NCP = JNC<MODE, T>::value;
    a[NCP - 1] = fma(ddTv, c[0], a[NCP - 1]);
    a[T] = fma(dTv, c[0], a[T]);
    if constexpr (MODE == 3) {
#pragma unroll
        for (int m = 1; m < T; m++) a[jet_sec3(T, m, T)] = fma(dTv, c[m < NCC ? m : 0], a[jet_sec3(T, m, T)]);
    }
#pragma unroll
    for (int j = 0; j < NCC; j++) a[jet_pm<T>(j)] = fma(Tv, c[j], a[jet_pm<T>(j)]);
}
template <int MODE, int T>
__device__ __forceinline__ void jet_fold_theta(double (&a)[JNC<MODE, T>::value], double c0, double Tv, double dTv, double ddTv) {
    constexpr int NCP = JNC<MODE, T>::value;
    a[NCP - 1] = fma(ddTv, c0, a[NCP - 1]);
    a[T] = fma(dTv, c0, a[T]);
    a[0] = fma(Tv, c0, a[0]);
}

template <int GK, int MODE, int T, int R, class CX>
struct JLeaf;

// indices I..ℓ(R) of dimension T, unrolled
template <int GK, int MODE, int T, int R, class CX, int I>
struct JLeafStep {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[JNC<MODE, T>::value], JetRec &rec) {
        if constexpr (I <= ellc<GK>(R)) {
            constexpr int b = blockof<GK>(I);
            double Tv, dTv, ddTv;
            if constexpr (I <= kTab + 1) {
                jtab_get<T, I - 2>(cx, Tv, dTv, ddTv);
            } else {
                double x2, x2d, x2dd;
                jtab_x2<T>(cx, x2, x2d, x2dd);
                if constexpr (I == kTab + 2) {
                    jtab_get<T, kTab - 1>(cx, rec.Tc, rec.dTc, rec.ddTc);
                    jtab_get<T, kTab - 2>(cx, rec.Tq, rec.dTq, rec.ddTq);
                }
                jet_rec_step(rec, x2, x2d, x2dd, Tv, dTv, ddTv);
            }
            if constexpr (R - b == 0) {
                const double c0 = th_get(th, off);
                off += 1;
                jet_fold_theta<MODE, T>(a, c0, Tv, dTv, ddTv);
            } else {
                double c[JNC<MODE, T - 1>::value];
                off = JLeaf<GK, MODE, T - 1, R - b, CX>::run(cx, th, off, c);
                jet_fold<MODE, T>(a, c, Tv, dTv, ddTv);
            }
            return JLeafStep<GK, MODE, T, R, CX, I + 1>::run(cx, th, off, a, rec);
        } else {
            return off;
        }
    }
};

// dimensions >= LOOPD (table in shared memory): one runtime loop per block, the child code exists once per block
template <int GK, int MODE, int T, int R, class CX, int B>
struct JLeafBlocks {
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[JNC<MODE, T>::value], JetRec &rec) {
        if constexpr (B <= R) {
            constexpr int lo = ellc<GK>(B - 1) + 1, hi = ellc<GK>(B);
            constexpr int CS = R - B == 0 ? 1 : even_up(leaf_end<GK>(CX::LOOPD, T - 1, R - B > 0 ? R - B : 0, 0));
            const int start = R - B == 0 ? off : even_up(off);
            const double *blk = th.base + start;
            const double2 *tb2 = cx.stab2 + (size_t)(T - 1 - CX::TREG) * (kTab + 1) * CX::NT;
            const double *tb1 = cx.stab1 + (size_t)(T - 1 - CX::TREG) * (kTab + 1) * CX::NT;
            double x2 = 0, x2d = 0, x2dd = 0;
            if constexpr (hi > kTab + 1) {
                const double2 v = tb2[kTab * CX::NT];
                x2 = v.x; x2d = v.y; x2dd = tb1[kTab * CX::NT];
                if constexpr (lo <= kTab + 2) {
                    const double2 e1 = tb2[(kTab - 1) * CX::NT], e0 = tb2[(kTab - 2) * CX::NT];
                    rec.Tc = e1.x; rec.dTc = e1.y; rec.ddTc = tb1[(kTab - 1) * CX::NT];
                    rec.Tq = e0.x; rec.dTq = e0.y; rec.ddTq = tb1[(kTab - 2) * CX::NT];
                }
            }
#pragma unroll 1
            for (int i = lo; i <= hi; i++) {
                double Tv, dTv, ddTv;
                if (hi <= kTab + 1 || i <= kTab + 1) {
                    const double2 v = tb2[(i - 2) * CX::NT];
                    Tv = v.x; dTv = v.y; ddTv = tb1[(i - 2) * CX::NT];
                } else {
                    jet_rec_step(rec, x2, x2d, x2dd, Tv, dTv, ddTv);
                }
                if constexpr (R - B == 0) {
                    jet_fold_theta<MODE, T>(a, blk[i - lo], Tv, dTv, ddTv);
                } else {
                    double c[JNC<MODE, T - 1>::value];
                    ThRd sub{blk + (i - lo) * CS, 0.0};
                    JLeaf<GK, MODE, T - 1, R - B, CX>::run(cx, sub, 0, c);
                    jet_fold<MODE, T>(a, c, Tv, dTv, ddTv);
                }
            }
            return JLeafBlocks<GK, MODE, T, R, CX, B + 1>::run(cx, th, even_up(start + (hi - lo + 1) * CS), a, rec);
        } else {
            return off;
        }
    }
};

template <int GK, int MODE, int R, class CX>
struct JLeaf<GK, MODE, 0, R, CX> {
    static __device__ __forceinline__ int run(const CX &, ThRd &th, int off, double (&a)[1]) {
        a[0] = th_get(th, off);
        return off + 1;
    }
};

template <int GK, int MODE, int T, int R, class CX>
struct JLeaf {
    static constexpr int NCP = JNC<MODE, T>::value;
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[NCP]) {
        if constexpr (R == 0) {
            a[0] = th_get(th, off);
#pragma unroll
            for (int m = 1; m < NCP; m++) a[m] = 0.0;
            return off + 1;
        } else {
            {   // index 1: T_0 = (1, 0, 0)
                double c[JNC<MODE, T - 1>::value];
                off = JLeaf<GK, MODE, T - 1, R, CX>::run(cx, th, off, c);
#pragma unroll
                for (int m = 0; m < NCP; m++) a[m] = 0.0;
#pragma unroll
                for (int j = 0; j < JNC<MODE, T - 1>::value; j++) a[jet_pm<T>(j)] = c[j];
            }
            JetRec rec{0, 0, 0, 0, 0, 0};
            if constexpr (T < CX::LOOPD) return JLeafStep<GK, MODE, T, R, CX, 2>::run(cx, th, off, a, rec);
            else return JLeafBlocks<GK, MODE, T, R, CX, 1>::run(cx, th, off, a, rec);
        }
    }
};

// ---- kernel: the basis is one leaf of LL = d dimensions and budget R
struct JetParams {
    const double *x;
    double *out;
    const double *theta;
    const uint32_t *thmap;        // shared-memory slot of every coefficient (layout of leaf_end with this kernel's LOOPD)
    int64_t n, K, Kslots;
    int ncomp;                    // requested components = doubles per output element
    int8_t col[32];               // jet component -> output column, -1 = not requested
    sk_transform tr[6];
};

constexpr size_t jet_smem_bytes(int NT, int LL, int TREG, int64_t Kslots) {
    return (size_t)(LL > TREG ? LL - TREG : 0) * (kTab + 1) * NT * 24 + (size_t)Kslots * 8;
}

template <int GK, int LL, int R, int MODE, int TREG, int LOOPD, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) smolyak_jet_kernel(const __grid_constant__ JetParams q) {
    static_assert(LOOPD > TREG, "looped dimensions read their table from shared memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    constexpr int SD = LL > TREG ? LL - TREG : 0;
    double2 *stab2 = reinterpret_cast<double2 *>(smem_raw) + tid;
    double *stab1 = reinterpret_cast<double *>(smem_raw + (size_t)SD * (kTab + 1) * NT * 16) + tid;
    double *s_theta = reinterpret_cast<double *>(smem_raw + (size_t)SD * (kTab + 1) * NT * 24);
    for (int64_t i = tid; i < q.K; i += NT) s_theta[q.thmap[i]] = q.theta[i];
    __syncthreads();
    constexpr int NCL = JNC<MODE, LL>::value;
    using CX = JetCtx<MODE, TREG, LOOPD, NT>;
    for (int64_t p0 = (int64_t)blockIdx.x * NT; p0 < q.n; p0 += (int64_t)gridDim.x * NT) {
        const bool live = p0 + tid < q.n;
        const int64_t p = live ? p0 + tid : q.n - 1;
        const double *xp = q.x + p * LL;
        CX cx;
        cx.stab2 = stab2; cx.stab1 = stab1;
#pragma unroll
        for (int j = 0; j < LL; j++) {
            double xj[3];
            skd::to_pm1_jet<3>(q.tr[j], __ldcs(xp + j), xj);
            double t[kTab], dt[kTab], ddt[kTab];
            const double x2 = 2.0 * xj[0], x2d = 2.0 * xj[1], x2dd = 2.0 * xj[2];
            t[0] = xj[0]; dt[0] = xj[1]; ddt[0] = xj[2];
            JetRec rec{xj[0], 1.0, xj[1], 0.0, xj[2], 0.0};
#pragma unroll
            for (int i = 1; i < kTab; i++) jet_rec_step(rec, x2, x2d, x2dd, t[i], dt[i], ddt[i]);
            if (j < TREG) {
                JetTab &tb = cx.tabs[j < TREG ? j : 0];
                tb.x2 = x2; tb.x2d = x2d; tb.x2dd = x2dd;
#pragma unroll
                for (int i = 0; i < kTab; i++) { tb.t[i] = t[i]; tb.dt[i] = dt[i]; tb.ddt[i] = ddt[i]; }
            } else {
                double2 *s2 = stab2 + (size_t)(j - TREG) * (kTab + 1) * NT;
                double *s1 = stab1 + (size_t)(j - TREG) * (kTab + 1) * NT;
#pragma unroll
                for (int i = 0; i < kTab; i++) { s2[i * NT] = make_double2(t[i], dt[i]); s1[i * NT] = ddt[i]; }
                s2[kTab * NT] = make_double2(x2, x2d); s1[kTab * NT] = x2dd;
            }
        }
        double a[NCL];
        ThRd th{s_theta, 0.0};
        JLeaf<GK, MODE, LL, R, CX>::run(cx, th, 0, a);
        if (live) {
            double *o = q.out + p * q.ncomp;
#pragma unroll
            for (int m = 0; m < NCL; m++) if (q.col[m] >= 0) __stcs(o + q.col[m], a[m]);
        }
    }
}

template <int GK, int LL, int R, int MODE, int TREG, int LOOPD, int NT, int MINB>
cudaError_t launch_jet(const JetParams &q, cudaStream_t stream) {
    auto kern = smolyak_jet_kernel<GK, LL, R, MODE, TREG, LOOPD, NT, MINB>;
    const size_t smem = jet_smem_bytes(NT, LL, TREG, q.Kslots);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1, dev = 0, sms = 148;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (q.n + NT - 1) / NT;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sms * std::max(per_sm, 1));
    kern<<<grid, NT, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

// ---- which (family, depth, budget) are built, and their shapes (host and device agree through these)
// budgets: up to 4 (2 561 terms at d = 6 for InteriorGrid2, 4 361 for InteriorGrid), 5 for EndpointGrid / InteriorGrid2 in
// d <= 4; larger bases keep the general kernel (their θ does not fit next to the shared tables)
constexpr bool jet_built_c(int gk, int LL, int R) {
    if (LL < 1 || LL > 6 || R < 1) return false;
    return gk == SK_GRID_INTERIOR ? R <= 4 : R <= (LL <= 4 ? 5 : 4);
}
constexpr size_t kJetSmemMax = 220 * 1024;      // the kernel has no static shared memory: 227 KB per CTA are available
constexpr int jet_treg_c(int LL, int) { return LL <= 2 ? LL : 2; }      // 42 registers per table; 256 threads × four shared tables fit
constexpr int jet_loopd_c(int gk, int LL, int R, int MODE) {
    const long n = gk == SK_GRID_INTERIOR ? nterms<SK_GRID_INTERIOR>(LL, R) : gk == SK_GRID_ENDPOINT ? nterms<SK_GRID_ENDPOINT>(LL, R) : nterms<SK_GRID_INTERIOR2>(LL, R);
#ifdef SK_JET_LOOPD6
    if (LL == 6) return SK_JET_LOOPD6;                      // experiments
#endif
    if (LL <= 2) return 99;                                 // both tables in registers: nothing to loop over
    if (n * (MODE == 3 ? 3 : 2) <= 1200) return 99;         // small leaves: straight-line code
    // larger ones loop the top dimension; the diagonal jet of the largest leaves also the one below it (measured at
    // (InteriorGrid2, 6, 4), first looped dimension 4 / 5 / 6: diagonal jet 804 / 954 / 775 Mpoints/s, full jet 545 / 613 / 668)
    const int want = (MODE == 2 && n > 2000) ? LL - 1 : LL;
    return want > jet_treg_c(LL, MODE) ? want : jet_treg_c(LL, MODE) + 1;
}
constexpr int jet_nt_c(int LL, int) { return LL >= 5 ? 256 : 256; }
constexpr int jet_minb_c(int LL, int R, int MODE) { return (LL <= 3 && R <= 3 && MODE == 2) ? 2 : 1; }

template <int GK, int MODE, int LL, int R>
cudaError_t launch_jet_built(const JetParams &q, cudaStream_t stream) {
    if constexpr (jet_built_c(GK, LL, R))
        return launch_jet<GK, LL, R, MODE, jet_treg_c(LL, MODE), jet_loopd_c(GK, LL, R, MODE), jet_nt_c(LL, MODE), jet_minb_c(LL, R, MODE)>(q, stream);
    else
        return cudaErrorNotSupported;
}
template <int GK, int MODE, int LL>
cudaError_t dispatch_jet_r(int R, const JetParams &q, cudaStream_t stream) {
    switch (R) {
    case 1: return launch_jet_built<GK, MODE, LL, 1>(q, stream);
    case 2: return launch_jet_built<GK, MODE, LL, 2>(q, stream);
    case 3: return launch_jet_built<GK, MODE, LL, 3>(q, stream);
    case 4: return launch_jet_built<GK, MODE, LL, 4>(q, stream);
    case 5: return launch_jet_built<GK, MODE, LL, 5>(q, stream);
    }
    return cudaErrorNotSupported;
}
template <int GK, int MODE>
cudaError_t dispatch_jet(int LL, int R, const JetParams &q, cudaStream_t stream) {
    switch (LL) {
    case 1: return dispatch_jet_r<GK, MODE, 1>(R, q, stream);
    case 2: return dispatch_jet_r<GK, MODE, 2>(R, q, stream);
    case 3: return dispatch_jet_r<GK, MODE, 3>(R, q, stream);
    case 4: return dispatch_jet_r<GK, MODE, 4>(R, q, stream);
    case 5: return dispatch_jet_r<GK, MODE, 5>(R, q, stream);
    case 6: return dispatch_jet_r<GK, MODE, 6>(R, q, stream);
    }
    return cudaErrorNotSupported;
}

}  // namespace skk
