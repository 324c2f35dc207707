// sm_100a kernels, part 3c — a few coefficient vectors at once on the unrolled leaves: linear_combination of a Smolyak basis
// with SVector coefficients (values only; src/generic_api.jl:114-128 with eltype(θ) <: SVector, test/test_generic_api.jl:95-102;
// the stacked policy functions of src/experimental.jl:156-166) for V = 2 or 4 vectors per launch.
//
// Why a kernel of its own: the caller used to launch the one-vector leaf kernel once per vector below the DMMA kernel's
// crossover (sk_block.cu).  With V coefficient vectors per launch the transformation, the tables, the recurrences and the
// table reads are paid once for V results, and the V coefficients of a term arrive in one or two LDS.128.  Measured at cfg 4a:
// +23 % (two vectors), +25 % (four) over per-vector launches.  What it does NOT deliver: in source order the Chebyshev value T
// is the same operand of V consecutive fmas a[v] += T·θ_v, which would keep it in the operand-reuse cache (two fresh register
// operands per fma instead of three: the register-file law of profiles/r2_leaf_register_file.md) — ptxas interleaves the
// chains and only 22 % of the DFMAs carry a .reuse flag in the SASS, as in the one-vector kernel.
//
// Same template recursion over the block structure, same coefficient layout (leaf_end) and per-block runtime loops as
// sk_leaf_impl.cuh; a slot holds the V coefficients of one basis term (one or two LDS.128).  The basis must be one leaf
// (d <= 6, block cap M = B).  Fast mode only (fma, nested summation order): within the stated tolerance of the reference.
#pragma once
#include "sk_leaf_impl.cuh"

namespace skk {

template <int TREG_, int LOOPD_, int NT_>
struct VecCtx {
    static constexpr int TREG = TREG_, LOOPD = LOOPD_, NT = NT_;
    double t[TREG_ > 0 ? TREG_ : 1][kTab];      // T_2..T_7 of the dimensions with register tables
    double x2[TREG_ > 0 ? TREG_ : 1];
    const double *stab;                         // shared tables of the dimensions >= TREG, offset by the thread id: [dim][entry | 2x][thread]
};

template <int V>
__device__ __forceinline__ void vth_get(const double *base, int off, double (&c)[V]) {
    static_assert(V == 2 || V == 4, "two or four coefficient vectors per launch");
    const double2 *p = reinterpret_cast<const double2 *>(base + (size_t)off * V);
    const double2 u = p[0];
    c[0] = u.x; c[1] = u.y;
    if constexpr (V == 4) { const double2 w = p[1]; c[2] = w.x; c[3] = w.y; }
}

template <int T, int E, class CX>
__device__ __forceinline__ double vtab_get(const CX &cx) {
    if constexpr (T - 1 < CX::TREG) return cx.t[T - 1][E];
    else return cx.stab[((T - 1 - CX::TREG) * (kTab + 1) + E) * CX::NT];
}
template <int T, class CX>
__device__ __forceinline__ double vtab_x2(const CX &cx) {
    if constexpr (T - 1 < CX::TREG) return cx.x2[T - 1];
    else return cx.stab[((T - 1 - CX::TREG) * (kTab + 1) + kTab) * CX::NT];
}

template <int GK, int V, int T, int R, class CX>
struct VLeaf;

template <int GK, int V, int T, int R, class CX, int I>
struct VLeafStep {
    static __device__ __forceinline__ int run(const CX &cx, const double *th, int off, double (&a)[V], double &Tc, double &Tq) {
        if constexpr (I <= ellc<GK>(R)) {
            constexpr int b = blockof<GK>(I);
            double Tv;
            if constexpr (I <= kTab + 1) {
                Tv = vtab_get<T, I - 2>(cx);
            } else {
                if constexpr (I == kTab + 2) { Tc = vtab_get<T, kTab - 1>(cx); Tq = vtab_get<T, kTab - 2>(cx); }
                Tv = fma(vtab_x2<T>(cx), Tc, -Tq);
                Tq = Tc; Tc = Tv;
            }
            double c[V];
            if constexpr (R - b == 0) { vth_get<V>(th, off, c); off += 1; }
            else off = VLeaf<GK, V, T - 1, R - b, CX>::run(cx, th, off, c);
#pragma unroll
            for (int v = 0; v < V; v++) a[v] = fma(Tv, c[v], a[v]);      // Tv: the same operand V times in a row
            return VLeafStep<GK, V, T, R, CX, I + 1>::run(cx, th, off, a, Tc, Tq);
        } else {
            return off;
        }
    }
};

template <int GK, int V, int T, int R, class CX, int B>
struct VLeafBlocks {
    static __device__ __forceinline__ int run(const CX &cx, const double *th, int off, double (&a)[V], double &Tc, double &Tq) {
        if constexpr (B <= R) {
            constexpr int lo = ellc<GK>(B - 1) + 1, hi = ellc<GK>(B);
            constexpr int CS = R - B == 0 ? 1 : even_up(leaf_end<GK>(CX::LOOPD, T - 1, R - B > 0 ? R - B : 0, 0));
            const int start = R - B == 0 ? off : even_up(off);
            const double *tb = cx.stab + (size_t)(T - 1 - CX::TREG) * (kTab + 1) * CX::NT;
            double x2 = 0;
            if constexpr (hi > kTab + 1) {
                x2 = tb[kTab * CX::NT];
                if constexpr (lo <= kTab + 2) { Tc = tb[(kTab - 1) * CX::NT]; Tq = tb[(kTab - 2) * CX::NT]; }
            }
#pragma unroll 1
            for (int i = lo; i <= hi; i++) {
                double Tv;
                if (hi <= kTab + 1 || i <= kTab + 1) Tv = tb[(i - 2) * CX::NT];
                else { Tv = fma(x2, Tc, -Tq); Tq = Tc; Tc = Tv; }
                double c[V];
                if constexpr (R - B == 0) vth_get<V>(th, start + (i - lo), c);
                else VLeaf<GK, V, T - 1, R - B, CX>::run(cx, th + (size_t)(start + (i - lo) * CS) * V, 0, c);
#pragma unroll
                for (int v = 0; v < V; v++) a[v] = fma(Tv, c[v], a[v]);
            }
            return VLeafBlocks<GK, V, T, R, CX, B + 1>::run(cx, th, even_up(start + (hi - lo + 1) * CS), a, Tc, Tq);
        } else {
            return off;
        }
    }
};

template <int GK, int V, int R, class CX>
struct VLeaf<GK, V, 0, R, CX> {
    static __device__ __forceinline__ int run(const CX &, const double *th, int off, double (&a)[V]) {
        vth_get<V>(th, off, a);
        return off + 1;
    }
};

template <int GK, int V, int T, int R, class CX>
struct VLeaf {
    static __device__ __forceinline__ int run(const CX &cx, const double *th, int off, double (&a)[V]) {
        if constexpr (R == 0) {
            vth_get<V>(th, off, a);
            return off + 1;
        } else {
            off = VLeaf<GK, V, T - 1, R, CX>::run(cx, th, off, a);      // index 1: T_0 = 1
            double Tc = 0, Tq = 0;
            if constexpr (T < CX::LOOPD) return VLeafStep<GK, V, T, R, CX, 2>::run(cx, th, off, a, Tc, Tq);
            else return VLeafBlocks<GK, V, T, R, CX, 1>::run(cx, th, off, a, Tc, Tq);
        }
    }
};

struct VecParams {
    const double *x;
    double *out;                  // out[p·out_stride + v]
    const double *theta;          // θ[k·theta_stride + v]
    const uint32_t *thmap;        // slot of every coefficient (layout of leaf_end with this kernel's LOOPD)
    int64_t n, K, Kslots, theta_stride, out_stride;
    sk_transform tr[6];
};

constexpr size_t vec_smem_bytes(int NT, int LL, int TREG, int V, int64_t Kslots) {
    return (size_t)(LL > TREG ? LL - TREG : 0) * (kTab + 1) * NT * 8 + (size_t)Kslots * V * 8;
}

template <int GK, int LL, int R, int V, int TREG, int LOOPD, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) smolyak_vec_kernel(const __grid_constant__ VecParams q) {
    static_assert(LOOPD > TREG, "looped dimensions read their table from shared memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    double *s_theta = reinterpret_cast<double *>(smem_raw);                             // first: 16-byte aligned slots
    double *stab = reinterpret_cast<double *>(smem_raw + (((size_t)q.Kslots * V * 8 + 15) & ~(size_t)15)) + tid;
    for (int64_t i = tid; i < q.K * V; i += NT) {
        const int64_t k = i / V;
        const int v = (int)(i - k * V);
        s_theta[(size_t)q.thmap[k] * V + v] = q.theta[k * q.theta_stride + v];
    }
    __syncthreads();
    using CX = VecCtx<TREG, LOOPD, NT>;
    for (int64_t p0 = (int64_t)blockIdx.x * NT; p0 < q.n; p0 += (int64_t)gridDim.x * NT) {
        const bool live = p0 + tid < q.n;
        const int64_t p = live ? p0 + tid : q.n - 1;
        const double *xp = q.x + p * LL;
        CX cx;
        cx.stab = stab;
#pragma unroll
        for (int j = 0; j < LL; j++) {
            const double x0 = skd::to_pm1(q.tr[j], __ldcs(xp + j));
            const double x2 = 2.0 * x0;
            double t[kTab];
            t[0] = x0;
            t[1] = fma(x2, x0, -1.0);
#pragma unroll
            for (int i = 2; i < kTab; i++) t[i] = fma(x2, t[i - 1], -t[i - 2]);
            if (j < TREG) {
                cx.x2[j < TREG ? j : 0] = x2;
#pragma unroll
                for (int i = 0; i < kTab; i++) cx.t[j < TREG ? j : 0][i] = t[i];
            } else {
                double *sd = stab + (size_t)(j - TREG) * (kTab + 1) * NT;
#pragma unroll
                for (int i = 0; i < kTab; i++) sd[i * NT] = t[i];
                sd[kTab * NT] = x2;
            }
        }
        double a[V];
        VLeaf<GK, V, LL, R, CX>::run(cx, s_theta, 0, a);
        if (live) {
            double *o = q.out + p * q.out_stride;
#pragma unroll
            for (int v = 0; v < V; v++) __stcs(o + v, a[v]);
        }
    }
}

template <int GK, int LL, int R, int V, int TREG, int LOOPD, int NT, int MINB>
cudaError_t launch_vec(const VecParams &q, cudaStream_t stream) {
    auto kern = smolyak_vec_kernel<GK, LL, R, V, TREG, LOOPD, NT, MINB>;
    const size_t smem = vec_smem_bytes(NT, LL, TREG, V, q.Kslots) + 16;
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
constexpr bool vec_built_c(int gk, int LL, int R) {
    if (LL < 1 || LL > 6 || R < 1) return false;
    return gk == SK_GRID_INTERIOR ? R <= (LL <= 4 ? 4 : 3) : R <= (LL <= 4 ? 5 : 4);
}
constexpr int vec_treg_c(int LL) { return LL < 3 ? LL : 3; }
constexpr long vec_nterms_c(int gk, int LL, int R) {
    return gk == SK_GRID_INTERIOR ? nterms<SK_GRID_INTERIOR>(LL, R) : gk == SK_GRID_ENDPOINT ? nterms<SK_GRID_ENDPOINT>(LL, R) : nterms<SK_GRID_INTERIOR2>(LL, R);
}
// straight-line code up to about 3 000 fmas; larger leaves loop the top dimension (its table is in shared memory: LL > TREG)
constexpr int vec_loopd_c(int gk, int LL, int R, int V) { return (LL > 3 && vec_nterms_c(gk, LL, R) * V > 3000) ? LL : 99; }
constexpr int vec_nt_c(int LL, int V = 2) { return LL >= 5 && V == 4 ? 384 : 256; }      // measured at (InteriorGrid2, 6, 4): 756 -> 826 Mpoints/s for V = 4, 1 613 -> 1 534 for V = 2
constexpr int vec_minb_c(int gk, int LL, int R, int V) { return vec_nterms_c(gk, LL, R) * V <= 1500 ? 2 : 1; }

template <int GK, int V, int LL, int R>
cudaError_t launch_vec_built(const VecParams &q, cudaStream_t stream) {
    if constexpr (vec_built_c(GK, LL, R))
        return launch_vec<GK, LL, R, V, vec_treg_c(LL), vec_loopd_c(GK, LL, R, V), vec_nt_c(LL, V), vec_minb_c(GK, LL, R, V)>(q, stream);
    else
        return cudaErrorNotSupported;
}
template <int GK, int V, int LL>
cudaError_t dispatch_vec_r(int R, const VecParams &q, cudaStream_t stream) {
    switch (R) {
    case 1: return launch_vec_built<GK, V, LL, 1>(q, stream);
    case 2: return launch_vec_built<GK, V, LL, 2>(q, stream);
    case 3: return launch_vec_built<GK, V, LL, 3>(q, stream);
    case 4: return launch_vec_built<GK, V, LL, 4>(q, stream);
    case 5: return launch_vec_built<GK, V, LL, 5>(q, stream);
    }
    return cudaErrorNotSupported;
}
template <int GK, int V>
cudaError_t dispatch_vec(int LL, int R, const VecParams &q, cudaStream_t stream) {
    switch (LL) {
    case 1: return dispatch_vec_r<GK, V, 1>(R, q, stream);
    case 2: return dispatch_vec_r<GK, V, 2>(R, q, stream);
    case 3: return dispatch_vec_r<GK, V, 3>(R, q, stream);
    case 4: return dispatch_vec_r<GK, V, 4>(R, q, stream);
    case 5: return dispatch_vec_r<GK, V, 5>(R, q, stream);
    case 6: return dispatch_vec_r<GK, V, 6>(R, q, stream);
    }
    return cudaErrorNotSupported;
}

}  // namespace skk
