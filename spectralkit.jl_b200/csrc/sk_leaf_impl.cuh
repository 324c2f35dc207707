// sm_100a kernels, part 3 — K2 (headline): Smolyak linear_combination with compile-time unrolled
// leaves.  Same nested (Horner-like) evaluation as sk_nested.cu, restructured so that the FP64
// pipe, not instruction issue, is the limit:
//
//   * The lowest LL (<= 6) dimensions of the index set are a "leaf": for a remaining block budget
//     r, the set {(i_1..i_LL) : Σ b_j <= r} is FIXED, so Leaf<GK, LL, r> is straight-line code
//     produced by template recursion over the block structure of the nested family GK
//     (smolyak_traversal.jl:69-84,112-116,146-148).  No index table, no state machine, no branch:
//     every fma has register operands and a θ read at a compile-time offset from the unit's base.
//     For d <= 6 the whole evaluation of a point is one straight-line leaf.
//   * Chebyshev values/derivatives for indices 2..7 (degrees 1..6: blocks 0-2 of all three
//     families) are per-point tables: in REGISTERS for the lowest TREG dimensions (where almost all
//     nodes are), in shared memory ([entry][thread] double2 = (T, T'), conflict-free LDS.128) for
//     the rest.  Higher indices continue the three-term recurrence in registers.
//   * Index 1 is T_0 = 1, T_0' = 0: the first child of every node is a register rename.
//   * Dimensions above LL (d > 6, or budgets without pre-built leaf code) are walked by a small
//     uniform interpreter over a unit program (one word per distinct tail (i_{LL+1}..i_d)); its
//     per-level state lives in shared memory, touched once per unit.
//   * θ and the unit program are staged in shared memory once per CTA (θ reads are broadcasts);
//     CTAs are persistent (grid = SMs × resident CTAs).
//
// Reference semantics reproduced (within tolerance; the summation order differs and fma is used):
// smolyak_api.jl:89-110, smolyak_traversal.jl:373-379, derivatives.jl:502-511, generic_api.jl:114-128.
#pragma once
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

// number of index tuples over T dimensions with block budget R
template <int GK>
__host__ __device__ constexpr long nterms(int T, int R) {
    if (T == 0) return 1;
    long s = 0;
    for (int i = 1; i <= ellc<GK>(R); i++) s += nterms<GK>(T - 1, R - blockof<GK>(i));
    return s;
}

// ---- θ in shared memory is read in PAIRS.  Coefficients are consumed strictly in traversal order, so inside
// straight-line code every even slot is one LDS.128 that also delivers the next coefficient (halves the
// shared-memory instructions of the kernel: one per coefficient was half of all LSU traffic).  Slots are
// offsets from a 16-byte aligned block start and are compile-time constants after inlining.  Wherever a
// runtime loop changes the base, blocks are re-aligned to even slots (leaf_end is the layout; the host
// mirrors it in smolyak_leaf_theta_map to scatter θ into shared memory).
struct ThRd {
    const double *base;
    double hold;
};
__device__ __forceinline__ double th_get(ThRd &t, int off) {
    if ((off & 1) == 0) {
        const double2 v = *reinterpret_cast<const double2 *>(t.base + off);
        t.hold = v.y;
        return v.x;
    }
    return t.hold;
}
__host__ __device__ constexpr int even_up(int x) { return (x + 1) & ~1; }
// slot after the last coefficient of Leaf<T, R> when it starts at slot `off` (LOOPD: first dimension walked by runtime loops)
template <int GK>
__host__ __device__ constexpr int leaf_end(int LOOPD, int T, int R, int off) {
    if (T == 0 || R == 0) return off + 1;
    off = leaf_end<GK>(LOOPD, T - 1, R, off);
    if (T < LOOPD) {
        for (int i = 2; i <= ellc<GK>(R); i++) {
            const int b = blockof<GK>(i);
            off = R - b == 0 ? off + 1 : leaf_end<GK>(LOOPD, T - 1, R - b, off);
        }
    } else {
        for (int B = 1; B <= R; B++) {
            const int cnt = ellc<GK>(B) - ellc<GK>(B - 1);
            if (R - B == 0) off = even_up(off + cnt);
            else off = even_up(off) + cnt * even_up(leaf_end<GK>(LOOPD, T - 1, R - B, 0));
        }
    }
    return off;
}

constexpr int kTab = 6;              // table entries: indices 2..7 (degrees 1..6)

template <bool GRAD>
struct DimTab {
    double t[kTab];
    double dt[GRAD ? kTab : 1];
    double x2, x2d;
};

// per-thread evaluation context; the type carries the compile-time configuration
template <bool GRAD_, int TREG_, int FENCE_, int LOOPD_, int NT_>
struct LeafCtx {
    static constexpr int NT = NT_;          // threads per CTA (stride of the shared per-thread arrays)
    static constexpr int LOOPD = LOOPD_;    // dimensions >= LOOPD run one runtime loop per block (must be > TREG)
    static constexpr bool GRAD = GRAD_;
    static constexpr int TREG = TREG_;      // dimensions with register tables
    static constexpr int SYNCT = FENCE_;    // > 0: the CTA re-converges (bar.sync) at the entry of every minimal
                                            // sub-leaf with at least SYNCT terms, so its warps walk the long
                                            // straight-line code together and share instruction-cache lines
    DimTab<GRAD_> tabs[TREG_ > 0 ? TREG_ : 1];
    const double2 *stab;                    // shared tables of dims >= TREG, already offset by the thread id
};

template <bool GRAD, int T>
struct NC { static constexpr int value = GRAD ? T + 1 : 1; };

// (T_i, T_i') of dimension T (1-based), table entry e = i - 2
template <int T, int E, class CX>
__device__ __forceinline__ void tab_get(const CX &cx, double &Tv, double &dTv) {
    if constexpr (T - 1 < CX::TREG) {
        Tv = cx.tabs[T - 1].t[E];
        dTv = CX::GRAD ? cx.tabs[T - 1].dt[CX::GRAD ? E : 0] : 0.0;
    } else {
        const double2 v = cx.stab[((T - 1 - CX::TREG) * (kTab + 1) + E) * CX::NT];
        Tv = v.x; dTv = v.y;
    }
}
template <int T, class CX>
__device__ __forceinline__ void tab_x2(const CX &cx, double &x2, double &x2d) {
    if constexpr (T - 1 < CX::TREG) { x2 = cx.tabs[T - 1].x2; x2d = cx.tabs[T - 1].x2d; }
    else { const double2 v = cx.stab[((T - 1 - CX::TREG) * (kTab + 1) + kTab) * CX::NT]; x2 = v.x; x2d = v.y; }
}

template <int GK, int T, int R, class CX>
struct Leaf;

// indices I..ℓ(R) of dimension T
template <int GK, int T, int R, class CX, int I>
struct LeafStep {
    static constexpr bool GRAD = CX::GRAD;
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[NC<GRAD, T>::value],
                                              double &Tc, double &Tq, double &dTc, double &dTq) {
        if constexpr (I <= ellc<GK>(R)) {
            constexpr int b = blockof<GK>(I);
            double Tv, dTv = 0.0;
            if constexpr (I <= kTab + 1) {
                tab_get<T, I - 2>(cx, Tv, dTv);
            } else {
                double x2, x2d;
                tab_x2<T>(cx, x2, x2d);
                if constexpr (I == kTab + 2) {
                    tab_get<T, kTab - 1>(cx, Tc, dTc);
                    tab_get<T, kTab - 2>(cx, Tq, dTq);
                }
                Tv = fma(x2, Tc, -Tq);
                if (GRAD) { dTv = fma(x2, dTc, fma(x2d, Tc, -dTq)); dTq = dTc; dTc = dTv; }
                Tq = Tc; Tc = Tv;
            }
            if constexpr (R - b == 0) {
                // no budget left below: the child is the single coefficient θ (all lower indices are 1,
                // every lower partial is structurally zero)
                const double c0 = th_get(th, off);
                off += 1;
                if (GRAD) a[GRAD ? T : 0] = fma(dTv, c0, a[GRAD ? T : 0]);
                a[0] = fma(Tv, c0, a[0]);
            } else {
                double c[NC<GRAD, T - 1>::value];
                off = Leaf<GK, T - 1, R - b, CX>::run(cx, th, off, c);
                if (GRAD) {
                    a[GRAD ? T : 0] = fma(dTv, c[0], a[GRAD ? T : 0]);
#pragma unroll
                    for (int m = 1; m < T; m++) a[GRAD ? m : 0] = fma(Tv, c[GRAD ? m : 0], a[GRAD ? m : 0]);
                }
                a[0] = fma(Tv, c[0], a[0]);
            }
            return LeafStep<GK, T, R, CX, I + 1>::run(cx, th, off, a, Tc, Tq, dTc, dTq);
        } else {
            return off;
        }
    }
};

// Dimensions whose table is in shared memory (T > TREG): one RUNTIME loop per block.  All indices of
// a block have the same child structure Leaf<T-1, R-b>, so the child code exists once per block and
// is re-executed from the instruction cache; the table entry is a dynamically indexed LDS.128.
template <int GK, int T, int R, class CX, int B>
struct LeafBlocks {
    static constexpr bool GRAD = CX::GRAD;
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[NC<GRAD, T>::value],
                                              double &Tc, double &Tq, double &dTc, double &dTq) {
        if constexpr (B <= R) {
            constexpr int lo = ellc<GK>(B - 1) + 1, hi = ellc<GK>(B);
            // θ of this block: single coefficients are contiguous scalars; multi-term children each start an
            // aligned block of CS slots (same layout as leaf_end)
            constexpr int CS = R - B == 0 ? 1 : even_up(leaf_end<GK>(CX::LOOPD, T - 1, R - B > 0 ? R - B : 0, 0));
            const int start = R - B == 0 ? off : even_up(off);
            const double *blk = th.base + start;
            const double2 *tb = cx.stab + (size_t)(T - 1 - CX::TREG) * (kTab + 1) * CX::NT;
            double x2 = 0, x2d = 0;
            if constexpr (hi > kTab + 1) {
                const double2 v = tb[kTab * CX::NT];
                x2 = v.x; x2d = v.y;
                if constexpr (lo <= kTab + 2) {          // the recurrence starts inside (or right at) this block
                    const double2 e1 = tb[(kTab - 1) * CX::NT], e0 = tb[(kTab - 2) * CX::NT];
                    Tc = e1.x; dTc = e1.y; Tq = e0.x; dTq = e0.y;
                }
            }
#pragma unroll 1
            for (int i = lo; i <= hi; i++) {
                double Tv, dTv;
                if (hi <= kTab + 1 || i <= kTab + 1) {
                    const double2 v = tb[(i - 2) * CX::NT];
                    Tv = v.x; dTv = v.y;
                } else {
                    Tv = fma(x2, Tc, -Tq);
                    dTv = GRAD ? fma(x2, dTc, fma(x2d, Tc, -dTq)) : 0.0;
                    dTq = dTc; dTc = dTv; Tq = Tc; Tc = Tv;
                }
                if constexpr (R - B == 0) {
                    const double c0 = blk[i - lo];
                    if (GRAD) a[GRAD ? T : 0] = fma(dTv, c0, a[GRAD ? T : 0]);
                    a[0] = fma(Tv, c0, a[0]);
                } else {
                    double c[NC<GRAD, T - 1>::value];
                    ThRd sub{blk + (i - lo) * CS, 0.0};
                    Leaf<GK, T - 1, R - B, CX>::run(cx, sub, 0, c);
                    if (GRAD) {
                        a[GRAD ? T : 0] = fma(dTv, c[0], a[GRAD ? T : 0]);
#pragma unroll
                        for (int m = 1; m < T; m++) a[GRAD ? m : 0] = fma(Tv, c[GRAD ? m : 0], a[GRAD ? m : 0]);
                    }
                    a[0] = fma(Tv, c[0], a[0]);
                }
            }
            return LeafBlocks<GK, T, R, CX, B + 1>::run(cx, th, even_up(start + (hi - lo + 1) * CS), a, Tc, Tq, dTc, dTq);
        } else {
            return off;
        }
    }
};

template <int GK, int R, class CX>
struct Leaf<GK, 0, R, CX> {
    static __device__ __forceinline__ int run(const CX &, ThRd &th, int off, double (&a)[1]) {
        a[0] = th_get(th, off);
        return off + 1;
    }
};

template <int GK, int T, int R, class CX>
struct Leaf {
    static constexpr bool GRAD = CX::GRAD;
    // a[0] = Σ θ·Π T, a[m] = ∂/∂y_m (m = 1..T) over {(i_1..i_T): Σ b_j <= R}; consumes θ in traversal order
    static __device__ __forceinline__ int run(const CX &cx, ThRd &th, int off, double (&a)[NC<GRAD, T>::value]) {
        if constexpr (R == 0) {
            a[0] = th_get(th, off);
            if (GRAD) {
#pragma unroll
                for (int m = 1; m <= T; m++) a[GRAD ? m : 0] = 0.0;
            }
            return off + 1;
        } else {
            if constexpr (CX::SYNCT > 0 && nterms<GK>(T, R) >= CX::SYNCT && nterms<GK>(T - 1, R) < CX::SYNCT) __syncthreads();
            {   // index 1: T_0 = 1, derivative 0
                double c[NC<GRAD, T - 1>::value];
                off = Leaf<GK, T - 1, R, CX>::run(cx, th, off, c);
                a[0] = c[0];
                if (GRAD) {
#pragma unroll
                    for (int m = 1; m < T; m++) a[GRAD ? m : 0] = c[GRAD ? m : 0];
                    a[GRAD ? T : 0] = 0.0;
                }
            }
            double Tc = 0, Tq = 0, dTc = 0, dTq = 0;
            if constexpr (T < CX::LOOPD) return LeafStep<GK, T, R, CX, 2>::run(cx, th, off, a, Tc, Tq, dTc, dTq);
            else return LeafBlocks<GK, T, R, CX, 1>::run(cx, th, off, a, Tc, Tq, dTc, dTq);
        }
    }
};

// ---- kernel
struct LeafParams {
    const double *x;
    double *out;
    const double *theta;
    const uint32_t *units;        // r | carry << 8
    const uint32_t *thmap;        // shared-memory slot of every coefficient (pair-aligned blocks, see leaf_end)
    int64_t n, K, n_units, Kslots;
    int small_cta;                // 0 full-size CTA; 2, 3: 128-thread variants (chosen by the plan: smolyak_leaf_cta)
    int64_t theta_stride, out_stride;   // coefficient vector v of a K×nvec block: θ[k·nvec + v] → out[p·nvec + v]
    int D, nup;
    int budget;                   // B of the plan: no unit has a larger r
    // fused gather (PEER instantiations, sk_multi.cu gather mode 2): the results are also stored into the full-output buffers
    // of up to 7 other devices, peer to peer over NVLink, by the threads that computed them
    int n_peer;
    double *peer[7];              // already offset to this launch's first point
    sk_transform tr[SK_MAX_DIM];
};

enum { F_X0 = 0, F_X1, F_T, F_TP, F_DT, F_DTP, F_ACC };   // upper-level fields, then the level's accumulators
// Upper level u (dimension LL+u+1) carries value + partials 1..LL+u+1 (gradient) or the value only: exactly that many
// accumulators are kept (the state of all levels is what limits the CTAs per SM for high-dimensional bases).
__host__ __device__ constexpr int up_ncomp(bool grad, int LL, int u) { return grad ? LL + u + 2 : 1; }
__host__ __device__ constexpr int up_offset(bool grad, int LL, int u) {      // doubles of the levels below u
    return grad ? u * (F_ACC + LL + 2) + u * (u - 1) / 2 : u * (F_ACC + 1);
}

// shared memory layout (bytes), shared by host and device
struct LeafSmem {
    size_t stab, up, units, theta, total;
    __host__ __device__ LeafSmem(int NT, int LL, int treg, int D, int nup, int64_t n_units, int64_t K, bool grad) {
        const int sdims = LL > treg ? LL - treg : 0;
        stab = 0;
        up = stab + (size_t)sdims * (kTab + 1) * NT * 16;
        units = up + (size_t)up_offset(grad, LL, nup) * NT * 8;
        theta = units + (size_t)((n_units + 3) & ~(int64_t)3) * 4;
        total = theta + (size_t)K * 8;
    }
};

// RONLY >= 0: single-budget kernel for bases that are one leaf of budget RONLY (d == LL).  The multi-budget kernel
// holds the code of every budget <= RMAX and is register-allocated for the largest: at (LL, R) = (3, 3) it ran with
// 254 registers and 268 spill stores per point (profiles/r1_leaf_variants.md, "single-budget kernels").
template <int GK, int LL, int RMAX, bool GRAD, int TREG, int MINB, int FENCE, int LOOPD, int NT, int RONLY = -1, bool PEER = false>
__global__ void __launch_bounds__(NT, MINB) smolyak_leaf_kernel(const __grid_constant__ LeafParams q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int D = q.D, NUP = q.nup;
    const LeafSmem lay(NT, LL, TREG, D, NUP, q.n_units, q.Kslots, GRAD);
    double2 *stab = reinterpret_cast<double2 *>(smem_raw + lay.stab) + tid;
    double *up = reinterpret_cast<double *>(smem_raw + lay.up);           // [level: fields, accumulators][threads]
    uint32_t *s_units = reinterpret_cast<uint32_t *>(smem_raw + lay.units);
    double *s_theta = reinterpret_cast<double *>(smem_raw + lay.theta);
    for (int64_t i = tid; i < q.n_units; i += NT) s_units[i] = q.units[i];
    for (int64_t i = tid; i < q.K; i += NT) s_theta[q.thmap[i]] = q.theta[i * q.theta_stride];
    __syncthreads();
#define UP(u, f) up[(size_t)(up_offset(GRAD, LL, (u)) + (f)) * NT + tid]
    constexpr int NCL = NC<GRAD, LL>::value;
    const int E = GRAD ? D + 1 : 1;

    // block-uniform tile loop (the leaf code contains CTA barriers): threads past the end evaluate the
    // last point again and skip the store
    for (int64_t p0 = (int64_t)blockIdx.x * NT; p0 < q.n; p0 += (int64_t)gridDim.x * NT) {
        const bool live = p0 + tid < q.n;
        const int64_t p = live ? p0 + tid : q.n - 1;
        const double *xp = q.x + p * D;                      // points and results are streamed (ld.cs / st.cs): they must not
                                                              // evict the spill lines, which live in L2
        using CX = LeafCtx<GRAD, TREG, FENCE, LOOPD, NT>;
        CX cx;
        cx.stab = stab;
#pragma unroll
        for (int j = 0; j < LL; j++) {
            double xj[GRAD ? 2 : 1];
            skd::to_pm1_jet<GRAD ? 2 : 1>(q.tr[j], __ldcs(xp + j), xj);
            const double x0 = xj[0], x1 = GRAD ? xj[GRAD ? 1 : 0] : 0.0;
            double t[kTab], dt[kTab];
            const double x2 = 2.0 * x0, x2d = 2.0 * x1;
            t[0] = x0; dt[0] = x1;
            t[1] = fma(x2, x0, -1.0); dt[1] = x2 * x2d;
#pragma unroll
            for (int i = 2; i < kTab; i++) {
                t[i] = fma(x2, t[i - 1], -t[i - 2]);
                dt[i] = GRAD ? fma(x2, dt[i - 1], fma(x2d, t[i - 1], -dt[i - 2])) : 0.0;
            }
            if (j < TREG) {
                DimTab<GRAD> &tb = cx.tabs[j < TREG ? j : 0];
                tb.x2 = x2; tb.x2d = x2d;
#pragma unroll
                for (int i = 0; i < kTab; i++) { tb.t[i] = t[i]; if (GRAD) tb.dt[GRAD ? i : 0] = dt[i]; }
            } else {
                double2 *sd = stab + (size_t)(j - TREG) * (kTab + 1) * NT;
#pragma unroll
                for (int i = 0; i < kTab; i++) sd[i * NT] = make_double2(t[i], dt[i]);
                sd[kTab * NT] = make_double2(x2, x2d);
            }
        }
        for (int u = 0; u < NUP; u++) {
            double xj[GRAD ? 2 : 1];
            skd::to_pm1_jet<GRAD ? 2 : 1>(q.tr[LL + u], __ldcs(xp + LL + u), xj);
            UP(u, F_X0) = xj[0];
            if (GRAD) UP(u, F_X1) = xj[GRAD ? 1 : 0];
        }
        uint32_t first = 0xFFFFFFFFu;                       // bit u: upper level u sits at index 1
        const double *ub = s_theta;                         // aligned block of the current unit
        double a[NCL];
        for (int64_t un = 0; un < q.n_units; un++) {
            const uint32_t w = s_units[un];
            const int r = RONLY >= 0 ? RONLY : (int)(w & 0xFF), c = (int)(w >> 8);
            if constexpr (RONLY >= 0) {
                ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, RONLY >= 0 ? RONLY : 0, CX>::run(cx, th, 0, a));
            } else
            switch (r) {
            case 0: { ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, 0, CX>::run(cx, th, 0, a)); } break;
            case 1: if constexpr (RMAX >= 1) { ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, RMAX >= 1 ? 1 : 0, CX>::run(cx, th, 0, a)); } break;
            case 2: if constexpr (RMAX >= 2) { ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, RMAX >= 2 ? 2 : 0, CX>::run(cx, th, 0, a)); } break;
            case 3: if constexpr (RMAX >= 3) { ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, RMAX >= 3 ? 3 : 0, CX>::run(cx, th, 0, a)); } break;
            case 4: if constexpr (RMAX >= 4) { ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, RMAX >= 4 ? 4 : 0, CX>::run(cx, th, 0, a)); } break;
            case 5: if constexpr (RMAX >= 5) { ThRd th{ub, 0.0}; ub += even_up(Leaf<GK, LL, RMAX >= 5 ? 5 : 0, CX>::run(cx, th, 0, a)); } break;
            default: break;
            }
            if (NUP > 0) {
                // fold the leaf into upper level 0 (dimension LL+1)
                if (first & 1u) {
#pragma unroll
                    for (int m = 0; m < NCL; m++) UP(0, F_ACC + m) = a[m];
                    if (GRAD) UP(0, F_ACC + LL + 1) = 0.0;
                } else if (r == 0) {
                    // no budget left for the leaf: it is the single coefficient a[0], all its partials are structurally
                    // zero (three quarters of the units of a high-dimensional basis)
                    UP(0, F_ACC) = fma(UP(0, F_T), a[0], UP(0, F_ACC));
                    if (GRAD) UP(0, F_ACC + LL + 1) = fma(UP(0, F_DT), a[0], UP(0, F_ACC + LL + 1));
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
                            const double X2 = 2.0 * UP(u, F_X0), Tc = UP(u, F_T), Tq = UP(u, F_TP);
                            UP(u, F_T) = fma(X2, Tc, -Tq); UP(u, F_TP) = Tc;
                            if (GRAD) {
                                const double dTc = UP(u, F_DT);
                                UP(u, F_DT) = fma(X2, dTc, fma(2.0 * UP(u, F_X1), Tc, -UP(u, F_DTP)));
                                UP(u, F_DTP) = dTc;
                            }
                        }
                        break;
                    }
                }
            }
        }
        double *o = q.out + p * q.out_stride;
        if (live) {
            if (NUP > 0) {
                for (int m = 0; m < E; m++) __stcs(o + m, UP(NUP - 1, F_ACC + m));
            } else {
#pragma unroll
                for (int m = 0; m < NCL; m++) __stcs(o + m, a[m]);
            }
        }
        if constexpr (PEER) {
            // Peer stores must be whole 32-byte sectors: a remote write is forwarded as it is issued (no merging in the local
            // L2), and the per-thread layout [point][NCL] makes every 8-byte store a partial sector (measured: 154 GB/s per
            // GPU at 8 GPUs).  A full warp therefore stages its 32·NCL results in the shared-memory slots of its own threads'
            // dimension-(TREG+1) table — free until the next tile rebuilds it — and writes them out as consecutive 16-byte
            // chunks.
            const int lane = tid & 31;
            const bool full = __all_sync(0xffffffffu, live);
            const int64_t pw = p - lane;                                   // first point of this warp
            bool vec = NCL > 1 && LL > TREG && full && ((NCL * 32) % 2 == 0);
            if (vec) for (int g = 0; g < q.n_peer; g++) vec = vec && ((reinterpret_cast<uintptr_t>(q.peer[g] + pw * q.out_stride) & 15) == 0);
            vec = __all_sync(0xffffffffu, vec);
            if (vec) {
                double *seg = reinterpret_cast<double *>(stab - lane);      // entry e of the warp: seg + e·NT·2, 64 doubles each
                __syncwarp();
#pragma unroll
                for (int m = 0; m < NCL; m++) { const int i = lane * NCL + m; seg[(size_t)(i >> 6) * NT * 2 + (i & 63)] = a[m]; }
                __syncwarp();
                for (int c = lane; c < NCL * 16; c += 32) {
                    const int i = 2 * c;
                    const double2 v = *reinterpret_cast<const double2 *>(seg + (size_t)(i >> 6) * NT * 2 + (i & 63));
                    for (int g = 0; g < q.n_peer; g++) __stcs(reinterpret_cast<double2 *>(q.peer[g] + pw * q.out_stride + i), v);
                }
                __syncwarp();
            } else if (live) {
                for (int g = 0; g < q.n_peer; g++) {
                    double *og = q.peer[g] + p * q.out_stride;
#pragma unroll
                    for (int m = 0; m < NCL; m++) __stcs(og + m, a[m]);
                }
            }
        }
    }
#undef UP
}

template <int GK, int LL, int RMAX, bool GRAD, int TREG, int MINB, int FENCE, int LOOPD, int NT, int RONLY = -1, bool PEER = false>
cudaError_t launch_leaf(const LeafParams &q, cudaStream_t stream) {
    auto kern = smolyak_leaf_kernel<GK, LL, RMAX, GRAD, TREG, MINB, FENCE, LOOPD, NT, RONLY, PEER>;
    const LeafSmem lay(NT, LL, TREG, q.D, q.nup, q.n_units, q.Kslots, GRAD);
    const size_t smem = lay.total;
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

// leaf depth / budget combinations that are built (code size and compile time bound the budget)
// Depths 12, 14 and 16 (budget <= 2, a few hundred terms) do the same for d = 11..16: d = 16, B = 2 went from 64 to
// 2 300 Mpoints/s (value + gradient) once no dimension was left to the interpreter.
// Depths 8 and 10 exist for the small budgets of high-dimensional bases: there the unit interpreter, not the
// arithmetic, is the cost (profiles/r1_leaf_variants.md), and a leaf over 8 or 10 dimensions with budget <= 3 or 2 is
// a few hundred terms of straight-line code.
constexpr int kLeafMaxLL = 16;
__host__ __device__ constexpr bool leaf_built_c(int LL) { return (LL >= 1 && LL <= 6) || LL == 8 || LL == 10 || LL == 12 || LL == 14 || LL == 16; }
__host__ __device__ constexpr int leaf_rmax_c(int gk, int LL) {
    if (LL >= 9) return 2;
    if (LL >= 7) return gk == SK_GRID_INTERIOR ? 2 : 3;
    if (gk == SK_GRID_INTERIOR) return LL <= 4 ? 4 : 3;             // ℓ = 3^b: 1 473 terms at (LL, R) = (4, 4), 737 at (6, 3)
    return LL <= 4 ? 5 : 4;
}
// tuning measured on B200 for the d = 6, B = 4 gradient case (profiles/r1_leaf_variants.md): 12 warps/SM at
// 168 registers with two register tables and only the top dimension looped is the best trade between
// latency hiding, spills and instruction-cache reuse
constexpr int leaf_treg_c(int LL, bool grad) { const int t = LL >= 7 ? (grad ? 3 : 4) : grad ? (LL >= 5 ? 2 : 3) : (LL >= 5 ? 3 : 4); return LL < t ? LL : t; }
constexpr int leaf_minb_c(int LL, bool grad) { return (LL >= 5 || (grad && LL >= 3)) ? 1 : 2; }
constexpr int leaf_nt_c(int LL, bool grad) { return LL >= 7 ? 128 : LL >= 5 ? (grad ? 384 : 512) : 256; }   // threads per CTA (deep leaves: 255 registers each)
// small-CTA variants for bases with dimensions above the leaf (d > LL): the upper-level state is (9 + d) doubles per
// thread and level, which does not fit next to the tables of a 256..512-thread CTA.  128 threads have 255 registers
// each, so one more dimension's table moves to registers.
constexpr int kLeafSmallNT = 128;
// register tables of the small CTAs: 3 (gradient) / 4 dimensions, or 4 for both when that lets one more CTA fit per SM
constexpr int leaf_treg_small_c(int LL, bool grad, bool lean = false) { const int t = grad && !lean ? 3 : 4; return LL < t ? LL : t; }
constexpr int leaf_sync_c(int) { return 0; }       // CTA re-convergence granularity (terms); measured: no gain
constexpr int leaf_loopd_c(int LL, bool grad) { return grad && LL >= 5 ? LL : 99; }   // first dimension walked by per-block runtime loops

// Bases with dimensions above the leaf: the kernel holds the leaf code of every budget <= RMAX and is register-allocated
// for the largest, so plans with budget 2 or 3 get instantiations that stop there (same reason as RONLY above).
template <int GK, int LL, bool GRAD, int TREG, int MINB, int LOOPD, int NT>
cudaError_t launch_leaf_budget(const LeafParams &q, cudaStream_t stream) {
    constexpr int RM = leaf_rmax_c(GK, LL);
    if constexpr (RM > 2) if (q.budget <= 2) return launch_leaf<GK, LL, 2, GRAD, TREG, MINB, 0, LOOPD, NT>(q, stream);
    if constexpr (RM > 3) if (q.budget <= 3) return launch_leaf<GK, LL, 3, GRAD, TREG, MINB, 0, LOOPD, NT>(q, stream);
    return launch_leaf<GK, LL, RM, GRAD, TREG, MINB, 0, LOOPD, NT>(q, stream);
}

// one dispatcher per (family, GRAD), each in its own translation unit (sk_leaf_g*.cu)
template <int GK, bool GRAD>
cudaError_t dispatch_leaf(int LL, const LeafParams &q, cudaStream_t stream) {
#ifdef SK_LEAF_VARIANTS
    if constexpr (GK == SK_GRID_INTERIOR2) {
        const char *e = getenv("SK_LEAF_VARIANT");
        const int v = e ? atoi(e) : 0;
        if (LL == 6 && v > 0) {
            switch (v) {      // NB: LOOPD must equal leaf_loopd_c (the host's coefficient map depends on it)
            case 1: return launch_leaf<GK, 6, 4, GRAD, 2, 1, 0, 6, 352>(q, stream);
            case 2: return launch_leaf<GK, 6, 4, GRAD, 3, 1, 0, 6, 256>(q, stream);
            case 3: return launch_leaf<GK, 6, 4, GRAD, 3, 1, 0, 6, 384>(q, stream);
            case 4: return launch_leaf<GK, 6, 4, GRAD, 4, 1, 0, 6, 256>(q, stream);
            }
        }
    }
#endif
    if (q.small_cta == 3) {      // gradient, four register tables: smaller shared-memory footprint, more CTAs per SM
        if constexpr (GRAD) {
            switch (LL) {
            case 4: return launch_leaf_budget<GK, 4, GRAD, leaf_treg_small_c(4, GRAD, true), 1, leaf_loopd_c(4, GRAD), kLeafSmallNT>(q, stream);
            case 6: return launch_leaf_budget<GK, 6, GRAD, leaf_treg_small_c(6, GRAD, true), 1, leaf_loopd_c(6, GRAD), kLeafSmallNT>(q, stream);
            }
        }
        return cudaErrorInvalidValue;
    }
    if (q.small_cta) {      // LL in {2, 4, 6} are the depths that occur together with upper dimensions (leaf_rmax_c)
        switch (LL) {
        case 2: return launch_leaf_budget<GK, 2, GRAD, leaf_treg_small_c(2, GRAD), 1, leaf_loopd_c(2, GRAD), kLeafSmallNT>(q, stream);
        case 4: return launch_leaf_budget<GK, 4, GRAD, leaf_treg_small_c(4, GRAD), 1, leaf_loopd_c(4, GRAD), kLeafSmallNT>(q, stream);
        case 6: return launch_leaf_budget<GK, 6, GRAD, leaf_treg_small_c(6, GRAD), 1, leaf_loopd_c(6, GRAD), kLeafSmallNT>(q, stream);
        }
        return cudaErrorInvalidValue;
    }
    switch (LL) {
    case 1: return launch_leaf<GK, 1, leaf_rmax_c(GK, 1), GRAD, leaf_treg_c(1, GRAD), leaf_minb_c(1, GRAD), leaf_sync_c(1), leaf_loopd_c(1, GRAD), leaf_nt_c(1, GRAD)>(q, stream);
    case 2: return launch_leaf<GK, 2, leaf_rmax_c(GK, 2), GRAD, leaf_treg_c(2, GRAD), leaf_minb_c(2, GRAD), leaf_sync_c(2), leaf_loopd_c(2, GRAD), leaf_nt_c(2, GRAD)>(q, stream);
    case 3: return launch_leaf<GK, 3, leaf_rmax_c(GK, 3), GRAD, leaf_treg_c(3, GRAD), leaf_minb_c(3, GRAD), leaf_sync_c(3), leaf_loopd_c(3, GRAD), leaf_nt_c(3, GRAD)>(q, stream);
    case 4: return launch_leaf<GK, 4, leaf_rmax_c(GK, 4), GRAD, leaf_treg_c(4, GRAD), leaf_minb_c(4, GRAD), leaf_sync_c(4), leaf_loopd_c(4, GRAD), leaf_nt_c(4, GRAD)>(q, stream);
    case 5: return launch_leaf<GK, 5, leaf_rmax_c(GK, 5), GRAD, leaf_treg_c(5, GRAD), leaf_minb_c(5, GRAD), leaf_sync_c(5), leaf_loopd_c(5, GRAD), leaf_nt_c(5, GRAD)>(q, stream);
    case 6:
        // six-dimensional bases (the headline): an instantiation that also stores into the other devices' buffers
        if (q.n_peer > 0 && q.nup == 0)
            return launch_leaf<GK, 6, leaf_rmax_c(GK, 6), GRAD, leaf_treg_c(6, GRAD), leaf_minb_c(6, GRAD), leaf_sync_c(6), leaf_loopd_c(6, GRAD), leaf_nt_c(6, GRAD), -1, true>(q, stream);
        return launch_leaf<GK, 6, leaf_rmax_c(GK, 6), GRAD, leaf_treg_c(6, GRAD), leaf_minb_c(6, GRAD), leaf_sync_c(6), leaf_loopd_c(6, GRAD), leaf_nt_c(6, GRAD)>(q, stream);
    case 8: return launch_leaf_budget<GK, 8, GRAD, leaf_treg_c(8, GRAD), leaf_minb_c(8, GRAD), leaf_loopd_c(8, GRAD), leaf_nt_c(8, GRAD)>(q, stream);
    case 10: return launch_leaf<GK, 10, leaf_rmax_c(GK, 10), GRAD, leaf_treg_c(10, GRAD), leaf_minb_c(10, GRAD), leaf_sync_c(10), leaf_loopd_c(10, GRAD), leaf_nt_c(10, GRAD)>(q, stream);
    case 12: return launch_leaf<GK, 12, leaf_rmax_c(GK, 12), GRAD, leaf_treg_c(12, GRAD), leaf_minb_c(12, GRAD), leaf_sync_c(12), leaf_loopd_c(12, GRAD), leaf_nt_c(12, GRAD)>(q, stream);
    case 14: return launch_leaf<GK, 14, leaf_rmax_c(GK, 14), GRAD, leaf_treg_c(14, GRAD), leaf_minb_c(14, GRAD), leaf_sync_c(14), leaf_loopd_c(14, GRAD), leaf_nt_c(14, GRAD)>(q, stream);
    case 16: return launch_leaf<GK, 16, leaf_rmax_c(GK, 16), GRAD, leaf_treg_c(16, GRAD), leaf_minb_c(16, GRAD), leaf_sync_c(16), leaf_loopd_c(16, GRAD), leaf_nt_c(16, GRAD)>(q, stream);
    }
    return cudaErrorInvalidValue;
}

// single-budget kernels (RONLY) for bases that are exactly one leaf (d == LL, one unit of budget B).  Built where they
// beat the multi-budget kernel of the same depth (profiles/r1_leaf_variants.md, "single-budget kernels"): every
// budget >= 2 for LL <= 4; budgets below the top one for LL = 5, 6 (the multi-budget kernels of those depths are
// tuned for their top budget), except the (6, 3) gradient where the 384-thread multi-budget kernel stays ahead.
// cudaErrorNotSupported = not built, the caller falls back to dispatch_leaf.
inline bool leaf_single_enabled() { static const bool on = getenv("SK_LEAF_NO_SINGLE") == nullptr; return on; }
// deep single leaves (EndpointGrid, InteriorGrid2; instantiated in sk_leaf_t*.cu): budget 3 at d = 7..12 except d = 8,
// which the depth-8 multi-budget kernel is tuned for (d = 10, B = 3, ten state variables at level 3, is the classic
// Smolyak setting), and budget 4 at d = 7, 8
constexpr bool leaf_single_deep_c(int gk, int LL, int R) {
    if (gk == SK_GRID_INTERIOR) return false;
    return (R == 3 && LL >= 7 && LL <= 12 && LL != 8) || (R == 4 && LL >= 7 && LL <= 10);
}
// budget 4 in nine and ten dimensions (9 361 / 13 441 terms for InteriorGrid2): every dimension from the fifth up runs
// per-block loops, so the code is the leaves of dimensions 1..4 plus a few dozen small loop bodies, and four register
// tables leave shared memory to θ (107 KB at d = 10)
constexpr bool leaf_single_deep_looped_c(int LL, int R) { return R == 4 && (LL == 9 || LL == 10); }
constexpr bool leaf_single_built_c(int gk, int LL, int R, bool grad) {
    if (leaf_single_deep_c(gk, LL, R)) return true;
    // InteriorGrid (ℓ = 3^b) budget 4 at five and six dimensions (2 641 / 4 361 terms) exists only here: inside a
    // multi-budget kernel it slowed every smaller budget (register allocation for the largest)
    if (gk == SK_GRID_INTERIOR && (LL == 5 || LL == 6) && R == 4) return true;
    if (gk != SK_GRID_INTERIOR && (LL == 5 || LL == 6) && R == 5) return true;      // 2 203..10 625 terms
    if (LL < 1 || LL > 6 || R < 1 || R > (LL <= 4 ? 5 : leaf_rmax_c(gk, LL))) return false;
    if (gk == SK_GRID_INTERIOR && R > 4) return false;      // ℓ(5) = 243 indices: beyond the front end's template recursion depth
    if (LL <= 4) return true;
#ifdef SK_SINGLE_HEADLINE
    if (gk == SK_GRID_INTERIOR2 && LL == 6 && R == 4 && grad) return true;
#endif
    if (R == leaf_rmax_c(gk, LL)) return false;
    return !(grad && LL == 6 && R == 3);
}
// measured choice per budget (variants A..E of the log): small budgets fit two CTAs of 256 threads per SM (<= 128
// registers); from budget 3 (gradients; 4 for values) two register tables instead of three or four leave the registers
// to the accumulators of the unrolled recursion
// Two-dimensional leaves with long index ranges (ℓ(R) > 27: InteriorGrid2 budget 5, InteriorGrid budget 4) walk
// dimension 2 with one runtime loop per block (LeafBlocks) instead of unrolling it index by index: +47..85 %.  For
// three and four dimensions, and for EndpointGrid, the unrolled code stays ahead (r1_leaf_variants.md).
#ifndef SK_SL_TREG
#define SK_SL_TREG 1
#endif
constexpr int leaf_ell_c(int gk, int b) {
    return gk == SK_GRID_ENDPOINT ? ellc<SK_GRID_ENDPOINT>(b) : gk == SK_GRID_INTERIOR ? ellc<SK_GRID_INTERIOR>(b) : ellc<SK_GRID_INTERIOR2>(b);
}
constexpr bool leaf_single_looped_c(int gk, int LL, int R) { return LL == 2 && gk != SK_GRID_ENDPOINT && leaf_ell_c(gk, R) > 27; }
constexpr int leaf_single_treg_c(int gk, int LL, int R, bool grad) {
    if (leaf_single_deep_looped_c(LL, R)) return 4;
    if (LL > 6) return leaf_treg_c(LL, grad);
    if (leaf_single_looped_c(gk, LL, R)) return LL - 1 < SK_SL_TREG ? LL - 1 : SK_SL_TREG;
    if (gk == SK_GRID_INTERIOR2 && LL == 6 && R == 5) return 3;      // θ (85 KB) + three shared-memory tables fit, four do not
    const int t = (R >= 4 || (R == 3 && (grad || gk == SK_GRID_INTERIOR))) ? 2 : leaf_treg_c(LL, grad);
    return LL < t ? LL : t;
}
// the largest single leaves (2 641..10 625 terms): their straight-line code is several hundred KB when only the top
// dimension is looped, far beyond the instruction cache, so the gradient kernels of InteriorGrid / InteriorGrid2 loop
// dimensions 4.. (+24..59 %; the EndpointGrid ones and all values-only kernels lose 15..45 % by it and stay as they were:
// profiles/r2_big_leaves.md).  SK_BIG_LOOPD_G / _V: first looped dimension for gradients / values (experiments)
#ifndef SK_BIG_LOOPD_G
#define SK_BIG_LOOPD_G 4
#endif
#ifndef SK_BIG_LOOPD_V
#define SK_BIG_LOOPD_V 99
#endif
constexpr bool leaf_single_big_c(int gk, int LL, int R) {
    return (LL == 5 || LL == 6) && ((gk == SK_GRID_INTERIOR && R == 4) || (gk == SK_GRID_INTERIOR2 && R == 5));
}
// first looped dimension; the host's coefficient layout (smolyak_leaf_theta_map) follows it
constexpr int leaf_single_loopd_c(int gk, int LL, int R, bool grad) {
    if (leaf_single_looped_c(gk, LL, R)) return leaf_single_treg_c(gk, LL, R, grad) + 1;
    if (leaf_single_deep_looped_c(LL, R)) return 5;
    const int base = leaf_loopd_c(LL, grad);
    if (leaf_single_big_c(gk, LL, R)) {
        const int lo = leaf_single_treg_c(gk, LL, R, grad) + 1, want = grad ? SK_BIG_LOOPD_G : SK_BIG_LOOPD_V;
        const int v = want < lo ? lo : want;
        return v < base ? v : base;
    }
    return base;
}
#ifdef SK_SINGLE_HEADLINE
constexpr int leaf_single_nt_c(int gk, int LL, int R, bool grad) { return gk == SK_GRID_INTERIOR2 && LL == 6 && R == 4 ? leaf_nt_c(LL, grad) : 256; }
#else
constexpr int leaf_single_nt_c(int, int LL, int, bool grad) { return LL > 6 ? leaf_nt_c(LL, grad) : 256; }
#endif
constexpr int leaf_single_minb_c(int gk, int R, int LL = 1) { return LL > 6 ? 1 : R <= 3 || (R == 4 && gk == SK_GRID_ENDPOINT) ? 2 : 1; }
template <int GK, bool GRAD, int LL, int R>
cudaError_t launch_leaf_single(const LeafParams &q, cudaStream_t stream) {
    if constexpr (leaf_single_built_c(GK, LL, R, GRAD))
        return launch_leaf<GK, LL, R, GRAD, leaf_single_treg_c(GK, LL, R, GRAD), leaf_single_minb_c(GK, R, LL), 0, leaf_single_loopd_c(GK, LL, R, GRAD), leaf_single_nt_c(GK, LL, R, GRAD), R>(q, stream);
    else
        return cudaErrorNotSupported;
}
template <int GK, bool GRAD, int LL>
cudaError_t dispatch_leaf_single_r(int R, const LeafParams &q, cudaStream_t stream) {
    switch (R) {
    case 1: return launch_leaf_single<GK, GRAD, LL, 1>(q, stream);
    case 2: return launch_leaf_single<GK, GRAD, LL, 2>(q, stream);
    case 3: return launch_leaf_single<GK, GRAD, LL, 3>(q, stream);
    case 4: return launch_leaf_single<GK, GRAD, LL, 4>(q, stream);
    case 5: return launch_leaf_single<GK, GRAD, LL, 5>(q, stream);
    }
    return cudaErrorNotSupported;
}
template <int GK, bool GRAD>
cudaError_t dispatch_leaf_single_deep(int LL, int R, const LeafParams &q, cudaStream_t stream) {
    if (R == 3) switch (LL) {
    case 7: return launch_leaf_single<GK, GRAD, 7, 3>(q, stream);
    case 9: return launch_leaf_single<GK, GRAD, 9, 3>(q, stream);
    case 10: return launch_leaf_single<GK, GRAD, 10, 3>(q, stream);
    case 11: return launch_leaf_single<GK, GRAD, 11, 3>(q, stream);
    case 12: return launch_leaf_single<GK, GRAD, 12, 3>(q, stream);
    }
    if (R == 4) switch (LL) {
    case 7: return launch_leaf_single<GK, GRAD, 7, 4>(q, stream);
    case 8: return launch_leaf_single<GK, GRAD, 8, 4>(q, stream);
    case 9: return launch_leaf_single<GK, GRAD, 9, 4>(q, stream);
    case 10: return launch_leaf_single<GK, GRAD, 10, 4>(q, stream);
    }
    return cudaErrorNotSupported;
}
template <int GK, bool GRAD>
cudaError_t dispatch_leaf_single(int LL, int R, const LeafParams &q, cudaStream_t stream) {
    switch (LL) {
    case 1: return dispatch_leaf_single_r<GK, GRAD, 1>(R, q, stream);
    case 2: return dispatch_leaf_single_r<GK, GRAD, 2>(R, q, stream);
    case 3: return dispatch_leaf_single_r<GK, GRAD, 3>(R, q, stream);
    case 4: return dispatch_leaf_single_r<GK, GRAD, 4>(R, q, stream);
    case 5: return dispatch_leaf_single_r<GK, GRAD, 5>(R, q, stream);
    case 6: return dispatch_leaf_single_r<GK, GRAD, 6>(R, q, stream);
    }
    return cudaErrorNotSupported;
}

}  // namespace skk
