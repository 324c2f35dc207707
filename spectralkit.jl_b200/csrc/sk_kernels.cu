// sm_100a kernels, part 1: univariate Chebyshev paths, the reference-order (direct) Smolyak
// paths, batched coordinate transformations and the FP64 pipe microbenchmarks.
//
// Compiled with -fmad=false (see sk_device.cuh): "exact" kernels reproduce the reference's
// unfused operation order bit for bit; "fast" kernels use explicit fma().
//
// Kernel map:
//   K1  uni_lincomb_kernel        fused transform + recurrence + dot (generic_api.jl:114-128,
//                                 chebyshev.jl:60-70, transformations.jl:183-333)
//   K3  uni_basis_kernel /        collocation_matrix rows (generic_api.jl:217-227), streamed out
//       smolyak_direct_kernel<BASIS>  with the bulk-copy engine (cp.async.bulk smem→global)
//   K2' smolyak_direct_*_kernel   direct-form Smolyak product + sequential dot in the reference's
//                                 order (smolyak_traversal.jl:373-379, derivatives.jl:491,502-511)
// Elsewhere: K2 unrolled-leaf nested kernel (sk_leaf*.cu, headline), generic nested kernel (sk_nested.cu), nested
// kernel for any ∂ specification (sk_general.cu), K4 coefficient-block DMMA kernel (sk_block.cu), fit (sk_solve.cu).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <mutex>
#include <cstdlib>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

using skd::Jet;

static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && cached[dev]) return cached[dev];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 64) cached[dev] = n;
    return n;
}

// -------------------------------------------------------------------------------------------
// bulk-copy (TMA engine, non-tensor form) helpers: shared::cta → global, bulk-group completion
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// -------------------------------------------------------------------------------------------
// K1: univariate linear combination, one point per thread, θ staged in shared memory
// -------------------------------------------------------------------------------------------
// θ of the fast kernels in the constant bank: θ is warp-uniform, and a uniform operand (LDCU.64 UR, c[3][UR+imm] →
// DFMA R, R, UR, R) leaves two register operands per fma — three fresh ones run at 57 % of the DFMA rate
// (profiles/micro/dfma_operands.cu).  Two slots, reused in stream order (events), so two streams can be in flight.
constexpr int kUniConstMax = 3840;
static __constant__ double c_uni_theta[2 * kUniConstMax];

template <int DP1, bool EXACT, bool THC = false>
__global__ void __launch_bounds__(256) uni_lincomb_kernel(int N, sk_transform tr, const double *__restrict__ theta,
                                                          int theta_in_smem, const double *__restrict__ x, int64_t n,
                                                          double *__restrict__ out) {
    extern __shared__ double s_theta[];
    const double *th = theta;
    if (!THC && theta_in_smem) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) s_theta[i] = theta[i];
        __syncthreads();
        th = s_theta;
    }
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        double xj[DP1];
        skd::to_pm1_jet<DP1>(tr, x[p], xj);
        double s[DP1];
        if (EXACT) {
            // generic_api.jl:114-128: s = θ₁·b₁; s = s + θ_k·b_k, jets in the reference's order
            Jet<DP1> X, fp = skd::jet_one<DP1>(), fpp;
#pragma unroll
            for (int r = 0; r < DP1; r++) { X.c[r] = xj[r]; fpp.c[r] = xj[r]; }
            double th0 = th[0];
#pragma unroll
            for (int r = 0; r < DP1; r++) s[r] = th0 * fp.c[r];
            for (int k = 1; k < N; k++) {
                Jet<DP1> f = skd::cheb_step_exact<DP1>(X, fp, fpp);
                fpp = fp; fp = f;
                double tk = th[k];
#pragma unroll
                for (int r = 0; r < DP1; r++) s[r] = s[r] + tk * f.c[r];
            }
        } else {
            // same recurrence with explicit fma contraction (DP1 <= 3)
            double t0 = 2.0 * xj[0];
            double t1 = DP1 > 1 ? 2.0 * xj[DP1 > 1 ? 1 : 0] : 0.0;
            double tt1 = 2.0 * t1;                                     // binom(2,1)·t1
            double t2 = DP1 > 2 ? 2.0 * xj[DP1 > 2 ? 2 : 0] : 0.0;
            double fp[DP1], fpp[DP1];
#pragma unroll
            for (int r = 0; r < DP1; r++) { fp[r] = r == 0 ? 1.0 : 0.0; fpp[r] = xj[r]; }
            const int cs = theta_in_smem;                              // THC: first constant-bank slot of θ
            double th0 = THC ? c_uni_theta[cs] : th[0];
#pragma unroll
            for (int r = 0; r < DP1; r++) s[r] = r == 0 ? th0 : 0.0;
#pragma unroll 4
            for (int k = 1; k < N; k++) {
                double f[DP1];
                f[0] = fma(t0, fp[0], -fpp[0]);
                if (DP1 > 1) f[DP1 > 1 ? 1 : 0] = fma(t0, fp[DP1 > 1 ? 1 : 0], fma(t1, fp[0], -fpp[DP1 > 1 ? 1 : 0]));
                if (DP1 > 2)
                    f[DP1 > 2 ? 2 : 0] = fma(t0, fp[DP1 > 2 ? 2 : 0],
                                             fma(tt1, fp[DP1 > 1 ? 1 : 0], fma(t2, fp[0], -fpp[DP1 > 2 ? 2 : 0])));
                const double tk = THC ? c_uni_theta[cs + k] : th[k];
#pragma unroll
                for (int r = 0; r < DP1; r++) { s[r] = fma(tk, f[r], s[r]); fpp[r] = fp[r]; fp[r] = f[r]; }
            }
        }
#pragma unroll
        for (int r = 0; r < DP1; r++) out[p * DP1 + r] = s[r];
    }
}

// Fast path for large batches, PP points per thread (the thread's points are one grid stride apart: loads and stores stay
// coalesced): θ comes from the constant bank through the uniform datapath, and every LDCU — which holds the scheduler for
// ≈ 2.25 cycles (profiles/micro/dfma_ldcu.cu) — now feeds the fmas of PP points instead of one.
template <int DP1, int PP>
__global__ void __launch_bounds__(256) uni_lincomb_pts_kernel(int N, sk_transform tr, int cs, const double *__restrict__ x, int64_t n,
                                                              double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p0 < n; p0 += stride * PP) {
        double t0[PP], t1[PP], tt1[PP], t2[PP], fp[PP][DP1], fpp[PP][DP1], s[PP][DP1];
#pragma unroll
        for (int v = 0; v < PP; v++) {
            double xj[DP1];
            skd::to_pm1_jet<DP1>(tr, x[min(p0 + v * stride, n - 1)], xj);
            t0[v] = 2.0 * xj[0];
            t1[v] = DP1 > 1 ? 2.0 * xj[DP1 > 1 ? 1 : 0] : 0.0;
            tt1[v] = 2.0 * t1[v];
            t2[v] = DP1 > 2 ? 2.0 * xj[DP1 > 2 ? 2 : 0] : 0.0;
#pragma unroll
            for (int r = 0; r < DP1; r++) { fp[v][r] = r == 0 ? 1.0 : 0.0; fpp[v][r] = xj[r]; }
        }
        const double th0 = c_uni_theta[cs];
#pragma unroll
        for (int v = 0; v < PP; v++)
#pragma unroll
            for (int r = 0; r < DP1; r++) s[v][r] = r == 0 ? th0 : 0.0;
#pragma unroll 4
        for (int k = 1; k < N; k++) {
            const double tk = c_uni_theta[cs + k];
#pragma unroll
            for (int v = 0; v < PP; v++) {
                double f[DP1];
                f[0] = fma(t0[v], fp[v][0], -fpp[v][0]);
                if (DP1 > 1) f[DP1 > 1 ? 1 : 0] = fma(t0[v], fp[v][DP1 > 1 ? 1 : 0], fma(t1[v], fp[v][0], -fpp[v][DP1 > 1 ? 1 : 0]));
                if (DP1 > 2)
                    f[DP1 > 2 ? 2 : 0] = fma(t0[v], fp[v][DP1 > 2 ? 2 : 0],
                                             fma(tt1[v], fp[v][DP1 > 1 ? 1 : 0], fma(t2[v], fp[v][0], -fpp[v][DP1 > 2 ? 2 : 0])));
#pragma unroll
                for (int r = 0; r < DP1; r++) { s[v][r] = fma(tk, f[r], s[v][r]); fpp[v][r] = fp[v][r]; fp[v][r] = f[r]; }
            }
        }
#pragma unroll
        for (int v = 0; v < PP; v++) {
            const int64_t p = p0 + v * stride;
            if (p < n) {
#pragma unroll
                for (int r = 0; r < DP1; r++) out[p * DP1 + r] = s[v][r];
            }
        }
    }
}

// SVector coefficients (nvec > 1), values only: θ is [N][nvec]; VC vectors per pass
template <bool EXACT, int VC>
__global__ void __launch_bounds__(256) uni_lincomb_nvec_kernel(int N, sk_transform tr, const double *__restrict__ theta,
                                                               int nvec, const double *__restrict__ x, int64_t n,
                                                               double *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        double x0 = skd::to_pm1(tr, x[p]);
        double t0 = 2.0 * x0;
        for (int v0 = 0; v0 < nvec; v0 += VC) {
            double s[VC];
            double fp = 1.0, fpp = x0;
#pragma unroll
            for (int v = 0; v < VC; v++) s[v] = (v0 + v < nvec) ? theta[v0 + v] * 1.0 : 0.0;
            for (int k = 1; k < N; k++) {
                double f = EXACT ? (t0 * fp - fpp) : fma(t0, fp, -fpp);
                fpp = fp; fp = f;
                const double *tk = theta + (size_t)k * nvec + v0;
#pragma unroll
                for (int v = 0; v < VC; v++)
                    if (v0 + v < nvec) s[v] = EXACT ? (s[v] + tk[v] * f) : fma(tk[v], f, s[v]);
            }
#pragma unroll
            for (int v = 0; v < VC; v++)
                if (v0 + v < nvec) out[p * nvec + v0 + v] = s[v];
        }
    }
}

// bookkeeping of the two constant-bank coefficient slots (c_uni_theta), shared by all launch_uni<DP1> instantiations
struct UniConstSlots {
    std::mutex mu;
    cudaEvent_t done[64][2] = {};      // per device and slot: recorded after the last kernel that reads the slot
    unsigned next[64] = {};
};
static UniConstSlots &uni_const_slots() { static UniConstSlots s; return s; }

template <int DP1>
static cudaError_t launch_uni(bool exact, int N, const sk_transform &tr, const double *theta, const double *x, int64_t n,
                              double *out, cudaStream_t stream) {
    int threads = 256;
    int64_t want = (n + threads - 1) / threads;
    int grid = (int)std::min<int64_t>(want, (int64_t)sm_count() * 16);
    size_t smem = (size_t)N * sizeof(double);
    int in_smem = smem <= 48 * 1024;
    if (!in_smem) smem = 0;
    static const int use_const = [] { const char *e = getenv("SK_THETA_CONST"); return e ? atoi(e) : 1; }();
    int dev = 0;
    cudaGetDevice(&dev);
    // (small batches stay on the shared-memory path: the symbol copy and the slot event cost ≈ 10 µs of launch latency)
    if (!exact && use_const && N <= kUniConstMax && dev < 64 && n >= ((int64_t)1 << 17)) {
        // ONE set of slot bookkeeping per device for every derivative order: all instantiations of this template write
        // the same __constant__ array, so the slot counter and the "last reader" events must be shared between them
        UniConstSlots &cs = uni_const_slots();
        std::lock_guard<std::mutex> lk(cs.mu);
        cudaEvent_t (&done)[64][2] = cs.done;
        const int slot = (int)(cs.next[dev]++ & 1u);
        cudaError_t e = cudaSuccess;
        if (!done[dev][slot]) e = cudaEventCreateWithFlags(&done[dev][slot], cudaEventDisableTiming);
        else e = cudaStreamWaitEvent(stream, done[dev][slot], 0);       // the launch that last read this slot
        if (e != cudaSuccess) return e;
        e = cudaMemcpyToSymbolAsync(c_uni_theta, theta, (size_t)N * sizeof(double), (size_t)slot * kUniConstMax * sizeof(double), cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
        // SK_UNI_PP = points per thread of the constant-bank kernel (experiments).  Measured at 2²⁶ points, cfg 2 / cfg 3:
        // one point 43 / 21.8 Gpoints/s, two 39.6 / 20.4, four 38.4 / 20.3 — unlike the Smolyak values kernel this one gains
        // nothing from sharing a coefficient between points (the recurrence, not θ, is most of its work): one point stays
        static const int pp = [] { const char *e2 = getenv("SK_UNI_PP"); return e2 ? atoi(e2) : 1; }();
        constexpr int D1 = DP1 <= 3 ? DP1 : 1;
        if (pp == 2 || pp == 4) {
            const int g2 = (int)std::min<int64_t>((want + pp - 1) / pp, (int64_t)sm_count() * 16);
            if (pp == 2) uni_lincomb_pts_kernel<D1, 2><<<g2, threads, 0, stream>>>(N, tr, slot * kUniConstMax, x, n, out);
            else uni_lincomb_pts_kernel<D1, 4><<<g2, threads, 0, stream>>>(N, tr, slot * kUniConstMax, x, n, out);
        } else
        uni_lincomb_kernel<D1, false, true><<<grid, threads, 0, stream>>>(N, tr, theta, slot * kUniConstMax, x, n, out);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return cudaEventRecord(done[dev][slot], stream);
    }
    if (exact) uni_lincomb_kernel<DP1, true><<<grid, threads, smem, stream>>>(N, tr, theta, in_smem, x, n, out);
    else uni_lincomb_kernel<(DP1 <= 3 ? DP1 : 1), false><<<grid, threads, smem, stream>>>(N, tr, theta, in_smem, x, n, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t uni_lincomb(int N, const sk_transform &tr, int order, bool exact, const double *theta, int nvec,
                        const double *x, int64_t n, double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (nvec > 1) {
        int threads = 256;
        int grid = (int)std::min<int64_t>((n + threads - 1) / threads, (int64_t)sm_count() * 16);
        if (exact) uni_lincomb_nvec_kernel<true, 8><<<grid, threads, 0, stream>>>(N, tr, theta, nvec, x, n, out);
        else uni_lincomb_nvec_kernel<false, 8><<<grid, threads, 0, stream>>>(N, tr, theta, nvec, x, n, out);
        count_launch();
        return cudaGetLastError();
    }
    if (order > 2) exact = true;   // the contracted recurrence is written out for D <= 2 only
    switch (order) {
    case 0: return launch_uni<1>(exact, N, tr, theta, x, n, out, stream);
    case 1: return launch_uni<2>(exact, N, tr, theta, x, n, out, stream);
    case 2: return launch_uni<3>(exact, N, tr, theta, x, n, out, stream);
    case 3: return launch_uni<4>(exact, N, tr, theta, x, n, out, stream);
    case 4: return launch_uni<5>(exact, N, tr, theta, x, n, out, stream);
    case 5: return launch_uni<6>(exact, N, tr, theta, x, n, out, stream);
    }
    return cudaErrorInvalidValue;
}

// -------------------------------------------------------------------------------------------
// K3 (univariate): collocation matrix.  One point per thread, basis functions produced in order by
// the exact recurrence; each block stages a tile of P points × E doubles per basis function in
// shared memory (triple buffered) and one elected thread streams it out with the bulk-copy engine.
// Falls back to plain coalesced stores when the tile is partial or the destination is not 16-byte
// aligned.  out[(j*ld + i)*E + r].
// -------------------------------------------------------------------------------------------
template <int DP1>
__global__ void __launch_bounds__(128) uni_basis_kernel(int N, sk_transform tr, const double *__restrict__ x, int64_t n,
                                                        double *__restrict__ out, int64_t ld) {
    constexpr int P = 128;
    __shared__ __align__(128) double stage[3][P * DP1];
    const int tid = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * P;
    const int64_t p = p0 + tid;
    const bool full = p0 + P <= n;
    const bool live = p < n;
    Jet<DP1> X, fp = skd::jet_one<DP1>(), fpp, f = fp;
    {
        double xj[DP1];
        skd::to_pm1_jet<DP1>(tr, live ? x[p] : 0.0, xj);
#pragma unroll
        for (int r = 0; r < DP1; r++) { X.c[r] = xj[r]; fpp.c[r] = xj[r]; }
    }
    for (int j = 0; j < N; j++) {
        if (j > 0) { f = skd::cheb_step_exact<DP1>(X, fp, fpp); fpp = fp; fp = f; }
        double *g = out + ((size_t)j * ld + p0) * DP1;
        const bool bulk = full && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);     // block-uniform
        if (bulk) {
            double *st = stage[j % 3];
#pragma unroll
            for (int r = 0; r < DP1; r++) st[tid * DP1 + r] = f.c[r];
            bulk_store_fence();
            __syncthreads();
            if (tid == 0) { bulk_store(g, st, P * DP1 * 8); bulk_store_wait_read_1(); }
        } else if (live) {
#pragma unroll
            for (int r = 0; r < DP1; r++) g[tid * DP1 + r] = f.c[r];
        }
    }
    if (tid == 0) bulk_store_wait_all();
}

cudaError_t uni_basis_at(int N, const sk_transform &tr, int order, const double *x, int64_t n, double *out, int64_t ld,
                         cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + 127) / 128);
    switch (order) {
    case 0: uni_basis_kernel<1><<<grid, 128, 0, stream>>>(N, tr, x, n, out, ld); break;
    case 1: uni_basis_kernel<2><<<grid, 128, 0, stream>>>(N, tr, x, n, out, ld); break;
    case 2: uni_basis_kernel<3><<<grid, 128, 0, stream>>>(N, tr, x, n, out, ld); break;
    case 3: uni_basis_kernel<4><<<grid, 128, 0, stream>>>(N, tr, x, n, out, ld); break;
    case 4: uni_basis_kernel<5><<<grid, 128, 0, stream>>>(N, tr, x, n, out, ld); break;
    case 5: uni_basis_kernel<6><<<grid, 128, 0, stream>>>(N, tr, x, n, out, ld); break;
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// K2' / K3 (Smolyak): direct form in the reference's operation order.
//   * d univariate tables of H jets per point (smolyak_api.jl:89-93) live in shared memory,
//     laid out [dim][index][order][point] so a warp's reads are conflict free;
//   * the index table (K×d uint16, traversal order, built once by the plan) replaces the
//     per-point `__inc` state machine (smolyak_traversal.jl:211-231); it is warp-uniform;
//   * each component is the left-to-right product over dimensions (derivatives.jl:502-511), and
//     the linear combination is the sequential s = s + θ_k·b_k (generic_api.jl:114-128).
// -------------------------------------------------------------------------------------------
struct DirectParams {
    int d, H, ncomp, E;          // ncomp == 0: plain point (E = 1)
    int64_t K, n, ld;
    const uint16_t *idx;
    const uint4 *fac;            // values-only plans: per term, the factors with index != 1 (8 × uint16: dim << 12 | index-1,
                                 // 0xFFFF = none), in dimension order; multiplying by T_0 = 1.0 is exact, so it is skipped
    const double *theta;
    int nvec;
    const double *x;
    double *out;
    int64_t k_chunk;             // BASIS: terms per blockIdx.y
    int no_bulk;                 // BASIS: plain coalesced stores instead of staged bulk copies
    sk_transform tr[SK_MAX_DIM];
    int8_t orders[SK_MAX_COMP][SK_MAX_DIM];
};

// tables of one tile: thread (tid, ty) builds the dimensions j ≡ ty (mod ny) of point tid
template <int EMAX>
__device__ __forceinline__ void build_tables(const DirectParams &q, const double *xp, bool live, double *tab, int P, int tid, int ty, int ny) {
    for (int j = ty; j < q.d; j += ny) {
        double xj[EMAX];
        skd::to_pm1_jet<EMAX>(q.tr[j], live ? xp[j] : 0.0, xj);
        Jet<EMAX> X, fp = skd::jet_one<EMAX>(), fpp, f = fp;
#pragma unroll
        for (int r = 0; r < EMAX; r++) { X.c[r] = xj[r]; fpp.c[r] = xj[r]; }
        for (int i = 0; i < q.H; i++) {
            if (i > 0) { f = skd::cheb_step_exact<EMAX>(X, fp, fpp); fpp = fp; fp = f; }
#pragma unroll
            for (int r = 0; r < EMAX; r++) tab[(((size_t)j * q.H + i) * EMAX + r) * P + tid] = f.c[r];
        }
    }
}

// product over dimensions of component c at term ik (left to right)
template <int EMAX>
__device__ __forceinline__ double term_product(const DirectParams &q, const uint16_t *ik, int c, const double *tab, int P, int tid) {
    int o = q.ncomp ? q.orders[c][0] : 0;
    double b = tab[(((size_t)0 * q.H + (ik[0] - 1)) * EMAX + o) * P + tid];
    for (int j = 1; j < q.d; j++) {
        o = q.ncomp ? q.orders[c][j] : 0;
        b = b * tab[(((size_t)j * q.H + (ik[j] - 1)) * EMAX + o) * P + tid];
    }
    return b;
}

// values only: the same left-to-right product over the factors that are not T_0 = 1 (x·1.0 == x bit for bit)
__device__ __forceinline__ double term_product_compact(const DirectParams &q, int64_t k, const double *tab, int P, int tid) {
    const uint4 w = __ldg(q.fac + k);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    double b = 1.0;
    bool first = true;
#pragma unroll
    for (int f = 0; f < 8; f++) {
        const uint32_t e = (ws[f >> 1] >> (16 * (f & 1))) & 0xFFFFu;
        if (e == 0xFFFFu) break;
        const double t = tab[((size_t)(e >> 12) * q.H + (e & 0xFFFu)) * P + tid];
        b = first ? t : b * t;
        first = false;
    }
    return b;
}
template <int EMAX>
__device__ __forceinline__ double term_value(const DirectParams &q, int64_t k, int c, const double *tab, int P, int tid) {
    if (EMAX == 1 && q.fac) return term_product_compact(q, k, tab, P, tid);
    return term_product<EMAX>(q, q.idx + (size_t)k * q.d, c, tab, P, tid);
}

// Block = P points (x) × ny component lanes (y).  The sequential sum s = s + θ_k·b_k of one component of one
// point is one thread's job (the reference's order), but components — and coefficient vectors — are
// independent, so they run on different threads: the tables limit a CTA to P = 32..128 points, and this is
// what fills the SM with warps.  A warp is 32 consecutive points of one component: table reads stay conflict
// free and the index table / θ reads stay warp-uniform.
template <int EMAX>
__global__ void smolyak_direct_lincomb_kernel(const __grid_constant__ DirectParams q) {
    extern __shared__ double tab[];
    const int P = blockDim.x, tid = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
    const int nc = q.ncomp ? q.ncomp : 1;
    for (int64_t p0 = (int64_t)blockIdx.x * P; p0 < q.n; p0 += (int64_t)gridDim.x * P) {
        const int64_t p = p0 + tid;
        const bool live = p < q.n;
        __syncthreads();                                   // the previous tile's tables are no longer read
        build_tables<EMAX>(q, q.x + (size_t)(live ? p : 0) * q.d, live, tab, P, tid, ty, ny);
        __syncthreads();
        if (q.nvec == 1) {
            for (int c = ty; c < nc; c += ny) {
                double s = 0.0;
                for (int64_t k = 0; k < q.K; k++) {
                    const double b = term_value<EMAX>(q, k, c, tab, P, tid);
                    const double th = q.theta[k];
                    s = k == 0 ? th * b : s + th * b;      // generic_api.jl:114-128: mul then add, never fused
                }
                if (live) q.out[(size_t)p * nc + c] = s;
            }
        } else {
            // SVector coefficients: plain points only (checked by the API layer); 8 vectors per pass
            for (int v0 = ty * 8; v0 < q.nvec; v0 += ny * 8) {
                double s[8];
                for (int64_t k = 0; k < q.K; k++) {
                    const double b = term_value<EMAX>(q, k, 0, tab, P, tid);
                    const double *tk = q.theta + (size_t)k * q.nvec + v0;
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        if (v0 + v < q.nvec) s[v] = k == 0 ? tk[v] * b : s[v] + tk[v] * b;
                }
                if (live)
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        if (v0 + v < q.nvec) q.out[(size_t)p * q.nvec + v0 + v] = s[v];
            }
        }
    }
}

// basis_at / collocation: blockIdx.x = tile of P points, blockIdx.y = chunk of terms; block = P points (x) × ny term
// lanes (y): lane ty produces the terms k0 + ty, k0 + ty + ny, … of its chunk.  (The tables allow 32..128 points per
// CTA; the term lanes are what fills the SM with warps — output is 8·E bytes per ≈ d multiplies, HBM-write-bound.)
// Each lane stages its column segment (P·E doubles, contiguous in the column-major result) in its own ring of 3
// buffers and one elected thread streams it out with the bulk-copy engine; lanes synchronise on their own named
// barrier.  Partial tiles / unaligned destinations use plain coalesced stores.
template <int EMAX>
__global__ void smolyak_direct_basis_kernel(const __grid_constant__ DirectParams q) {
    extern __shared__ double smem[];
    const int P = blockDim.x, tid = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
    const int E = q.E;
    double *tab = smem;
    double *stage = smem + (size_t)q.d * q.H * EMAX * P + (size_t)ty * 3 * P * E;
    const int64_t p0 = (int64_t)blockIdx.x * P, p = p0 + tid;
    const bool live = p < q.n, full = p0 + P <= q.n;
    build_tables<EMAX>(q, q.x + (size_t)(live ? p : 0) * q.d, live, tab, P, tid, ty, ny);
    __syncthreads();
    const int64_t k0 = (int64_t)blockIdx.y * q.k_chunk, k1 = min(q.K, k0 + q.k_chunk);
    int buf = 0;
    for (int64_t k = k0 + ty; k < k1; k += ny) {
        double *g = q.out + ((size_t)k * q.ld + p0) * E;
        const bool bulk = !q.no_bulk && full && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) && (((size_t)P * E * 8) % 16 == 0);
        double *st = stage + (size_t)buf * P * E;
        for (int c = 0; c < E; c++) {
            double b = term_value<EMAX>(q, k, c, tab, P, tid);
            if (bulk) st[tid * E + c] = b;
            else if (live) __stcs(g + (size_t)tid * E + c, b);
        }
        if (bulk) {
            bulk_store_fence();
            asm volatile("bar.sync %0, %1;" ::"r"(ty + 1), "r"(P) : "memory");      // this lane's P threads
            if (tid == 0) { bulk_store(g, st, (uint32_t)(P * E * 8)); bulk_store_wait_read_1(); }
            buf = buf == 2 ? 0 : buf + 1;
            // the buffer written two terms from now was handed to the copy engine one term ago and its read has
            // completed (wait_group.read 1) before thread 0 passes the next barrier
        }
    }
    if (tid == 0) bulk_store_wait_all();
}

static void fill_direct_params(const sk_plan &plan, DirectParams &q) {
    q.d = plan.basis.d; q.H = (int)plan.H; q.ncomp = plan.ncomp; q.E = plan.ncomp ? plan.ncomp : 1;
    q.K = plan.K; q.idx = plan.d_idx; q.fac = reinterpret_cast<const uint4 *>(plan.d_fac);
    for (int j = 0; j < q.d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    for (int c = 0; c < plan.ncomp; c++) for (int j = 0; j < q.d; j++) q.orders[c][j] = (int8_t)plan.orders[c][j];
}
static int plan_emax(const sk_plan &plan) {
    int e = 1;
    for (int j = 0; j < plan.basis.d; j++) e = std::max(e, plan.deg[j] + 1);
    return e <= 3 ? e : 6;
}
// points per block such that the tables (+ staging) fit in shared memory
static int pick_points_per_block(size_t bytes_per_point, size_t limit) {
    for (int P = 128; P >= 32; P >>= 1) if (bytes_per_point * P <= limit) return P;
    return 0;
}

template <typename K>
static cudaError_t set_smem(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// The direct kernels keep every dimension's full Chebyshev table (d·ℓ(B)·(order + 1) doubles) per point in shared
// memory, 32 points per CTA at least (one warp per named barrier): bases beyond that are refused up front (sk_api.cu).
static constexpr size_t kDirectSmemLimit = 224 * 1024;
bool smolyak_direct_fits(const sk_plan &plan, bool for_basis) {
    const size_t tab_point = (size_t)plan.basis.d * (size_t)plan.H * plan_emax(plan) * sizeof(double);
    const size_t extra = for_basis ? 3 * (size_t)plan.E * sizeof(double) : 0;
    return (tab_point + extra) * 32 <= kDirectSmemLimit;
}

cudaError_t smolyak_direct_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n,
                                   double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    DirectParams q{};
    fill_direct_params(plan, q);
    q.theta = theta; q.nvec = nvec; q.x = x; q.out = out; q.n = n; q.ld = n; q.k_chunk = q.K;
    const int emax = plan_emax(plan);
    size_t per_point = (size_t)q.d * q.H * emax * sizeof(double);
    // independent sums per point: components, or groups of 8 coefficient vectors
    const int lanes_wanted = nvec > 1 ? (nvec + 7) / 8 : (q.ncomp ? q.ncomp : 1);
    int P = pick_points_per_block(per_point, kDirectSmemLimit);
    if (!P) return cudaErrorInvalidConfiguration;
    if (lanes_wanted == 1 && n <= (int64_t)P * sm_count() / 2) while (P > 32 && n <= (int64_t)(P / 2) * sm_count()) P >>= 1;   // few points: more, smaller tiles
    int ny = std::max(1, std::min(lanes_wanted, 1024 / P));
    size_t smem = per_point * P;
    int grid = (int)std::min<int64_t>((n + P - 1) / P, (int64_t)sm_count() * 8);
    cudaError_t e = cudaSuccess;
#define SK_LAUNCH_DL(EM)                                                                 \
    do {                                                                                 \
        cudaFuncAttributes fa;                                                           \
        e = cudaFuncGetAttributes(&fa, smolyak_direct_lincomb_kernel<EM>);               \
        if (e == cudaSuccess) ny = std::max(1, std::min(ny, fa.maxThreadsPerBlock / P));   /* register file */ \
        if (e == cudaSuccess) e = set_smem(smolyak_direct_lincomb_kernel<EM>, smem);     \
        if (e == cudaSuccess) smolyak_direct_lincomb_kernel<EM><<<grid, dim3(P, ny), smem, stream>>>(q); \
    } while (0)
    if (emax == 1) SK_LAUNCH_DL(1); else if (emax == 2) SK_LAUNCH_DL(2); else if (emax == 3) SK_LAUNCH_DL(3); else SK_LAUNCH_DL(6);
#undef SK_LAUNCH_DL
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t smolyak_direct_basis_at(const sk_plan &plan, const double *x, int64_t n, double *out, int64_t ld,
                                    cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    DirectParams q{};
    fill_direct_params(plan, q);
    q.theta = nullptr; q.nvec = 1; q.x = x; q.out = out; q.n = n; q.ld = ld;
    // E == 1: a warp's 32 values are already contiguous in the column, plain streaming stores win (1.73 vs 1.53 TB/s);
    // E > 1: the staged bulk copy makes the interleaved components one contiguous segment (189 vs 174 GB/s).
    // SK_BASIS_NO_BULK=0/1 forces either (experiments)
    static const int force = [] { const char *e = getenv("SK_BASIS_NO_BULK"); return e ? atoi(e) : -1; }();
    q.no_bulk = force >= 0 ? force : (q.E == 1);
    const int emax = plan_emax(plan);
    const size_t tab_point = (size_t)q.d * q.H * emax * sizeof(double);
    const size_t limit = kDirectSmemLimit;
    int P = 0, ny = 1;
    // points per CTA: P = 64 (two warps per lane) with as many term lanes as fit (<= 8: 15 named barriers, 512+
    // threads); fall back to fewer points when the tables are large
    for (int cand : {64, 32, 128}) {
        for (int lanes = 8; lanes >= 1; lanes >>= 1) {
            if (cand * lanes > 1024) continue;
            if ((tab_point + 3 * (size_t)lanes * q.E * sizeof(double)) * cand <= limit) { P = cand; ny = lanes; break; }
        }
        if (P) break;
    }
    if (!P) return cudaErrorInvalidConfiguration;
    const size_t smem = (tab_point + 3 * (size_t)ny * q.E * sizeof(double)) * P;
    int64_t tiles = (n + P - 1) / P;
    // enough blocks to fill the machine: split the K terms when there are few point tiles
    int64_t want_chunks = std::max<int64_t>(1, ((int64_t)sm_count() * 2 + tiles - 1) / tiles);
    int64_t chunks = std::min<int64_t>(want_chunks, std::max<int64_t>(1, q.K / (16 * ny)));
    chunks = std::min<int64_t>(chunks, 65535);
    q.k_chunk = (q.K + chunks - 1) / chunks;
    dim3 grid((unsigned)tiles, (unsigned)((q.K + q.k_chunk - 1) / q.k_chunk));
    cudaError_t e;
#define SK_LAUNCH_DB(EM)                                                   \
    do {                                                                   \
        cudaFuncAttributes fa;                                             \
        e = cudaFuncGetAttributes(&fa, smolyak_direct_basis_kernel<EM>);   \
        if (e == cudaSuccess) while (ny > 1 && P * ny > fa.maxThreadsPerBlock) ny >>= 1;   /* register file */ \
        if (e == cudaSuccess) e = set_smem(smolyak_direct_basis_kernel<EM>, smem);               \
        if (e == cudaSuccess) smolyak_direct_basis_kernel<EM><<<grid, dim3(P, ny), smem, stream>>>(q); \
    } while (0)
    if (emax == 1) SK_LAUNCH_DB(1); else if (emax == 2) SK_LAUNCH_DB(2); else if (emax == 3) SK_LAUNCH_DB(3); else SK_LAUNCH_DB(6);
#undef SK_LAUNCH_DB
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// batched coordinate transformations (transformations.jl:112-139)
// -------------------------------------------------------------------------------------------
struct TrParams { int d; sk_transform tr[SK_MAX_DIM]; };

template <int DP1>
__global__ void transform_to_kernel(const __grid_constant__ TrParams q, const double *__restrict__ y, int64_t total,
                                    double *__restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(e % q.d);
        double xj[DP1];
        skd::to_pm1_jet<DP1>(q.tr[j], y[e], xj);
#pragma unroll
        for (int r = 0; r < DP1; r++) out[e * DP1 + r] = xj[r];
    }
}
__global__ void transform_from_kernel(const __grid_constant__ TrParams q, const double *__restrict__ x, int64_t total,
                                      double *__restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = skd::from_pm1(q.tr[(int)(e % q.d)], x[e]);
}

cudaError_t transform_to(const sk_transform *tr, int d, int order, const double *y, int64_t n, double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    TrParams q{};
    q.d = d;
    for (int j = 0; j < d; j++) q.tr[j] = tr[j];
    int64_t total = n * d;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    switch (order) {
    case 0: transform_to_kernel<1><<<grid, 256, 0, stream>>>(q, y, total, out); break;
    case 1: transform_to_kernel<2><<<grid, 256, 0, stream>>>(q, y, total, out); break;
    case 2: transform_to_kernel<3><<<grid, 256, 0, stream>>>(q, y, total, out); break;
    case 3: transform_to_kernel<4><<<grid, 256, 0, stream>>>(q, y, total, out); break;
    case 4: transform_to_kernel<5><<<grid, 256, 0, stream>>>(q, y, total, out); break;
    case 5: transform_to_kernel<6><<<grid, 256, 0, stream>>>(q, y, total, out); break;
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}
cudaError_t transform_from(const sk_transform *tr, int d, const double *x, int64_t n, double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    TrParams q{};
    q.d = d;
    for (int j = 0; j < d; j++) q.tr[j] = tr[j];
    int64_t total = n * d;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    transform_from_kernel<<<grid, 256, 0, stream>>>(q, x, total, out);
    count_launch();
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// FP64 pipe microbenchmarks: the roofline denominators MEASURED_PEAKS.json does not carry.
//   which == 0: DFMA  — 16 independent fma chains per thread
//   which == 1: DMMA  — mma.sync.aligned.m8n8k4.row.col.f64, 8 independent accumulator tiles/warp
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double *sink) {
    double a[16];
    const double m = 1.0000001, c = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double *sink) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

// which == 3: DFMA with THREE distinct register operands per instruction (a[i] = fma(a[i], b[i], c[i])).  The peak
// kernel above has two loop-invariant operands (operand-reuse cache); real evaluation code mostly does not, and the
// register file cannot deliver three fresh 64-bit operands at the pipe's issue rate: 20.6 vs 36 TFLOP/s on B200
// (profiles/micro/dfma_operands.cu).  Reported next to the peak as the practical ceiling of such code.
__global__ void __launch_bounds__(256) dfma3_peak_kernel(int iters, double *sink) {
    double a[16], b[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6; b[i] = 1.0 - (i + threadIdx.x) * 1e-9; c[i] = 1e-9 * (i + 1 + threadIdx.x); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = __longlong_as_double(__double_as_longlong(b[i]) ^ (long long)(it & 1));   // keeps b live, no FP64 work
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i] + b[i] + c[i];
    if (s == 123.456) sink[0] = s;
}

// which == 2: even warps run the DFMA chains, odd warps the DMMA tiles — do the two share one pipe?
__global__ void __launch_bounds__(256) dmix_peak_kernel(int iters, double *sink) {
    double s = 0;
    if ((threadIdx.x >> 5) & 1) {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
        double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    } else {
        double a[16];
        const double m = 1.0000001, c = 1e-9 * (threadIdx.x + 1);
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) s += a[i];
    }
    if (s == 123.456) sink[0] = s;
}

cudaError_t fp64_peak_launch(int which, int iters, double *sink, double *flops, cudaStream_t stream) {
    int blocks = sm_count() * 8, threads = 256;
    if (which == 3) {
        dfma3_peak_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
        *flops = 2.0 * 16 * (double)iters * blocks * threads;
    } else if (which == 2) {
        dmix_peak_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
        *flops = (2.0 * 16 * 32 + 2.0 * 8 * 8 * 4 * 8) * (double)iters * blocks * (threads / 64);
    } else if (which == 0) {
        dfma_peak_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
        *flops = 2.0 * 16 * (double)iters * blocks * threads;
    } else {
        dmma_peak_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
        *flops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * blocks * (threads / 32);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace skk

extern "C" int64_t sk_launch_count(void) { return skk::launch_count(); }
