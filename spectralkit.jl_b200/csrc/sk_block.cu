// sm_100a kernels, part 4 — K4: Smolyak linear_combination with a K×nvec coefficient BLOCK
// (SVector coefficients: src/derivatives.jl:15,19 with src/generic_api.jl:114-128; the several-
// policy-functions caller of src/experimental.jl:228-233).  out[n][nvec] = A[n][K] · Θ[K][nvec], where
// A (the collocation matrix, generic_api.jl:217-227) is never materialised: its tiles are generated
// on chip and contracted with FP64 tensor-core tiles (DMMA, mma.sync.m8n8k4.f64).
//
// One persistent CTA per SM, three warp roles connected by an mbarrier ring of NS stages; a stage
// holds KC = 16 consecutive basis terms:
//   * PRODUCER warps — one POINT per thread.  The column-major traversal (smolyak_traversal.jl:211-231)
//     only ever increments one index and resets the lower ones to 1, so every dimension's T_i is
//     carried by the three-term recurrence (chebyshev.jl:60-70) in registers and the term is
//     T_{i_c}·(product of the dimensions above c): 1 fma + 1 mul per term per point, no table at all.
//     The plan's "advance program" (one byte per term: which dimension moved) replaces `__inc`; it is
//     uniform over points, so the switch on it never diverges.  Values go to smem as A[k][point].
//   * LOADER warp — streams the KC×NVT slab of Θ (rows are contiguous in the caller's layout) with
//     the bulk-copy engine (cp.async.bulk global→shared, mbarrier complete_tx), one copy per row into
//     rows padded to NVT+4 doubles so B-fragment loads are bank-conflict free.  Θ (159 MB at cfg 5)
//     is read by all CTAs in the same order, i.e. once from HBM and 147 times from L2.
//   * CONSUMER warps (8) — warp tile (PT/WM)×(NVT/WN), accumulators in registers, per k4 step
//     MT + NT fragment LDS.64 feed MT·NT DMMAs.
//
// Algorithmic work: 2·K·nvec flop per point in the contraction + 3·K in the producer.
// Summation order differs from the reference (k4-blocked, fma inside the tensor core): SK_MODE_FAST
// tolerance; SK_MODE_EXACT requests are served by the direct kernel (sk_kernels.cu).
#include <algorithm>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

namespace {

constexpr int kKC = 16;              // terms per pipeline stage
constexpr int kConsumerWarps = 8;

struct BlockParams {
    const double *x;                 // [n][d]
    const double *theta;             // [K][nvec]
    double *out;                     // [n][nvec]
    const uint8_t *adv;              // advance program, padded to a multiple of kKC
    int64_t n, K;
    int d, nvec;
    int n_vtiles;
    int64_t n_ptiles;
    int bulk_ok;                     // Θ rows can be bulk-copied (16-byte aligned base, even nvec)
    sk_transform tr[SK_MAX_DIM];
};

// ---- mbarrier / bulk-copy primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n"
        " bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- producer state: per-dimension recurrence (T_cur, T_prev, 2x) and the product of the dimensions above
template <int DMAX>
struct Walk {
    double T[DMAX], Tp[DMAX], X2[DMAX], Q[DMAX];
    // dimension C (0-based) moves to its next index, the lower ones return to index 1 (T_0 = 1)
    template <int C>
    __device__ __forceinline__ double advance() {
        const double tn = fma(X2[C], T[C], -Tp[C]);      // chebyshev.jl:65-70; the first step gives 2x·1 − x = x
        Tp[C] = T[C];
        T[C] = tn;
        const double a = tn * Q[C];
#pragma unroll
        for (int l = 0; l < C; l++) { T[l] = 1.0; Tp[l] = 0.5 * X2[l]; Q[l] = a; }
        return a;
    }
    __device__ __forceinline__ double step(uint32_t c) {
        switch (c) {
#define SK_ADV(C) case C: if constexpr (C < DMAX) return advance<C < DMAX ? C : 0>(); else return 0.0;
            SK_ADV(0) SK_ADV(1) SK_ADV(2) SK_ADV(3) SK_ADV(4) SK_ADV(5) SK_ADV(6) SK_ADV(7)
            SK_ADV(8) SK_ADV(9) SK_ADV(10) SK_ADV(11) SK_ADV(12) SK_ADV(13) SK_ADV(14) SK_ADV(15)
#undef SK_ADV
        case 0xFE: return 1.0;                            // first term: every index is 1
        default: return 0.0;                              // padding past K
        }
    }
};

template <int PT, int NVT>
struct BlockSmem {
    static constexpr int LDP = PT + 4;                    // ≡ 4 (mod 16) doubles: conflict-free A fragments
    static constexpr int LDB = NVT + 4;                   // same for B fragments; rows stay 16-byte aligned
    static constexpr size_t a_bytes = (size_t)kKC * LDP * 8;
    static constexpr size_t b_bytes = (size_t)kKC * LDB * 8;
    static constexpr size_t stage_bytes = a_bytes + b_bytes;
    static constexpr int NS = (int)((200 * 1024) / stage_bytes) < 6 ? (int)((200 * 1024) / stage_bytes) : 6;
    static constexpr size_t bars = (size_t)NS * stage_bytes;
    static constexpr size_t total = bars + 2 * NS * 8;
};

template <int DMAX, int PT, int NVT, int WM, int WN>
__global__ void __launch_bounds__((kConsumerWarps + PT / 32 + 1) * 32, 1) smolyak_block_kernel(const __grid_constant__ BlockParams q) {
    static_assert(WM * WN == kConsumerWarps, "eight consumer warps");
    using L = BlockSmem<PT, NVT>;
    constexpr int NS = L::NS, LDP = L::LDP, LDB = L::LDB;
    constexpr int PW = PT / 32;                           // producer warps
    constexpr int WTM = PT / WM, WTN = NVT / WN;          // warp tile
    constexpr int MT = WTM / 8, NT = WTN / 8;
    static_assert(WTM % 8 == 0 && WTN % 8 == 0 && MT >= 1 && NT >= 1, "warp tile must be made of 8x8 DMMA tiles");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::bars);
    uint64_t *empty = full + NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], PT + 1); mbar_init(&empty[s], kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_chunks = (q.K + kKC - 1) / kKC;
    const int64_t n_tiles = q.n_ptiles * q.n_vtiles;
    uint32_t it = 0;                                      // pipeline iteration counter (continues across tiles)

    if (warp < kConsumerWarps) {
        // ------------------------------------------------------------------ consumers
        const int wm = warp / WN, wn = warp % WN;
        const int g = lane >> 2, t = lane & 3;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t p0 = (tile / q.n_vtiles) * PT;
            const int v0 = (int)(tile % q.n_vtiles) * NVT;
            double acc[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int64_t ch = 0; ch < n_chunks; ch++, it++) {
                const int s = it % NS;
                mbar_wait(&full[s], (it / NS) & 1);
                const double *As = reinterpret_cast<const double *>(smem + (size_t)s * L::stage_bytes) + wm * WTM + g;
                const double *Bs = reinterpret_cast<const double *>(smem + (size_t)s * L::stage_bytes + L::a_bytes) + wn * WTN + g;
#pragma unroll
                for (int kk = 0; kk < kKC / 4; kk++) {
                    double a[MT], b[NT];
#pragma unroll
                    for (int i = 0; i < MT; i++) a[i] = As[(kk * 4 + t) * LDP + i * 8];
#pragma unroll
                    for (int j = 0; j < NT; j++) b[j] = Bs[(kk * 4 + t) * LDB + j * 8];
#pragma unroll
                    for (int i = 0; i < MT; i++)
#pragma unroll
                        for (int j = 0; j < NT; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            // epilogue: C fragment (row g, columns 2t, 2t+1) → out[p][v]
            const bool vec2 = (q.nvec & 1) == 0 && (reinterpret_cast<uintptr_t>(q.out) & 15) == 0;
#pragma unroll
            for (int i = 0; i < MT; i++) {
                const int64_t p = p0 + wm * WTM + i * 8 + g;
                if (p >= q.n) continue;
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    const int v = v0 + wn * WTN + j * 8 + 2 * t;
                    double *o = q.out + (size_t)p * q.nvec + v;
                    if (vec2 && v + 1 < q.nvec) *reinterpret_cast<double2 *>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
                    else {
                        if (v < q.nvec) o[0] = acc[i][j][0];
                        if (v + 1 < q.nvec) o[1] = acc[i][j][1];
                    }
                }
            }
        }
    } else if (warp < kConsumerWarps + PW) {
        // ------------------------------------------------------------------ producers: one point per thread
        const int pt = threadIdx.x - kConsumerWarps * 32;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t p0 = (tile / q.n_vtiles) * PT;
            const int64_t p = min(p0 + pt, q.n - 1);          // threads past the end recompute the last point
            Walk<DMAX> w;
#pragma unroll
            for (int j = 0; j < DMAX; j++) {
                double xj = 0.0;
                if (j < q.d) xj = skd::to_pm1(q.tr[j], q.x[(size_t)p * q.d + j]);
                w.X2[j] = 2.0 * xj; w.T[j] = 1.0; w.Tp[j] = xj; w.Q[j] = 1.0;
            }
            uint4 prog = __ldg(reinterpret_cast<const uint4 *>(q.adv));
            for (int64_t ch = 0; ch < n_chunks; ch++, it++) {
                const int s = it % NS;
                const uint4 cur = prog;
                if (ch + 1 < n_chunks) prog = __ldg(reinterpret_cast<const uint4 *>(q.adv) + ch + 1);
                mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
                double *As = reinterpret_cast<double *>(smem + (size_t)s * L::stage_bytes) + pt;
                const uint32_t words[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
                for (int r = 0; r < kKC; r++) As[r * LDP] = w.step((words[r >> 2] >> (8 * (r & 3))) & 0xFFu);
                mbar_arrive(&full[s]);
            }
        }
    } else {
        // ------------------------------------------------------------------ loader: Θ slabs
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int v0 = (int)(tile % q.n_vtiles) * NVT;
            const int cols = min(NVT, q.nvec - v0);
            const bool bulk = q.bulk_ok && (cols & 1) == 0;       // v0 is a multiple of NVT (even): rows stay 16-byte aligned
            for (int64_t ch = 0; ch < n_chunks; ch++, it++) {
                const int s = it % NS;
                mbar_wait(&empty[s], ((it / NS) & 1) ^ 1);
                double *Bs = reinterpret_cast<double *>(smem + (size_t)s * L::stage_bytes + L::a_bytes);
                const int64_t k0 = ch * kKC;
                if (bulk) {
                    if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(kKC * cols * 8));
                    __syncwarp();
                    if (lane < kKC) {
                        const int64_t k = min(k0 + lane, q.K - 1);   // rows past K: their A entries are exactly 0
                        bulk_load(Bs + lane * LDB, q.theta + (size_t)k * q.nvec + v0, (uint32_t)(cols * 8), &full[s]);
                    }
                } else {
                    for (int r = 0; r < kKC; r++) {
                        const int64_t k = min(k0 + r, q.K - 1);
                        const double *src = q.theta + (size_t)k * q.nvec + v0;
                        for (int c = lane; c < cols; c += 32) Bs[r * LDB + c] = __ldg(src + c);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                }
            }
        }
    }
}

template <int DMAX, int PT, int NVT, int WM, int WN>
cudaError_t launch_block(const BlockParams &q0, cudaStream_t stream) {
    BlockParams q = q0;
    using L = BlockSmem<PT, NVT>;
    auto kern = smolyak_block_kernel<DMAX, PT, NVT, WM, WN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total);
    if (e != cudaSuccess) return e;
    q.n_ptiles = (q.n + PT - 1) / PT;
    q.n_vtiles = (q.nvec + NVT - 1) / NVT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = q.n_ptiles * q.n_vtiles;
    const int grid = (int)std::min<int64_t>(tiles, sms);
    kern<<<grid, (kConsumerWarps + PT / 32 + 1) * 32, L::total, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

// tile shapes: 8 consumer warps as WM×WN, warp tile (PT/WM) points × (NVT/WN) vectors.  With many dimensions the
// producer's register state is large, so those variants keep 64 points (11 warps, 168 registers per thread).
template <int DMAX>
cudaError_t dispatch_block(const BlockParams &q, cudaStream_t stream) {
    if (q.nvec > 128) return launch_block<DMAX, 64, 256, 2, 4>(q, stream);     // warp tile 32 points × 64 vectors
    if (q.nvec > 64) return launch_block<DMAX, 64, 128, 2, 4>(q, stream);      // 32 × 32
    if constexpr (DMAX <= 10) {
        if (q.nvec > 32) return launch_block<DMAX, 128, 64, 4, 2>(q, stream);  // 32 × 32
        if (q.nvec > 16) return launch_block<DMAX, 128, 32, 4, 2>(q, stream);  // 32 × 16
        if (q.nvec > 8) return launch_block<DMAX, 128, 16, 8, 1>(q, stream);   // 16 × 16
        return launch_block<DMAX, 128, 8, 8, 1>(q, stream);                    // 16 × 8
    } else {
        if (q.nvec > 32) return launch_block<DMAX, 64, 64, 2, 4>(q, stream);   // 32 × 16
        if (q.nvec > 16) return launch_block<DMAX, 64, 32, 4, 2>(q, stream);   // 16 × 16
        if (q.nvec > 8) return launch_block<DMAX, 64, 16, 8, 1>(q, stream);    // 8 × 16
        return launch_block<DMAX, 64, 8, 8, 1>(q, stream);                     // 8 × 8
    }
}

}  // namespace

bool smolyak_block_supported(const sk_plan &plan) {
    return plan.basis.kind == SK_BASIS_SMOLYAK && plan.ncomp == 0 && plan.d_adv != nullptr && plan.basis.d <= 16;
}

// advance program: byte k = the (0-based) dimension whose index moved to reach term k from term k-1
// (all lower dimensions returned to index 1); 0xFE for the first term; 0xFF padding to a multiple of 16
void smolyak_block_program(const sk_plan &plan, std::vector<uint8_t> &adv) {
    const skh::SmolyakSet &S = plan.set;
    const size_t K = (size_t)S.K, d = (size_t)S.d;
    adv.assign((K + kKC - 1) / kKC * kKC, 0xFF);
    adv[0] = 0xFE;
    for (size_t k = 1; k < K; k++) {
        size_t c = 0;
        for (size_t j = 0; j < d; j++) if (S.idx[k * d + j] != S.idx[(k - 1) * d + j]) c = j;
        adv[k] = (uint8_t)c;
    }
}

double smolyak_block_flops(const sk_plan &plan, int nvec) { return 2.0 * (double)plan.K * nvec + 3.0 * (double)plan.K + 12.0 * plan.basis.d; }

cudaError_t smolyak_block_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n,
                                  double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    BlockParams q{};
    q.x = x; q.theta = theta; q.out = out; q.adv = plan.d_adv;
    q.n = n; q.K = plan.K; q.d = plan.basis.d; q.nvec = nvec;
    q.bulk_ok = (nvec & 1) == 0 && (reinterpret_cast<uintptr_t>(theta) & 15) == 0;
    for (int j = 0; j < q.d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    return q.d <= 6 ? dispatch_block<6>(q, stream) : q.d <= 10 ? dispatch_block<10>(q, stream) : dispatch_block<16>(q, stream);
}

}  // namespace skk
