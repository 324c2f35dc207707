// sm_100a kernels, part 4 — K4: Smolyak linear_combination with a K×nvec coefficient BLOCK
// (SVector coefficients: src/derivatives.jl:15,19 with src/generic_api.jl:114-128; the several-
// policy-functions caller of src/experimental.jl:228-233).  out[n][nvec] = A[n][K] · Θ[K][nvec], where
// A (the collocation matrix, generic_api.jl:217-227) is never materialised: its tiles are generated
// on chip and contracted with FP64 tensor-core tiles (DMMA, mma.sync.m8n8k4.f64).
//
// Measured on B200 (profiles/micro/fp64_mix.cu): DMMA and the FP64 CUDA-core instructions share one
// pipe, and as soon as two warps per scheduler stream DMMAs, a warp issuing DFMA/DMUL gets NO slot
// until they stop.  A concurrent "producer warp" design therefore degenerates into T_prod + T_cons
// with the producer at its latency-bound worst (first version of this file: 56 % of the DMMA peak).
// So the CTA alternates explicitly between two phases and makes the FP64 part of phase A short and
// free of dependency chains:
//
//   tile prologue   per point, the Chebyshev values T_i(x_j) of every dimension (chebyshev.jl:60-70 by
//                   recurrence, after transform_to, transformations.jl:183-320) go to a TABLE:
//                   indices <= HL in shared memory [row][point], the rare higher ones (one per term at
//                   most, a few % of the terms) to an L2-resident global scratch.
//   phase A         64 basis terms × 64 points: every thread forms 16 products of at most F = min(d, B)
//                   table entries (a term has at most B indices != 1).  The plan's factor program (byte
//                   offsets of the table rows, built once from the traversal order,
//                   smolyak_traversal.jl:211-231) replaces `__inc`; products are independent, so the
//                   loads pipeline and the DMULs issue back to back.
//   phase B         4 × 16 terms (wide tiles) of DMMA: 8 warps, warp tile 32 points × NVT/4 vectors, accumulators in
//                   registers, A fragments from the phase-A tile, B fragments from the Θ ring.
//   loader warp     (asynchronous to the phases; its warpgroup runs at 40 registers so that the workers get
//                   232, see setmaxnreg below) streams 16×NVT slabs of Θ — rows are contiguous in the
//                   caller's layout — with the bulk-copy engine (cp.async.bulk + mbarrier complete_tx)
//                   into rows padded to NVT+4 doubles (conflict-free fragment loads), and the factor
//                   program of the next 64 terms.  Θ (159 MB at cfg 5) is read by all CTAs in the same
//                   order: once from HBM, 147 times from L2.
//
// Algorithmic work per point: 2·K·nvec flop in the contraction + (F−1)·K in phase A.
// Summation order differs from the reference (k4-blocked, fma inside the tensor core) and products are
// formed as a tree: SK_MODE_FAST tolerance; SK_MODE_EXACT requests go to the direct kernel.
#include <algorithm>
#include <cstdlib>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

namespace {

constexpr int kPT = 64;              // points per tile
constexpr int kKA = 64;              // terms per phase-A chunk
// terms per Θ stage: 16 for wide vector tiles (a 3-stage ring of 33 KB slabs); narrow tiles stage more terms per
// slab, otherwise the loop runs at the latency of one bulk copy per 16 terms (measured: 2.7 µs per 64 terms)
__host__ __device__ constexpr int stage_terms(int NVT) { return NVT >= 128 ? 16 : NVT >= 64 ? 32 : 64; }
constexpr int kNS = 3;               // Θ stages
constexpr int kGT = 8;               // terms a thread forms together in phase A (independent products in flight)
constexpr int kWorkers = 8;          // worker warps (phase A and DMMA)
constexpr int kLDP = kPT + 4;        // ≡ 4 (mod 16) doubles: conflict-free A fragments
constexpr uint32_t kHigh = 0x80000000u;   // factor lives in the global scratch

struct BlockParams {
    const double *x;                 // [n][d]
    const double *theta;             // [K][nvec]
    double *out;                     // [n][nvec]
    const uint32_t *prog;            // factor program, one block of prog_chunk_words per 64 terms
    double *scratch;                 // [grid][hrows][kPT] high-index table entries (or nullptr)
    int64_t n, K;
    int d, nvec, H, HL;              // table rows in smem: indices 2..HL of every dimension
    int n_vtiles;
    int64_t n_ptiles;
    int bulk_ok;                     // Θ rows can be bulk-copied (16-byte aligned base, even nvec)
    int dbg;                         // experiments (SK_BLOCK_DBG): 1 = skip the DMMAs, 2 = skip phase A
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
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory"); }
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// factor-program geometry shared by host and device: FP slots (4 or 8 words) per term, then one 16-byte
// trailer per 64-term chunk (word 0: bit g set = group g of kGT terms has a factor in the global scratch)
__host__ __device__ constexpr int slots_padded(int F) { return F <= 4 ? 4 : 8; }
__host__ __device__ constexpr int prog_chunk_words(int F) { return kKA * slots_padded(F) + 4; }

template <int NVT>
struct BlockSmem {
    static constexpr int LDB = NVT + 4;                       // ≡ 4 (mod 16) for NVT >= 16; rows stay 16-byte aligned
    static constexpr size_t a_bytes = (size_t)kKA * kLDP * 8;
    static constexpr int KC = stage_terms(NVT);
    static constexpr size_t b_stage = (size_t)KC * LDB * 8;
    static constexpr size_t prog_bytes = (size_t)prog_chunk_words(8) * 4;
    static constexpr size_t off_a = 0;
    static constexpr size_t off_b = off_a + a_bytes;
    static constexpr size_t off_prog = off_b + kNS * b_stage;
    static constexpr size_t off_bar = off_prog + 2 * prog_bytes;
    static constexpr size_t off_tab = off_bar + 128;
    static size_t total(int rows) { return off_tab + (size_t)rows * kPT * 8; }
};
// table rows that fit next to the largest variant's buffers (the program depends on HL, so HL depends on d only)
constexpr size_t kSmemMax = 227 * 1024;
inline int table_rows_max() { return (int)((kSmemMax - BlockSmem<256>::off_tab) / (kPT * 8)); }

// product of F table entries (tree order); `slot` holds byte offsets relative to the thread's table column
template <int F, bool HIGH>
__device__ __forceinline__ double term_value(const uint32_t *slot, const char *tabp, const char *scrp) {
    double f[F];
#pragma unroll
    for (int i = 0; i < F; i++) {
        const uint32_t o = slot[i];
        if (HIGH && (o & kHigh)) f[i] = __ldcg(reinterpret_cast<const double *>(scrp + (o & ~kHigh)));
        else f[i] = *reinterpret_cast<const double *>(tabp + o);
    }
#pragma unroll
    for (int w = 1; w < F; w <<= 1)
#pragma unroll
        for (int i = 0; i + w < F; i += 2 * w) f[i] = f[i] * f[i + w];
    return f[0];
}

template <int F, int NVT, int WM, int WN>
__global__ void __launch_bounds__((kWorkers + 4) * 32, 1) smolyak_block_kernel(const __grid_constant__ BlockParams q) {
    static_assert(WM * WN == kWorkers, "eight worker warps");
    using L = BlockSmem<NVT>;
    constexpr int LDB = L::LDB, KC = L::KC;
    constexpr int WTM = kPT / WM, WTN = NVT / WN;         // warp tile
    constexpr int MT = WTM / 8, NT = WTN / 8;
    constexpr int FP = slots_padded(F);
    static_assert(WTM % 8 == 0 && WTN % 8 == 0 && MT >= 1 && NT >= 1, "warp tile must be made of 8x8 DMMA tiles");
    extern __shared__ __align__(128) unsigned char smem[];
    double *As = reinterpret_cast<double *>(smem + L::off_a);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::off_bar);     // Θ stage filled
    uint64_t *empty = full + kNS;                                          // Θ stage consumed by all worker warps
    uint64_t *pfull = empty + kNS;                                         // factor program chunk arrived
    uint64_t *pempty = pfull + 2;                                          // ... and was read by all worker warps
    double *tab = reinterpret_cast<double *>(smem + L::off_tab);           // [row][point]; row 0 = ones, row 1 = zeros
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kNS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWorkers); }
        for (int s = 0; s < 2; s++) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], kWorkers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < kPT) { tab[threadIdx.x] = 1.0; tab[kPT + threadIdx.x] = 0.0; }
    __syncthreads();
    const int64_t n_chunks = (q.K + kKA - 1) / kKA;        // phase-A chunks per tile
    const int64_t n_tiles = q.n_ptiles * q.n_vtiles;
    const int hrows = q.d * (q.H - q.HL);                 // scratch rows per CTA
    uint32_t it = 0;                                      // Θ stage counter (continues across tiles)
    uint32_t pit = 0;                                     // program chunk counter

    if (warp < kWorkers) {
        // two worker warpgroups at 232 registers: setmaxnreg moves registers between whole warpgroups of a CTA
        // (the register file is per scheduler: with a third full-size warp there, everyone is capped at 168)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int wm = warp / WN, wn = warp % WN;
        const int g = lane >> 2, t = lane & 3;
        const int pt = threadIdx.x & (kPT - 1), qd = threadIdx.x / kPT;       // phase A: point, quarter of the chunk
        const char *tabp = reinterpret_cast<const char *>(tab + pt);
        double *scr = q.scratch ? q.scratch + (size_t)blockIdx.x * hrows * kPT + pt : nullptr;
        int64_t built_tile = -1;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t ptile = tile / q.n_vtiles;
            const int64_t p0 = ptile * kPT;
            const int v0 = (int)(tile % q.n_vtiles) * NVT;
            // ---- prologue: Chebyshev table of this tile's points (dimension j handled by quarter j mod 4)
            if (ptile != built_tile) {
                const int64_t p = min(p0 + pt, q.n - 1);      // threads past the end recompute the last point
                for (int j = qd; j < q.d; j += 4) {
                    const double xj = skd::to_pm1(q.tr[j], q.x[(size_t)p * q.d + j]);
                    const double x2 = 2.0 * xj;
                    double tc = xj, tp = 1.0;                 // T_1, T_0 (basis indices 2, 1)
                    double *lo = tab + (size_t)(2 + j * (q.HL - 1)) * kPT + pt;
                    double *hi = scr + (size_t)j * (q.H - q.HL) * kPT;
                    for (int i = 2; i <= q.H; i++) {
                        if (i <= q.HL) lo[(size_t)(i - 2) * kPT] = tc;
                        else hi[(size_t)(i - q.HL - 1) * kPT] = tc;
                        const double tn = fma(x2, tc, -tp);   // chebyshev.jl:65-70
                        tp = tc; tc = tn;
                    }
                }
                built_tile = ptile;
                workers_sync();
            }
            double acc[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int64_t ch = 0; ch < n_chunks; ch++, pit++) {
                // ---- phase A: 16 terms per thread
                const int pb = pit & 1;
                mbar_wait(&pfull[pb], (pit >> 1) & 1);
                const uint32_t *pg = reinterpret_cast<const uint32_t *>(smem + L::off_prog + (size_t)pb * L::prog_bytes);
                const uint32_t himask = pg[kKA * FP];
                if (!(q.dbg & 2)) {
#pragma unroll
                    for (int grp = 0; grp < 16 / kGT; grp++) {
                        const int r0 = qd * 16 + grp * kGT;
                        uint32_t slot[kGT][FP];
#pragma unroll
                        for (int u = 0; u < kGT; u++)
#pragma unroll
                            for (int w = 0; w < FP; w += 4) {
                                const uint4 v = *reinterpret_cast<const uint4 *>(pg + (r0 + u) * FP + w);
                                slot[u][w] = v.x; slot[u][w + 1] = v.y; slot[u][w + 2] = v.z; slot[u][w + 3] = v.w;
                            }
                        double a[kGT];
                        if ((himask >> (qd * (16 / kGT) + grp)) & 1u) {
#pragma unroll
                            for (int u = 0; u < kGT; u++) a[u] = term_value<F, true>(slot[u], tabp, reinterpret_cast<const char *>(scr));
                        } else {
#pragma unroll
                            for (int u = 0; u < kGT; u++) a[u] = term_value<F, false>(slot[u], tabp, nullptr);
                        }
#pragma unroll
                        for (int u = 0; u < kGT; u++) As[(r0 + u) * kLDP + pt] = a[u];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&pempty[pb]);
                workers_sync();                               // the A tile is complete
                // ---- phase B: 4 Θ stages of 16 terms
#pragma unroll 1
                for (int sub = 0; sub < kKA / KC; sub++, it++) {
                    const int s = it % kNS;
                    mbar_wait(&full[s], (it / kNS) & 1);
                    const double *Ap = As + sub * KC * kLDP + wm * WTM + g;
                    const double *Bp = reinterpret_cast<const double *>(smem + L::off_b + (size_t)s * L::b_stage) + wn * WTN + g;
                    if (!(q.dbg & 1)) {
                        double a[2][MT], b[2][NT];
#pragma unroll
                        for (int i = 0; i < MT; i++) a[0][i] = Ap[t * kLDP + i * 8];
#pragma unroll
                        for (int j = 0; j < NT; j++) b[0][j] = Bp[t * LDB + j * 8];
#pragma unroll
                        for (int kk = 0; kk < KC / 4; kk++) {
                            if (kk + 1 < KC / 4) {           // fragments of the next k4 step fly while this one computes
#pragma unroll
                                for (int i = 0; i < MT; i++) a[(kk + 1) & 1][i] = Ap[((kk + 1) * 4 + t) * kLDP + i * 8];
#pragma unroll
                                for (int j = 0; j < NT; j++) b[(kk + 1) & 1][j] = Bp[((kk + 1) * 4 + t) * LDB + j * 8];
                            }
#pragma unroll
                            for (int i = 0; i < MT; i++)
#pragma unroll
                                for (int j = 0; j < NT; j++) dmma884(acc[i][j][0], acc[i][j][1], a[kk & 1][i], b[kk & 1][j]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                workers_sync();                               // everyone finished reading the A tile (and, at the
                                                              // last chunk, the table may be rebuilt)
            }
            // ---- epilogue: C fragment (row g, columns 2t, 2t+1) → out[p][v]
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
    } else {
        // ------------------------------------------------------------------ loader warpgroup: gives its registers
        // to the workers (128·(168−40) = 2·128·(232−168)); one warp streams Θ and the program, three retire
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp > kWorkers) return;
        constexpr uint32_t chunk_bytes = (uint32_t)prog_chunk_words(F) * 4;
        auto load_prog = [&](int64_t ch, uint32_t pi) {
            const int pb = pi & 1;
            if (lane == 0) {
                mbar_wait(&pempty[pb], ((pi >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&pfull[pb], chunk_bytes);
                bulk_load(smem + L::off_prog + (size_t)pb * L::prog_bytes, q.prog + (size_t)ch * prog_chunk_words(F), chunk_bytes, &pfull[pb]);
            }
        };
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int v0 = (int)(tile % q.n_vtiles) * NVT;
            const int cols = min(NVT, q.nvec - v0);
            const bool bulk = q.bulk_ok && (cols & 1) == 0;       // v0 is a multiple of NVT (even): rows stay 16-byte aligned
            load_prog(0, pit);
            for (int64_t ch = 0; ch < n_chunks; ch++, pit++) {
                if (ch + 1 < n_chunks) load_prog(ch + 1, pit + 1);
                for (int sub = 0; sub < kKA / KC; sub++, it++) {
                    const int s = it % kNS;
                    mbar_wait(&empty[s], ((it / kNS) & 1) ^ 1);
                    double *Bs = reinterpret_cast<double *>(smem + L::off_b + (size_t)s * L::b_stage);
                    const int64_t k0 = ch * kKA + sub * KC;
                    if (bulk) {
                        if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(KC * cols * 8));
                        __syncwarp();
                        for (int r = lane; r < KC; r += 32) {
                            const int64_t k = min(k0 + r, q.K - 1);      // rows past K: their A entries are exactly 0
                            bulk_load(Bs + r * LDB, q.theta + (size_t)k * q.nvec + v0, (uint32_t)(cols * 8), &full[s]);
                        }
                    } else {
                        for (int r = 0; r < KC; r++) {
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
}

template <int F, int NVT, int WM, int WN>
cudaError_t launch_block(const BlockParams &q0, int grid, int rows, cudaStream_t stream) {
    BlockParams q = q0;
    using L = BlockSmem<NVT>;
    auto kern = smolyak_block_kernel<F, NVT, WM, WN>;
    const size_t smem = L::total(rows);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    q.n_vtiles = (q.nvec + NVT - 1) / NVT;
    const int64_t tiles = q.n_ptiles * q.n_vtiles;
    kern<<<(int)std::min<int64_t>(tiles, grid), (kWorkers + 4) * 32, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

// vector-tile shapes: 8 worker warps as WM×WN, warp tile (64/WM) points × (NVT/WN) vectors
template <int F>
cudaError_t dispatch_block(const BlockParams &q, int grid, int rows, cudaStream_t stream) {
    if (q.nvec > 128) return launch_block<F, 256, 2, 4>(q, grid, rows, stream);    // 32 points × 64 vectors per warp
    if (q.nvec > 64) return launch_block<F, 128, 2, 4>(q, grid, rows, stream);     // 32 × 32
    if (q.nvec > 32) return launch_block<F, 64, 2, 4>(q, grid, rows, stream);      // 32 × 16
    if (q.nvec > 16) return launch_block<F, 32, 4, 2>(q, grid, rows, stream);      // 16 × 16
    if (q.nvec > 8) return launch_block<F, 16, 8, 1>(q, grid, rows, stream);       // 8 × 16
    return launch_block<F, 8, 8, 1>(q, grid, rows, stream);                        // 8 × 8
}

int block_table_HL(int d, int64_t H) { return (int)std::min<int64_t>(H, 1 + (table_rows_max() - 2) / d); }

}  // namespace

bool smolyak_block_supported(const sk_plan &plan) {
    return plan.basis.kind == SK_BASIS_SMOLYAK && plan.ncomp == 0 && plan.d_bprog != nullptr;
}

// Factor program.  Table rows: 0 = ones, 1 = zeros, 2 + j·(HL−1) + (i−2) = T_{i−1}(x_j) for 2 <= i <= HL (shared
// memory), scratch row j·(H−HL) + (i−HL−1) for i > HL (flagged with bit 31).  Entries are BYTE offsets
// row·64·8 relative to the thread's column.  Terms past K multiply the zeros row.
bool smolyak_block_program(const sk_plan &plan, std::vector<uint32_t> &prog, int *F_out, int *HL_out) {
    const skh::SmolyakSet &S = plan.set;
    const size_t K = (size_t)S.K, d = (size_t)S.d;
    int F = 1;
    for (size_t k = 0; k < K; k++) {
        int nf = 0;
        for (size_t j = 0; j < d; j++) nf += S.idx[k * d + j] != 1;
        F = std::max(F, nf);
    }
    if (F > 8) return false;
    const int Fk = F <= 3 ? 3 : F <= 5 ? 5 : 8;                     // kernel variants
    const int FP = slots_padded(Fk), CW = prog_chunk_words(Fk);
    const int HL = block_table_HL((int)d, S.H);
    if (HL < 1 || (uint64_t)d * (uint64_t)(S.H - HL) * kPT * 8 >= kHigh) return false;
    const size_t n_chunks = (K + kKA - 1) / kKA;
    prog.assign(n_chunks * CW, 0u);
    for (size_t c = 0; c < n_chunks; c++) {
        uint32_t *blk = prog.data() + c * CW;
        for (int r = 0; r < kKA; r++) {
            const size_t k = c * kKA + r;
            uint32_t *sl = blk + r * FP;                           // unused slots stay 0 = ones row
            if (k >= K) { sl[0] = 1u * kPT * 8; continue; }
            int nf = 0;
            for (size_t j = 0; j < d; j++) {
                const int i = S.idx[k * d + j];
                if (i == 1) continue;
                if (i <= HL) sl[nf++] = (uint32_t)((2 + j * (HL - 1) + (i - 2)) * kPT * 8);
                else {
                    sl[nf++] = kHigh | (uint32_t)((j * (S.H - HL) + (i - HL - 1)) * kPT * 8);
                    blk[kKA * FP] |= 1u << (r / kGT);
                }
            }
        }
    }
    *F_out = Fk; *HL_out = HL;
    return true;
}

double smolyak_block_flops(const sk_plan &plan, int nvec) {
    return 2.0 * (double)plan.K * nvec + (double)(plan.bprog_F - 1) * (double)plan.K + 2.0 * plan.basis.d * (double)(plan.H - 1) + 10.0 * plan.basis.d;
}

cudaError_t smolyak_block_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n,
                                  double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    BlockParams q{};
    q.x = x; q.theta = theta; q.out = out; q.prog = plan.d_bprog;
    q.n = n; q.K = plan.K; q.d = plan.basis.d; q.nvec = nvec; q.H = (int)plan.H; q.HL = plan.bprog_HL;
    q.bulk_ok = (nvec & 1) == 0 && (reinterpret_cast<uintptr_t>(theta) & 15) == 0;
    static const int dbg = [] { const char *e = getenv("SK_BLOCK_DBG"); return e ? atoi(e) : 0; }();
    q.dbg = dbg;
    for (int j = 0; j < q.d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    q.n_ptiles = (n + kPT - 1) / kPT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int rows = 2 + q.d * (q.HL - 1);
    const size_t scratch_bytes = (size_t)sms * q.d * (q.H - q.HL) * kPT * 8;
    cudaError_t e = cudaSuccess;
    if (scratch_bytes) {
        e = cudaMallocAsync(reinterpret_cast<void **>(&q.scratch), scratch_bytes, stream);
        if (e != cudaSuccess) return e;
    }
    switch (plan.bprog_F) {
    case 3: e = dispatch_block<3>(q, sms, rows, stream); break;
    case 5: e = dispatch_block<5>(q, sms, rows, stream); break;
    default: e = dispatch_block<8>(q, sms, rows, stream); break;
    }
    if (q.scratch) { cudaError_t e2 = cudaFreeAsync(q.scratch, stream); if (e == cudaSuccess) e = e2; }
    return e;
}

}  // namespace skk
