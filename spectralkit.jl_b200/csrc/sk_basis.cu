// sm_100a kernels, part 6 — K3: batched basis_at / collocation_matrix of a Smolyak basis (generic_api.jl:217-227,
// smolyak_api.jl:89-110), bit-identical to the reference's operation order, HBM-write-bound.
//
// out is Julia's Matrix{eltype}: column-major n×K of E-double elements, out[(k·ld + i)·E + c] — seen from the kernel a
// row-major [K][ld·E] array of doubles whose rows are the basis terms.  What the round-1 kernel got wrong was on-chip:
// it kept every dimension's FULL Chebyshev table (d·ℓ(B)·(order+1) doubles per point) in shared memory, which capped a
// CTA at 32 points, refused large bases, and left the gradient-valued matrix at 3 % of HBM.  Here:
//
//   * The K terms are cut into CHUNKS of consecutive terms (plan time, host).  A chunk only needs the table rows
//     (dimension, index) its terms touch — index 1 is T_0 = 1 and is never stored for plain values — so the table of a
//     chunk is a few dozen rows whatever d·ℓ(B) is: the per-dimension table is STREAMED chunk by chunk (the recurrence
//     walks up to the chunk's highest index of each dimension and keeps only the rows on the chunk's list).  No basis
//     is refused any more.
//   * One CTA = 32 points (lane = point) × 8 warps.  Build phase: warp w runs the exact recurrence
//     (chebyshev.jl:60-70, derivatives.jl:125-134 order) of dimensions w, w+8, …; compute phase: warp w owns TW
//     consecutive terms of every chunk slice, forms the left-to-right products (derivatives.jl:502-511) from the rows
//     and stages a [TW terms][32·E] tile in its own double-buffered shared-memory slot.
//   * Tiles leave through TMA: one tensor map over the [K][n·E] view (row pitch ld·E·8 bytes) per call, one
//     cp.async.bulk.tensor.2d store per tile (SASS UTMASTG).  The descriptor clips partial point tiles and the last
//     terms; ld ≠ n is just the pitch.  Destinations the descriptor cannot express (pitch or base not 16-byte aligned)
//     take coalesced st.global.cs from the same staging tile.
//
// Algorithmic bytes: 8·E per matrix element written + 8·d per point read per CTA pass; flops ≈ (d−1)·E multiplies per
// element: HBM-write-bound by design.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <vector>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

using skd::Jet;

namespace {

constexpr int kP = 32;            // points per CTA (lane = point)
constexpr int kWarps = 8;

int sm_count() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

// ---- chunk program (host, once per plan and jet width)
// words (uint32) per chunk header: term0, nterms, nrows, build_off (uint16 units into `build`)
struct ChunkHdr { uint32_t term0, nterms, nrows, build_off; };
// build list of a chunk, uint16: for every dimension j: maxidx_j, cnt_j, then cnt_j × (index, row) ascending
// factor table: per term kFacW uint16 row slots; values-only plans: the factors with index != 1 in dimension order,
// 0xFFFF-terminated (x·1.0 == x bit for bit); ∂ plans: all d factors (row 0 = the jet of T_0 = (1, 0, …))
struct BasisProgram {
    std::vector<ChunkHdr> hdr;
    std::vector<uint16_t> build, fac;
    int facw = 8, max_rows = 0, tk = 0;
};

// terms per chunk: as many as keep the chunk's rows within `row_cap`, in multiples of kWarps·TW
void build_program(const skh::SmolyakSet &S, bool jets, int row_cap, int tw, BasisProgram &bp) {
    const int d = S.d;
    bp.facw = d <= 8 ? 8 : 16;
    bp.fac.assign((size_t)S.K * bp.facw, 0);                        // row 0 = the jet of T_0 = (1, 0, …): padding multiplies by 1.0
    const int step = kWarps * tw;
    std::vector<int> slot((size_t)d * ((size_t)S.H + 1), -1);      // row of (dim, index) in the open chunk
    std::vector<std::pair<int, int>> used;                          // (dim, index) of the open chunk
    int64_t k = 0;
    while (k < S.K) {
        ChunkHdr h{(uint32_t)k, 0, 0, (uint32_t)bp.build.size()};
        int rows = 1;                                               // row 0: the constant jet of index 1
        used.clear();
        int64_t kend = k;
        for (;;) {
            const int64_t kn = std::min<int64_t>(S.K, kend + step);
            // rows the next slice of terms would add
            std::vector<std::pair<int, int>> add;
            for (int64_t t = kend; t < kn; t++)
                for (int j = 0; j < d; j++) {
                    const int i = S.idx[(size_t)t * d + j];
                    if (i == 1) continue;
                    int &s = slot[(size_t)j * (S.H + 1) + i];
                    if (s == -1) { s = -2; add.emplace_back(j, i); }
                }
            if (kend > k && rows + (int)add.size() > row_cap) {     // close the chunk before this slice
                for (auto &a : add) slot[(size_t)a.first * (S.H + 1) + a.second] = -1;
                break;
            }
            for (auto &a : add) { slot[(size_t)a.first * (S.H + 1) + a.second] = rows++; used.push_back(a); }
            kend = kn;
            if (kend >= S.K || (int)(kend - k) >= 4096) break;
        }
        h.nterms = (uint32_t)(kend - k); h.nrows = (uint32_t)rows;
        bp.max_rows = std::max(bp.max_rows, rows);
        // factor slots of the chunk's terms
        for (int64_t t = k; t < kend; t++) {
            uint16_t *f = &bp.fac[(size_t)t * bp.facw];
            int nf = 0;
            for (int j = 0; j < d; j++) {
                const int i = S.idx[(size_t)t * d + j];
                if (jets) f[j] = (uint16_t)(i == 1 ? 0 : slot[(size_t)j * (S.H + 1) + i]);
                else if (i != 1) f[nf++] = (uint16_t)slot[(size_t)j * (S.H + 1) + i];
            }
        }
        // build list, per dimension ascending by index
        std::sort(used.begin(), used.end());
        size_t u = 0;
        for (int j = 0; j < d; j++) {
            size_t u0 = u;
            while (u < used.size() && used[u].first == j) u++;
            const int cnt = (int)(u - u0);
            bp.build.push_back((uint16_t)(cnt ? used[u - 1].second : 0));
            bp.build.push_back((uint16_t)cnt);
            for (size_t v = u0; v < u; v++) {
                bp.build.push_back((uint16_t)used[v].second);
                bp.build.push_back((uint16_t)slot[(size_t)j * (S.H + 1) + used[v].second]);
            }
        }
        for (auto &a : used) slot[(size_t)a.first * (S.H + 1) + a.second] = -1;
        bp.hdr.push_back(h);
        k = kend;
    }
}

struct BasisParams {
    const double *x;
    double *out;
    int64_t n, ld, K;
    int d, E, ncomp, facw, tw, n_chunks, max_rows, use_tma, pb, nf;  // pb: points per TMA box (pb·E <= 256); nf: factors per term (values)
    int substride;                                                   // doubles between the sub-boxes of a staging tile (128-byte multiples: TMA source alignment)
    const double *theta;                                             // exact linear_combination: θ[K][nvec]
    int nvec;
    const ChunkHdr *hdr;
    const uint16_t *build, *fac;
    sk_transform tr[SK_MAX_DIM];
    int8_t orders[SK_MAX_COMP][SK_MAX_DIM];
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// slot m (0..15) of a term's factor words
template <int NW>
__device__ __forceinline__ uint32_t fac_slot(const uint32_t (&w)[NW], int m) { return (w[m >> 1] >> (16 * (m & 1))) & 0xFFFFu; }
// 32-bit shared-memory addressing (one IMAD per table read instead of 64-bit index arithmetic)
__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds128(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }

// products of TW consecutive terms for this lane's point, staged as [sub-box][term][point in box][component]
//   values-only plans: NF = min(d, B) compact factors per term, padded with row 0 (ones): x·1.0 == x bit for bit
//   gradient plans (GD > 0 = d): rows are (T, T') pairs, one LDS.128 per factor, prefix products shared between the
//     components — the same multiplications in the same left-to-right order as derivatives.jl:502-511
//   other ∂ plans: all d factors of every component, left to right
// rows_s: shared address of the rows + this lane's offset; st_s: shared address of this lane's slot in term 0 of the
// staging tile, tstride: bytes between consecutive terms there
template <int JW, int TW, bool BIG, int GD, int PPL>
__device__ __forceinline__ void tile_products(const BasisParams &q, uint32_t rows_s, const uint4 *fw, uint32_t st_s, uint32_t tstride) {
    const int E = q.E, d = q.d;
    constexpr int NW = BIG ? 8 : 4, NS = 2 * NW;      // factor words / slots per term (d <= 8: four words)
    constexpr int P = kP * PPL;                       // points per CTA: PPL per lane (lane, lane + 32)
    constexpr uint32_t ROWB = (uint32_t)JW * P * 8;   // bytes per table row
    static_assert(PPL == 1 || (JW == 1 && GD == 0), "two points per lane: plain values only");
    // every lane holds the factor words of term `lane` of the tile (lanes >= TW: unused); broadcast term by term.  Two words
    // (four slots) cover values-only plans with B <= 4 and ∂ plans with d <= 4: the rest sits behind one uniform branch
    const int nwords = q.ncomp == 0 ? (q.nf + 1) >> 1 : (d + 1) >> 1;
    uint32_t w[TW][NW];
#pragma unroll
    for (int t = 0; t < TW; t++) {
        w[t][0] = __shfl_sync(0xffffffffu, fw[0].x, t);
        w[t][1] = __shfl_sync(0xffffffffu, fw[0].y, t);
#pragma unroll
        for (int i = 2; i < NW; i++) w[t][i] = 0u;
    }
    if (nwords > 2) {
#pragma unroll
        for (int t = 0; t < TW; t++) {
            w[t][2] = __shfl_sync(0xffffffffu, fw[0].z, t);
            w[t][3] = __shfl_sync(0xffffffffu, fw[0].w, t);
            if constexpr (BIG) {
                w[t][4] = __shfl_sync(0xffffffffu, fw[1].x, t); w[t][5] = __shfl_sync(0xffffffffu, fw[1].y, t);
                w[t][6] = __shfl_sync(0xffffffffu, fw[1].z, t); w[t][7] = __shfl_sync(0xffffffffu, fw[1].w, t);
            }
        }
    }
    uint32_t dst[TW];
#pragma unroll
    for (int t = 0; t < TW; t++) dst[t] = st_s + (uint32_t)t * tstride;
    if (q.ncomp == 0) {
        // the PPL points of a lane (lane, lane + 32: 256 bytes apart in every row and in the staging tile) share the
        // factor words and every address computation — 40 % of the instructions of the one-point version
        double b[TW][PPL];
#pragma unroll
        for (int t = 0; t < TW; t++) {
            const uint32_t a0 = rows_s + fac_slot(w[t], 0) * ROWB;
#pragma unroll
            for (int pp = 0; pp < PPL; pp++) b[t][pp] = lds64(a0 + 256u * pp);
        }
#pragma unroll
        for (int t = 0; t < TW; t++) {
            const uint32_t a1 = rows_s + fac_slot(w[t], 1) * ROWB;      // padding = row 0 = 1.0: exact
#pragma unroll
            for (int pp = 0; pp < PPL; pp++) b[t][pp] = b[t][pp] * lds64(a1 + 256u * pp);
        }
#pragma unroll
        for (int m = 2; m < NS; m += 2) {
            if (m < q.nf) {
#pragma unroll
                for (int t = 0; t < TW; t++) {
                    const uint32_t am = rows_s + fac_slot(w[t], m) * ROWB;
#pragma unroll
                    for (int pp = 0; pp < PPL; pp++) b[t][pp] = b[t][pp] * lds64(am + 256u * pp);
                }
#pragma unroll
                for (int t = 0; t < TW; t++) {
                    const uint32_t am = rows_s + fac_slot(w[t], m + 1) * ROWB;
#pragma unroll
                    for (int pp = 0; pp < PPL; pp++) b[t][pp] = b[t][pp] * lds64(am + 256u * pp);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TW; t++)
#pragma unroll
            for (int pp = 0; pp < PPL; pp++) sts64(dst[t] + 256u * pp, b[t][pp]);
    } else if constexpr (GD > 0) {
        // rows_s already points at this lane's (T, T') pair: [row][lane] double2
#pragma unroll
        for (int t = 0; t < TW; t++) {
            double2 v[GD];
#pragma unroll
            for (int j = 0; j < GD; j++) v[j] = lds128(rows_s + fac_slot(w[t], j) * ROWB);
            double pre[GD];
            pre[0] = v[0].x;
#pragma unroll
            for (int j = 1; j < GD; j++) pre[j] = pre[j - 1] * v[j].x;
            sts64(dst[t], pre[GD - 1]);
#pragma unroll
            for (int m = 0; m < GD; m++) {
                double bb = m == 0 ? v[0].y : pre[m > 0 ? m - 1 : 0] * v[m].y;
#pragma unroll
                for (int j = m + 1; j < GD; j++) bb = bb * v[j].x;
                sts64(dst[t] + 8u * (m + 1), bb);
            }
        }
    } else {
        uint32_t ra[TW][NS];
#pragma unroll
        for (int t = 0; t < TW; t++) {
#pragma unroll
            for (int j = 0; j < NS; j++) ra[t][j] = rows_s + fac_slot(w[t], j) * ROWB;
        }
        for (int cc = 0; cc < E; cc++) {
            double b[TW];
            const uint32_t o0 = (uint32_t)q.orders[cc][0] * (P * 8);
#pragma unroll
            for (int t = 0; t < TW; t++) b[t] = lds64(ra[t][0] + o0);
#pragma unroll
            for (int j = 1; j < NS; j++) {
                if (j < d) {
                    const uint32_t o = (uint32_t)q.orders[cc][j] * (P * 8);
#pragma unroll
                    for (int t = 0; t < TW; t++) b[t] = b[t] * lds64(ra[t][j] + o);
                }
            }
#pragma unroll
            for (int t = 0; t < TW; t++) sts64(dst[t] + 8u * cc, b[t]);
        }
    }
}

// shared memory: [x jets d·JW·32][rows max_rows·JW·32][staging kWarps·2·tw·32·E], 128-byte aligned pieces
template <int JW, int TW, bool BIG, int GD, int PPL>
__global__ void __launch_bounds__(kP * kWarps, BIG ? 1 : 2) smolyak_tile_basis_kernel(const __grid_constant__ BasisParams q,
                                                                          const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int d = q.d, E = q.E;
    constexpr int P = kP * PPL;                                                         // points per CTA (PPL per lane)
    double *xj = reinterpret_cast<double *>(smem_raw);                                  // [d][JW][P]
    const size_t xj_bytes = ((size_t)d * JW * P * 8 + 127) & ~(size_t)127;
    double *rows = reinterpret_cast<double *>(smem_raw + xj_bytes);                     // [row][JW][P]
    const size_t rows_bytes = ((size_t)q.max_rows * JW * P * 8 + 127) & ~(size_t)127;
    const size_t tile_bytes = ((size_t)(P / q.pb) * q.substride * 8 + 127) & ~(size_t)127;
    double *stage = reinterpret_cast<double *>(smem_raw + xj_bytes + rows_bytes + (size_t)w * 2 * tile_bytes);
    const int nsub = P / q.pb;                                                          // TMA boxes per tile along the points
    // this lane's slot in term 0 of a staging tile ([sub-box][term][point in box][component]) and the stride between terms
    const uint32_t lane_off = (uint32_t)((lane / q.pb) * q.substride + (lane % q.pb) * E) * 8u, tstride = (uint32_t)(q.pb * E) * 8u;
    int buf = 0;

    for (int64_t p0 = (int64_t)blockIdx.x * P; p0 < q.n; p0 += (int64_t)gridDim.x * P) {
        __syncthreads();                                            // previous tile's rows and jets are no longer read
        for (int j = w; j < d; j += kWarps) {
#pragma unroll
            for (int pp = 0; pp < PPL; pp++) {
                const int64_t p = min(p0 + lane + 32 * pp, q.n - 1);
                double v[JW];
                skd::to_pm1_jet<JW>(q.tr[j], __ldcs(q.x + p * d + j), v);
#pragma unroll
                for (int r = 0; r < JW; r++) xj[((size_t)j * JW + r) * P + lane + 32 * pp] = v[r];
            }
        }
        if (w == kWarps - 1) {                                      // row 0: the jet of T_0 = (1, 0, …)
            if constexpr (GD > 0) reinterpret_cast<double2 *>(rows)[lane] = make_double2(1.0, 0.0);
            else {
#pragma unroll
                for (int pp = 0; pp < PPL; pp++)
#pragma unroll
                    for (int r = 0; r < JW; r++) rows[(size_t)r * P + lane + 32 * pp] = r == 0 ? 1.0 : 0.0;
            }
        }
        for (int c = 0; c < q.n_chunks; c++) {
            const ChunkHdr h = q.hdr[c];
            __syncthreads();                                        // jets written / previous chunk's rows consumed
            // ---- build: warp w walks the recurrence of dimensions w, w+8, … and keeps the rows on the chunk's list
            {
                const uint16_t *bl = q.build + h.build_off;
                for (int j = 0; j < d; j++) {
                    const int maxidx = bl[0], cnt = bl[1];
                    if (cnt > 0 && (j % kWarps) == w) {
                        Jet<JW> X[PPL], fp[PPL], fpp[PPL], f[PPL];
#pragma unroll
                        for (int pp = 0; pp < PPL; pp++) {
                            fp[pp] = skd::jet_one<JW>();
#pragma unroll
                            for (int r = 0; r < JW; r++) { X[pp].c[r] = xj[((size_t)j * JW + r) * P + lane + 32 * pp]; fpp[pp].c[r] = X[pp].c[r]; }
                        }
                        int nx = 0, want = bl[2];
                        for (int i = 2; i <= maxidx; i++) {
#pragma unroll
                            for (int pp = 0; pp < PPL; pp++) { f[pp] = skd::cheb_step_exact<JW>(X[pp], fp[pp], fpp[pp]); fpp[pp] = fp[pp]; fp[pp] = f[pp]; }
                            if (i == want) {
                                if constexpr (GD > 0) reinterpret_cast<double2 *>(rows)[(size_t)bl[3 + 2 * nx] * kP + lane] = make_double2(f[0].c[0], f[0].c[JW > 1 ? 1 : 0]);
                                else {
#pragma unroll
                                    for (int pp = 0; pp < PPL; pp++) {
                                        double *r0 = rows + (size_t)bl[3 + 2 * nx] * JW * P + lane + 32 * pp;
#pragma unroll
                                        for (int r = 0; r < JW; r++) r0[(size_t)r * P] = f[pp].c[r];
                                    }
                                }
                                nx++;
                                want = nx < cnt ? bl[2 + 2 * nx] : 0;
                            }
                        }
                    }
                    bl += 2 + 2 * cnt;
                }
            }
            __syncthreads();
            // ---- compute: warp w owns TW consecutive terms of every slice of kWarps·TW terms; the factor words of the next
            // tile are fetched (one coalesced 16-byte load per term, lanes 0..TW-1) while the current one is computed
            const uint4 zero4 = make_uint4(0, 0, 0, 0);
            auto fetch = [&](uint32_t s0, uint4 (&fw)[2]) {
                fw[0] = zero4; fw[1] = zero4;
                if (lane < TW && s0 + lane < h.nterms) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(q.fac + (size_t)(h.term0 + s0 + lane) * q.facw);
                    fw[0] = __ldg(src);
                    if (BIG) fw[1] = __ldg(src + 1);
                }
            };
            uint4 fw[2], fwn[2];
            uint32_t s0 = (uint32_t)w * TW;
            if (s0 < h.nterms) fetch(s0, fw);
            for (; s0 < h.nterms; s0 += kWarps * TW) {
                const uint32_t s1 = s0 + kWarps * TW;
                if (s1 < h.nterms) fetch(s1, fwn);
                double *st = stage + (size_t)buf * (tile_bytes / 8);
                const int nt = min((uint32_t)TW, h.nterms - s0);
                const int64_t k0 = (int64_t)h.term0 + s0;
                tile_products<JW, TW, BIG, GD, PPL>(q, (uint32_t)__cvta_generic_to_shared(rows) + (uint32_t)lane * (GD > 0 ? 16u : 8u), fw, (uint32_t)__cvta_generic_to_shared(st) + lane_off, tstride);
                if (q.use_tma) {
                    proxy_fence();
                    __syncwarp();
                    if (lane == 0) {
                        for (int sb = 0; sb < nsub; sb++)
                            tma_store_2d(&tmap, st + (size_t)sb * q.substride, (int)((p0 + (int64_t)sb * q.pb) * E), (int)k0);
                        tma_commit();
                        tma_wait_read_1();                          // the other buffer's read has completed
                    }
                    __syncwarp();
                } else {
                    __syncwarp();
                    const int64_t row_elems = (int64_t)P * E;
                    for (int t = 0; t < nt; t++) {
                        double *g = q.out + ((size_t)(k0 + t) * q.ld + p0) * E;
                        for (int e = lane; e < row_elems; e += 32) {
                            const int pt = e / E, cc = e - pt * E;
                            if (p0 + pt < q.n)
                                __stcs(g + e, st[(size_t)(pt / q.pb) * q.substride + (size_t)t * (q.pb * E) + (size_t)(pt % q.pb) * E + cc]);
                        }
                    }
                    __syncwarp();
                }
                buf ^= 1;
                fw[0] = fwn[0]; fw[1] = fwn[1];
            }
        }
    }
    if (q.use_tma && lane == 0) tma_wait_all();
}

// ---- bit-identical linear_combination on the same chunked tables (SK_MODE_EXACT for bases of any size):
// the reference's sequential s = s + θ_k·b_k (generic_api.jl:114-128: multiply, then add, never fused) in term order.
// One thread owns the sums of one point (lane) for the items of its warp — components of a ∂ request (item = component,
// warp w takes items w, w+8, …) or groups of 8 coefficient vectors (SVector coefficients, values only) — and walks ALL
// terms of every chunk in order; only the table rows are shared.  A checker-speed path (the products are re-formed per
// item), like the direct kernel it replaces for large bases.
constexpr int kMaxItems = 8;      // items per warp: 64 components, or 64 groups of 8 = 512 coefficient vectors
template <int JW, bool BIG>
__global__ void __launch_bounds__(kP * kWarps) smolyak_tile_lincomb_kernel(const __grid_constant__ BasisParams q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int d = q.d;
    double *xj = reinterpret_cast<double *>(smem_raw);
    const size_t xj_bytes = ((size_t)d * JW * kP * 8 + 127) & ~(size_t)127;
    double *rows = reinterpret_cast<double *>(smem_raw + xj_bytes);
    const bool vecs = q.nvec > 1;
    const int n_items = vecs ? (q.nvec + 7) / 8 : q.E;
    constexpr int NW = BIG ? 8 : 4;
    for (int64_t p0 = (int64_t)blockIdx.x * kP; p0 < q.n; p0 += (int64_t)gridDim.x * kP) {
        const int64_t p = min(p0 + lane, q.n - 1);
        __syncthreads();
        for (int j = w; j < d; j += kWarps) {
            double v[JW];
            skd::to_pm1_jet<JW>(q.tr[j], q.x[p * d + j], v);
#pragma unroll
            for (int r = 0; r < JW; r++) xj[((size_t)j * JW + r) * kP + lane] = v[r];
        }
        if (w == kWarps - 1) {
#pragma unroll
            for (int r = 0; r < JW; r++) rows[(size_t)r * kP + lane] = r == 0 ? 1.0 : 0.0;
        }
        double s[kMaxItems][8];                              // [item of this warp][vector of the group] (components: [.][0])
#pragma unroll
        for (int i = 0; i < kMaxItems; i++)
#pragma unroll
            for (int v = 0; v < 8; v++) s[i][v] = 0.0;
        for (int c = 0; c < q.n_chunks; c++) {
            const ChunkHdr h = q.hdr[c];
            __syncthreads();
            {
                const uint16_t *bl = q.build + h.build_off;
                for (int j = 0; j < d; j++) {
                    const int maxidx = bl[0], cnt = bl[1];
                    if (cnt > 0 && (j % kWarps) == w) {
                        Jet<JW> X, fp = skd::jet_one<JW>(), fpp, f;
#pragma unroll
                        for (int r = 0; r < JW; r++) { X.c[r] = xj[((size_t)j * JW + r) * kP + lane]; fpp.c[r] = X.c[r]; }
                        int nx = 0, want = bl[2];
                        for (int i = 2; i <= maxidx; i++) {
                            f = skd::cheb_step_exact<JW>(X, fp, fpp); fpp = fp; fp = f;
                            if (i == want) {
                                double *r0 = rows + (size_t)bl[3 + 2 * nx] * JW * kP + lane;
#pragma unroll
                                for (int r = 0; r < JW; r++) r0[(size_t)r * kP] = f.c[r];
                                nx++;
                                want = nx < cnt ? bl[2 + 2 * nx] : 0;
                            }
                        }
                    }
                    bl += 2 + 2 * cnt;
                }
            }
            __syncthreads();
            for (uint32_t t = 0; t < h.nterms; t++) {
                const int64_t k = (int64_t)h.term0 + t;
                const uint4 *src = reinterpret_cast<const uint4 *>(q.fac + (size_t)k * q.facw);
                uint32_t fwd[NW];
                { const uint4 a = __ldg(src); fwd[0] = a.x; fwd[1] = a.y; fwd[2] = a.z; fwd[3] = a.w; }
                if constexpr (BIG) { const uint4 b = __ldg(src + 1); fwd[4] = b.x; fwd[5] = b.y; fwd[6] = b.z; fwd[7] = b.w; }
                if (vecs || q.ncomp == 0) {
                    // plain values: the compact factors (row 0 pads with 1.0, exact)
                    double b = rows[(size_t)fac_slot(fwd, 0) * kP + lane];
#pragma unroll
                    for (int m = 1; m < 2 * NW; m++) if (m < q.nf) b = b * rows[(size_t)fac_slot(fwd, m) * kP + lane];
#pragma unroll
                    for (int i = 0; i < kMaxItems; i++) {
                        const int item = w + i * kWarps;
                        if (item < n_items) {
                            if (!vecs) {
                                const double th = q.theta[k];
                                s[i][0] = k == 0 ? th * b : s[i][0] + th * b;
                            } else {
                                const double *tk = q.theta + (size_t)k * q.nvec + item * 8;
#pragma unroll
                                for (int v = 0; v < 8; v++)
                                    if (item * 8 + v < q.nvec) s[i][v] = k == 0 ? tk[v] * b : s[i][v] + tk[v] * b;
                            }
                        }
                    }
                } else {
                    const double th = q.theta[k];
#pragma unroll
                    for (int i = 0; i < kMaxItems; i++) {
                        const int cc = w + i * kWarps;
                        if (cc < n_items) {
                            double b = rows[((size_t)fac_slot(fwd, 0) * JW + q.orders[cc][0]) * kP + lane];
#pragma unroll
                            for (int j = 1; j < 2 * NW; j++) if (j < d) b = b * rows[((size_t)fac_slot(fwd, j) * JW + q.orders[cc][j]) * kP + lane];
                            s[i][0] = k == 0 ? th * b : s[i][0] + th * b;
                        }
                    }
                }
            }
        }
        if (p0 + lane < q.n) {
#pragma unroll
            for (int i = 0; i < kMaxItems; i++) {
                const int item = w + i * kWarps;
                if (item < n_items) {
                    if (!vecs) q.out[(size_t)(p0 + lane) * q.E + item] = s[i][0];
                    else {
#pragma unroll
                        for (int v = 0; v < 8; v++) if (item * 8 + v < q.nvec) q.out[(size_t)(p0 + lane) * q.nvec + item * 8 + v] = s[i][v];
                    }
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// device-resident chunk programs of a plan, one per jet width, built on first use
// ---- bit-identical linear_combination, one coefficient vector, plain values or the full gradient [value, ∂1..∂d], d <= 8:
// ONE THREAD PER POINT walks all terms in order (the reference's sequential sum) and forms all components of a term from
// one set of table reads, sharing the prefix products between the components exactly as derivatives.jl:502-511 forms them
// (((T1·T2)·…·T_{m-1})·T_m')·T_{m+1}·… — the same multiplications in the same order, so every bit agrees).  A CTA is one warp
// with its own chunked table ([row][lane], (T, T') pairs for gradients): no CTA barrier anywhere, eight CTAs per SM.
// The kernel above splits the components over warps (one product chain per component and warp, 7× the table reads) and the
// direct kernel keeps full tables (two warps per SM): this one is 5–8× faster at cfg 4 (profiles/r2_exact_mode.jsonl).
template <int D, bool GRADIENT>
__global__ void __launch_bounds__(32) smolyak_warp_exact_kernel(const __grid_constant__ BasisParams q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    constexpr int JW = GRADIENT ? 2 : 1, NC = GRADIENT ? D + 1 : 1;
    double2 *rows2 = reinterpret_cast<double2 *>(smem_raw) + lane;      // gradient: [row][lane] (T, T')
    double *rows1 = reinterpret_cast<double *>(smem_raw) + lane;        // values:   [row][lane] T
    for (int64_t p0 = (int64_t)blockIdx.x * kP; p0 < q.n; p0 += (int64_t)gridDim.x * kP) {
        const int64_t p = min(p0 + lane, q.n - 1);
        Jet<JW> X[D];
#pragma unroll
        for (int j = 0; j < D; j++) skd::to_pm1_jet<JW>(q.tr[j], q.x[p * D + j], X[j].c);
        __syncwarp();                                                   // the previous tile's rows are no longer read
        if constexpr (GRADIENT) rows2[0] = make_double2(1.0, 0.0); else rows1[0] = 1.0;      // row 0: index 1, T_0 = (1, 0)
        double s[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) s[c] = 0.0;
        for (int ch = 0; ch < q.n_chunks; ch++) {
            const ChunkHdr h = q.hdr[ch];
            __syncwarp();                                               // the previous chunk's rows are no longer read
            const uint16_t *bl = q.build + h.build_off;
#pragma unroll
            for (int j = 0; j < D; j++) {
                const int maxidx = bl[0], cnt = bl[1];
                if (cnt > 0) {
                    Jet<JW> fp = skd::jet_one<JW>(), fpp = X[j], f;
                    int nx = 0, want = bl[2];
                    for (int i = 2; i <= maxidx; i++) {
                        f = skd::cheb_step_exact<JW>(X[j], fp, fpp); fpp = fp; fp = f;
                        if (i == want) {
                            const int r = bl[3 + 2 * nx];
                            if constexpr (GRADIENT) rows2[(size_t)r * kP] = make_double2(f.c[0], f.c[JW - 1]); else rows1[(size_t)r * kP] = f.c[0];
                            nx++;
                            want = nx < cnt ? bl[2 + 2 * nx] : 0;
                        }
                    }
                }
                bl += 2 + 2 * cnt;
            }
            __syncwarp();
            // Terms are taken four at a time: their products are independent (only the sums are sequential, and they are added
            // in term order), which multiplies the instruction-level parallelism of a warp that has the SM's schedulers almost to
            // itself; the factor words and coefficients of the next four are fetched while this one is computed (warp-uniform:
            // one broadcast each).
            const uint4 *facp = reinterpret_cast<const uint4 *>(q.fac) + h.term0;
            const double *thp = q.theta + h.term0;
            auto products = [&](const uint4 &a4, double (&b)[NC]) {
                const uint32_t fwd[4] = {a4.x, a4.y, a4.z, a4.w};
                if constexpr (GRADIENT) {
                    double2 v[D];
#pragma unroll
                    for (int j = 0; j < D; j++) v[j] = rows2[(size_t)fac_slot(fwd, j) * kP];
                    double pre[D];
                    pre[0] = v[0].x;
#pragma unroll
                    for (int j = 1; j < D; j++) pre[j] = pre[j - 1] * v[j].x;
                    b[0] = pre[D - 1];
#pragma unroll
                    for (int m = 0; m < D; m++) {
                        double bb = m == 0 ? v[0].y : pre[m > 0 ? m - 1 : 0] * v[m].y;
#pragma unroll
                        for (int j = m + 1; j < D; j++) bb = bb * v[j].x;
                        b[m + 1 < NC ? m + 1 : 0] = bb;
                    }
                } else {
                    double bb = rows1[(size_t)fac_slot(fwd, 0) * kP];
#pragma unroll
                    for (int m = 1; m < 8; m++) if (m < q.nf) bb = bb * rows1[(size_t)fac_slot(fwd, m) * kP];      // padding = row 0 = 1.0: exact
                    b[0] = bb;
                }
            };
            // generic_api.jl:114-128: multiply, then add, never fused; the first term starts the sum
            auto accumulate = [&](int64_t k, double th, const double (&b)[NC]) {
#pragma unroll
                for (int c = 0; c < NC; c++) s[c] = k == 0 ? th * b[c] : s[c] + th * b[c];
            };
            const uint32_t nt = h.nterms;
            constexpr int U = 4;                                        // terms per iteration
            uint4 fq[U];
            double tq[U];
#pragma unroll
            for (int u = 0; u < U; u++) { const uint32_t tt = min((uint32_t)u, nt - 1); fq[u] = __ldg(facp + tt); tq[u] = __ldg(thp + tt); }
            uint32_t t = 0;
            for (; t + U <= nt; t += U) {
                uint4 fa[U];
                double ta[U];
#pragma unroll
                for (int u = 0; u < U; u++) { fa[u] = fq[u]; ta[u] = tq[u]; }
#pragma unroll
                for (int u = 0; u < U; u++) { const uint32_t tt = min(t + U + u, nt - 1); fq[u] = __ldg(facp + tt); tq[u] = __ldg(thp + tt); }
                double b[U][NC];
#pragma unroll
                for (int u = 0; u < U; u++) products(fa[u], b[u]);
#pragma unroll
                for (int u = 0; u < U; u++) accumulate((int64_t)h.term0 + t + u, ta[u], b[u]);
            }
            // the last nt - t < U terms: their factor words are fq[0..], in order
#pragma unroll
            for (int u = 0; u < U - 1; u++) {
                if (t + u < nt) {
                    double b0[NC];
                    products(fq[u], b0);
                    accumulate((int64_t)h.term0 + t + u, tq[u], b0);
                }
            }
        }
        if (p0 + lane < q.n) {
#pragma unroll
            for (int c = 0; c < NC; c++) q.out[(size_t)(p0 + lane) * NC + c] = s[c];
        }
    }
}

struct ProgramCache {
    std::mutex mu;
    struct Entry { const sk_plan *plan; int jw, device, pts; BasisProgram host; ChunkHdr *d_hdr; uint16_t *d_build, *d_fac; int tw, row_cap; };
    std::list<Entry> entries;      // a list: entries keep their address while other threads add theirs
};
ProgramCache &cache() { static ProgramCache c; return c; }

}  // namespace

// forget (and free) the chunk programs of a plan that is being destroyed
void smolyak_tile_basis_forget(const sk_plan *plan) {
    ProgramCache &pc = cache();
    std::lock_guard<std::mutex> lk(pc.mu);
    for (auto it = pc.entries.begin(); it != pc.entries.end();) {
        if (it->plan == plan) {
            cudaFree(it->d_hdr); cudaFree(it->d_build); cudaFree(it->d_fac);
            it = pc.entries.erase(it);
        } else ++it;
    }
}


// one coefficient vector, plain values or the full gradient, d <= 8 (factor words of 8 slots): the one-thread-per-point kernel
bool smolyak_warp_exact_serves(const sk_plan &plan, int nvec) {
    static const int off = [] { const char *e = getenv("SK_EXACT_OLD"); return e ? atoi(e) : 0; }();
    if (off || plan.basis.kind != SK_BASIS_SMOLYAK || nvec != 1 || plan.basis.d > 8 || plan.H > 4096) return false;
    return plan.ncomp == 0 || plan.fast_mode == 1;
}
// exact-mode linear_combination for any basis size; false if the request is outside what the kernel's per-warp item
// table holds (more than 64 components or 512 coefficient vectors)
bool smolyak_tile_lincomb_supported(const sk_plan &plan, int nvec) {
    if (plan.basis.kind != SK_BASIS_SMOLYAK) return false;
    return nvec > 1 ? (plan.ncomp == 0 && nvec <= 8 * kMaxItems * kWarps) : plan.E <= kMaxItems * kWarps;
}
cudaError_t smolyak_tile_basis_at_impl(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out, int64_t ld, cudaStream_t stream);
cudaError_t smolyak_tile_lincomb_exact(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out, cudaStream_t stream) {
    return smolyak_tile_basis_at_impl(plan, theta, nvec, x, n, out, n, stream);
}
cudaError_t smolyak_tile_basis_at(const sk_plan &plan, const double *x, int64_t n, double *out, int64_t ld, cudaStream_t stream) {
    return smolyak_tile_basis_at_impl(plan, nullptr, 1, x, n, out, ld, stream);
}
// theta == nullptr: basis_at (TMA tiles); otherwise the bit-identical linear combination with the same tables
cudaError_t smolyak_tile_basis_at_impl(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out, int64_t ld, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int d = plan.basis.d, E = plan.E;
    int jw = 1;
    for (int j = 0; j < d; j++) jw = std::max(jw, plan.deg[j] + 1);
    if (plan.ncomp == 0) jw = 1;
    const int JW = jw <= 3 ? jw : 6;
    // staging tile of one warp: TW terms × 32 points × E doubles, 2-4 KB (one 16 KB buffer pair per warp at E = 64)
    // plain values (one double per element): two points per lane, 64 per CTA (SK_BASIS_PPL1=1: one; experiments)
    static const int ppl1 = [] { const char *e = getenv("SK_BASIS_PPL1"); return e ? atoi(e) : 0; }();
    const bool two = theta == nullptr && plan.ncomp == 0 && E == 1 && JW == 1 && !ppl1;
    // exact linear_combination, one coefficient vector, values or the full gradient in d <= 8: one warp per CTA, one thread
    // per point (smolyak_warp_exact_kernel); its tables are small chunks (eight CTAs per SM) under their own cache tag
    const bool warp_exact = smolyak_warp_exact_serves(plan, nvec) && theta != nullptr;
    const int P = two ? 2 * kP : warp_exact ? 1 : kP;                // (P doubles as the tag of the chunk-program cache)
    int tw = two ? 4 : 8;                                           // TW terms per warp tile: a power of two, 2-4 KB per tile
    while (tw > 1 && tw * E > 16) tw >>= 1;
    // TMA box: pb points × E components per row (<= 256 elements), TW rows; the sub-boxes of a staging tile start on
    // 128-byte boundaries (source alignment of cp.async.bulk.tensor)
    int pb = P;
    while (pb * E > 256) pb >>= 1;
    const int substride = (tw * pb * E + 15) & ~15;
    const size_t tile_bytes = (((size_t)(P / pb) * substride * 8) + 127) & ~(size_t)127;
    const size_t xj_bytes = ((size_t)d * JW * P * 8 + 127) & ~(size_t)127;
    const size_t stage_bytes = (size_t)kWarps * 2 * tile_bytes;
    const size_t budget = 200 * 1024;
    if (xj_bytes + stage_bytes + (size_t)(d + 2) * JW * P * 8 > budget) return cudaErrorInvalidConfiguration;
    // rows: enough for one slice of terms at least; more rows = larger chunks = fewer recurrence re-runs.  Two CTAs per SM
    // when the whole table is small.
    int row_cap = (int)((budget - xj_bytes - stage_bytes) / ((size_t)JW * P * 8));
    row_cap = std::min(row_cap, 1 + d * (int)plan.H);
    const int half_cap = (int)((budget / 2 - xj_bytes - stage_bytes > 0 ? budget / 2 - xj_bytes - stage_bytes : 0) / ((size_t)JW * P * 8));
    if (half_cap >= 1 + d * (int)std::min<int64_t>(plan.H, 16)) row_cap = std::min(row_cap, half_cap);
    // one slice of kWarps·TW terms must fit in any case: every term adds at most min(d, B) rows (its factors with index
    // != 1; index 1 is the shared row 0 for every kind of request).  (The bound used to be d rows per term, which pushed
    // 9- and 10-dimensional InteriorGrid bases with second-order requests past the shared memory of an SM: found by the
    // parity sweep, profiles/r2_fuzz_parity.jsonl.)
    row_cap = std::max(row_cap, kWarps * tw * std::min(d, std::max(plan.basis.B, 1)) + 1);
    if (warp_exact) {
        tw = 1;
        row_cap = std::max(JW == 2 ? 56 : 112, kWarps * std::min(d, std::max(plan.basis.B, 1)) + 1);      // 28 KB per CTA
        row_cap = std::min(row_cap, 1 + d * (int)plan.H);
    }

    ProgramCache &pc = cache();
    ProgramCache::Entry *en = nullptr;
    {
        std::lock_guard<std::mutex> lk(pc.mu);
        int dev = 0;
        cudaGetDevice(&dev);
        for (auto &e : pc.entries) if (e.plan == &plan && e.jw == JW && e.device == dev && e.pts == P) { en = &e; break; }
        if (!en) {
            ProgramCache::Entry e{&plan, JW, dev, P, {}, nullptr, nullptr, nullptr, tw, row_cap};
            build_program(plan.set, plan.ncomp > 0, row_cap, tw, e.host);
            cudaError_t ce = cudaMalloc(&e.d_hdr, e.host.hdr.size() * sizeof(ChunkHdr));
            if (ce == cudaSuccess) ce = cudaMalloc(&e.d_build, std::max<size_t>(2, e.host.build.size() * 2));
            if (ce == cudaSuccess) ce = cudaMalloc(&e.d_fac, e.host.fac.size() * 2);
            if (ce == cudaSuccess) ce = cudaMemcpy(e.d_hdr, e.host.hdr.data(), e.host.hdr.size() * sizeof(ChunkHdr), cudaMemcpyHostToDevice);
            if (ce == cudaSuccess && !e.host.build.empty()) ce = cudaMemcpy(e.d_build, e.host.build.data(), e.host.build.size() * 2, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess) ce = cudaMemcpy(e.d_fac, e.host.fac.data(), e.host.fac.size() * 2, cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) { cudaFree(e.d_hdr); cudaFree(e.d_build); cudaFree(e.d_fac); return ce; }
            pc.entries.push_back(std::move(e));
            en = &pc.entries.back();
        }
    }
    BasisParams q{};
    q.x = x; q.out = out; q.n = n; q.ld = ld; q.K = plan.K; q.d = d; q.E = E; q.ncomp = plan.ncomp;
    q.nf = std::max(1, std::min(std::min(d, plan.basis.B), en->host.facw));
    q.facw = en->host.facw; q.tw = tw; q.n_chunks = (int)en->host.hdr.size(); q.max_rows = en->host.max_rows;
    q.hdr = en->d_hdr; q.build = en->d_build; q.fac = en->d_fac;
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    for (int c = 0; c < plan.ncomp; c++) for (int j = 0; j < d; j++) q.orders[c][j] = (int8_t)plan.orders[c][j];
    q.pb = pb; q.substride = substride;
    q.theta = theta; q.nvec = nvec;
    if (theta && warp_exact) {
        const size_t smem_w = (size_t)q.max_rows * kP * (JW == 2 ? 16 : 8);
        const int64_t tiles_w = (n + kP - 1) / kP;
        cudaError_t ew = cudaSuccess;
#define SK_LAUNCH_WE(DD, G)                                                                                               \
    do {                                                                                                                  \
        ew = cudaFuncSetAttribute(smolyak_warp_exact_kernel<DD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w); \
        int per_sm = 1;                                                                                                   \
        if (ew == cudaSuccess) ew = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smolyak_warp_exact_kernel<DD, G>, kP, smem_w); \
        const int grid = (int)std::min<int64_t>(tiles_w, (int64_t)sm_count() * std::max(per_sm, 1));                      \
        if (ew == cudaSuccess) smolyak_warp_exact_kernel<DD, G><<<grid, kP, smem_w, stream>>>(q);                         \
    } while (0)
#define SK_LAUNCH_WED(DD) do { if (JW == 2) SK_LAUNCH_WE(DD, true); else SK_LAUNCH_WE(DD, false); } while (0)
        switch (d) {
        case 1: SK_LAUNCH_WED(1); break; case 2: SK_LAUNCH_WED(2); break; case 3: SK_LAUNCH_WED(3); break; case 4: SK_LAUNCH_WED(4); break;
        case 5: SK_LAUNCH_WED(5); break; case 6: SK_LAUNCH_WED(6); break; case 7: SK_LAUNCH_WED(7); break; case 8: SK_LAUNCH_WED(8); break;
        }
#undef SK_LAUNCH_WED
#undef SK_LAUNCH_WE
        if (ew != cudaSuccess) return ew;
        count_launch();
        return cudaGetLastError();
    }
    if (theta) {
        const size_t rows_b = ((size_t)q.max_rows * JW * kP * 8 + 127) & ~(size_t)127;
        const size_t smem_l = xj_bytes + rows_b;
        const int64_t tiles_l = (n + kP - 1) / kP;
        cudaError_t el = cudaSuccess;
#define SK_LAUNCH_TL(J, G)                                                                                                \
    do {                                                                                                                  \
        el = cudaFuncSetAttribute(smolyak_tile_lincomb_kernel<J, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l); \
        int per_sm = 1;                                                                                                   \
        if (el == cudaSuccess) el = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smolyak_tile_lincomb_kernel<J, G>, kP * kWarps, smem_l); \
        const int grid = (int)std::min<int64_t>(tiles_l, (int64_t)sm_count() * std::max(per_sm, 1));                      \
        if (el == cudaSuccess) smolyak_tile_lincomb_kernel<J, G><<<grid, kP * kWarps, smem_l, stream>>>(q);               \
    } while (0)
#define SK_LAUNCH_TLJ(J) do { if (q.facw > 8) SK_LAUNCH_TL(J, true); else SK_LAUNCH_TL(J, false); } while (0)
        if (JW == 1) SK_LAUNCH_TLJ(1); else if (JW == 2) SK_LAUNCH_TLJ(2); else if (JW == 3) SK_LAUNCH_TLJ(3); else SK_LAUNCH_TLJ(6);
#undef SK_LAUNCH_TLJ
#undef SK_LAUNCH_TL
        if (el != cudaSuccess) return el;
        count_launch();
        return cudaGetLastError();
    }
    static const int no_tma = [] { const char *e = getenv("SK_BASIS_NO_TMA"); return e ? atoi(e) : 0; }();
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof tmap);
    q.use_tma = 0;
    const uint64_t pitch = (uint64_t)ld * E * 8;
    if (!no_tma && pb >= 1 && encode_tiled() && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && pitch % 16 == 0 && pitch < ((uint64_t)1 << 40) &&
        (uint64_t)n * E < ((uint64_t)1 << 32) && (((size_t)pb * E * 8) % 16) == 0) {
        const cuuint64_t gdim[2] = {(cuuint64_t)n * E, (cuuint64_t)plan.K};
        const cuuint64_t gstr[1] = {pitch};
        const cuuint32_t box[2] = {(cuuint32_t)(pb * E), (cuuint32_t)tw};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        q.use_tma = r == CUDA_SUCCESS;
    }
    const size_t rows_bytes = ((size_t)q.max_rows * JW * P * 8 + 127) & ~(size_t)127;
    const size_t smem = xj_bytes + rows_bytes + stage_bytes;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    const int64_t tiles = (n + P - 1) / P;
    cudaError_t e = cudaSuccess;
#define SK_LAUNCH_TBP(J, T, G, D, PL)                                                                                     \
    do {                                                                                                                  \
        e = cudaFuncSetAttribute(smolyak_tile_basis_kernel<J, T, G, D, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        int per_sm = 1;                                                                                                   \
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smolyak_tile_basis_kernel<J, T, G, D, PL>, kP * kWarps, smem); \
        const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * std::max(per_sm, 1));                        \
        if (e == cudaSuccess) smolyak_tile_basis_kernel<J, T, G, D, PL><<<grid, kP * kWarps, smem, stream>>>(q, tmap);    \
    } while (0)
#define SK_LAUNCH_TB(J, T, G, D) SK_LAUNCH_TBP(J, T, G, D, 1)
#define SK_LAUNCH_TBT(J, G) do { if (tw == 8) SK_LAUNCH_TB(J, 8, G, 0); else if (tw == 4) SK_LAUNCH_TB(J, 4, G, 0); else if (tw == 2) SK_LAUNCH_TB(J, 2, G, 0); else SK_LAUNCH_TB(J, 1, G, 0); } while (0)
#define SK_LAUNCH_TBJ(J) do { if (q.facw > 8) SK_LAUNCH_TBT(J, true); else SK_LAUNCH_TBT(J, false); } while (0)
    // full gradient [value, ∂1..∂d] (the plan's fast_mode 1) in d <= 8 dimensions: (T, T') rows, shared prefix products
    const bool grad = plan.fast_mode == 1 && JW == 2 && d >= 1 && d <= 8 && tw <= 4;
    if (two) {
        if (q.facw > 8) SK_LAUNCH_TBP(1, 4, true, 0, 2); else SK_LAUNCH_TBP(1, 4, false, 0, 2);
    } else if (grad) {
        switch (d) {
#define SK_G(D) case D: if (tw == 4) SK_LAUNCH_TB(2, 4, false, D); else if (tw == 2) SK_LAUNCH_TB(2, 2, false, D); else SK_LAUNCH_TB(2, 1, false, D); break;
        SK_G(1) SK_G(2) SK_G(3) SK_G(4) SK_G(5) SK_G(6) SK_G(7) SK_G(8)
#undef SK_G
        }
    } else if (JW == 1) SK_LAUNCH_TBJ(1); else if (JW == 2) SK_LAUNCH_TBJ(2); else if (JW == 3) SK_LAUNCH_TBJ(3); else SK_LAUNCH_TBJ(6);
#undef SK_LAUNCH_TBT
#undef SK_LAUNCH_TBJ
#undef SK_LAUNCH_TB
#undef SK_LAUNCH_TBP
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace skk
