// SURVEY.md §8(f) rank 1 — the step before evaluation in every workflow: fitting θ on the basis' own grid,
//     θ = collocation_matrix(basis) \ y        (docs/src/index.md:92, test/test_smolyak.jl:20-21,
//                                               test/test_generic_api.jl:20)
// The collocation matrix is generated on the device by the library's own basis_at kernels (bit-identical to the
// reference's values) and never leaves it.  Julia's `\` on a square dense matrix is an LU factorisation with partial
// pivoting followed by two triangular solves; this file does the same with its own kernels (no cuSOLVER, no cuBLAS):
//
//   blocked right-looking LU, panels of NB = 64 columns of the column-major K×K matrix:
//     lu_panel_cluster_kernel  a thread-block cluster of up to 8 CTAs factors 8 columns of the panel at a time with the
//                        sub-panel in distributed shared memory: per column an arg-max (first maximal |a|, as LAPACK's
//                        idamax) agreed on through DSMEM, row swap across the panel, reciprocal scaling, rank-1 update
//                        inside the 8 columns (lu_panel_kernel: the same on global memory, for sub-panels beyond 25 600
//                        rows); lu_panel_update_kernel brings the panel's other columns up to date on all SMs
//     lu_swap_kernel     the panel's row interchanges applied to the columns left and right of it
//     lu_trsm_kernel     U12 = L11⁻¹·A12 (unit lower 64×64 triangle in shared memory, one matrix column per thread)
//     lu_gemm_kernel     A22 −= L21·U12: the (2/3)·K³ flops of the factorisation, on the FP64 tensor path —
//                        mma.sync.m8n8k4.f64 (SASS DMMA), 64×64 tiles, operands staged in shared memory with the
//                        conflict-free pitch ≡ 4 (mod 16) doubles the coefficient-block kernel uses (sk_block.cu)
//   then P·y, L·z = P·y and U·θ = z block by block (64-row triangle per CTA + a row-parallel update).
//
// An initialisation-time step, not the hot path; K ≤ 46 340 (the matrix must fit the device).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cooperative_groups.h>

#include "sk_launch.h"

namespace cg = cooperative_groups;
using skh::set_error;

namespace {

// global -> shared copies that do not pass through registers: every thread issues all of its copies, then waits once, so a
// tile costs ONE memory round trip instead of one per loop iteration (these kernels are latency-bound, not bandwidth-bound)
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid = true) {
    const int sz = valid ? 8 : 0;             // src-size 0: the destination is zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int NB = 64;           // panel width
#ifndef SK_LU_IB
#define SK_LU_IB 8
#endif
constexpr int IB = SK_LU_IB;     // sub-panel width; measured LU times at K = 2 561 / 4 096 / 8 192: IB = 8 14.3 / 24.8 / 82.1 ms,
                                 // 16: 13.7 / 24.7 / 82.3, 32: 15.9 / 28.1 / 178 — the per-column latency dominates either way
constexpr int kPitch = 68;       // shared-memory pitch of the GEMM operand tiles (≡ 4 mod 16: conflict-free fragment loads)

// ---- panel factorisation: columns [k0, k0+nb) of rows [k0, K); one CTA of 1024 threads
// (sub-panel [jb, jb+ib) of the panel: pivot search, swap across the WHOLE panel, scaling, rank-1 updates inside the
// sub-panel only; the panel's columns to the right are brought up to date by lu_panel_update_kernel on all SMs)
__global__ void __launch_bounds__(1024) lu_panel_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int nb, int jb, int ib,
                                                        int *__restrict__ ipiv, int *__restrict__ info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ double s_rcp;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = jb; j < jb + ib; j++) {
        const int col = k0 + j;
        double *cj = A + (size_t)col * lda;
        // arg-max of |a| over rows col..K-1 (first maximal element)
        double best = -1.0;
        int bi = col;
        for (int r = col + tid; r < K; r += blockDim.x) {
            const double v = fabs(cj[r]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            best = lane < (int)(blockDim.x >> 5) ? s_val[lane] : -1.0;
            bi = lane < (int)(blockDim.x >> 5) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                s_piv = bi;
                ipiv[col] = bi;
                if (best == 0.0 && *info == 0) *info = col + 1;      // exactly singular (LAPACK's info > 0)
                s_rcp = best == 0.0 ? 0.0 : 1.0 / cj[bi];
            }
        }
        __syncthreads();
        const int piv = s_piv;
        // swap rows col <-> piv inside the panel
        if (piv != col && tid < nb) {
            double *c = A + (size_t)(k0 + tid) * lda;
            const double t = c[col]; c[col] = c[piv]; c[piv] = t;
        }
        __syncthreads();
        // scale the column below the diagonal, then rank-1 update of the panel's remaining columns
        const double rcp = s_rcp;
        for (int r = col + 1 + tid; r < K; r += blockDim.x) cj[r] *= rcp;
        __syncthreads();
        const int rem = jb + ib - j - 1, rows = K - col - 1;
        for (int64_t e = tid; e < (int64_t)rem * rows; e += blockDim.x) {
            const int c = (int)(e / rows), r = (int)(e - (int64_t)c * rows) + col + 1;
            double *cc = A + (size_t)(col + 1 + c) * lda;
            cc[r] -= cj[r] * cc[col];
        }
        __syncthreads();
    }
    // rows [jb, jb+ib) of the panel's columns to the right become rows of U: u = L_bb^{-1} a_b, one column per thread,
    // in place (lu_panel_update_kernel then reads them)
    const int right = nb - jb - ib, rb = k0 + jb;
    if (tid < right) {
        double *c = A + (size_t)(rb + ib + tid) * lda + rb;
        double x[IB];
        for (int i = 0; i < ib; i++) {
            double sacc = c[i];
            for (int jj = 0; jj < i; jj++) sacc -= A[(size_t)(rb + jj) * lda + rb + i] * x[jj];
            x[i] = sacc;
        }
        for (int i = 1; i < ib; i++) c[i] = x[i];
    }
}

// ---- the same sub-panel factorisation on a thread-block CLUSTER: the ib <= 8 columns of the sub-panel (rows rb..K-1) live
// in the shared memory of up to 8 CTAs, `slice` rows each, so the per-column steps (arg-max, swap, scaling, rank-1 update)
// run at shared-memory latency instead of one L2 round trip each — the global-memory kernel above spends 6.6 µs per
// column.  Per column: every CTA finds the first maximal |a| of its rows and publishes it; after one cluster barrier all
// CTAs read the candidates through distributed shared memory and agree on the pivot (LAPACK's first-maximum rule); rank 0,
// which owns the diagonal rows, exchanges the pivot row with its own through DSMEM and publishes it; after a second cluster
// barrier every CTA scales and updates its own rows.  The panel's other columns are swapped in global memory as before.
__device__ __forceinline__ void argmax_combine(double &best, int &bi, double ov, int oi) {
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
}
struct PanelCand { double val; int idx; int pad; };
__global__ void __launch_bounds__(1024) lu_panel_cluster_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int nb, int jb, int ib,
                                                               int slice, int *__restrict__ ipiv, int *__restrict__ info) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) double psm[];
    double *s = psm;                                                        // [IB][slice], column-major
    PanelCand *s_cand = reinterpret_cast<PanelCand *>(psm + (size_t)IB * slice);   // this CTA's candidate (read by the cluster)
    double *s_u = reinterpret_cast<double *>(s_cand + 1);                   // [IB] pivot row (rank 0's copy is the published one)
    double *s_ul = s_u + IB;                                                // [IB] local copy
    double *s_wv = s_ul + IB;                                               // [32] per-warp candidates
    int *s_wi = reinterpret_cast<int *>(s_wv + 32);                         // [32]
    int *s_piv = s_wi + 32;                                                 // [1]
    int *s_pivs = s_piv + 1;                                                // [IB] pivots of the sub-panel (rank 0)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rb = k0 + jb, rows = K - rb;
    const int r_lo = rank * slice, r_hi = min(rows, r_lo + slice);          // this CTA's rows, relative to rb
    const int nth = (int)blockDim.x, nwarp = nth >> 5;
    for (int c = 0; c < ib; c++) {
        const double *g = A + (size_t)(rb + c) * lda + rb;
        for (int r = r_lo + tid; r < r_hi; r += nth) cp_async8(&s[c * slice + r - r_lo], g + r);
    }
    if (rank == 0 && tid < IB) s_pivs[tid] = tid;
    cp_async_wait_all();
    __syncthreads();
    for (int j = 0; j < ib; j++) {
        // 1. first maximal |a| among this CTA's rows >= j of column j
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int r = r_lo + tid; r < r_hi; r += nth) {
            if (r < j) continue;
            const double v = fabs(s[j * slice + r - r_lo]);
            if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) argmax_combine(best, bi, __shfl_down_sync(0xffffffffu, best, o), __shfl_down_sync(0xffffffffu, bi, o));
        if (lane == 0) { s_wv[wid] = best; s_wi[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            best = lane < nwarp ? s_wv[lane] : -1.0;
            bi = lane < nwarp ? s_wi[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) argmax_combine(best, bi, __shfl_down_sync(0xffffffffu, best, o), __shfl_down_sync(0xffffffffu, bi, o));
            if (lane == 0) { s_cand->val = best; s_cand->idx = bi; }
        }
        cluster.sync();
        // 2. the pivot: warp 0 of every CTA reads all candidates through distributed shared memory
        if (wid == 0) {
            best = -1.0; bi = 0x7fffffff;
            if (lane < C) {
                const PanelCand *rc = cluster.map_shared_rank(s_cand, lane);
                best = rc->val; bi = rc->idx;
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) argmax_combine(best, bi, __shfl_down_sync(0xffffffffu, best, o), __shfl_down_sync(0xffffffffu, bi, o));
            best = __shfl_sync(0xffffffffu, best, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
            if (bi == 0x7fffffff) bi = j;                                   // a column of NaNs: no interchange (as idamax returns the first)
            if (lane == 0) { *s_piv = bi; if (rank == 0) s_pivs[j] = bi; }
            // 3. rank 0 owns the diagonal rows (j < ib <= slice): exchange row j with the pivot row, publish the pivot row
            if (rank == 0) {
                const int po = bi / slice;
                if (lane < ib) {
                    double *mine = s + lane * slice + j;
                    double *theirs = cluster.map_shared_rank(s, po) + lane * slice + (bi - po * slice);
                    const double a = *mine, b = *theirs;
                    *theirs = a; *mine = b;
                    s_u[lane] = b;
                }
                if (lane == 0) {
                    ipiv[rb + j] = rb + bi;
                    if (best == 0.0 && *info == 0) *info = rb + j + 1;      // exactly singular (LAPACK's info > 0)
                }
            }
        }
        cluster.sync();
        // 4. scale column j and update columns j+1.. of this CTA's rows below the diagonal
        if (tid < ib) s_ul[tid] = cluster.map_shared_rank(s_u, 0)[tid];
        __syncthreads();
        const double pv = s_ul[j], rcp = pv == 0.0 ? 0.0 : 1.0 / pv;
        for (int r = r_lo + tid; r < r_hi; r += nth) {
            if (r <= j) continue;
            const int o = r - r_lo;
            const double l = s[j * slice + o] * rcp;
            s[j * slice + o] = l;
            for (int c = j + 1; c < ib; c++) s[c * slice + o] -= l * s_ul[c];
        }
        // (the next column's arg-max reads the rows this thread has just updated: same row-to-thread map, no barrier needed)
    }
    for (int c = 0; c < ib; c++) {
        double *g = A + (size_t)(rb + c) * lda + rb;
        for (int r = r_lo + tid; r < r_hi; r += nth) g[r] = s[c * slice + r - r_lo];
    }
    __syncthreads();
    // the panel's other columns (rank 0, one thread per column): the ib row interchanges, applied in order on a register
    // copy of the <= 2·ib rows they touch (all loads in flight at once, instead of one L2 round trip per interchange);
    // for the columns to the right, rows [jb, jb+ib) then become rows of U: u = L_bb⁻¹ a_b with L_bb from shared memory
    if (rank == 0 && tid < nb && (tid < jb || tid >= jb + ib)) {
        double *col = A + (size_t)(k0 + tid) * lda + rb;
        double top[IB], low[IB];
        int pv[IB];
#pragma unroll
        for (int i = 0; i < IB; i++) { pv[i] = i < ib ? s_pivs[i] : i; top[i] = i < ib ? col[i] : 0.0; }
#pragma unroll
        for (int i = 0; i < IB; i++) low[i] = (i < ib && pv[i] >= ib) ? col[pv[i]] : 0.0;
        // value currently stored in relative row r: diagonal rows live in top[], pivot rows below in low[] (slot of the FIRST
        // interchange that touches that row)
#pragma unroll
        for (int i = 0; i < IB; i++) {
            if (i < ib && pv[i] != i) {
                if (pv[i] < ib) {                                     // both rows inside the diagonal block
                    double t = top[i];
#pragma unroll
                    for (int m = 0; m < IB; m++) if (m == pv[i]) { top[i] = top[m]; top[m] = t; }
                } else {
                    int slot = i;
#pragma unroll
                    for (int m = IB - 1; m >= 0; m--) if (m < i && pv[m] == pv[i]) slot = m;      // same row picked before
                    double t = top[i];
#pragma unroll
                    for (int m = 0; m < IB; m++) if (m == slot) { top[i] = low[m]; low[m] = t; }
                }
            }
        }
        if (tid >= jb + ib) {                                         // right of the sub-panel: rows of U
#pragma unroll
            for (int i = 1; i < IB; i++) {
                if (i < ib) {
                    double sacc = top[i];
#pragma unroll
                    for (int jj = 0; jj < IB; jj++) if (jj < i) sacc -= s[jj * slice + i] * top[jj];
                    top[i] = sacc;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < IB; i++) if (i < ib) col[i] = top[i];
#pragma unroll
        for (int i = 0; i < IB; i++) {
            bool first = i < ib && pv[i] >= ib;
#pragma unroll
            for (int m = 0; m < IB; m++) if (m < i && pv[m] == pv[i]) first = false;
            if (first) col[pv[i]] = low[i];
        }
    }
    cluster.sync();           // nobody reads remote shared memory any more
}

// ---- the panel's columns right of a factored sub-panel: a_r -= L_rb u with u = the ib rows of U the panel kernel
// solved in place; grid: (row tiles of 256, columns)
__global__ void __launch_bounds__(256) lu_panel_update_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int jb, int ib) {
    __shared__ double u[IB];
    const int c = k0 + jb + ib + blockIdx.y;            // column being updated
    const int rb = k0 + jb;                             // first row / column of the sub-panel
    double *col = A + (size_t)c * lda;
    if (threadIdx.x < ib) u[threadIdx.x] = col[rb + threadIdx.x];      // rows of U, solved by the panel kernel
    __syncthreads();
    const int r = rb + ib + blockIdx.x * 256 + threadIdx.x;
    if (r < K) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < IB; t++) if (t < ib) s += A[(size_t)(rb + t) * lda + r] * u[t];
        col[r] -= s;
    }
}

// ---- row interchanges of the panel applied to every other column (one thread per column, swaps in order)
__global__ void lu_swap_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int nb, const int *__restrict__ ipiv) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K - nb) return;
    if (c >= k0) c += nb;                           // skip the panel's own columns
    double *col = A + (size_t)c * lda;
    for (int j = 0; j < nb; j++) {
        const int p = ipiv[k0 + j];
        if (p != k0 + j) { const double t = col[k0 + j]; col[k0 + j] = col[p]; col[p] = t; }
    }
}

// ---- the same interchanges as ONE permutation of the <= 2·nb rows they touch: lu_perm_kernel (one warp) composes the nb
// swaps into (row, source row) pairs, lu_permute_kernel gathers every column's affected entries and scatters them — two
// global round trips per column instead of 2·nb dependent ones (46 -> a few µs per panel at K = 2 561)
struct PanelPerm { int n; int row[2 * NB]; int src[2 * NB]; };
__global__ void __launch_bounds__(32) lu_perm_kernel(int k0, int nb, const int *__restrict__ ipiv, PanelPerm *__restrict__ out) {
    __shared__ int row[2 * NB], cur[2 * NB];      // affected rows; cur[i] = original row whose content sits in row[i] now
    __shared__ int n;
    __shared__ int piv[NB];
    const int lane = threadIdx.x;
    for (int i = lane; i < nb; i += 32) { row[i] = k0 + i; cur[i] = k0 + i; piv[i] = ipiv[k0 + i]; }
    if (lane == 0) n = nb;
    __syncwarp();
    for (int j = 0; j < nb; j++) {
        const int p = piv[j];
        if (p == k0 + j) continue;                 // warp-uniform
        int slot = -1;
        if (p < k0 + nb) slot = p - k0;
        else {
            for (int i = nb + lane; i < n; i += 32) if (row[i] == p) slot = i;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) slot = max(slot, __shfl_xor_sync(0xffffffffu, slot, o));
            if (slot < 0) {
                slot = n;
                if (lane == 0) { row[slot] = p; cur[slot] = p; n = slot + 1; }
            }
        }
        __syncwarp();
        if (lane == 0) { const int t = cur[j]; cur[j] = cur[slot]; cur[slot] = t; }
        __syncwarp();
    }
    if (lane == 0) out->n = n;
    for (int i = lane; i < n; i += 32) { out->row[i] = row[i]; out->src[i] = cur[i]; }
}
__global__ void __launch_bounds__(2 * NB) lu_permute_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int nb, const PanelPerm *__restrict__ pp) {
    const int n = pp->n, i = threadIdx.x;
    const int dst = i < n ? pp->row[i] : 0, src = i < n ? pp->src[i] : 0;
    for (int c = blockIdx.x; c < K - nb; c += gridDim.x) {
        double *col = A + (size_t)(c >= k0 ? c + nb : c) * lda;      // skip the panel's own columns
        double v = 0.0;
        if (i < n && src != dst) v = col[src];
        __syncthreads();
        if (i < n && src != dst) col[dst] = v;
        __syncthreads();
    }
}

// ---- U12 = L11^{-1} A12: the nb×nb unit lower triangle in shared memory, one column of A12 per thread
__global__ void __launch_bounds__(64) lu_trsm_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int nb) {
    __shared__ double L[NB][NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const int r = e % nb, c = e / nb;
        cp_async8(&L[r][c], &A[(size_t)(k0 + c) * lda + k0 + r]);
    }
    cp_async_wait_all();
    __syncthreads();
    const int c = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    double *col = A + (size_t)c * lda + k0;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = i < nb ? col[i] : 0.0;
    // column-oriented: after step j, x[j] is final and every x[i], i > j, has had its j-th term subtracted — the same
    // subtractions in the same order per entry as the row-oriented loop, but NB-1-j independent operations per step
#pragma unroll
    for (int j = 0; j < NB - 1; j++) {
#pragma unroll
        for (int i = j + 1; i < NB; i++) x[i] -= L[i][j] * x[j];
    }
#pragma unroll
    for (int i = 0; i < NB; i++) if (i < nb) col[i] = x[i];
}

// ---- A22 -= L21 · U12 on the FP64 tensor path: CTA tile 64×64, four warps of 32×32 (4×4 mma tiles), k = nb
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) lu_gemm_kernel(double *__restrict__ A, int64_t lda, int K, int k0, int nb) {
    extern __shared__ __align__(16) double gemm_smem[];
    double *As = gemm_smem;                  // As[k][row]: L21 tile, rows contiguous (as in memory)
    double *Bs = gemm_smem + NB * kPitch;    // Bs[col][k]: U12 tile, k contiguous (as in memory)
    const int r0 = k0 + nb + blockIdx.x * 64, c0 = k0 + nb + blockIdx.y * 64;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int e = tid; e < 64 * NB; e += 128) {
        const int i = e & 63, k = e >> 6;
        const bool ok = k < nb && r0 + i < K;
        cp_async8(&As[k * kPitch + i], ok ? &A[(size_t)(k0 + k) * lda + r0 + i] : A, ok);
    }
    for (int e = tid; e < 64 * NB; e += 128) {
        const int k = e & 63, j = e >> 6;
        const bool ok = k < nb && c0 + j < K;
        cp_async8(&Bs[j * kPitch + k], ok ? &A[(size_t)(c0 + j) * lda + k0 + k] : A, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    const int wr = (w & 1) * 32, wc = (w >> 1) * 32;      // this warp's 32×32 corner inside the tile
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll 4
    for (int k = 0; k < NB; k += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = As[(k + fk) * kPitch + wr + 8 * i + fr];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = Bs[(wc + 8 * j + fr) * kPitch + k + fk];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = r0 + wr + 8 * i + fr;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int c = c0 + wc + 8 * j + 2 * fk + q;
                if (r < K && c < K) A[(size_t)c * lda + r] -= acc[i][j][q];
            }
        }
}

// ---- right-hand sides, row-major Y[K][nrhs]
__global__ void rhs_permute_kernel(const double *__restrict__ in, double *__restrict__ out, int K, int nrhs, const int *__restrict__ ipiv) {
    // the interchanges are sequential by definition: one thread per right-hand-side column applies them in order
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nrhs) return;
    for (int i = 0; i < K; i++) out[(size_t)i * nrhs + v] = in[(size_t)i * nrhs + v];
    for (int i = 0; i < K; i++) {
        const int p = ipiv[i];
        if (p != i) { const double t = out[(size_t)i * nrhs + v]; out[(size_t)i * nrhs + v] = out[(size_t)p * nrhs + v]; out[(size_t)p * nrhs + v] = t; }
    }
}
// one diagonal block: forward (unit lower) or backward (upper, non-unit) substitution; one thread per right-hand side
__global__ void rhs_tri_kernel(const double *__restrict__ A, int64_t lda, int k0, int nb, double *__restrict__ Y, int nrhs, int upper) {
    __shared__ double T[NB][NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const int r = e % nb, c = e / nb;
        T[r][c] = A[(size_t)(k0 + c) * lda + k0 + r];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nrhs; v += blockDim.x) {
        double *y = Y + (size_t)k0 * nrhs + v;
        if (!upper) {
            for (int i = 1; i < nb; i++) {
                double s = y[(size_t)i * nrhs];
                for (int j = 0; j < i; j++) s -= T[i][j] * y[(size_t)j * nrhs];
                y[(size_t)i * nrhs] = s;
            }
        } else {
            for (int i = nb - 1; i >= 0; i--) {
                double s = y[(size_t)i * nrhs];
                for (int j = i + 1; j < nb; j++) s -= T[i][j] * y[(size_t)j * nrhs];
                y[(size_t)i * nrhs] = s / T[i][i];
            }
        }
    }
}
// Y[rows r_lo..r_hi) -= A[rows, k0..k0+nb) · Y[k0..k0+nb): one thread per (row, right-hand side)
__global__ void rhs_update_kernel(const double *__restrict__ A, int64_t lda, int k0, int nb, int r_lo, int r_hi, double *__restrict__ Y, int nrhs) {
    extern __shared__ double yb[];               // [nb][nrhs]
    for (int e = threadIdx.x; e < nb * nrhs; e += blockDim.x) yb[e] = Y[(size_t)k0 * nrhs + e];
    __syncthreads();
    const int64_t total = (int64_t)(r_hi - r_lo) * nrhs;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(e / (r_hi - r_lo)), r = r_lo + (int)(e % (r_hi - r_lo));      // consecutive threads: consecutive rows
        double s = 0.0;
        for (int k = 0; k < nb; k++) s += A[(size_t)(k0 + k) * lda + r] * yb[k * nrhs + v];
        Y[(size_t)r * nrhs + v] -= s;
    }
}

constexpr size_t kPanelSmemMax = 208 * 1024;
inline size_t panel_cluster_smem(int slice) { return ((size_t)IB * slice + 2 + 2 * IB + 32) * sizeof(double) + (48 + IB) * sizeof(int); }

constexpr size_t kLuScratchBytes = sizeof(PanelPerm);

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

}  // namespace

#define CKS(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return SK_ERR_CUDA; } \
    } while (0)

namespace skk {

// In-place LU with partial pivoting of the column-major K×K matrix A (leading dimension lda); ipiv[i] = row swapped with
// row i (0-based), *info = index + 1 of the first exactly-zero pivot or 0.  ms_gemm (optional): device time of the trailing
// updates (the DMMA part).
// `scratch` (optional): kLuScratchBytes of device memory for the composed row permutation of a panel; without it the
// interchanges are applied one by one.
cudaError_t lu_factor(double *A, int64_t lda, int K, int *ipiv, int *info, cudaStream_t stream, float *ms_gemm = nullptr, void *scratch = nullptr) {
    constexpr size_t kGemmSmem = (size_t)2 * NB * kPitch * sizeof(double);      // 69 632 bytes
    cudaError_t e = cudaFuncSetAttribute(lu_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lu_panel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSmemMax);
    static const int no_cluster = [] { const char *v = getenv("SK_LU_NO_CLUSTER"); return v ? atoi(v) : 0; }();
    PanelPerm *perm = no_cluster ? nullptr : static_cast<PanelPerm *>(scratch);
    if (e == cudaSuccess) e = cudaMemsetAsync(info, 0, sizeof(int), stream);
    for (int k0 = 0; k0 < K && e == cudaSuccess; k0 += NB) {
        const int nb = std::min(NB, K - k0);
        for (int jb = 0; jb < nb; jb += IB) {
            const int ib = std::min(IB, nb - jb);
            // sub-panel in the shared memory of a cluster of 1..8 CTAs (about 512 rows each); beyond 8 × 3 200 rows, or with
            // SK_LU_NO_CLUSTER=1 (experiments), the global-memory kernel
            const int rows = K - (k0 + jb);
            // about 512 rows per CTA: measured faster than one CTA holding the whole sub-panel (36 vs 27 µs per sub-panel at
            // K = 2 561) — a 160 KB dynamic allocation between kernels with small ones makes the SMs reconfigure their
            // shared-memory carve-out at every launch
            int C = 1;
            while (C < 8 && rows > C * 512) C *= 2;
            const int slice = (((rows + C - 1) / C) + 31) & ~31;
            const size_t psmem = panel_cluster_smem(slice);
            if (!no_cluster && psmem <= kPanelSmemMax) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)C); cfg.blockDim = dim3(slice <= 512 ? 256 : slice <= 1024 ? 512 : 1024); cfg.dynamicSmemBytes = psmem; cfg.stream = stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                e = cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel, A, lda, K, k0, nb, jb, ib, slice, ipiv, info);
                if (e != cudaSuccess) {      // no cluster launch on this device / partition: the global-memory kernel does the same
                    (void)cudaGetLastError();
                    lu_panel_kernel<<<1, 1024, 0, stream>>>(A, lda, K, k0, nb, jb, ib, ipiv, info);
                }
            } else
            lu_panel_kernel<<<1, 1024, 0, stream>>>(A, lda, K, k0, nb, jb, ib, ipiv, info);
            const int right = nb - jb - ib, below = K - (k0 + jb + ib);
            if (right > 0) {
                dim3 ug((unsigned)std::max(1, (below + 255) / 256), (unsigned)right);
                lu_panel_update_kernel<<<ug, 256, 0, stream>>>(A, lda, K, k0, jb, ib);
                count_launch();
            }
            count_launch();
        }
        if (K - nb > 0) {
            if (perm) {
                lu_perm_kernel<<<1, 32, 0, stream>>>(k0, nb, ipiv, perm);
                lu_permute_kernel<<<std::min(K - nb, 2 * 148), 2 * NB, 0, stream>>>(A, lda, K, k0, nb, perm);
                count_launch();
            } else
                lu_swap_kernel<<<(K - nb + 127) / 128, 128, 0, stream>>>(A, lda, K, k0, nb, ipiv);
        }
        const int rest = K - k0 - nb;
        if (rest > 0) {
            lu_trsm_kernel<<<(rest + 63) / 64, 64, 0, stream>>>(A, lda, K, k0, nb);
            dim3 grid((unsigned)((rest + 63) / 64), (unsigned)((rest + 63) / 64));
            cudaEvent_t g0 = nullptr, g1 = nullptr;
            if (ms_gemm) { cudaEventCreate(&g0); cudaEventCreate(&g1); cudaEventRecord(g0, stream); }
            lu_gemm_kernel<<<grid, 128, kGemmSmem, stream>>>(A, lda, K, k0, nb);
            if (ms_gemm) {      // measurement only: serialises the factorisation
                cudaEventRecord(g1, stream); cudaEventSynchronize(g1);
                float t = 0; cudaEventElapsedTime(&t, g0, g1); *ms_gemm += t;
                cudaEventDestroy(g0); cudaEventDestroy(g1);
            }
            count_launch(2);
        }
        count_launch(1);      // the swap kernel
        e = cudaGetLastError();
    }
    return e;
}

// solves A·X = Y for the factored A; Y (device, row-major [K][nrhs]) is overwritten by X; `work` has K·nrhs doubles
cudaError_t lu_solve(const double *A, int64_t lda, int K, const int *ipiv, double *Y, double *work, int nrhs, cudaStream_t stream) {
    rhs_permute_kernel<<<(nrhs + 63) / 64, 64, 0, stream>>>(Y, work, K, nrhs, ipiv);
    count_launch();
    const size_t sm = (size_t)NB * nrhs * sizeof(double);
    if (sm > 48 * 1024) return cudaErrorInvalidValue;
    for (int k0 = 0; k0 < K; k0 += NB) {               // L z = P y
        const int nb = std::min(NB, K - k0);
        rhs_tri_kernel<<<1, 128, 0, stream>>>(A, lda, k0, nb, work, nrhs, 0);
        count_launch();
        if (k0 + nb < K) {
            const int64_t tot = (int64_t)(K - k0 - nb) * nrhs;
            rhs_update_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, 1184), 256, sm, stream>>>(A, lda, k0, nb, k0 + nb, K, work, nrhs);
            count_launch();
        }
    }
    for (int k0 = ((K - 1) / NB) * NB; k0 >= 0; k0 -= NB) {      // U x = z
        const int nb = std::min(NB, K - k0);
        rhs_tri_kernel<<<1, 128, 0, stream>>>(A, lda, k0, nb, work, nrhs, 1);
        count_launch();
        if (k0 > 0) {
            const int64_t tot = (int64_t)k0 * nrhs;
            rhs_update_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, 1184), 256, sm, stream>>>(A, lda, k0, nb, 0, k0, work, nrhs);
            count_launch();
        }
    }
    cudaError_t e = cudaMemcpyAsync(Y, work, (size_t)K * nrhs * sizeof(double), cudaMemcpyDeviceToDevice, stream);
    return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace skk

extern "C" int sk_collocation_solve(const sk_plan *plan, const double *y, int nrhs, double *theta, uint32_t flags, void *stream_) {
    if (!plan || !y || !theta) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    if (nrhs < 1) SK_FAIL(SK_ERR_ARGUMENT, "nrhs < 1");
    if (nrhs > 96) SK_FAIL(SK_ERR_UNSUPPORTED, "more than 96 right-hand sides: solve them in groups");
    if (plan->E != 1) SK_FAIL(SK_ERR_ARGUMENT, "collocation solve needs a plan without derivatives");
    const int64_t K = plan->K;
    if (K * K >= ((int64_t)1 << 31)) SK_FAIL(SK_ERR_UNSUPPORTED, "dimension(basis) above 46340: the dense factorisation is limited to 32-bit indexing");
    int prev = 0;
    CKS(cudaGetDevice(&prev));
    CKS(cudaSetDevice(plan->device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool uni = plan->basis.kind == SK_BASIS_CHEBYSHEV;
    const int d = uni ? 1 : plan->basis.d;

    // the basis' own grid (host, correctly rounded; sk_grid) → device
    std::vector<double> grid((size_t)K * d);
    int rc = sk_grid(&plan->basis, plan->has_tr ? plan->tr : nullptr, grid.data());
    if (rc != SK_OK) return rc;
    DevBuf dx, dA, dB, dwork, dpiv, dinfo, dscr;
    CKS(dscr.alloc(kLuScratchBytes));
    CKS(dx.alloc(grid.size() * sizeof(double)));
    CKS(dA.alloc((size_t)K * K * sizeof(double)));
    CKS(dB.alloc((size_t)K * nrhs * sizeof(double)));
    CKS(dwork.alloc((size_t)K * nrhs * sizeof(double)));
    CKS(dpiv.alloc((size_t)K * sizeof(int)));
    CKS(dinfo.alloc(sizeof(int)));
    CKS(cudaMemcpyAsync(dx.p, grid.data(), grid.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    // collocation matrix, column-major K×K: A[i + k·K] = b_k(x_i)   (generic_api.jl:217-227)
    rc = sk_basis_at(plan, (const double *)dx.p, K, (double *)dA.p, K, SK_MEM_DEVICE, stream);
    if (rc != SK_OK) return rc;
    // right-hand sides y[K][nrhs] (the layout of Vector{SVector{nrhs}}) stay row-major on the device
    CKS(cudaMemcpyAsync(dB.p, y, (size_t)K * nrhs * sizeof(double), (flags & SK_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    CKS(skk::lu_factor((double *)dA.p, K, (int)K, (int *)dpiv.p, (int *)dinfo.p, stream, nullptr, dscr.p));
    int info = 0;
    CKS(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CKS(cudaStreamSynchronize(stream));
    if (info != 0) SK_FAIL(SK_ERR_ARGUMENT, "collocation matrix is singular (LU pivot " + std::to_string(info) + " is zero)");   // Julia: SingularException
    CKS(skk::lu_solve((const double *)dA.p, K, (int)K, (const int *)dpiv.p, (double *)dB.p, (double *)dwork.p, nrhs, stream));
    CKS(cudaMemcpyAsync(theta, dB.p, (size_t)K * nrhs * sizeof(double), (flags & SK_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    CKS(cudaStreamSynchronize(stream));
    return SK_OK;
}

// Measurement helper: factors a K×K matrix with the library's LU `reps` times (device pointer, overwritten) and returns the
// mean device time; *tflops = (2/3)·K³ / time.  Used by profiles/run_solve.py.
extern "C" int sk_lu_benchmark(double *A_dev, int64_t K, int reps, double *ms, double *tflops, double *ms_gemm) {
    if (!A_dev || K < 1 || K > 46340 || reps < 1 || !ms || !tflops) SK_FAIL(SK_ERR_ARGUMENT, "bad argument");
    DevBuf copy, piv, info, scr;
    CKS(scr.alloc(kLuScratchBytes));
    CKS(copy.alloc((size_t)K * K * sizeof(double)));
    CKS(piv.alloc((size_t)K * sizeof(int)));
    CKS(info.alloc(sizeof(int)));
    CKS(cudaMemcpy(copy.p, A_dev, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToDevice));
    cudaEvent_t e0, e1;
    CKS(cudaEventCreate(&e0)); CKS(cudaEventCreate(&e1));
    float total = 0;
    for (int r = 0; r < reps; r++) {
        CKS(cudaMemcpy(A_dev, copy.p, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToDevice));
        CKS(cudaEventRecord(e0, nullptr));
        const auto h0 = std::chrono::steady_clock::now();
        CKS(skk::lu_factor(A_dev, K, (int)K, (int *)piv.p, (int *)info.p, nullptr, nullptr, scr.p));
        const auto h1 = std::chrono::steady_clock::now();
        if (getenv("SK_LU_TRACE")) fprintf(stderr, "lu_factor K=%lld: host enqueue %.3f ms\n", (long long)K, std::chrono::duration<double, std::milli>(h1 - h0).count());
        CKS(cudaEventRecord(e1, nullptr));
        CKS(cudaEventSynchronize(e1));
        float t = 0;
        CKS(cudaEventElapsedTime(&t, e0, e1));
        total += t;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms = total / reps;
    *tflops = (2.0 / 3.0) * (double)K * K * K / (*ms * 1e-3) * 1e-12;
    if (ms_gemm) {
        float g = 0;
        CKS(cudaMemcpy(A_dev, copy.p, (size_t)K * K * sizeof(double), cudaMemcpyDeviceToDevice));
        CKS(skk::lu_factor(A_dev, K, (int)K, (int *)piv.p, (int *)info.p, nullptr, &g, scr.p));
        CKS(cudaDeviceSynchronize());
        *ms_gemm = g;
    }
    return SK_OK;
}
