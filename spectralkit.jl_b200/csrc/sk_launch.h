// Kernel launchers (implemented in sk_kernels.cu / sk_nested.cu), called by the C-ABI layer.
// All pointers are device pointers on the current device; all launches go to `stream`.
#pragma once
#include <cuda_runtime.h>

#include "sk_internal.h"

namespace skk {

// ---- univariate Chebyshev ∘ transformation
cudaError_t uni_lincomb(int N, const sk_transform &tr, int order, bool exact, const double *theta, int nvec,
                        const double *x, int64_t n, double *out, cudaStream_t stream);
cudaError_t uni_basis_at(int N, const sk_transform &tr, int order, const double *x, int64_t n, double *out,
                         int64_t ld, cudaStream_t stream);

// ---- Smolyak, reference operation order (direct form, index table)
bool smolyak_direct_fits(const sk_plan &plan, bool for_basis);      // tables of 32 points fit in shared memory
cudaError_t smolyak_direct_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n,
                                   double *out, cudaStream_t stream);
cudaError_t smolyak_direct_basis_at(const sk_plan &plan, const double *x, int64_t n, double *out, int64_t ld,
                                    cudaStream_t stream);

// ---- Smolyak basis_at / collocation_matrix: chunked tables, TMA tensor-map stores (sk_basis.cu); any basis, any ∂ request
cudaError_t smolyak_tile_basis_at(const sk_plan &plan, const double *x, int64_t n, double *out, int64_t ld, cudaStream_t stream);
void smolyak_tile_basis_forget(const sk_plan *plan);
// bit-identical linear_combination (reference order) on the same chunked tables: any basis size
bool smolyak_tile_lincomb_supported(const sk_plan &plan, int nvec);
bool smolyak_warp_exact_serves(const sk_plan &plan, int nvec);      // values / full gradient, one vector, d <= 8: one thread per point
cudaError_t smolyak_tile_lincomb_exact(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out, cudaStream_t stream);

// ---- Smolyak, nested run-structured evaluation (values, or value + full gradient)
bool smolyak_nested_supported(const sk_plan &plan);
cudaError_t smolyak_nested_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n,
                                   double *out, cudaStream_t stream);
// closed-form FP64 flop count per point of the nested kernel for this plan
double smolyak_nested_flops(const sk_plan &plan);

// ---- Smolyak, nested evaluation with compile-time unrolled leaves (sk_leaf.cu): the headline kernel
bool smolyak_leaf_supported(sk_plan &plan);      // also fixes plan.leaf_LL / plan.leaf_cta
int smolyak_leaf_cta(const sk_plan &plan);
void smolyak_leaf_units(const sk_plan &plan, std::vector<uint32_t> &units);
int64_t smolyak_leaf_theta_map(const sk_plan &plan, std::vector<uint32_t> *map);
cudaError_t smolyak_leaf_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n,
                                 double *out, cudaStream_t stream);
double smolyak_leaf_flops(const sk_plan &plan);

// fused gather: peer buffers for the calling thread's next leaf launch (sk_leaf.cu; used by sk_multi.cu)
void smolyak_leaf_set_peers(int n, double *const *ptrs);
bool smolyak_leaf_peers_used();
bool smolyak_leaf_serves_peers(const sk_plan &plan);

// shared-memory slot of every coefficient of Leaf<T, R> under the layout of leaf_end (sk_leaf_impl.cuh); returns the next slot
int smolyak_leaf_map_rec(const skh::SmolyakSet &S, const std::vector<int> &blk_of, int LOOPD, int T, int R, int off, int64_t base,
                         std::vector<uint32_t> *map);

// one coefficient vector of a K×nvec block (θ[k·theta_stride], out[p·out_stride]) through the one-vector leaf kernel
cudaError_t smolyak_leaf_lincomb_strided(const sk_plan &plan, const double *theta, int64_t theta_stride, const double *x, int64_t n,
                                         double *out, int64_t out_stride, cudaStream_t stream);

// ---- Smolyak, two or four coefficient vectors per launch on unrolled leaves (sk_vec.cu): values, one-leaf bases
bool smolyak_vec_supported(const sk_plan &plan);
int64_t smolyak_vec_theta_map(const sk_plan &plan, int V, std::vector<uint32_t> *map);
cudaError_t smolyak_vec_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out,
                                cudaStream_t stream);

// ---- Smolyak, plain values with two points per thread on unrolled leaves (sk_pts.cu): one vector, one-leaf bases
bool smolyak_pts_supported(const sk_plan &plan);
int64_t smolyak_pts_theta_map(const sk_plan &plan, std::vector<uint32_t> *map);
cudaError_t smolyak_pts_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out, cudaStream_t stream);

// ---- Smolyak, second-order jets on unrolled leaves (sk_jet.cu): ∂ requests of total order <= 2 on bases that are one leaf
bool smolyak_jet_supported(sk_plan &plan);       // also fixes plan.jet_mode / plan.jet_col
int64_t smolyak_jet_theta_map(const sk_plan &plan, std::vector<uint32_t> *map);
cudaError_t smolyak_jet_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n, double *out, cudaStream_t stream);
double smolyak_jet_flops(const sk_plan &plan);

// ---- Smolyak, nested evaluation for any ∂ specification with per-coordinate orders <= 2 (sk_general.cu)
bool smolyak_general_supported(const sk_plan &plan);
bool smolyak_general_tree(const sk_plan &plan, std::vector<uint16_t> &tree, std::vector<int> &lvl_off);
cudaError_t smolyak_general_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n, double *out,
                                    cudaStream_t stream);
double smolyak_general_flops(const sk_plan &plan);

// ---- Smolyak with a K×nvec coefficient block: on-chip collocation tiles × Θ with FP64 DMMA (sk_block.cu)
bool smolyak_block_supported(const sk_plan &plan);
bool smolyak_block_program(const sk_plan &plan, std::vector<uint32_t> &prog, int *F, int *HL);
cudaError_t smolyak_block_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n,
                                  double *out, cudaStream_t stream);
double smolyak_block_flops(const sk_plan &plan, int nvec);

// ---- coordinate transformations over n×d points
cudaError_t transform_to(const sk_transform *tr, int d, int order, const double *y, int64_t n, double *out,
                         cudaStream_t stream);
cudaError_t transform_from(const sk_transform *tr, int d, const double *x, int64_t n, double *out,
                           cudaStream_t stream);

// ---- FP64 pipe microbenchmarks; returns flops executed by one launch
cudaError_t fp64_peak_launch(int which, int iters, double *sink, double *flops, cudaStream_t stream);

void count_launch(int n = 1);

}  // namespace skk
