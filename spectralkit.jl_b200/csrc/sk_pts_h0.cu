// Instantiates the two-points-per-thread values kernels (sk_pts_impl.cuh) for one nested family.
#include "sk_pts_impl.cuh"
namespace skk {
cudaError_t dispatch_pts_h0(int LL, int R, const PtsParams &q, cudaStream_t stream) { return dispatch_pts<SK_GRID_INTERIOR>(LL, R, q, stream); }
}  // namespace skk
