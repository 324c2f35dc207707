// Instantiates the two-points-per-thread values kernels (sk_pts_impl.cuh) for one nested family.
#include "sk_pts_impl.cuh"
namespace skk {
cudaError_t dispatch_pts_h1(int LL, int R, const PtsParams &q, cudaStream_t stream) { return dispatch_pts<SK_GRID_ENDPOINT>(LL, R, q, stream); }
}  // namespace skk
