// Instantiates the deep single-budget leaf kernels (sk_leaf_impl.cuh: leaf_single_deep_c) for one nested
// family and one output mode; split per translation unit so the build parallelises.
#include "sk_leaf_impl.cuh"
namespace skk {
cudaError_t dispatch_leaf_t2g(int LL, int R, const LeafParams &q, cudaStream_t stream) {
    return dispatch_leaf_single_deep<SK_GRID_INTERIOR2, true>(LL, R, q, stream);
}
}  // namespace skk
