// Instantiates the unrolled-leaf Smolyak kernels (sk_leaf_impl.cuh) for one nested family and one
// output mode; split per translation unit so the build parallelises.
#include "sk_leaf_impl.cuh"
namespace skk {
cudaError_t dispatch_leaf_g1v(int LL, const LeafParams &q, cudaStream_t stream) {
    return dispatch_leaf<SK_GRID_ENDPOINT, false>(LL, q, stream);
}
}  // namespace skk
