// Instantiates the second-order-jet leaf kernels (sk_jet_impl.cuh) for one nested family and one jet layout
// (d: value, gradient, diagonal second partials; f: full Hessian); split per translation unit so the build parallelises.
#include "sk_jet_impl.cuh"
namespace skk {
cudaError_t dispatch_jet_h2f(int LL, int R, const JetParams &q, cudaStream_t stream) {
    return dispatch_jet<SK_GRID_INTERIOR2, 3>(LL, R, q, stream);
}
}  // namespace skk
