// Instantiates the multi-vector leaf kernels (sk_vec_impl.cuh; two and four coefficient vectors per launch) for one nested
// family; split per translation unit so the build parallelises.
#include "sk_vec_impl.cuh"
namespace skk {
cudaError_t dispatch_vec_h1(int V, int LL, int R, const VecParams &q, cudaStream_t stream) {
    return V == 4 ? dispatch_vec<SK_GRID_ENDPOINT, 4>(LL, R, q, stream) : dispatch_vec<SK_GRID_ENDPOINT, 2>(LL, R, q, stream);
}
}  // namespace skk
