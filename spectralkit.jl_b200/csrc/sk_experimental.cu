// SURVEY.md §8(f) rank 3 — the caller of the hot path in the reference's `Experimental` module (src/experimental.jl):
//   make_policy_functions (:156-166): policy function m is  t_m ∘ linear_combination(basis, coefficients[m·K+1 : (m+1)·K]),
//       several functions over ONE basis from one stacked coefficient vector = the K×M coefficient block of the hot path;
//   bridge (:242-266): t = transform_from(PM1, outer, ·) ∘ transform_to(PM1, inner, ·) on the policy value;
//   sum_of_squared_residuals (:228-233): mapreduce(+) of sum_abs2(residuals) over the grid.
// The residual function itself is user code in the host language; what is batched here is everything around it:
//   sk_policy_functions       one pass: stacked coefficients → [K][M] block → linear_combination (leaf / DMMA kernels) →
//                             per-policy bridge in place (the reference's operations, unfused: bit-identical to Bridge)
//   sk_sum_abs2               deterministic sum of squares of a residual array (fixed partition, fixed tree)
//   sk_policy_residual_ssq    the fused least-squares special case r = policy(x) − target: evaluation, bridge, residual and
//                             reduction without materialising the residuals
#include <algorithm>
#include <string>
#include <vector>

#include "sk_device.cuh"
#include "sk_launch.h"

using skh::set_error;

namespace {

#define CKE(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return SK_ERR_CUDA; } \
    } while (0)

constexpr int kMaxPol = 64;
struct BridgeParams { int npol; sk_bridge b[kMaxPol]; };

// coefficients[m·K + k] → block[k·M + m]
__global__ void stack_to_block_kernel(const double *__restrict__ c, double *__restrict__ block, int64_t K, int M) {
    const int64_t total = K * M;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e / M;
        const int m = (int)(e - k * M);
        block[e] = c[(size_t)m * K + k];
    }
}

__device__ __forceinline__ double bridge_apply(const sk_bridge &b, double v) {
    // (b::Bridge)(x) = transform_from(PM1(), outer, transform_to(PM1(), inner, x))   (experimental.jl:256-260)
    return skd::from_pm1(b.outer, skd::to_pm1(b.inner, v));
}

__global__ void bridge_kernel(const __grid_constant__ BridgeParams q, double *__restrict__ v, int64_t n) {
    const int64_t total = n * q.npol;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        v[e] = bridge_apply(q.b[(int)(e % q.npol)], v[e]);
}

// partial[b] = Σ over this block's fixed slice of f(e); the slices and the in-block tree do not depend on timing
template <bool RESID>
__global__ void __launch_bounds__(256) ssq_partial_kernel(const __grid_constant__ BridgeParams q, const double *__restrict__ v, const double *__restrict__ target,
                                                          int64_t count, double *__restrict__ partial) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
        double r = v[e];
        if (RESID) r = bridge_apply(q.b[(int)(e % q.npol)], r) - target[e];
        s = fma(r, r, s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) ssq_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

constexpr int kSsqBlocks = 592;      // 4 per SM on a B200: a fixed partition, so the result is reproducible

struct Dev {
    void *p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) { return cudaMalloc(&p, b ? b : 1); }
};

int fill_bridges(const sk_bridge *bridges, int npol, BridgeParams &q) {
    q.npol = npol;
    for (int m = 0; m < npol; m++) {
        if (bridges) q.b[m] = bridges[m];
        else { q.b[m].outer.kind = SK_TR_NONE; q.b[m].outer.p0 = 0; q.b[m].outer.p1 = 1; q.b[m].inner = q.b[m].outer; q.b[m].outer.reserved = q.b[m].inner.reserved = 0; }
        for (const sk_transform *t : {&q.b[m].outer, &q.b[m].inner})
            if (t->kind < SK_TR_NONE || t->kind > SK_TR_INF_RATIONAL) SK_FAIL(SK_ERR_ARGUMENT, "invalid transformation kind in bridge");
    }
    return SK_OK;
}

// evaluation shared by the two entry points: device result [n][npol] (bridges NOT applied) in *dout
int evaluate_block(const sk_plan *plan, const double *coefficients, int64_t coefficients_len, int npol, const double *x, int64_t n,
                   uint32_t flags, cudaStream_t stream, Dev &dcoef, Dev &dblock, Dev &dx, double **dout_owned, Dev &dout) {
    sk_plan_info info{};
    int rc = sk_plan_get_info(plan, &info);
    if (rc != SK_OK) return rc;
    if (npol < 1 || npol > kMaxPol) SK_FAIL(SK_ERR_ARGUMENT, "need 1..64 policy functions");
    if (info.E != 1) SK_FAIL(SK_ERR_UNSUPPORTED, "policy functions take a plan without derivatives");
    if (coefficients_len != info.K * npol) SK_FAIL(SK_ERR_ARGUMENT, "lastindex(coefficients) == last(last(ranges)) must hold.");   // experimental.jl:162
    if (n < 0 || !coefficients || (n > 0 && !x)) SK_FAIL(SK_ERR_ARGUMENT, "bad buffers");
    const bool host = !(flags & SK_MEM_DEVICE);
    const double *dc = coefficients, *dxp = x;
    if (host) {
        CKE(dcoef.alloc((size_t)coefficients_len * 8));
        CKE(cudaMemcpyAsync(dcoef.p, coefficients, (size_t)coefficients_len * 8, cudaMemcpyHostToDevice, stream));
        dc = (const double *)dcoef.p;
        CKE(dx.alloc((size_t)n * info.d * 8));
        CKE(cudaMemcpyAsync(dx.p, x, (size_t)n * info.d * 8, cudaMemcpyHostToDevice, stream));
        dxp = (const double *)dx.p;
    }
    CKE(dblock.alloc((size_t)coefficients_len * 8));
    const int64_t tot = info.K * npol;
    stack_to_block_kernel<<<(int)std::min<int64_t>((tot + 255) / 256, 2048), 256, 0, stream>>>(dc, (double *)dblock.p, info.K, npol);
    skk::count_launch();
    CKE(cudaGetLastError());
    CKE(dout.alloc((size_t)std::max<int64_t>(n, 1) * npol * 8));
    *dout_owned = (double *)dout.p;
    if (n == 0) return SK_OK;
    return sk_linear_combination(plan, (const double *)dblock.p, info.K, npol, dxp, n, (double *)dout.p, (flags & ~SK_MEM_HOST) | SK_MEM_DEVICE, stream);
}

}  // namespace

extern "C" {

int sk_policy_functions(const sk_plan *plan, const double *coefficients, int64_t coefficients_len, int npol, const sk_bridge *bridges,
                        const double *x, int64_t n, double *out, uint32_t flags, void *stream_) {
    if (!plan || (n > 0 && !out)) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    sk_plan_info info{};
    int rc = sk_plan_get_info(plan, &info);
    if (rc != SK_OK) return rc;
    int prev = 0;
    CKE(cudaGetDevice(&prev));
    CKE(cudaSetDevice(info.device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    cudaStream_t stream = (cudaStream_t)stream_;
    BridgeParams q{};
    if (npol >= 1 && npol <= kMaxPol) { rc = fill_bridges(bridges, npol, q); if (rc != SK_OK) return rc; }
    Dev dcoef, dblock, dx, dout;
    double *res = nullptr;
    rc = evaluate_block(plan, coefficients, coefficients_len, npol, x, n, flags, stream, dcoef, dblock, dx, &res, dout);
    if (rc != SK_OK) return rc;
    if (n == 0) return SK_OK;
    if (bridges) {
        bridge_kernel<<<(int)std::min<int64_t>((n * npol + 255) / 256, 4096), 256, 0, stream>>>(q, res, n);
        skk::count_launch();
        CKE(cudaGetLastError());
    }
    CKE(cudaMemcpyAsync(out, res, (size_t)n * npol * 8, (flags & SK_MEM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    CKE(cudaStreamSynchronize(stream));
    return SK_OK;
}

int sk_sum_abs2(const double *r, int64_t count, double *result_host, uint32_t flags, int device, void *stream_) {
    if (!result_host || count < 0 || (count > 0 && !r)) SK_FAIL(SK_ERR_ARGUMENT, "bad argument");
    *result_host = 0.0;
    if (count == 0) return SK_OK;
    int ndev = 0;
    int rc = sk_device_count(&ndev);
    if (rc != SK_OK) return rc;
    if (device < 0 || device >= ndev) SK_FAIL(SK_ERR_ARGUMENT, "device index out of range");
    int prev = 0;
    CKE(cudaGetDevice(&prev));
    CKE(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    cudaStream_t stream = (cudaStream_t)stream_;
    Dev dr, dpart;
    const double *src = r;
    if (!(flags & SK_MEM_DEVICE)) {
        CKE(dr.alloc((size_t)count * 8));
        CKE(cudaMemcpyAsync(dr.p, r, (size_t)count * 8, cudaMemcpyHostToDevice, stream));
        src = (const double *)dr.p;
    }
    CKE(dpart.alloc((kSsqBlocks + 1) * 8));
    BridgeParams q{};
    q.npol = 1;
    ssq_partial_kernel<false><<<kSsqBlocks, 256, 0, stream>>>(q, src, nullptr, count, (double *)dpart.p);
    ssq_final_kernel<<<1, 256, 0, stream>>>((const double *)dpart.p, kSsqBlocks, (double *)dpart.p + kSsqBlocks);
    skk::count_launch(2);
    CKE(cudaGetLastError());
    CKE(cudaMemcpyAsync(result_host, (double *)dpart.p + kSsqBlocks, 8, cudaMemcpyDeviceToHost, stream));
    CKE(cudaStreamSynchronize(stream));
    return SK_OK;
}

int sk_policy_residual_ssq(const sk_plan *plan, const double *coefficients, int64_t coefficients_len, int npol, const sk_bridge *bridges,
                           const double *x, int64_t n, const double *target, double *result_host, uint32_t flags, void *stream_) {
    if (!plan || !result_host || (n > 0 && !target)) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    *result_host = 0.0;
    sk_plan_info info{};
    int rc = sk_plan_get_info(plan, &info);
    if (rc != SK_OK) return rc;
    int prev = 0;
    CKE(cudaGetDevice(&prev));
    CKE(cudaSetDevice(info.device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    cudaStream_t stream = (cudaStream_t)stream_;
    BridgeParams q{};
    if (npol >= 1 && npol <= kMaxPol) { rc = fill_bridges(bridges, npol, q); if (rc != SK_OK) return rc; }
    Dev dcoef, dblock, dx, dout, dt, dpart;
    double *res = nullptr;
    rc = evaluate_block(plan, coefficients, coefficients_len, npol, x, n, flags, stream, dcoef, dblock, dx, &res, dout);
    if (rc != SK_OK) return rc;
    if (n == 0) return SK_OK;
    const double *tg = target;
    if (!(flags & SK_MEM_DEVICE)) {
        CKE(dt.alloc((size_t)n * npol * 8));
        CKE(cudaMemcpyAsync(dt.p, target, (size_t)n * npol * 8, cudaMemcpyHostToDevice, stream));
        tg = (const double *)dt.p;
    }
    CKE(dpart.alloc((kSsqBlocks + 1) * 8));
    ssq_partial_kernel<true><<<kSsqBlocks, 256, 0, stream>>>(q, res, tg, n * npol, (double *)dpart.p);
    ssq_final_kernel<<<1, 256, 0, stream>>>((const double *)dpart.p, kSsqBlocks, (double *)dpart.p + kSsqBlocks);
    skk::count_launch(2);
    CKE(cudaGetLastError());
    CKE(cudaMemcpyAsync(result_host, (double *)dpart.p + kSsqBlocks, 8, cudaMemcpyDeviceToHost, stream));
    CKE(cudaStreamSynchronize(stream));
    return SK_OK;
}

}  // extern "C"
