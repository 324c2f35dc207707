// SURVEY.md §8(f) rank 1 — the step before evaluation in every workflow: fitting θ on the basis' own grid,
//     θ = collocation_matrix(basis) \ y        (docs/src/index.md:92, test/test_smolyak.jl:20-21,
//                                               test/test_generic_api.jl:20)
// The collocation matrix is generated on the device by the library's own basis_at kernels (bit-identical to
// the reference's values) and never leaves it; the dense LU factorisation and the triangular solves — an
// initialisation-time step, not on the hot path — are cuSOLVER's (getrf/getrs, partial pivoting like Julia's
// `\` on a square matrix), loaded lazily with dlopen so that evaluation-only users never map the library.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "sk_launch.h"

using skh::set_error;

namespace {

typedef void *solver_handle;
struct Solver {
    void *lib = nullptr;
    int (*create)(solver_handle *) = nullptr;
    int (*destroy)(solver_handle) = nullptr;
    int (*set_stream)(solver_handle, cudaStream_t) = nullptr;
    int (*getrf_buffer)(solver_handle, int, int, double *, int, int *) = nullptr;
    int (*getrf)(solver_handle, int, int, double *, int, double *, int *, int *) = nullptr;
    int (*getrs)(solver_handle, int /*cublasOperation_t*/, int, int, const double *, int, const int *, double *, int, int *) = nullptr;
    std::string error;
};

Solver &solver() {
    static Solver s;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"}) {
            s.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (s.lib) break;
        }
        if (!s.lib) { s.error = std::string("cannot load libcusolver: ") + dlerror(); return; }
        auto sym = [&](const char *n) { void *p = dlsym(s.lib, n); if (!p && s.error.empty()) s.error = std::string("libcusolver lacks ") + n; return p; };
        s.create = reinterpret_cast<decltype(s.create)>(sym("cusolverDnCreate"));
        s.destroy = reinterpret_cast<decltype(s.destroy)>(sym("cusolverDnDestroy"));
        s.set_stream = reinterpret_cast<decltype(s.set_stream)>(sym("cusolverDnSetStream"));
        s.getrf_buffer = reinterpret_cast<decltype(s.getrf_buffer)>(sym("cusolverDnDgetrf_bufferSize"));
        s.getrf = reinterpret_cast<decltype(s.getrf)>(sym("cusolverDnDgetrf"));
        s.getrs = reinterpret_cast<decltype(s.getrs)>(sym("cusolverDnDgetrs"));
    });
    return s;
}

// [rows][cols] row-major  <->  column-major with leading dimension rows
__global__ void transpose_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t rows, int cols, int to_colmajor) {
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / cols;
        const int r = (int)(e % cols);
        if (to_colmajor) out[i + (int64_t)r * rows] = in[e];
        else out[e] = in[i + (int64_t)r * rows];
    }
}

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

extern "C" int sk_collocation_solve(const sk_plan *plan, const double *y, int nrhs, double *theta, uint32_t flags, void *stream_) {
    if (!plan || !y || !theta) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    if (nrhs < 1) SK_FAIL(SK_ERR_ARGUMENT, "nrhs < 1");
    if (plan->E != 1) SK_FAIL(SK_ERR_ARGUMENT, "collocation solve needs a plan without derivatives");
    const int64_t K = plan->K;
    if (K * K >= ((int64_t)1 << 31)) SK_FAIL(SK_ERR_UNSUPPORTED, "dimension(basis) above 46340: the dense factorisation is limited to 32-bit indexing");
    Solver &S = solver();
    if (!S.error.empty()) SK_FAIL(SK_ERR_UNSUPPORTED, S.error);
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
    DevBuf dx, dA, dB, dwork, dpiv, dinfo, dy;
    CKS(dx.alloc(grid.size() * sizeof(double)));
    CKS(dA.alloc((size_t)K * K * sizeof(double)));
    CKS(dB.alloc((size_t)K * nrhs * sizeof(double)));
    CKS(dpiv.alloc((size_t)K * sizeof(int)));
    CKS(dinfo.alloc(sizeof(int)));
    CKS(cudaMemcpyAsync(dx.p, grid.data(), grid.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    // collocation matrix, column-major K×K: A[i + k·K] = b_k(x_i)   (generic_api.jl:217-227)
    rc = sk_basis_at(plan, (const double *)dx.p, K, (double *)dA.p, K, SK_MEM_DEVICE, stream);
    if (rc != SK_OK) return rc;
    // right-hand sides: y[K][nrhs] → column-major K×nrhs
    const double *ysrc = y;
    if (!(flags & SK_MEM_DEVICE)) {
        CKS(dy.alloc((size_t)K * nrhs * sizeof(double)));
        CKS(cudaMemcpyAsync(dy.p, y, (size_t)K * nrhs * sizeof(double), cudaMemcpyHostToDevice, stream));
        ysrc = (const double *)dy.p;
    }
    const int tgrid = (int)std::min<int64_t>((K * nrhs + 255) / 256, 4096);
    transpose_kernel<<<tgrid, 256, 0, stream>>>(ysrc, (double *)dB.p, K, nrhs, 1);
    skk::count_launch();
    CKS(cudaGetLastError());

    solver_handle h = nullptr;
    if (S.create(&h) != 0) SK_FAIL(SK_ERR_CUDA, "cusolverDnCreate failed");
    struct HandleGuard { Solver &S; solver_handle h; ~HandleGuard() { S.destroy(h); } } hg{S, h};
    if (S.set_stream(h, stream) != 0) SK_FAIL(SK_ERR_CUDA, "cusolverDnSetStream failed");
    int lwork = 0;
    if (S.getrf_buffer(h, (int)K, (int)K, (double *)dA.p, (int)K, &lwork) != 0) SK_FAIL(SK_ERR_CUDA, "cusolverDnDgetrf_bufferSize failed");
    CKS(dwork.alloc((size_t)lwork * sizeof(double)));
    if (S.getrf(h, (int)K, (int)K, (double *)dA.p, (int)K, (double *)dwork.p, (int *)dpiv.p, (int *)dinfo.p) != 0) SK_FAIL(SK_ERR_CUDA, "cusolverDnDgetrf failed");
    int info = 0;
    CKS(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CKS(cudaStreamSynchronize(stream));
    if (info != 0) SK_FAIL(SK_ERR_ARGUMENT, "collocation matrix is singular (LU pivot " + std::to_string(info) + " is zero)");   // Julia: SingularException
    if (S.getrs(h, 0 /*CUBLAS_OP_N*/, (int)K, nrhs, (const double *)dA.p, (int)K, (const int *)dpiv.p, (double *)dB.p, (int)K, (int *)dinfo.p) != 0)
        SK_FAIL(SK_ERR_CUDA, "cusolverDnDgetrs failed");
    // θ[K][nrhs]
    if (flags & SK_MEM_DEVICE) {
        transpose_kernel<<<tgrid, 256, 0, stream>>>((const double *)dB.p, theta, K, nrhs, 0);
        skk::count_launch();
        CKS(cudaGetLastError());
        CKS(cudaStreamSynchronize(stream));
    } else {
        if (!dy.p) CKS(dy.alloc((size_t)K * nrhs * sizeof(double)));
        transpose_kernel<<<tgrid, 256, 0, stream>>>((const double *)dB.p, (double *)dy.p, K, nrhs, 0);
        skk::count_launch();
        CKS(cudaGetLastError());
        CKS(cudaMemcpyAsync(theta, dy.p, (size_t)K * nrhs * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CKS(cudaStreamSynchronize(stream));
    }
    return SK_OK;
}
