// Host side of the two-points-per-thread values kernels (sk_pts_impl.cuh): which plans they serve, the coefficient map and
// the dispatch to the per-family translation units (sk_pts_h*.cu).
#include <algorithm>

#include "sk_pts_impl.cuh"

namespace skk {

cudaError_t dispatch_pts_h0(int LL, int R, const PtsParams &q, cudaStream_t stream);
cudaError_t dispatch_pts_h1(int LL, int R, const PtsParams &q, cudaStream_t stream);
cudaError_t dispatch_pts_h2(int LL, int R, const PtsParams &q, cudaStream_t stream);
cudaError_t smolyak_pts_lincomb_one(const sk_plan &plan, const double *theta, int64_t theta_stride, const double *x, int64_t n, double *out,
                                    int64_t out_stride, cudaStream_t stream);

bool smolyak_pts_supported(const sk_plan &plan) {
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.ncomp != 0 || plan.basis.M != plan.basis.B) return false;
    const int d = plan.basis.d, B = plan.basis.B, gk = plan.basis.grid_kind;
    if (d > 8 || !pts_built_c(gk, d, B)) return false;
    return pts_smem_bytes(pts_nt_c(d), d, pts_treg_c(d), smolyak_pts_theta_map(plan, nullptr)) <= 200 * 1024;
}

int64_t smolyak_pts_theta_map(const sk_plan &plan, std::vector<uint32_t> *map) {
    const skh::SmolyakSet &S = plan.set;
    std::vector<int> blk_of((size_t)S.H + 2, S.M);
    for (int b = S.M - 1; b >= 0; b--) for (int64_t i = 1; i <= S.ell[(size_t)b]; i++) blk_of[(size_t)i] = b;
    if (map) { map->clear(); map->reserve((size_t)S.K); }
    return even_up(smolyak_leaf_map_rec(S, blk_of, pts_loopd_c(S.grid_kind, S.d, S.B), S.d, S.B, 0, 0, map));
}

// nvec > 1: one launch per coefficient vector of the K×nvec block (θ and the results strided by nvec)
cudaError_t smolyak_pts_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (nvec > 1) {
        for (int v = 0; v < nvec; v++) {
            cudaError_t e = smolyak_pts_lincomb_one(plan, theta + v, nvec, x, n, out + v, nvec, stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    return smolyak_pts_lincomb_one(plan, theta, 1, x, n, out, 1, stream);
}
cudaError_t smolyak_pts_lincomb_one(const sk_plan &plan, const double *theta, int64_t theta_stride, const double *x, int64_t n, double *out,
                                    int64_t out_stride, cudaStream_t stream) {
    const int d = plan.basis.d, R = plan.basis.B;
    PtsParams q{};
    q.x = x; q.out = out; q.theta = theta; q.thmap = plan.d_pthmap; q.n = n; q.K = plan.K; q.Kslots = plan.pslots;
    q.theta_stride = theta_stride; q.out_stride = out_stride;
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    switch (plan.basis.grid_kind) {
    case SK_GRID_INTERIOR: return dispatch_pts_h0(d, R, q, stream);
    case SK_GRID_ENDPOINT: return dispatch_pts_h1(d, R, q, stream);
    case SK_GRID_INTERIOR2: return dispatch_pts_h2(d, R, q, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace skk
