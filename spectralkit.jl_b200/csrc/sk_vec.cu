// Host side of the multi-vector leaf kernels (sk_vec_impl.cuh): which plans they serve, the coefficient maps, and the
// dispatch to the per-family translation units (sk_vec_h*.cu).
#include <algorithm>

#include "sk_vec_impl.cuh"

namespace skk {

cudaError_t dispatch_vec_h0(int V, int LL, int R, const VecParams &q, cudaStream_t stream);
cudaError_t dispatch_vec_h1(int V, int LL, int R, const VecParams &q, cudaStream_t stream);
cudaError_t dispatch_vec_h2(int V, int LL, int R, const VecParams &q, cudaStream_t stream);

// plain values on a basis that is one leaf (d <= 6, block cap M = B, budgets the kernels are built for)
bool smolyak_vec_supported(const sk_plan &plan) {
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.ncomp != 0 || plan.basis.M != plan.basis.B) return false;
    const int d = plan.basis.d, B = plan.basis.B, gk = plan.basis.grid_kind;
    if (d > 6 || !vec_built_c(gk, d, B)) return false;
    for (int V = 2; V <= 4; V += 2) {
        const int64_t slots = smolyak_vec_theta_map(plan, V, nullptr);
        if (vec_smem_bytes(vec_nt_c(d, V), d, vec_treg_c(d), V, slots) + 16 > 200 * 1024) return false;
    }
    return true;
}

int64_t smolyak_vec_theta_map(const sk_plan &plan, int V, std::vector<uint32_t> *map) {
    const skh::SmolyakSet &S = plan.set;
    std::vector<int> blk_of((size_t)S.H + 2, S.M);
    for (int b = S.M - 1; b >= 0; b--) for (int64_t i = 1; i <= S.ell[(size_t)b]; i++) blk_of[(size_t)i] = b;
    if (map) { map->clear(); map->reserve((size_t)S.K); }
    return even_up(smolyak_leaf_map_rec(S, blk_of, vec_loopd_c(S.grid_kind, S.d, S.B, V), S.d, S.B, 0, 0, map));
}

// nvec coefficient vectors: groups of four, then two; a last single vector goes to the one-vector leaf kernel
cudaError_t smolyak_vec_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int d = plan.basis.d, R = plan.basis.B;
    VecParams q{};
    q.x = x; q.n = n; q.K = plan.K; q.theta_stride = nvec; q.out_stride = nvec;
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    int v0 = 0;
    while (nvec - v0 >= 2) {
        const int V = nvec - v0 >= 4 ? 4 : 2, w = V == 4 ? 1 : 0;
        q.theta = theta + v0; q.out = out + v0; q.thmap = plan.d_vthmap[w]; q.Kslots = plan.vslots[w];
        cudaError_t e = cudaErrorInvalidValue;
        switch (plan.basis.grid_kind) {
        case SK_GRID_INTERIOR: e = dispatch_vec_h0(V, d, R, q, stream); break;
        case SK_GRID_ENDPOINT: e = dispatch_vec_h1(V, d, R, q, stream); break;
        case SK_GRID_INTERIOR2: e = dispatch_vec_h2(V, d, R, q, stream); break;
        }
        if (e != cudaSuccess) return e;
        v0 += V;
    }
    if (v0 < nvec) return smolyak_leaf_lincomb_strided(plan, theta + v0, nvec, x, n, out + v0, nvec, stream);
    return cudaSuccess;
}

}  // namespace skk
