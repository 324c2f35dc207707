// Host side of the second-order-jet leaf kernels (sk_jet_impl.cuh): which plans they serve, the component -> column map,
// the closed-form flop count and the dispatch to the per-family translation units (sk_jet_h*.cu).
#include <algorithm>

#include "sk_jet_impl.cuh"

namespace skk {

cudaError_t dispatch_jet_h0d(int LL, int R, const JetParams &q, cudaStream_t stream);
cudaError_t dispatch_jet_h0f(int LL, int R, const JetParams &q, cudaStream_t stream);
cudaError_t dispatch_jet_h1d(int LL, int R, const JetParams &q, cudaStream_t stream);
cudaError_t dispatch_jet_h1f(int LL, int R, const JetParams &q, cudaStream_t stream);
cudaError_t dispatch_jet_h2d(int LL, int R, const JetParams &q, cudaStream_t stream);
cudaError_t dispatch_jet_h2f(int LL, int R, const JetParams &q, cudaStream_t stream);

// jet component of one requested partial (orders over the d coordinates): -1 when its total order exceeds 2
static int jet_component(int mode, int d, const int32_t *o) {
    int m = 0, n = 0, total = 0;
    for (int j = 0; j < d; j++) {
        if (o[j] < 0 || o[j] > 2) return -1;
        total += o[j];
        if (o[j] == 2) { m = n = j + 1; }
        else if (o[j] == 1) { if (!m) m = j + 1; else n = j + 1; }
    }
    if (total == 0) return 0;
    if (total == 1) return m;
    if (total != 2) return -1;
    if (m == n) return mode == 2 ? d + m : jet_sec3(d, m, m);
    return mode == 3 ? jet_sec3(d, m, n) : -1;
}

// smallest jet that holds every requested component: 2 (no cross partials) or 3; 0 = the request is not a second-order jet
static int jet_mode_of(const sk_plan &plan) {
    const int d = plan.basis.d;
    int mode = 2;
    for (int c = 0; c < plan.ncomp; c++) {
        if (jet_component(3, d, plan.orders[c]) < 0) return 0;
        if (jet_component(2, d, plan.orders[c]) < 0) mode = 3;
    }
    return mode;
}

static int jet_loopd(int gk, int LL, int R, int mode) {
    return gk == SK_GRID_INTERIOR ? jet_loopd_c(SK_GRID_INTERIOR, LL, R, mode) : gk == SK_GRID_ENDPOINT ? jet_loopd_c(SK_GRID_ENDPOINT, LL, R, mode)
                                                                                 : jet_loopd_c(SK_GRID_INTERIOR2, LL, R, mode);
}

// Serves the plan?  On success fixes plan.jet_mode and plan.jet_col.
bool smolyak_jet_supported(sk_plan &plan) {
    plan.jet_mode = 0;
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.ncomp < 1) return false;
    const int d = plan.basis.d, B = plan.basis.B, gk = plan.basis.grid_kind;
    if (plan.basis.M != B || d > 6 || !jet_built_c(gk, d, B)) return false;
    const int mode = jet_mode_of(plan);
    if (!mode) return false;
    for (int m = 0; m < 32; m++) plan.jet_col[m] = -1;
    for (int c = 0; c < plan.ncomp; c++) {
        const int m = jet_component(mode, d, plan.orders[c]);
        if (m < 0 || m >= 32 || plan.jet_col[m] >= 0) return false;      // rows of a canonical request are distinct
        plan.jet_col[m] = (int8_t)c;
    }
    plan.jet_mode = mode;
    const int64_t slots = smolyak_jet_theta_map(plan, nullptr);
    if (jet_smem_bytes(jet_nt_c(d, mode), d, jet_treg_c(d, mode), slots) > kJetSmemMax) { plan.jet_mode = 0; return false; }
    return true;
}

int64_t smolyak_jet_theta_map(const sk_plan &plan, std::vector<uint32_t> *map) {
    const skh::SmolyakSet &S = plan.set;
    std::vector<int> blk_of((size_t)S.H + 2, S.M);
    for (int b = S.M - 1; b >= 0; b--) for (int64_t i = 1; i <= S.ell[(size_t)b]; i++) blk_of[(size_t)i] = b;
    if (map) { map->clear(); map->reserve((size_t)S.K); }
    return even_up(smolyak_leaf_map_rec(S, blk_of, jet_loopd(S.grid_kind, S.d, S.B, plan.jet_mode), S.d, S.B, 0, 0, map));
}

// closed-form FP64 flop count per point (fma = 2), node by node like the template recursion
static double jet_flops_rec(const skh::SmolyakSet &S, int mode, int T, int R, std::vector<std::vector<double>> &memo) {
    if (T == 0 || R == 0) return 0.0;
    double &m = memo[(size_t)T][(size_t)R];
    if (m >= 0) return m;
    const auto nc = [&](int t) { return t == 0 ? 1 : mode == 2 ? 2 * t + 1 : (t + 1) * (t + 2) / 2; };
    double f = jet_flops_rec(S, mode, T - 1, R, memo);
    int b = 1;
    for (int64_t i = 2; i <= S.ell[(size_t)R]; i++) {
        while (S.ell[(size_t)b] < i) b++;
        if (i > kTab + 1) f += 13;
        if (R - b == 0) f += 6;
        else f += jet_flops_rec(S, mode, T - 1, R - b, memo) + 2.0 * (nc(T - 1) + 2 + (mode == 3 ? T - 1 : 0));
    }
    return m = f;
}
double smolyak_jet_flops(const sk_plan &plan) {
    const skh::SmolyakSet &S = plan.set;
    std::vector<std::vector<double>> memo((size_t)S.d + 1, std::vector<double>((size_t)S.B + 1, -1.0));
    return 12.0 * S.d + S.d * (kTab - 1) * 13.0 + jet_flops_rec(S, plan.jet_mode, S.d, S.B, memo);
}

cudaError_t smolyak_jet_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n, double *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int d = plan.basis.d;
    JetParams q{};
    q.x = x; q.out = out; q.theta = theta; q.thmap = plan.d_thmap;
    q.n = n; q.K = plan.K; q.Kslots = plan.theta_slots; q.ncomp = plan.ncomp;
    for (int m = 0; m < 32; m++) q.col[m] = plan.jet_col[m];
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    const int R = plan.basis.B;
    const bool full = plan.jet_mode == 3;
    switch (plan.basis.grid_kind) {
    case SK_GRID_INTERIOR: return full ? dispatch_jet_h0f(d, R, q, stream) : dispatch_jet_h0d(d, R, q, stream);
    case SK_GRID_ENDPOINT: return full ? dispatch_jet_h1f(d, R, q, stream) : dispatch_jet_h1d(d, R, q, stream);
    case SK_GRID_INTERIOR2: return full ? dispatch_jet_h2f(d, R, q, stream) : dispatch_jet_h2d(d, R, q, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace skk
