// Host side of the unrolled-leaf Smolyak kernels (sk_leaf_impl.cuh): which plans they serve, the
// unit program, the closed-form flop count, and the dispatch to the per-family translation units.
#include <algorithm>
#include <cstdlib>

#include "sk_leaf_impl.cuh"

namespace skk {

#ifdef SK_GPTS_HOOK
cudaError_t smolyak_gpts_hook(const LeafParams &q, cudaStream_t stream);      // experiments (exp/gpts_variant.cu)
#endif
cudaError_t dispatch_leaf_g0v(int LL, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_g0g(int LL, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_g1v(int LL, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_g1g(int LL, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_g2v(int LL, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_g2g(int LL, const LeafParams &q, cudaStream_t stream);
// single-budget kernels (sk_leaf_s*.cu); cudaErrorNotSupported when (LL, R) is not built
cudaError_t dispatch_leaf_s0v(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_s0g(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_s1v(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_s1g(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_s2v(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_s2g(int LL, int R, const LeafParams &q, cudaStream_t stream);
// deep single-budget kernels (sk_leaf_t*.cu)
cudaError_t dispatch_leaf_t1v(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_t1g(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_t2v(int LL, int R, const LeafParams &q, cudaStream_t stream);
cudaError_t dispatch_leaf_t2g(int LL, int R, const LeafParams &q, cudaStream_t stream);

// number of unrolled leaf dimensions of a plan: chosen once by smolyak_leaf_supported (the deepest built leaf that
// covers the budget B and whose state fits in shared memory) and stored in the plan
static int leaf_depth(const sk_plan &plan) { return plan.leaf_LL; }

// shared-memory slot of every coefficient: mirrors leaf_end / LeafBlocks of sk_leaf_impl.cuh (pair-aligned blocks
// wherever a runtime loop changes the base); returns the slot after the last coefficient of Leaf<T, R>
int smolyak_leaf_map_rec(const skh::SmolyakSet &S, const std::vector<int> &blk_of, int LOOPD, int T, int R, int off, int64_t base,
                         std::vector<uint32_t> *map) {
    if (T == 0 || R == 0) { if (map) map->push_back((uint32_t)(base + off)); return off + 1; }
    off = smolyak_leaf_map_rec(S, blk_of, LOOPD, T - 1, R, off, base, map);
    if (T < LOOPD) {
        for (int64_t i = 2; i <= S.ell[(size_t)R]; i++) {
            const int b = blk_of[(size_t)i];
            if (R - b == 0) { if (map) map->push_back((uint32_t)(base + off)); off++; }
            else off = smolyak_leaf_map_rec(S, blk_of, LOOPD, T - 1, R - b, off, base, map);
        }
    } else {
        for (int B = 1; B <= R; B++) {
            const int cnt = (int)(S.ell[(size_t)B] - S.ell[(size_t)B - 1]);
            if (R - B == 0) {
                if (map) for (int c = 0; c < cnt; c++) map->push_back((uint32_t)(base + off + c));
                off = even_up(off + cnt);
            } else {
                const int start = even_up(off);
                const int CS = even_up(smolyak_leaf_map_rec(S, blk_of, LOOPD, T - 1, R - B, 0, 0, nullptr));
                if (map) for (int c = 0; c < cnt; c++) smolyak_leaf_map_rec(S, blk_of, LOOPD, T - 1, R - B, 0, base + start + (int64_t)c * CS, map);
                off = start + cnt * CS;
            }
        }
    }
    return off;
}

// the basis is one leaf of budget B and a kernel for that budget alone is built (sk_leaf_s*.cu)
static bool leaf_single_kernel(const sk_plan &plan) {
    const int LL = leaf_depth(plan);
    return leaf_single_enabled() && plan.basis.d == LL && leaf_single_built_c(plan.basis.grid_kind, LL, plan.basis.B, plan.fast_mode == 1);
}

// θ scatter map of the whole plan (units in traversal order, each starting an aligned block); returns the slot count
int64_t smolyak_leaf_theta_map(const sk_plan &plan, std::vector<uint32_t> *map) {
    const skh::SmolyakSet &S = plan.set;
    const int LL = leaf_depth(plan);
    const bool g = plan.fast_mode == 1;
    const int LOOPD = leaf_single_kernel(plan) ? leaf_single_loopd_c(plan.basis.grid_kind, LL, plan.basis.B, g) : leaf_loopd_c(LL, g);
    std::vector<int> blk_of((size_t)S.H + 2, S.M);
    for (int b = S.M - 1; b >= 0; b--) for (int64_t i = 1; i <= S.ell[(size_t)b]; i++) blk_of[(size_t)i] = b;
    std::vector<uint32_t> units;
    smolyak_leaf_units(plan, units);
    if (map) { map->clear(); map->reserve((size_t)S.K); }
    int64_t pos = 0;
    for (uint32_t w : units) pos += even_up(smolyak_leaf_map_rec(S, blk_of, LOOPD, LL, (int)(w & 0xFF), 0, pos, map));
    return pos;
}

static cudaError_t smolyak_leaf_lincomb_one(const sk_plan &plan, const double *theta, int64_t theta_stride, const double *x, int64_t n,
                                            double *out, int64_t out_stride, cudaStream_t stream);
cudaError_t smolyak_leaf_lincomb_strided(const sk_plan &plan, const double *theta, int64_t theta_stride, const double *x, int64_t n,
                                         double *out, int64_t out_stride, cudaStream_t stream) {
    return n == 0 ? cudaSuccess : smolyak_leaf_lincomb_one(plan, theta, theta_stride, x, n, out, out_stride, stream);
}

// Decides whether the unrolled-leaf kernel serves the plan and, if so, fixes its leaf depth (plan.leaf_LL) and CTA
// shape (plan.leaf_cta).  SK_LEAF_LL=n caps the depth (experiments).
bool smolyak_leaf_supported(sk_plan &plan) {
    plan.leaf_LL = 0; plan.leaf_cta = 0;
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.fast_mode < 0) return false;
    if (plan.basis.M != plan.basis.B) return false;            // block cap below the budget: generic kernel
    static const int cap = [] {
        const char *e = getenv("SK_LEAF_LL");
        int x = e ? atoi(e) : kLeafMaxLL;
        return x >= 1 && x <= kLeafMaxLL ? x : kLeafMaxLL;
    }();
    const int d = plan.basis.d, B = plan.basis.B;
    for (int LL = std::min(d, cap); LL >= 1; LL--) {
        const bool single = LL == d && leaf_single_enabled() && leaf_single_built_c(plan.basis.grid_kind, LL, B, plan.fast_mode == 1);
        if (!single && (!leaf_built_c(LL) || B > leaf_rmax_c(plan.basis.grid_kind, LL))) continue;
        const int nup = d - LL;
        if (nup > 12) break;
        int64_t n_units = 1;
        if (nup > 0) n_units = skh::smolyak_count(plan.basis.grid_kind, nup, plan.basis.B, plan.basis.M);
        if (n_units * 2 + plan.K > 64 * 1024) continue;        // cannot fit anyway (and keeps the map cheap)
        plan.leaf_LL = LL;
        plan.leaf_cta = smolyak_leaf_cta(plan);
        if (plan.leaf_cta) return true;                          // deepest leaf whose state fits in shared memory
    }
    plan.leaf_LL = 0;
    return false;
}

// which CTA shape serves the plan: 1 = the tuned full-size CTA, 2 = the 128-thread variant (bases with upper
// dimensions whose state does not fit otherwise), 3 = the same with four register tables, 0 = neither (θ or the state too large for shared memory)
int smolyak_leaf_cta(const sk_plan &plan) {
    const int LL = leaf_depth(plan);
    const int d = plan.basis.d, nup = d - LL;
    int64_t n_units = 1;
    if (nup > 0) n_units = skh::smolyak_count(plan.basis.grid_kind, nup, plan.basis.B, plan.basis.M);
    const bool g = plan.fast_mode == 1;
    const int64_t slots = smolyak_leaf_theta_map(plan, nullptr);
    const size_t limit = 200 * 1024;
    if (leaf_single_kernel(plan)) {
        const int gk = plan.basis.grid_kind, R = plan.basis.B;
        return LeafSmem(leaf_single_nt_c(gk, LL, R, g), LL, leaf_single_treg_c(gk, LL, R, g), d, 0, 1, slots, g).total <= limit ? 1 : 0;
    }
    if (LeafSmem(leaf_nt_c(LL, g), LL, leaf_treg_c(LL, g), d, nup, n_units, slots, g).total <= limit) return 1;
    if (nup > 0 && (LL == 2 || LL == 4 || LL == 6)) {
        const size_t sm_smem = 227 * 1024;
        const size_t a = LeafSmem(kLeafSmallNT, LL, leaf_treg_small_c(LL, g), d, nup, n_units, slots, g).total;
        const size_t b = LeafSmem(kLeafSmallNT, LL, leaf_treg_small_c(LL, g, true), d, nup, n_units, slots, g).total;
        // gradient with LL >= 4: a fourth register table costs speed per CTA (measured −14 % at d = 7, 8) but may
        // let one more CTA fit per SM (+62 % at d = 10, B = 3)
        // (at 255 registers per thread the register file holds two 128-thread CTAs, whatever shared memory allows)
        const auto ctas = [&](size_t bytes) { return std::min<size_t>(2, sm_smem / bytes); };
        const size_t cap = 226 * 1024;      // a CTA may use up to 227 KB of dynamic shared memory
        if (g && LL >= 4 && b <= cap && (a > cap || ctas(b) > ctas(a))) return 3;
        if (a <= cap) return 2;
    }
    return 0;
}

// units: distinct tails (i_{LL+1}..i_d) in traversal order; word = r | carry << 8
void smolyak_leaf_units(const sk_plan &plan, std::vector<uint32_t> &units) {
    const skh::SmolyakSet &S = plan.set;
    const int d = S.d, LL = leaf_depth(plan), nup = d - LL;
    units.clear();
    if (nup == 0) { units.push_back((uint32_t)S.B); return; }
    std::vector<int> blk_of((size_t)S.H + 2, S.M);
    for (int b = S.M - 1; b >= 0; b--) for (int64_t i = 1; i <= S.ell[(size_t)b]; i++) blk_of[(size_t)i] = b;
    std::vector<int> idx((size_t)nup, 1), blk((size_t)nup, 0);
    int used = 0;
    for (;;) {
        const uint32_t r = (uint32_t)(S.B - used);
        int j = 0;
        for (; j < nup; j++) {
            int cand = idx[(size_t)j] + 1;
            if (cand <= S.H) {
                int nb = blk_of[(size_t)cand];
                if (used - blk[(size_t)j] + nb <= S.B) { used += nb - blk[(size_t)j]; idx[(size_t)j] = cand; blk[(size_t)j] = nb; break; }
            }
            used -= blk[(size_t)j]; idx[(size_t)j] = 1; blk[(size_t)j] = 0;
        }
        units.push_back(r | ((uint32_t)j << 8));      // j levels wrapped (j == nup: end)
        if (j == nup) break;
    }
}

// closed-form FP64 flop count per point of the leaf kernel (fma = 2, mul/add = 1; transforms 10/coordinate),
// following the template recursion of sk_leaf_impl.cuh node by node
static double leaf_flops_rec(const skh::SmolyakSet &S, int T, int R, bool g, std::vector<std::vector<double>> &memo) {
    if (T == 0 || R == 0) return 0.0;
    double &m = memo[(size_t)T][(size_t)R];
    if (m >= 0) return m;
    double f = leaf_flops_rec(S, T - 1, R, g, memo);          // index 1: rename
    int b = 1;
    for (int64_t i = 2; i <= S.ell[(size_t)R]; i++) {
        while (S.ell[(size_t)b] < i) b++;
        if (i > kTab + 1) f += g ? 6 : 2;                     // recurrence step past the table
        if (R - b == 0) f += g ? 4 : 2;                       // single-coefficient child: value + own partial
        else f += leaf_flops_rec(S, T - 1, R - b, g, memo) + (g ? 2.0 * (T + 1) : 2.0);
    }
    return m = f;
}
double smolyak_leaf_flops(const sk_plan &plan) {
    const skh::SmolyakSet &S = plan.set;
    const bool g = plan.fast_mode == 1;
    const int d = S.d, LL = leaf_depth(plan), nup = d - LL;
    std::vector<std::vector<double>> memo((size_t)LL + 1, std::vector<double>((size_t)S.B + 1, -1.0));
    std::vector<uint32_t> units;
    smolyak_leaf_units(plan, units);
    double f = 12.0 * d + LL * (g ? 3 + (kTab - 2) * 6 : (kTab - 1) * 2);
    std::vector<char> first((size_t)std::max(nup, 1), 1);
    for (uint32_t w : units) {
        const int r = (int)(w & 0xFF), c = (int)(w >> 8);
        f += leaf_flops_rec(S, LL, r, g, memo);
        if (nup == 0) continue;
        if (!first[0]) f += g ? (r == 0 ? 4.0 : 2.0 * (LL + 2)) : 2.0;     // budget-0 leaf: value and the new partial only
        for (int u = 0; u < nup; u++) {
            if (u < c) {
                if (u + 1 < nup && !first[(size_t)u + 1]) f += g ? 2.0 * (LL + u + 3) : 2.0;
                first[(size_t)u] = 1;
            } else {
                if (!first[(size_t)u]) f += g ? 6 : 2;
                first[(size_t)u] = 0;
                break;
            }
        }
    }
    return f;
}

// nvec > 1 (values only): one launch per coefficient vector of the K×nvec block, θ read with stride nvec.  For a
// handful of vectors this beats the DMMA kernel (sk_block.cu), whose narrowest tile is 8 vectors wide.
cudaError_t smolyak_leaf_lincomb(const sk_plan &plan, const double *theta, int nvec, const double *x, int64_t n, double *out,
                                 cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (nvec > 1) {
        for (int v = 0; v < nvec; v++) {
            cudaError_t e = smolyak_leaf_lincomb_one(plan, theta + v, nvec, x, n, out + v, nvec, stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    return smolyak_leaf_lincomb_one(plan, theta, 1, x, n, out, plan.E, stream);
}

// ---- fused gather (sk_multi.cu, gather mode 2): the calling thread names up to 7 peer buffers; the next one-vector launch of
// a plan the PEER instantiation serves stores its results there too and reports it
namespace { thread_local int t_n_peer = 0; thread_local double *t_peer[7] = {}; thread_local bool t_peer_used = false; }
void smolyak_leaf_set_peers(int n, double *const *ptrs) {
    t_n_peer = n < 0 ? 0 : n > 7 ? 0 : n;
    for (int g = 0; g < t_n_peer; g++) t_peer[g] = ptrs[g];
    t_peer_used = false;
}
bool smolyak_leaf_peers_used() { return t_peer_used; }
// the multi-budget six-dimensional kernel with the full-size CTA (cfg 4: the headline) has a PEER instantiation
bool smolyak_leaf_serves_peers(const sk_plan &plan) {
    return plan.fast_kernel == 2 && plan.basis.d == 6 && plan.leaf_LL == 6 && plan.leaf_cta == 1 && !leaf_single_kernel(plan);
}

static cudaError_t smolyak_leaf_lincomb_one(const sk_plan &plan, const double *theta, int64_t theta_stride, const double *x, int64_t n,
                                            double *out, int64_t out_stride, cudaStream_t stream) {
    const int d = plan.basis.d, LL = leaf_depth(plan), nup = d - LL;
    LeafParams q{};
    if (t_n_peer > 0 && smolyak_leaf_serves_peers(plan)) {
        q.n_peer = t_n_peer;
        for (int g = 0; g < t_n_peer; g++) q.peer[g] = t_peer[g];
        t_peer_used = true;
    }
    q.x = x; q.out = out; q.theta = theta; q.units = plan.d_units; q.thmap = plan.d_thmap;
    q.theta_stride = theta_stride; q.out_stride = out_stride; q.small_cta = plan.leaf_cta >= 2 ? plan.leaf_cta : 0;
    q.n = n; q.K = plan.K; q.Kslots = plan.theta_slots; q.n_units = plan.n_units; q.D = d; q.nup = nup; q.budget = plan.basis.B;
    for (int j = 0; j < d; j++) {
        if (plan.has_tr) q.tr[j] = plan.tr[j];
        else { q.tr[j].kind = SK_TR_NONE; q.tr[j].reserved = 0; q.tr[j].p0 = 0; q.tr[j].p1 = 1; }
    }
    const bool g = plan.fast_mode == 1;
#ifdef SK_GPTS_HOOK
    if (g && d == 6 && plan.basis.B == 4 && plan.basis.grid_kind == SK_GRID_INTERIOR2 && nup == 0 && getenv("SK_GPTS")) return smolyak_gpts_hook(q, stream);
#endif
    if (leaf_single_kernel(plan)) {      // the basis is one leaf of budget B: single-budget kernel when built
        const int R = plan.basis.B;
        cudaError_t e = cudaErrorNotSupported;
        if (LL > 6) {
            switch (plan.basis.grid_kind) {
            case SK_GRID_ENDPOINT: e = g ? dispatch_leaf_t1g(LL, R, q, stream) : dispatch_leaf_t1v(LL, R, q, stream); break;
            case SK_GRID_INTERIOR2: e = g ? dispatch_leaf_t2g(LL, R, q, stream) : dispatch_leaf_t2v(LL, R, q, stream); break;
            }
        } else
        switch (plan.basis.grid_kind) {
        case SK_GRID_INTERIOR: e = g ? dispatch_leaf_s0g(LL, R, q, stream) : dispatch_leaf_s0v(LL, R, q, stream); break;
        case SK_GRID_ENDPOINT: e = g ? dispatch_leaf_s1g(LL, R, q, stream) : dispatch_leaf_s1v(LL, R, q, stream); break;
        case SK_GRID_INTERIOR2: e = g ? dispatch_leaf_s2g(LL, R, q, stream) : dispatch_leaf_s2v(LL, R, q, stream); break;
        }
        if (e != cudaErrorNotSupported) return e;
    }
    switch (plan.basis.grid_kind) {
    case SK_GRID_INTERIOR: return g ? dispatch_leaf_g0g(LL, q, stream) : dispatch_leaf_g0v(LL, q, stream);
    case SK_GRID_ENDPOINT: return g ? dispatch_leaf_g1g(LL, q, stream) : dispatch_leaf_g1v(LL, q, stream);
    case SK_GRID_INTERIOR2: return g ? dispatch_leaf_g2g(LL, q, stream) : dispatch_leaf_g2v(LL, q, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace skk
