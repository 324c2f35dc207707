// sm_100a kernels, part 5 — K2-general: Smolyak linear_combination for ANY ∂ specification (cross partials,
// second orders: ∂(1,1), ∂(2,2) ∪ …, src/derivatives.jl:336-421, test/test_smolyak.jl:39,
// test/test_derivatives.jl:107-112) with the nested, run-structured evaluation of sk_nested.cu.
//
// The reference forms, for each of the C requested components and each of the K terms, the product over
// dimensions of the univariate derivative of the order that component asks for (derivatives.jl:502-511).
// Evaluated dimension by dimension instead, level t only has to carry one partial sum per distinct
// order-PREFIX (o_1..o_t) that occurs among the components — the plan's "prefix tree" (value + gradient:
// t + 1 prefixes at level t; a full Hessian in d = 6: 3, 6, 10, 15, 21, 28):
//     acc_t[p] += T^{(o_t(p))}_{i_t}(x_t) · acc_{t−1}[parent(p)]
// One point per thread.  All per-level state (coordinate jets, the three-term recurrence of the jets
// (T, T', T'') of every level, the accumulators) lives in shared memory as [slot][thread] — conflict free,
// indexed at run time, so one kernel serves every specification with per-coordinate orders <= 2; the level-1
// run is summed in registers.  Control flow depends only on the index set: warp-uniform.
// Fast mode (fma, nested summation order): within the stated tolerance; exact mode uses the direct kernel.
#include <algorithm>

#include "sk_device.cuh"
#include "sk_launch.h"

namespace skk {

namespace {

constexpr int kMaxDeg = 2;
constexpr int kJ = kMaxDeg + 1;

struct GeneralParams {
    const double *x;
    double *out;
    const double *theta;
    const uint32_t *runs;            // run length | (levels that wrap afterwards) << 24
    const uint16_t *tree;            // per prefix: parent | order << 12, levels concatenated
    int64_t n, K, n_runs;
    int d, ncomp, total_pref, stage;
    int deg[SK_MAX_DIM];
    int lvl_off[SK_MAX_DIM + 1];     // first prefix of level t
    sk_transform tr[SK_MAX_DIM];
};

// one step of the recurrence on jets of raw derivatives, fma-contracted (chebyshev.jl:65-70 with derivatives.jl:125-134)
__device__ __forceinline__ void jet_step(const double (&xj)[kJ], const double (&c)[kJ], const double (&q)[kJ], double (&n)[kJ], int deg) {
    const double t0 = 2.0 * xj[0];
    n[0] = fma(t0, c[0], -q[0]);
    if (deg >= 1) {
        const double t1 = 2.0 * xj[1];
        n[1] = fma(t0, c[1], fma(t1, c[0], -q[1]));
        if (deg >= 2) n[2] = fma(t0, c[2], fma(2.0 * t1, c[1], fma(2.0 * xj[2], c[0], -q[2])));
    }
}

__global__ void __launch_bounds__(256) smolyak_general_kernel(const __grid_constant__ GeneralParams q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x, tid = threadIdx.x, d = q.d;
    // layout: tree | θ, runs (if staged) | per-thread slots
    uint16_t *s_tree = reinterpret_cast<uint16_t *>(smem_raw);
    size_t off = ((size_t)q.total_pref * 2 + 15) & ~(size_t)15;
    const double *theta = q.theta;
    const uint32_t *runs = q.runs;
    if (q.stage) {
        double *s_theta = reinterpret_cast<double *>(smem_raw + off);
        off += (size_t)q.K * 8;
        uint32_t *s_runs = reinterpret_cast<uint32_t *>(smem_raw + off);
        off += (((size_t)q.n_runs * 4) + 15) & ~(size_t)15;
        for (int64_t i = tid; i < q.K; i += NT) s_theta[i] = q.theta[i];
        for (int64_t i = tid; i < q.n_runs; i += NT) s_runs[i] = q.runs[i];
        theta = s_theta;
        runs = s_runs;
    }
    for (int i = tid; i < q.total_pref; i += NT) s_tree[i] = q.tree[i];
    double *slots = reinterpret_cast<double *>(smem_raw + off) + tid;
#define XJ(t, k) slots[(size_t)((t) * kJ + (k)) * NT]
#define TC(t, k) slots[(size_t)((d + (t)) * kJ + (k)) * NT]
#define TQ(t, k) slots[(size_t)((2 * d + (t)) * kJ + (k)) * NT]
#define ACC(i) slots[(size_t)(3 * d * kJ + (i)) * NT]
    __syncthreads();

    for (int64_t p0 = (int64_t)blockIdx.x * NT; p0 < q.n; p0 += (int64_t)gridDim.x * NT) {
        const bool live = p0 + tid < q.n;
        const int64_t p = live ? p0 + tid : q.n - 1;
        double x0j[kJ];                                      // level-1 coordinate jet stays in registers
        for (int t = 0; t < d; t++) {
            double xj[kJ];
            skd::to_pm1_jet<kJ>(q.tr[t], __ldcs(q.x + p * d + t), xj);
            if (t == 0) { x0j[0] = xj[0]; x0j[1] = xj[1]; x0j[2] = xj[2]; }
#pragma unroll
            for (int k = 0; k < kJ; k++) XJ(t, k) = xj[k];
        }
        uint32_t first = 0xFFFFFFFFu;                        // bit t: level t sits at index 1 (T = (1, 0, 0))
        const double *th = theta;
        const int deg0 = q.deg[0];
        for (int64_t r = 0; r < q.n_runs; r++) {
            const uint32_t w = runs[r];
            const int L = (int)(w & 0xFFFFFFu), c = (int)(w >> 24);
            // ---- level 1: a[k] = Σ_i θ_i · T_i^{(k)}(x_1) over the run, in registers
            double a[kJ] = {th[0], 0.0, 0.0};
            if (L > 1) {
                double tq[kJ] = {1.0, 0.0, 0.0}, tc[kJ] = {x0j[0], x0j[1], x0j[2]};
                double ti = th[1];
                a[0] = fma(ti, tc[0], a[0]);
                if (deg0 >= 1) a[1] = ti * tc[1];
                if (deg0 >= 2) a[2] = ti * tc[2];
                for (int i = 2; i < L; i++) {
                    double tn[kJ] = {0.0, 0.0, 0.0};
                    jet_step(x0j, tc, tq, tn, deg0);
                    ti = th[i];
                    a[0] = fma(ti, tn[0], a[0]);
                    if (deg0 >= 1) a[1] = fma(ti, tn[1], a[1]);
                    if (deg0 >= 2) a[2] = fma(ti, tn[2], a[2]);
#pragma unroll
                    for (int k = 0; k < kJ; k++) { tq[k] = tc[k]; tc[k] = tn[k]; }
                }
            }
            th += L;
            for (int pi = 0; pi < q.lvl_off[1]; pi++) {
                const int o = s_tree[pi] >> 12;
                ACC(pi) = o == 0 ? a[0] : o == 1 ? a[1] : a[2];
            }
            // ---- upper levels: fold the finished level t-1 into level t; levels 1..c wrap, level c+1 advances
            for (int t = 1; t < d; t++) {
                const int b0 = q.lvl_off[t], b1 = q.lvl_off[t + 1], cb = q.lvl_off[t - 1];
                if (first & (1u << t)) {
                    for (int pi = b0; pi < b1; pi++) {
                        const uint32_t e = s_tree[pi];
                        ACC(pi) = (e >> 12) == 0 ? ACC(cb + (e & 0xFFFu)) : 0.0;
                    }
                } else {
                    const double t0 = TC(t, 0), t1 = TC(t, 1), t2 = TC(t, 2);
                    for (int pi = b0; pi < b1; pi++) {       // (batching the loads ahead of the stores did not help: measured)
                        const uint32_t e = s_tree[pi];
                        const int o = (int)(e >> 12);
                        ACC(pi) = fma(o == 0 ? t0 : o == 1 ? t1 : t2, ACC(cb + (e & 0xFFFu)), ACC(pi));
                    }
                }
                if (c < t) {
                    // level t advances: step its jet recurrence
                    if (first & (1u << t)) {
#pragma unroll
                        for (int k = 0; k < kJ; k++) { TQ(t, k) = k == 0 ? 1.0 : 0.0; TC(t, k) = XJ(t, k); }
                        first &= ~(1u << t);
                    } else {
                        double xj[kJ], tc[kJ], tq[kJ], tn[kJ] = {0.0, 0.0, 0.0};
#pragma unroll
                        for (int k = 0; k < kJ; k++) { xj[k] = XJ(t, k); tc[k] = TC(t, k); tq[k] = TQ(t, k); }
                        jet_step(xj, tc, tq, tn, q.deg[t]);
#pragma unroll
                        for (int k = 0; k < kJ; k++) { TQ(t, k) = tc[k]; TC(t, k) = tn[k]; }
                    }
                    break;
                }
                first |= 1u << t;                            // level t wraps to index 1; its sum was folded upwards
            }
        }
        if (live) {
            const int top = q.lvl_off[d - 1];
            for (int cidx = 0; cidx < q.ncomp; cidx++) __stcs(q.out + p * q.ncomp + cidx, ACC(top + cidx));
        }
    }
#undef XJ
#undef TC
#undef TQ
#undef ACC
}

}  // namespace

// prefix tree of a canonical component table: level t holds the distinct (o_1..o_t) in order of first
// appearance; the top level is in component order (rows of a canonical expansion are distinct)
static bool build_tree(const sk_plan &plan, std::vector<uint16_t> &tree, int *lvl_off) {
    const int d = plan.basis.d, C = plan.ncomp;
    if (C < 1 || d > SK_MAX_DIM) return false;
    std::vector<std::vector<int>> pref_of((size_t)d, std::vector<int>((size_t)C, 0));   // prefix index of component c at level t
    tree.clear();
    lvl_off[0] = 0;
    for (int t = 0; t < d; t++) {
        std::vector<std::pair<int, int>> seen;              // (parent prefix, order) of the prefixes of this level
        for (int c = 0; c < C; c++) {
            const int o = plan.orders[c][t];
            if (o > kMaxDeg) return false;
            const int par = t == 0 ? 0 : pref_of[(size_t)t - 1][(size_t)c];
            int found = -1;
            for (size_t s = 0; s < seen.size(); s++) if (seen[s].first == par && seen[s].second == o) { found = (int)s; break; }
            if (found < 0) { found = (int)seen.size(); seen.push_back({par, o}); }
            pref_of[(size_t)t][(size_t)c] = found;
        }
        if (seen.size() > 4095) return false;
        for (auto &e : seen) tree.push_back((uint16_t)(e.first | (e.second << 12)));
        lvl_off[t + 1] = (int)tree.size();
    }
    // the top level must be the components themselves, in order
    if (lvl_off[d] - lvl_off[d - 1] != C) return false;
    for (int c = 0; c < C; c++) if (pref_of[(size_t)d - 1][(size_t)c] != c) return false;
    return true;
}

bool smolyak_general_supported(const sk_plan &plan) {
    if (plan.basis.kind != SK_BASIS_SMOLYAK || plan.ncomp < 1) return false;
    std::vector<uint16_t> tree;
    int lvl[SK_MAX_DIM + 1];
    if (!build_tree(plan, tree, lvl)) return false;
    const size_t per_thread = ((size_t)3 * plan.basis.d * kJ + tree.size()) * 8;
    return per_thread * 64 + tree.size() * 2 + 64 <= 200 * 1024;      // at least two warps per CTA
}

bool smolyak_general_tree(const sk_plan &plan, std::vector<uint16_t> &tree, std::vector<int> &lvl_off) {
    lvl_off.assign((size_t)plan.basis.d + 1, 0);
    return build_tree(plan, tree, lvl_off.data());
}

// closed-form FP64 flop count per point (fma = 2): walks the run program like the kernel
double smolyak_general_flops(const sk_plan &plan) {
    std::vector<uint16_t> tree;
    std::vector<int> lvl;
    if (!smolyak_general_tree(plan, tree, lvl)) return plan.flops_direct;
    const int d = plan.basis.d;
    auto rec = [&](int t) { return plan.deg[t] == 0 ? 2.0 : plan.deg[t] == 1 ? 6.0 : 13.0; };
    double f = 12.0 * d;
    std::vector<char> first((size_t)d, 1);
    for (int64_t r = 0; r < plan.n_runs; r++) {
        const int64_t L = plan.set.run_len[(size_t)r];
        const int c = plan.set.run_carry[(size_t)r];
        if (L > 1) f += 2.0 * (plan.deg[0] + 1) + (double)(L - 2) * (rec(0) + 2.0 * (plan.deg[0] + 1));
        for (int t = 1; t < d; t++) {
            if (!first[(size_t)t]) f += 2.0 * (lvl[(size_t)t + 1] - lvl[(size_t)t]);
            if (c < t) { if (!first[(size_t)t]) f += rec(t); first[(size_t)t] = 0; break; }
            first[(size_t)t] = 1;
        }
    }
    return f;
}

cudaError_t smolyak_general_lincomb(const sk_plan &plan, const double *theta, const double *x, int64_t n, double *out,
                                    cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    GeneralParams q{};
    q.x = x; q.out = out; q.theta = theta; q.runs = plan.d_runs; q.tree = plan.d_tree;
    q.n = n; q.K = plan.K; q.n_runs = plan.n_runs; q.d = plan.basis.d; q.ncomp = plan.ncomp;
    q.total_pref = plan.tree_lvl[(size_t)q.d];
    for (int t = 0; t < q.d; t++) {
        q.deg[t] = plan.deg[t];
        q.lvl_off[t] = plan.tree_lvl[(size_t)t];
        if (plan.has_tr) q.tr[t] = plan.tr[t];
        else { q.tr[t].kind = SK_TR_NONE; q.tr[t].reserved = 0; q.tr[t].p0 = 0; q.tr[t].p1 = 1; }
    }
    q.lvl_off[q.d] = q.total_pref;
    const size_t per_thread = ((size_t)3 * q.d * kJ + q.total_pref) * 8;
    const size_t tree_bytes = ((size_t)q.total_pref * 2 + 15) & ~(size_t)15;
    const size_t stage_bytes = (size_t)q.K * 8 + ((((size_t)q.n_runs * 4) + 15) & ~(size_t)15);
    const size_t limit = 200 * 1024;
    q.stage = tree_bytes + stage_bytes + per_thread * 64 <= limit;
    const size_t fixed = tree_bytes + (q.stage ? stage_bytes : 0);
    int NT = (int)std::min<size_t>(256, (limit - fixed) / per_thread / 32 * 32);
    if (NT < 32) return cudaErrorInvalidConfiguration;
    const size_t smem = fixed + per_thread * NT;
    cudaError_t e = cudaFuncSetAttribute(smolyak_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1, dev = 0, sms = 148;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smolyak_general_kernel, NT, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>((n + NT - 1) / NT, (int64_t)sms * std::max(per_sm, 1));
    smolyak_general_kernel<<<grid, NT, smem, stream>>>(q);
    count_launch();
    return cudaGetLastError();
}

}  // namespace skk
