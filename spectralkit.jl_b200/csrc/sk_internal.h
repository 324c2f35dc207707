// Internal declarations shared by the host plan builder (sk_host.cpp), the kernels
// (sk_kernels.cu) and the C-ABI layer (sk_api.cu).  Not part of the public boundary.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "spectralkit_b200.h"

namespace skh {

// ---- errors (thread-local message, status code returned through the C ABI)
void set_error(const std::string &msg);
const char *last_error();
#define SK_FAIL(code, msg) do { ::skh::set_error(msg); return (code); } while (0)

// ---- block structure of the nested univariate families (smolyak_traversal.jl:69-84,112-116,146-148)
int64_t total_length(int grid_kind, int b);   // ℓ(b); -1 on bad kind
int64_t block_length(int grid_kind, int b);
// permutation of 1..len putting nested grids first (smolyak_traversal.jl:86-168); false if len is
// not a cumulative block length of that family
bool grid_shuffle(int grid_kind, int64_t len, std::vector<int64_t> &out);
// correctly rounded Chebyshev grid point i (1-based) of an N-point grid (chebyshev.jl:88-108)
double gridpoint(int grid_kind, int64_t N, int64_t i);

// ---- Smolyak index set in the reference's traversal order (column-major, i_1 fastest)
struct SmolyakSet {
    int grid_kind = 0, d = 0, B = 0, M = 0;
    int64_t H = 0;                 // ℓ(M)
    int64_t K = 0;                 // number of index tuples
    std::vector<int64_t> ell;      // ℓ(0..M)
    std::vector<uint16_t> idx;     // [K][d], 1-based
    // run structure: one run per distinct tail (i_2..i_d)
    //   run_len[r]   = ℓ(min(slack, M)) terms of θ, contiguous
    //   run_carry[r] = number of levels 2,3,… that wrap to index 1 when advancing to run r+1
    std::vector<uint32_t> run_len;
    std::vector<uint8_t> run_carry;
    std::vector<int64_t> Kj;       // K_j, j = 1..d (number of distinct suffixes (i_j..i_d)), Kj[0] = K
};
// count only (no enumeration): the number of tuples
int64_t smolyak_count(int grid_kind, int d, int B, int M);
bool smolyak_enumerate(int grid_kind, int d, int B, int M, SmolyakSet &out, bool want_idx = true);

// ---- ∂ specification algebra (derivatives.jl:228-316)
// partials: np rows of d ints (trailing zeros allowed).  Produces the canonical component table
// (row-major [ncomp][d]) and per-coordinate degrees.  Returns false on invalid input.
bool partials_canonical(int d, int np, const int32_t *partials, std::vector<int32_t> &orders,
                        std::vector<int32_t> &degrees);

}  // namespace skh

// ---- staging for calls with HOST buffers: device chunk buffers, streams and events of the copy/compute pipeline.
// A call borrows one from its plan (creating it on first use) and returns it, so repeated calls do not pay
// cudaMalloc/cudaFree (which also synchronise the device); concurrent calls on one plan get one each.
struct sk_workspace {
    static constexpr int NBUF = 3;
    void *theta = nullptr; size_t theta_cap = 0;
    void *xb[NBUF] = {}, *ob[NBUF] = {};
    size_t x_cap = 0, o_cap = 0;              // bytes per buffer
    void *s[3] = {};                           // cudaStream_t: H2D, kernels, D2H
    void *in[NBUF] = {}, *done[NBUF] = {}, *outc[NBUF] = {};   // cudaEvent_t
    void *e0 = nullptr;
};

// ---- the plan object behind the opaque handle
struct sk_plan {
    sk_basis_desc basis{};       // normalised (M clamped to B)
    int device = 0;
    bool has_tr = false;
    sk_transform tr[SK_MAX_DIM]{};
    int order = 0;               // univariate D
    int ncomp = 0;               // Smolyak: 0 = plain point
    int32_t orders[SK_MAX_COMP][SK_MAX_DIM]{};
    int32_t deg[SK_MAX_DIM]{};   // per-coordinate jet order
    int allow_d2 = 0;
    int E = 1;                   // doubles per output element (nvec == 1)
    int64_t K = 0, H = 0;
    int fast_mode = -1;          // Smolyak nested kernel: -1 unavailable, 0 values, 1 value+gradient
    skh::SmolyakSet set;
    // device tables (Smolyak)
    uint16_t *d_idx = nullptr;   // [K][d]
    uint16_t *d_fac = nullptr;   // [K][8] non-unit factors per term (values-only plans; direct kernels)
    uint32_t *d_runs = nullptr;  // [n_runs] run_len | carry << 24   (generic nested kernel)
    int64_t n_runs = 0;
    uint32_t *d_units = nullptr; // [n_units] r | carry << 8             (unrolled-leaf kernel)
    int64_t n_units = 0;
    uint32_t *d_thmap = nullptr; // [K] shared-memory slot of every coefficient (unrolled-leaf kernel)
    int64_t theta_slots = 0;
    int leaf_LL = 0;             // dimensions covered by the unrolled leaf (sk_leaf.cu: smolyak_leaf_supported)
    int leaf_cta = 0;            // 1 full-size CTA, 2 the 128-thread variant (sk_leaf.cu: smolyak_leaf_cta)
    uint32_t *d_bprog = nullptr; // factor program of the coefficient-block kernel (sk_block.cu)
    int bprog_F = 0, bprog_HL = 0;
    uint16_t *d_tree = nullptr;  // prefix tree of the component table (general nested kernel)
    std::vector<int> tree_lvl;   // first prefix of every level, then the total
    uint32_t *d_vthmap[2] = {nullptr, nullptr};   // coefficient maps of the two- and four-vector leaf kernels (sk_vec.cu)
    int64_t vslots[2] = {0, 0};
    uint32_t *d_pthmap = nullptr;  // coefficient map of the two-points-per-thread values kernel (sk_pts.cu)
    int64_t pslots = 0;
    bool pts_ok = false;
    bool vec_ok = false;         // plain values on a one-leaf basis: the multi-vector leaf kernels serve K×nvec blocks
    int jet_mode = 0;            // second-order-jet leaf kernel (sk_jet.cu): 2 value/gradient/diagonal, 3 full Hessian jet
    int8_t jet_col[32]{};        // jet component -> output column (-1: not requested)
    int fast_kernel = 0;         // 0 none (direct only), 1 generic nested, 2 unrolled leaves, 3 general nested (any ∂ spec),
                                 // 4 second-order-jet leaves (∂ requests of total order <= 2 on one-leaf bases)
    double flops_fast = 0, flops_direct = 0;
    // host-buffer staging pool (the only mutable state of a plan; guarded by ws_mu)
    mutable std::mutex ws_mu;
    mutable std::vector<sk_workspace *> ws_free;
    sk_plan() = default;
    sk_plan(const sk_plan &) = delete;
    sk_plan &operator=(const sk_plan &) = delete;
    ~sk_plan();                  // frees the device tables and the staging pool (csrc/sk_api.cu)
};
