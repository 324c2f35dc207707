// C-ABI entry points (include/spectralkit_b200.h): argument checking with the reference's error
// behaviour, plan lifetime, host<->device staging, and the single-process multi-GPU driver.
// No evaluation happens on the host: every hot-path entry point fails if there is no device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "sk_launch.h"

using skh::set_error;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                \
            return e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? SK_ERR_NO_DEVICE : SK_ERR_CUDA; \
        }                                                                                 \
    } while (0)

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void workspace_destroy(sk_workspace *w) {
    if (!w) return;
    cudaFree(w->theta);
    for (int i = 0; i < sk_workspace::NBUF; i++) {
        cudaFree(w->xb[i]); cudaFree(w->ob[i]);
        if (w->in[i]) cudaEventDestroy((cudaEvent_t)w->in[i]);
        if (w->done[i]) cudaEventDestroy((cudaEvent_t)w->done[i]);
        if (w->outc[i]) cudaEventDestroy((cudaEvent_t)w->outc[i]);
    }
    if (w->e0) cudaEventDestroy((cudaEvent_t)w->e0);
    for (int i = 0; i < 3; i++) if (w->s[i]) cudaStreamDestroy((cudaStream_t)w->s[i]);
    delete w;
}

bool valid_grid_kind(int g) { return g == SK_GRID_INTERIOR || g == SK_GRID_ENDPOINT || g == SK_GRID_INTERIOR2; }

int normalise_basis(const sk_basis_desc *in, sk_basis_desc *out) {
    if (!in) SK_FAIL(SK_ERR_ARGUMENT, "basis descriptor is NULL");
    *out = *in;
    if (!valid_grid_kind(in->grid_kind)) SK_FAIL(SK_ERR_ARGUMENT, "invalid grid kind");
    if (in->kind == SK_BASIS_CHEBYSHEV) {
        if (in->N < 1) SK_FAIL(SK_ERR_ARGUMENT, "N ≥ 1 must hold.");                 // chebyshev.jl:29
        out->d = 1; out->B = 0; out->M = 0;
        return SK_OK;
    }
    if (in->kind != SK_BASIS_SMOLYAK) SK_FAIL(SK_ERR_ARGUMENT, "invalid basis kind");
    if (in->d < 1) SK_FAIL(SK_ERR_ARGUMENT, "N ≥ 1 must hold.");                      // smolyak_api.jl:65
    if (in->d > SK_MAX_DIM) SK_FAIL(SK_ERR_UNSUPPORTED, "Smolyak dimension above SK_MAX_DIM");
    if (in->B < 0) SK_FAIL(SK_ERR_ARGUMENT, "B isa Int && B ≥ 0 must hold.");          // smolyak_traversal.jl:13
    if (in->M < 0) SK_FAIL(SK_ERR_ARGUMENT, "M isa Int && M ≥ 0 must hold.");          // :14
    out->M = in->M > in->B ? in->B : in->M;                                            // :15-16 (the shim warns)
    if (out->M > SK_MAX_BLOCKS) SK_FAIL(SK_ERR_UNSUPPORTED, "block cap M above SK_MAX_BLOCKS");
    if (skh::total_length(out->grid_kind, out->M) > 65535) SK_FAIL(SK_ERR_UNSUPPORTED, "parent dimension above 65535");
    // tables are sized by K: refuse index sets beyond SK_MAX_TERMS up front (d = 16, B = M = 8, InteriorGrid would be 4.3e8)
    if (skh::smolyak_count(out->grid_kind, out->d, out->B, out->M) > (int64_t)SK_MAX_TERMS) SK_FAIL(SK_ERR_UNSUPPORTED, "basis dimension above SK_MAX_TERMS");
    out->N = 0;
    return SK_OK;
}

int check_transform(const sk_transform &t) {
    switch (t.kind) {
    case SK_TR_NONE: return SK_OK;
    case SK_TR_BOUNDED_LINEAR:
        if (!(std::isfinite(t.p0) && std::isfinite(t.p1) && t.p1 > 0)) SK_FAIL(SK_ERR_DOMAIN, "BoundedLinear: need finite m and s > 0");
        return SK_OK;
    case SK_TR_SEMI_INF_RATIONAL:
        if (!(std::isfinite(t.p0) && std::isfinite(t.p1) && t.p1 != 0)) SK_FAIL(SK_ERR_DOMAIN, "SemiInfRational: need finite A and finite L ≠ 0");
        return SK_OK;
    case SK_TR_INF_RATIONAL:
        if (!(std::isfinite(t.p0) && std::isfinite(t.p1) && t.p1 > 0)) SK_FAIL(SK_ERR_DOMAIN, "InfRational: need finite A and finite L > 0");
        return SK_OK;
    }
    SK_FAIL(SK_ERR_ARGUMENT, "invalid transformation kind");
}

bool is_rational(const sk_transform &t) { return t.kind == SK_TR_SEMI_INF_RATIONAL || t.kind == SK_TR_INF_RATIONAL; }

// ------------------------------------------------------------------------------ exception boundary
// Nothing may unwind through the C ABI into ctypes or a Julia ccall: the entry points that build tables sized by the
// basis run behind this guard and turn allocation failures into status codes.
template <class F>
int guarded(const char *what, F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        set_error(std::string(what) + ": out of host memory");
        return SK_ERR_UNSUPPORTED;
    } catch (const std::exception &e) {
        set_error(std::string(what) + ": " + e.what());
        return SK_ERR_ARGUMENT;
    } catch (...) {
        set_error(std::string(what) + ": unknown exception");
        return SK_ERR_ARGUMENT;
    }
}
}  // namespace

extern "C" {

const char *sk_version(void) { return "spectralkit-b200 0.1.0 (sm_100a)"; }
const char *sk_last_error(void) { return skh::last_error(); }

int sk_device_count(int *count) {
    if (!count) SK_FAIL(SK_ERR_ARGUMENT, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) { *count = 0; SK_FAIL(SK_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    *count = n;
    return SK_OK;
}

// ------------------------------------------------------------------------------ constructors
int sk_transform_make(int kind, double a, double b, sk_transform *out) {
    if (!out) SK_FAIL(SK_ERR_ARGUMENT, "out is NULL");
    out->kind = kind; out->reserved = 0;
    switch (kind) {
    case SK_TR_NONE: out->p0 = 0; out->p1 = 1; return SK_OK;
    case SK_TR_BOUNDED_LINEAR: {                                                       // transformations.jl:154-161
        if (!(std::isfinite(a) && std::isfinite(b))) SK_FAIL(SK_ERR_DOMAIN, "isfinite(a) && isfinite(b) must hold.");
        double s = (b - a) / 2, m = (a + b) / 2;
        if (!(s > 0)) SK_FAIL(SK_ERR_DOMAIN, "Need `a < b`.");
        out->p0 = m; out->p1 = s;
        return SK_OK;
    }
    case SK_TR_SEMI_INF_RATIONAL:                                                      // :211-215
        if (!std::isfinite(a)) SK_FAIL(SK_ERR_DOMAIN, "isfinite(A) must hold.");
        if (!(std::isfinite(b) && b != 0)) SK_FAIL(SK_ERR_DOMAIN, "isfinite(L) && L ≠ 0 must hold.");
        out->p0 = a; out->p1 = b;
        return SK_OK;
    case SK_TR_INF_RATIONAL:                                                           // :285-289
        if (!std::isfinite(a)) SK_FAIL(SK_ERR_DOMAIN, "isfinite(A) must hold.");
        if (!(std::isfinite(b) && b > 0)) SK_FAIL(SK_ERR_DOMAIN, "isfinite(L) && L > 0 must hold.");
        out->p0 = a; out->p1 = b;
        return SK_OK;
    }
    SK_FAIL(SK_ERR_ARGUMENT, "invalid transformation kind");
}

int sk_nesting_total_length(int grid_kind, int b, int64_t *out) {
    if (!out || !valid_grid_kind(grid_kind) || b < 0 || b > 30) SK_FAIL(SK_ERR_ARGUMENT, "bad argument");
    *out = skh::total_length(grid_kind, b);
    return SK_OK;
}
int sk_nesting_block_length(int grid_kind, int b, int64_t *out) {
    if (!out || !valid_grid_kind(grid_kind) || b < 0 || b > 30) SK_FAIL(SK_ERR_ARGUMENT, "bad argument");
    *out = skh::block_length(grid_kind, b);
    return SK_OK;
}
int sk_grid_shuffle(int grid_kind, int64_t len, int64_t *out) {
    if (!out || !valid_grid_kind(grid_kind)) SK_FAIL(SK_ERR_ARGUMENT, "bad argument");
    std::vector<int64_t> v;
    if (!skh::grid_shuffle(grid_kind, len, v)) SK_FAIL(SK_ERR_ARGUMENT, "len is not a cumulative block length of this grid family");
    std::memcpy(out, v.data(), v.size() * sizeof(int64_t));
    return SK_OK;
}

int sk_dimension(const sk_basis_desc *basis, int64_t *K) {
    sk_basis_desc b;
    int rc = normalise_basis(basis, &b);
    if (rc != SK_OK) return rc;
    if (!K) SK_FAIL(SK_ERR_ARGUMENT, "K is NULL");
    *K = b.kind == SK_BASIS_CHEBYSHEV ? b.N : skh::smolyak_count(b.grid_kind, b.d, b.B, b.M);
    return SK_OK;
}
int sk_parent_dimension(const sk_basis_desc *basis, int64_t *H) {
    sk_basis_desc b;
    int rc = normalise_basis(basis, &b);
    if (rc != SK_OK) return rc;
    if (!H) SK_FAIL(SK_ERR_ARGUMENT, "H is NULL");
    *H = b.kind == SK_BASIS_CHEBYSHEV ? b.N : skh::total_length(b.grid_kind, b.M);
    return SK_OK;
}

static int sk_smolyak_indices_impl(const sk_basis_desc *basis, int32_t *out) {
    sk_basis_desc b;
    int rc = normalise_basis(basis, &b);
    if (rc != SK_OK) return rc;
    if (b.kind != SK_BASIS_SMOLYAK || !out) SK_FAIL(SK_ERR_ARGUMENT, "need a Smolyak basis and an output buffer");
    skh::SmolyakSet S;
    if (!skh::smolyak_enumerate(b.grid_kind, b.d, b.B, b.M, S)) SK_FAIL(SK_ERR_ARGUMENT, "cannot enumerate index set");
    for (size_t i = 0; i < S.idx.size(); i++) out[i] = S.idx[i];
    return SK_OK;
}

static int sk_grid_impl(const sk_basis_desc *basis, const sk_transform *tr, double *out) {
    sk_basis_desc b;
    int rc = normalise_basis(basis, &b);
    if (rc != SK_OK) return rc;
    if (!out) SK_FAIL(SK_ERR_ARGUMENT, "out is NULL");
    const int d = b.kind == SK_BASIS_CHEBYSHEV ? 1 : b.d;
    if (tr) for (int j = 0; j < d; j++) { rc = check_transform(tr[j]); if (rc != SK_OK) return rc; }
    // transform_from on the host, reference order (transformations.jl:178-181, 244, 309)
    auto from = [](const sk_transform &t, double x) {
        switch (t.kind) {
        case SK_TR_BOUNDED_LINEAR: return x * t.p1 + t.p0;
        case SK_TR_SEMI_INF_RATIONAL: return t.p0 + (t.p1 * (1 + x)) / (1 - x);
        case SK_TR_INF_RATIONAL: return t.p0 + (t.p1 * x) / std::sqrt(1 - x * x);
        default: return x;
        }
    };
    if (b.kind == SK_BASIS_CHEBYSHEV) {
        for (int64_t i = 1; i <= b.N; i++) {
            double x = skh::gridpoint(b.grid_kind, b.N, i);
            out[i - 1] = tr ? from(tr[0], x) : x;
        }
        return SK_OK;
    }
    skh::SmolyakSet S;
    if (!skh::smolyak_enumerate(b.grid_kind, b.d, b.B, b.M, S)) SK_FAIL(SK_ERR_ARGUMENT, "cannot enumerate index set");
    std::vector<int64_t> sh;
    if (!skh::grid_shuffle(b.grid_kind, S.H, sh)) SK_FAIL(SK_ERR_ARGUMENT, "cannot build grid shuffle");
    std::vector<double> src((size_t)S.H);
    for (int64_t k = 0; k < S.H; k++) src[(size_t)k] = skh::gridpoint(b.grid_kind, S.H, sh[(size_t)k]);   // smolyak_api.jl:124-125
    for (int64_t k = 0; k < S.K; k++)
        for (int j = 0; j < d; j++) {
            double x = src[(size_t)S.idx[(size_t)k * d + j] - 1];
            out[(size_t)k * d + j] = tr ? from(tr[j], x) : x;
        }
    return SK_OK;
}

int sk_partials_canonical(int d, int np, const int32_t *partials, int32_t *out_orders, int cap, int32_t *ncomp,
                          int32_t *degrees) {
    if (!partials && np > 0) SK_FAIL(SK_ERR_ARGUMENT, "partials is NULL");
    std::vector<int32_t> ord, deg;
    if (!skh::partials_canonical(d, np, partials, ord, deg)) SK_FAIL(SK_ERR_ARGUMENT, "invalid partial derivative specification");
    int nc = (int)(ord.size() / (size_t)d);
    if (ncomp) *ncomp = nc;
    if (out_orders) {
        if (nc > cap) SK_FAIL(SK_ERR_ARGUMENT, "output capacity too small");
        std::memcpy(out_orders, ord.data(), ord.size() * sizeof(int32_t));
    }
    if (degrees) std::memcpy(degrees, deg.data(), (size_t)d * sizeof(int32_t));
    return SK_OK;
}

// ------------------------------------------------------------------------------ refinement (host, index merge)
namespace {
bool same_transforms(const sk_transform *a, const sk_transform *b, int d) {
    if (!a || !b) return a == b;                                    // a transformed and a plain basis never match
    for (int j = 0; j < d; j++)
        if (a[j].kind != b[j].kind || std::memcmp(&a[j].p0, &b[j].p0, sizeof(double)) != 0 || std::memcmp(&a[j].p1, &b[j].p1, sizeof(double)) != 0) return false;
    return true;
}
// is_subset_basis (chebyshev.jl:140-142, smolyak_api.jl:141-153, generic_api.jl:254,314-317)
bool subset_basis(const sk_basis_desc &a, const sk_transform *ta, const sk_basis_desc &b, const sk_transform *tb) {
    if (a.kind != b.kind) return false;
    const int d = a.kind == SK_BASIS_CHEBYSHEV ? 1 : a.d;
    if (a.kind == SK_BASIS_SMOLYAK && a.d != b.d) return false;
    if (!same_transforms(ta, tb, d)) return false;
    if (a.kind == SK_BASIS_CHEBYSHEV) return b.N >= a.N;             // grid kinds may differ
    return b.B >= a.B && b.M >= a.M && a.grid_kind == b.grid_kind;   // shared univariate indices need the same nested family
}
}  // namespace

static int sk_is_subset_basis_impl(const sk_basis_desc *basis1, const sk_transform *tr1, const sk_basis_desc *basis2, const sk_transform *tr2, int *out) {
    sk_basis_desc a, b;
    int rc = normalise_basis(basis1, &a);
    if (rc != SK_OK) return rc;
    rc = normalise_basis(basis2, &b);
    if (rc != SK_OK) return rc;
    if (!out) SK_FAIL(SK_ERR_ARGUMENT, "out is NULL");
    *out = subset_basis(a, tr1, b, tr2) ? 1 : 0;
    return SK_OK;
}

static int sk_augment_coefficients_impl(const sk_basis_desc *basis1, const sk_transform *tr1, const sk_basis_desc *basis2, const sk_transform *tr2,
                            const double *theta1, int64_t theta1_len, int nvec, double *theta2) {
    sk_basis_desc a, b;
    int rc = normalise_basis(basis1, &a);
    if (rc != SK_OK) return rc;
    rc = normalise_basis(basis2, &b);
    if (rc != SK_OK) return rc;
    if (!theta1 || !theta2 || nvec < 1) SK_FAIL(SK_ERR_ARGUMENT, "bad buffers");
    if (!subset_basis(a, tr1, b, tr2)) SK_FAIL(SK_ERR_ARGUMENT, "is_subset_basis(basis1, basis2) must hold.");          // smolyak_api.jl:203
    const size_t w = (size_t)nvec;
    if (a.kind == SK_BASIS_CHEBYSHEV) {                                                                       // chebyshev.jl:133-138
        if (theta1_len != a.N) SK_FAIL(SK_ERR_ARGUMENT, "length(θ1) == dimension(basis1) must hold.");
        std::memcpy(theta2, theta1, (size_t)a.N * w * sizeof(double));
        std::memset(theta2 + (size_t)a.N * w, 0, (size_t)(b.N - a.N) * w * sizeof(double));
        return SK_OK;
    }
    skh::SmolyakSet S1, S2;
    if (!skh::smolyak_enumerate(a.grid_kind, a.d, a.B, a.M, S1) || !skh::smolyak_enumerate(b.grid_kind, b.d, b.B, b.M, S2))
        SK_FAIL(SK_ERR_ARGUMENT, "cannot enumerate index set");
    if (theta1_len != S1.K) SK_FAIL(SK_ERR_ARGUMENT, "dimension(basis1) == length(θ1) must hold.");                // smolyak_api.jl:204
    // PaddingIterator (smolyak_api.jl:155-200): one pass over the larger traversal, the smaller one advancing
    // only on a match; after its last element nothing matches any more
    const size_t d = (size_t)a.d;
    int64_t i1 = 0;
    for (int64_t k = 0; k < S2.K; k++) {
        double *dst = theta2 + (size_t)k * w;
        if (i1 < S1.K && std::memcmp(&S1.idx[(size_t)i1 * d], &S2.idx[(size_t)k * d], d * sizeof(uint16_t)) == 0) {
            std::memcpy(dst, theta1 + (size_t)i1 * w, w * sizeof(double));
            i1++;
        } else std::memset(dst, 0, w * sizeof(double));
    }
    return SK_OK;
}

// ------------------------------------------------------------------------------ plans
static int sk_plan_create_impl(const sk_basis_desc *basis, const sk_transform *tr, const sk_deriv_spec *spec, int device,
                   sk_plan **out) {
    if (!out) SK_FAIL(SK_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    std::unique_ptr<sk_plan> p(new (std::nothrow) sk_plan());
    if (!p) SK_FAIL(SK_ERR_ARGUMENT, "out of memory");
    int rc = normalise_basis(basis, &p->basis);
    if (rc != SK_OK) return rc;
    int ndev = 0;
    rc = sk_device_count(&ndev);
    if (rc != SK_OK) return rc;
    if (device < 0 || device >= ndev) SK_FAIL(SK_ERR_ARGUMENT, "device index out of range");
    p->device = device;
    const bool uni = p->basis.kind == SK_BASIS_CHEBYSHEV;
    const int d = uni ? 1 : p->basis.d;
    p->has_tr = tr != nullptr;
    for (int j = 0; j < d; j++) {
        if (tr) { rc = check_transform(tr[j]); if (rc != SK_OK) return rc; p->tr[j] = tr[j]; }
        else { p->tr[j].kind = SK_TR_NONE; p->tr[j].reserved = 0; p->tr[j].p0 = 0; p->tr[j].p1 = 1; }
    }
    p->allow_d2 = spec ? spec->allow_rational_d2 : 0;
    if (uni) {
        p->order = spec ? spec->order : 0;
        if (p->order < 0) SK_FAIL(SK_ERR_ARGUMENT, "D ≥ 0 must hold.");                                  // derivatives.jl:69
        if (p->order > SK_MAX_ORDER) SK_FAIL(SK_ERR_UNSUPPORTED, "derivative order above SK_MAX_ORDER");
        p->deg[0] = p->order;
        p->K = p->H = p->basis.N;
        p->E = p->order + 1;
        p->flops_direct = p->flops_fast = (double)(p->K - 1) * (p->order == 0 ? 2 : p->order == 1 ? 6 : 12) + (double)p->K * 2 * (p->order + 1);
    } else {
        p->ncomp = spec ? spec->ncomp : 0;
        if (p->ncomp < 0 || p->ncomp > SK_MAX_COMP) SK_FAIL(SK_ERR_ARGUMENT, "invalid number of derivative components");
        if (p->ncomp > 0 && !spec->orders) SK_FAIL(SK_ERR_ARGUMENT, "orders is NULL");
        for (int c = 0; c < p->ncomp; c++)
            for (int j = 0; j < d; j++) {
                int o = spec->orders[c * d + j];
                if (o < 0 || o > SK_MAX_ORDER) SK_FAIL(SK_ERR_ARGUMENT, "derivative order out of range");
                p->orders[c][j] = o;
                if (o > p->deg[j]) p->deg[j] = o;
            }
        p->E = p->ncomp ? p->ncomp : 1;
        if (!skh::smolyak_enumerate(p->basis.grid_kind, d, p->basis.B, p->basis.M, p->set)) SK_FAIL(SK_ERR_ARGUMENT, "cannot enumerate index set");
        p->K = p->set.K; p->H = p->set.H;
        p->n_runs = (int64_t)p->set.run_len.size();
        // fast (nested) kernel: plain values, or the full gradient in canonical order [value, ∂1..∂d]
        if (p->ncomp == 0) p->fast_mode = 0;
        else if (p->ncomp == d + 1) {
            bool grad = true;
            for (int c = 0; c <= d && grad; c++)
                for (int j = 0; j < d; j++) if (p->orders[c][j] != (c >= 1 && j == c - 1 ? 1 : 0)) { grad = false; break; }
            p->fast_mode = grad ? 1 : -1;
        } else if (p->ncomp == 1) {
            bool zero = true;
            for (int j = 0; j < d; j++) zero &= p->orders[0][j] == 0;
            p->fast_mode = zero ? 0 : -1;
        }
        const double C = p->E;
        double ftab = 0;
        for (int j = 0; j < d; j++) ftab += (double)(p->H - 1) * (p->deg[j] == 0 ? 2 : p->deg[j] == 1 ? 6 : 12);
        p->flops_direct = (double)p->K * (C * (d - 1) + 2 * C) + ftab;
        p->fast_kernel = skk::smolyak_leaf_supported(*p) ? 2 : skk::smolyak_nested_supported(*p) ? 1 : 0;
        // any ∂ request no specialised kernel serves (general specifications; gradients with M < B or d > 10 that neither
        // the leaf nor the run-program kernel takes) goes to the general nested kernel when its state fits
        // ∂ requests of total order <= 2 (Hessian entries, diagonals, cross partials) on a basis that is one leaf: unrolled
        // second-order-jet leaves (SK_NO_JET=1 keeps them on the general kernel: experiments)
        if (p->ncomp > 0 && p->fast_mode < 0 && getenv("SK_NO_JET") == nullptr && skk::smolyak_jet_supported(*p)) p->fast_kernel = 4;
        if (!p->fast_kernel && p->ncomp > 0 && skk::smolyak_general_supported(*p)) p->fast_kernel = 3;
        // gradients the leaf kernel cannot serve (block cap M < B, θ too large): the generic nested kernel keeps its state
        // in registers, which spill beyond 8 dimensions (d = 10, B = 3: 3.6 vs 8.9 Mpoints/s); the general kernel keeps it
        // in shared memory
        if (p->fast_kernel == 1 && p->fast_mode == 1 && d >= 9 && skk::smolyak_general_supported(*p)) p->fast_kernel = 3;
        {   // experiments: SK_FORCE_GENERAL=1 sends every ∂ request the general kernel can serve to it
            static const int force = [] { const char *e = getenv("SK_FORCE_GENERAL"); return e ? atoi(e) : 0; }();
            if (force && p->ncomp > 0 && skk::smolyak_general_supported(*p)) { p->fast_kernel = 3; p->jet_mode = 0; }
        }
        p->n_runs = (int64_t)p->set.run_len.size();
        p->flops_fast = p->fast_kernel == 2 ? skk::smolyak_leaf_flops(*p) : p->fast_kernel == 1 ? skk::smolyak_nested_flops(*p)
                        : p->fast_kernel == 3 ? skk::smolyak_general_flops(*p) : p->fast_kernel == 4 ? skk::smolyak_jet_flops(*p) : p->flops_direct;
    }
    // the reference throws for order >= 2 through the rational maps (transformations.jl:266,332)
    for (int j = 0; j < d; j++)
        if (is_rational(p->tr[j]) && (p->deg[j] > 2 || (p->deg[j] == 2 && !p->allow_d2)))
            SK_FAIL(SK_ERR_UNSUPPORTED, std::to_string(p->deg[j]) + "th derivative not implemented yet, open an issue.");
    if (!uni) {
        DeviceGuard g(device);
        if (!g.ok) SK_FAIL(SK_ERR_CUDA, "cudaSetDevice failed");
        std::vector<uint32_t> prog((size_t)p->n_runs);
        for (size_t r = 0; r < prog.size(); r++) prog[r] = p->set.run_len[r] | ((uint32_t)p->set.run_carry[r] << 24);
        CK(cudaMalloc(&p->d_idx, p->set.idx.size() * sizeof(uint16_t)));
        CK(cudaMemcpy(p->d_idx, p->set.idx.data(), p->set.idx.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&p->d_runs, prog.size() * sizeof(uint32_t)));
        CK(cudaMemcpy(p->d_runs, prog.data(), prog.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (p->ncomp == 0 && p->H <= 4096 && d <= 16) {
            // compact factor lists for the direct kernels: dim << 12 | index-1 of the indices != 1, 0xFFFF-terminated
            std::vector<uint16_t> fac((size_t)p->K * 8, 0xFFFF);
            bool ok = true;
            for (int64_t k = 0; k < p->K && ok; k++) {
                int nf = 0;
                for (int j = 0; j < d; j++) {
                    const int i = p->set.idx[(size_t)k * d + j];
                    if (i == 1) continue;
                    if (nf == 8) { ok = false; break; }
                    fac[(size_t)k * 8 + nf++] = (uint16_t)((j << 12) | (i - 1));
                }
            }
            if (ok) {
                CK(cudaMalloc(&p->d_fac, fac.size() * sizeof(uint16_t)));
                CK(cudaMemcpy(p->d_fac, fac.data(), fac.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
            }
        }
        if (p->ncomp == 0) {
            std::vector<uint32_t> bprog;
            if (skk::smolyak_block_program(*p, bprog, &p->bprog_F, &p->bprog_HL)) {
                CK(cudaMalloc(&p->d_bprog, bprog.size() * sizeof(uint32_t)));
                CK(cudaMemcpy(p->d_bprog, bprog.data(), bprog.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            }
        }
        if (p->fast_kernel == 3) {
            std::vector<uint16_t> tree;
            if (!skk::smolyak_general_tree(*p, tree, p->tree_lvl)) SK_FAIL(SK_ERR_ARGUMENT, "internal error: prefix tree");
            CK(cudaMalloc(&p->d_tree, tree.size() * sizeof(uint16_t)));
            CK(cudaMemcpy(p->d_tree, tree.data(), tree.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        }
        if (p->ncomp == 0 && p->fast_kernel == 2 && skk::smolyak_vec_supported(*p)) {
            for (int w = 0; w < 2; w++) {
                std::vector<uint32_t> vmap;
                p->vslots[w] = skk::smolyak_vec_theta_map(*p, w ? 4 : 2, &vmap);
                if ((int64_t)vmap.size() != p->K) SK_FAIL(SK_ERR_ARGUMENT, "internal error: coefficient map does not cover the basis");
                CK(cudaMalloc(&p->d_vthmap[w], vmap.size() * sizeof(uint32_t)));
                CK(cudaMemcpy(p->d_vthmap[w], vmap.data(), vmap.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            }
            p->vec_ok = true;
        }
        if (p->ncomp == 0 && p->fast_kernel == 2 && skk::smolyak_pts_supported(*p)) {
            std::vector<uint32_t> pmap;
            p->pslots = skk::smolyak_pts_theta_map(*p, &pmap);
            if ((int64_t)pmap.size() != p->K) SK_FAIL(SK_ERR_ARGUMENT, "internal error: coefficient map does not cover the basis");
            CK(cudaMalloc(&p->d_pthmap, pmap.size() * sizeof(uint32_t)));
            CK(cudaMemcpy(p->d_pthmap, pmap.data(), pmap.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            p->pts_ok = true;
        }
        if (p->fast_kernel == 4) {
            std::vector<uint32_t> thmap;
            p->theta_slots = skk::smolyak_jet_theta_map(*p, &thmap);
            if ((int64_t)thmap.size() != p->K) SK_FAIL(SK_ERR_ARGUMENT, "internal error: coefficient map does not cover the basis");
            CK(cudaMalloc(&p->d_thmap, thmap.size() * sizeof(uint32_t)));
            CK(cudaMemcpy(p->d_thmap, thmap.data(), thmap.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        if (p->fast_kernel == 2) {
            std::vector<uint32_t> units;
            skk::smolyak_leaf_units(*p, units);
            p->n_units = (int64_t)units.size();
            CK(cudaMalloc(&p->d_units, units.size() * sizeof(uint32_t)));
            CK(cudaMemcpy(p->d_units, units.data(), units.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            std::vector<uint32_t> thmap;
            p->theta_slots = skk::smolyak_leaf_theta_map(*p, &thmap);
            if ((int64_t)thmap.size() != p->K) SK_FAIL(SK_ERR_ARGUMENT, "internal error: coefficient map does not cover the basis");
            CK(cudaMalloc(&p->d_thmap, thmap.size() * sizeof(uint32_t)));
            CK(cudaMemcpy(p->d_thmap, thmap.data(), thmap.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    *out = p.release();
    return SK_OK;
}

// ------------------------------------------------------------------------------ guarded entry points (see `guarded` above)
int sk_smolyak_indices(const sk_basis_desc *basis, int32_t *out) { return guarded("sk_smolyak_indices", [&] { return sk_smolyak_indices_impl(basis, out); }); }
int sk_grid(const sk_basis_desc *basis, const sk_transform *tr, double *out) { return guarded("sk_grid", [&] { return sk_grid_impl(basis, tr, out); }); }
int sk_is_subset_basis(const sk_basis_desc *basis1, const sk_transform *tr1, const sk_basis_desc *basis2, const sk_transform *tr2, int *out) {
    return guarded("sk_is_subset_basis", [&] { return sk_is_subset_basis_impl(basis1, tr1, basis2, tr2, out); });
}
int sk_augment_coefficients(const sk_basis_desc *basis1, const sk_transform *tr1, const sk_basis_desc *basis2, const sk_transform *tr2,
                            const double *theta1, int64_t theta1_len, int nvec, double *theta2) {
    return guarded("sk_augment_coefficients", [&] { return sk_augment_coefficients_impl(basis1, tr1, basis2, tr2, theta1, theta1_len, nvec, theta2); });
}
int sk_plan_create(const sk_basis_desc *basis, const sk_transform *tr, const sk_deriv_spec *spec, int device, sk_plan **out) {
    return guarded("sk_plan_create", [&] { return sk_plan_create_impl(basis, tr, spec, device, out); });
}

// frees everything the plan owns on its device; also runs when sk_plan_create fails half-way (unique_ptr)
sk_plan::~sk_plan() {
    DeviceGuard g(device);
    skk::smolyak_tile_basis_forget(this);
    if (d_idx) cudaFree(d_idx);
    if (d_runs) cudaFree(d_runs);
    if (d_units) cudaFree(d_units);
    if (d_thmap) cudaFree(d_thmap);
    for (int w = 0; w < 2; w++) if (d_vthmap[w]) cudaFree(d_vthmap[w]);
    if (d_pthmap) cudaFree(d_pthmap);
    if (d_fac) cudaFree(d_fac);
    if (d_tree) cudaFree(d_tree);
    if (d_bprog) cudaFree(d_bprog);
    for (sk_workspace *w : ws_free) workspace_destroy(w);
    ws_free.clear();
}

int sk_plan_destroy(sk_plan *plan) {
    delete plan;
    return SK_OK;
}

double sk_coefficient_block_flops(const sk_plan *plan, int nvec) {
    if (!plan || nvec < 1) return 0.0;
    if (plan->basis.kind == SK_BASIS_SMOLYAK && nvec > 1 && skk::smolyak_block_supported(*plan)) return skk::smolyak_block_flops(*plan, nvec);
    return plan->flops_fast * nvec;
}

int sk_plan_get_info(const sk_plan *plan, sk_plan_info *info) {
    if (!plan || !info) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    std::memset(info, 0, sizeof *info);
    info->kind = plan->basis.kind; info->grid_kind = plan->basis.grid_kind;
    info->d = plan->basis.kind == SK_BASIS_CHEBYSHEV ? 1 : plan->basis.d;
    info->B = plan->basis.B; info->M = plan->basis.M; info->device = plan->device;
    info->K = plan->K; info->H = plan->H; info->E = plan->E;
    info->fast_path = plan->basis.kind == SK_BASIS_CHEBYSHEV ? 1 : plan->fast_kernel;
    info->n_runs = plan->n_runs;
    info->flops_per_point_fast = plan->flops_fast;
    info->flops_per_point_direct = plan->flops_direct;
    info->bytes_per_point = 8.0 * info->d + 8.0 * plan->E;
    return SK_OK;
}

// ------------------------------------------------------------------------------ evaluation
namespace {

// crossover with the DMMA kernel measured on B200 (profiles/r1_block_kernel.md); SK_LEAF_PER_VECTOR_MAX overrides (experiments)
static int leaf_per_vector_max() {
    static const int v = [] { const char *e = getenv("SK_LEAF_PER_VECTOR_MAX"); return e ? atoi(e) : 24; }();
    return v;
}

// crossover of the multi-vector leaf kernels with the DMMA kernel (measured: profiles/r2_small_nvec.jsonl); SK_VEC_MAX
// overrides (0 disables them: experiments)
static int vec_max() {
    static const int v = [] { const char *e = getenv("SK_VEC_MAX"); return e ? atoi(e) : 40; }();
    return v;
}

// smallest batch for the two-points-per-thread values kernel (its tiles are twice as large); SK_PTS_MIN overrides,
// SK_PTS_MIN=-1 disables the kernel (experiments)
static int64_t pts_min_points() {
    static const int64_t v = [] { const char *e = getenv("SK_PTS_MIN"); const int64_t x = e ? atoll(e) : (int64_t)1 << 16; return x < 0 ? INT64_MAX : x; }();
    return v;
}

static int pts_vec_max() {
    static const int v = [] { const char *e = getenv("SK_PTS_VEC_MAX"); return e ? atoi(e) : 56; }();
    return v;
}

// one device-resident evaluation on `stream`
int run_lincomb_device(const sk_plan *plan, const double *d_theta, int nvec, const double *d_x, int64_t n,
                       double *d_out, bool exact, cudaStream_t stream) {
    if (plan->basis.kind == SK_BASIS_CHEBYSHEV) {
        CK(skk::uni_lincomb(plan->basis.N, plan->tr[0], plan->order, exact, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    // plain values, one vector, large batches on a one-leaf basis: two points per thread (sk_pts_impl.cuh)
    // (also per vector of a K×nvec block up to pts_vec_max() vectors: one such launch per vector beats the multi-vector leaf
    // kernel and, below ≈ 56 vectors, the DMMA kernel: profiles/r2_small_nvec.jsonl)
    if (!exact && nvec >= 1 && nvec <= pts_vec_max() && plan->pts_ok && n >= pts_min_points()) {
        CK(skk::smolyak_pts_lincomb(*plan, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    // a few coefficient vectors on a one-leaf basis: four (then two) vectors per launch share every Chebyshev value
    // (sk_vec_impl.cuh); the DMMA kernel takes over above vec_max() vectors
    if (!exact && nvec >= 2 && nvec <= vec_max() && plan->vec_ok) {
        CK(skk::smolyak_vec_lincomb(*plan, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    // a few coefficient vectors: one unrolled-leaf launch per vector beats the 8-wide DMMA tile (values only)
    if (!exact && plan->fast_kernel == 2 && (nvec == 1 || (nvec <= leaf_per_vector_max() && plan->fast_mode == 0))) {
        CK(skk::smolyak_leaf_lincomb(*plan, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    if (!exact && nvec == 1 && plan->fast_kernel == 4) {
        CK(skk::smolyak_jet_lincomb(*plan, d_theta, d_x, n, d_out, stream));
        return SK_OK;
    }
    if (!exact && nvec == 1 && plan->fast_kernel == 3) {
        CK(skk::smolyak_general_lincomb(*plan, d_theta, d_x, n, d_out, stream));
        return SK_OK;
    }
    if (!exact && nvec == 1 && plan->fast_kernel == 1) {
        CK(skk::smolyak_nested_lincomb(*plan, d_theta, d_x, n, d_out, stream));
        return SK_OK;
    }
    if (!exact && nvec > 1 && skk::smolyak_block_supported(*plan)) {
        CK(skk::smolyak_block_lincomb(*plan, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    if (skk::smolyak_warp_exact_serves(*plan, nvec)) {      // values / full gradient, one vector: one thread per point, shared prefix products
        CK(skk::smolyak_tile_lincomb_exact(*plan, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    if (!skk::smolyak_direct_fits(*plan, false)) {
        // full tables of 32 points exceed shared memory: the chunked-table kernel keeps the reference's order too
        if (!skk::smolyak_tile_lincomb_supported(*plan, nvec))
            SK_FAIL(SK_ERR_UNSUPPORTED, "bit-identical evaluation of this request needs more than 64 components or 512 coefficient vectors on a basis whose tables exceed shared memory (fast mode serves it)");
        CK(skk::smolyak_tile_lincomb_exact(*plan, d_theta, nvec, d_x, n, d_out, stream));
        return SK_OK;
    }
    CK(skk::smolyak_direct_lincomb(*plan, d_theta, nvec, d_x, n, d_out, stream));
    return SK_OK;
}

// points per chunk of the host-buffer pipeline (SK_HOST_CHUNK overrides; experiments)
int64_t host_chunk_points() {
    static const int64_t v = [] { const char *e = getenv("SK_HOST_CHUNK"); int64_t x = e ? atoll(e) : 0; return x > 0 ? x : (int64_t)1 << 21; }();
    return v;
}

// borrows a staging workspace from the plan for the duration of one call
struct WorkspaceLease {
    const sk_plan *plan;
    sk_workspace *ws = nullptr;
    explicit WorkspaceLease(const sk_plan *p) : plan(p) {
        std::lock_guard<std::mutex> lk(plan->ws_mu);
        if (!plan->ws_free.empty()) { ws = plan->ws_free.back(); plan->ws_free.pop_back(); }
    }
    bool poisoned = false;      // a failed call may leave copies or kernels in flight on the workspace's streams
    ~WorkspaceLease() {
        if (!ws) return;
        if (poisoned) {        // drain and destroy instead of handing half-used streams and buffers to the next call
            for (int i = 0; i < 3; i++) if (ws->s[i]) cudaStreamSynchronize((cudaStream_t)ws->s[i]);
            workspace_destroy(ws);
            return;
        }
        std::unique_lock<std::mutex> lk(plan->ws_mu);
        if (plan->ws_free.size() < 2) { plan->ws_free.push_back(ws); return; }
        lk.unlock();
        workspace_destroy(ws);
    }
    int ensure(size_t theta_bytes, size_t x_bytes, size_t o_bytes) {
        if (!ws) {
            ws = new (std::nothrow) sk_workspace();
            if (!ws) SK_FAIL(SK_ERR_ARGUMENT, "out of memory");
            for (int i = 0; i < 3; i++) CK(cudaStreamCreateWithFlags((cudaStream_t *)&ws->s[i], cudaStreamNonBlocking));
            for (int i = 0; i < sk_workspace::NBUF; i++) {
                CK(cudaEventCreateWithFlags((cudaEvent_t *)&ws->in[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags((cudaEvent_t *)&ws->done[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags((cudaEvent_t *)&ws->outc[i], cudaEventDisableTiming));
            }
            CK(cudaEventCreateWithFlags((cudaEvent_t *)&ws->e0, cudaEventDisableTiming));
        }
        if (theta_bytes > ws->theta_cap) { cudaFree(ws->theta); ws->theta = nullptr; ws->theta_cap = 0; CK(cudaMalloc(&ws->theta, theta_bytes)); ws->theta_cap = theta_bytes; }
        if (x_bytes > ws->x_cap) {
            for (int i = 0; i < sk_workspace::NBUF; i++) { cudaFree(ws->xb[i]); ws->xb[i] = nullptr; }
            ws->x_cap = 0;
            for (int i = 0; i < sk_workspace::NBUF; i++) CK(cudaMalloc(&ws->xb[i], x_bytes));
            ws->x_cap = x_bytes;
        }
        if (o_bytes > ws->o_cap) {
            for (int i = 0; i < sk_workspace::NBUF; i++) { cudaFree(ws->ob[i]); ws->ob[i] = nullptr; }
            ws->o_cap = 0;
            for (int i = 0; i < sk_workspace::NBUF; i++) CK(cudaMalloc(&ws->ob[i], o_bytes));
            ws->o_cap = o_bytes;
        }
        return SK_OK;
    }
};

int check_lincomb_args(const sk_plan *plan, const double *theta, int64_t theta_len, int nvec, const double *x,
                       int64_t n, double *out) {
    if (!plan) SK_FAIL(SK_ERR_ARGUMENT, "plan is NULL");
    if (n < 0) SK_FAIL(SK_ERR_ARGUMENT, "n < 0");
    if (nvec < 1) SK_FAIL(SK_ERR_ARGUMENT, "nvec < 1");
    if (theta_len < plan->K) SK_FAIL(SK_ERR_ARGUMENT, "not enough coefficients");          // generic_api.jl:100
    if (theta_len > plan->K) SK_FAIL(SK_ERR_ARGUMENT, "too many coefficients");            // generic_api.jl:101
    if (nvec > 1 && plan->E != 1) SK_FAIL(SK_ERR_UNSUPPORTED, "SVector coefficients need real basis values (no method _mul(::SVector, ::𝑑Expansion))");
    if (!theta || (n > 0 && (!x || !out))) SK_FAIL(SK_ERR_ARGUMENT, "NULL buffer");
    return SK_OK;
}

}  // namespace

int sk_linear_combination(const sk_plan *plan, const double *theta, int64_t theta_len, int nvec, const double *x,
                          int64_t n, double *out, uint32_t flags, void *stream_) {
    int rc = check_lincomb_args(plan, theta, theta_len, nvec, x, n, out);
    if (rc != SK_OK) return rc;
    DeviceGuard g(plan->device);
    if (!g.ok) SK_FAIL(SK_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool exact = (flags & SK_MODE_EXACT) != 0;
    const int d = plan->basis.kind == SK_BASIS_CHEBYSHEV ? 1 : plan->basis.d;
    const int Eout = nvec > 1 ? nvec : plan->E;
    if (flags & SK_MEM_DEVICE) return run_lincomb_device(plan, theta, nvec, x, n, out, exact, stream);

    // host buffers: θ once, then points in chunks through a 3-deep pipeline: the H2D copy of chunk i+1 and the
    // D2H copy of chunk i-1 overlap the kernel of chunk i when the caller's buffers are pinned (pageable
    // buffers still work, the copies just serialise).  PCIe is full duplex: the bound is max(in, out) bytes.
    if (n == 0) return SK_OK;
    const size_t theta_bytes = (size_t)plan->K * nvec * sizeof(double);
    const int64_t chunk = std::min<int64_t>(n, host_chunk_points());
    WorkspaceLease lease(plan);
    lease.poisoned = true;      // cleared when the call completes: every early return below drains and drops the workspace
    rc = lease.ensure(theta_bytes, (size_t)chunk * d * sizeof(double), (size_t)chunk * Eout * sizeof(double));
    if (rc != SK_OK) return rc;
    sk_workspace &w = *lease.ws;
    constexpr int NB = sk_workspace::NBUF;
    cudaStream_t s_in = (cudaStream_t)w.s[0], s_k = (cudaStream_t)w.s[1], s_out = (cudaStream_t)w.s[2];
    // order after work already queued on the caller's stream
    CK(cudaEventRecord((cudaEvent_t)w.e0, stream));
    CK(cudaStreamWaitEvent(s_in, (cudaEvent_t)w.e0, 0));
    CK(cudaStreamWaitEvent(s_k, (cudaEvent_t)w.e0, 0));
    CK(cudaMemcpyAsync(w.theta, theta, theta_bytes, cudaMemcpyHostToDevice, s_k));
    int64_t done_pts = 0;
    for (int it = 0; done_pts < n; it++) {
        const int b = it % NB;
        const int64_t m = std::min<int64_t>(chunk, n - done_pts);
        if (it >= NB) {                                                   // buffer reuse
            CK(cudaStreamWaitEvent(s_in, (cudaEvent_t)w.done[b], 0));     // the kernel that read xb[b] finished
            CK(cudaStreamWaitEvent(s_k, (cudaEvent_t)w.outc[b], 0));      // the D2H copy that read ob[b] finished
        }
        CK(cudaMemcpyAsync(w.xb[b], x + (size_t)done_pts * d, (size_t)m * d * sizeof(double), cudaMemcpyHostToDevice, s_in));
        CK(cudaEventRecord((cudaEvent_t)w.in[b], s_in));
        CK(cudaStreamWaitEvent(s_k, (cudaEvent_t)w.in[b], 0));
        rc = run_lincomb_device(plan, (const double *)w.theta, nvec, (const double *)w.xb[b], m, (double *)w.ob[b], exact, s_k);
        if (rc != SK_OK) { cudaDeviceSynchronize(); return rc; }
        CK(cudaEventRecord((cudaEvent_t)w.done[b], s_k));
        CK(cudaStreamWaitEvent(s_out, (cudaEvent_t)w.done[b], 0));
        CK(cudaMemcpyAsync(out + (size_t)done_pts * Eout, w.ob[b], (size_t)m * Eout * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        CK(cudaEventRecord((cudaEvent_t)w.outc[b], s_out));
        done_pts += m;
    }
    CK(cudaStreamSynchronize(s_out));
    CK(cudaStreamSynchronize(s_k));
    CK(cudaStreamSynchronize(s_in));
    lease.poisoned = false;
    return SK_OK;
}

int sk_basis_at(const sk_plan *plan, const double *x, int64_t n, double *out, int64_t ld, uint32_t flags, void *stream_) {
    if (!plan) SK_FAIL(SK_ERR_ARGUMENT, "plan is NULL");
    if (n < 0 || ld < n) SK_FAIL(SK_ERR_ARGUMENT, "need 0 <= n <= ld");
    if (n > 0 && (!x || !out)) SK_FAIL(SK_ERR_ARGUMENT, "NULL buffer");
    if (n == 0) return SK_OK;
    DeviceGuard g(plan->device);
    if (!g.ok) SK_FAIL(SK_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool uni = plan->basis.kind == SK_BASIS_CHEBYSHEV;
    const int d = uni ? 1 : plan->basis.d;
    auto run = [&](const double *dx, double *dout, int64_t nn, int64_t ldd) -> int {
        if (uni) CK(skk::uni_basis_at(plan->basis.N, plan->tr[0], plan->order, dx, nn, dout, ldd, stream));
        else {
            static const int old_kernel = [] { const char *e = getenv("SK_BASIS_OLD"); return e ? atoi(e) : 0; }();   // A/B against the round-1 kernel
            if (old_kernel && skk::smolyak_direct_fits(*plan, true)) CK(skk::smolyak_direct_basis_at(*plan, dx, nn, dout, ldd, stream));
            else CK(skk::smolyak_tile_basis_at(*plan, dx, nn, dout, ldd, stream));
        }
        return SK_OK;
    };
    if (flags & SK_MEM_DEVICE) return run(x, out, n, ld);
    // host buffers: points in chunks through the plan's staging workspace (the same 3-deep pipeline as
    // sk_linear_combination): chunk i+1 is copied in and chunk i-1 copied out (one strided 2-D copy into the caller's
    // column-major matrix) while chunk i is evaluated.  A 10^6 × 2 561 matrix never exists on the device in one piece.
    const size_t col_bytes = (size_t)plan->K * plan->E * sizeof(double);          // bytes of one point's K elements
    int64_t chunk = (int64_t)(((size_t)128 << 20) / col_bytes) & ~(int64_t)31;    // about 128 MB of results per chunk, whole 32-point tiles
    chunk = std::max<int64_t>(32, std::min<int64_t>(chunk, n));
    WorkspaceLease lease(plan);
    lease.poisoned = true;
    int rc = lease.ensure(0, (size_t)chunk * d * sizeof(double), (size_t)chunk * col_bytes);
    if (rc != SK_OK) return rc;
    sk_workspace &w = *lease.ws;
    constexpr int NB = sk_workspace::NBUF;
    cudaStream_t s_in = (cudaStream_t)w.s[0], s_k = (cudaStream_t)w.s[1], s_out = (cudaStream_t)w.s[2];
    CK(cudaEventRecord((cudaEvent_t)w.e0, stream));
    CK(cudaStreamWaitEvent(s_in, (cudaEvent_t)w.e0, 0));
    CK(cudaStreamWaitEvent(s_k, (cudaEvent_t)w.e0, 0));
    int64_t done_pts = 0;
    for (int it = 0; done_pts < n; it++) {
        const int bi = it % NB;
        const int64_t m = std::min<int64_t>(chunk, n - done_pts);
        if (it >= NB) {
            CK(cudaStreamWaitEvent(s_in, (cudaEvent_t)w.done[bi], 0));
            CK(cudaStreamWaitEvent(s_k, (cudaEvent_t)w.outc[bi], 0));
        }
        CK(cudaMemcpyAsync(w.xb[bi], x + (size_t)done_pts * d, (size_t)m * d * sizeof(double), cudaMemcpyHostToDevice, s_in));
        CK(cudaEventRecord((cudaEvent_t)w.in[bi], s_in));
        CK(cudaStreamWaitEvent(s_k, (cudaEvent_t)w.in[bi], 0));
        stream = s_k;
        rc = run((const double *)w.xb[bi], (double *)w.ob[bi], m, m);      // device chunk: leading dimension m
        if (rc != SK_OK) return rc;
        CK(cudaEventRecord((cudaEvent_t)w.done[bi], s_k));
        CK(cudaStreamWaitEvent(s_out, (cudaEvent_t)w.done[bi], 0));
        CK(cudaMemcpy2DAsync(out + (size_t)done_pts * plan->E, (size_t)ld * plan->E * sizeof(double), w.ob[bi], (size_t)m * plan->E * sizeof(double),
                             (size_t)m * plan->E * sizeof(double), (size_t)plan->K, cudaMemcpyDeviceToHost, s_out));
        CK(cudaEventRecord((cudaEvent_t)w.outc[bi], s_out));
        done_pts += m;
    }
    CK(cudaStreamSynchronize(s_out));
    CK(cudaStreamSynchronize(s_k));
    CK(cudaStreamSynchronize(s_in));
    lease.poisoned = false;
    return SK_OK;
}

static int transform_common(bool to, const sk_transform *tr, int d, int order, int allow_d2, const double *in, int64_t n,
                            double *out, uint32_t flags, int device, void *stream_) {
    if (!tr || d < 1 || d > SK_MAX_DIM) SK_FAIL(SK_ERR_ARGUMENT, "need 1..SK_MAX_DIM transformations");
    if (n < 0 || (n > 0 && (!in || !out))) SK_FAIL(SK_ERR_ARGUMENT, "bad buffers");
    if (order < 0) SK_FAIL(SK_ERR_ARGUMENT, "D ≥ 0 must hold.");
    if (order > SK_MAX_ORDER) SK_FAIL(SK_ERR_UNSUPPORTED, "derivative order above SK_MAX_ORDER");
    for (int j = 0; j < d; j++) {
        int rc = check_transform(tr[j]);
        if (rc != SK_OK) return rc;
        if (to && is_rational(tr[j]) && (order > 2 || (order == 2 && !allow_d2)))
            SK_FAIL(SK_ERR_UNSUPPORTED, std::to_string(order) + "th derivative not implemented yet, open an issue.");
    }
    if (n == 0) return SK_OK;
    int ndev = 0;
    int rc = sk_device_count(&ndev);
    if (rc != SK_OK) return rc;
    if (device < 0 || device >= ndev) SK_FAIL(SK_ERR_ARGUMENT, "device index out of range");
    DeviceGuard g(device);
    if (!g.ok) SK_FAIL(SK_ERR_NO_DEVICE, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t ib = (size_t)n * d * sizeof(double), ob = ib * (to ? order + 1 : 1);
    auto run = [&](const double *di, double *dout) -> int {
        if (to) CK(skk::transform_to(tr, d, order, di, n, dout, stream));
        else CK(skk::transform_from(tr, d, di, n, dout, stream));
        return SK_OK;
    };
    if (flags & SK_MEM_DEVICE) return run(in, out);
    double *di = nullptr, *dout = nullptr;
    CK(cudaMalloc(&di, ib));
    cudaError_t e = cudaMalloc(&dout, ob);
    if (e != cudaSuccess) { cudaFree(di); CK(e); }
    e = cudaMemcpyAsync(di, in, ib, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) rc = run(di, dout);
    if (e == cudaSuccess && rc == SK_OK) e = cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && rc == SK_OK) e = cudaStreamSynchronize(stream);
    cudaFree(di); cudaFree(dout);
    if (rc != SK_OK) return rc;
    CK(e);
    return SK_OK;
}

int sk_transform_to(const sk_transform *tr, int d, int order, int allow_rational_d2, const double *y, int64_t n,
                    double *out, uint32_t flags, int device, void *stream) {
    return transform_common(true, tr, d, order, allow_rational_d2, y, n, out, flags, device, stream);
}
int sk_transform_from(const sk_transform *tr, int d, const double *x, int64_t n, double *out, uint32_t flags, int device,
                      void *stream) {
    return transform_common(false, tr, d, 0, 0, x, n, out, flags, device, stream);
}

// ------------------------------------------------------------------------------ multi-GPU
int sk_linear_combination_sharded(sk_plan *const *plans, int ngpu, const double *theta, int64_t theta_len, int nvec,
                                  const double *x, int64_t n, double *out, uint32_t flags) {
    if (!plans || ngpu < 1) SK_FAIL(SK_ERR_ARGUMENT, "need at least one plan");
    if (flags & SK_MEM_DEVICE) SK_FAIL(SK_ERR_ARGUMENT, "sharded evaluation takes host buffers");
    for (int g = 0; g < ngpu; g++) {
        if (!plans[g]) SK_FAIL(SK_ERR_ARGUMENT, "plan is NULL");
        for (int h = 0; h < g; h++) if (plans[h]->device == plans[g]->device) SK_FAIL(SK_ERR_ARGUMENT, "plans must live on distinct devices");
        if (plans[g]->K != plans[0]->K || plans[g]->E != plans[0]->E) SK_FAIL(SK_ERR_ARGUMENT, "plans must describe the same basis");
    }
    int rc = check_lincomb_args(plans[0], theta, theta_len, nvec, x, n, out);
    if (rc != SK_OK) return rc;
    const int d = plans[0]->basis.kind == SK_BASIS_CHEBYSHEV ? 1 : plans[0]->basis.d;
    const int Eout = nvec > 1 ? nvec : plans[0]->E;
    const int64_t per = (n + ngpu - 1) / ngpu;           // contiguous slices of ceil(n/G) points
    std::vector<int> rcs((size_t)ngpu, SK_OK);
    std::vector<std::string> msgs((size_t)ngpu);
    std::vector<std::thread> workers;
    for (int g = 0; g < ngpu; g++) {
        workers.emplace_back([&, g]() {
            const int64_t lo = std::min<int64_t>(n, per * g), hi = std::min<int64_t>(n, lo + per);
            rcs[(size_t)g] = sk_linear_combination(plans[g], theta, theta_len, nvec, x + (size_t)lo * d, hi - lo,
                                                   out + (size_t)lo * Eout, flags, nullptr);
            if (rcs[(size_t)g] != SK_OK) msgs[(size_t)g] = sk_last_error();
        });
    }
    for (auto &w : workers) w.join();
    for (int g = 0; g < ngpu; g++) if (rcs[(size_t)g] != SK_OK) SK_FAIL(rcs[(size_t)g], "device " + std::to_string(plans[g]->device) + ": " + msgs[(size_t)g]);
    return SK_OK;
}

// ------------------------------------------------------------------------------ measurement
int sk_fp64_peak(int device, int which, double seconds, double *tflops_burst, double *tflops_sustained) {
    if (which < 0 || which > 3 || !tflops_burst || !tflops_sustained) SK_FAIL(SK_ERR_ARGUMENT, "bad argument");
    int ndev = 0;
    int rc = sk_device_count(&ndev);
    if (rc != SK_OK) return rc;
    if (device < 0 || device >= ndev) SK_FAIL(SK_ERR_ARGUMENT, "device index out of range");
    DeviceGuard g(device);
    double *sink = nullptr;
    CK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1, t0, t1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    const int iters = 4096;
    double flops = 0, best = 0;
    for (int w = 0; w < 3; w++) CK(skk::fp64_peak_launch(which, iters, sink, &flops, nullptr));
    CK(cudaDeviceSynchronize());
    // burst: the best single launch, timed alone (half of the budget)
    CK(cudaEventRecord(t0, nullptr));
    int launches = 0;
    float total_ms = 0, best_ms = 1e30f;
    do {
        CK(cudaEventRecord(e0, nullptr));
        CK(skk::fp64_peak_launch(which, iters, sink, &flops, nullptr));
        CK(cudaEventRecord(e1, nullptr));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, flops / (ms * 1e-3));
        best_ms = std::min(best_ms, ms);
        launches++;
        CK(cudaEventRecord(t1, nullptr));
        CK(cudaEventSynchronize(t1));
        CK(cudaEventElapsedTime(&total_ms, t0, t1));
    } while (total_ms < seconds * 0.5e3 && launches < 100000);
    // sustained: launches queued back to back between two events, no host synchronisation in between
    const int queued = (int)std::min(20000.0, std::max(8.0, seconds * 0.5e3 / best_ms));
    CK(cudaEventRecord(t0, nullptr));
    for (int i = 0; i < queued; i++) CK(skk::fp64_peak_launch(which, iters, sink, &flops, nullptr));
    CK(cudaEventRecord(t1, nullptr));
    CK(cudaEventSynchronize(t1));
    CK(cudaEventElapsedTime(&total_ms, t0, t1));
    *tflops_burst = best * 1e-12;
    *tflops_sustained = flops * queued / (total_ms * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    return SK_OK;
}

}  // extern "C"
