/*
 * sk_oracle.c — CPU ORACLE for the SpectralKit.jl batched basis-evaluation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (spectralkit.jl_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's *pointwise* algorithm, in the reference's
 * own operation order (Julia never contracts a*b+c into an fma, so this file must be built
 * with -ffp-contract=off), citing the reference file:line each function follows
 * (paths relative to /root/reference/).  A batch is a loop over points (OpenMP), each point
 * re-running the whole pointwise path — including the Smolyak index state machine — exactly
 * as a Julia caller broadcasting `linear_combination` would.
 *
 * Parity status: Julia is not installed in the build image or on the GPU box, so the oracle
 * cannot be pinned against outputs of the reference itself.  It IS pinned against every
 * known-answer value the reference's docstrings/tests contain (tests/test_oracle_known_answers.py)
 * and against an independent pure-Python/mpmath restatement (oracle/py_oracle.py).
 * Arithmetic that lives in Julia Base (sinpi/cospi/hypot) is restated as: correctly rounded
 * sinpi/cospi of the already-rounded double argument (binary128 evaluation), and Julia's
 * `Base.Math._hypot` fma branch (documented as correctly rounded).
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXD1 6   /* jets up to order 5 (test/test_chebyshev.jl:93 uses 𝑑^5) */
#define ORC_MAXDIM 16 /* Smolyak dimension */
#define ORC_MAXM 8    /* block cap M */
#define ORC_MAXCOMP 64

enum { ORC_OK = 0, ORC_ERR_ARGUMENT = -1, ORC_ERR_DOMAIN = -2, ORC_ERR_UNSUPPORTED = -3 };
enum { TR_NONE = 0, TR_BOUNDED_LINEAR = 1, TR_SEMI_INF = 2, TR_INF = 3 };
enum { GRID_INTERIOR = 0, GRID_ENDPOINT = 1, GRID_INTERIOR2 = 2 };

typedef struct { int32_t kind; int32_t pad; double p0, p1; } orc_transform;

/* ------------------------------------------------------------------------------------------
 * Jet arithmetic — src/derivatives.jl:13-21,105-134.  A jet holds raw derivatives (f, f', …).
 * ---------------------------------------------------------------------------------------- */
typedef struct { double c[ORC_MAXD1]; } jet;

static const int BINOM[6][6] = {
    {1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0}, {1, 2, 1, 0, 0, 0},
    {1, 3, 3, 1, 0, 0}, {1, 4, 6, 4, 1, 0}, {1, 5, 10, 10, 5, 1}};

/* src/derivatives.jl:109-111 */
static jet jet_one(int dp1) { jet r; memset(&r, 0, sizeof r); r.c[0] = 1.0; (void)dp1; return r; }
/* src/derivatives.jl:105-107: (𝑑^D)(x) = (x, 1, 0, …) */
static jet jet_seed(double x, int dp1) {
    jet r; memset(&r, 0, sizeof r); r.c[0] = x; if (dp1 > 1) r.c[1] = 1.0; return r;
}
/* src/derivatives.jl:113-115 */
static jet jet_add(jet x, jet y, int dp1) { jet r = x; for (int k = 0; k < dp1; k++) r.c[k] = x.c[k] + y.c[k]; return r; }
/* src/derivatives.jl:117-119 */
static jet jet_sub(jet x, jet y, int dp1) { jet r = x; for (int k = 0; k < dp1; k++) r.c[k] = x.c[k] - y.c[k]; return r; }
/* src/derivatives.jl:121-123: real × jet */
static jet jet_scale(double a, jet y, int dp1) { jet r = y; for (int k = 0; k < dp1; k++) r.c[k] = a * y.c[k]; return r; }
/* src/derivatives.jl:125-134: Leibniz, term = (binom*x_i)*y_{k-i}, summed left to right */
static jet jet_mul(jet x, jet y, int dp1) {
    jet r; memset(&r, 0, sizeof r);
    for (int k = 0; k < dp1; k++) {
        double s = ((double)BINOM[k][0] * x.c[0]) * y.c[k];
        for (int i = 1; i <= k; i++) s = s + ((double)BINOM[k][i] * x.c[i]) * y.c[k - i];
        r.c[k] = s;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Julia Base.Math._hypot(x::Float64, y::Float64), the native-fma branch (documented there as
 * "correctly rounded").  Call site: src/transformations.jl:314.
 * ---------------------------------------------------------------------------------------- */
double orc_hypot(double x, double y) {
    double ax = fabs(x), ay = fabs(y);
    if (isinf(ax) || isinf(ay)) return INFINITY;
    if (isnan(ax) || isnan(ay)) return NAN;
    if (ay > ax) { double t = ax; ax = ay; ay = t; }
    const double eps = 2.220446049250313e-16;
    if (ay <= ax * sqrt(eps / 2)) return ax;
    double scale = eps * sqrt(2.2250738585072014e-308);
    if (ax > sqrt(1.7976931348623157e308 / 2)) { ax = ax * scale; ay = ay * scale; scale = 1.0 / scale; }
    else if (ay < sqrt(2.2250738585072014e-308)) { ax = ax / scale; ay = ay / scale; }
    else scale = 1.0;
    double h = sqrt(fma(ax, ax, ay * ay));
    double hsq = h * h, axsq = ax * ax;
    h -= (fma(-ay, ay, hsq - axsq) + fma(h, h, -hsq) - fma(ax, ax, -axsq)) / (2 * h);
    return h * scale;
}

/* ------------------------------------------------------------------------------------------
 * Transformations — src/transformations.jl.
 * Stored fields are the reference's: BoundedLinear (m, s) :150-153; rational maps (A, L).
 * ---------------------------------------------------------------------------------------- */
/* constructors with the reference's DomainError checks (:154-161, :211-215, :285-289) */
int orc_transform_make(int kind, double a, double b, orc_transform *out) {
    out->kind = kind; out->pad = 0;
    switch (kind) {
    case TR_NONE: out->p0 = 0; out->p1 = 1; return ORC_OK;
    case TR_BOUNDED_LINEAR: {
        if (!(isfinite(a) && isfinite(b))) return ORC_ERR_DOMAIN;
        double s = (b - a) / 2, m = (a + b) / 2;
        if (!(s > 0)) return ORC_ERR_DOMAIN;
        out->p0 = m; out->p1 = s; return ORC_OK; }
    case TR_SEMI_INF:
        if (!isfinite(a)) return ORC_ERR_DOMAIN;
        if (!(isfinite(b) && b != 0)) return ORC_ERR_DOMAIN;
        out->p0 = a; out->p1 = b; return ORC_OK;
    case TR_INF:
        if (!isfinite(a)) return ORC_ERR_DOMAIN;
        if (!(isfinite(b) && b > 0)) return ORC_ERR_DOMAIN;
        out->p0 = a; out->p1 = b; return ORC_OK;
    }
    return ORC_ERR_ARGUMENT;
}

/* transform_from: :178-181 (x*s+m), :244 (A + L*(1+x)/(1-x), `*` and `/` left-assoc), :309 */
double orc_transform_from(const orc_transform *t, double x) {
    switch (t->kind) {
    case TR_BOUNDED_LINEAR: return x * t->p1 + t->p0;
    case TR_SEMI_INF: return t->p0 + (t->p1 * (1 + x)) / (1 - x);
    case TR_INF: return t->p0 + (t->p1 * x) / sqrt(1 - x * x);
    default: return x;
    }
}

/* transform_to on a real: :183-186, :246-255, :311-320 */
static double transform_to_real(const orc_transform *t, double y) {
    switch (t->kind) {
    case TR_BOUNDED_LINEAR: return (y - t->p0) / t->p1;
    case TR_SEMI_INF: {
        double z = y - t->p0; double x = (z - t->p1) / (z + t->p1);
        if (y == INFINITY || y == -INFINITY) return 1.0; /* sic: +1 for both, :250-251 */
        return x; }
    case TR_INF: {
        double z = y - t->p0; double x = z / orc_hypot(z, t->p1);
        if (isinf(y)) return y > 0 ? 1.0 : -1.0;
        return x; }
    default: return y;
    }
}

/* transform_to on a jet: :188-195, :257-267, :322-333.
 * `ext2` != 0 enables the 2nd-derivative EXTENSION for the rational maps (the reference throws,
 * :266,332); formulas from SURVEY.md Appendix A.3, x2 = x''(y)*y1^2 + x'(y)*y2. */
int orc_transform_to_jet(const orc_transform *t, const double *y, int dp1, int ext2, double *x) {
    switch (t->kind) {
    case TR_NONE: for (int k = 0; k < dp1; k++) x[k] = y[k]; return ORC_OK;
    case TR_BOUNDED_LINEAR:
        x[0] = transform_to_real(t, y[0]);
        for (int k = 1; k < dp1; k++) x[k] = y[k] / t->p1;
        return ORC_OK;
    case TR_SEMI_INF: {
        double L = t->p1;
        double x0 = transform_to_real(t, y[0]); x[0] = x0;
        if (dp1 == 1) return ORC_OK;
        double Q = (x0 - 1) * (x0 - 1);           /* abs2(x0 - 1) */
        x[1] = (y[1] * Q) / (2 * L);
        if (dp1 == 2) return ORC_OK;
        if (!ext2 || dp1 > 3) return ORC_ERR_UNSUPPORTED;
        {   /* extension: x'(y) = Q/(2L); x''(y) = -(1-x0)^3/(2 L^2) */
            double omx = 1 - x0;
            double xpp = -((omx * omx) * omx) / ((2 * L) * L);
            x[2] = xpp * (y[1] * y[1]) + (Q / (2 * L)) * y[2];
        }
        return ORC_OK; }
    case TR_INF: {
        double L = t->p1;
        double x0 = transform_to_real(t, y[0]); x[0] = x0;
        if (dp1 == 1) return ORC_OK;
        double Q = 1 - x0 * x0;                    /* 1 - abs2(x0) */
        double sQ = sqrt(Q);
        x[1] = ((y[1] * Q) * sQ) / L;
        if (dp1 == 2) return ORC_OK;
        if (!ext2 || dp1 > 3) return ORC_ERR_UNSUPPORTED;
        {   /* extension: x'(y) = Q*sqrt(Q)/L; x''(y) = -3 x0 Q^2 / L^2 */
            double xpp = -(((3 * x0) * Q) * Q) / (L * L);
            x[2] = xpp * (y[1] * y[1]) + ((Q * sQ) / L) * y[2];
        }
        return ORC_OK; }
    }
    return ORC_ERR_ARGUMENT;
}

/* ------------------------------------------------------------------------------------------
 * Chebyshev — src/chebyshev.jl:56-70.  out[k*dp1 + r], k = 0..N-1 (basis function T_k).
 * ---------------------------------------------------------------------------------------- */
static void cheb_table(int N, int dp1, jet x, jet *out) {
    jet one = jet_one(dp1);
    out[0] = one;
    jet fp = one, fpp = x;
    for (int i = 2; i <= N; i++) {
        jet t = jet_scale(2.0, x, dp1);                  /* _mul(2, x) */
        jet f = jet_sub(jet_mul(t, fp, dp1), fpp, dp1);  /* _sub(_mul(_mul(2,x), fp), fpp) */
        out[i - 1] = f;
        fpp = fp; fp = f;
    }
}

/* real-valued fast path of the same iterator (T = Float64) */
static void cheb_table_real(int N, double x, double *out) {
    out[0] = 1.0;
    double fp = 1.0, fpp = x;
    for (int i = 2; i <= N; i++) {
        double f = (2 * x) * fp - fpp;
        out[i - 1] = f; fpp = fp; fp = f;
    }
}

int orc_cheb_basis_at(int N, int dp1, const double *xjet, double *out) {
    if (N < 1 || dp1 < 1 || dp1 > ORC_MAXD1) return ORC_ERR_ARGUMENT;
    jet x; memset(&x, 0, sizeof x);
    for (int k = 0; k < dp1; k++) x.c[k] = xjet[k];
    jet *tab = (jet *)malloc(sizeof(jet) * (size_t)N);
    cheb_table(N, dp1, x, tab);
    for (int i = 0; i < N; i++) for (int k = 0; k < dp1; k++) out[(size_t)i * dp1 + k] = tab[i].c[k];
    free(tab);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Grid points — src/chebyshev.jl:88-108.  sinpi/cospi are Julia Base; restated as the
 * correctly rounded value at the already-rounded Float64 argument (binary128 evaluation with
 * an exact reduction so that cospi(1/2) == 0 and cospi(1) == -1 exactly).
 * ---------------------------------------------------------------------------------------- */
static double sinpi_cr(double t) { /* |t| <= 1/2 in all call sites */
    if (t == 0.0) return t;
    return (double)sinq(M_PIq * (__float128)t);
}
static double cospi_cr(double t) { /* 0 <= t <= 1 in all call sites */
    __float128 u = 0.5Q - (__float128)t;  /* exact in binary128 */
    if (u == 0) return 0.0;
    return (double)sinq(M_PIq * u);
}

double orc_gridpoint(int grid_kind, int N, int i) {
    switch (grid_kind) {
    case GRID_INTERIOR: return sinpi_cr((double)(N - 2 * i + 1) / (double)(2 * N));
    case GRID_ENDPOINT:
        if (N == 1) return cospi_cr(1 / 2.0);
        return cospi_cr((double)(N - i) / (double)(N - 1));
    case GRID_INTERIOR2: return cospi_cr((double)(N - i + 1) / (double)(N + 1));
    }
    return NAN;
}

/* ------------------------------------------------------------------------------------------
 * Nesting lengths and grid shuffles — src/smolyak_traversal.jl:69-84,112-116,146-148 and the
 * three `iterate` pairs :86-106, :118-140, :150-168.
 * ---------------------------------------------------------------------------------------- */
static int64_t ipow3(int b) { int64_t r = 1; while (b-- > 0) r *= 3; return r; }

int64_t orc_nesting_total_length(int grid_kind, int b) {
    switch (grid_kind) {
    case GRID_ENDPOINT: return b == 0 ? 1 : (((int64_t)1 << b) + 1);
    case GRID_INTERIOR: return ipow3(b);
    case GRID_INTERIOR2: return ((int64_t)1 << (b + 1)) - 1;
    }
    return -1;
}
int64_t orc_nesting_block_length(int grid_kind, int b) {
    switch (grid_kind) {
    case GRID_ENDPOINT: return b <= 1 ? b + 1 : ((int64_t)1 << (b - 1));
    case GRID_INTERIOR: return b == 0 ? 1 : 2 * ipow3(b - 1);
    case GRID_INTERIOR2: return (int64_t)1 << b;
    }
    return -1;
}

/* returns the number of indices written (== len for valid cumulative block lengths) */
int64_t orc_grid_shuffle(int grid_kind, int64_t len, int64_t *out) {
    int64_t n = 0;
    if (grid_kind == GRID_ENDPOINT) {
        /* :86-106 — state (i, step); step == 0 special-cased through i == 0 */
        int64_t i = 0, step = 0;
        out[n++] = (len + 1) / 2;
        for (;;) {
            if (i == 0) {
                if (len > 1) { out[n++] = 1; i = 1; step = len - 1; continue; }
                break;
            }
            int64_t i2 = i + step;
            if (i2 <= len) { out[n++] = i2; i = i2; }
            else {
                int64_t step2 = step / 2;
                if (step2 >= 2) { i2 = step2 / 2 + 1; out[n++] = i2; i = i2; step = step2; }
                else break;
            }
        }
    } else if (grid_kind == GRID_INTERIOR) {
        /* :118-140 — state (i, i0, Δ, a) */
        int64_t i0 = (len + 1) / 2, D = len, a = 2, i = i0;
        out[n++] = i0;
        for (;;) {
            int64_t i2 = i + a * D;
            if (i2 <= len) { out[n++] = i2; i = i2; a = 3 - a; }
            else {
                if (D == 1) break;
                D = D / 3; i0 -= D; out[n++] = i0; i = i0; a = 2;
            }
        }
    } else {
        /* :150-168 — state (i, step) */
        int64_t i = (len + 1) / 2, step = 2 * i;
        out[n++] = i;
        for (;;) {
            int64_t i2 = i + step;
            if (i2 <= len) { out[n++] = i2; i = i2; }
            else {
                int64_t step2 = step / 2;
                if (step2 >= 2) { i2 = step2 / 2; out[n++] = i2; i = i2; step = step2; }
                else break;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Smolyak index traversal — src/smolyak_traversal.jl:174-231 (state machine),
 * :238-256 (length DP), :294-334 (iterator).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int N;                      /* dimension */
    int Mp1;
    int64_t ntl[ORC_MAXM + 2];  /* nesting_total_lengths, index b (0-based) */
    int64_t slack;
    int64_t indices[ORC_MAXDIM], blocks[ORC_MAXDIM], limits[ORC_MAXDIM];
} smolyak_state;

/* :174-181 */
static void inc_init(smolyak_state *s, int B) {
    for (int j = 0; j < s->N; j++) { s->indices[j] = 1; s->blocks[j] = 0; s->limits[j] = s->ntl[0]; }
    s->slack = B;
}

/* :211-231, recursion over the tail expressed with an offset `j`; returns valid and *delta */
static int inc_rec(smolyak_state *s, int j, int64_t slack, int64_t *delta) {
    int64_t i1 = s->indices[j], b1 = s->blocks[j], l1 = s->limits[j];
    if (i1 < l1) { s->indices[j] = i1 + 1; *delta = 0; return 1; }
    else if (b1 < (s->Mp1 - 1) && slack > 0) {
        s->indices[j] = i1 + 1; s->blocks[j] = b1 + 1; s->limits[j] = s->ntl[b1 + 1]; *delta = -1; return 1;
    } else {
        if (j == s->N - 1) { *delta = 0; return 0; }
        int64_t d1 = b1, dt = 0;
        int valid = inc_rec(s, j + 1, slack + d1, &dt);
        s->indices[j] = 1; s->blocks[j] = 0; s->limits[j] = s->ntl[0];
        *delta = d1 + dt;
        return valid;
    }
}
/* :328-334 */
static int inc(smolyak_state *s) {
    int64_t delta = 0;
    int valid = inc_rec(s, 0, s->slack, &delta);
    if (!valid) return 0;
    s->slack += delta;
    return 1;
}

static int smolyak_state_setup(smolyak_state *s, int grid_kind, int N, int B, int M) {
    if (N < 1 || N > ORC_MAXDIM || B < 0 || M < 0) return ORC_ERR_ARGUMENT;
    if (M > B) M = B;                                   /* :15-16 (warning + clamp) */
    if (M > ORC_MAXM) return ORC_ERR_ARGUMENT;
    s->N = N; s->Mp1 = M + 1;
    for (int b = 0; b <= M; b++) s->ntl[b] = orc_nesting_total_length(grid_kind, b);
    inc_init(s, B);
    return ORC_OK;
}

/* :238-256 */
int64_t orc_smolyak_length(int grid_kind, int N, int B, int M) {
    if (N < 1 || B < 0 || M < 0) return -1;
    if (M > B) M = B;
    int64_t *c = (int64_t *)calloc((size_t)B + 1, sizeof(int64_t));
    for (int b = 0; b <= M; b++) c[b] = orc_nesting_block_length(grid_kind, b);
    for (int n = 2; n <= N; n++)
        for (int b = B; b >= 0; b--) {
            int64_t s = 0;
            int amax = b < M ? b : M;
            for (int a = 0; a <= amax; a++) s += orc_nesting_block_length(grid_kind, a) * c[b - a];
            c[b] = s;
        }
    int64_t tot = 0;
    for (int b = 0; b <= B; b++) tot += c[b];
    free(c);
    return tot;
}

int64_t orc_smolyak_parent_dimension(int grid_kind, int B, int M) {
    if (M > B) M = B;
    return orc_nesting_total_length(grid_kind, M);   /* H = last(nesting_total_lengths), :307 */
}

/* collect(SmolyakIndices): out[k*N + j], 1-based indices, traversal order; returns count */
int64_t orc_smolyak_indices(int grid_kind, int N, int B, int M, int64_t *out, int64_t cap) {
    smolyak_state s;
    if (smolyak_state_setup(&s, grid_kind, N, B, M) != ORC_OK) return -1;
    int64_t k = 0;
    do {
        if (k < cap) for (int j = 0; j < N; j++) out[k * N + j] = s.indices[j];
        k++;
    } while (inc(&s));
    return k;
}

/* Smolyak grid — src/smolyak_api.jl:121-135: sources[k] = gridpoint(parent, shuffle[k]) */
int64_t orc_smolyak_grid(int grid_kind, int N, int B, int M, const orc_transform *tr, double *out, int64_t cap) {
    smolyak_state s;
    if (smolyak_state_setup(&s, grid_kind, N, B, M) != ORC_OK) return -1;
    int H = (int)orc_smolyak_parent_dimension(grid_kind, B, M);
    int64_t *sh = (int64_t *)malloc(sizeof(int64_t) * (size_t)(H + 2));
    double *src = (double *)malloc(sizeof(double) * (size_t)H);
    int64_t ns = orc_grid_shuffle(grid_kind, H, sh);
    if (ns != H) { free(sh); free(src); return -1; }
    for (int k = 0; k < H; k++) src[k] = orc_gridpoint(grid_kind, H, (int)sh[k]);
    int64_t k = 0;
    do {
        if (k < cap)
            for (int j = 0; j < N; j++) {
                double x = src[s.indices[j] - 1];
                /* transformed grid: src/generic_api.jl:296-300 + transformations.jl:129-134 */
                out[k * N + j] = tr ? orc_transform_from(&tr[j], x) : x;
            }
        k++;
    } while (inc(&s));
    free(sh); free(src);
    return k;
}

/* univariate grid (optionally transformed, src/generic_api.jl:296-300) */
int orc_cheb_grid(int grid_kind, int N, const orc_transform *tr, double *out) {
    if (N < 1) return ORC_ERR_ARGUMENT;
    for (int i = 1; i <= N; i++) {
        double x = orc_gridpoint(grid_kind, N, i);
        out[i - 1] = tr ? orc_transform_from(tr, x) : x;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * ∂ specification algebra — src/derivatives.jl:188-323.
 * A `Partials` is (len, I[len]) with no trailing zero.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int len; int I[ORC_MAXDIM]; } partials;

/* :228-238 */
static int is_strict_subset(const partials *p1, const partials *p2) {
    if (p1->len > p2->len) return 0;
    int strict = 0;
    int m = p1->len < p2->len ? p1->len : p2->len;
    for (int i = 0; i < m; i++) {
        if (p1->I[i] > p2->I[i]) return 0;
        strict |= p1->I[i] < p2->I[i];
    }
    return strict || p1->len < p2->len;
}
/* tuple `<` on NTuples of possibly different length (Julia: lexicographic, shorter prefix first) */
static int tuple_less(const partials *a, const partials *b) {
    int m = a->len < b->len ? a->len : b->len;
    for (int i = 0; i < m; i++) { if (a->I[i] < b->I[i]) return 1; if (a->I[i] > b->I[i]) return 0; }
    return a->len < b->len;
}
/* :247-251 */
static int partials_isless(const partials *p1, const partials *p2) {
    if (is_strict_subset(p1, p2)) return 1;
    if (is_strict_subset(p2, p1)) return 0;
    return tuple_less(p1, p2);
}

/* :266-275 — sort descending (stable insertion sort; `sort!(…; rev=true)` is stable), then drop
 * elements that are strict subsets of the previously kept one.  in/out: row-major [np][d] with
 * trailing zeros allowed (they are stripped, :214-220).  Returns the number kept. */
int orc_partials_minimal(int d, int np, const int32_t *in, int32_t *out) {
    if (d < 0 || d > ORC_MAXDIM || np < 0 || np > ORC_MAXCOMP) return ORC_ERR_ARGUMENT;
    partials ps[ORC_MAXCOMP];
    for (int p = 0; p < np; p++) {
        int len = d;
        while (len > 0 && in[p * d + len - 1] == 0) len--;
        ps[p].len = len;
        for (int i = 0; i < ORC_MAXDIM; i++) ps[p].I[i] = i < len ? in[p * d + i] : 0;
        for (int i = 0; i < len; i++) if (ps[p].I[i] < 0) return ORC_ERR_ARGUMENT;
    }
    /* stable sort with order rev(isless): a before b iff isless(b, a) */
    for (int i = 1; i < np; i++) {
        partials key = ps[i]; int j = i - 1;
        while (j >= 0 && partials_isless(&ps[j], &key)) { ps[j + 1] = ps[j]; j--; }
        ps[j + 1] = key;
    }
    int nk = 0; partials kept[ORC_MAXCOMP];
    for (int p = 0; p < np; p++)
        if (nk == 0 || !is_strict_subset(&ps[p], &kept[nk - 1])) kept[nk++] = ps[p];
    for (int p = 0; p < nk; p++) for (int i = 0; i < d; i++) out[p * d + i] = kept[p].I[i];
    return nk;
}

/* :284-300 — insertion-ordered union of CartesianIndices(I .+ 1) .- 1 (first index fastest).
 * `min` is a minimal representation [nk][d].  out: [ncomp][d]; returns ncomp. */
int orc_partials_canonical(int d, int nk, const int32_t *min, int32_t *out, int cap) {
    int nc = 0;
    for (int p = 0; p < nk; p++) {
        int idx[ORC_MAXDIM] = {0};
        for (;;) {
            int found = 0;
            for (int c = 0; c < nc && !found; c++) {
                int same = 1;
                for (int i = 0; i < d; i++) if (out[c * d + i] != idx[i]) { same = 0; break; }
                found = same;
            }
            if (!found) { if (nc >= cap) return ORC_ERR_ARGUMENT; for (int i = 0; i < d; i++) out[nc * d + i] = idx[i]; nc++; }
            int j = 0;
            while (j < d) { if (idx[j] < min[p * d + j]) { idx[j]++; break; } idx[j] = 0; j++; }
            if (j == d) break;
        }
    }
    return nc;
}

/* :308-316 */
void orc_partials_degrees(int d, int nk, const int32_t *min, int32_t *deg) {
    for (int i = 0; i < d; i++) deg[i] = 0;
    for (int p = 0; p < nk; p++) for (int i = 0; i < d; i++) if (min[p * d + i] > deg[i]) deg[i] = min[p * d + i];
}

/* ------------------------------------------------------------------------------------------
 * Univariate linear_combination — src/generic_api.jl:114-128 driven by chebyshev.jl:60-70,
 * after transform_to (generic_api.jl:291-294).
 *   theta: [N][nvec] (nvec fastest; nvec > 1 is the SVector-coefficient case, value only —
 *          derivatives.jl:15,19 have no jet method for SVector coefficients)
 *   y: [n]; out: [n][dp1] (nvec == 1) or [n][nvec] (dp1 == 1)
 *   absmode: accumulate |θ_k|·|b_k| instead (the scale S_r(x) of SURVEY.md Appendix C)
 * ---------------------------------------------------------------------------------------- */
int orc_uni_lincomb(int N, const orc_transform *tr, int dp1, int ext2, const double *theta, int nvec,
                    const double *y, int64_t n, double *out, int absmode, int nthreads) {
    if (N < 1 || dp1 < 1 || dp1 > ORC_MAXD1 || nvec < 1) return ORC_ERR_ARGUMENT;
    if (nvec > 1 && dp1 > 1) return ORC_ERR_UNSUPPORTED;
    if (tr && (tr->kind == TR_SEMI_INF || tr->kind == TR_INF) && (dp1 > 3 || (dp1 == 3 && !ext2))) return ORC_ERR_UNSUPPORTED;
    orc_transform ident = {TR_NONE, 0, 0.0, 1.0};
    if (!tr) tr = &ident;
    int err = 0;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int64_t p = 0; p < n; p++) {
        if (dp1 == 1) {
            /* Real path: B iterates Float64 */
            double x = transform_to_real(tr, y[p]);
            double fp = 1.0, fpp = x;
            double s[64]; double *sv = nvec <= 64 ? s : (double *)malloc(sizeof(double) * (size_t)nvec);
            for (int v = 0; v < nvec; v++) sv[v] = absmode ? fabs(theta[v]) * 1.0 : theta[v] * 1.0;
            for (int i = 2; i <= N; i++) {
                double f = (2 * x) * fp - fpp;
                fpp = fp; fp = f;
                double b = absmode ? fabs(f) : f;
                for (int v = 0; v < nvec; v++) {
                    double th = theta[(size_t)(i - 1) * nvec + v];
                    if (absmode) th = fabs(th);
                    sv[v] = sv[v] + th * b;
                }
            }
            for (int v = 0; v < nvec; v++) out[(size_t)p * nvec + v] = sv[v];
            if (sv != s) free(sv);
        } else {
            jet yj = jet_seed(y[p], dp1);
            jet x; memset(&x, 0, sizeof x);
            int rc = orc_transform_to_jet(tr, yj.c, dp1, ext2, x.c);
            if (rc != ORC_OK) { err = rc; continue; }
            jet one = jet_one(dp1);
            jet fp = one, fpp = x;
            jet b = one;
            if (absmode) for (int k = 0; k < dp1; k++) b.c[k] = fabs(b.c[k]);
            jet s = jet_scale(absmode ? fabs(theta[0]) : theta[0], b, dp1);
            for (int i = 2; i <= N; i++) {
                jet t = jet_scale(2.0, x, dp1);
                jet f = jet_sub(jet_mul(t, fp, dp1), fpp, dp1);
                fpp = fp; fp = f;
                b = f;
                double th = theta[i - 1];
                if (absmode) { th = fabs(th); for (int k = 0; k < dp1; k++) b.c[k] = fabs(b.c[k]); }
                s = jet_add(s, jet_scale(th, b, dp1), dp1);
            }
            for (int k = 0; k < dp1; k++) out[(size_t)p * dp1 + k] = s.c[k];
        }
    }
    return err;
}

/* Univariate basis_at / collocation_matrix — src/generic_api.jl:217-227.
 * out: column-major n×N matrix of elements of dp1 doubles: out[(j*n + i)*dp1 + r]. */
int orc_uni_basis_at(int N, const orc_transform *tr, int dp1, int ext2, const double *y, int64_t n, double *out) {
    if (N < 1 || dp1 < 1 || dp1 > ORC_MAXD1) return ORC_ERR_ARGUMENT;
    orc_transform ident = {TR_NONE, 0, 0.0, 1.0};
    if (!tr) tr = &ident;
    jet *tab = (jet *)malloc(sizeof(jet) * (size_t)N);
    for (int64_t p = 0; p < n; p++) {
        jet yj = jet_seed(y[p], dp1), x; memset(&x, 0, sizeof x);
        int rc = orc_transform_to_jet(tr, yj.c, dp1, ext2, x.c);
        if (rc != ORC_OK) { free(tab); return rc; }
        cheb_table(N, dp1, x, tab);
        for (int j = 0; j < N; j++) for (int k = 0; k < dp1; k++) out[((size_t)j * n + p) * dp1 + k] = tab[j].c[k];
    }
    free(tab);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Smolyak linear_combination / basis_at — src/smolyak_api.jl:89-110, smolyak_traversal.jl:373-379,
 * derivatives.jl:437-445 (coordinate expansion), :491,:502-511 (_product), :467-473, and
 * generic_api.jl:114-128.
 *   spec: ncomp × d orders (canonical expansion).  ncomp == 0 means "plain point" (no ∂): real
 *         tables and `prod`, output 1 value (or nvec values for SVector coefficients).
 *   y: [n][d]; theta: [K][nvec]; out: [n][ncomp] | [n][nvec]
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int grid_kind, d, B, M, H;
    int64_t K;
    int ncomp;                         /* 0: plain reals */
    int32_t orders[ORC_MAXCOMP][ORC_MAXDIM];
    int deg[ORC_MAXDIM];               /* per-coordinate jet order M_j */
    int ext2;
    orc_transform tr[ORC_MAXDIM];
} smolyak_cfg;

static int smolyak_cfg_make(smolyak_cfg *c, int grid_kind, int d, int B, int M, const orc_transform *tr,
                            int ncomp, const int32_t *orders, int ext2) {
    if (d < 1 || d > ORC_MAXDIM || ncomp < 0 || ncomp > ORC_MAXCOMP) return ORC_ERR_ARGUMENT;
    if (M > B) M = B;
    c->grid_kind = grid_kind; c->d = d; c->B = B; c->M = M; c->ext2 = ext2;
    c->H = (int)orc_smolyak_parent_dimension(grid_kind, B, M);
    c->K = orc_smolyak_length(grid_kind, d, B, M);
    c->ncomp = ncomp;
    for (int j = 0; j < d; j++) { c->deg[j] = 0; if (tr) c->tr[j] = tr[j]; else { c->tr[j].kind = TR_NONE; c->tr[j].p0 = 0; c->tr[j].p1 = 1; } }
    for (int q = 0; q < ncomp; q++) for (int j = 0; j < d; j++) {
        int o = orders[q * d + j];
        if (o < 0 || o >= ORC_MAXD1) return ORC_ERR_ARGUMENT;
        c->orders[q][j] = o; if (o > c->deg[j]) c->deg[j] = o;
    }
    for (int j = 0; j < d; j++)
        if ((c->tr[j].kind == TR_SEMI_INF || c->tr[j].kind == TR_INF) && (c->deg[j] > 2 || (c->deg[j] == 2 && !ext2)))
            return ORC_ERR_UNSUPPORTED;
    return ORC_OK;
}

/* one point: fills tab[j][i] (jets of order deg[j]) — smolyak_api.jl:89-93 */
static int smolyak_tables(const smolyak_cfg *c, const double *y, jet *tab /* [d][H] */) {
    for (int j = 0; j < c->d; j++) {
        int dp1 = c->deg[j] + 1;
        jet yj = jet_seed(y[j], dp1), x; memset(&x, 0, sizeof x);
        int rc = orc_transform_to_jet(&c->tr[j], yj.c, dp1, c->ext2, x.c);
        if (rc != ORC_OK) return rc;
        cheb_table(c->H, dp1, x, tab + (size_t)j * c->H);
    }
    return ORC_OK;
}

/* Full gradient [value, ∂1..∂d] (the headline request): the same per-term arithmetic as the generic loop below — every
 * component's product formed left to right over the dimensions (derivatives.jl:502-511), then s = s + θ·b unfused
 * (generic_api.jl:114-128) — with the component structure known at compile time, as Julia's generated code for a concrete
 * ∂ type has it: no order lookups, the (T, T') of the term's d indices fetched once.  Bit-identical to the generic loop (the
 * golden fixtures were written by it); only faster, so that the timed CPU baseline is not handicapped by interpretation. */
static inline __attribute__((always_inline)) void grad_term(const int D, const jet *tab, int64_t H, const int64_t *idx, double th, double *sv, int first) {
    /* component q in lane q of a short vector (what an SVector-valued ∂Expansion is in the reference): per dimension one
     * vector multiply by (T, …, T, T', T, …) with T' in lane j + 1 — every lane still multiplies left to right */
    enum { W = ORC_MAXDIM + 1 };
    double b[W], f[W];
    const jet *e0 = tab + idx[0] - 1;
    for (int q = 0; q <= D; q++) b[q] = e0->c[0];
    b[1] = e0->c[1];
    for (int j = 1; j < D; j++) {
        const jet *e = tab + (size_t)j * H + idx[j] - 1;
        for (int q = 0; q <= D; q++) f[q] = e->c[0];
        f[j + 1] = e->c[1];
        for (int q = 0; q <= D; q++) b[q] = b[q] * f[q];
    }
    if (first) for (int q = 0; q <= D; q++) sv[q] = th * b[q];
    else for (int q = 0; q <= D; q++) sv[q] = sv[q] + th * b[q];
}
static int is_full_gradient(const smolyak_cfg *c, int d, int ncomp) {
    if (ncomp != d + 1) return 0;
    for (int q = 0; q <= d; q++) for (int j = 0; j < d; j++) if (c->orders[q][j] != (q == j + 1 ? 1 : 0)) return 0;
    return 1;
}

int orc_smolyak_lincomb(int grid_kind, int d, int B, int M, const orc_transform *tr,
                        int ncomp, const int32_t *orders, int ext2,
                        const double *theta, int nvec, const double *y, int64_t n, double *out,
                        int absmode, int nthreads) {
    smolyak_cfg c;
    int rc = smolyak_cfg_make(&c, grid_kind, d, B, M, tr, ncomp, orders, ext2);
    if (rc != ORC_OK) return rc;
    if (nvec < 1 || nvec > 4096) return ORC_ERR_ARGUMENT;
    if (nvec > 1 && ncomp > 0) return ORC_ERR_UNSUPPORTED;   /* no _mul(::SVector, ::∂Expansion) method */
    int err = 0;
    const int gradfast = !absmode && ncomp > 0 && getenv("SK_ORACLE_GENERIC") == NULL && is_full_gradient(&c, d, ncomp);
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        jet *tab = (jet *)malloc(sizeof(jet) * (size_t)d * (size_t)c.H);
        double *rtab = (double *)malloc(sizeof(double) * (size_t)d * (size_t)c.H);
        double *sv = (double *)malloc(sizeof(double) * (size_t)(nvec > ORC_MAXCOMP ? nvec : ORC_MAXCOMP));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t p = 0; p < n; p++) {
            const double *yp = y + (size_t)p * d;
            smolyak_state s;
            smolyak_state_setup(&s, grid_kind, d, B, M);
            if (ncomp == 0) {
                /* plain point: real tables, _product(nothing, x) = prod(x) (left to right) */
                for (int j = 0; j < d; j++) cheb_table_real(c.H, transform_to_real(&c.tr[j], yp[j]), rtab + (size_t)j * c.H);
                int64_t k = 0;
                do {
                    double b = rtab[s.indices[0] - 1];
                    for (int j = 1; j < d; j++) b = b * rtab[(size_t)j * c.H + s.indices[j] - 1];
                    if (absmode) b = fabs(b);
                    for (int v = 0; v < nvec; v++) {
                        double th = theta[(size_t)k * nvec + v];
                        if (absmode) th = fabs(th);
                        sv[v] = k == 0 ? th * b : sv[v] + th * b;
                    }
                    k++;
                } while (inc(&s));
                for (int v = 0; v < nvec; v++) out[(size_t)p * nvec + v] = sv[v];
            } else {
                int rc2 = smolyak_tables(&c, yp, tab);
                if (rc2 != ORC_OK) { err = rc2; continue; }
                int64_t k = 0;
                if (gradfast) {
                    do {
                        switch (d) {      /* constant trip counts: the compiler unrolls the products */
                        case 2: grad_term(2, tab, c.H, s.indices, theta[k], sv, k == 0); break;
                        case 3: grad_term(3, tab, c.H, s.indices, theta[k], sv, k == 0); break;
                        case 4: grad_term(4, tab, c.H, s.indices, theta[k], sv, k == 0); break;
                        case 6: grad_term(6, tab, c.H, s.indices, theta[k], sv, k == 0); break;
                        default: grad_term(d, tab, c.H, s.indices, theta[k], sv, k == 0); break;
                        }
                        k++;
                    } while (inc(&s));
                } else
                do {
                    double th = absmode ? fabs(theta[k]) : theta[k];
                    for (int q = 0; q < ncomp; q++) {
                        double b = tab[s.indices[0] - 1].c[c.orders[q][0]];
                        for (int j = 1; j < d; j++) b = b * tab[(size_t)j * c.H + s.indices[j] - 1].c[c.orders[q][j]];
                        if (absmode) b = fabs(b);
                        sv[q] = k == 0 ? th * b : sv[q] + th * b;
                    }
                    k++;
                } while (inc(&s));
                for (int q = 0; q < ncomp; q++) out[(size_t)p * ncomp + q] = sv[q];
            }
        }
        free(tab); free(rtab); free(sv);
    }
    return err;
}

/* Smolyak basis_at / collocation_matrix: out column-major n×K of elements of E doubles
 * (E = 1 for plain points, ncomp otherwise): out[(k*n + i)*E + q]. */
int orc_smolyak_basis_at(int grid_kind, int d, int B, int M, const orc_transform *tr,
                         int ncomp, const int32_t *orders, int ext2,
                         const double *y, int64_t n, double *out) {
    smolyak_cfg c;
    int rc = smolyak_cfg_make(&c, grid_kind, d, B, M, tr, ncomp, orders, ext2);
    if (rc != ORC_OK) return rc;
    int E = ncomp == 0 ? 1 : ncomp;
    jet *tab = (jet *)malloc(sizeof(jet) * (size_t)d * (size_t)c.H);
    double *rtab = (double *)malloc(sizeof(double) * (size_t)d * (size_t)c.H);
    for (int64_t p = 0; p < n; p++) {
        const double *yp = y + (size_t)p * d;
        smolyak_state s;
        smolyak_state_setup(&s, grid_kind, d, B, M);
        int64_t k = 0;
        if (ncomp == 0) {
            for (int j = 0; j < d; j++) cheb_table_real(c.H, transform_to_real(&c.tr[j], yp[j]), rtab + (size_t)j * c.H);
            do {
                double b = rtab[s.indices[0] - 1];
                for (int j = 1; j < d; j++) b = b * rtab[(size_t)j * c.H + s.indices[j] - 1];
                out[((size_t)k * n + p) * E] = b;
                k++;
            } while (inc(&s));
        } else {
            rc = smolyak_tables(&c, yp, tab);
            if (rc != ORC_OK) break;
            do {
                for (int q = 0; q < ncomp; q++) {
                    double b = tab[s.indices[0] - 1].c[c.orders[q][0]];
                    for (int j = 1; j < d; j++) b = b * tab[(size_t)j * c.H + s.indices[j] - 1].c[c.orders[q][j]];
                    out[((size_t)k * n + p) * E + q] = b;
                }
                k++;
            } while (inc(&s));
        }
    }
    free(tab); free(rtab);
    return rc;
}

/* batched transform_to / transform_from over coordinates (transformations.jl:112-139) */
int orc_transform_to_batch(const orc_transform *tr, int d, int dp1, int ext2, const double *y, int64_t n, double *out) {
    for (int64_t p = 0; p < n; p++)
        for (int j = 0; j < d; j++) {
            jet yj = jet_seed(y[(size_t)p * d + j], dp1);
            int rc = orc_transform_to_jet(&tr[j], yj.c, dp1, ext2, out + ((size_t)p * d + j) * dp1);
            if (rc != ORC_OK) return rc;
        }
    return ORC_OK;
}
int orc_transform_from_batch(const orc_transform *tr, int d, const double *x, int64_t n, double *out) {
    for (int64_t p = 0; p < n; p++)
        for (int j = 0; j < d; j++) out[(size_t)p * d + j] = orc_transform_from(&tr[j], x[(size_t)p * d + j]);
    return ORC_OK;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
