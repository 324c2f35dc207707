/*
 * spectralkit_b200.h — C ABI of the B200-native (sm_100a) batched basis-evaluation library.
 *
 * The reference (tpapp/SpectralKit.jl v0.16.2) has no FFI or plugin interface: its boundary is
 * the set of exported Julia generic functions (src/generic_api.jl:5-7) plus the constructors of
 * bases, transformations and derivative requests.  This header is the C boundary a Julia
 * `ccall` shim (spectralkit.jl_b200/julia/SpectralKitB200.jl), the Python `ctypes` mirror
 * (spectralkit.jl_b200/api.py) or any C caller binds.  Every entry point cites the reference
 * lines whose semantics it reproduces (paths relative to the reference checkout).
 *
 * Conventions
 *   - all floating point is IEEE binary64; all indices returned are 1-based like Julia's;
 *   - every function returns an int status (SK_OK == 0); sk_last_error() gives a thread-local
 *     message.  SK_ERR_ARGUMENT ↔ Julia ArgumentError, SK_ERR_DOMAIN ↔ DomainError,
 *     SK_ERR_UNSUPPORTED ↔ `error("… not implemented yet")` / MethodError;
 *   - the caller owns every buffer; the library keeps nothing but what is inside a plan;
 *   - plans are immutable after creation and may be used concurrently from several threads,
 *     each call on its own stream;
 *   - there is NO CPU fallback: evaluation entry points fail with SK_ERR_NO_DEVICE /
 *     SK_ERR_CUDA if no usable sm_100 device exists.  Only integer metadata and grids
 *     (initialisation-time, tiny) are computed on the host.
 *
 * Memory layouts (identical to the Julia object layouts, so a shim passes pointers through)
 *   points   x   : univariate [n]; Smolyak [n][d]   (Vector{NTuple{d,Float64}} / d×n Matrix)
 *   coeffs   θ   : [K][nvec], nvec fastest          (Vector{Float64} / Vector{SVector{nvec}})
 *   lincomb  out : [n][E]  with E = D+1 (𝑑Expansion), ncomp (∂Expansion), or nvec (SVector)
 *   basis_at out : column-major n×K matrix of E-double elements: out[(k*n + i)*E + q]
 *                  (= Matrix{Float64|𝑑Expansion|∂Expansion}, src/generic_api.jl:217-227)
 */
#ifndef SPECTRALKIT_B200_H
#define SPECTRALKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_MAX_DIM 16        /* Smolyak dimension d */
#define SK_MAX_BLOCKS 8      /* block cap M */
#define SK_MAX_ORDER 5       /* univariate derivative order D (test/test_chebyshev.jl:93 uses 5) */
#define SK_MAX_COMP 64       /* components of a ∂ specification */
#define SK_MAX_TERMS (1 << 24) /* dimension(basis) of a Smolyak basis: tables are sized by it (SK_ERR_UNSUPPORTED above) */

enum sk_status {
    SK_OK = 0,
    SK_ERR_ARGUMENT = -1,
    SK_ERR_DOMAIN = -2,
    SK_ERR_UNSUPPORTED = -3,
    SK_ERR_CUDA = -4,
    SK_ERR_NO_DEVICE = -5,
    SK_ERR_NCCL = -6
};

enum sk_basis_kind { SK_BASIS_CHEBYSHEV = 0, SK_BASIS_SMOLYAK = 1 };
/* src/generic_api.jl:174-192 */
enum sk_grid_kind { SK_GRID_INTERIOR = 0, SK_GRID_ENDPOINT = 1, SK_GRID_INTERIOR2 = 2 };
/* src/transformations.jl:149,206,280 */
enum sk_transform_kind { SK_TR_NONE = 0, SK_TR_BOUNDED_LINEAR = 1, SK_TR_SEMI_INF_RATIONAL = 2, SK_TR_INF_RATIONAL = 3 };

/* Chebyshev(grid_kind, N) (src/chebyshev.jl:12-32) or
 * smolyak_basis(Chebyshev, grid_kind, SmolyakParameters(B, M), d) (src/smolyak_api.jl:63-75). */
typedef struct sk_basis_desc {
    int32_t kind;       /* sk_basis_kind */
    int32_t grid_kind;  /* sk_grid_kind */
    int32_t N;          /* Chebyshev: number of basis functions, >= 1 */
    int32_t d;          /* Smolyak: dimension, 1..SK_MAX_DIM */
    int32_t B;          /* Smolyak: sum of block indices <= B */
    int32_t M;          /* Smolyak: each block index <= M (M > B is clamped, smolyak_traversal.jl:15-16) */
} sk_basis_desc;

/* The reference's stored fields: BoundedLinear (m, s) (transformations.jl:150-153);
 * SemiInfRational / InfRational (A, L) (:207-210, :281-284). */
typedef struct sk_transform {
    int32_t kind;       /* sk_transform_kind */
    int32_t reserved;
    double p0;          /* m or A */
    double p1;          /* s or L */
} sk_transform;

/* Derivative request.
 *   Chebyshev: `order` = D of 𝑑^D (src/derivatives.jl:66-107); ncomp/orders ignored.
 *   Smolyak:   ncomp == 0 → a plain point (no ∂); otherwise the canonical expansion of a
 *              ∂ specification (src/derivatives.jl:284-300): orders[c*d + j] = derivative order
 *              of component c along coordinate j.  sk_partials_canonical builds it. */
typedef struct sk_deriv_spec {
    int32_t order;
    int32_t ncomp;
    const int32_t *orders;
    int32_t allow_rational_d2;  /* EXTENSION: 2nd derivatives through the rational maps (the
                                   reference throws, transformations.jl:266,332); 0 = reference behaviour */
} sk_deriv_spec;

typedef struct sk_plan sk_plan;  /* opaque */

typedef struct sk_plan_info {
    int32_t kind, grid_kind, d, B, M, device;
    int64_t K;              /* dimension(basis) */
    int64_t H;              /* parent dimension (Smolyak) or N */
    int32_t E;              /* doubles per output element for nvec == 1 */
    int32_t fast_path;      /* linear_combination fast kernel: 0 direct only, 1 nested (run program),
                               2 nested with compile-time unrolled leaves, 3 nested for any ∂ specification,
                               4 unrolled second-order-jet leaves (∂ requests of total order <= 2) (univariate: 1) */
    int64_t n_runs;         /* level-1 runs of the traversal (K_2) */
    double flops_per_point_fast;    /* closed-form FP64 flop count of the fast kernel, per point */
    double flops_per_point_direct;  /* F_direct of SURVEY.md §8(d): what the reference executes */
    double bytes_per_point;         /* algorithmic HBM bytes per point: 8*d + 8*E */
} sk_plan_info;

/* flags for the evaluation entry points */
#define SK_MEM_HOST 0u          /* x, theta, out are host pointers; copies happen inside the call */
#define SK_MEM_DEVICE 1u        /* x, theta, out are device pointers on the plan's device */
#define SK_MODE_FAST 0u         /* FMA-contracted, nested evaluation; within the stated tolerance */
#define SK_MODE_EXACT 2u        /* reference operation order, no contraction: bit-identical to the
                                   reference restatement (sequential sum, mul-then-add), for bases of
                                   any size (full tables on chip where they fit, tables streamed chunk by
                                   chunk otherwise).  sk_basis_at / collocation matrices are always
                                   bit-identical */

/* ---------------------------------------------------------------- library */
const char *sk_version(void);
const char *sk_last_error(void);           /* thread-local */
int sk_device_count(int *count);           /* SK_ERR_NO_DEVICE if no CUDA device */

/* ---------------------------------------------------------------- constructors, metadata (host) */
/* BoundedLinear(a,b) / SemiInfRational(A,L) / InfRational(A,L) with the reference's DomainError
 * checks (transformations.jl:154-161, 211-215, 285-289). */
int sk_transform_make(int kind, double a, double b, sk_transform *out);
/* nesting_total_length / nesting_block_length (smolyak_traversal.jl:69-84,112-116,146-148) */
int sk_nesting_total_length(int grid_kind, int b, int64_t *out);
int sk_nesting_block_length(int grid_kind, int b, int64_t *out);
/* collect(SmolyakGridShuffle(grid_kind, len)) (smolyak_traversal.jl:86-168); out[len] */
int sk_grid_shuffle(int grid_kind, int64_t len, int64_t *out);
/* dimension(basis) (chebyshev.jl:34, smolyak_api.jl:82, smolyak_traversal.jl:238-256) */
int sk_dimension(const sk_basis_desc *basis, int64_t *K);
/* highest_visited_index / parent Chebyshev dimension (smolyak_traversal.jl:307,317) */
int sk_parent_dimension(const sk_basis_desc *basis, int64_t *H);
/* collect(SmolyakIndices) (smolyak_traversal.jl:294-334): out[K][d], 1-based, traversal order */
int sk_smolyak_indices(const sk_basis_desc *basis, int32_t *out);
/* collect(grid(basis ∘ transformations)) (chebyshev.jl:88-127, smolyak_api.jl:121-135,
 * generic_api.jl:296-300).  tr == NULL → untransformed.  out[N] or out[K][d].
 * sinpi/cospi are correctly rounded at the rounded Float64 argument. */
int sk_grid(const sk_basis_desc *basis, const sk_transform *tr, double *out);
/* ∂ algebra: minimal representation + canonical expansion + per-coordinate degrees
 * (derivatives.jl:266-316).  partials[np][d] (trailing zeros allowed); out_orders[cap][d]. */
int sk_partials_canonical(int d, int np, const int32_t *partials, int32_t *out_orders, int cap,
                          int32_t *ncomp, int32_t *degrees);

/* ---------------------------------------------------------------- refinement between bases (host; index merge) */
/* is_subset_basis(basis1, basis2) (chebyshev.jl:140-142, smolyak_api.jl:141-153, generic_api.jl:254,314-317):
 * tr1/tr2 NULL for untransformed bases; transformed bases need identical transformations. */
int sk_is_subset_basis(const sk_basis_desc *basis1, const sk_transform *tr1, const sk_basis_desc *basis2,
                       const sk_transform *tr2, int *out);
/* augment_coefficients(basis1, basis2, θ1) (chebyshev.jl:133-138, smolyak_api.jl:155-208, generic_api.jl:319-322):
 * θ2[dimension(basis2)][nvec] gets θ1's entries at the shared indices (one ordered pass, the reference's
 * PaddingIterator) and zeros elsewhere.  SK_ERR_ARGUMENT unless is_subset_basis and theta1_len == dimension(basis1). */
int sk_augment_coefficients(const sk_basis_desc *basis1, const sk_transform *tr1, const sk_basis_desc *basis2,
                            const sk_transform *tr2, const double *theta1, int64_t theta1_len, int nvec, double *theta2);

/* ---------------------------------------------------------------- plans */
/* tr: NULL (untransformed basis) or 1 (Chebyshev) / d (Smolyak) transformations: basis ∘ t
 * (generic_api.jl:263-274).  spec: NULL = values only. */
int sk_plan_create(const sk_basis_desc *basis, const sk_transform *tr, const sk_deriv_spec *spec,
                   int device, sk_plan **out);
int sk_plan_destroy(sk_plan *plan);
int sk_plan_get_info(const sk_plan *plan, sk_plan_info *info);

/* ---------------------------------------------------------------- the hot path (GPU only) */
/* linear_combination(basis, θ, x) for n points (generic_api.jl:114-138).  nvec > 1 is the
 * SVector-coefficient case (derivatives.jl:15,19; values only).  `theta_len` is checked against
 * dimension(basis): SK_ERR_ARGUMENT "not enough/too many coefficients" (generic_api.jl:99-102). */
int sk_linear_combination(const sk_plan *plan, const double *theta, int64_t theta_len, int nvec,
                          const double *x, int64_t n, double *out, uint32_t flags, void *stream);
/* collocation_matrix(basis, x) == basis_at at n points (generic_api.jl:217-227); always in the
 * reference's operation order (bit-identical basis values), for any basis and any ∂ request.  ld = leading
 * dimension in elements (>= n); only the first n entries of every column are written.  Host buffers are
 * streamed in chunks of points (the full matrix never has to fit on the device in one piece). */
int sk_basis_at(const sk_plan *plan, const double *x, int64_t n, double *out, int64_t ld,
                uint32_t flags, void *stream);
/* transform_to / transform_from over n points of d coordinates (transformations.jl:112-139,
 * 178-195, 244-267, 309-333).  order = D of the seeded jets for transform_to:
 * out[n][d][D+1]; transform_from: out[n][d]. */
int sk_transform_to(const sk_transform *tr, int d, int order, int allow_rational_d2, const double *y,
                    int64_t n, double *out, uint32_t flags, int device, void *stream);
int sk_transform_from(const sk_transform *tr, int d, const double *x, int64_t n, double *out,
                      uint32_t flags, int device, void *stream);

/* ---------------------------------------------------------------- fitting (initialisation time) */
/* θ = collocation_matrix(basis) \ y on the basis' own grid (docs/src/index.md:92, test/test_smolyak.jl:20-21,
 * test/test_generic_api.jl:20).  The collocation matrix is generated on the device by the basis_at kernels; the
 * dense LU with partial pivoting (what Julia's `\` does on a square matrix) and the triangular solves are the library's
 * own kernels (blocked right-looking factorisation, trailing updates on the FP64 tensor path; csrc/sk_solve.cu).  The
 * plan must carry no derivative request; nrhs <= 96.  y and theta: [K][nrhs], nrhs fastest (the layout of Vector{SVector{nrhs}}; nrhs = 1
 * is a plain vector).  SK_ERR_ARGUMENT if the matrix is singular (Julia: SingularException).  Synchronises `stream`. */
int sk_collocation_solve(const sk_plan *plan, const double *y, int nrhs, double *theta, uint32_t flags, void *stream);

/* ---------------------------------------------------------------- Experimental: policy functions and residual sums */
/* bridge(outer, inner) (src/experimental.jl:242-266): x ↦ transform_from(PM1, outer, transform_to(PM1, inner, x));
 * kind SK_TR_NONE on both sides is the identity. */
typedef struct sk_bridge {
    sk_transform outer, inner;
} sk_bridge;
/* make_policy_functions (src/experimental.jl:156-166) evaluated for n points in one pass: policy m is
 * bridges[m] ∘ linear_combination(basis, coefficients[m·K : (m+1)·K]) with K = dimension(basis); the stacked vector
 * becomes the K×npol coefficient block of the hot path.  out[n][npol].  bridges == NULL: no transformation.
 * SK_ERR_ARGUMENT unless coefficients_len == K·npol (the reference's @argcheck).  1 <= npol <= 64.  Synchronises. */
int sk_policy_functions(const sk_plan *plan, const double *coefficients, int64_t coefficients_len, int npol,
                        const sk_bridge *bridges, const double *x, int64_t n, double *out, uint32_t flags, void *stream);
/* sum_abs2 reduced over a residual array (src/experimental.jl:218-233): Σ r², fixed partition and fixed reduction tree
 * (reproducible; the reference adds left to right, so the last bits may differ).  r: host or device by flags. */
int sk_sum_abs2(const double *r, int64_t count, double *result_host, uint32_t flags, int device, void *stream);
/* The least-squares special case fused end to end: Σ_{i,m} (policy_m(x_i) − target[i][m])² without materialising the
 * policy values or the residuals on the host. */
int sk_policy_residual_ssq(const sk_plan *plan, const double *coefficients, int64_t coefficients_len, int npol,
                           const sk_bridge *bridges, const double *x, int64_t n, const double *target,
                           double *result_host, uint32_t flags, void *stream);

/* ---------------------------------------------------------------- multi-GPU (one process, G devices) */
/* Points are independent: contiguous slices of ceil(n/G) points per device, plans replicated,
 * θ copied to each device, no reduction, outputs land in the caller's host buffer.
 * plans[g] must live on distinct devices.  Host pointers only.  (The reference is single-threaded; the
 * caller-side loop this replaces is the broadcast of test/test_generic_api.jl:22.) */
int sk_linear_combination_sharded(sk_plan *const *plans, int ngpu, const double *theta, int64_t theta_len,
                                  int nvec, const double *x, int64_t n, double *out, uint32_t flags);

/* Device-resident shards (SURVEY.md §8(b) "sk_*_sharded(comm, …)", §8(e)).  A communicator is a single-process
 * NCCL clique over `devices` (NCCL is bound at run time: SK_ERR_NCCL if libnccl.so.2 cannot be loaded or
 * fails); it also owns one stream per device on which the calls below run. */
typedef struct sk_comm sk_comm;  /* opaque */
int sk_comm_create(int ngpu, const int *devices, sk_comm **out);
int sk_comm_destroy(sk_comm *comm);
int sk_comm_size(const sk_comm *comm, int *ngpu, int *nccl_version);
/* C1: one broadcast of the coefficients from device `root` to theta_dev[g] on every device (count doubles). */
int sk_broadcast_coefficients(sk_comm *comm, double *const *theta_dev, int64_t count, int root);
/* linear_combination over device-resident slices: device g holds points [g·per, min(n, (g+1)·per)), per = ceil(n/G),
 * in x_dev[g], and the coefficients in theta_dev[g]; plans[g] lives on the communicator's device g.
 * gather == 0: out_dev[g] = that slice's results [slice][E] (outputs stay sharded: the default of SURVEY.md §8(e)).
 * gather == 2 (C2 fused with the compute): the same full buffers, filled without NCCL and without a pass after the compute.
 *   Plans served by the six-dimensional leaf kernel (cfg 4) run ONE kernel per device whose threads store every result into
 *   all devices' buffers themselves, peer to peer over NVLink (staged per warp so that remote writes are whole sectors);
 *   other plans are evaluated in chunks whose results the copy engines push to the peers while the next chunk is
 *   evaluated.  ms_gather then is what remains after the last kernel.  Needs peer access between all devices of the
 *   communicator; otherwise it behaves like gather == 1.
 * gather != 0 (C2): every out_dev[g] is a full buffer of per·G rows; device g writes its slice at row g·per and an
 *   in-place ncclAllGather over NVLink completes all G copies (rows >= n are padding).
 * ms_compute / ms_gather: optional [G] arrays, device-timed duration of the two phases per device.  Synchronises. */
int sk_linear_combination_sharded_device(sk_comm *comm, sk_plan *const *plans, const double *const *theta_dev,
                                         int64_t theta_len, int nvec, const double *const *x_dev, int64_t n,
                                         double *const *out_dev, int gather, uint32_t flags, float *ms_compute,
                                         float *ms_gather);

/* The fused gather for callers that hold ONE PROCESS PER GPU (torchrun, MPI): linear_combination over this rank's
 * device-resident slice whose results are ALSO stored into n_peer (<= 7) buffers of other GPUs — device pointers mapped into
 * this process (CUDA IPC, torch symmetric memory, cuMem*), already offset to this slice's first row.  One coefficient vector,
 * fast mode.  *fused = 1: the plan's kernel stored to the peers itself (plans served by the six-dimensional leaf kernel: cfg 4);
 * *fused = 0: only `out` was written and the caller copies.  The writes are complete when the work on `stream` is.  No
 * reference counterpart (the reference is single-threaded); north_star's "results gathered … over NVLink". */
int sk_linear_combination_peers(const sk_plan *plan, const double *theta, int64_t theta_len, const double *x, int64_t n,
                                double *out, double *const *peer_out, int n_peer, uint32_t flags, void *stream, int *fused);

/* ---------------------------------------------------------------- measurement helpers */
/* Measures the FP64 pipe on `device`: which = 0 DFMA (CUDA-core), 1 DMMA (mma.sync m8n8k4 f64),
 * 2 both at once (half the warps each: shows whether they share one pipe), 3 DFMA with three distinct register
 * operands per instruction (the register-file-bound rate of code without operand reuse).
 * Runs for about `seconds`; returns TFLOP/s of the best single launch timed alone (burst) and of launches queued back to
 * back between two events with no host synchronisation in between (sustained). */
int sk_fp64_peak(int device, int which, double seconds, double *tflops_burst, double *tflops_sustained);
/* Closed-form FP64 flop count per point of linear_combination with a K×nvec coefficient block (the DMMA kernel:
 * 2·K·nvec in the contraction + the generation of the collocation tile); nvec == 1: flops_per_point_fast. */
double sk_coefficient_block_flops(const sk_plan *plan, int nvec);
/* Times the library's LU factorisation on a K×K column-major device matrix (overwritten): mean device milliseconds over
 * `reps` runs, (2/3)·K³ flop / time in TFLOP/s, and the milliseconds spent in the DMMA trailing updates alone. */
int sk_lu_benchmark(double *A_dev, int64_t K, int reps, double *ms, double *tflops, double *ms_gemm);
/* Number of kernel launches issued by this library since load (all threads). */
int64_t sk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRALKIT_B200_H */
