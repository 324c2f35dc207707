"""Host-side mirror of the SpectralKit.jl API for the batched hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference (the executable stand-in for
the Julia `ccall` shim in julia/SpectralKitB200.jl, which cannot run in this environment):

    basis = smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(4), 6)
    ct    = coordinate_transformations(BoundedLinear(0, 4), SemiInfRational(1, 2), ...)
    y     = linear_combination(basis @ ct, θ, gradient(6)(x))     # x: (n, 6) -> (n, 7)

Differences forced by batching: points come in arrays — univariate `(n,)`,
Smolyak `(n, d)`; a Python scalar / length-d vector is one point.  Results with derivatives are
arrays whose last axis is the reference's container order: `𝑑Expansion` → `[f, f′, …]`
(derivatives.jl:27-57), `∂Expansion` → canonical expansion order (derivatives.jl:284-300).
`@` stands for Julia's `∘`.

numpy arrays are evaluated through host buffers (copies inside the call); torch CUDA tensors are
evaluated in place on their device and a CUDA tensor is returned.  Every evaluation runs in the
CUDA library; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
import math
import warnings
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from ._lib import ArgumentError, DomainError, UnsupportedError, CudaError  # noqa: F401  (re-exported)

try:  # torch is only the device-memory plumbing
    import torch
except Exception:  # pragma: no cover
    torch = None


# ----------------------------------------------------------------------------- grids (generic_api.jl:174-192)
@dataclass(frozen=True)
class _GridKind:
    code: int
    name: str

    def __repr__(self):
        return f"{self.name}()"

    def __call__(self):
        return self


InteriorGrid = _GridKind(L.SK_GRID_INTERIOR, "InteriorGrid")
EndpointGrid = _GridKind(L.SK_GRID_ENDPOINT, "EndpointGrid")
InteriorGrid2 = _GridKind(L.SK_GRID_INTERIOR2, "InteriorGrid2")


def _grid_code(g) -> int:
    if not isinstance(g, _GridKind):
        raise TypeError(f"grid kind must be InteriorGrid(), EndpointGrid() or InteriorGrid2(), got {g!r}")
    return g.code


# ----------------------------------------------------------------------------- domains (domains.jl)
@dataclass(frozen=True)
class UnivariateDomain:
    min: float
    max: float

    def __iter__(self):
        return iter((self.min, self.max))


PM1 = UnivariateDomain(-1.0, 1.0)


@dataclass(frozen=True)
class CoordinateDomains:
    domains: Tuple[UnivariateDomain, ...]

    def __len__(self):
        return len(self.domains)

    def __getitem__(self, i):
        return self.domains[i]


def coordinate_domains(*domains):
    if len(domains) == 1 and isinstance(domains[0], (tuple, list)):
        domains = tuple(domains[0])
    return CoordinateDomains(tuple(domains))


def extrema(dom: UnivariateDomain):
    return (dom.min, dom.max)


# ----------------------------------------------------------------------------- transformations (transformations.jl)
class _UnivariateTransformation:
    kind = L.SK_TR_NONE

    def __init__(self, a, b):
        t = L.Transform()
        L.check(L.lib.sk_transform_make(self.kind, float(a), float(b), C.byref(t)))
        self._t = t

    @property
    def p0(self):
        return self._t.p0

    @property
    def p1(self):
        return self._t.p1

    def __eq__(self, other):
        return type(self) is type(other) and (self.p0, self.p1) == (other.p0, other.p1)

    def __hash__(self):
        return hash((type(self).__name__, self.p0, self.p1))


class BoundedLinear(_UnivariateTransformation):
    """y = x·s + m on (a, b) (transformations.jl:149-200); DomainError unless finite a < b."""
    kind = L.SK_TR_BOUNDED_LINEAR
    m = property(lambda self: self.p0)
    s = property(lambda self: self.p1)

    def __repr__(self):
        return f"({self.m - self.s},{self.m + self.s}) ↔ domain [linear transformation]"


class SemiInfRational(_UnivariateTransformation):
    """y = A + L(1+x)/(1−x) on [A, ∞) (L > 0) or (−∞, A] (L < 0) (transformations.jl:206-274)."""
    kind = L.SK_TR_SEMI_INF_RATIONAL
    A = property(lambda self: self.p0)
    L = property(lambda self: self.p1)

    def __repr__(self):
        dom = f"({self.A},∞)" if self.L > 0 else "(-∞,A)"
        return f"{dom} ↔ domain [rational transformation with scale {self.L}]"


class InfRational(_UnivariateTransformation):
    """y = A + L·x/√(1−x²) on (−∞, ∞), L > 0 (transformations.jl:280-335)."""
    kind = L.SK_TR_INF_RATIONAL
    A = property(lambda self: self.p0)
    L = property(lambda self: self.p1)

    def __repr__(self):
        return f"(-∞,∞) ↔ domain [rational transformation with center {self.A}, scale {self.L}]"


class CoordinateTransformations:
    """Per-coordinate transformations (transformations.jl:57-110)."""

    def __init__(self, transformations: Sequence[_UnivariateTransformation]):
        self.transformations = tuple(transformations)
        for t in self.transformations:
            if not isinstance(t, _UnivariateTransformation):
                raise TypeError(f"not a univariate transformation: {t!r}")

    def __len__(self):
        return len(self.transformations)

    def __iter__(self):
        return iter(self.transformations)

    def __eq__(self, other):
        return isinstance(other, CoordinateTransformations) and self.transformations == other.transformations

    def __hash__(self):
        return hash(self.transformations)

    def __repr__(self):
        return "coordinate transformations" + "".join(f"\n  {t!r}" for t in self.transformations)


def coordinate_transformations(*ts):
    if len(ts) == 1 and isinstance(ts[0], (tuple, list)):
        ts = tuple(ts[0])
    return CoordinateTransformations(ts)


def _domain_of_transformation(t):
    if isinstance(t, BoundedLinear):
        return UnivariateDomain(t.m - t.s, t.m + t.s)
    if isinstance(t, SemiInfRational):
        return UnivariateDomain(t.A, math.inf) if t.L > 0 else UnivariateDomain(-math.inf, t.A)
    if isinstance(t, InfRational):
        return UnivariateDomain(-math.inf, math.inf)
    if isinstance(t, CoordinateTransformations):
        return coordinate_domains(*[_domain_of_transformation(u) for u in t])
    raise TypeError(f"not a transformation: {t!r}")


def _tr_array(ts: Sequence[_UnivariateTransformation]):
    arr = (L.Transform * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = t._t
    return arr


# ----------------------------------------------------------------------------- derivative requests (derivatives.jl)
class Derivatives:
    """`𝑑^D` (derivatives.jl:66-107): `d(x)` asks for the value and first D derivatives at x.

    `𝑑` is NFKC-normalised to `d` by Python, so `𝑑(x)`, `(𝑑**2)(x)` work verbatim.
    """

    def __init__(self, D: int = 1):
        if not isinstance(D, int) or D < 0:
            raise ArgumentError("D ≥ 0 must hold.")
        self.D = D

    def __mul__(self, other):
        return Derivatives(self.D + other.D)

    def __pow__(self, y):
        if not isinstance(y, int) or y < 0:
            raise ArgumentError("Y isa Int && Y ≥ 0 must hold.")
        return Derivatives(self.D * y)

    def __lshift__(self, n):
        """`𝑑^D << Val(n)`: order D along coordinate n (derivatives.jl:409-413)."""
        if not isinstance(n, int) or n < 1:
            raise ArgumentError("N isa Int && N ≥ 1 must hold.")
        if self.D < 1:
            raise ArgumentError("D ≥ 1 must hold.")
        return Partial(*([0] * (n - 1) + [self.D]))

    def __call__(self, x):
        return _JetPoints(self.D, x)

    def __repr__(self):
        return f"𝑑^{self.D}"


d = Derivatives(1)


@dataclass
class _JetPoints:
    D: int
    x: object


class Partial:
    """`∂(i1, i2, …)` and unions (derivatives.jl:336-421).  `|` is Julia's `∪`."""

    def __init__(self, *I, _partials=None):
        if _partials is not None:
            self.partials = tuple(_partials)
            return
        if len(I) == 1 and isinstance(I[0], tuple):
            I = I[0]
            if len(I) and I[-1] == 0:                     # only the vararg form strips trailing zeros
                raise ArgumentError("I ≡ () || last(I) ≠ 0 must hold.")
        I = [int(i) for i in I]
        if any(i < 0 for i in I):
            raise ArgumentError("all(i -> i ≥ 0, I) must hold.")
        while I and I[-1] == 0:
            I.pop()
        self.partials = (tuple(I),)

    def __or__(self, other):
        return Partial(_partials=self.partials + other.partials)

    union = __or__

    def expansion(self, dim: int):
        """(component table [ncomp][dim], per-coordinate degrees) via sk_partials_canonical."""
        need = max((len(p) for p in self.partials), default=0)
        if dim < need:
            raise ArgumentError("N ≥ _partials_minimum_input_dimension(_Ps) must hold.")
        arr = np.zeros((max(len(self.partials), 1), dim), dtype=np.int32)
        for i, p in enumerate(self.partials):
            arr[i, : len(p)] = p
        out = np.zeros((L.SK_MAX_COMP, dim), dtype=np.int32)
        nc = C.c_int32(0)
        deg = np.zeros(dim, dtype=np.int32)
        L.check(L.lib.sk_partials_canonical(dim, len(self.partials), arr.ctypes.data_as(L._ip), out.ctypes.data_as(L._ip),
                                            L.SK_MAX_COMP, C.byref(nc), deg.ctypes.data_as(L._ip)))
        return out[: nc.value].copy(), deg

    def __call__(self, x):
        return _PartialPoints(self, x)

    def __repr__(self):
        rs = ["∂(" + ", ".join(map(str, p)) + ")" for p in self.partials]
        return rs[0] if len(rs) == 1 else "union(" + ", ".join(rs) + ")"


def partial(*I):
    """`∂(I...)`."""
    return Partial(*I)


def gradient(dim: int) -> Partial:
    """`∂(1) ∪ ∂(0,1) ∪ … ∪ ∂(0,…,0,1)`: components [value, ∂₁, …, ∂_dim]."""
    out = Partial(1)
    for j in range(1, dim):
        out = out | Partial(*([0] * j + [1]))
    return out


@dataclass
class _PartialPoints:
    spec: Partial
    x: object


# ----------------------------------------------------------------------------- bases
class FunctionBasis:
    def __matmul__(self, transformation):          # Julia's `basis ∘ transformation`
        return TransformedBasis(self, transformation)


class Chebyshev(FunctionBasis):
    """First N Chebyshev polynomials of the first kind on [-1, 1] (chebyshev.jl:12-32)."""
    multivariate = False

    def __init__(self, grid_kind, N: int):
        self.grid_kind = grid_kind
        _grid_code(grid_kind)
        if not isinstance(N, (int, np.integer)):
            raise TypeError("N must be an integer")
        if N < 1:
            raise ArgumentError("N ≥ 1 must hold.")
        self.N = int(N)

    def _desc(self):
        return L.BasisDesc(L.SK_BASIS_CHEBYSHEV, self.grid_kind.code, self.N, 1, 0, 0)

    def __eq__(self, other):
        return isinstance(other, Chebyshev) and (self.grid_kind, self.N) == (other.grid_kind, other.N)

    def __hash__(self):
        return hash(("Chebyshev", self.grid_kind.code, self.N))

    def __repr__(self):
        return f"Chebyshev polynomials (1st kind), {self.grid_kind!r}, dimension: {self.N}"


class SmolyakParameters:
    """Σ bᵢ ≤ B, all bᵢ ≤ M (smolyak_traversal.jl:11-37); M > B is clamped with a warning."""

    def __init__(self, B: int, M: Optional[int] = None):
        M = B if M is None else M
        if not isinstance(B, (int, np.integer)) or B < 0:
            raise ArgumentError("B isa Int && B ≥ 0 must hold.")
        if not isinstance(M, (int, np.integer)) or M < 0:
            raise ArgumentError("M isa Int && M ≥ 0 must hold.")
        if M > B:
            warnings.warn("M > B replaced with M = B")
        self.B, self.M = int(B), int(min(B, M))

    def __repr__(self):
        return f"Smolyak parameters, ∑bᵢ ≤ {self.B}, all bᵢ ≤ {self.M}"


class SmolyakBasis(FunctionBasis):
    multivariate = True

    def __init__(self, grid_kind, params: SmolyakParameters, N: int):
        if N < 1:
            raise ArgumentError("N ≥ 1 must hold.")
        self.grid_kind, self.params, self.dim = grid_kind, params, int(N)
        H = C.c_int64(0)
        L.check(L.lib.sk_parent_dimension(C.byref(self._desc()), C.byref(H)))
        self.univariate_parent = Chebyshev(grid_kind, H.value)

    def _desc(self):
        return L.BasisDesc(L.SK_BASIS_SMOLYAK, _grid_code(self.grid_kind), 0, self.dim, self.params.B, self.params.M)

    def __len__(self):
        return self.dim

    def __getitem__(self, i):
        if not 1 <= i <= self.dim:                    # 1-based like the reference (smolyak_api.jl:20-23)
            raise IndexError(i)
        return self.univariate_parent

    def __eq__(self, other):
        return isinstance(other, SmolyakBasis) and (self.grid_kind, self.params.B, self.params.M, self.dim) == \
            (other.grid_kind, other.params.B, other.params.M, other.dim)

    def __hash__(self):
        return hash(("Smolyak", self.grid_kind.code, self.params.B, self.params.M, self.dim))

    def __repr__(self):
        return (f"Sparse multivariate basis on ℝ^{self.dim}\n  Smolyak indexing, ∑bᵢ ≤ {self.params.B}, all bᵢ ≤ "
                f"{self.params.M}, dimension {dimension(self)}\n  using {self.univariate_parent!r}")


def smolyak_basis(univariate_family, grid_kind, smolyak_parameters: SmolyakParameters, N: int) -> SmolyakBasis:
    """smolyak_api.jl:63-75.  Only `Chebyshev` exists as a univariate family."""
    if univariate_family is not Chebyshev:
        raise TypeError("univariate_family must be Chebyshev")
    _grid_code(grid_kind)
    return SmolyakBasis(grid_kind, smolyak_parameters, N)


class TransformedBasis(FunctionBasis):
    """`parent ∘ transformation` (generic_api.jl:263-274)."""

    def __init__(self, parent: FunctionBasis, transformation):
        if isinstance(parent, TransformedBasis):
            raise ArgumentError("domain_kind(domain(parent)) ≡ domain_kind(T) must hold.")
        multi = isinstance(transformation, CoordinateTransformations)
        if multi != parent.multivariate:
            raise ArgumentError("domain_kind(domain(parent)) ≡ domain_kind(T) must hold.")
        if multi and len(transformation) != parent.dim:
            raise ArgumentError("length(domains) == length(transformations) == length(x) must hold.")
        self.parent, self.transformation = parent, transformation

    multivariate = property(lambda self: self.parent.multivariate)

    def __len__(self):
        return len(self.parent)

    def __getitem__(self, i):
        return TransformedBasis(self.parent[i], self.transformation.transformations[i - 1])

    def __eq__(self, other):
        return isinstance(other, TransformedBasis) and (self.parent, self.transformation) == (other.parent, other.transformation)

    def __hash__(self):
        return hash((self.parent, self.transformation))


def is_function_basis(b) -> bool:
    return isinstance(b, FunctionBasis) or (isinstance(b, type) and issubclass(b, FunctionBasis))


def parent(basis):
    return basis.parent


def transformation(basis):
    return basis.transformation if isinstance(basis, TransformedBasis) else None


def _split(basis):
    """-> (untransformed basis, list of univariate transformations or None)"""
    if isinstance(basis, TransformedBasis):
        t = basis.transformation
        return basis.parent, (list(t) if isinstance(t, CoordinateTransformations) else [t])
    return basis, None


def dimension(basis) -> int:
    base, _ = _split(basis)
    K = C.c_int64(0)
    L.check(L.lib.sk_dimension(C.byref(base._desc()), C.byref(K)))
    return K.value


def domain(obj):
    if isinstance(obj, TransformedBasis):
        return _domain_of_transformation(obj.transformation)
    if isinstance(obj, Chebyshev):
        return PM1
    if isinstance(obj, SmolyakBasis):
        return coordinate_domains(*([PM1] * obj.dim))
    return _domain_of_transformation(obj)


def grid(basis, dtype=np.float64) -> np.ndarray:
    """collect(grid(basis)): (N,) or (K, d) (chebyshev.jl:88-127, smolyak_api.jl:121-135)."""
    base, trs = _split(basis)
    K = dimension(basis)
    out = np.zeros(K if not base.multivariate else (K, base.dim), dtype=np.float64)
    tr = _tr_array(trs) if trs is not None else None
    L.check(L.lib.sk_grid(C.byref(base._desc()), tr, out.ctypes.data_as(L._dp)))
    return out.astype(dtype, copy=False)


def smolyak_indices(basis) -> np.ndarray:
    """collect(basis.smolyak_indices): (K, d) 1-based tuples in traversal order."""
    base, _ = _split(basis)
    out = np.zeros((dimension(base), base.dim), dtype=np.int32)
    L.check(L.lib.sk_smolyak_indices(C.byref(base._desc()), out.ctypes.data_as(L._ip)))
    return out


def nesting_total_length(grid_kind, b: int) -> int:
    o = C.c_int64(0)
    L.check(L.lib.sk_nesting_total_length(_grid_code(grid_kind), b, C.byref(o)))
    return o.value


def nesting_block_length(grid_kind, b: int) -> int:
    o = C.c_int64(0)
    L.check(L.lib.sk_nesting_block_length(_grid_code(grid_kind), b, C.byref(o)))
    return o.value


def grid_shuffle(grid_kind, length: int) -> np.ndarray:
    out = np.zeros(max(length, 1), dtype=np.int64)
    L.check(L.lib.sk_grid_shuffle(_grid_code(grid_kind), length, out.ctypes.data_as(L._lp)))
    return out


# ----------------------------------------------------------------------------- plans
class _Plan:
    def __init__(self, handle):
        self.handle = handle
        info = L.PlanInfo()
        L.check(L.lib.sk_plan_get_info(handle, C.byref(info)))
        self.info = info

    def __del__(self):
        try:
            L.lib.sk_plan_destroy(self.handle)
        except Exception:
            pass


_plan_cache = {}


def get_plan(basis, order: int = 0, spec: Optional[Partial] = None, device: int = 0, allow_rational_d2: bool = False) -> _Plan:
    """Plans are cached per (basis, derivative request, device)."""
    base, trs = _split(basis)
    key = (base, tuple(trs) if trs else None, order, spec.partials if spec else None, device, bool(allow_rational_d2))
    p = _plan_cache.get(key)
    if p is not None:
        return p
    ds = L.DerivSpec(order, 0, None, int(allow_rational_d2))
    keep = None
    if spec is not None:
        if not base.multivariate:
            raise UnsupportedError("∂ specifications need a multivariate basis")
        table, _ = spec.expansion(base.dim)
        keep = np.ascontiguousarray(table, dtype=np.int32)
        ds = L.DerivSpec(0, len(keep), keep.ctypes.data_as(L._ip), int(allow_rational_d2))
    h = C.c_void_p()
    tr = _tr_array(trs) if trs is not None else None
    L.check(L.lib.sk_plan_create(C.byref(base._desc()), tr, C.byref(ds), device, C.byref(h)))
    p = _Plan(h)
    _plan_cache[key] = p
    return p


# ----------------------------------------------------------------------------- buffers
def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _unwrap_points(basis, x):
    """-> (raw points, order, spec)"""
    if isinstance(x, _JetPoints):
        if _split(basis)[0].multivariate:
            raise UnsupportedError("use ∂ specifications (partial/gradient) with multivariate bases")
        return x.x, x.D, None
    if isinstance(x, _PartialPoints):
        return x.x, 0, x.spec
    return x, 0, None


def _points(base, x):
    """-> (array (n,) or (n, d) contiguous float64, single-point flag, device index or None)"""
    dim = base.dim if base.multivariate else 1
    if _is_torch(x) and x.is_cuda:
        single = x.dim() == (1 if base.multivariate else 0)
        t = x.reshape(1, -1) if (single and base.multivariate) else (x.reshape(1) if single else x)
        if t.dtype != torch.float64:
            raise TypeError("device points must be float64")
        if base.multivariate and (t.dim() != 2 or t.shape[1] != dim):
            raise ArgumentError("length(x) == N must hold.")
        if not base.multivariate and t.dim() != 1:
            raise ArgumentError("univariate points must be a vector")
        return t.contiguous(), single, t.device.index
    a = np.asarray(x.cpu() if _is_torch(x) else x, dtype=np.float64)
    if base.multivariate:
        single = a.ndim == 1
        a = a.reshape(1, -1) if single else a
        if a.ndim != 2 or a.shape[1] != dim:
            raise ArgumentError("length(x) == N must hold.")                     # smolyak_api.jl:97
    else:
        single = a.ndim == 0
        a = a.reshape(1) if single else a
        if a.ndim != 1:
            raise ArgumentError("univariate points must be a scalar or a vector")
    return np.ascontiguousarray(a), single, None


def _ptr(a):
    return C.c_void_p(a.data_ptr()) if _is_torch(a) else C.c_void_p(a.ctypes.data)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if dev is not None else None


# ----------------------------------------------------------------------------- the hot path
class LinearCombination:
    """`linear_combination(basis, θ)` callable (generic_api.jl:141-160)."""

    def __init__(self, basis, theta):
        self.basis, self.theta = basis, theta
        n = theta.shape[0] if hasattr(theta, "shape") else len(theta)
        if dimension(basis) != n:
            raise ArgumentError("dimension(basis) == length(θ) must hold.")

    def __call__(self, x, **kw):
        return linear_combination(self.basis, self.theta, x, **kw)

    def __matmul__(self, transformation):          # `linear_combination(basis, θ) ∘ t` (generic_api.jl:302-305)
        return LinearCombination(self.basis @ transformation, self.theta)


def linear_combination(basis, theta, x=None, *, exact: bool = False, device: int = 0, allow_rational_d2: bool = False):
    """Σ θₖ·fₖ(x) (generic_api.jl:114-138) for a batch of points, on the GPU.

    θ: (K,) or (K, M) (M coefficient vectors = `Vector{SVector{M}}`).  Returns (n,) for plain
    values, (n, D+1) / (n, ncomp) with derivatives, (n, M) for SVector coefficients; a single
    point drops the leading axis.  `exact=True` selects the reference operation order
    (bit-identical to the oracle); the default is the fused/nested kernel.
    """
    if x is None:
        return LinearCombination(basis, theta)
    raw, order, spec = _unwrap_points(basis, x)
    base, _ = _split(basis)
    pts, single, dev = _points(base, raw)
    plan = get_plan(basis, order, spec, dev if dev is not None else device, allow_rational_d2)
    on_device = dev is not None
    if on_device:
        th = theta if _is_torch(theta) else torch.as_tensor(np.asarray(theta, dtype=np.float64))
        th = th.to(device=pts.device, dtype=torch.float64).contiguous()
    else:
        th = np.ascontiguousarray(theta.cpu().numpy() if _is_torch(theta) else theta, dtype=np.float64)
    if th.ndim not in (1, 2):
        raise ArgumentError("θ must be a vector of reals or of fixed-length vectors")
    nvec = 1 if th.ndim == 1 else int(th.shape[1])
    n = int(pts.shape[0])
    E = nvec if th.ndim == 2 else plan.info.E
    out = torch.empty((n, E), dtype=torch.float64, device=pts.device) if on_device else np.empty((n, E), dtype=np.float64)
    flags = (L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST) | (L.SK_MODE_EXACT if exact else L.SK_MODE_FAST)
    L.check(L.lib.sk_linear_combination(plan.handle, _ptr(th), int(th.shape[0]), nvec, _ptr(pts), n, _ptr(out), flags, _stream(dev)))
    if th.ndim == 1 and order == 0 and spec is None:
        out = out[:, 0]
    return out[0] if single else out


def stacked_linear_combination(basis, coefficients, x, **kw):
    """Several functions over one basis from ONE stacked coefficient vector, the layout of the reference's
    `Experimental.make_policy_functions` (src/experimental.jl:156-166: function p owns `coefficients[(p-1)·K+1 : p·K]`,
    K = dimension(basis)).  Returns (n, P): column p is `linear_combination(basis, coefficients[p·K:(p+1)·K], x)`,
    evaluated in one pass as a K×P coefficient block."""
    K = dimension(basis)
    th = coefficients
    total = int(th.shape[0]) if hasattr(th, "shape") else len(th)
    if total == 0 or total % K != 0:
        raise ArgumentError("length(coefficients) must be a positive multiple of dimension(basis)")
    P = total // K
    if _is_torch(th):
        block = th.reshape(P, K).t().contiguous()
    else:
        block = np.ascontiguousarray(np.asarray(th, dtype=np.float64).reshape(P, K).T)
    return linear_combination(basis, block, x, **kw)


def basis_at(basis, x, *, device: int = 0, allow_rational_d2: bool = False):
    """Basis functions at the points: array (n, K[, E]) — row i is `collect(basis_at(basis, x[i]))`.

    Memory is the reference's column-major `Matrix` (generic_api.jl:217-227): the returned array is
    a transposed view of the (K, n, E) buffer the kernel writes.
    """
    raw, order, spec = _unwrap_points(basis, x)
    base, _ = _split(basis)
    pts, single, dev = _points(base, raw)
    plan = get_plan(basis, order, spec, dev if dev is not None else device, allow_rational_d2)
    n, K, E = int(pts.shape[0]), plan.info.K, plan.info.E
    on_device = dev is not None
    out = torch.empty((K, n, E), dtype=torch.float64, device=pts.device) if on_device else np.empty((K, n, E), dtype=np.float64)
    flags = L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST
    L.check(L.lib.sk_basis_at(plan.handle, _ptr(pts), n, _ptr(out), n, flags, _stream(dev)))
    res = out.permute(1, 0, 2) if on_device else out.transpose(1, 0, 2)
    if order == 0 and spec is None:
        res = res[:, :, 0]
    return res[0] if single else res


def collocation_matrix(basis, x=None, **kw):
    """`collocation_matrix(basis, x = grid(basis))` (generic_api.jl:217-227): (M, N[, E])."""
    if x is None:
        x = grid(basis)
    return basis_at(basis, x, **kw)


def _desc_and_tr(basis):
    base, trs = _split(basis)
    return base._desc(), (_tr_array(trs) if trs is not None else None)


def is_subset_basis(basis1, basis2) -> bool:
    """Can coefficients of `basis1` be augmented to `basis2`? (generic_api.jl:247-254, chebyshev.jl:140-142,
    smolyak_api.jl:141-153, generic_api.jl:314-317)"""
    for b in (basis1, basis2):
        if not isinstance(_split(b)[0], (Chebyshev, SmolyakBasis)):
            return False
    d1, t1 = _desc_and_tr(basis1)
    d2, t2 = _desc_and_tr(basis2)
    if (t1 is None) != (t2 is None) or (t1 is not None and len(t1) != len(t2)):
        return False
    out = C.c_int(0)
    L.check(L.lib.sk_is_subset_basis(C.byref(d1), t1, C.byref(d2), t2, C.byref(out)))
    return bool(out.value)


def augment_coefficients(basis1, basis2, theta1):
    """Coefficients θ2 of `basis2` with `linear_combination(basis1, θ1, x) == linear_combination(basis2, θ2, x)`
    (generic_api.jl:229-242; chebyshev.jl:133-138; smolyak_api.jl:155-208).  θ1: (K1,) or (K1, M)."""
    if not is_subset_basis(basis1, basis2):
        raise ArgumentError("is_subset_basis(basis1, basis2) must hold.")
    th = np.ascontiguousarray(theta1.cpu().numpy() if _is_torch(theta1) else theta1, dtype=np.float64)
    if th.ndim not in (1, 2):
        raise ArgumentError("θ must be a vector of reals or of fixed-length vectors")
    nvec = 1 if th.ndim == 1 else int(th.shape[1])
    d1, t1 = _desc_and_tr(basis1)
    d2, t2 = _desc_and_tr(basis2)
    out = np.empty((dimension(basis2),) + th.shape[1:], dtype=np.float64)
    L.check(L.lib.sk_augment_coefficients(C.byref(d1), t1, C.byref(d2), t2, _ptr(th), int(th.shape[0]), nvec, _ptr(out)))
    return out


def collocation_solve(basis, y, *, device: int = 0):
    """θ = `collocation_matrix(basis) \\ y` on the basis' own grid (docs/src/index.md:92,
    test/test_smolyak.jl:20-21): y holds the function values at `grid(basis)`, shape (K,) or (K, M) for M
    functions at once; returns θ of the same shape.  The matrix is generated and factorised on the GPU."""
    on_device = _is_torch(y) and y.is_cuda
    dev = y.device.index if on_device else None
    plan = get_plan(basis, 0, None, dev if dev is not None else device)
    if on_device:
        yy = y.to(dtype=torch.float64).contiguous()
    else:
        yy = np.ascontiguousarray(y.cpu().numpy() if _is_torch(y) else y, dtype=np.float64)
    if yy.ndim not in (1, 2) or int(yy.shape[0]) != plan.info.K:
        raise ArgumentError("y must hold one value (or one fixed-length vector) per grid point of the basis")
    nrhs = 1 if yy.ndim == 1 else int(yy.shape[1])
    out = torch.empty_like(yy) if on_device else np.empty_like(yy)
    flags = L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST
    L.check(L.lib.sk_collocation_solve(plan.handle, _ptr(yy), nrhs, _ptr(out), flags, _stream(dev)))
    return out


def _transform_common(to, dom_or_basis, t, x, device):
    ts = list(t) if isinstance(t, CoordinateTransformations) else [t]
    order = 0
    if isinstance(x, _JetPoints):
        if not to:
            raise UnsupportedError("transform_from of derivatives")
        x, order = x.x, x.D
    multi = isinstance(t, CoordinateTransformations)
    on_device = _is_torch(x) and x.is_cuda
    a = x if on_device else np.asarray(x, dtype=np.float64)
    single = (a.ndim == 1) if multi else (a.ndim == 0)
    a = a.reshape(1, -1) if (single and multi) else (a.reshape(1) if single else a)
    if multi and a.shape[-1] != len(ts):
        raise ArgumentError("length(domains) == length(transformations) == length(x) must hold.")
    a = a.contiguous() if on_device else np.ascontiguousarray(a)
    n = int(a.shape[0])
    shape = tuple(a.shape) + ((order + 1,) if (to and order > 0) else ())
    dev = a.device.index if on_device else None
    out = torch.empty(shape, dtype=torch.float64, device=a.device) if on_device else np.empty(shape, dtype=np.float64)
    flags = L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST
    d_ = len(ts)
    if to:
        L.check(L.lib.sk_transform_to(_tr_array(ts), d_, order, 0, _ptr(a), n, _ptr(out), flags, dev if on_device else device, _stream(dev)))
    else:
        L.check(L.lib.sk_transform_from(_tr_array(ts), d_, _ptr(a), n, _ptr(out), flags, dev if on_device else device, _stream(dev)))
    return out[0] if single else out


def transform_to(dom_or_basis, t, x, *, device: int = 0):
    """`transform_to(domain, t, x)` for batches (transformations.jl:112-127,183-195,246-267,311-333)."""
    return _transform_common(True, dom_or_basis, t, x, device)


def transform_from(dom_or_basis, t, x, *, device: int = 0):
    """`transform_from(domain, t, x)` for batches (transformations.jl:129-139,178-181,244,309)."""
    return _transform_common(False, dom_or_basis, t, x, device)


def fp64_peak(device: int = 0, which: int = 0, seconds: float = 0.5):
    """(burst, sustained) TFLOP/s of the FP64 pipe: which = 0 DFMA, 1 DMMA, 2 both at once, 3 DFMA with three
    distinct register operands (register-file bound)."""
    b, s = C.c_double(0), C.c_double(0)
    L.check(L.lib.sk_fp64_peak(device, which, seconds, C.byref(b), C.byref(s)))
    return b.value, s.value


def launch_count() -> int:
    return int(L.lib.sk_launch_count())


def linear_combination_sharded(basis, theta, x, devices: Sequence[int], *, order_or_spec=None, exact=False):
    """One process, several GPUs: contiguous point slices per device. Host arrays."""
    raw, order, spec = _unwrap_points(basis, x)
    base, _ = _split(basis)
    pts, single, dev = _points(base, raw)
    if dev is not None:
        raise ArgumentError("sharded evaluation takes host arrays")
    plans = [get_plan(basis, order, spec, g) for g in devices]
    th = np.ascontiguousarray(theta, dtype=np.float64)
    nvec = 1 if th.ndim == 1 else int(th.shape[1])
    n = int(pts.shape[0])
    E = nvec if th.ndim == 2 else plans[0].info.E
    out = np.empty((n, E), dtype=np.float64)
    handles = (C.c_void_p * len(plans))(*[p.handle for p in plans])
    flags = L.SK_MEM_HOST | (L.SK_MODE_EXACT if exact else L.SK_MODE_FAST)
    L.check(L.lib.sk_linear_combination_sharded(handles, len(plans), _ptr(th), int(th.shape[0]), nvec, _ptr(pts), n, _ptr(out), flags))
    if th.ndim == 1 and order == 0 and spec is None:
        out = out[:, 0]
    return out[0] if single else out
