"""Host-side mirror of the batched part of the reference's `Experimental` module (src/experimental.jl): the caller of
the hot path.  Policy functions are several linear combinations over ONE basis from one stacked coefficient vector
(`make_policy_functions`, :156-166), each followed by a univariate `bridge` (:242-266); the model's residual function is
user code; `sum_of_squared_residuals` (:228-233) reduces it over the grid.  Everything except the user's residual
function runs in the library (C ABI: sk_policy_functions, sk_sum_abs2, sk_policy_residual_ssq).
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional

import numpy as np

from . import _lib as L
from . import api

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class Bridge:
    """`bridge(outer_transformation, inner_transformation)`: x ↦ transform_from(PM1, outer, transform_to(PM1, inner, x))
    (src/experimental.jl:242-266); `inverse()` swaps the two (InverseFunctions.inverse)."""

    def __init__(self, outer, inner):
        for t in (outer, inner):
            if not isinstance(t, api._UnivariateTransformation):
                raise api.ArgumentError("bridge takes univariate transformations")
        self.outer, self.inner = outer, inner

    def inverse(self):
        return Bridge(self.inner, self.outer)

    def _c(self):
        return L.Bridge(self.outer._t, self.inner._t)


def bridge(outer_transformation, inner_transformation) -> Bridge:
    return Bridge(outer_transformation, inner_transformation)


def _identity():
    t = L.Transform(L.SK_TR_NONE, 0, 0.0, 1.0)
    return L.Bridge(t, t)


class PolicyFunctions:
    """What `make_policy_functions(model_family, policy_transformations, approximation_basis, coefficients)` returns in
    the reference is a NamedTuple of callables; here it is one object evaluating ALL of them for a batch of points in
    one pass: `pf(x)` → (n, M) in the order of `policy_transformations`, `pf.named(x)` → dict of columns."""

    def __init__(self, policy_transformations: Mapping[str, Optional[Bridge]], approximation_basis, coefficients, device: int = 0):
        self.names = list(policy_transformations.keys())
        self.basis = approximation_basis
        K = api.dimension(approximation_basis)
        M = len(self.names)
        total = int(coefficients.shape[0]) if hasattr(coefficients, "shape") else len(coefficients)
        if M < 1 or total != K * M:
            raise api.ArgumentError("lastindex(coefficients) == last(last(ranges)) must hold.")      # experimental.jl:162
        self.coefficients = coefficients
        self.M, self.K, self.device = M, K, device
        self._bridges = (L.Bridge * M)(*[(b._c() if b is not None else _identity()) for b in policy_transformations.values()])
        self._any = any(b is not None for b in policy_transformations.values())

    def _prep(self, x):
        base, _ = api._split(self.basis)
        pts, single, dev = api._points(base, x)
        plan = api.get_plan(self.basis, 0, None, dev if dev is not None else self.device)
        on_device = dev is not None
        co = self.coefficients
        if on_device:
            co = torch.as_tensor(co, dtype=torch.float64, device=pts.device).contiguous()
        else:
            co = np.ascontiguousarray(co.cpu().numpy() if api._is_torch(co) else co, dtype=np.float64)
        return pts, single, dev, plan, on_device, co

    def __call__(self, x):
        pts, single, dev, plan, on_device, co = self._prep(x)
        n = int(pts.shape[0])
        out = torch.empty((n, self.M), dtype=torch.float64, device=pts.device) if on_device else np.empty((n, self.M))
        flags = (L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST) | L.SK_MODE_FAST
        L.check(L.lib.sk_policy_functions(plan.handle, api._ptr(co), self.K * self.M, self.M, C.cast(self._bridges, C.c_void_p) if self._any else None,
                                          api._ptr(pts), n, api._ptr(out), flags, api._stream(dev)))
        return out[0] if single else out

    def named(self, x):
        v = self(x)
        return {name: v[..., i] for i, name in enumerate(self.names)}

    def residual_ssq(self, x, target) -> float:
        """Σ (policy(x) − target)² fused on the device (sk_policy_residual_ssq)."""
        pts, single, dev, plan, on_device, co = self._prep(x)
        n = int(pts.shape[0])
        tg = torch.as_tensor(target, dtype=torch.float64, device=pts.device).contiguous() if on_device else np.ascontiguousarray(target, dtype=np.float64)
        if int(np.prod(tg.shape)) != n * self.M:
            raise api.ArgumentError("target must have one value per point and policy function")
        res = C.c_double(0)
        flags = (L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST) | L.SK_MODE_FAST
        L.check(L.lib.sk_policy_residual_ssq(plan.handle, api._ptr(co), self.K * self.M, self.M, C.cast(self._bridges, C.c_void_p) if self._any else None,
                                             api._ptr(pts), n, api._ptr(tg), C.byref(res), flags, api._stream(dev)))
        return res.value


def make_policy_functions(policy_transformations: Mapping[str, Optional[Bridge]], approximation_basis, coefficients, device: int = 0) -> PolicyFunctions:
    return PolicyFunctions(policy_transformations, approximation_basis, coefficients, device)


def policy_coefficients_dimension(policy_transformations, approximation_basis) -> int:
    """dimension(approximation_basis) · length(policy_transformations) (src/experimental.jl:152-154)"""
    return api.dimension(approximation_basis) * len(policy_transformations)


def sum_abs2(residuals, device: int = 0) -> float:
    """Σ r² over an array of residuals (`sum_abs2` mapreduced over the grid, src/experimental.jl:218-233); deterministic."""
    on_device = api._is_torch(residuals) and residuals.is_cuda
    r = residuals.contiguous() if on_device else np.ascontiguousarray(residuals.cpu().numpy() if api._is_torch(residuals) else residuals, dtype=np.float64)
    count = int(r.numel()) if on_device else int(r.size)
    res = C.c_double(0)
    dev = r.device.index if on_device else device
    L.check(L.lib.sk_sum_abs2(api._ptr(r), count, C.byref(res), L.SK_MEM_DEVICE if on_device else L.SK_MEM_HOST, dev, api._stream(dev if on_device else None)))
    return res.value


def sum_of_squared_residuals(calculate_residuals, policy_functions: PolicyFunctions, grid) -> float:
    """`sum_of_squared_residuals(model_family, model_parameters, policy_functions, grid)`: the policy functions are
    evaluated for the whole grid in one pass, `calculate_residuals(named policy values, grid)` (user code, vectorised
    over the grid) returns an array or a dict of arrays of residuals, and their squares are summed on the device."""
    vals = policy_functions.named(grid)
    r = calculate_residuals(vals, grid)
    if isinstance(r, Mapping):
        parts = list(r.values())
        r = torch.stack([torch.as_tensor(p) for p in parts]) if any(api._is_torch(p) for p in parts) else np.stack([np.asarray(p) for p in parts])
    return sum_abs2(r, policy_functions.device)
