"""ctypes front-end of the CPU oracle (oracle/sk_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Every function mirrors one C entry point; see sk_oracle.c for the reference file:line each
follows.  Arrays are numpy float64/int32/int64, C-contiguous.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsk_oracle.so")

OK, ERR_ARGUMENT, ERR_DOMAIN, ERR_UNSUPPORTED = 0, -1, -2, -3
TR_NONE, TR_BOUNDED_LINEAR, TR_SEMI_INF, TR_INF = 0, 1, 2, 3
GRID_INTERIOR, GRID_ENDPOINT, GRID_INTERIOR2 = 0, 1, 2


class OracleError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"oracle {what}: status {code}")
        self.code = code


class Transform(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("p0", C.c_double), ("p1", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile if the .so is missing (or stale)."""
    src = os.path.join(_HERE, "sk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.orc_hypot.restype = C.c_double
        L.orc_hypot.argtypes = [C.c_double, C.c_double]
        L.orc_transform_make.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(Transform)]
        L.orc_transform_from.restype = C.c_double
        L.orc_transform_from.argtypes = [C.POINTER(Transform), C.c_double]
        L.orc_transform_to_jet.argtypes = [C.POINTER(Transform), dp, C.c_int, C.c_int, dp]
        L.orc_cheb_basis_at.argtypes = [C.c_int, C.c_int, dp, dp]
        L.orc_gridpoint.restype = C.c_double
        L.orc_gridpoint.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_nesting_total_length.restype = C.c_int64
        L.orc_nesting_total_length.argtypes = [C.c_int, C.c_int]
        L.orc_nesting_block_length.restype = C.c_int64
        L.orc_nesting_block_length.argtypes = [C.c_int, C.c_int]
        L.orc_grid_shuffle.restype = C.c_int64
        L.orc_grid_shuffle.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_smolyak_length.restype = C.c_int64
        L.orc_smolyak_length.argtypes = [C.c_int] * 4
        L.orc_smolyak_parent_dimension.restype = C.c_int64
        L.orc_smolyak_parent_dimension.argtypes = [C.c_int] * 3
        L.orc_smolyak_indices.restype = C.c_int64
        L.orc_smolyak_indices.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int64), C.c_int64]
        L.orc_smolyak_grid.restype = C.c_int64
        L.orc_smolyak_grid.argtypes = [C.c_int] * 4 + [C.POINTER(Transform), dp, C.c_int64]
        L.orc_cheb_grid.argtypes = [C.c_int, C.c_int, C.POINTER(Transform), dp]
        ip = C.POINTER(C.c_int32)
        L.orc_partials_minimal.argtypes = [C.c_int, C.c_int, ip, ip]
        L.orc_partials_canonical.argtypes = [C.c_int, C.c_int, ip, ip, C.c_int]
        L.orc_partials_degrees.restype = None
        L.orc_partials_degrees.argtypes = [C.c_int, C.c_int, ip, ip]
        L.orc_uni_lincomb.argtypes = [C.c_int, C.POINTER(Transform), C.c_int, C.c_int, dp, C.c_int,
                                      dp, C.c_int64, dp, C.c_int, C.c_int]
        L.orc_uni_basis_at.argtypes = [C.c_int, C.POINTER(Transform), C.c_int, C.c_int, dp, C.c_int64, dp]
        L.orc_smolyak_lincomb.argtypes = [C.c_int] * 4 + [C.POINTER(Transform), C.c_int, ip, C.c_int,
                                                          dp, C.c_int, dp, C.c_int64, dp, C.c_int, C.c_int]
        L.orc_smolyak_basis_at.argtypes = [C.c_int] * 4 + [C.POINTER(Transform), C.c_int, ip, C.c_int,
                                                           dp, C.c_int64, dp]
        L.orc_transform_to_batch.argtypes = [C.POINTER(Transform), C.c_int, C.c_int, C.c_int, dp, C.c_int64, dp]
        L.orc_transform_from_batch.argtypes = [C.POINTER(Transform), C.c_int, dp, C.c_int64, dp]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _check(rc, what):
    if rc != OK:
        raise OracleError(rc, what)


# ----------------------------------------------------------------------------- transformations
def make_transform(kind, a=0.0, b=1.0) -> Transform:
    t = Transform()
    _check(lib().orc_transform_make(kind, float(a), float(b), C.byref(t)), "transform_make")
    return t


def BoundedLinear(a, b):
    return make_transform(TR_BOUNDED_LINEAR, a, b)


def SemiInfRational(A, L):
    return make_transform(TR_SEMI_INF, A, L)


def InfRational(A, L):
    return make_transform(TR_INF, A, L)


def Identity():
    return make_transform(TR_NONE)


def _tr_array(trs):
    if trs is None:
        return None
    arr = (Transform * len(trs))()
    for i, t in enumerate(trs):
        arr[i].kind, arr[i].pad, arr[i].p0, arr[i].p1 = t.kind, 0, t.p0, t.p1
    return arr


def hypot(x, y):
    return lib().orc_hypot(float(x), float(y))


def transform_from(t, x):
    return lib().orc_transform_from(C.byref(t), float(x))


def transform_to(t, y, D=0, ext2=False):
    """transform_to of the jet (y, 1, 0, …) of order D (or of an explicit jet if `y` is a sequence)."""
    yj = np.zeros(D + 1) if np.isscalar(y) else np.asarray(y, dtype=np.float64).copy()
    if np.isscalar(y):
        yj[0] = y
        if D >= 1:
            yj[1] = 1.0
    out = np.zeros(len(yj))
    _check(lib().orc_transform_to_jet(C.byref(t), _dp(yj), len(yj), int(ext2), _dp(out)), "transform_to")
    return out


def transform_to_batch(trs, y, D=0, ext2=False):
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, d = y.shape
    out = np.zeros((n, d, D + 1))
    _check(lib().orc_transform_to_batch(_tr_array(trs), d, D + 1, int(ext2), _dp(y), n, _dp(out)), "transform_to_batch")
    return out


def transform_from_batch(trs, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, d = x.shape
    out = np.zeros((n, d))
    _check(lib().orc_transform_from_batch(_tr_array(trs), d, _dp(x), n, _dp(out)), "transform_from_batch")
    return out


# ----------------------------------------------------------------------------- Chebyshev / grids
def cheb_basis_at(N, xjet):
    xjet = np.atleast_1d(np.asarray(xjet, dtype=np.float64)).copy()
    out = np.zeros((N, len(xjet)))
    _check(lib().orc_cheb_basis_at(N, len(xjet), _dp(xjet), _dp(out)), "cheb_basis_at")
    return out


def gridpoint(grid_kind, N, i):
    return lib().orc_gridpoint(grid_kind, N, i)


def cheb_grid(grid_kind, N, tr=None):
    out = np.zeros(N)
    _check(lib().orc_cheb_grid(grid_kind, N, C.byref(tr) if tr is not None else None, _dp(out)), "cheb_grid")
    return out


def nesting_total_length(grid_kind, b):
    return lib().orc_nesting_total_length(grid_kind, b)


def nesting_block_length(grid_kind, b):
    return lib().orc_nesting_block_length(grid_kind, b)


def grid_shuffle(grid_kind, length):
    out = np.zeros(length + 2, dtype=np.int64)
    n = lib().orc_grid_shuffle(grid_kind, length, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out[:n].copy()


# ----------------------------------------------------------------------------- Smolyak
def smolyak_length(grid_kind, d, B, M=None):
    return lib().orc_smolyak_length(grid_kind, d, B, B if M is None else M)


def smolyak_parent_dimension(grid_kind, B, M=None):
    return lib().orc_smolyak_parent_dimension(grid_kind, B, B if M is None else M)


def smolyak_indices(grid_kind, d, B, M=None):
    M = B if M is None else M
    K = smolyak_length(grid_kind, d, B, M)
    out = np.zeros((K, d), dtype=np.int64)
    n = lib().orc_smolyak_indices(grid_kind, d, B, M, out.ctypes.data_as(C.POINTER(C.c_int64)), K)
    assert n == K, (n, K)
    return out


def smolyak_grid(grid_kind, d, B, M=None, trs=None):
    M = B if M is None else M
    K = smolyak_length(grid_kind, d, B, M)
    out = np.zeros((K, d))
    n = lib().orc_smolyak_grid(grid_kind, d, B, M, _tr_array(trs), _dp(out), K)
    assert n == K, (n, K)
    return out


# ----------------------------------------------------------------------------- ∂ specs
def partials_minimal(partials, d):
    """partials: list of tuples (trailing zeros allowed) -> minimal representation [nk][d]."""
    a = np.zeros((len(partials), d), dtype=np.int32)
    for i, p in enumerate(partials):
        a[i, : len(p)] = p
    out = np.zeros_like(a)
    nk = lib().orc_partials_minimal(d, len(partials), _ip(a), _ip(out))
    if nk < 0:
        raise OracleError(nk, "partials_minimal")
    return out[:nk].copy()


def partials_canonical(minimal, d):
    m = np.ascontiguousarray(minimal, dtype=np.int32).reshape(-1, d)
    out = np.zeros((64, d), dtype=np.int32)
    nc = lib().orc_partials_canonical(d, len(m), _ip(m), _ip(out), 64)
    if nc < 0:
        raise OracleError(nc, "partials_canonical")
    return out[:nc].copy()


def partials_degrees(minimal, d):
    m = np.ascontiguousarray(minimal, dtype=np.int32).reshape(-1, d)
    deg = np.zeros(d, dtype=np.int32)
    lib().orc_partials_degrees(d, len(m), _ip(m), _ip(deg))
    return deg


def deriv_spec(partials, d):
    """∂(…) ∪ ∂(…) … -> canonical component table [ncomp][d] (derivatives.jl:266-300)."""
    return partials_canonical(partials_minimal(partials, d), d)


def gradient_spec(d):
    """∂(1) ∪ ∂(0,1) ∪ … : [value, ∂1, …, ∂d]."""
    return deriv_spec([tuple([0] * j + [1]) for j in range(d)], d)


# ----------------------------------------------------------------------------- evaluation
def uni_lincomb(N, theta, y, tr=None, D=0, ext2=False, absmode=False, nthreads=0):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    nvec = 1 if theta.ndim == 1 else theta.shape[1]
    if theta.shape[0] != N:
        raise OracleError(ERR_ARGUMENT, "uni_lincomb: coefficient length")  # generic_api.jl:99-102
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    out = np.zeros((n, nvec if nvec > 1 else D + 1))
    _check(lib().orc_uni_lincomb(N, C.byref(tr) if tr is not None else None, D + 1, int(ext2), _dp(theta), nvec,
                                 _dp(y), n, _dp(out), int(absmode), nthreads), "uni_lincomb")
    return out


def uni_basis_at(N, y, tr=None, D=0, ext2=False):
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    out = np.zeros((N, n, D + 1))          # column-major n×N of (D+1)-element entries
    _check(lib().orc_uni_basis_at(N, C.byref(tr) if tr is not None else None, D + 1, int(ext2), _dp(y), n, _dp(out)),
           "uni_basis_at")
    return out


def smolyak_lincomb(grid_kind, d, B, M, theta, y, trs=None, spec=None, ext2=False, absmode=False, nthreads=0):
    M = B if M is None else M
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    nvec = 1 if theta.ndim == 1 else theta.shape[1]
    K = smolyak_length(grid_kind, d, B, M)
    if theta.shape[0] != K:
        raise OracleError(ERR_ARGUMENT, "smolyak_lincomb: coefficient length")
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    assert y.shape[1] == d
    if spec is None:
        ncomp, sp = 0, None
    else:
        sp = np.ascontiguousarray(spec, dtype=np.int32).reshape(-1, d)
        ncomp = len(sp)
    out = np.zeros((n, ncomp if ncomp else nvec))
    _check(lib().orc_smolyak_lincomb(grid_kind, d, B, M, _tr_array(trs), ncomp, _ip(sp) if sp is not None else None,
                                     int(ext2), _dp(theta), nvec, _dp(y), n, _dp(out), int(absmode), nthreads),
           "smolyak_lincomb")
    return out


def smolyak_basis_at(grid_kind, d, B, M, y, trs=None, spec=None, ext2=False):
    M = B if M is None else M
    K = smolyak_length(grid_kind, d, B, M)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    if spec is None:
        ncomp, sp, E = 0, None, 1
    else:
        sp = np.ascontiguousarray(spec, dtype=np.int32).reshape(-1, d)
        ncomp = E = len(sp)
    out = np.zeros((K, n, E))              # column-major n×K of E-element entries
    _check(lib().orc_smolyak_basis_at(grid_kind, d, B, M, _tr_array(trs), ncomp, _ip(sp) if sp is not None else None,
                                      int(ext2), _dp(y), n, _dp(out)), "smolyak_basis_at")
    return out


def max_threads():
    """Host threads the oracle may use: the CPUs this process may run on (torchrun exports OMP_NUM_THREADS=1,
    which would otherwise make the timed CPU baseline single-threaded), never below OpenMP's own default."""
    try:
        avail = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        avail = os.cpu_count() or 1
    return max(int(lib().orc_max_threads()), avail)
