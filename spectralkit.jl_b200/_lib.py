"""ctypes binding of libspectralkit_b200.so (include/spectralkit_b200.h).

There is no fallback of any kind: if the shared library is missing this module raises at import,
and every evaluation call fails with the library's status code when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SK_B200_LIB") or os.path.join(_HERE, "lib", "libspectralkit_b200.so")   # override: tuning builds

SK_OK, SK_ERR_ARGUMENT, SK_ERR_DOMAIN, SK_ERR_UNSUPPORTED, SK_ERR_CUDA, SK_ERR_NO_DEVICE, SK_ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
SK_BASIS_CHEBYSHEV, SK_BASIS_SMOLYAK = 0, 1
SK_GRID_INTERIOR, SK_GRID_ENDPOINT, SK_GRID_INTERIOR2 = 0, 1, 2
SK_TR_NONE, SK_TR_BOUNDED_LINEAR, SK_TR_SEMI_INF_RATIONAL, SK_TR_INF_RATIONAL = 0, 1, 2, 3
SK_MEM_HOST, SK_MEM_DEVICE, SK_MODE_FAST, SK_MODE_EXACT = 0, 1, 0, 2
SK_MAX_DIM, SK_MAX_COMP, SK_MAX_ORDER = 16, 64, 5


class ArgumentError(ValueError):
    """Julia's ArgumentError (e.g. "not enough coefficients", generic_api.jl:99-102)."""


class DomainError(ValueError):
    """Julia's DomainError (transformation constructors, transformations.jl:155-158,212-213,286-287)."""


class UnsupportedError(NotImplementedError):
    """`error("… not implemented yet")` / MethodError in the reference."""


class NcclError(RuntimeError):
    """NCCL could not be loaded or a collective failed (SK_ERR_NCCL)."""


class CudaError(RuntimeError):
    """CUDA failure or no usable device: there is no CPU fallback."""


class BasisDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("grid_kind", C.c_int32), ("N", C.c_int32), ("d", C.c_int32), ("B", C.c_int32), ("M", C.c_int32)]


class Transform(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p0", C.c_double), ("p1", C.c_double)]


class DerivSpec(C.Structure):
    _fields_ = [("order", C.c_int32), ("ncomp", C.c_int32), ("orders", C.POINTER(C.c_int32)), ("allow_rational_d2", C.c_int32)]


class Bridge(C.Structure):
    _fields_ = [("outer", Transform), ("inner", Transform)]


class PlanInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("grid_kind", C.c_int32), ("d", C.c_int32), ("B", C.c_int32), ("M", C.c_int32), ("device", C.c_int32),
                ("K", C.c_int64), ("H", C.c_int64), ("E", C.c_int32), ("fast_path", C.c_int32), ("n_runs", C.c_int64),
                ("flops_per_point_fast", C.c_double), ("flops_per_point_direct", C.c_double), ("bytes_per_point", C.c_double)]


# every symbol include/spectralkit_b200.h declares (tests check that the .so exports all of them)
EXPORTS = ["sk_version", "sk_last_error", "sk_device_count", "sk_transform_make", "sk_nesting_total_length",
           "sk_nesting_block_length", "sk_grid_shuffle", "sk_dimension", "sk_parent_dimension", "sk_smolyak_indices",
           "sk_grid", "sk_partials_canonical", "sk_plan_create", "sk_plan_destroy", "sk_plan_get_info",
           "sk_linear_combination", "sk_basis_at", "sk_transform_to", "sk_transform_from",
           "sk_linear_combination_sharded", "sk_comm_create", "sk_comm_destroy", "sk_comm_size", "sk_broadcast_coefficients",
           "sk_linear_combination_sharded_device", "sk_linear_combination_peers", "sk_collocation_solve", "sk_is_subset_basis", "sk_augment_coefficients", "sk_coefficient_block_flops", "sk_policy_functions", "sk_sum_abs2", "sk_policy_residual_ssq",
           "sk_fp64_peak", "sk_lu_benchmark", "sk_launch_count"]

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                      "(make -C spectralkit.jl_b200/csrc). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)
_dp, _ip, _lp, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p
lib.sk_version.restype = C.c_char_p
lib.sk_last_error.restype = C.c_char_p
lib.sk_device_count.argtypes = [C.POINTER(C.c_int)]
lib.sk_transform_make.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(Transform)]
lib.sk_nesting_total_length.argtypes = [C.c_int, C.c_int, _lp]
lib.sk_nesting_block_length.argtypes = [C.c_int, C.c_int, _lp]
lib.sk_grid_shuffle.argtypes = [C.c_int, C.c_int64, _lp]
lib.sk_dimension.argtypes = [C.POINTER(BasisDesc), _lp]
lib.sk_parent_dimension.argtypes = [C.POINTER(BasisDesc), _lp]
lib.sk_smolyak_indices.argtypes = [C.POINTER(BasisDesc), _ip]
lib.sk_grid.argtypes = [C.POINTER(BasisDesc), C.POINTER(Transform), _dp]
lib.sk_partials_canonical.argtypes = [C.c_int, C.c_int, _ip, _ip, C.c_int, _ip, _ip]
lib.sk_plan_create.argtypes = [C.POINTER(BasisDesc), C.POINTER(Transform), C.POINTER(DerivSpec), C.c_int, C.POINTER(_vp)]
lib.sk_plan_destroy.argtypes = [_vp]
lib.sk_plan_get_info.argtypes = [_vp, C.POINTER(PlanInfo)]
lib.sk_linear_combination.argtypes = [_vp, _vp, C.c_int64, C.c_int, _vp, C.c_int64, _vp, C.c_uint32, _vp]
lib.sk_basis_at.argtypes = [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_uint32, _vp]
lib.sk_transform_to.argtypes = [C.POINTER(Transform), C.c_int, C.c_int, C.c_int, _vp, C.c_int64, _vp, C.c_uint32, C.c_int, _vp]
lib.sk_transform_from.argtypes = [C.POINTER(Transform), C.c_int, _vp, C.c_int64, _vp, C.c_uint32, C.c_int, _vp]
lib.sk_linear_combination_sharded.argtypes = [C.POINTER(_vp), C.c_int, _vp, C.c_int64, C.c_int, _vp, C.c_int64, _vp, C.c_uint32]
lib.sk_comm_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]
lib.sk_comm_destroy.argtypes = [_vp]
lib.sk_comm_size.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.sk_broadcast_coefficients.argtypes = [_vp, C.POINTER(_vp), C.c_int64, C.c_int]
lib.sk_linear_combination_sharded_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.c_int64, C.c_int, C.POINTER(_vp), C.c_int64,
                                                     C.POINTER(_vp), C.c_int, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
lib.sk_linear_combination_peers.argtypes = [_vp, _vp, C.c_int64, _vp, C.c_int64, _vp, C.POINTER(_vp), C.c_int, C.c_uint32, _vp, C.POINTER(C.c_int)]
lib.sk_is_subset_basis.argtypes = [C.POINTER(BasisDesc), C.POINTER(Transform), C.POINTER(BasisDesc), C.POINTER(Transform), C.POINTER(C.c_int)]
lib.sk_augment_coefficients.argtypes = [C.POINTER(BasisDesc), C.POINTER(Transform), C.POINTER(BasisDesc), C.POINTER(Transform), _vp, C.c_int64, C.c_int, _vp]
lib.sk_coefficient_block_flops.argtypes = [_vp, C.c_int]
lib.sk_coefficient_block_flops.restype = C.c_double
lib.sk_collocation_solve.argtypes = [_vp, _vp, C.c_int, _vp, C.c_uint32, _vp]
lib.sk_fp64_peak.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp]
lib.sk_lu_benchmark.argtypes = [_vp, C.c_int64, C.c_int, _dp, _dp, _dp]
lib.sk_policy_functions.argtypes = [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.c_int64, _vp, C.c_uint32, _vp]
lib.sk_sum_abs2.argtypes = [_vp, C.c_int64, _dp, C.c_uint32, C.c_int, _vp]
lib.sk_policy_residual_ssq.argtypes = [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.c_int64, _vp, _dp, C.c_uint32, _vp]
lib.sk_launch_count.restype = C.c_int64


def check(rc: int) -> None:
    if rc == SK_OK:
        return
    msg = (lib.sk_last_error() or b"").decode("utf-8", "replace")
    if rc == SK_ERR_ARGUMENT:
        raise ArgumentError(msg)
    if rc == SK_ERR_DOMAIN:
        raise DomainError(msg)
    if rc == SK_ERR_UNSUPPORTED:
        raise UnsupportedError(msg)
    if rc == SK_ERR_NCCL:
        raise NcclError(msg)
    raise CudaError(f"status {rc}: {msg}")
