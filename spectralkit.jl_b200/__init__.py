"""spectralkit_b200 — B200-native batched basis evaluation behind the SpectralKit.jl API.

The directory is named `spectralkit.jl_b200` (not a valid Python identifier); import it through
`__graft_entry__.load_package()`, which registers it as the module `spectralkit_b200`.

Layout
    csrc/    CUDA kernels (sm_100a), host plan builder, the C ABI (include/spectralkit_b200.h)
    lib/     libspectralkit_b200.so (built in-tree; git-ignored)
    api.py   host-side mirror of the reference's Julia API (executable stand-in for the shim)
    julia/   the Julia `ccall` shim (written against the same ABI; cannot be run here)
"""
from .api import *  # noqa: F401,F403
from .api import d, get_plan, fp64_peak, launch_count, smolyak_indices, grid_shuffle  # noqa: F401
from . import _lib  # noqa: F401
from . import experimental  # noqa: F401
from .experimental import bridge, make_policy_functions, sum_of_squared_residuals, sum_abs2  # noqa: F401
from .sharding import shard_bounds, broadcast_coefficients, gather_outputs, linear_combination_peers, Comm  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
