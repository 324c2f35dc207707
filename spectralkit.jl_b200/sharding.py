"""Multi-GPU plumbing: one process per GPU, points sharded, θ broadcast once,
outputs gathered only on request.  torch.distributed is the transport (NCCL on GPUs, gloo in the CPU
tests); there is no reduction anywhere on this path, so nothing here touches the kernels.
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ceil(n/world) points for `rank` (same rule as
    sk_linear_combination_sharded in csrc/sk_api.cu)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    per = (n + world - 1) // world
    lo = min(n, per * rank)
    return lo, min(n, lo + per)


def broadcast_coefficients(theta, src: int = 0):
    """C1: one broadcast of θ (20 KB at cfg 4, 159 MB at cfg 5), amortised over all calls."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(theta, src)
    return theta


def gather_outputs(out_local, n_total: int):
    """C2 (only when the caller wants a device-resident full output): all-gather of the per-rank
    result slices, padded to the common slice length and trimmed to n_total rows."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out_local
    world = dist.get_world_size()
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(out_local.shape[1:]), dtype=out_local.dtype, device=out_local.device)
    pad[: out_local.shape[0]] = out_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat(parts, 0)[:n_total]
