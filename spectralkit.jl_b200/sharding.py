"""Multi-GPU plumbing: points sharded, θ broadcast once, outputs gathered only on request.

Two deployment shapes, same slicing rule (contiguous slices of ceil(n/G) points):
  * one process per GPU (torchrun): `torch.distributed` is the transport (NCCL on GPUs, gloo in the CPU tests):
    `broadcast_coefficients`, `gather_outputs`;
  * one process, G devices: the library's own communicator (`Comm`, C ABI `sk_comm_*`) with the device-pointer
    entry `sk_linear_combination_sharded_device` and its in-library NCCL all-gather.
There is no reduction anywhere on this path, so nothing here touches the kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple


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


def gather_outputs(out_local, n_total: int, full=None):
    """C2 (only when the caller wants a device-resident full output): all-gather of the per-rank result slices.

    The slices are the `shard_bounds` slices, so all but the trailing ones have ceil(n/world) rows; the collective
    runs in place on one padded buffer (`all_gather_into_tensor`: a single ncclAllGather, no concatenation copy) and
    the result is the first n_total rows of it.  `full` may be a preallocated (ceil(n/world)·world, …) buffer."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out_local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_total + world - 1) // world
    tail = tuple(out_local.shape[1:])
    if full is None:
        full = torch.empty((per * world,) + tail, dtype=out_local.dtype, device=out_local.device)
    mine = full[rank * per:(rank + 1) * per]
    if out_local.data_ptr() != mine.data_ptr():
        mine[: out_local.shape[0]].copy_(out_local)
    if out_local.shape[0] < per:
        mine[out_local.shape[0]:].zero_()
    try:
        dist.all_gather_into_tensor(full, mine)
    except (RuntimeError, NotImplementedError):      # backends without the flat collective (old gloo)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous())
        full = torch.cat(parts, 0)
    return full[:n_total]


class Comm:
    """Single-process NCCL clique over `devices` (C ABI `sk_comm_create`)."""

    def __init__(self, devices: Sequence[int]):
        from . import _lib as L
        self._L = L
        self.devices = [int(g) for g in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        L.check(L.lib.sk_comm_create(len(self.devices), arr, C.byref(h)))
        self.handle = h

    @property
    def nccl_version(self) -> int:
        n, v = C.c_int(0), C.c_int(0)
        self._L.check(self._L.lib.sk_comm_size(self.handle, C.byref(n), C.byref(v)))
        return v.value

    def close(self):
        if self.handle:
            self._L.lib.sk_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def broadcast_coefficients(self, thetas, root: int = 0):
        """thetas[g]: a float64 CUDA tensor on devices[g]; the root's content is copied to all of them."""
        L = self._L
        ptrs = (C.c_void_p * len(thetas))(*[int(t.data_ptr()) for t in thetas])
        L.check(L.lib.sk_broadcast_coefficients(self.handle, ptrs, int(thetas[0].numel()), int(root)))
        return thetas

    def linear_combination(self, plans, thetas, xs, n_total: int, outs, *, nvec: int = 1, gather: bool = False, exact: bool = False):
        """Device-resident sharded evaluation (`sk_linear_combination_sharded_device`).  plans[g] / thetas[g] / xs[g] /
        outs[g] live on devices[g]; xs[g] is the g-th `shard_bounds` slice; with gather=True every outs[g] is a full
        (ceil(n/G)·G, E) buffer.  Returns (ms_compute[g], ms_gather[g]) device-timed per device."""
        L = self._L
        G = len(self.devices)
        hp = (C.c_void_p * G)(*[p.handle for p in plans])
        tp = (C.c_void_p * G)(*[int(t.data_ptr()) for t in thetas])
        xp = (C.c_void_p * G)(*[int(x.data_ptr()) for x in xs])
        op = (C.c_void_p * G)(*[int(o.data_ptr()) for o in outs])
        msc, msg = (C.c_float * G)(), (C.c_float * G)()
        theta_len = int(thetas[0].shape[0])
        flags = L.SK_MEM_DEVICE | (L.SK_MODE_EXACT if exact else L.SK_MODE_FAST)
        L.check(L.lib.sk_linear_combination_sharded_device(self.handle, hp, tp, theta_len, int(nvec), xp, int(n_total), op,
                                                           2 if gather == 2 else (1 if gather else 0), flags, msc, msg))
        return list(msc), list(msg)


def linear_combination_peers(plan, theta, x, out, peer_ptrs, stream=None) -> bool:
    """`sk_linear_combination_peers`: evaluate this rank's slice (torch tensors on this rank's device: theta (K,), x (n, d),
    out (n, E)) and store the results into the peers' buffers as well (`peer_ptrs`: device addresses of the other ranks'
    buffers mapped into this process, already offset to this slice's first row — e.g. torch symmetric memory
    `rendezvous(...).buffer_ptrs`).  Returns True when the kernel stored to the peers itself (the fused gather)."""
    from . import _lib as L
    import torch
    n = int(x.shape[0])
    ptrs = (C.c_void_p * max(1, len(peer_ptrs)))(*[int(p) for p in peer_ptrs])
    fused = C.c_int(0)
    st = C.c_void_p(int(stream.cuda_stream)) if stream is not None else C.c_void_p(int(torch.cuda.current_stream(x.device).cuda_stream))
    L.check(L.lib.sk_linear_combination_peers(plan.handle, C.c_void_p(int(theta.data_ptr())), int(theta.shape[0]), C.c_void_p(int(x.data_ptr())), n,
                                              C.c_void_p(int(out.data_ptr())), ptrs, len(peer_ptrs), L.SK_MEM_DEVICE | L.SK_MODE_FAST, st, C.byref(fused)))
    return bool(fused.value)
