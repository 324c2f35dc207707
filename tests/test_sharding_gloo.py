"""World-size-2 CPU test (gloo) of the multi-GPU host logic: point sharding, the θ broadcast (C1) and
the optional output all-gather (C2).  The evaluator here is the CPU oracle — this test is about the
plumbing, which is identical on GPUs (NCCL) — and checks that sharded == unsharded, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle as orc
    sk = ge.load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d, B = 3, 3
    K = orc.smolyak_length(2, d, B, B)
    theta = torch.zeros(K, dtype=torch.float64)
    if rank == 0:
        theta.copy_(torch.as_tensor(np.random.default_rng(4).standard_normal(K)))
    sk.broadcast_coefficients(theta, 0)
    x = np.random.default_rng(9).uniform(-1, 1, (n, d))               # every rank can derive the full input
    lo, hi = sk.shard_bounds(n, world, rank)
    spec = orc.gradient_spec(d)
    local = orc.smolyak_lincomb(2, d, B, B, theta.numpy(), x[lo:hi], spec=spec) if hi > lo else np.zeros((0, d + 1))
    full = sk.gather_outputs(torch.as_tensor(local), n).numpy()
    if rank == 0:
        q.put((theta.numpy().copy(), full, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 64, 1])
def test_sharded_equals_unsharded_world2(n):
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    theta, full, bounds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want_theta = np.random.default_rng(4).standard_normal(len(theta))
    assert (theta == want_theta).all()
    x = np.random.default_rng(9).uniform(-1, 1, (n, 3))
    want = orc.smolyak_lincomb(2, 3, 3, 3, want_theta, x, spec=orc.gradient_spec(3))
    assert full.shape == want.shape and (full == want).all()
    assert bounds == (0, (n + 1) // 2)


def test_shard_bounds_cover_everything():
    import __graft_entry__ as ge
    sk = ge.load_package()
    for n in (0, 1, 7, 8, 100, 10**8):
        for world in (1, 2, 3, 4, 8):
            cuts = [sk.shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) <= (n + world - 1) // world
    with pytest.raises(ValueError):
        sk.shard_bounds(10, 2, 2)
