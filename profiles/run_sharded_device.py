"""Single-process multi-GPU through the C ABI: sk_comm_create + sk_broadcast_coefficients + sk_linear_combination_sharded_device
(device-resident shards, optional in-library NCCL all-gather), for cfg 4b (value + gradient) and cfg 5 (K×256 block).
usage: python profiles/run_sharded_device.py [cfg4|cfg5] [n_total]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
from spectralkit_b200.sharding import Comm
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
G = torch.cuda.device_count()
devices = list(range(G))
comm = Comm(devices)
if cfg == "cfg4":
    n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]
    ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
    ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
    d = 6
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), d) @ ct
    spec, nvec, E = sk.gradient(d), 1, d + 1
else:
    n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    d = 10
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(5), d)
    ct, spec, nvec, E = None, None, 256, 256
K = sk.dimension(basis)
plans = [sk.get_plan(basis, 0, spec, g) for g in devices]
per = (n_total + G - 1) // G
thetas, xs, outs_s, outs_f = [], [], [], []
for g in devices:
    with torch.cuda.device(g):
        dev = torch.device("cuda", g)
        shape = (K,) if nvec == 1 else (K, nvec)
        th = torch.zeros(shape, dtype=torch.float64, device=dev)
        if g == 0:
            th.copy_(torch.as_tensor(np.random.default_rng(4).standard_normal(shape)))
        thetas.append(th)
        lo, hi = sk.shard_bounds(n_total, G, g)
        gen = torch.Generator(device=dev).manual_seed(104 + g)
        u = torch.rand((hi - lo, d), generator=gen, device=dev, dtype=torch.float64) * 1.98 - 0.99
        xs.append(sk.transform_from(None, ct, u) if ct is not None else u)
        outs_s.append(torch.empty((hi - lo, E), dtype=torch.float64, device=dev))
for g in devices: torch.cuda.synchronize(g)
comm.broadcast_coefficients(thetas, 0)
assert all(torch.equal(thetas[0].cpu(), t.cpu()) for t in thetas), "broadcast failed"
res = {"config": cfg + " through sk_linear_combination_sharded_device (one process, in-library NCCL)", "gpus": G, "points_total": n_total, "nccl_version": comm.nccl_version}
for rep in range(2):
    msc, _ = comm.linear_combination(plans, thetas, xs, n_total, outs_s, nvec=nvec, gather=False)
res["sharded_ms_max"] = max(msc); res["sharded_Mpoints_per_s"] = n_total / max(msc) / 1e3
# gather: full buffers on every device
need = per * G * E * 8
free = min(torch.cuda.mem_get_info(g)[0] for g in devices)
if need < free * 0.9:
    for g in devices:
        outs_f.append(torch.empty((per * G, E), dtype=torch.float64, device=torch.device("cuda", g)))
    for rep in range(2):
        msc, msg = comm.linear_combination(plans, thetas, xs, n_total, outs_f, nvec=nvec, gather=True)
    recv = per * (G - 1) * E * 8
    res.update({"gather_compute_ms_max": max(msc), "gather_ms_max": max(msg), "gather_received_bytes_per_gpu": recv,
                "gather_recv_gbs_per_gpu": recv / (max(msg) * 1e-3) / 1e9 if max(msg) > 0 else None, "full_output_bytes": need})
    # every device must hold every slice
    ok = True
    for g in devices:
        lo, hi = sk.shard_bounds(n_total, G, g)
        m = min(hi - lo, 1000)
        for h in (0, G - 1):
            ok &= bool(torch.equal(outs_f[h][lo:lo + m].cpu(), outs_s[g][:m].cpu()))
    res["gathered_equals_sharded"] = ok
    # gather mode 2: slices pushed peer to peer over NVLink while the next chunk is computed
    for o in outs_f: o.zero_()
    for g in devices: torch.cuda.synchronize(g)
    for rep in range(2):
        msc2, msg2 = comm.linear_combination(plans, thetas, xs, n_total, outs_f, nvec=nvec, gather=2)
    res.update({"overlapped_compute_ms_max": max(msc2), "overlapped_tail_ms_max": max(msg2),
                "overlapped_total_ms_max": max(a + b for a, b in zip(msc2, msg2)), "nccl_total_ms_max": max(a + b for a, b in zip(msc, msg))})
    ok2 = True
    for g in devices:
        lo, hi = sk.shard_bounds(n_total, G, g)
        m = min(hi - lo, 1000)
        for h in range(G):
            ok2 &= bool(torch.equal(outs_f[h][lo:lo + m].cpu(), outs_s[g][:m].cpu()))
            ok2 &= bool(torch.equal(outs_f[h][hi - m:hi].cpu(), outs_s[g][hi - lo - m:].cpu()))
    res["overlapped_equals_sharded"] = ok2
else:
    res["gather_skipped"] = f"full output of {need / 1e9:.1f} GB does not fit next to the shards"
print(json.dumps(res), flush=True)
comm.close()
