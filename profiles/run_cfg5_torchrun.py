"""cfg 5 as BASELINE.json states it: Smolyak d=10, B=5 (InteriorGrid2, K = 77 505) × Θ (K×256) at 10^7 points across the
GPUs of one box.  One process per GPU (torchrun), device-resident shards of ceil(n/G) points, Θ generated on rank 0 and
broadcast once over NCCL (C1, timed), outputs stay sharded; optional in-place all-gather (C2, timed).  Rank 0 checks a
strided sample of its slice against the CPU oracle.
usage: torchrun --nproc-per-node G profiles/run_cfg5_torchrun.py [n_total] [--gather]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as ge
sk = ge.load_package()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_total = int(args[0]) if args else 10_000_000
want_gather = "--gather" in sys.argv
d, B, nvec = 10, 5, 256
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
K = sk.dimension(basis)
lo, hi = sk.shard_bounds(n_total, world, rank)
n = hi - lo
theta = torch.empty((K, nvec), dtype=torch.float64, device=dev)
if rank == 0:
    theta.copy_(torch.as_tensor(np.random.default_rng(5).standard_normal((K, nvec))))
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sk.broadcast_coefficients(theta, 0); e1.record(); barrier()
bcast_ms = e0.elapsed_time(e1)
g = torch.Generator(device=dev).manual_seed(105 + rank)
x = torch.rand((n, d), generator=g, device=dev, dtype=torch.float64) * 2 - 1
plan = sk.get_plan(basis, 0, None, local)
import ctypes as C
L = sk._lib
out = torch.empty((n, nvec), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev)
def step():
    L.check(L.lib.sk_linear_combination(plan.handle, C.c_void_p(theta.data_ptr()), K, nvec, C.c_void_p(x.data_ptr()), n,
                                        C.c_void_p(out.data_ptr()), L.SK_MEM_DEVICE | L.SK_MODE_FAST, C.c_void_p(stream.cuda_stream)))
small = min(n, 148 * 64)
L.check(L.lib.sk_linear_combination(plan.handle, C.c_void_p(theta.data_ptr()), K, nvec, C.c_void_p(x.data_ptr()), small,
                                    C.c_void_p(out.data_ptr()), L.SK_MEM_DEVICE | L.SK_MODE_FAST, C.c_void_p(stream.cuda_stream)))   # warm-up
barrier()
times = []
for rep in range(2):
    e0.record(stream); step(); e1.record(stream); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = min(times)
flops_pt = float(L.lib.sk_coefficient_block_flops(plan.handle, nvec))
dmma_burst, dmma_sust = sk.fp64_peak(local, 1, 0.3)
res = {"config": "cfg5: Smolyak d=10 B=5 InteriorGrid2 x Theta(K x 256), device-resident shards (torchrun)", "gpus": world, "points_total": n_total,
       "points_per_gpu": n, "K": K, "nvec": nvec, "ms_max_over_ranks": ms, "all_reps_ms": times, "Mpoints_per_s": n_total / ms / 1e3,
       "flops_per_point": flops_pt, "tflops_per_gpu": n * flops_pt / (ms * 1e-3) / 1e12, "tflops_aggregate": n_total * flops_pt / (ms * 1e-3) / 1e12,
       "dmma_peak_burst_tflops": dmma_burst, "dmma_peak_sustained_tflops": dmma_sust, "frac_of_dmma_burst_per_gpu": n * flops_pt / (ms * 1e-3) / 1e12 / dmma_burst,
       "theta_broadcast_ms": bcast_ms, "theta_bytes": K * nvec * 8, "output_bytes_total": n_total * nvec * 8}
if want_gather and world > 1:
    per = (n_total + world - 1) // world
    full = torch.empty((per * world, nvec), dtype=torch.float64, device=dev)
    mine = full[rank * per:(rank + 1) * per]
    mine[:n].copy_(out)
    dist.all_gather_into_tensor(full, mine); barrier()
    e0.record(stream); dist.all_gather_into_tensor(full, mine); e1.record(stream); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    recv = per * (world - 1) * nvec * 8
    res["gather"] = {"ms": float(t.item()), "received_bytes_per_gpu": recv, "recv_gbs_per_gpu": recv / (float(t.item()) * 1e-3) / 1e9}
    del full
if rank == 0:
    from oracle import oracle as orc
    idx = np.linspace(0, n - 1, 24).astype(np.int64)
    xs = x[idx].cpu().numpy()
    th = theta.cpu().numpy()
    want = orc.smolyak_lincomb(2, d, B, B, th, xs)
    scale = orc.smolyak_lincomb(2, d, B, B, th, xs, absmode=True)
    got = out[idx].cpu().numpy()
    res["oracle_check"] = {"points": int(len(idx)), "stride": int(idx[1] - idx[0]) if len(idx) > 1 else 0, "max_scaled_error": float(np.max(np.abs(got - want) / scale)), "tolerance": 1e-13}
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
