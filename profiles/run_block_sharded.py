"""cfg 5 through the single-process sharded driver (sk_linear_combination_sharded): Smolyak d=10, B=5, Θ = K×256,
points split over all visible GPUs, HOST buffers in and out (the copies are inside the call).  A strided sample of the
result is checked against the CPU oracle.  usage: python profiles/run_block_sharded.py [points_per_gpu] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
G = torch.cuda.device_count()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = per * G
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(5), 10)
K, nvec = sk.dimension(basis), 256
rng = np.random.default_rng(105)
theta = rng.standard_normal((K, nvec))
x = rng.uniform(-1, 1, (n, 10))
for rep in range(reps):
    t0 = time.perf_counter()
    out = sk.linear_combination_sharded(basis, theta, x, devices=list(range(G)))
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "cfg5 through sk_linear_combination_sharded (host buffers, pageable numpy arrays)", "gpus": G, "points": n, "seconds": dt,
                      "Mpoints_per_s": n / dt / 1e6, "tflops_aggregate": n * 2.0 * K * nvec / dt / 1e12,
                      "host_bytes_in": x.nbytes + theta.nbytes, "host_bytes_out": out.nbytes}), flush=True)
from oracle import oracle as orc
idx = np.linspace(0, n - 1, 16).astype(np.int64)
want = orc.smolyak_lincomb(2, 10, 5, 5, theta, x[idx])
scale = orc.smolyak_lincomb(2, 10, 5, 5, theta, x[idx], absmode=True)
print(json.dumps({"oracle_check_points": int(len(idx)), "max_scaled_error": float(np.max(np.abs(out[idx] - want) / scale)), "tolerance": 1e-13}))
