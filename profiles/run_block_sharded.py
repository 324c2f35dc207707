"""cfg 5 through the single-process sharded driver (sk_linear_combination_sharded): Smolyak d=10, B=5, Θ = K×256,
points split over all visible GPUs, host buffers in and out.  usage: python profiles/run_block_sharded.py [points_per_gpu]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
G = torch.cuda.device_count()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 8
n = per * G
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(5), 10)
K, nvec = sk.dimension(basis), 256
rng = np.random.default_rng(105)
theta = rng.standard_normal((K, nvec))
x = rng.uniform(-1, 1, (n, 10))
for rep in range(2):
    t0 = time.perf_counter()
    out = sk.linear_combination_sharded(basis, theta, x, devices=list(range(G)))
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "cfg5 sharded driver", "gpus": G, "points": n, "seconds": dt, "Mpoints_per_s": n / dt / 1e6,
                      "tflops_aggregate": n * 2.0 * K * nvec / dt / 1e12}), flush=True)
one = sk.linear_combination(basis, theta, x[:64])
print(json.dumps({"max_abs_diff_vs_single_gpu_call": float(np.abs(out[:64] - one).max())}))
