"""End-to-end (host buffers, pinned) throughput of the headline call, as bench.py's `e2e` measures it.
usage: python profiles/run_e2e.py [n_points] [reps]      env SK_HOST_CHUNK=points per pipeline chunk"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]
ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6) @ ct
theta = np.random.default_rng(4).standard_normal(2561)
xh = torch.empty((n, 6), dtype=torch.float64, pin_memory=True)
u = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 1.98 - 0.99
xh.copy_(sk.transform_from(None, ct, u)); del u
oh = torch.empty((n, 7), dtype=torch.float64, pin_memory=True)
plan = sk.get_plan(basis, 0, sk.gradient(6), 0)
import ctypes as C
L = sk._lib
def step():
    L.check(L.lib.sk_linear_combination(plan.handle, C.c_void_p(theta.ctypes.data), 2561, 1, C.c_void_p(xh.data_ptr()), n,
                                        C.c_void_p(oh.data_ptr()), L.SK_MEM_HOST | L.SK_MODE_FAST, None))
step(); torch.cuda.synchronize()
for i in range(reps):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"points": n, "chunk": os.environ.get("SK_HOST_CHUNK", "default"), "ms": dt * 1e3, "Mpoints_per_s": n / dt / 1e6, "GBps_aggregate": n * 104 / dt / 1e9}), flush=True)
