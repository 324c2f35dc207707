"""Small driver for ncu captures: runs the headline Smolyak linear_combination (cfg 4b, value+gradient)
a few times on device-resident points.  usage: python profiles/run_kernel.py [n_points] [reps] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "grad"
MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]
ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6) @ ct
theta = torch.as_tensor(np.random.default_rng(4).standard_normal(2561)).cuda()
u = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 1.98 - 0.99
y = sk.transform_from(None, ct, u)
pts = sk.gradient(6)(y) if mode == "grad" else y
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record(); out = sk.linear_combination(basis, theta, pts); e1.record(); torch.cuda.synchronize()
    print(f"rep {i}: {e0.elapsed_time(e1):.3f} ms  {n / e0.elapsed_time(e1) / 1e3:.1f} Mpoints/s")
