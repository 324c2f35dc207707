"""Few coefficient vectors (the several-policy-functions case): one unrolled-leaf launch per vector vs the DMMA
block kernel.  usage: SK_LEAF_PER_VECTOR_MAX=<0|99> python profiles/run_small_nvec.py [n_points]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
K = sk.dimension(basis)
gen = torch.Generator(device="cuda").manual_seed(7)
x = torch.rand((n, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nvec in (2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64):
    th = torch.randn((K, nvec), generator=gen, device="cuda", dtype=torch.float64)
    for _ in range(2):
        sk.linear_combination(basis, th, x)
    torch.cuda.synchronize()
    e0.record(); out = sk.linear_combination(basis, th, x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"path_max": os.environ.get("SK_LEAF_PER_VECTOR_MAX", "default"), "vec_max": os.environ.get("SK_VEC_MAX", "default"), "nvec": nvec, "points": n, "ms": ms, "Mpoints_per_s": n / ms / 1e3}), flush=True)
