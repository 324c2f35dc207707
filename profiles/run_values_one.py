"""One plain-values evaluation of cfg 4a at 2^22 points (two-points-per-thread kernel; SK_PTS_MIN=-1: the one-point leaf kernel), for ncu.
usage: python profiles/run_values_one.py"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << 22
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
theta = torch.as_tensor(np.random.default_rng(4).standard_normal(sk.dimension(basis))).cuda()
x = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 1.98 - 0.99
for _ in range(3):
    out = sk.linear_combination(basis, theta, x)
torch.cuda.synchronize()
print(float(out.sum()))
