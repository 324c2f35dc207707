"""One bit-identical (exact) evaluation of cfg 4b, value + gradient (for ncu).  usage: python profiles/run_exact_one.py [log2_points=19]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 19)
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
theta = torch.as_tensor(np.random.default_rng(4).standard_normal(sk.dimension(basis))).cuda()
x = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 1.98 - 0.99
for _ in range(2):
    out = sk.linear_combination(basis, theta, sk.gradient(6)(x), exact=True)
th4 = torch.randn((sk.dimension(basis), 4), device="cuda", dtype=torch.float64)
for _ in range(2):
    out4 = sk.linear_combination(basis, th4, x)
torch.cuda.synchronize()
print(float(out.sum()), float(out4.sum()))
