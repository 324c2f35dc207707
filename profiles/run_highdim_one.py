import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
d, B, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
th = torch.as_tensor(np.random.default_rng(1).standard_normal(sk.dimension(basis))).cuda()
x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
for _ in range(2):
    sk.linear_combination(basis, th, sk.gradient(d)(x))
torch.cuda.synchronize()
