import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
d, B, n = 6, 4, 1 << 24
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
th = torch.as_tensor(np.random.default_rng(1).standard_normal(sk.dimension(basis))).cuda()
x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    sk.linear_combination(basis, th, sk.gradient(d)(x))
torch.cuda.synchronize()
r = []
for _ in range(5):
    e0.record(); sk.linear_combination(basis, th, sk.gradient(d)(x)); e1.record(); torch.cuda.synchronize()
    r.append(n / e0.elapsed_time(e1) / 1e3)
print(os.environ.get("SK_B200_LIB", "shipped"), [round(v) for v in r])
