import sys, json, os; sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << 24
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
th = torch.as_tensor(np.random.default_rng(1).standard_normal(sk.dimension(basis))).cuda()
x = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 1.98 - 0.99
pts = sk.gradient(6)(x)
small = sk.gradient(6)(x[:100003].contiguous())
fast = sk.linear_combination(basis, th, small); exact = sk.linear_combination(basis, th, small, exact=True)
err = float((fast - exact).abs().max() / exact.abs().max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2): sk.linear_combination(basis, th, pts)
torch.cuda.synchronize()
best = 0
for _ in range(4):
    e0.record(); sk.linear_combination(basis, th, pts); e1.record(); torch.cuda.synchronize()
    best = max(best, n / e0.elapsed_time(e1) / 1e3)
print(json.dumps({"lib": os.path.basename(os.environ.get("SK_B200_LIB", "shipped")), "gpts": os.environ.get("SK_GPTS", "0"), "err": err, "grad_Mpoints_s": round(best, 1)}), flush=True)
