import sys, json, os; sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << 22
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
GR = {1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}
for code, d, B in ((2, 7, 3), (2, 8, 3), (1, 8, 3), (2, 8, 2), (2, 4, 5), (1, 4, 5), (2, 3, 5)):
    basis = sk.smolyak_basis(sk.Chebyshev, GR[code], sk.SmolyakParameters(B), d)
    th = torch.as_tensor(np.random.default_rng(1).standard_normal(sk.dimension(basis))).cuda()
    x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    for _ in range(2): sk.linear_combination(basis, th, x)
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        e0.record(); sk.linear_combination(basis, th, x); e1.record(); torch.cuda.synchronize()
        best = max(best, n / e0.elapsed_time(e1) / 1e3)
    print(json.dumps({"pts_min": os.environ.get("SK_PTS_MIN", "default"), "grid": code, "d": d, "B": B, "K": sk.dimension(basis), "values_Mpoints_s": round(best, 1)}), flush=True)
