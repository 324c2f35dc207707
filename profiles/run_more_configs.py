"""Value + gradient over other grid families, dimensions and budgets (2^20 points): which kernel serves them and how fast.
usage: python profiles/run_more_configs.py > profiles/r1_more_configs.jsonl"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for code, d, B in ((2, 6, 5), (2, 4, 5), (2, 8, 4), (2, 10, 4), (1, 6, 4), (0, 6, 3), (0, 6, 4), (0, 4, 4), (1, 10, 3), (0, 10, 2)):
    basis = sk.smolyak_basis(sk.Chebyshev, {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}[code], sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    th = torch.as_tensor(np.random.default_rng(1).standard_normal(K)).cuda()
    x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    info = sk.get_plan(basis, 0, sk.gradient(d)).info
    for _ in range(2):
        sk.linear_combination(basis, th, sk.gradient(d)(x))
    torch.cuda.synchronize()
    e0.record(); sk.linear_combination(basis, th, sk.gradient(d)(x)); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"grid": code, "d": d, "B": B, "K": K, "fast_path": int(info.fast_path), "Mpoints_per_s": round(n / ms / 1e3, 1), "tflops": round(n * info.flops_per_point_fast / ms / 1e9, 2)}), flush=True)
