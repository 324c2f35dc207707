"""Higher-dimensional Smolyak bases (leaf kernel with the unit interpreter above 6 dimensions, generic nested kernel
otherwise): value + gradient and values.  usage: python profiles/run_highdim.py [n_points]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for d, B in ((6, 4), (7, 2), (7, 3), (7, 4), (8, 2), (8, 3), (8, 4), (9, 2), (9, 3), (10, 3), (10, 2), (11, 2), (11, 3), (12, 2), (12, 3), (13, 2), (14, 2), (15, 2), (16, 2), (5, 5)):
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    th = torch.as_tensor(np.random.default_rng(1).standard_normal(K)).cuda()
    x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    for name, pts, spec in (("value+gradient", sk.gradient(d)(x), sk.gradient(d)), ("values", x, None)):
        info = sk.get_plan(basis, 0, spec).info
        for _ in range(2):
            sk.linear_combination(basis, th, pts)
        torch.cuda.synchronize()
        e0.record(); sk.linear_combination(basis, th, pts); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"d": d, "B": B, "K": K, "request": name, "fast_path": int(info.fast_path), "points": n, "ms": ms, "Mpoints_per_s": n / ms / 1e3,
                          "flops_per_point": info.flops_per_point_fast, "tflops": n * info.flops_per_point_fast / ms / 1e9}), flush=True)
