"""Coverage matrix: value + gradient (and values only) over every grid family, d = 2..6, B = 2..5: kernel and rate.
usage: python profiles/run_matrix.py [log2_points=20] > profiles/r1_matrix.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
GR = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}
for code in (0, 1, 2):
    for d in (2, 3, 4, 5, 6):
        for B in (2, 3, 4, 5):
            basis = sk.smolyak_basis(sk.Chebyshev, GR[code], sk.SmolyakParameters(B), d)
            K = sk.dimension(basis)
            if K > 30000:
                continue
            th = torch.as_tensor(np.random.default_rng(1).standard_normal(K)).cuda()
            x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
            info = sk.get_plan(basis, 0, sk.gradient(d)).info
            for _ in range(2):
                sk.linear_combination(basis, th, sk.gradient(d)(x))
            torch.cuda.synchronize()
            e0.record(); sk.linear_combination(basis, th, sk.gradient(d)(x)); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            for _ in range(2):
                sk.linear_combination(basis, th, x)
            torch.cuda.synchronize()
            e0.record(); sk.linear_combination(basis, th, x); e1.record(); torch.cuda.synchronize()
            msv = e0.elapsed_time(e1)
            print(json.dumps({"grid": code, "d": d, "B": B, "K": K, "fast_path": int(info.fast_path), "Mpoints_per_s": round(n / ms / 1e3, 1),
                              "tflops": round(n * info.flops_per_point_fast / ms / 1e9, 2), "values_Mpoints_per_s": round(n / msv / 1e3, 1)}), flush=True)
