"""General ∂ specifications (sk_general.cu) on the cfg 4a basis (Smolyak d=6, B=4): fast general nested kernel vs the
bit-exact direct kernel.  usage: python profiles/run_general.py [n_points]"""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
theta = torch.as_tensor(np.random.default_rng(1).standard_normal(sk.dimension(basis))).cuda()
x = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 1.98 - 0.99
def union(parts):
    s = sk.partial(*parts[0])
    for q in parts[1:]:
        s = s | sk.partial(*q)
    return s
e = lambda j, o=1: tuple([0] * j + [o])
specs = {"d(1,1): 4 components": [(1, 1)],
         "value + Hessian diagonal: 7": [e(j, 2)[:j + 1] for j in range(6)],
         "value + gradient + Hessian diagonal: 13": [e(j, 2)[:j + 1] for j in range(6)] + [e(j, 1)[:j + 1] for j in range(6)],
         "full Hessian (upper triangle) + gradient + value: 28": [tuple(1 if m in (a, b) else 0 for m in range(max(a, b) + 1)) if a != b else e(a, 2)[:a + 1] for a in range(6) for b in range(a, 6)]}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, parts in specs.items():
    spec = union(parts)
    info = sk.get_plan(basis, 0, spec).info
    for exact, m in ((False, n), (True, n // 16)):
        pts = spec(x[:m])
        for _ in range(2):
            out = sk.linear_combination(basis, theta, pts, exact=exact)
        torch.cuda.synchronize()
        e0.record(); out = sk.linear_combination(basis, theta, pts, exact=exact); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fl = info.flops_per_point_direct if exact else info.flops_per_point_fast
        print(json.dumps({"spec": name, "components": int(out.shape[1]), "mode": "exact" if exact else "fast", "fast_path": int(info.fast_path), "points": m,
                          "ms": ms, "Mpoints_per_s": m / ms / 1e3, "flops_per_point": fl, "tflops": m * fl / ms / 1e9}), flush=True)
