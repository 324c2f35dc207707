"""Bit-identical (exact) mode throughput: linear_combination of the cfg-4 basis, values and value + gradient, plain and with the
cfg-4b transformations; SK_EXACT_OLD=1 runs the kernels used before the one-thread-per-point kernel.
usage: python profiles/run_exact.py [log2_points=20]"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]
ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for code, d, B, tr in ((2, 6, 4, False), (2, 6, 4, True), (1, 4, 5, False), (0, 6, 4, False), (2, 8, 3, False)):
    grid = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}[code]
    basis = sk.smolyak_basis(sk.Chebyshev, grid, sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    theta = torch.as_tensor(np.random.default_rng(4).standard_normal(K)).cuda()
    u = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    if tr:
        ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
        y = sk.transform_from(None, ct, u); basis = basis @ ct
    else:
        y = u
    for name, pts in (("values", y), ("value+gradient", sk.gradient(d)(y))):
        for _ in range(2): out = sk.linear_combination(basis, theta, pts, exact=True)
        torch.cuda.synchronize()
        e0.record(); out = sk.linear_combination(basis, theta, pts, exact=True); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"kernels": "before" if os.environ.get("SK_EXACT_OLD") else "one thread per point", "grid": code, "d": d, "B": B, "K": K, "transformed": tr,
                          "request": name, "points": n, "ms": ms, "Mpoints_per_s": n / ms / 1e3, "checksum": float(out.sum())}), flush=True)
