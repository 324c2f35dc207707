"""Budget 4 in 8..10 dimensions (looped deep single leaves, sk_leaf_impl.cuh: leaf_single_deep_looped_c) vs the leaf + interpreted upper
dimensions (SK_LEAF_NO_SINGLE=1): value + gradient and values, each checked against the bit-identical mode.
usage: python profiles/run_budget4_deep.py"""
import sys, json; sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
GR = {1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}
for code, d, B in ((2, 10, 4), (2, 9, 4), (1, 10, 4), (1, 9, 4), (2, 8, 4), (2, 10, 3)):
    basis = sk.smolyak_basis(sk.Chebyshev, GR[code], sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    th = torch.as_tensor(np.random.default_rng(1).standard_normal(K)).cuda()
    x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    rec = {"grid": code, "d": d, "B": B, "K": K}
    for name, wrap, spec in (("grad", sk.gradient(d), sk.gradient(d)), ("values", (lambda t: t), None)):
        info = sk.get_plan(basis, 0, spec).info
        small = x[:1024].contiguous()
        fast = sk.linear_combination(basis, th, wrap(small)); exact = sk.linear_combination(basis, th, wrap(small), exact=True)
        rec[name + "_err"] = float((fast - exact).abs().max() / exact.abs().max())
        for _ in range(2): sk.linear_combination(basis, th, wrap(x))
        torch.cuda.synchronize()
        best = 0
        for _ in range(3):
            e0.record(); sk.linear_combination(basis, th, wrap(x)); e1.record(); torch.cuda.synchronize()
            best = max(best, n / e0.elapsed_time(e1) / 1e3)
        rec[name + "_Mpoints_s"] = round(best, 1); rec[name + "_tflops"] = round(best * 1e6 * info.flops_per_point_fast / 1e12, 2); rec[name + "_path"] = int(info.fast_path)
    print(json.dumps(rec), flush=True)
