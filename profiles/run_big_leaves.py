"""Times the largest single-budget leaves (d = 6: InteriorGrid2 budget 5, K = 10 625; InteriorGrid budget 4, K = 4 361), value +
gradient and values only, for one or more builds of the library (SK_B200_LIB), one subprocess each; every build is checked
against the bit-exact mode on 2048 points.  usage: python profiles/run_big_leaves.py [--n log2] lib_a.so ... ("shipped")"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << int(sys.argv[1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
GR = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}
CFGS = [tuple(int(v) for v in c.split(",")) for c in os.environ.get("SK_BIG_CFGS", "2,6,5;0,6,4;1,6,5").split(";")]
for code, d, B in CFGS:
    basis = sk.smolyak_basis(sk.Chebyshev, GR[code], sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    th = torch.as_tensor(np.random.default_rng(1).standard_normal(K)).cuda()
    x = torch.rand((n, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    rec = {"lib": os.path.basename(os.environ.get("SK_B200_LIB", "shipped")), "grid": code, "d": d, "B": B, "K": K}
    for name, wrap in (("grad", sk.gradient(d)), ("values", lambda t: t))[:1 if os.environ.get("SK_BIG_GRAD_ONLY") else 2]:
        try:
            info = sk.get_plan(basis, 0, sk.gradient(d)).info if name == "grad" else None
            small = x[:2048].contiguous()
            fast = sk.linear_combination(basis, th, wrap(small))
            exact = sk.linear_combination(basis, th, wrap(small), exact=True)
            rec[name + "_err"] = float((fast - exact).abs().max() / exact.abs().max())
            for _ in range(2): sk.linear_combination(basis, th, wrap(x))
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                e0.record(); sk.linear_combination(basis, th, wrap(x)); e1.record(); torch.cuda.synchronize()
                best = max(best, n / e0.elapsed_time(e1) / 1e3)
            rec[name + "_Mpoints_s"] = round(best, 1)
            if info is not None: rec["grad_tflops"] = round(best * 1e6 * info.flops_per_point_fast / 1e12, 2)
        except Exception as ex:
            rec[name + "_error"] = str(ex)[:200]
    print(json.dumps(rec), flush=True)
''' % ROOT
lg = 20
args = sys.argv[1:]
if args and args[0] == "--n": lg = int(args[1]); args = args[2:]
for lib in args:
    env = dict(os.environ)
    if lib != "shipped":
        env["SK_B200_LIB"] = os.path.join(ROOT, lib) if not os.path.isabs(lib) else lib
    p = subprocess.run([sys.executable, "-c", CHILD, str(lg)], env=env, capture_output=True, text=True)
    print(p.stdout.strip() or ("FAILED " + lib + ": " + p.stderr[-800:]), flush=True)
    if p.returncode != 0 and p.stdout.strip(): print("STDERR " + p.stderr[-400:], flush=True)
