"""Times the headline kernel (cfg 4b: d=6, B=4, InteriorGrid2, mixed transformations, value + gradient) for several
builds of the library, one subprocess each (SK_B200_LIB=<lib>), and checks each build against the bit-exact mode on
4096 points.  usage: python profiles/run_variants.py [--n N] lib_a.so lib_b.so ...   ("shipped" = the in-tree library)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1])
MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]
ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6) @ ct
theta = torch.as_tensor(np.random.default_rng(4).standard_normal(2561)).cuda()
g = torch.Generator(device="cuda"); g.manual_seed(7)
u = torch.rand((n, 6), device="cuda", dtype=torch.float64, generator=g) * 1.98 - 0.99
y = sk.transform_from(None, ct, u)
pts = sk.gradient(6)(y)
small = sk.gradient(6)(y[:4096].contiguous())
fast = sk.linear_combination(basis, theta, small)
exact = sk.linear_combination(basis, theta, small, exact=True)
err = float((fast - exact).abs().max() / exact.abs().max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2): sk.linear_combination(basis, theta, pts)
torch.cuda.synchronize()
r = []
for _ in range(5):
    e0.record(); out = sk.linear_combination(basis, theta, pts); e1.record(); torch.cuda.synchronize()
    r.append(n / e0.elapsed_time(e1) / 1e3)
print(json.dumps({"lib": os.environ.get("SK_B200_LIB", "shipped"), "n": n, "mpoints_s": [round(v, 1) for v in r], "best": round(max(r), 1), "rel_err_vs_exact": err, "checksum": float(out.sum())}))
''' % ROOT
n = 1 << 24
args = sys.argv[1:]
if args and args[0] == "--n": n = int(args[1]); args = args[2:]
for lib in args:
    env = dict(os.environ)
    if lib != "shipped":
        env["SK_B200_LIB"] = os.path.join(ROOT, lib) if not os.path.isabs(lib) else lib
        env["SK_LEAF_NO_SINGLE"] = "1"
    p = subprocess.run([sys.executable, "-c", CHILD, str(n)], env=env, capture_output=True, text=True)
    print(p.stdout.strip() or ("FAILED " + lib + ": " + p.stderr[-800:]), flush=True)
