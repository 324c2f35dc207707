"""Univariate driver for ncu: cfg 2 (Chebyshev N=64 ∘ BoundedLinear, value + d1) or cfg 3 (N=128 ∘ SemiInfRational).
usage: python profiles/run_uni.py [n_points] [reps] [cfg]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if cfg == 2:
    N, t = 64, sk.BoundedLinear(-2, 3)
    y = torch.rand(n, device="cuda", dtype=torch.float64) * 5 - 2
else:
    N, t = 128, sk.SemiInfRational(0.7, 0.3)
    y = sk.transform_from(None, t, torch.rand(n, device="cuda", dtype=torch.float64) * 1.98 - 0.99)
b = sk.Chebyshev(sk.InteriorGrid(), N) @ t
th = torch.as_tensor(np.random.default_rng(2).standard_normal(N)).cuda()
flops = (N - 1) * 6 + N * 4
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record(); out = sk.linear_combination(b, th, sk.d(y)); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {i}: {ms:.3f} ms  {n / ms / 1e3:.0f} Mpoints/s  {n * flops / ms / 1e9:.2f} TFLOP/s")
