"""Driver for the coefficient-block kernel K4 (cfg 5 geometry: Smolyak d=10, B=5, InteriorGrid2, K=77 505,
Θ = K×256): device-resident points, CUDA events.  usage: python profiles/run_block.py [n_points] [reps] [nvec] [d] [B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nvec = int(sys.argv[3]) if len(sys.argv) > 3 else 256
d = int(sys.argv[4]) if len(sys.argv) > 4 else 10
B = int(sys.argv[5]) if len(sys.argv) > 5 else 5
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
K = sk.dimension(basis)
gen = torch.Generator(device="cuda").manual_seed(105)
theta = torch.randn((K, nvec), generator=gen, device="cuda", dtype=torch.float64)
x = torch.rand((n, d), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
flops = sk._lib.lib.sk_coefficient_block_flops(sk.get_plan(basis).handle, nvec)      # closed form of the kernel: 2·K·nvec + (F−1)·K + tables
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record(); out = sk.linear_combination(basis, theta, x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"kernel": "smolyak_block_kernel", "d": d, "B": B, "K": K, "nvec": nvec, "points": n, "rep": i, "ms": ms,
                      "Mpoints_per_s": n / ms / 1e3, "tflops": n * flops / ms / 1e9}), flush=True)
if len(sys.argv) > 6:      # spot check against the bit-exact direct kernel on a few points
    ref = sk.linear_combination(basis, theta, x[:256].contiguous(), exact=True)
    err = (out[:256] - ref).abs().max().item() / ref.abs().max().item()
    print(json.dumps({"max_rel_err_vs_exact_kernel": err}))
