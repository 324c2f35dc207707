import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = 1 << 22
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
K = sk.dimension(basis)
gen = torch.Generator(device="cuda").manual_seed(7)
x = torch.rand((n, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nvec in (8, 64, 256):
    th = torch.randn((K, nvec), generator=gen, device="cuda", dtype=torch.float64)
    for _ in range(2): sk.linear_combination(basis, th, x)
    torch.cuda.synchronize()
    e0.record(); out = sk.linear_combination(basis, th, x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"dbg": os.environ.get("SK_BLOCK_DBG", "0"), "nvec": nvec, "ms": ms, "tflops": n * 2.0 * K * nvec / ms / 1e9}), flush=True)
