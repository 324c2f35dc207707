import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
gen = torch.Generator(device="cuda").manual_seed(3)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
b4 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
for n in (20000, 200000):
    u = torch.rand((n, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    ms = timed(lambda: sk.basis_at(b4, u))
    print(json.dumps({"config": "Smolyak d=6 B=4 basis_at values", "points": n, "ms": ms, "hbm_write_gbs": n * 2561 * 8 / ms / 1e6}))
u = torch.rand((50000, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
ms = timed(lambda: sk.basis_at(b4, sk.gradient(6)(u)))
print(json.dumps({"config": "Smolyak d=6 B=4 basis_at value+gradient", "points": 50000, "ms": ms, "hbm_write_gbs": 50000 * 2561 * 56 / ms / 1e6}))
b = sk.Chebyshev(sk.InteriorGrid(), 128) @ sk.BoundedLinear(-2, 3)
y = torch.rand(4000000, generator=gen, device="cuda", dtype=torch.float64) * 5 - 2
ms = timed(lambda: sk.basis_at(b, y))
print(json.dumps({"config": "Chebyshev N=128 basis_at values", "points": 4000000, "ms": ms, "hbm_write_gbs": 4e6 * 128 * 8 / ms / 1e6}))
