"""basis_at / collocation_matrix throughput (HBM-write-bound): Smolyak d=6 B=4 values and value+gradient, InteriorGrid d=6 B=4
gradient (beyond the old 896-double limit), univariate N=128.  SK_BASIS_OLD=1 runs the round-1 kernel where it applies.
usage: python profiles/run_basis_at.py"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
gen = torch.Generator(device="cuda").manual_seed(3)
peak = 6553.0
try: peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception: pass
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def line(cfg, n, K, E, ms):
    gbs = n * K * E * 8 / ms / 1e6
    print(json.dumps({"config": cfg, "kernel": "old" if os.environ.get("SK_BASIS_OLD") else "tile+TMA", "points": n, "K": K, "E": E, "ms": ms,
                      "hbm_write_gbs": gbs, "frac_of_measured_hbm": gbs / peak}), flush=True)
b4 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
K = sk.dimension(b4)
for n in (20000, 200000, 1000000):
    u = torch.rand((n, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    line("Smolyak d=6 B=4 InteriorGrid2 basis_at values", n, K, 1, timed(lambda: sk.basis_at(b4, u)))
for n in (50000, 200000):
    u = torch.rand((n, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    line("Smolyak d=6 B=4 InteriorGrid2 basis_at value+gradient", n, K, 7, timed(lambda: sk.basis_at(b4, sk.gradient(6)(u))))
bi = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(4), 6)
Ki = sk.dimension(bi)
try:
    u = torch.rand((60000, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    line("Smolyak d=6 B=4 InteriorGrid (l=81) basis_at value+gradient", 60000, Ki, 7, timed(lambda: sk.basis_at(bi, sk.gradient(6)(u))))
    u = torch.rand((400000, 6), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    line("Smolyak d=6 B=4 InteriorGrid (l=81) basis_at values", 400000, Ki, 1, timed(lambda: sk.basis_at(bi, u)))
except Exception as e:
    print(json.dumps({"config": "InteriorGrid d=6 B=4", "error": str(e)[:200]}))
b = sk.Chebyshev(sk.InteriorGrid(), 128) @ sk.BoundedLinear(-2, 3)
y = torch.rand(4000000, generator=gen, device="cuda", dtype=torch.float64) * 5 - 2
line("Chebyshev N=128 basis_at values", 4000000, 128, 1, timed(lambda: sk.basis_at(b, y)))
