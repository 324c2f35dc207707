"""Secondary measurements for the other BASELINE.json configs (device-resident, CUDA events, 3 warm-ups,
inputs >> L2).  usage: python profiles/bench_configs.py [n_points] > profiles/r1_configs.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dfma_burst, dfma_sust = sk.fp64_peak(0, 0, 0.4)
dmma_burst, dmma_sust = sk.fp64_peak(0, 1, 0.4)
print(json.dumps({"fp64_peak_tflops": {"dfma_burst": dfma_burst, "dfma_sustained": dfma_sust, "dmma_burst": dmma_burst, "dmma_sustained": dmma_sust}}), flush=True)


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, basis, theta, pts, npts, info, exact=False, **kw):
    th = torch.as_tensor(theta).cuda()
    ms = timed(lambda: sk.linear_combination(basis, th, pts, exact=exact, **kw))
    flops = info.flops_per_point_direct if exact else info.flops_per_point_fast
    print(json.dumps({"config": name, "mode": "exact" if exact else "fast", "points": npts, "ms": ms, "Mpoints_per_s": npts / ms / 1e3,
                      "flops_per_point": flops, "tflops": npts * flops / ms / 1e9, "frac_of_measured_dfma": npts * flops / ms / 1e9 / dfma_sust,
                      "hbm_gbs": npts * info.bytes_per_point / ms / 1e6}), flush=True)


rng = np.random.default_rng(0)
gen = torch.Generator(device="cuda").manual_seed(100)
# cfg 1: latency of small calls
b20 = sk.Chebyshev(sk.InteriorGrid(), 20)
th = rng.standard_normal(20)
x = torch.rand(10_000, generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
thc = torch.as_tensor(th).cuda()
ms = timed(lambda: sk.linear_combination(b20, thc, x), reps=200, warm=20)
g = torch.as_tensor(sk.grid(b20)).cuda()
ms_c = timed(lambda: sk.collocation_matrix(b20, g), reps=200, warm=20)
print(json.dumps({"config": "cfg1 Chebyshev(InteriorGrid,20): lincomb at 1e4 points / collocation 20x20", "us_per_call_lincomb": ms * 1e3, "us_per_call_collocation": ms_c * 1e3}), flush=True)
# cfg 2
t = sk.BoundedLinear(-2, 3)
b = sk.Chebyshev(sk.InteriorGrid(), 64) @ t
y = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * 5 - 2
report("cfg2 Chebyshev N=64 o BoundedLinear(-2,3), value+d1", b, rng.standard_normal(64), sk.d(y), n, sk.get_plan(b, 1).info)
report("cfg2 Chebyshev N=64 o BoundedLinear(-2,3), value+d1", b, rng.standard_normal(64), sk.d(y), n, sk.get_plan(b, 1).info, exact=True)
# cfg 3
u = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * 1.98 - 0.99
for name, t in (("SemiInfRational(0.7,0.3)", sk.SemiInfRational(0.7, 0.3)), ("InfRational(0.4,0.9)", sk.InfRational(0.4, 0.9))):
    b = sk.Chebyshev(sk.InteriorGrid(), 128) @ t
    y = sk.transform_from(None, t, u)
    report(f"cfg3 Chebyshev N=128 o {name}, value+d1", b, rng.standard_normal(128), sk.d(y), n, sk.get_plan(b, 1).info)
    report(f"cfg3 Chebyshev N=128 o {name}, value+d1+d2 (extension)", b, rng.standard_normal(128), (sk.d ** 2)(y), n,
           sk.get_plan(b, 2, allow_rational_d2=True).info, allow_rational_d2=True)
del y, u
# cfg 4a / 4b
m = min(n, 50_000_000)
u = torch.rand((m, 6), generator=gen, device="cuda", dtype=torch.float64) * 1.98 - 0.99
b4 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
th = rng.standard_normal(2561)
report("cfg4a Smolyak d=6 B=4 untransformed, value+gradient", b4, th, sk.gradient(6)(u), m, sk.get_plan(b4, 0, sk.gradient(6)).info)
report("cfg4a Smolyak d=6 B=4 untransformed, values", b4, th, u, m, sk.get_plan(b4).info)
mm = 2_000_000
report("cfg4a Smolyak d=6 B=4 untransformed, value+gradient", b4, th, sk.gradient(6)(u[:mm]), mm, sk.get_plan(b4, 0, sk.gradient(6)).info, exact=True)
# collocation (HBM-write-bound): n x K x 8 bytes
mc = 20_000
ms = timed(lambda: sk.basis_at(b4, u[:mc]), reps=3, warm=1)
print(json.dumps({"config": "Smolyak d=6 B=4 basis_at (collocation rows), values", "points": mc, "ms": ms, "hbm_write_gbs": mc * 2561 * 8 / ms / 1e6}), flush=True)
del u
# cfg 5: Smolyak d=10 B=5 (InteriorGrid2, K = 77 505) with a K×256 coefficient block (K4, FP64 DMMA).  The spec's 10^7
# points over 8 GPUs are 1.25·10^6 per GPU; the rate does not depend on n beyond a few waves of 148 tiles of 64
# points, so 16 waves are timed here and the per-GPU time for 1.25·10^6 points is extrapolated from the rate.
b5 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(5), 10)
K5, nv = sk.dimension(b5), 256
th5 = torch.randn((K5, nv), generator=gen, device="cuda", dtype=torch.float64)
n5 = 148 * 64 * 16
x5 = torch.rand((n5, 10), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
ms = timed(lambda: sk.linear_combination(b5, th5, x5), reps=2, warm=1)
f5 = sk._lib.lib.sk_coefficient_block_flops(sk.get_plan(b5).handle, nv)
print(json.dumps({"config": "cfg5 Smolyak d=10 B=5 x (K x 256) coefficient block, values (K4, DMMA)", "mode": "fast", "points": n5, "ms": ms,
                  "Mpoints_per_s": n5 / ms / 1e3, "flops_per_point": f5, "tflops": n5 * f5 / ms / 1e9,
                  "frac_of_measured_dmma": n5 * f5 / ms / 1e9 / dmma_sust, "seconds_for_1.25e6_points_per_gpu": 1.25e6 / (n5 / ms * 1e3),
                  "hbm_gbs": n5 * (80 + 2048) / ms / 1e6}), flush=True)
