"""The library's own LU (csrc/sk_solve.cu): collocation solve of the headline basis (K = 2561) against LAPACK on the same
matrix, and the factorisation rate on random diagonally-unfriendly matrices.  usage: python profiles/run_solve.py"""
import ctypes as C, json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as ge
sk = ge.load_package()
L = sk._lib
dmma_burst, _ = sk.fp64_peak(0, 1, 0.3)
for K in (2561, 4096, 8192):
    A = torch.randn((K, K), dtype=torch.float64, device="cuda")
    ms, tf = C.c_double(0), C.c_double(0)
    mg = C.c_double(0); L.check(L.lib.sk_lu_benchmark(C.c_void_p(A.data_ptr()), K, 3, C.byref(ms), C.byref(tf), C.byref(mg)))
    print(json.dumps({"config": "LU factorisation (partial pivoting), random K x K", "K": K, "ms": ms.value, "tflops": tf.value,
                      "ms_trailing_updates_dmma": mg.value, "trailing_update_tflops": (2.0 / 3.0) * K ** 3 / (mg.value * 1e-3) * 1e-12 if mg.value else None, "trailing_update_frac_of_dmma_burst": (2.0 / 3.0) * K ** 3 / (mg.value * 1e-3) * 1e-12 / dmma_burst if mg.value else None}), flush=True)
basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
K = sk.dimension(basis)
g = sk.grid(basis)
f = lambda x: np.exp(-0.3 * (x ** 2).sum(1)) + 0.2 * x[:, 0] * x[:, 3]
Y = np.stack([f(g), np.cos(g.sum(1))], 1)
t0 = time.perf_counter(); th = sk.collocation_solve(basis, Y); dt = time.perf_counter() - t0
t0 = time.perf_counter(); th = sk.collocation_solve(basis, Y); dt2 = time.perf_counter() - t0
Cm = sk.collocation_matrix(basis)
want = np.linalg.solve(Cm, Y)
cond = np.linalg.cond(Cm)
err = float(np.max(np.abs(th - want)) / np.max(np.abs(want)))
res = float(np.max(np.abs(Cm @ th - Y)))
print(json.dumps({"config": "collocation_solve cfg-4 basis (grid + matrix + LU + solves), 2 right-hand sides", "K": K, "seconds_first": dt, "seconds": dt2,
                  "rel_diff_vs_lapack": err, "cond": cond, "bound_50_cond_eps": 50 * cond * 2.2e-16, "max_residual": res}))
