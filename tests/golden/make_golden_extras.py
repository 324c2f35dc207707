"""Golden fixtures for the paths added after the first set (run from the repo root:
python tests/golden/make_golden_extras.py) → tests/golden/extras.npz.  Oracle-generated like lincomb.npz (the reference
cannot be run here); every Smolyak vector is cross-checked against the pure-Python restatement on a few points.

  k4_*    coefficient block (SVector coefficients): Smolyak d=5, B=3, EndpointGrid, 40 vectors, mixed transformations
  gs_*    general ∂ specification ∂(1,1) ∪ ∂(0,0,2) ∪ ∂(2): Smolyak d=3, B=3, InteriorGrid, transformations (order 2
          only through BoundedLinear, the reference's behaviour)
  ag_*    augment_coefficients through the PaddingIterator restatement: InteriorGrid2 d=3, (B,M) = (2,1) → (4,3)
  cs_*    fit on the basis' own grid: collocation matrix of the oracle, right-hand sides, LAPACK solution
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from oracle import py_oracle as po        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260102)
    out = {}
    # ---- k4
    d, B, nvec = 5, 3, 40
    trs = [orc.BoundedLinear(0, 4), orc.SemiInfRational(1, 2), orc.InfRational(0, 4), orc.BoundedLinear(-1, 2), orc.SemiInfRational(-3, 3)]
    ptr = [po.Tr(1, 0, 4), po.Tr(2, 1, 2), po.Tr(3, 0, 4), po.Tr(1, -1, 2), po.Tr(2, -3, 3)]
    K = orc.smolyak_length(1, d, B, B)
    th = rng.standard_normal((K, nvec))
    y = orc.transform_from_batch(trs, rng.uniform(-0.95, 0.95, (24, d)))
    res = orc.smolyak_lincomb(1, d, B, B, th, y, trs=trs)
    for p in range(2):
        for v in (0, nvec - 1):
            want = po.smolyak_lincomb(1, d, B, B, th[:, v].tolist(), y[p].tolist(), trs=ptr)
            assert [res[p, v]] == list(np.atleast_1d(want)), (res[p, v], want)
    out["k4_theta"], out["k4_y"], out["k4_out"] = th, y, res
    out["k4_scale"] = orc.smolyak_lincomb(1, d, B, B, th, y, trs=trs, absmode=True)
    # ---- general spec
    d, B = 3, 3
    trs = [orc.BoundedLinear(0, 4), orc.SemiInfRational(1, 2), orc.BoundedLinear(-2, 2)]
    ptr = [po.Tr(1, 0, 4), po.Tr(2, 1, 2), po.Tr(1, -2, 2)]
    parts = [(1, 1), (0, 0, 2), (2,)]
    spec = orc.deriv_spec(parts, d)
    K = orc.smolyak_length(0, d, B, B)
    th = rng.standard_normal(K)
    y = orc.transform_from_batch(trs, rng.uniform(-0.95, 0.95, (24, d)))
    res = orc.smolyak_lincomb(0, d, B, B, th, y, trs=trs, spec=spec)
    for p in range(2):
        want = po.smolyak_lincomb(0, d, B, B, th.tolist(), y[p].tolist(), trs=ptr, spec=[tuple(q) for q in spec.tolist()])
        assert res[p].tolist() == want
    out["gs_theta"], out["gs_y"], out["gs_out"], out["gs_spec"] = th, y, res, spec
    out["gs_scale"] = orc.smolyak_lincomb(0, d, B, B, th, y, trs=trs, spec=spec, absmode=True)
    # ---- augment
    th1 = rng.standard_normal(orc.smolyak_length(2, 3, 2, 1))
    out["ag_theta1"] = th1
    out["ag_theta2"] = np.array(po.augment_coefficients(2, 3, 2, 1, 4, 3, th1.tolist()))
    # ---- collocation solve
    d, B = 3, 3
    g = orc.smolyak_grid(2, d, B, B)
    C = orc.smolyak_basis_at(2, d, B, B, g)[:, :, 0].T
    Y = np.stack([np.exp(g[:, 0]) * np.cos(g[:, 1] + g[:, 2]), 1.0 / (2.0 + g[:, 0] * g[:, 1]) + g[:, 2] ** 3], axis=1)
    out["cs_y"], out["cs_theta"], out["cs_cond"] = Y, np.linalg.solve(C, Y), np.array(np.linalg.cond(C))
    np.savez_compressed(os.path.join(HERE, "extras.npz"), **out)
    print("written extras.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
