"""Generates the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

  grids.npz     correctly rounded Chebyshev grids from 240-bit mpmath (oracle/py_oracle.py) — pins the
                sinpi/cospi the reference delegates to Julia Base
  indices.npz   Smolyak index sets from the NAIVE filtered Cartesian enumeration of the reference's
                own test (test/test_smolyak_traversal.jl:74-90)
  lincomb.npz   inputs + outputs of the C oracle for one small case per BASELINE config family,
                cross-checked here against the pure-Python restatement before being written

The reference itself cannot be run (no Julia), so these are oracle-generated vectors; the oracle is
pinned separately by tests/test_oracle_known_answers.py.
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
    grids = {}
    for code in (0, 1, 2):
        for N in (20, 31, 63, 128, 243):
            grids[f"g{code}_N{N}"] = np.array([po.gridpoint(code, N, i) for i in range(1, N + 1)])
    np.savez_compressed(os.path.join(HERE, "grids.npz"), **grids)

    idx = {}
    for code in (0, 1, 2):
        for (d, B, M) in ((2, 3, 3), (3, 3, 2), (4, 2, 2), (3, 4, 4)):
            if po.total_length(code, min(B, M)) ** d > 3_000_000:
                continue
            idx[f"g{code}_d{d}_B{B}_M{M}"] = np.array(po.naive_indices(code, d, B, M), dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "indices.npz"), **idx)

    rng = np.random.default_rng(20260101)
    out = {}
    # cfg 1: Chebyshev(InteriorGrid(), 20), values at 64 points + collocation at own grid
    th = rng.standard_normal(20)
    x = rng.uniform(-1, 1, 64)
    out["c1_theta"], out["c1_x"] = th, x
    out["c1_val"] = orc.uni_lincomb(20, th, x)[:, 0]
    out["c1_colloc"] = orc.uni_basis_at(20, orc.cheb_grid(0, 20))[:, :, 0].T
    # cfg 2: N=64 ∘ BoundedLinear(-2,3), value + d1, endpoint atoms
    th = rng.standard_normal(64)
    u = np.clip((rng.random(64) - 0.5) * 2.5, -1, 1)
    t = orc.BoundedLinear(-2, 3)
    y = np.array([orc.transform_from(t, v) for v in u])
    out["c2_theta"], out["c2_y"] = th, y
    out["c2_out"] = orc.uni_lincomb(64, th, y, tr=t, D=1)
    # cfg 3: N=128 ∘ SemiInfRational(0.7,0.3) / InfRational(0.4,0.9), value + d1 (+ d2 extension)
    th = rng.standard_normal(128)
    u = rng.uniform(-0.99, 0.99, 64)
    out["c3_theta"] = th
    for name, t in (("semi", orc.SemiInfRational(0.7, 0.3)), ("inf", orc.InfRational(0.4, 0.9))):
        y = np.array([orc.transform_from(t, v) for v in u])
        out[f"c3_{name}_y"] = y
        out[f"c3_{name}_d1"] = orc.uni_lincomb(128, th, y, tr=t, D=1)
        out[f"c3_{name}_d2"] = orc.uni_lincomb(128, th, y, tr=t, D=2, ext2=True)
    # cfg 4 (reduced: d=4, B=3) with mixed transformations, gradient
    d, B = 4, 3
    trs = [orc.BoundedLinear(0, 4), orc.SemiInfRational(1, 2), orc.InfRational(0, 4), orc.SemiInfRational(-3, 3)]
    ptr = [po.Tr(1, 0, 4), po.Tr(2, 1, 2), po.Tr(3, 0, 4), po.Tr(2, -3, 3)]
    K = orc.smolyak_length(2, d, B, B)
    th = rng.standard_normal(K)
    y = orc.transform_from_batch(trs, rng.uniform(-0.95, 0.95, (16, d)))
    spec = orc.gradient_spec(d)
    res = orc.smolyak_lincomb(2, d, B, B, th, y, trs=trs, spec=spec)
    for p in range(4):   # independent pure-Python restatement must agree bit for bit
        want = po.smolyak_lincomb(2, d, B, B, th.tolist(), y[p].tolist(), trs=ptr, spec=[tuple(q) for q in spec.tolist()])
        assert res[p].tolist() == want
    out["c4_theta"], out["c4_y"], out["c4_grad"] = th, y, res
    out["c4_val"] = orc.smolyak_lincomb(2, d, B, B, th, y, trs=trs)[:, 0]
    # cfg 5 (reduced: d=4, B=3, 8 coefficient vectors, values only)
    th = rng.standard_normal((K, 8))
    y = rng.uniform(-1, 1, (16, d))
    out["c5_theta"], out["c5_y"] = th, y
    out["c5_out"] = orc.smolyak_lincomb(2, d, B, B, th, y)
    np.savez_compressed(os.path.join(HERE, "lincomb.npz"), **out)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
