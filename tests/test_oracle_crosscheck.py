"""C oracle vs the independent pure-Python restatement: must agree bit for bit."""
import numpy as np
import pytest

from oracle import py_oracle as po

INT, END, INT2 = 0, 1, 2


def _pair(orc, kind, a, b):
    return ({1: orc.BoundedLinear, 2: orc.SemiInfRational, 3: orc.InfRational}[kind](a, b), po.Tr(kind, a, b))


@pytest.mark.parametrize("D", [0, 1, 2, 3])
def test_univariate_bitwise(orc, D):
    rng = np.random.default_rng(100 + D)
    N = 24
    theta = rng.standard_normal(N)
    for kind, a, b in ((1, -2.0, 3.0), (2, 0.7, 0.3), (3, 0.4, 0.9), (2, -3.0, -1.5)):
        if kind != 1 and D > 2:
            continue
        ct, pt = _pair(orc, kind, a, b)
        u = np.clip((rng.random(40) - 0.5) * 2.4, -1, 1)
        with np.errstate(all="ignore"):
            y = np.array([po.transform_from(pt, x) for x in u])
            assert (y == np.array([orc.transform_from(ct, x) for x in u])).all()
        got = orc.uni_lincomb(N, theta, y, tr=ct, D=D, ext2=True)
        for p in range(len(y)):
            want = po.uni_lincomb(N, theta.tolist(), float(y[p]), tr=pt, D=D, ext2=True)
            assert got[p].tolist() == want or (np.isnan(got[p]).all() and np.isnan(want).all()), (kind, p, y[p])


@pytest.mark.parametrize("kind", [INT, END, INT2])
def test_smolyak_bitwise(orc, kind):
    rng = np.random.default_rng(200 + kind)
    d, B, M = 3, 3, 2
    K = orc.smolyak_length(kind, d, B, M)
    theta = rng.standard_normal(K)
    ctr, ptr = zip(_pair(orc, 1, 0.0, 4.0), _pair(orc, 2, 1.0, 2.0), _pair(orc, 3, 0.0, 4.0))
    u = rng.uniform(-0.95, 0.95, (5, d))
    y = orc.transform_from_batch(list(ctr), u)
    spec = orc.deriv_spec([(1, 1), (0, 0, 2)], d)
    got_v = orc.smolyak_lincomb(kind, d, B, M, theta, y, trs=list(ctr))
    got_s = orc.smolyak_lincomb(kind, d, B, M, theta, y, trs=list(ctr), spec=spec, ext2=True)
    for p in range(len(y)):
        assert got_v[p].tolist() == po.smolyak_lincomb(kind, d, B, M, theta.tolist(), y[p].tolist(), trs=list(ptr))
        want = po.smolyak_lincomb(kind, d, B, M, theta.tolist(), y[p].tolist(), trs=list(ptr),
                                  spec=[tuple(q) for q in spec.tolist()], ext2=True)
        assert got_s[p].tolist() == want
