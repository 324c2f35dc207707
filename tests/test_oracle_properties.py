"""The reference's own property tests, ported to the oracle.

Each test names the reference test it restates.  Together with the known-answer tests these
are what pins the oracle (Julia itself cannot run in this environment).
"""
import math

import numpy as np
import pytest

from oracle import py_oracle as po

INT, END, INT2 = 0, 1, 2
GRIDS = (END, INT, INT2)


def rand_pm1(rng, n):
    # test/utilities.jl:28 — atoms on both endpoints
    return np.clip((rng.random(n) - 0.5) * 2.5, -1, 1)


def cheb_cos(x, n):          # test/utilities.jl:7 (n is 1-based)
    return math.cos((n - 1) * math.acos(x))


def cheb_cos_deriv(x, n):    # test/utilities.jl:9-19
    z = float((n - 1) ** 2)
    if x == -1:
        return -z if n % 2 == 1 else z
    if x == 1:
        return z
    t = math.acos(x)
    return (n - 1) * math.sin((n - 1) * t) / math.sin(t)


@pytest.mark.parametrize("N", range(1, 11))
def test_chebyshev_values_and_derivatives(orc, N):
    # test/test_chebyshev.jl:18-31
    rng = np.random.default_rng(N)
    xs = rand_pm1(rng, 100)
    B = orc.uni_basis_at(N, xs)[:, :, 0]                       # [N][n]
    want = np.array([[cheb_cos(x, i) for x in xs] for i in range(1, N + 1)])
    np.testing.assert_allclose(B, want, rtol=1e-8, atol=1e-12)
    theta = rng.random(N)
    out = orc.uni_lincomb(N, theta, xs, D=1)
    np.testing.assert_allclose(out[:, 0], theta @ want, rtol=1e-8, atol=1e-12)
    dwant = np.array([[cheb_cos_deriv(x, i) for x in xs] for i in range(1, N + 1)])
    np.testing.assert_allclose(out[:, 1], theta @ dwant, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("N", range(1, 11))
def test_chebyshev_grids(orc, N):
    # test/test_chebyshev.jl:33-44
    g = orc.cheb_grid(INT, N)
    assert all(abs(cheb_cos(x, N + 1)) <= 1e-14 for x in g)
    assert (np.diff(g) < 0).all()      # sinpi((N-2i+1)/2N) decreases in i (src/chebyshev.jl:91)
    if N >= 2:
        g = orc.cheb_grid(END, N)
        assert all(abs(cheb_cos_deriv(x, N)) <= 1e-13 for x in g[1:-1])
        assert g[0] == -1 and g[-1] == 1


def test_grids_are_correctly_rounded_sinpi_cospi(orc):
    # the product pins grids to correctly rounded sinpi/cospi; the oracle's
    # binary128 evaluation must agree with 240-bit mpmath at every size a Smolyak parent can have.
    for kind in GRIDS:
        for N in sorted({1, 2, 3, 5, 7, 9, 15, 17, 20, 27, 31, 33, 63, 64, 81, 127, 128, 243}):
            got = orc.cheb_grid(kind, N)
            want = [po.gridpoint(kind, N, i) for i in range(1, N + 1)]
            assert got.tolist() == want, (kind, N)


def test_hypot_is_correctly_rounded(orc):
    # Base.hypot (fma branch) is documented as correctly rounded; check the restatement against mpmath.
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.standard_normal(3000) * 10.0 ** rng.integers(-8, 8, 3000), [0.0, 1e-300, 1e300, 3.0]])
    ys = np.concatenate([rng.standard_normal(3000) * 10.0 ** rng.integers(-8, 8, 3000), [0.9, 1e-300, 1e300, 4.0]])
    for x, y in zip(xs, ys):
        assert orc.hypot(x, y) == po.hypot_cr(x, y), (x, y)


@pytest.mark.parametrize("mk", ["bl", "semi", "inf"])
def test_transformation_round_trips(orc, mk):
    # test/test_transformations.jl:12-23,37-48,72-83
    tr = {"bl": lambda: orc.BoundedLinear(1, 5), "semi": lambda: orc.SemiInfRational(3.0, 4.0),
          "inf": lambda: orc.InfRational(0.0, 1.0)}[mk]()
    rng = np.random.default_rng(11)
    for x in rand_pm1(rng, 200):
        with np.errstate(all="ignore"):
            y = orc.transform_from(tr, x)
        if mk == "bl":
            assert 1 <= y <= 5
        elif mk == "semi":
            assert (y == math.inf) if x == 1 else (3.0 <= y < math.inf)
        else:
            assert (y == math.copysign(math.inf, x)) if abs(x) == 1 else math.isfinite(y)
        assert orc.transform_to(tr, y)[0] == pytest.approx(x, rel=1e-12, abs=1e-12)


def test_derivatives_through_transformations_vs_finite_differences(orc):
    # test/test_chebyshev.jl:91-107 (atol 1e-6): BoundedLinear up to order 5, rational maps order 1
    rng = np.random.default_rng(3)
    N = 5
    for tr, D in ((orc.BoundedLinear(-2, 3), 5), (orc.SemiInfRational(0.7, 0.3), 1), (orc.InfRational(0.4, 0.9), 1)):
        theta = rng.standard_normal(N)
        for u in rng.uniform(-0.9, 0.9, 20):
            y = orc.transform_from(tr, u)
            got = orc.uni_lincomb(N, theta, np.array([y]), tr=tr, D=D)[0]
            f = lambda v: orc.uni_lincomb(N, theta, np.array([v]), tr=tr)[0, 0]
            # central differences with a step suited to each order
            import mpmath
            for r in range(0, D + 1):
                if r == 0:
                    want = f(y)
                else:
                    h = 1e-3 if r > 2 else 1e-5
                    want = float(mpmath.diff(lambda v: mpmath.mpf(f(float(v))), y, r, h=h * max(1.0, abs(y)), direction=0)) if r <= 2 else None
                if want is not None:
                    assert got[r] == pytest.approx(want, abs=1e-5 * max(1.0, abs(want)))


def test_second_derivative_extension_vs_mpmath(orc):
    # SURVEY.md Appendix A.3 (extension: the reference throws at order 2)
    for tr, ptr in ((orc.SemiInfRational(0.7, 0.3), po.Tr(po.SEMI_INF, 0.7, 0.3)),
                    (orc.SemiInfRational(-3.0, 3.0), po.Tr(po.SEMI_INF, -3.0, 3.0)),
                    (orc.InfRational(0.4, 0.9), po.Tr(po.INF, 0.4, 0.9))):
        for u in (-0.9, -0.3, 0.0, 0.2, 0.77):
            y = orc.transform_from(tr, u)
            got = orc.transform_to(tr, y, D=2, ext2=True)
            for r in range(3):
                want = float(po.transform_to_truth(ptr, y, r))
                assert got[r] == pytest.approx(want, rel=1e-12, abs=1e-14)


def test_nesting_lengths_and_nested_grids(orc):
    # test/test_smolyak_traversal.jl:21-32
    tol = math.sqrt(np.finfo(float).eps)
    for kind in GRIDS:
        for b in range(0, 6):
            nA = orc.nesting_total_length(kind, b)
            gA = orc.cheb_grid(kind, nA)
            gB = orc.cheb_grid(kind, orc.nesting_total_length(kind, b + 1))
            assert all(np.min(np.abs(gB - x)) <= tol for x in gA)
            assert sum(orc.nesting_block_length(kind, a) for a in range(b + 1)) == nA


def test_shuffle_equals_blockwise_set_differences(orc):
    # test/test_smolyak_traversal.jl:38-64
    for kind in GRIDS:
        for b in range(0, 6):
            n = orc.nesting_total_length(kind, b)
            sh = orc.grid_shuffle(kind, n)
            assert len(sh) == n
            assert sh.tolist() == po.shuffle_by_blocks(kind, b), (kind, b)


def test_smolyak_indices_equal_naive_filter(orc):
    # test/test_smolyak_traversal.jl:92-107 (bit-exact requirement), widened to B <= 4
    for kind in GRIDS:
        for B in range(0, 5):
            for M in range(0, B + 1):
                for d in range(1, 5):
                    if total := po.total_length(kind, M) ** d > 200000:
                        continue
                    got = orc.smolyak_indices(kind, d, B, M)
                    want = po.naive_indices(kind, d, B, M)
                    assert len(want) == orc.smolyak_length(kind, d, B, M) == len(got)
                    assert got.tolist() == [list(t) for t in want], (kind, B, M, d)


def test_smolyak_product_equals_prod(orc):
    # test/test_smolyak_traversal.jl:109-127 via basis_at on a value-only point
    rng = np.random.default_rng(5)
    for kind in GRIDS:
        for (d, B, M) in ((1, 2, 2), (2, 3, 2), (3, 3, 3), (4, 2, 1)):
            y = rng.uniform(-1, 1, (3, d))
            H = orc.smolyak_parent_dimension(kind, B, M)
            idx = orc.smolyak_indices(kind, d, B, M)
            Bv = orc.smolyak_basis_at(kind, d, B, M, y)[:, :, 0]           # [K][n]
            for p in range(3):
                tabs = [orc.uni_basis_at(H, y[p:p + 1, j])[:, 0, 0] for j in range(d)]
                want = np.array([np.prod([tabs[j][ix[j] - 1] for j in range(d)]) for ix in idx])
                np.testing.assert_allclose(Bv[:, p], want, rtol=1e-14, atol=0)


def test_smolyak_bilinear_fit_has_four_coefficients(orc):
    # test/test_smolyak.jl:13-30
    f = lambda x: (x[:, 0] - 3) * (x[:, 1] + 5)
    g = orc.smolyak_grid(INT, 2, 3, 3)
    Cm = orc.smolyak_basis_at(INT, 2, 3, 3, g)[:, :, 0].T                 # n×K
    theta = np.linalg.solve(Cm, f(g))
    assert int((np.abs(theta) > 1e-8).sum()) == 4
    yy = np.stack(np.meshgrid(np.linspace(-1, 1, 25), np.linspace(-1, 1, 25)), -1).reshape(-1, 2)
    np.testing.assert_allclose(orc.smolyak_lincomb(INT, 2, 3, 3, theta, yy)[:, 0], f(yy), rtol=1e-9, atol=1e-9)


def test_smolyak_grid_nesting(orc):
    # test/test_smolyak.jl:82-96 (subset of the (B, M) pairs)
    tol = math.sqrt(np.finfo(float).eps)
    for kind in GRIDS:
        for (B1, M1, B2, M2) in ((0, 0, 1, 1), (1, 1, 2, 2), (1, 2, 3, 3), (2, 3, 3, 4), (2, 2, 4, 5)):
            g1 = orc.smolyak_grid(kind, 2, B1, M1)
            g2 = orc.smolyak_grid(kind, 2, B2, M2)
            for x in g1:
                assert np.min(np.max(np.abs(g2 - x), axis=1)) <= tol


def test_transformed_basis_equalities(orc):
    # test/test_generic_api.jl:32-51,62-82: l1(transform_to(x)) == l2(x) with `==`
    rng = np.random.default_rng(9)
    N = 10
    t = orc.BoundedLinear(1.0, 2.0)
    theta = rng.standard_normal(N)
    xs = rng.random(20) + 1.0
    l2 = orc.uni_lincomb(N, theta, xs, tr=t)[:, 0]
    l1 = orc.uni_lincomb(N, theta, orc.transform_to_batch([t], xs[:, None])[:, 0, 0])[:, 0]
    assert (l1 == l2).all()
    assert (orc.cheb_grid(END, N, tr=t) == [orc.transform_from(t, x) for x in orc.cheb_grid(END, N)]).all()
    trs = [orc.BoundedLinear(1.0, 2.0), orc.SemiInfRational(0, 1)]
    K = orc.smolyak_length(INT, 2, 2, 2)
    theta = rng.standard_normal(K)
    xs = rng.random((20, 2)) + 1.0
    l2 = orc.smolyak_lincomb(INT, 2, 2, 2, theta, xs, trs=trs)
    l1 = orc.smolyak_lincomb(INT, 2, 2, 2, theta, orc.transform_to_batch(trs, xs)[:, :, 0])
    assert (l1 == l2).all()
    assert (orc.smolyak_grid(INT, 2, 2, 2, trs=trs) == orc.transform_from_batch(trs, orc.smolyak_grid(INT, 2, 2, 2))).all()


def test_svector_coefficients_passthrough(orc):
    # test/test_generic_api.jl:95-102: `≡` column-wise
    rng = np.random.default_rng(13)
    N, M = 10, 3
    tr = orc.SemiInfRational(0.0, 1.0)
    theta = rng.random((N, M))
    x = np.array([2.0, 0.3, 17.0])
    got = orc.uni_lincomb(N, theta, x, tr=tr)
    for v in range(M):
        assert (got[:, v] == orc.uni_lincomb(N, theta[:, v].copy(), x, tr=tr)[:, 0]).all()
    K = orc.smolyak_length(INT2, 3, 2, 2)
    theta = rng.standard_normal((K, 4))
    y = rng.uniform(-1, 1, (5, 3))
    got = orc.smolyak_lincomb(INT2, 3, 2, 2, theta, y)
    for v in range(4):
        assert (got[:, v] == orc.smolyak_lincomb(INT2, 3, 2, 2, theta[:, v].copy(), y)[:, 0]).all()


def test_length_checks(orc):
    # test/test_generic_api.jl:25-30
    with pytest.raises(orc.OracleError):
        orc.uni_lincomb(5, np.zeros(6), np.array([0.0]))
    with pytest.raises(orc.OracleError):
        orc.smolyak_lincomb(INT, 2, 2, 2, np.zeros(3), np.zeros((1, 2)))


def test_smolyak_gradient_vs_analytic_polynomial(orc):
    # The reference does not test multivariate derivative correctness (test/test_smolyak.jl:32-42);
    # SURVEY.md §4 take-away (iii): check against an analytic polynomial.  f = prod_j (1 + a_j y_j + c_j y_j^2)
    # lies in the span of the d=3, B=3 basis only approximately, so instead differentiate the basis
    # expansion itself with mpmath-free central differences of the oracle's own value path.
    rng = np.random.default_rng(17)
    d, B = 3, 3
    trs = [orc.BoundedLinear(0, 4), orc.SemiInfRational(1, 2), orc.InfRational(0, 4)]
    K = orc.smolyak_length(INT2, d, B, B)
    theta = rng.standard_normal(K)
    spec = orc.gradient_spec(d)
    u = rng.uniform(-0.8, 0.8, (6, d))
    y = orc.transform_from_batch(trs, u)
    got = orc.smolyak_lincomb(INT2, d, B, B, theta, y, trs=trs, spec=spec)
    val = orc.smolyak_lincomb(INT2, d, B, B, theta, y, trs=trs)
    assert (got[:, 0] == val[:, 0]).all()
    for j in range(d):
        h = 1e-6 * np.maximum(1.0, np.abs(y[:, j]))
        yp, ym = y.copy(), y.copy()
        yp[:, j] += h
        ym[:, j] -= h
        fd = (orc.smolyak_lincomb(INT2, d, B, B, theta, yp, trs=trs)[:, 0]
              - orc.smolyak_lincomb(INT2, d, B, B, theta, ym, trs=trs)[:, 0]) / (yp[:, j] - ym[:, j])
        np.testing.assert_allclose(got[:, 1 + j], fd, rtol=2e-6, atol=2e-6)
