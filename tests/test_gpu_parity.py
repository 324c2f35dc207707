"""GPU parity tests: the CUDA library (through its C ABI / the Python mirror) against the CPU oracle.

Parity bar (BASELINE.json north_star):
  * exact mode (`exact=True`, and every basis_at / collocation_matrix / transform call) must be
    BIT-IDENTICAL to the oracle, NaNs matching position for position;
  * fast mode (fma-contracted, nested Smolyak evaluation) must satisfy
        |gpu − oracle| ≤ TAU · S_r(x),   S_r(x) = Σ_k |θ_k|·|b_k^{(r)}(x)|   (oracle, absmode)
    with TAU = 1e-13 for every derivative order r, away from the
    transformation singularities (|x| ≤ 0.99 on the rational maps).
"""
import math
import os

import numpy as np
import pytest

import __graft_entry__ as ge

pytestmark = pytest.mark.gpu

sk = ge.load_package()
TAU = 1e-13
GRIDS = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}


def same(a, b):
    """bit-for-bit equality with NaNs matching"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def within(got, want, scale, tau=TAU):
    """|got − want| ≤ tau·scale element-wise (0 ≤ 0 passes where the scale vanishes)"""
    got, want, scale = np.asarray(got), np.asarray(want), np.asarray(scale)
    ok = np.abs(got - want) <= tau * scale
    if not ok.all():
        with np.errstate(all="ignore"):
            print("max scaled error", np.nanmax(np.abs(got - want) / scale))
    return bool(ok.all())


def rand_pm1(rng, shape):
    return np.clip((rng.random(shape) - 0.5) * 2.5, -1, 1)      # test/utilities.jl:28


def mk_tr(orc, kind, a, b):
    return ({1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}[kind](a, b),
            {1: orc.BoundedLinear, 2: orc.SemiInfRational, 3: orc.InfRational}[kind](a, b))


MIXED6 = [(1, 0, 4), (1, -2, 2), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3)]     # cfg 4b transformations


# ------------------------------------------------------------------------------------------ univariate
@pytest.mark.parametrize("N", [1, 2, 3, 20, 64, 128])
@pytest.mark.parametrize("D", [0, 1, 2, 3, 5])
def test_uni_bounded_linear_exact_and_fast(orc, N, D):
    rng = np.random.default_rng(1000 * N + D)
    t, ot = mk_tr(orc, 1, -2, 3)
    theta = rng.standard_normal(N)
    y = np.array([orc.transform_from(ot, u) for u in rand_pm1(rng, 1000)])
    want = orc.uni_lincomb(N, theta, y, tr=ot, D=D)
    basis = sk.Chebyshev(sk.InteriorGrid(), N) @ t
    pts = (sk.d ** D)(y) if D else y
    got = sk.linear_combination(basis, theta, pts, exact=True)
    assert same(got.reshape(want.shape), want)
    fast = sk.linear_combination(basis, theta, pts).reshape(want.shape)
    scale = orc.uni_lincomb(N, theta, y, tr=ot, D=D, absmode=True)
    assert within(fast, want, scale)


@pytest.mark.parametrize("kind,a,b", [(2, 0.7, 0.3), (3, 0.4, 0.9), (2, -3.0, -1.5)])
@pytest.mark.parametrize("D", [0, 1, 2])
def test_uni_rational_maps(orc, kind, a, b, D):
    rng = np.random.default_rng(7 + D)
    N = 128
    t, ot = mk_tr(orc, kind, a, b)
    theta = rng.standard_normal(N)
    u = np.concatenate([rng.uniform(-0.99, 0.99, 2000), [0.0]])
    y = np.array([orc.transform_from(ot, v) for v in u])
    basis = sk.Chebyshev(sk.InteriorGrid(), N) @ t
    pts = (sk.d ** D)(y) if D else y
    if D == 2:      # the reference throws (transformations.jl:266,332); the extension is opt-in
        with pytest.raises(sk.UnsupportedError):
            sk.linear_combination(basis, theta, pts)
    kw = dict(allow_rational_d2=True) if D == 2 else {}
    want = orc.uni_lincomb(N, theta, y, tr=ot, D=D, ext2=True)
    got = sk.linear_combination(basis, theta, pts, exact=True, **kw).reshape(want.shape)
    assert same(got, want)
    fast = sk.linear_combination(basis, theta, pts, **kw).reshape(want.shape)
    scale = orc.uni_lincomb(N, theta, y, tr=ot, D=D, ext2=True, absmode=True)
    assert within(fast, want, scale)


def test_uni_infinite_endpoints_exact_limits(orc):
    # test/test_chebyshev.jl:109-146 through the batched path, both modes
    N = 10
    y = np.array([math.inf, -math.inf])
    for t, minus in ((sk.SemiInfRational(2.3, 0.7), lambda i: 1.0), (sk.InfRational(0.3, 3.0), lambda i: (-1.0) ** (i + 1))):
        basis = sk.Chebyshev(sk.InteriorGrid(), N) @ t
        for i in range(1, N + 1):
            theta = np.zeros(N)
            theta[i - 1] = 1.0
            for exact in (True, False):
                out = sk.linear_combination(basis, theta, sk.d(y), exact=exact)
                assert out[0, 0] == 1 and out[0, 1] == 0
                assert out[1, 0] == minus(i) and out[1, 1] == 0


def test_uni_endpoint_atoms_and_closed_forms(orc):
    # test/test_chebyshev.jl:18-31: values/derivatives vs cos((n-1) acos x), incl. x = ±1
    rng = np.random.default_rng(5)
    for N in (5, 10):
        x = rand_pm1(rng, 500)
        theta = rng.random(N)
        out = sk.linear_combination(sk.Chebyshev(sk.EndpointGrid(), N), theta, sk.d(x))
        want = sum(theta[i] * np.cos(i * np.arccos(x)) for i in range(N))
        np.testing.assert_allclose(out[:, 0], want, rtol=1e-10, atol=1e-12)
        assert same(sk.linear_combination(sk.Chebyshev(sk.EndpointGrid(), N), theta, sk.d(x), exact=True),
                    orc.uni_lincomb(N, theta, x, D=1))


def test_uni_svector_coefficients(orc):
    # test/test_generic_api.jl:95-102
    rng = np.random.default_rng(13)
    N, M = 10, 3
    t, ot = mk_tr(orc, 2, 0.0, 1.0)
    theta = rng.random((N, M))
    x = rng.uniform(0.01, 30, 300)
    basis = sk.Chebyshev(sk.InteriorGrid(), N) @ t
    got = sk.linear_combination(basis, theta, x, exact=True)
    assert same(got, orc.uni_lincomb(N, theta, x, tr=ot))
    for v in range(M):
        assert same(got[:, v], sk.linear_combination(basis, theta[:, v].copy(), x, exact=True))
    fast = sk.linear_combination(basis, theta, x)
    np.testing.assert_allclose(fast, got, rtol=1e-12, atol=1e-13)
    theta = rng.standard_normal((64, 21))
    got = sk.linear_combination(sk.Chebyshev(sk.InteriorGrid(), 64), theta, x / 30 - 0.5, exact=True)
    assert same(got, orc.uni_lincomb(64, theta, x / 30 - 0.5))
    with pytest.raises(sk.UnsupportedError):
        sk.linear_combination(basis, theta[:10], sk.d(x))


@pytest.mark.parametrize("n", [1, 20, 127, 128, 129, 1000])
@pytest.mark.parametrize("D", [0, 1, 2])
def test_uni_basis_at_bit_identical(orc, n, D):
    # generic_api.jl:217-227; odd n exercises the unaligned (plain store) path, n >= 128 the bulk-copy path
    rng = np.random.default_rng(n + D)
    t, ot = mk_tr(orc, 1, -2, 3)
    y = rng.uniform(-2, 3, n)
    for N in (1, 20, 33):
        want = orc.uni_basis_at(N, y, tr=ot, D=D)                          # (N, n, D+1)
        got = sk.basis_at(sk.Chebyshev(sk.InteriorGrid(), N) @ t, (sk.d ** D)(y) if D else y)
        assert same(np.asarray(got).reshape(n, N, D + 1), want.transpose(1, 0, 2))


def test_collocation_default_grid_and_fit(orc):
    # test/test_generic_api.jl:6-23
    basis = sk.Chebyshev(sk.EndpointGrid(), 10)
    assert same(sk.collocation_matrix(basis), sk.collocation_matrix(basis, sk.grid(basis)))
    x = sk.grid(sk.Chebyshev(sk.EndpointGrid(), 20))
    Cm = sk.collocation_matrix(basis, x)
    assert Cm.shape == (20, 10) and np.isfinite(Cm).all()
    phi = np.linalg.lstsq(Cm, np.exp(x), rcond=None)[0]
    yy = np.linspace(-1, 1, 50)
    assert np.max(np.abs(np.exp(yy) - sk.linear_combination(basis, phi, yy))) <= 1e-9
    # cfg 1 of BASELINE.json: Chebyshev(InteriorGrid(), 20) collocation at its own grid
    b20 = sk.Chebyshev(sk.InteriorGrid(), 20)
    assert same(sk.collocation_matrix(b20), orc.uni_basis_at(20, orc.cheb_grid(0, 20))[:, :, 0].T)


def test_length_checks_and_errors():
    # test/test_generic_api.jl:25-30; generic_api.jl:99-102
    basis = sk.Chebyshev(sk.InteriorGrid(), 5)
    with pytest.raises(sk.ArgumentError, match="too many coefficients"):
        sk.linear_combination(basis, np.zeros(6), 0.0)
    with pytest.raises(sk.ArgumentError, match="not enough coefficients"):
        sk.linear_combination(basis, np.zeros(4), 0.0)
    sb = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(2), 2)
    with pytest.raises(sk.ArgumentError):
        sk.linear_combination(sb, np.zeros(sk.dimension(sb)), np.zeros(3))          # smolyak_api.jl:97
    with pytest.raises(sk.ArgumentError, match="too many coefficients"):
        sk.linear_combination(sb, np.zeros(sk.dimension(sb) + 1), np.zeros(2))
    with pytest.raises(sk.UnsupportedError):                                         # transformations.jl:266
        sk.linear_combination(sb @ sk.coordinate_transformations(sk.SemiInfRational(0, 1), sk.BoundedLinear(0, 1)),
                              np.zeros(sk.dimension(sb)), sk.partial(2)(np.ones(2)))
    assert sk.linear_combination(basis, np.ones(5), 0.3).shape == ()                 # single point -> scalar


def test_transformed_basis_equalities(orc):
    # test/test_generic_api.jl:32-51,62-82: l1(transform_to(x)) == l2(x) == l3(x) with `==`
    rng = np.random.default_rng(9)
    basis = sk.Chebyshev(sk.EndpointGrid(), 10)
    t = sk.BoundedLinear(1.0, 2.0)
    theta = rng.standard_normal(10)
    x = rng.random(20) + 1.0
    l1 = sk.linear_combination(basis, theta)
    l2 = sk.linear_combination(basis @ t, theta)
    l3 = sk.linear_combination(basis, theta) @ t
    assert same(l1(sk.transform_to(sk.domain(basis), t, x), exact=True), l2(x, exact=True))
    assert same(l2(x, exact=True), l3(x, exact=True))
    b0 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(2, 2), 2)
    ct = sk.coordinate_transformations(sk.BoundedLinear(1.0, 2.0), sk.SemiInfRational(0, 1))
    theta = rng.standard_normal(sk.dimension(b0))
    x = rng.random((20, 2)) + 1.0
    l1 = sk.linear_combination(b0, theta)
    l2 = sk.linear_combination(b0 @ ct, theta)
    assert same(l1(sk.transform_to(sk.domain(b0), ct, x), exact=True), l2(x, exact=True))
    assert same(sk.grid(b0 @ ct), sk.transform_from(sk.domain(b0), ct, sk.grid(b0)))


def test_batched_transforms_bit_identical(orc):
    rng = np.random.default_rng(21)
    pairs = [mk_tr(orc, *k) for k in MIXED6]
    ct = sk.coordinate_transformations(*[p[0] for p in pairs])
    otr = [p[1] for p in pairs]
    u = rand_pm1(rng, (4000, 6))
    with np.errstate(all="ignore"):
        y = sk.transform_from(None, ct, u)
        assert same(y, orc.transform_from_batch(otr, u))
        y = np.where(np.isnan(y), 0.5, y)
        assert same(sk.transform_to(None, ct, y), orc.transform_to_batch(otr, y)[:, :, 0])
        assert same(sk.transform_to(None, ct, sk.d(y)), orc.transform_to_batch(otr, y, D=1))
    # docstring known answer, transformations.jl:88-104
    ct2 = sk.coordinate_transformations(sk.BoundedLinear(0, 2), sk.SemiInfRational(2, 3))
    x = sk.transform_from(None, ct2, np.array([0.4, 0.5]))
    assert x.tolist() == [1.4, 11.0]
    assert sk.transform_to(None, ct2, x).tolist() == [0.3999999999999999, 0.5]
    # ±Inf limits, test/test_transformations.jl:50-57,85-92
    for t, vals in ((sk.SemiInfRational(3.0, 4.0), (1.0, 1.0)), (sk.InfRational(0.0, 1.0), (1.0, -1.0))):
        out = sk.transform_to(None, t, sk.d(np.array([math.inf, -math.inf])))
        assert out[:, 0].tolist() == list(vals) and out[:, 1].tolist() == [0.0, 0.0]


# ------------------------------------------------------------------------------------------ Smolyak
SMOLYAK_CASES = [(0, 2, 3, 3), (1, 3, 3, 2), (2, 3, 4, 4), (2, 6, 2, 2), (1, 4, 3, 3), (0, 3, 2, 1), (2, 1, 4, 4), (2, 2, 0, 0)]


@pytest.mark.parametrize("code,d,B,M", SMOLYAK_CASES)
def test_smolyak_exact_bit_identical(orc, code, d, B, M):
    rng = np.random.default_rng(code * 100 + d * 10 + B)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B, M), d)
    K = sk.dimension(basis)
    theta = rng.standard_normal(K)
    x = rand_pm1(rng, (300, d))
    assert same(sk.linear_combination(basis, theta, x, exact=True), orc.smolyak_lincomb(code, d, B, M, theta, x)[:, 0])
    spec, ospec = sk.gradient(d), orc.gradient_spec(d)
    assert same(sk.linear_combination(basis, theta, spec(x), exact=True),
                orc.smolyak_lincomb(code, d, B, M, theta, x, spec=ospec))
    if d >= 2:
        spec = sk.partial(1, 1) | sk.partial(0, 2)
        ospec = orc.deriv_spec([(1, 1), (0, 2)], d)
        assert same(sk.linear_combination(basis, theta, spec(x), exact=True),
                    orc.smolyak_lincomb(code, d, B, M, theta, x, spec=ospec))


def test_smolyak_exact_mode_has_no_table_limit(orc):
    # round 1 refused bit-identical evaluation of bases whose full tables (d·ℓ(B)·(order+1) doubles per point for 32
    # points) exceed shared memory — InteriorGrid, B = 4: ℓ = 81, six dimensions, gradient.  The chunked-table kernel
    # serves them in the reference's order; the fast path stays within tolerance
    code, d, B = 0, 6, 4
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
    rng = np.random.default_rng(99)
    theta = rng.standard_normal(sk.dimension(basis))
    x = rand_pm1(rng, (50, d))
    ospec = orc.gradient_spec(d)
    want = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec)
    assert same(sk.linear_combination(basis, theta, sk.gradient(d)(x), exact=True), want)
    scale = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec, absmode=True)
    assert within(sk.linear_combination(basis, theta, sk.gradient(d)(x)), want, scale)
    assert same(sk.linear_combination(basis, theta, x, exact=True), orc.smolyak_lincomb(code, d, B, B, theta, x)[:, 0])


@pytest.mark.parametrize("code", [0, 1, 2])
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6])
def test_smolyak_single_budget_leaf_kernels(orc, code, d):
    # bases that are one leaf (d <= 6) run budget-specific kernels for B = 1..5 (sk_leaf_impl.cuh: leaf_single_built_c);
    # every one of them, values and gradient, ragged n and n = 1
    for B in (1, 2, 3, 4, 5):
        basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
        K = sk.dimension(basis)
        if K > 12000:
            continue
        rng = np.random.default_rng(7000 + code * 100 + d * 10 + B)
        theta = rng.standard_normal(K)
        for n in (1, 777 if K <= 3000 else 259):
            x = rand_pm1(rng, (n, d))
            want = orc.smolyak_lincomb(code, d, B, B, theta, x)[:, 0]
            scale = orc.smolyak_lincomb(code, d, B, B, theta, x, absmode=True)[:, 0]
            assert within(sk.linear_combination(basis, theta, x), want, scale)
            ospec = orc.gradient_spec(d)
            want = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec)
            scale = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec, absmode=True)
            assert within(sk.linear_combination(basis, theta, sk.gradient(d)(x)), want, scale)


@pytest.mark.parametrize("code,d,B,M", SMOLYAK_CASES + [(2, 6, 4, 4), (1, 8, 3, 3), (0, 5, 3, 3), (2, 10, 3, 3), (2, 4, 5, 5),
                                                       (2, 12, 2, 2), (0, 13, 2, 2), (1, 16, 2, 2), (2, 16, 2, 2), (2, 11, 2, 2), (2, 14, 2, 2), (1, 15, 2, 2),
                                                       (2, 7, 3, 3), (1, 9, 3, 3), (1, 10, 3, 3), (2, 11, 3, 3), (2, 12, 3, 3), (1, 12, 3, 3), (2, 7, 4, 4), (1, 8, 4, 4), (2, 8, 4, 4),
                                                       (2, 9, 4, 4), (1, 9, 4, 4), (2, 10, 4, 4), (1, 10, 4, 4)])      # budget 4 at d = 9, 10: looped deep leaves
def test_smolyak_fast_within_tolerance(orc, code, d, B, M):
    rng = np.random.default_rng(code * 100 + d * 10 + B + 1)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B, M), d)
    assert sk.get_plan(basis).info.fast_path >= 1
    K = sk.dimension(basis)
    theta = rng.standard_normal(K)
    n = 2000 if K < 5000 else 300 if K < 9000 else 131
    x = rand_pm1(rng, (n, d))
    want = orc.smolyak_lincomb(code, d, B, M, theta, x)[:, 0]
    scale = orc.smolyak_lincomb(code, d, B, M, theta, x, absmode=True)[:, 0]
    got = sk.linear_combination(basis, theta, x)
    assert within(got, want, scale)
    ospec = orc.gradient_spec(d)
    want = orc.smolyak_lincomb(code, d, B, M, theta, x, spec=ospec)
    scale = orc.smolyak_lincomb(code, d, B, M, theta, x, spec=ospec, absmode=True)
    got = sk.linear_combination(basis, theta, sk.gradient(d)(x))
    assert got.shape == (n, d + 1)
    assert within(got, want, scale)


def test_smolyak_headline_config_with_transformations(orc):
    # BASELINE cfg 4b: d=6, B=4, InteriorGrid2, mixed transformations, value + gradient
    rng = np.random.default_rng(104)
    d, B = 6, 4
    pairs = [mk_tr(orc, *k) for k in MIXED6]
    ct = sk.coordinate_transformations(*[p[0] for p in pairs])
    otr = [p[1] for p in pairs]
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d) @ ct
    K = sk.dimension(basis)
    assert K == 2561
    theta = rng.standard_normal(K)
    y = orc.transform_from_batch(otr, rng.uniform(-0.99, 0.99, (1500, d)))
    ospec = orc.gradient_spec(d)
    want = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, spec=ospec)
    scale = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, spec=ospec, absmode=True)
    assert same(sk.linear_combination(basis, theta, sk.gradient(d)(y), exact=True), want)
    got = sk.linear_combination(basis, theta, sk.gradient(d)(y))
    assert within(got, want, scale)
    info = sk.get_plan(basis, 0, sk.gradient(d)).info
    assert info.fast_path >= 1 and info.K == 2561 and info.H == 31 and info.n_runs == 1471 and info.E == 7
    assert info.flops_per_point_direct == 126569.0                        # SURVEY.md §8(d) F_direct
    # ±Inf on the rational coordinates: finite results, value/gradient limits exact in both modes
    yinf = y[:8].copy()
    yinf[:, 2] = math.inf
    yinf[:, 3] = -math.inf
    w = orc.smolyak_lincomb(2, d, B, B, theta, yinf, trs=otr, spec=ospec)
    assert same(sk.linear_combination(basis, theta, sk.gradient(d)(yinf), exact=True), w)
    g = sk.linear_combination(basis, theta, sk.gradient(d)(yinf))
    assert np.isfinite(g).all() and (g[:, 3] == 0).all() and (g[:, 4] == 0).all()


def test_smolyak_gradient_correctness_vs_analytic(orc):
    # the reference never checks multivariate derivative values (test/test_smolyak.jl:32-42): do it here.
    # Fit f(x) = (x1 - 3)(x2 + 5) exactly (test/test_smolyak.jl:13-30), then compare the gradient analytically.
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(3), 2)
    g = sk.grid(basis)
    Cm = sk.collocation_matrix(basis, g)
    theta = np.linalg.solve(Cm, (g[:, 0] - 3) * (g[:, 1] + 5))
    assert int((np.abs(theta) > 1e-8).sum()) == 4
    yy = np.stack(np.meshgrid(np.linspace(-1, 1, 100), np.linspace(-1, 1, 100)), -1).reshape(-1, 2)
    out = sk.linear_combination(basis, theta, sk.gradient(2)(yy))
    np.testing.assert_allclose(out[:, 0], (yy[:, 0] - 3) * (yy[:, 1] + 5), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(out[:, 1], yy[:, 1] + 5, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out[:, 2], yy[:, 0] - 3, rtol=1e-10, atol=1e-10)
    full = sk.linear_combination(basis, theta, sk.partial(1, 1)(yy), exact=True)       # ∂(1,1): [f, ∂1, ∂2, ∂1∂2]
    np.testing.assert_allclose(full[:, 3], 1.0, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 200])
def test_smolyak_basis_at_bit_identical(orc, n):
    rng = np.random.default_rng(n)
    for code, d, B, M in [(2, 3, 3, 3), (0, 2, 3, 3), (1, 4, 2, 2)]:
        basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B, M), d)
        x = rand_pm1(rng, (n, d))
        K = sk.dimension(basis)
        want = orc.smolyak_basis_at(code, d, B, M, x)                                  # (K, n, 1)
        assert same(np.asarray(sk.basis_at(basis, x)), want[:, :, 0].T)
        ospec = orc.gradient_spec(d)
        want = orc.smolyak_basis_at(code, d, B, M, x, spec=ospec)
        got = np.asarray(sk.basis_at(basis, sk.gradient(d)(x)))
        assert got.shape == (n, K, d + 1) and same(got, want.transpose(1, 0, 2))


def test_smolyak_basis_at_large_tables_bit_identical(orc):
    """Bases whose full Chebyshev tables exceeded the round-1 kernel's shared memory (d·ℓ(B)·(order+1) > 896): the
    chunked-table kernel has no such limit (generic_api.jl:217-227 has none either).  InteriorGrid d=6 B=4 with a
    gradient (ℓ = 81: 972 doubles per point), InteriorGrid2 d=10 B=3 gradient, a general ∂ spec with order 2, and a
    leading dimension larger than n through the raw C ABI."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(77)
    for code, d, B, n in [(0, 6, 4, 37), (2, 10, 3, 70), (0, 3, 5, 33)]:
        basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
        K = sk.dimension(basis)
        x = rand_pm1(rng, (n, d))
        want = orc.smolyak_basis_at(code, d, B, B, x)
        assert same(np.asarray(sk.basis_at(basis, x)), want[:, :, 0].T)
        want = orc.smolyak_basis_at(code, d, B, B, x, spec=orc.gradient_spec(d))
        got = np.asarray(sk.basis_at(basis, sk.gradient(d)(x)))
        assert got.shape == (n, K, d + 1) and same(got, want.transpose(1, 0, 2))
    # general specification with second order: ∂(2) ∪ ∂(0,1,1) in d = 3
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 3)
    spec = sk.partial(2) | sk.partial(0, 1, 1)
    x = rand_pm1(rng, (45, 3))
    table, _ = spec.expansion(3)
    want = orc.smolyak_basis_at(2, 3, 4, 4, x, spec=np.asarray(table, dtype=np.int32))
    got = np.asarray(sk.basis_at(basis, spec(x)))
    assert same(got, want.transpose(1, 0, 2))
    # ld > n (odd pitch: the tensor map cannot express it, plain stores) and ld > n with an even pitch (TMA)
    L = sk._lib
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3), 4)
    K = sk.dimension(basis)
    plan = sk.get_plan(basis, 0, None, 0)
    x = rand_pm1(rng, (50, 4))
    want = orc.smolyak_basis_at(2, 4, 3, 3, x)[:, :, 0]                                 # (K, n)
    xd = torch.as_tensor(x).cuda()
    for ld in (53, 64):
        out = torch.full((K, ld), -7.0, dtype=torch.float64, device="cuda")
        L.check(L.lib.sk_basis_at(plan.handle, C.c_void_p(xd.data_ptr()), 50, C.c_void_p(out.data_ptr()), ld, L.SK_MEM_DEVICE, None))
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        assert same(o[:, :50], want) and (o[:, 50:] == -7.0).all()


def test_smolyak_basis_at_second_order_request_on_nine_dimensional_interior_grid(orc):
    # regression (parity sweep, round 2): the chunk-size floor counted d table rows per term instead of min(d, B) and asked
    # for more shared memory than an SM has — InteriorGrid d = 9, 10 with B = 3 and a second-order request was refused
    rng = np.random.default_rng(910)
    for d, parts in ((9, [(0, 0, 0, 0, 0, 0, 0, 0, 2)]), (10, [(2,), (0, 0, 0, 1)])):
        basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(3), d)
        spec = sk.partial(*parts[0])
        for q in parts[1:]:
            spec = spec | sk.partial(*q)
        x = rand_pm1(rng, (40, d))
        want = orc.smolyak_basis_at(0, d, 3, 3, x, spec=orc.deriv_spec(parts, d))      # (K, n, E)
        assert same(sk.basis_at(basis, spec(x)), want.transpose(1, 0, 2))


def test_smolyak_collocation_at_own_grid(orc):
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3), 3)
    Cm = sk.collocation_matrix(basis)
    K = sk.dimension(basis)
    assert Cm.shape == (K, K)
    assert same(Cm, orc.smolyak_basis_at(2, 3, 3, 3, orc.smolyak_grid(2, 3, 3, 3))[:, :, 0].T)
    assert np.linalg.cond(Cm) < 1e6


def test_smolyak_svector_coefficients(orc):
    rng = np.random.default_rng(33)
    d, B = 4, 3
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    theta = rng.standard_normal((K, 11))
    x = rng.uniform(-1, 1, (100, d))
    got = sk.linear_combination(basis, theta, x, exact=True)
    assert same(got, orc.smolyak_lincomb(2, d, B, B, theta, x))
    for v in (0, 10):
        assert same(got[:, v], sk.linear_combination(basis, theta[:, v].copy(), x, exact=True))


# ---- general ∂ specifications through the fast general nested kernel (fast_path 3)
# Requests whose components all have total order <= 2 on a basis that is one leaf (d <= 6) take the second-order-jet
# leaves (fast_path 4, sk_jet_impl.cuh); everything else the general nested kernel (fast_path 3).
@pytest.mark.parametrize("code,d,B,parts,path", [
    (2, 2, 3, [(1, 1)], 4),                                  # test/test_smolyak.jl:39: ∂(1, 1) → [f, ∂1, ∂2, ∂1∂2]
    (0, 3, 3, [(1, 1), (0, 2)], 4),
    (1, 4, 3, [(2,), (0, 2), (0, 0, 2), (0, 0, 0, 2)], 4),   # value + diagonal of the Hessian
    (2, 6, 4, [(1, 1), (0, 1, 1), (0, 0, 0, 2, 1), (1, 0, 0, 0, 0, 1)], 3),   # a third-order component
    (2, 3, 4, [(2, 2, 2)], 3),                               # every order up to 2 in every coordinate: 27 components
    (2, 1, 4, [(2,)], 4),
    (2, 8, 3, [(0, 0, 0, 0, 0, 0, 1, 1), (2,)], 3),          # more than six dimensions
    (0, 6, 4, [(1, 1), (0, 0, 2)], 4),                       # 4 361 terms: the largest jet leaf
    (2, 6, 5, [(1, 1), (0, 0, 2)], 3),                       # 10 625 terms: θ does not fit next to the shared tables
])
def test_smolyak_general_spec_fast(orc, code, d, B, parts, path):
    rng = np.random.default_rng(4242 + d + B)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
    spec = sk.partial(*parts[0])
    for q in parts[1:]:
        spec = spec | sk.partial(*q)
    ospec = orc.deriv_spec(parts, d)
    assert sk.get_plan(basis, 0, spec).info.fast_path == path
    theta = rng.standard_normal(sk.dimension(basis))
    x = rand_pm1(rng, (333, d))
    want = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec)
    scale = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec, absmode=True)
    got = sk.linear_combination(basis, theta, spec(x))
    assert got.shape == want.shape and within(got, want, scale)
    assert same(sk.linear_combination(basis, theta, spec(x), exact=True), want)


def hessian_parts(d, diag_only=False):
    """∂ requests whose union is the Hessian (with its lower orders: value and gradient) or its diagonal"""
    parts = []
    for n in range(d):
        for m in range(n + 1):
            if diag_only and m != n:
                continue
            o = [0] * (n + 1)
            o[m] += 1
            o[n] += 1
            parts.append(tuple(o))
    return parts


@pytest.mark.parametrize("code,d,B", [
    (0, 1, 4), (0, 2, 4), (0, 3, 3), (0, 4, 4), (0, 5, 3), (0, 6, 3), (0, 5, 4), (0, 6, 4),
    (1, 1, 5), (1, 2, 5), (1, 3, 4), (1, 4, 2), (1, 5, 4), (1, 6, 4),
    (2, 1, 5), (2, 2, 5), (2, 3, 5), (2, 4, 4), (2, 5, 4), (2, 6, 4), (2, 6, 1), (2, 4, 5),
])
@pytest.mark.parametrize("diag", [False, True])
def test_smolyak_second_order_jet_leaves(orc, code, d, B, diag):
    # full Hessian jets (value, gradient, all second partials: 28 components at d = 6) and their diagonal variants
    # through smolyak_jet_kernel, against the oracle's direct evaluation (derivatives.jl:502-511)
    rng = np.random.default_rng(7000 + 100 * code + 10 * d + B)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
    parts = hessian_parts(d, diag)
    spec = sk.partial(*parts[0])
    for q in parts[1:]:
        spec = spec | sk.partial(*q)
    ospec = orc.deriv_spec(parts, d)
    info = sk.get_plan(basis, 0, spec).info
    assert info.fast_path == 4 and info.E == (2 * d + 1 if diag else (d + 1) * (d + 2) // 2)
    theta = rng.standard_normal(sk.dimension(basis))
    x = rand_pm1(rng, (161, d))
    want = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec)
    scale = orc.smolyak_lincomb(code, d, B, B, theta, x, spec=ospec, absmode=True)
    got = sk.linear_combination(basis, theta, spec(x))
    assert got.shape == want.shape and within(got, want, scale)


def test_smolyak_second_order_jet_subsets_and_transformations(orc):
    # a request that names only some second partials stores only those columns, in the request's canonical order; mixed
    # transformations (second derivatives through the rational maps are the opt-in extension)
    rng = np.random.default_rng(7777)
    d, B = 6, 4
    trs = [mk_tr(orc, k, a, b) for k, a, b in MIXED6]
    ct = sk.coordinate_transformations(*[t[0] for t in trs])
    otr = [t[1] for t in trs]
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    theta = rng.standard_normal(sk.dimension(basis))
    y = sk.transform_from(basis, ct, rng.uniform(-0.95, 0.95, (150, d)))
    for parts in ([(1, 0, 0, 0, 0, 1)], [(0, 0, 2), (0, 1, 0, 1)], [(2,), (0, 0, 0, 0, 0, 2)], hessian_parts(d)):
        spec = sk.partial(*parts[0])
        for q in parts[1:]:
            spec = spec | sk.partial(*q)
        ospec = orc.deriv_spec(parts, d)
        assert sk.get_plan(basis @ ct, 0, spec, allow_rational_d2=True).info.fast_path == 4
        want = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, spec=ospec, ext2=True)
        scale = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, spec=ospec, ext2=True, absmode=True)
        got = sk.linear_combination(basis @ ct, theta, spec(y), allow_rational_d2=True)
        assert got.shape == want.shape and within(got, want, scale)
        assert same(sk.linear_combination(basis @ ct, theta, spec(y), allow_rational_d2=True, exact=True), want)


@pytest.mark.parametrize("n", [1, 31, 255, 257, 1000])
def test_smolyak_second_order_jet_ragged_batches(orc, n):
    # partial CTAs (256 points each), a single point; the request's columns must be the only memory written
    rng = np.random.default_rng(n)
    d, B = 3, 3
    basis = sk.smolyak_basis(sk.Chebyshev, sk.EndpointGrid(), sk.SmolyakParameters(B), d)
    parts = [(0, 1, 1), (2,)]
    spec = sk.partial(*parts[0]) | sk.partial(*parts[1])
    ospec = orc.deriv_spec(parts, d)
    assert sk.get_plan(basis, 0, spec).info.fast_path == 4
    theta = rng.standard_normal(sk.dimension(basis))
    x = rand_pm1(rng, (n, d))
    want = orc.smolyak_lincomb(1, d, B, B, theta, x, spec=ospec)
    scale = orc.smolyak_lincomb(1, d, B, B, theta, x, spec=ospec, absmode=True)
    got = sk.linear_combination(basis, theta, spec(x))
    assert got.shape == want.shape and within(got, want, scale)


def test_smolyak_general_spec_fast_transformed(orc):
    # mixed transformations; second derivatives through the rational maps are the opt-in extension
    rng = np.random.default_rng(99)
    d, B = 4, 3
    trs = [mk_tr(orc, 1, 0, 4), mk_tr(orc, 2, 1, 2), mk_tr(orc, 3, 0, 4), mk_tr(orc, 1, -1, 2)]
    ct = sk.coordinate_transformations(*[t[0] for t in trs])
    otr = [t[1] for t in trs]
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    theta = rng.standard_normal(sk.dimension(basis))
    y = sk.transform_from(basis, ct, rng.uniform(-0.95, 0.95, (200, d)))
    for parts, ext in (([(1, 1), (0, 0, 1, 1)], False), ([(2,), (0, 1, 1), (0, 0, 0, 2)], False), ([(1, 2), (0, 0, 2)], True)):
        spec = sk.partial(*parts[0])
        for q in parts[1:]:
            spec = spec | sk.partial(*q)
        ospec = orc.deriv_spec(parts, d)
        want = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, spec=ospec, ext2=ext)
        scale = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, spec=ospec, ext2=ext, absmode=True)
        got = sk.linear_combination(basis @ ct, theta, spec(y), allow_rational_d2=ext)
        assert within(got, want, scale)


# ---- K4: coefficient block (SVector coefficients) through the DMMA kernel, fast mode
@pytest.mark.parametrize("code,d,B,nvec,n", [
    (2, 4, 3, 11, 100),      # odd nvec: generic Θ loader, scalar epilogue stores; partial point tile
    (2, 4, 3, 2, 300),       # smallest block, 8-wide tile
    (0, 3, 2, 12, 257),      # InteriorGrid, 16-wide tile
    (1, 5, 3, 24, 129),      # EndpointGrid, 32-wide tile
    (2, 6, 4, 40, 200),      # 64-wide tile
    (2, 6, 4, 100, 130),     # 128-wide tile, partial vector tile
    (2, 3, 5, 256, 70),      # 256-wide tile, high budget in few dimensions
    (2, 6, 4, 300, 65),      # two vector tiles (256 + 44)
    (1, 12, 2, 16, 90),      # more than 10 dimensions
    (2, 1, 4, 8, 50),        # one-dimensional Smolyak basis
])
def test_smolyak_block_fast(orc, code, d, B, nvec, n):
    rng = np.random.default_rng(7000 + 13 * d + nvec)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    theta = rng.standard_normal((K, nvec))
    x = rand_pm1(rng, (n, d))
    want = orc.smolyak_lincomb(code, d, B, B, theta, x)
    scale = orc.smolyak_lincomb(code, d, B, B, theta, x, absmode=True)
    got = sk.linear_combination(basis, theta, x)
    assert got.shape == (n, nvec) and within(got, want, scale)
    # device-resident, and consistent with the bit-exact direct kernel
    import torch
    gdev = sk.linear_combination(basis, torch.as_tensor(theta).cuda(), torch.as_tensor(x).cuda())
    assert same(gdev.cpu().numpy(), got)
    assert same(sk.linear_combination(basis, theta, x, exact=True), want)


@pytest.mark.parametrize("code,d,B", [(0, 1, 4), (0, 3, 3), (0, 4, 4), (0, 6, 3), (1, 2, 5), (1, 4, 3), (1, 5, 4), (1, 6, 4),
                                      (2, 1, 5), (2, 2, 5), (2, 3, 4), (2, 4, 5), (2, 5, 4), (2, 6, 4), (2, 6, 2)])
def test_smolyak_few_coefficient_vectors(orc, code, d, B):
    # SVector coefficients with 2..9 components on one-leaf bases: the four- and two-vector leaf kernels plus the one-vector
    # kernel for an odd remainder (sk_vec_impl.cuh); against the oracle's generic_api.jl:114-128 with SVector θ
    rng = np.random.default_rng(8800 + 100 * code + 10 * d + B)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    for nvec, n in ((2, 300), (3, 257), (4, 1), (5, 255), (7, 64), (9, 513)):
        theta = rng.standard_normal((K, nvec))
        x = rand_pm1(rng, (n, d))
        want = orc.smolyak_lincomb(code, d, B, B, theta, x)
        scale = orc.smolyak_lincomb(code, d, B, B, theta, x, absmode=True)
        got = sk.linear_combination(basis, theta, x)
        assert got.shape == (n, nvec) and within(got, want, scale)
        # every column equals the one-vector evaluation within the same tolerance (different summation grouping is not allowed
        # to matter beyond it)
        one = sk.linear_combination(basis, theta[:, nvec - 1].copy(), x)
        assert within(got[:, nvec - 1], one, scale[:, nvec - 1], tau=2 * TAU)


@pytest.mark.parametrize("code,d,B", [(2, 6, 4), (2, 5, 4), (2, 3, 4), (1, 6, 4), (1, 4, 3), (0, 6, 3), (0, 4, 4), (2, 6, 2), (2, 4, 5), (1, 3, 5),
                                      (2, 7, 3), (1, 8, 3), (2, 8, 2)])
def test_smolyak_values_two_points_per_thread(orc, code, d, B):
    # batches of 2^16 points and more take the two-points-per-thread values kernel (sk_pts_impl.cuh): ragged sizes (last tile
    # partial, odd), checked against the oracle on a sample of rows and against the bit-identical mode on all of them
    rng = np.random.default_rng(9100 + 100 * code + 10 * d + B)
    basis = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
    theta = rng.standard_normal(sk.dimension(basis))
    for n in (65536, 70001):
        x = rand_pm1(rng, (n, d))
        got = sk.linear_combination(basis, theta, x)
        rows = np.r_[0:150, n // 2:n // 2 + 150, n - 150:n]
        want = orc.smolyak_lincomb(code, d, B, B, theta, x[rows])[:, 0]
        scale = orc.smolyak_lincomb(code, d, B, B, theta, x[rows], absmode=True)[:, 0]
        assert got.shape == (n,) and within(got[rows], want, scale)
        exact = sk.linear_combination(basis, theta, x, exact=True)
        assert np.max(np.abs(got - exact)) <= 1e-12 * max(1.0, np.max(np.abs(exact)))


def test_smolyak_few_coefficient_vectors_transformed(orc):
    rng = np.random.default_rng(8899)
    d, B = 6, 4
    trs = [mk_tr(orc, k, a, b) for k, a, b in MIXED6]
    ct = sk.coordinate_transformations(*[t[0] for t in trs])
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    y = sk.transform_from(basis, ct, rng.uniform(-0.95, 0.95, (700, d)))
    for nvec in (2, 4, 6):
        theta = rng.standard_normal((sk.dimension(basis), nvec))
        want = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=[t[1] for t in trs])
        scale = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=[t[1] for t in trs], absmode=True)
        assert within(sk.linear_combination(basis @ ct, theta, y), want, scale)
        assert same(sk.linear_combination(basis @ ct, theta, y, exact=True), want)


def test_smolyak_block_transformed_cfg5_shape(orc):
    # cfg 5 geometry (d = 10, B = 5, K = 77 505, 256 coefficient vectors) on a handful of points, with
    # coordinate transformations mixed in; Θ is 159 MB
    rng = np.random.default_rng(505)
    d, B, nvec, n = 10, 5, 256, 70
    kinds = [(1, -1, 1), (1, 0, 4), (2, 1, 2), (3, 0, 4), (1, -1, 2), (2, -3, 3), (1, -2, 2), (3, 1, 1), (1, 0, 1), (2, 0.5, 1)]
    trs = [mk_tr(orc, *k) for k in kinds]
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    ct = sk.coordinate_transformations(*[t[0] for t in trs])
    K = sk.dimension(basis)
    assert K == 77505
    theta = rng.standard_normal((K, nvec))
    u = rng.uniform(-0.99, 0.99, (n, d))
    y = sk.transform_from(basis, ct, u)
    otr = [t[1] for t in trs]
    want = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr)
    scale = orc.smolyak_lincomb(2, d, B, B, theta, y, trs=otr, absmode=True)
    got = sk.linear_combination(basis @ ct, theta, y)
    assert within(got, want, scale)
    # linearity in Θ (size-independent property): columns scaled and summed
    th2 = np.ascontiguousarray(theta[:, :8] @ rng.standard_normal((8, 8)))
    g8 = sk.linear_combination(basis @ ct, np.ascontiguousarray(theta[:, :8]), y)
    w = np.linalg.solve(theta[:8, :8], th2[:8, :8])          # th2 = theta[:, :8] @ w
    np.testing.assert_allclose(sk.linear_combination(basis @ ct, th2, y), g8 @ w, rtol=1e-9, atol=1e-9 * np.abs(scale[:, :8]).max())


# ---- §8(f) rank 1: fit on the basis' own grid, θ = collocation_matrix(basis) \ y, then evaluate
def test_collocation_solve_smolyak_roundtrip(orc):
    import torch
    rng = np.random.default_rng(808)
    trs = [mk_tr(orc, 1, 0, 4), mk_tr(orc, 2, 1, 2), mk_tr(orc, 3, 0, 4)]
    ct = sk.coordinate_transformations(*[t[0] for t in trs])
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 3) @ ct
    K = sk.dimension(basis)
    g = np.asarray(sk.grid(basis))
    f = lambda y: np.stack([np.exp(-y[:, 0] / 4) * (y[:, 1] - 1) / (y[:, 1] + 1) + np.tanh(y[:, 2] / 4),
                            1.0 / (1.0 + y[:, 0]) + 0.3 * np.tanh(y[:, 2] / 3) ** 2,
                            np.cos(y[:, 0]) / (y[:, 1] + 2)], axis=1)
    Y = f(g)                                                    # three functions at once: (K, 3)
    theta = sk.collocation_solve(basis, Y)
    assert theta.shape == (K, 3)
    C = sk.collocation_matrix(basis)                            # bit-identical to the oracle (tested above)
    want = np.linalg.solve(C, Y)                                # LAPACK getrf/getrs = Julia's `\` on a square matrix
    tol = 50 * np.linalg.cond(C) * np.finfo(float).eps
    assert np.max(np.abs(theta - want)) <= tol * np.max(np.abs(want))
    # interpolation property (test/test_smolyak.jl:20-30): the fit reproduces the values at the grid
    back = sk.linear_combination(basis, theta, g)
    np.testing.assert_allclose(back, Y, rtol=0, atol=1e-9 * np.abs(Y).max())
    # one right-hand side, device-resident
    th1 = sk.collocation_solve(basis, torch.as_tensor(Y[:, 0].copy()).cuda())
    assert th1.is_cuda and np.max(np.abs(th1.cpu().numpy() - want[:, 0])) <= tol * np.max(np.abs(want))


def test_collocation_solve_univariate_and_errors(orc):
    b = sk.Chebyshev(sk.EndpointGrid(), 20) @ sk.BoundedLinear(-2, 3)
    g = np.asarray(sk.grid(b))
    theta = sk.collocation_solve(b, np.exp(g / 3))
    x = np.linspace(-2, 3, 101)
    np.testing.assert_allclose(sk.linear_combination(b, theta, x), np.exp(x / 3), atol=1e-13, rtol=1e-12)
    with pytest.raises(sk.ArgumentError):
        sk.collocation_solve(b, np.ones(19))


def test_augmented_coefficients_evaluate_equally(orc):
    # test/test_smolyak.jl:70-79 and test/test_chebyshev.jl:46-54: same function on the refined basis
    rng = np.random.default_rng(919)
    b1 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3, 2), 4)
    b2 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4, 4), 4)
    th1 = rng.standard_normal(sk.dimension(b1))
    th2 = sk.augment_coefficients(b1, b2, th1)
    x = rand_pm1(rng, (500, 4))
    assert same(sk.linear_combination(b1, th1, x, exact=True), orc.smolyak_lincomb(2, 4, 3, 2, th1, x)[:, 0])
    scale = orc.smolyak_lincomb(2, 4, 3, 2, th1, x, spec=orc.gradient_spec(4), absmode=True)
    assert within(sk.linear_combination(b2, th2, sk.gradient(4)(x)), sk.linear_combination(b1, th1, sk.gradient(4)(x)), scale, 2 * TAU)
    t = sk.BoundedLinear(-2, 3)
    c1, c2 = sk.Chebyshev(sk.InteriorGrid(), 7) @ t, sk.Chebyshev(sk.EndpointGrid(), 12) @ t
    th = rng.standard_normal(7)
    y = rng.uniform(-2, 3, 300)
    np.testing.assert_allclose(sk.linear_combination(c2, sk.augment_coefficients(c1, c2, th), y), sk.linear_combination(c1, th, y), rtol=1e-13, atol=1e-13)


def test_edge_cases_empty_single_and_degenerate(orc):
    import torch
    rng = np.random.default_rng(5)
    b6 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
    K = sk.dimension(b6)
    th = rng.standard_normal(K)
    # empty batches (host and device) for every entry point of the path
    assert sk.linear_combination(b6, th, np.empty((0, 6))).shape == (0,)
    assert sk.linear_combination(b6, th, sk.gradient(6)(np.empty((0, 6)))).shape == (0, 7)
    assert sk.linear_combination(b6, rng.standard_normal((K, 3)), np.empty((0, 6))).shape == (0, 3)
    assert sk.linear_combination(b6, torch.as_tensor(th).cuda(), torch.empty((0, 6), dtype=torch.float64, device="cuda")).shape == (0,)
    assert sk.basis_at(b6, np.empty((0, 6))).shape[:2] == (0, K)
    c20 = sk.Chebyshev(sk.InteriorGrid(), 20)
    assert sk.linear_combination(c20, np.ones(20), np.empty(0)).shape == (0,)
    # a single point drops the leading axis, like the reference's pointwise call
    x1 = rng.uniform(-1, 1, 6)
    v = sk.linear_combination(b6, th, x1)
    assert np.ndim(v) == 0 and abs(v - orc.smolyak_lincomb(2, 6, 4, 4, th, x1[None, :])[0, 0]) <= 1e-13 * orc.smolyak_lincomb(2, 6, 4, 4, th, x1[None, :], absmode=True)[0, 0]
    # degenerate bases: B = 0 (one basis function), through the leaf, block and exact kernels
    b0 = sk.smolyak_basis(sk.Chebyshev, sk.EndpointGrid(), sk.SmolyakParameters(0), 3)
    assert sk.dimension(b0) == 1
    x = rand_pm1(rng, (70, 3))
    assert same(sk.linear_combination(b0, np.array([2.5]), x), np.full(70, 2.5))
    assert same(sk.linear_combination(b0, np.array([[2.5, -1.0, 3.0]]), x), np.tile([2.5, -1.0, 3.0], (70, 1)))
    th40 = rng.standard_normal((1, 40))
    assert same(sk.linear_combination(b0, th40, x), np.tile(th40, (70, 1)))
    g = sk.linear_combination(b0, np.array([2.5]), sk.gradient(3)(x))
    assert same(g, np.tile([2.5, 0, 0, 0], (70, 1)))
    # N = 1 Chebyshev
    c1 = sk.Chebyshev(sk.EndpointGrid(), 1)
    assert same(sk.linear_combination(c1, np.array([3.0]), sk.d(np.linspace(-1, 1, 5))), np.tile([3.0, 0.0], (5, 1)))


def test_no_out_of_bounds_writes_around_caller_buffers(orc):
    """compute-sanitizer is not available on the GPU pool: the kernels write into the middle of a sentinel-filled
    device buffer through the raw C ABI, and everything around the result region must stay untouched.  Ragged sizes
    (partial tiles in points, vectors and terms) for every kernel family."""
    import ctypes as C
    import torch
    L = sk._lib
    rng = np.random.default_rng(31)
    SENT = -7.25e300
    PAD = 4099                                           # doubles of guard band on each side (odd: misaligns by 8 bytes)

    def run(basis, spec, theta, x, E, exact=False, basis_at=False):
        plan = sk.get_plan(basis, 0 if not isinstance(spec, int) else spec, None if isinstance(spec, int) else spec, 0)
        n = x.shape[0]
        K = plan.info.K
        words = n * K * E if basis_at else n * E
        buf = torch.full((words + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        xd = torch.as_tensor(np.ascontiguousarray(x)).cuda()
        out_ptr = C.c_void_p(buf.data_ptr() + PAD * 8)
        if basis_at:
            L.check(L.lib.sk_basis_at(plan.handle, C.c_void_p(xd.data_ptr()), n, out_ptr, n, L.SK_MEM_DEVICE, None))
        else:
            th = torch.as_tensor(np.ascontiguousarray(theta)).cuda()
            nvec = 1 if theta.ndim == 1 else theta.shape[1]
            flags = L.SK_MEM_DEVICE | (L.SK_MODE_EXACT if exact else L.SK_MODE_FAST)
            L.check(L.lib.sk_linear_combination(plan.handle, C.c_void_p(th.data_ptr()), K, nvec, C.c_void_p(xd.data_ptr()), n, out_ptr, flags, None))
        torch.cuda.synchronize()
        h = buf.cpu().numpy()
        assert (h[:PAD] == SENT).all() and (h[PAD + words:] == SENT).all(), "write outside the result buffer"
        assert not (h[PAD:PAD + words] == SENT).any(), "result element never written"

    b6 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
    K6 = sk.dimension(b6)
    b8 = sk.smolyak_basis(sk.Chebyshev, sk.EndpointGrid(), sk.SmolyakParameters(3), 8)
    K8 = sk.dimension(b8)
    for n in (1, 383, 385, 1000):
        x6 = rand_pm1(rng, (n, 6))
        run(b6, sk.gradient(6), rng.standard_normal(K6), x6, 7)                      # leaf kernel, gradient
        run(b6, None, rng.standard_normal(K6), x6, 1)                                # leaf kernel, values
        run(b6, None, rng.standard_normal((K6, 5)), x6, 5)                           # multi-vector leaf kernels (4 + 1 vectors, strided θ / out)
        run(b6, None, rng.standard_normal((K6, 37)), x6, 37)                         # block kernel, 64-wide tile, odd nvec
        run(b6, None, rng.standard_normal((K6, 130)), x6, 130)                       # block kernel, 256-wide tile, partial
        run(b6, sk.partial(1, 1) | sk.partial(0, 0, 2), rng.standard_normal(K6), x6, 6)   # second-order-jet leaves (column map)
        run(b6, sk.gradient(6), rng.standard_normal(K6), x6, 7, exact=True)          # one-thread-per-point exact kernel
        run(b8, sk.gradient(8), rng.standard_normal(K8), rand_pm1(rng, (n, 8)), 9)   # leaf kernel with the unit interpreter
    for n in (65537, 66049):                                                         # two-points-per-thread values kernel (tiles of 512 points)
        x6 = rand_pm1(rng, (n, 6))
        run(b6, None, rng.standard_normal(K6), x6, 1)
        run(b6, None, rng.standard_normal((K6, 3)), x6, 3)                           # one launch per vector, strided θ / out
        run(b8, None, rng.standard_normal(K8), rand_pm1(rng, (n, 8)), 1)
    for n in (1, 63, 65, 200):
        run(b6, None, None, rand_pm1(rng, (n, 6)), 1, basis_at=True)                 # basis_at, values (streaming stores)
        run(b6, sk.gradient(6), None, rand_pm1(rng, (n, 6)), 7, basis_at=True)       # basis_at, staged bulk copies
    c = sk.Chebyshev(sk.InteriorGrid(), 33) @ sk.BoundedLinear(-2, 3)
    for n in (1, 255, 257, 1 << 17):
        run(c, 1, rng.standard_normal(33), rng.uniform(-2, 3, n), 2)                 # univariate (constant-bank θ above 2^17)
        run(c, 0, rng.standard_normal((33, 3)), rng.uniform(-2, 3, n), 3)


def test_golden_fixtures(orc):
    g = np.load(os.path.join(ge.ROOT, "tests", "golden", "lincomb.npz"))
    b20 = sk.Chebyshev(sk.InteriorGrid(), 20)
    assert same(sk.linear_combination(b20, g["c1_theta"], g["c1_x"], exact=True), g["c1_val"])
    assert same(sk.collocation_matrix(b20), g["c1_colloc"])
    b64 = sk.Chebyshev(sk.InteriorGrid(), 64) @ sk.BoundedLinear(-2, 3)
    assert same(sk.linear_combination(b64, g["c2_theta"], sk.d(g["c2_y"]), exact=True), g["c2_out"])
    for name, t in (("semi", sk.SemiInfRational(0.7, 0.3)), ("inf", sk.InfRational(0.4, 0.9))):
        b128 = sk.Chebyshev(sk.InteriorGrid(), 128) @ t
        assert same(sk.linear_combination(b128, g["c3_theta"], sk.d(g[f"c3_{name}_y"]), exact=True), g[f"c3_{name}_d1"])
        assert same(sk.linear_combination(b128, g["c3_theta"], (sk.d ** 2)(g[f"c3_{name}_y"]), exact=True,
                                          allow_rational_d2=True), g[f"c3_{name}_d2"])
    ct = sk.coordinate_transformations(sk.BoundedLinear(0, 4), sk.SemiInfRational(1, 2), sk.InfRational(0, 4), sk.SemiInfRational(-3, 3))
    b4 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3), 4)
    assert same(sk.linear_combination(b4 @ ct, g["c4_theta"], sk.gradient(4)(g["c4_y"]), exact=True), g["c4_grad"])
    assert same(sk.linear_combination(b4 @ ct, g["c4_theta"], g["c4_y"], exact=True), g["c4_val"])
    np.testing.assert_allclose(sk.linear_combination(b4 @ ct, g["c4_theta"], sk.gradient(4)(g["c4_y"])), g["c4_grad"], rtol=1e-11, atol=1e-11)
    assert same(sk.linear_combination(b4, g["c5_theta"], g["c5_y"], exact=True), g["c5_out"])


def test_golden_fixtures_extras(orc):
    # committed vectors for the later paths (tests/golden/make_golden_extras.py): coefficient block, general ∂
    # specification, refinement, fit
    g = np.load(os.path.join(ge.ROOT, "tests", "golden", "extras.npz"))
    ct = sk.coordinate_transformations(sk.BoundedLinear(0, 4), sk.SemiInfRational(1, 2), sk.InfRational(0, 4), sk.BoundedLinear(-1, 2), sk.SemiInfRational(-3, 3))
    b5 = sk.smolyak_basis(sk.Chebyshev, sk.EndpointGrid(), sk.SmolyakParameters(3), 5) @ ct
    assert same(sk.linear_combination(b5, g["k4_theta"], g["k4_y"], exact=True), g["k4_out"])
    assert within(sk.linear_combination(b5, g["k4_theta"], g["k4_y"]), g["k4_out"], g["k4_scale"])                 # block kernel
    assert within(sk.linear_combination(b5, np.ascontiguousarray(g["k4_theta"][:, :6]), g["k4_y"]), g["k4_out"][:, :6], g["k4_scale"][:, :6])   # leaf per vector
    ct3 = sk.coordinate_transformations(sk.BoundedLinear(0, 4), sk.SemiInfRational(1, 2), sk.BoundedLinear(-2, 2))
    b3 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(3), 3) @ ct3
    spec = sk.partial(1, 1) | sk.partial(0, 0, 2) | sk.partial(2)
    assert same(sk.linear_combination(b3, g["gs_theta"], spec(g["gs_y"]), exact=True), g["gs_out"])
    assert within(sk.linear_combination(b3, g["gs_theta"], spec(g["gs_y"])), g["gs_out"], g["gs_scale"])            # general nested kernel
    bs = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3), 3)
    th = sk.collocation_solve(bs, g["cs_y"])
    assert np.max(np.abs(th - g["cs_theta"])) <= 50 * float(g["cs_cond"]) * np.finfo(float).eps * np.max(np.abs(g["cs_theta"]))


def test_device_pointers_and_sharded_driver(orc):
    import torch
    rng = np.random.default_rng(55)
    d, B = 6, 4
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    theta = rng.standard_normal(sk.dimension(basis))
    x = rng.uniform(-1, 1, (5000, d))
    host = sk.linear_combination(basis, theta, sk.gradient(d)(x))
    dev = sk.linear_combination(basis, torch.as_tensor(theta).cuda(), sk.gradient(d)(torch.as_tensor(x).cuda()))
    assert dev.is_cuda and same(dev.cpu().numpy(), host)
    sh = sk.linear_combination_sharded(basis, theta, sk.gradient(d)(x), devices=list(range(torch.cuda.device_count())))
    assert same(sh, host)
    Bm = sk.basis_at(basis, torch.as_tensor(x[:64]).cuda())
    assert same(Bm.cpu().numpy(), np.asarray(sk.basis_at(basis, x[:64])))


def test_sharded_gather_fused_into_the_kernel_two_gpus(orc):
    # gather mode 2 (needs two GPUs with peer access; skipped on one): every device's full buffer must equal the sharded results,
    # for a plan the peer-storing leaf kernel serves (d = 6 gradient) and for one that goes through the copy engines (d = 4)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from spectralkit_b200.sharding import Comm
    comm = Comm([0, 1])
    try:
        for d, B, n in ((6, 4, 100_004), (6, 4, 100_001), (4, 3, 2_500_001)):      # aligned slices, odd slice length (scalar peer stores), copy engines
            basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
            K = sk.dimension(basis)
            plans = [sk.get_plan(basis, 0, sk.gradient(d), g) for g in (0, 1)]
            per = (n + 1) // 2
            thetas, xs, sh, full = [], [], [], []
            th0 = torch.as_tensor(np.random.default_rng(d).standard_normal(K))
            for g in (0, 1):
                dev = torch.device("cuda", g)
                thetas.append(th0.to(dev))
                lo, hi = sk.shard_bounds(n, 2, g)
                gen = torch.Generator(device=dev).manual_seed(50 + g)
                xs.append(torch.rand((hi - lo, d), generator=gen, device=dev, dtype=torch.float64) * 1.98 - 0.99)
                sh.append(torch.empty((hi - lo, d + 1), dtype=torch.float64, device=dev))
                full.append(torch.zeros((per * 2, d + 1), dtype=torch.float64, device=dev))
            comm.linear_combination(plans, thetas, xs, n, sh, gather=False)
            comm.linear_combination(plans, thetas, xs, n, full, gather=2)
            for g in (0, 1):
                lo, hi = sk.shard_bounds(n, 2, g)
                for h in (0, 1):
                    assert torch.equal(full[h][lo:hi].cpu(), sh[g].cpu())
    finally:
        comm.close()


def test_linear_combination_peers_entry_on_one_gpu(orc):
    # sk_linear_combination_peers without peers is sk_linear_combination; argument checks of the entry
    import ctypes as C
    import torch
    from spectralkit_b200.sharding import linear_combination_peers
    L = sk._lib
    d = 6
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), d)
    plan = sk.get_plan(basis, 0, sk.gradient(d), 0)
    theta = torch.as_tensor(np.random.default_rng(3).standard_normal(sk.dimension(basis))).cuda()
    x = torch.rand((1000, d), device="cuda", dtype=torch.float64) * 1.98 - 0.99
    out = torch.empty((1000, d + 1), device="cuda", dtype=torch.float64)
    assert linear_combination_peers(plan, theta, x, out, []) is False
    torch.cuda.synchronize()
    assert torch.equal(out, sk.linear_combination(basis, theta, sk.gradient(d)(x)))
    # a buffer of this device passed as its own "peer": the kernel stores there as well
    twin = torch.zeros_like(out)
    assert linear_combination_peers(plan, theta, x, out, [twin.data_ptr()]) is True
    torch.cuda.synchronize()
    assert torch.equal(twin, out)
    fused = C.c_int(0)
    rc = L.lib.sk_linear_combination_peers(plan.handle, C.c_void_p(theta.data_ptr()), int(theta.shape[0]), C.c_void_p(x.data_ptr()), 1000,
                                           C.c_void_p(out.data_ptr()), None, 8, L.SK_MEM_DEVICE, None, C.byref(fused))
    assert rc != 0      # more than seven peers / NULL array


def test_large_host_batch_goes_through_the_chunk_pipeline(orc):
    # > 2^22 points: the host path double-buffers chunks; results must not depend on chunking
    rng = np.random.default_rng(77)
    N = 16
    t, ot = mk_tr(orc, 1, -2, 3)
    theta = rng.standard_normal(N)
    n = (1 << 22) * 2 + 12345
    y = rng.uniform(-2, 3, n)
    got = sk.linear_combination(sk.Chebyshev(sk.InteriorGrid(), N) @ t, theta, sk.d(y), exact=True)
    sel = np.concatenate([np.arange(0, 1000), np.arange((1 << 22) - 500, (1 << 22) + 500), np.arange(n - 1000, n)])
    assert same(got[sel], orc.uni_lincomb(N, theta, y[sel], tr=ot, D=1))


def test_full_size_headline_properties(orc):
    # BASELINE cfg 4 at its full size (10^8 points, device resident): properties that do not need the
    # oracle on every point — linearity in θ, and a strided subsample against the oracle.
    import torch
    n = int(os.environ.get("SK_FULL_N", 100_000_000))
    d, B = 6, 4
    pairs = [mk_tr(orc, *k) for k in MIXED6]
    ct = sk.coordinate_transformations(*[p[0] for p in pairs])
    otr = [p[1] for p in pairs]
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d) @ ct
    gen = torch.Generator(device="cuda").manual_seed(104)
    u = torch.rand((n, d), generator=gen, device="cuda", dtype=torch.float64) * 1.98 - 0.99
    y = sk.transform_from(None, ct, u)
    del u
    rng = np.random.default_rng(4)
    th1, th2 = rng.standard_normal(2561), rng.standard_normal(2561)
    spec = sk.gradient(d)
    o1 = sk.linear_combination(basis, torch.as_tensor(th1).cuda(), spec(y))
    o2 = sk.linear_combination(basis, torch.as_tensor(th2).cuda(), spec(y))
    o12 = sk.linear_combination(basis, torch.as_tensor(th1 + th2).cuda(), spec(y))
    assert torch.isfinite(o12).all()
    lin = (o12 - (o1 + o2)).abs().amax(dim=0).cpu().numpy()
    ref = (o1.abs() + o2.abs()).amax(dim=0).cpu().numpy()
    assert (lin <= 1e-9 * ref).all(), (lin, ref)
    sel = torch.arange(0, n, max(1, n // 1500), device="cuda")
    ys = y[sel].cpu().numpy()
    ospec = orc.gradient_spec(d)
    want = orc.smolyak_lincomb(2, d, B, B, th1, ys, trs=otr, spec=ospec)
    scale = orc.smolyak_lincomb(2, d, B, B, th1, ys, trs=otr, spec=ospec, absmode=True)
    assert within(o1[sel].cpu().numpy(), want, scale)


def test_stacked_coefficients_like_make_policy_functions(orc):
    # src/experimental.jl:156-166: one coefficient vector, a range of dimension(basis) entries per policy function
    rng = np.random.default_rng(12)
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3), 4) @ sk.coordinate_transformations(
        sk.BoundedLinear(0, 4), sk.BoundedLinear(-2, 2), sk.SemiInfRational(1, 2), sk.InfRational(0, 4))
    K, P = sk.dimension(basis), 3
    coef = rng.standard_normal(K * P)
    y = sk.transform_from(basis, basis.transformation, rng.uniform(-0.9, 0.9, (200, 4)))
    got = sk.stacked_linear_combination(basis, coef, y)
    assert got.shape == (200, P)
    for p in range(P):
        assert same(got[:, p], sk.linear_combination(basis, coef[p * K:(p + 1) * K].copy(), y))
    with pytest.raises(sk.ArgumentError):
        sk.stacked_linear_combination(basis, coef[:-1], y)


def test_randomised_parity_sweep():
    # profiles/fuzz_parity.py walks the whole dispatch table with random configurations (20 000 of them are logged in
    # profiles/r1_fuzz_parity.jsonl); a short sweep runs with the suite
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ge.ROOT, "profiles", "fuzz_parity.py"), "200", "77"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    summary = json.loads(out.stdout.strip().splitlines()[-1])
    assert summary["failures"] == 0 and summary["worst_fast_scaled_error"] <= TAU, out.stdout[-2000:]


def test_experimental_policy_functions_and_residual_sums(orc):
    """The batched part of src/experimental.jl: make_policy_functions (:156-166: several functions over one basis from
    one stacked coefficient vector, each followed by a bridge :242-266) and sum_of_squared_residuals (:228-233)."""
    import torch
    rng = np.random.default_rng(2026)
    d, B = 3, 3
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(B), d)
    K = sk.dimension(basis)
    # three policies: identity, a bridge between two bounded domains, a bridge from a bounded to a semi-infinite domain
    pt = {"c": None, "k": sk.bridge(sk.BoundedLinear(0.5, 3.0), sk.BoundedLinear(-4.0, 4.0)),
          "h": sk.bridge(sk.SemiInfRational(0.0, 2.0), sk.BoundedLinear(-6.0, 6.0))}
    coef = rng.standard_normal(3 * K) * 0.3
    x = rand_pm1(rng, (257, d))
    pf = sk.make_policy_functions(pt, basis, coef)
    got = np.asarray(pf(x))
    raw = np.stack([orc.smolyak_lincomb(2, d, B, B, coef[m * K:(m + 1) * K], x)[:, 0] for m in range(3)], 1)
    scale = np.stack([orc.smolyak_lincomb(2, d, B, B, coef[m * K:(m + 1) * K], x, absmode=True)[:, 0] for m in range(3)], 1)
    assert np.max(np.abs(got[:, 0] - raw[:, 0]) / scale[:, 0]) <= 1e-13
    # bridges applied to the library's own raw values are the reference's unfused operations: bit-identical to the oracle
    raw_gpu = np.asarray(sk.stacked_linear_combination(basis, coef, x))
    def obridge(outer, inner, v):
        return np.array([orc.transform_from(outer, orc.transform_to(inner, float(t))[0]) for t in v])
    assert same(got[:, 0], raw_gpu[:, 0])
    assert same(got[:, 1], obridge(orc.BoundedLinear(0.5, 3.0), orc.BoundedLinear(-4.0, 4.0), raw_gpu[:, 1]))
    assert same(got[:, 2], obridge(orc.SemiInfRational(0.0, 2.0), orc.BoundedLinear(-6.0, 6.0), raw_gpu[:, 2]))
    assert set(pf.named(x)) == {"c", "k", "h"}
    # device-resident inputs give the same numbers
    got_dev = pf(torch.as_tensor(x).cuda()).cpu().numpy()
    assert same(got_dev, got)
    # length check of the reference (@argcheck lastindex(coefficients) == last(last(ranges)))
    with pytest.raises(sk.ArgumentError):
        sk.make_policy_functions(pt, basis, coef[:-1])
    # residual sums: deterministic, equal to the float64 sum within rounding; the fused least-squares form agrees
    target = got + rng.standard_normal(got.shape) * 1e-3
    want = float(np.sum((got - target) ** 2))
    s1 = sk.sum_abs2(got - target)
    s2 = sk.sum_abs2(torch.as_tensor(got - target).cuda())
    s3 = pf.residual_ssq(x, target)
    assert s1 == s2 and abs(s1 - want) <= 1e-12 * want and abs(s3 - want) <= 1e-12 * want
    # sum_of_squared_residuals with a user residual function vectorised over the grid
    g = sk.grid(basis)
    r = sk.sum_of_squared_residuals(lambda p, grid: {"euler": p["c"] - 0.1 * grid[:, 0], "budget": p["k"] - 1.0}, pf, g)
    vals = np.asarray(pf(g))
    assert abs(r - float(np.sum((vals[:, 0] - 0.1 * g[:, 0]) ** 2) + np.sum((vals[:, 1] - 1.0) ** 2))) <= 1e-12 * r


def test_uni_constant_bank_slots_shared_between_derivative_orders(orc):
    """the constant-bank θ slots of the univariate fast kernels are one set per device for every derivative
    order.  Two streams evaluate order 0 and order 1 with different θ concurrently, many times: each result must be the
    one its own θ gives (a private slot counter per template instantiation let one stream overwrite the other's θ)."""
    import torch
    n, N = 1 << 18, 64                                                   # n >= 2^17: the constant-bank path
    b = sk.Chebyshev(sk.InteriorGrid(), N) @ sk.BoundedLinear(-2, 3)
    gen = torch.Generator(device="cuda").manual_seed(5)
    y = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * 5 - 2
    th0 = torch.randn(N, generator=gen, device="cuda", dtype=torch.float64)
    th1 = torch.randn(N, generator=gen, device="cuda", dtype=torch.float64) * 3 + 1
    want0 = sk.linear_combination(b, th0, y).clone()
    want1 = sk.linear_combination(b, th1, sk.d(y)).clone()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    outs0, outs1 = [], []
    for _ in range(12):
        with torch.cuda.stream(s0):
            outs0.append(sk.linear_combination(b, th0, y))
        with torch.cuda.stream(s1):
            outs1.append(sk.linear_combination(b, th1, sk.d(y)))
    torch.cuda.synchronize()
    assert all(torch.equal(o, want0) for o in outs0)
    assert all(torch.equal(o, want1) for o in outs1)


def test_exact_mode_for_bases_beyond_shared_memory_tables(orc):
    """SK_MODE_EXACT for bases whose full tables exceed shared memory (round 1 returned SK_ERR_UNSUPPORTED): the
    chunked-table kernel keeps the reference's sequential s = s + θ_k·b_k (generic_api.jl:114-128) and is bit-identical."""
    rng = np.random.default_rng(909)
    for code, d, B, n, tr in [(0, 6, 4, 41, False), (2, 12, 3, 35, False), (0, 3, 5, 50, True)]:
        basis0 = sk.smolyak_basis(sk.Chebyshev, GRIDS[code], sk.SmolyakParameters(B), d)
        K = sk.dimension(basis0)
        trs_sk = trs_or = None
        basis = basis0
        if tr:
            trs_sk = [sk.BoundedLinear(-1, 2), sk.SemiInfRational(0.5, 2.0), sk.InfRational(0.0, 1.5)]
            trs_or = [orc.BoundedLinear(-1, 2), orc.SemiInfRational(0.5, 2.0), orc.InfRational(0.0, 1.5)]
            basis = basis0 @ sk.coordinate_transformations(*trs_sk)
        u = np.clip((rng.random((n, d)) - 0.5) * 2.2, -0.97, 0.97)
        y = orc.transform_from_batch(trs_or, u) if tr else u
        theta = rng.standard_normal(K)
        assert same(np.asarray(sk.linear_combination(basis, theta, y, exact=True)), orc.smolyak_lincomb(code, d, B, B, theta, y, trs=trs_or)[:, 0])
        want = orc.smolyak_lincomb(code, d, B, B, theta, y, trs=trs_or, spec=orc.gradient_spec(d))
        assert same(np.asarray(sk.linear_combination(basis, theta, sk.gradient(d)(y), exact=True)), want)
        # SVector coefficients (values only), 11 vectors: two groups of 8
        th = rng.standard_normal((K, 11))
        assert same(np.asarray(sk.linear_combination(basis, th, y, exact=True)), orc.smolyak_lincomb(code, d, B, B, th, y, trs=trs_or))
