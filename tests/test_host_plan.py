"""CPU-side checks of the product library: it loads, exports the whole C ABI, fails loudly without a
GPU, and its host plan builder (integers + grids) is BIT-EXACT against the oracle."""
import ctypes
import itertools
import math
import re
import os

import numpy as np
import pytest

import __graft_entry__ as ge

sk = ge.load_package()
GRIDS = [(sk.InteriorGrid(), 0), (sk.EndpointGrid(), 1), (sk.InteriorGrid2(), 2)]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ge.ROOT, "include", "spectralkit_b200.h")).read()
    declared = set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(sk._lib.EXPORTS)
    lib = ctypes.CDLL(sk._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert "sm_100a" in sk._lib.lib.sk_version().decode()


def test_library_contains_only_sm100a_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", sk._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    basis = sk.Chebyshev(sk.InteriorGrid(), 5)
    with pytest.raises(sk.CudaError):
        sk.linear_combination(basis, np.zeros(5), np.zeros(3))
    with pytest.raises(sk.CudaError):
        sk.basis_at(basis, np.zeros(3))
    with pytest.raises(sk.CudaError):
        sk.transform_to(None, sk.BoundedLinear(0, 1), np.zeros(3))


def test_nesting_lengths(orc):
    for g, code in GRIDS:
        for b in range(0, 8):
            assert sk.nesting_total_length(g, b) == orc.nesting_total_length(code, b)
            assert sk.nesting_block_length(g, b) == orc.nesting_block_length(code, b)


def test_grid_shuffle_bit_exact(orc):
    # closed-form shuffle (product) vs the reference's iterator state machines (oracle)
    for g, code in GRIDS:
        for b in range(0, 8):
            n = orc.nesting_total_length(code, b)
            assert sk.grid_shuffle(g, n).tolist() == orc.grid_shuffle(code, n).tolist(), (code, b)
        with pytest.raises(sk.ArgumentError):
            sk.grid_shuffle(g, 4)


@pytest.mark.parametrize("g,code", GRIDS)
def test_smolyak_indices_bit_exact(orc, g, code):
    # constrained odometer (product) vs `__inc` state machine (oracle): identical tuples, identical order
    cases = [(d, B, M) for B in range(0, 5) for M in range(0, B + 1) for d in range(1, 5)]
    cases += [(6, 4, 4), (6, 4, 2), (5, 5, 3), (10, 3, 3), (7, 2, 5)]
    for d, B, M in cases:
        K = orc.smolyak_length(code, d, B, M)
        if K > 60000:
            continue
        basis = sk.smolyak_basis(sk.Chebyshev, g, _params(B, M), d)
        assert sk.dimension(basis) == K
        assert basis.univariate_parent.N == orc.smolyak_parent_dimension(code, B, M)
        assert (sk.smolyak_indices(basis) == orc.smolyak_indices(code, d, B, M)).all(), (code, d, B, M)


def _params(B, M):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return sk.SmolyakParameters(B, M)


def test_smolyak_indices_golden_naive_filter():
    # fixtures from the naive filtered Cartesian enumeration (test/test_smolyak_traversal.jl:74-90)
    gold = np.load(os.path.join(ge.ROOT, "tests", "golden", "indices.npz"))
    for key in gold.files:
        code, d, B, M = [int(v[1:]) for v in key.split("_")]
        g = [g for g, c in GRIDS if c == code][0]
        basis = sk.smolyak_basis(sk.Chebyshev, g, _params(B, M), d)
        assert (sk.smolyak_indices(basis) == gold[key]).all(), key


def test_headline_sizes():
    b4 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
    assert sk.dimension(b4) == 2561 and b4.univariate_parent.N == 31
    b5 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(5), 10)
    assert sk.dimension(b5) == 77505 and b5.univariate_parent.N == 63
    # docstring example smolyak_api.jl:46-49
    b = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(3), 2)
    assert sk.dimension(b) == 81 and b.univariate_parent.N == 27


def test_smolyak_parameters_warning():
    with pytest.warns(UserWarning, match="M > B replaced with M = B"):
        p = sk.SmolyakParameters(2, 4)
    assert (p.B, p.M) == (2, 2)


@pytest.mark.parametrize("g,code", GRIDS)
def test_grids_bit_exact(orc, g, code):
    for N in [1, 2, 3, 5, 7, 9, 15, 17, 20, 27, 31, 33, 63, 64, 65, 81, 127, 128, 243]:
        assert sk.grid(sk.Chebyshev(g, N)).tolist() == orc.cheb_grid(code, N).tolist(), (code, N)
    t, ot = sk.SemiInfRational(0.7, 0.3), orc.SemiInfRational(0.7, 0.3)
    with np.errstate(all="ignore"):
        assert sk.grid(sk.Chebyshev(g, 20) @ t).tolist() == orc.cheb_grid(code, 20, tr=ot).tolist()
    for d, B, M in [(2, 3, 3), (3, 3, 2), (6, 4, 4), (4, 5, 3)]:
        if orc.smolyak_length(code, d, B, M) > 60000:
            continue
        basis = sk.smolyak_basis(sk.Chebyshev, g, _params(B, M), d)
        assert (sk.grid(basis) == orc.smolyak_grid(code, d, B, M)).all()
    basis = sk.smolyak_basis(sk.Chebyshev, g, sk.SmolyakParameters(3), 3)
    ct = sk.coordinate_transformations(sk.BoundedLinear(0, 4), sk.SemiInfRational(1, 2), sk.InfRational(0, 4))
    otr = [orc.BoundedLinear(0, 4), orc.SemiInfRational(1, 2), orc.InfRational(0, 4)]
    with np.errstate(all="ignore"):
        got, want = sk.grid(basis @ ct), orc.smolyak_grid(code, 3, 3, 3, trs=otr)
    assert ((got == want) | (np.isnan(got) & np.isnan(want))).all()


def test_grid_golden_mpmath():
    # correctly rounded sinpi/cospi pinned by 240-bit mpmath (tests/golden/make_golden.py)
    gold = np.load(os.path.join(ge.ROOT, "tests", "golden", "grids.npz"))
    for g, code in GRIDS:
        for N in (20, 31, 63, 128, 243):
            assert sk.grid(sk.Chebyshev(g, N)).tolist() == gold[f"g{code}_N{N}"].tolist()


def test_partials_algebra_matches_oracle(orc):
    rng = np.random.default_rng(0)
    for _ in range(300):
        dim = int(rng.integers(1, 5))
        specs = []
        for _ in range(int(rng.integers(1, 4))):
            ln = int(rng.integers(1, dim + 1))
            p = [int(v) for v in rng.integers(0, 3, ln)]
            specs.append(tuple(p))
        spec = sk.partial(*specs[0])
        for p in specs[1:]:
            spec = spec | sk.partial(*p)
        try:
            want = orc.deriv_spec(specs, dim)
        except orc.OracleError:
            continue
        if len(want) > 64:
            continue
        table, deg = spec.expansion(dim)
        assert table.tolist() == want.tolist(), specs
        assert deg.tolist() == orc.partials_degrees(orc.partials_minimal(specs, dim), dim).tolist()
    assert sk.gradient(6).expansion(6)[0].tolist() == orc.gradient_spec(6).tolist()


def test_transformation_constructors():
    # test/test_transformations.jl:4-5,27-29,61-64
    for ctor, args in [(sk.BoundedLinear, (-1.0, math.inf)), (sk.BoundedLinear, (-1.0, -2.0)),
                       (sk.SemiInfRational, (-1.0, math.inf)), (sk.SemiInfRational, (-1.0, 0.0)),
                       (sk.SemiInfRational, (math.nan, 2.0)), (sk.InfRational, (1.0, math.inf)),
                       (sk.InfRational, (1.0, 0.0)), (sk.InfRational, (1.0, -2.0)), (sk.InfRational, (math.nan, 1))]:
        with pytest.raises(sk.DomainError):
            ctor(*args)
    t = sk.BoundedLinear(1, 5)
    assert (t.m, t.s) == (3.0, 2.0) and sk.extrema(sk.domain(t)) == (1.0, 5.0)
    assert sk.extrema(sk.domain(sk.SemiInfRational(3.0, 4.0))) == (3.0, math.inf)
    assert sk.extrema(sk.domain(sk.InfRational(0.0, 1.0))) == (-math.inf, math.inf)
    # test/test_transformations.jl:113-127 reprs
    assert repr(sk.BoundedLinear(2.0, 3)) == "(2.0,3.0) ↔ domain [linear transformation]"
    assert repr(sk.SemiInfRational(7.0, 1)) == "(7.0,∞) ↔ domain [rational transformation with scale 1.0]"
    assert repr(sk.InfRational(0.5, 1)) == "(-∞,∞) ↔ domain [rational transformation with center 0.5, scale 1.0]"


def test_basis_constructors_and_api_surface():
    with pytest.raises(sk.ArgumentError):
        sk.Chebyshev(sk.InteriorGrid(), 0)                                    # test/test_chebyshev.jl:6-7
    with pytest.raises(TypeError):
        sk.Chebyshev("invalid_grid", 10)                                      # :8
    with pytest.raises(TypeError):
        sk.smolyak_basis(sk.Chebyshev, "invalid_grid", sk.SmolyakParameters(3), 2)   # test/test_smolyak.jl:9
    basis = sk.Chebyshev(sk.EndpointGrid(), 10)
    t = sk.BoundedLinear(1.0, 2.0)
    tb = basis @ t
    assert sk.is_function_basis(tb) and sk.is_function_basis(sk.Chebyshev) and not sk.is_function_basis("a fish")
    assert sk.domain(tb) == sk.domain(t) and sk.dimension(tb) == 10 and sk.parent(tb) is basis and sk.transformation(tb) is t
    assert sk.transformation(basis) is None
    b0 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid(), sk.SmolyakParameters(2, 2), 2)
    ct = sk.coordinate_transformations(sk.BoundedLinear(1.0, 2.0), sk.SemiInfRational(0, 1))
    b = b0 @ ct
    assert len(b) == 2 and b[2] == b0.univariate_parent @ ct.transformations[1]   # test/test_generic_api.jl:70
    with pytest.raises(IndexError):
        b0[3]
    with pytest.raises(sk.ArgumentError):
        b0 @ sk.coordinate_transformations(sk.BoundedLinear(1.0, 2.0))
    with pytest.raises(sk.ArgumentError):
        sk.linear_combination(basis, np.zeros(11))                            # test/test_generic_api.jl:29
    # 𝑑 algebra: test/test_derivatives.jl:14-18
    assert (sk.d * sk.d).D == 2 and (sk.d ** 3).D == 3 and ((sk.d ** 2) << 3).partials == ((0, 0, 2),)
    assert repr(sk.partial(1, 2) | sk.partial(2, 1)) == "union(∂(1, 2), ∂(2, 1))"
    assert sk.partial(1, 0).partials == ((1,),)
    with pytest.raises(sk.ArgumentError):
        sk.partial((1, 0))


# ---- refinement between bases: is_subset_basis / augment_coefficients (host index merge through the C ABI)
def test_is_subset_basis_rules():
    cheb = lambda g, n: sk.Chebyshev(g, n)
    assert sk.is_subset_basis(cheb(sk.InteriorGrid(), 5), cheb(sk.EndpointGrid(), 6))          # test/test_chebyshev.jl:63-65
    assert not sk.is_subset_basis(cheb(sk.InteriorGrid(), 5), cheb(sk.InteriorGrid(), 4))      # :67-68
    sm = lambda g, B, M, d=2: sk.smolyak_basis(sk.Chebyshev, g, sk.SmolyakParameters(B, M), d)
    b1 = sm(sk.InteriorGrid(), 2, 2)
    assert not sk.is_subset_basis(b1, sm(sk.EndpointGrid(), 2, 3))                              # test/test_smolyak.jl:60-63 (after the M clamp: B2 = 2, M2 = 2, other grid)
    assert not sk.is_subset_basis(b1, sm(sk.InteriorGrid(), 2, 1))                              # :65-68
    assert sk.is_subset_basis(b1, sm(sk.InteriorGrid(), 3, 2))                                  # :70-71
    assert not sk.is_subset_basis(b1, sm(sk.InteriorGrid(), 3, 2, d=3))
    assert not sk.is_subset_basis(b1, cheb(sk.InteriorGrid(), 50))
    t, t2 = sk.SemiInfRational(0.3, 0.9), sk.SemiInfRational(0.3, 0.8)
    assert sk.is_subset_basis(cheb(sk.InteriorGrid(), 5) @ t, cheb(sk.InteriorGrid(), 8) @ t)   # test/test_chebyshev.jl:75-82
    assert not sk.is_subset_basis(cheb(sk.InteriorGrid(), 5) @ t, cheb(sk.InteriorGrid(), 8) @ t2)
    assert not sk.is_subset_basis(cheb(sk.InteriorGrid(), 5) @ t, cheb(sk.InteriorGrid(), 8))
    with pytest.raises(sk.ArgumentError):
        sk.augment_coefficients(cheb(sk.InteriorGrid(), 5), cheb(sk.InteriorGrid(), 4), np.ones(5))
    with pytest.raises(sk.ArgumentError):
        sk.augment_coefficients(cheb(sk.InteriorGrid(), 5), cheb(sk.InteriorGrid(), 5), np.ones(4))   # too few coefficients (:70-71)


@pytest.mark.parametrize("code,d,p1,p2", [(0, 2, (2, 2), (3, 2)), (2, 3, (2, 1), (4, 3)), (1, 2, (1, 1), (3, 3)), (2, 3, (0, 0), (2, 2)), (0, 2, (2, 2), (2, 2))])
def test_augment_coefficients_matches_padding_iterator(code, d, p1, p2):
    from oracle import py_oracle as po
    grid = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}[code]
    b1 = sk.smolyak_basis(sk.Chebyshev, grid, sk.SmolyakParameters(*p1), d)
    b2 = sk.smolyak_basis(sk.Chebyshev, grid, sk.SmolyakParameters(*p2), d)
    rng = np.random.default_rng(17)
    th = rng.standard_normal(sk.dimension(b1))
    want = np.array(po.augment_coefficients(code, d, p1[0], p1[1], p2[0], p2[1], list(th)))
    got = sk.augment_coefficients(b1, b2, th)
    assert got.shape == (sk.dimension(b2),) and (got == want).all()
    assert np.count_nonzero(got) == np.count_nonzero(th)          # nothing lost: the smaller traversal is a subsequence
    thv = rng.standard_normal((sk.dimension(b1), 3))                # SVector coefficients
    gv = sk.augment_coefficients(b1, b2, thv)
    for v in range(3):
        assert (gv[:, v] == np.array(po.augment_coefficients(code, d, p1[0], p1[1], p2[0], p2[1], list(thv[:, v])))).all()
    c = sk.augment_coefficients(sk.Chebyshev(grid, 4), sk.Chebyshev(grid, 9), np.arange(4.0))
    assert (c == np.array([0, 1, 2, 3, 0, 0, 0, 0, 0.0])).all()


def test_augment_coefficients_golden():
    g = np.load(os.path.join(ge.ROOT, "tests", "golden", "extras.npz"))
    b1 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(2, 1), 3)
    b2 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4, 3), 3)
    assert (sk.augment_coefficients(b1, b2, g["ag_theta1"]) == g["ag_theta2"]).all()


def test_oversized_bases_are_refused_with_a_status_code():
    """nothing may unwind through the C ABI; index sets beyond SK_MAX_TERMS are refused up front
    (d = 16, B = M = 8, InteriorGrid would be 4.3e8 terms: 13.7 GB of index table)."""
    import ctypes as C
    L = sk._lib
    desc = L.BasisDesc(L.SK_BASIS_SMOLYAK, L.SK_GRID_INTERIOR, 0, 16, 8, 8)
    K = C.c_int64(0)
    assert L.lib.sk_dimension(C.byref(desc), C.byref(K)) in (L.SK_OK, L.SK_ERR_UNSUPPORTED)
    h = C.c_void_p()
    ds = L.DerivSpec(0, 0, None, 0)
    rc = L.lib.sk_plan_create(C.byref(desc), None, C.byref(ds), 0, C.byref(h))
    assert rc == L.SK_ERR_UNSUPPORTED and b"SK_MAX_TERMS" in L.lib.sk_last_error()
    out = (C.c_int32 * 16)()
    assert L.lib.sk_smolyak_indices(C.byref(desc), out) == L.SK_ERR_UNSUPPORTED


def test_multi_gpu_and_experimental_entries_fail_cleanly_without_a_device():
    """The new entry points validate their arguments and report SK_ERR_NO_DEVICE / SK_ERR_ARGUMENT on a box without
    GPUs (no CPU fallback, no crash)."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    L = sk._lib
    devs = (C.c_int * 2)(0, 1)
    h = C.c_void_p()
    assert L.lib.sk_comm_create(2, devs, C.byref(h)) == L.SK_ERR_NO_DEVICE
    assert L.lib.sk_comm_create(0, devs, C.byref(h)) == L.SK_ERR_ARGUMENT
    assert L.lib.sk_comm_destroy(None) == L.SK_OK
    r = (C.c_double * 4)(1, 2, 3, 4)
    res = C.c_double(-1)
    assert L.lib.sk_sum_abs2(r, 4, C.byref(res), L.SK_MEM_HOST, 0, None) == L.SK_ERR_NO_DEVICE
    assert L.lib.sk_sum_abs2(r, 0, C.byref(res), L.SK_MEM_HOST, 0, None) == L.SK_OK and res.value == 0.0
    assert L.lib.sk_policy_functions(None, r, 4, 1, None, r, 1, r, L.SK_MEM_HOST, None) == L.SK_ERR_ARGUMENT
    with pytest.raises(sk.CudaError):
        sk.Comm([0])


def test_bridge_and_policy_function_host_side_checks():
    """Argument checking of the Experimental mirror happens before any device work (src/experimental.jl:160-162)."""
    basis = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(2), 2)
    K = sk.dimension(basis)
    pt = {"a": None, "b": sk.bridge(sk.BoundedLinear(0, 1), sk.BoundedLinear(-1, 1))}
    with pytest.raises(sk.ArgumentError):
        sk.make_policy_functions(pt, basis, np.zeros(2 * K - 1))
    with pytest.raises(sk.ArgumentError):
        sk.bridge(sk.BoundedLinear(0, 1), "not a transformation")
    pf = sk.make_policy_functions(pt, basis, np.zeros(2 * K))
    assert pf.names == ["a", "b"] and pf.M == 2 and pf.K == K
    b = sk.bridge(sk.BoundedLinear(0, 1), sk.SemiInfRational(0, 2)).inverse()
    assert isinstance(b.outer, sk.SemiInfRational) and isinstance(b.inner, sk.BoundedLinear)
    assert sk.experimental.policy_coefficients_dimension(pt, basis) == 2 * K
