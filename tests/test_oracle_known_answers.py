"""Pin the oracle against every known-answer value the reference contains.

Julia cannot run here, so these docstring/test constants are the only reference-produced
numbers available; the oracle must reproduce each one exactly as the reference states it.
"""
import math

import numpy as np
import pytest

INT, END, INT2 = 0, 1, 2


def test_coordinate_transformations_docstring(orc):
    # src/transformations.jl:88-104
    trs = [orc.BoundedLinear(0, 2), orc.SemiInfRational(2, 3)]
    x = orc.transform_from_batch(trs, np.array([[0.4, 0.5]]))
    assert x.tolist() == [[1.4, 11.0]]
    y = orc.transform_to_batch(trs, x)[:, :, 0]
    assert y.tolist() == [[0.3999999999999999, 0.5]]


def test_semiinf_docstring_mappings(orc):
    # src/transformations.jl:236-240: -1/2 -> A + L/3, 0 -> A + L, 1/2 -> A + 3L
    A, L = 2.0, 3.0
    t = orc.SemiInfRational(A, L)
    assert orc.transform_from(t, -0.5) == pytest.approx(A + L / 3, rel=1e-15)
    assert orc.transform_from(t, 0.0) == A + L
    assert orc.transform_from(t, 0.5) == A + 3 * L


def test_infrational_docstring_mappings(orc):
    # src/transformations.jl:302-305: 0 -> A, ±0.5 -> A ± L/√3
    A, L = 0.7, 1.3
    t = orc.InfRational(A, L)
    assert orc.transform_from(t, 0.0) == A
    for s in (1, -1):
        assert orc.transform_from(t, s * 0.5) == pytest.approx(A + s * L / math.sqrt(3), rel=1e-15)


def test_smolyak_docstring_dimension(orc):
    # src/smolyak_api.jl:46-49
    assert orc.smolyak_length(INT, 2, 3, 3) == 81
    assert orc.smolyak_parent_dimension(INT, 3, 3) == 27


def test_survey_target_sizes(orc):
    # cfg 4 and cfg 5 sizes, K_j tables
    assert orc.smolyak_length(INT2, 6, 4, 4) == 2561
    assert orc.smolyak_parent_dimension(INT2, 4, 4) == 31
    assert [orc.smolyak_length(INT2, j, 4, 4) for j in range(6, 0, -1)] == [2561, 1471, 769, 351, 129, 31]
    assert orc.smolyak_length(INT2, 10, 5, 5) == 77505
    assert orc.smolyak_length(END, 10, 5, 5) == 41265
    assert orc.smolyak_length(INT, 10, 5, 5) == 141165


def test_jet_seed_docstring(orc):
    # src/derivatives.jl:80-86: 𝑑(2.0) = (2, 1); (𝑑^3)(2.0) = (2, 1, 0, 0) — through the identity map
    t = orc.Identity()
    assert orc.transform_to(t, 2.0, D=1).tolist() == [2.0, 1.0]
    assert orc.transform_to(t, 2.0, D=3).tolist() == [2.0, 1.0, 0.0, 0.0]


def test_chebyshev_jet_docstring(orc):
    # src/derivatives.jl:170-177: Chebyshev(InteriorGrid(), 3) at the jet (0.1, 1)
    tab = orc.cheb_basis_at(3, [0.1, 1.0])
    assert tab[0].tolist() == [1.0, 0.0]
    assert tab[1].tolist() == [0.1, 1.0]
    assert tab[2].tolist() == [-0.98, 0.4]


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_infinite_endpoint_limits(orc, sign):
    # test/test_chebyshev.jl:109-146 and test/test_transformations.jl:50-57,85-92
    N = 10
    y = np.array([sign * math.inf])
    for tr, minus_val in ((orc.SemiInfRational(2.3, 0.7), lambda i: 1.0),
                          (orc.InfRational(0.3, 3.0), lambda i: (-1.0) ** (i + 1))):
        x = orc.transform_to(tr, sign * math.inf, D=1)
        assert x[1] == 0
        for i in range(1, N + 1):
            theta = np.zeros(N)
            theta[i - 1] = 1.0
            out = orc.uni_lincomb(N, theta, y, tr=tr, D=1)[0]
            want = 1.0 if sign > 0 else minus_val(i)
            assert out[0] == want
            assert out[1] == 0


def test_endpoint_grid_exact_ends(orc):
    # test/test_chebyshev.jl:42-43; src/chebyshev.jl:97-98
    for N in range(2, 11):
        g = orc.cheb_grid(END, N)
        assert g[0] == -1 and g[-1] == 1
    assert orc.cheb_grid(END, 1).tolist() == [0.0]


def test_smolyak_parameters_clamp(orc):
    # src/smolyak_traversal.jl:15-16; test/test_smolyak.jl:10 — M > B becomes M = B
    assert orc.smolyak_length(INT, 2, 2, 4) == orc.smolyak_length(INT, 2, 2, 2)
    assert (orc.smolyak_indices(INT, 2, 2, 4) == orc.smolyak_indices(INT, 2, 2, 2)).all()


def test_shuffle_examples(orc):
    # SURVEY.md a22 / Appendix A.8 restated shuffles
    assert orc.grid_shuffle(INT2, 7).tolist() == [4, 2, 6, 1, 3, 5, 7]
    assert orc.grid_shuffle(END, 9).tolist() == [5, 1, 9, 3, 7, 2, 4, 6, 8]
    assert orc.grid_shuffle(INT, 9).tolist() == [5, 2, 8, 1, 3, 4, 6, 7, 9]
    assert orc.grid_shuffle(INT2, 15).tolist() == [8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15]


def test_transform_constructor_domain_errors(orc):
    # test/test_transformations.jl:4-5,27-29,61-64
    bad = [(orc.BoundedLinear, (-1.0, math.inf)), (orc.BoundedLinear, (-1.0, -2.0)),
           (orc.SemiInfRational, (-1.0, math.inf)), (orc.SemiInfRational, (-1.0, 0.0)),
           (orc.SemiInfRational, (math.nan, 2.0)), (orc.InfRational, (1.0, math.inf)),
           (orc.InfRational, (1.0, 0.0)), (orc.InfRational, (1.0, -2.0)), (orc.InfRational, (math.nan, 1.0))]
    for ctor, args in bad:
        with pytest.raises(orc.OracleError) as e:
            ctor(*args)
        assert e.value.code == orc.ERR_DOMAIN


def test_rational_second_derivative_throws_without_extension(orc):
    # src/transformations.jl:266,332
    for tr in (orc.SemiInfRational(0.7, 0.3), orc.InfRational(0.4, 0.9)):
        with pytest.raises(orc.OracleError) as e:
            orc.transform_to(tr, 1.0, D=2)
        assert e.value.code == orc.ERR_UNSUPPORTED
        assert len(orc.transform_to(tr, 1.0, D=2, ext2=True)) == 3


def test_gradient_spec_order(orc):
    # SURVEY.md Appendix A.5 worked example: [value, ∂1, …, ∂6]
    spec = orc.gradient_spec(6)
    want = np.vstack([np.zeros((1, 6), dtype=int), np.eye(6, dtype=int)])
    assert (spec == want).all()
    assert orc.partials_degrees(orc.partials_minimal([tuple([0] * j + [1]) for j in range(6)], 6), 6).tolist() == [1] * 6


def test_partials_examples(orc):
    # src/derivatives.jl:396-401: ∂(1,2) ∪ ∂(2,1) prints union(∂(2,1), ∂(1,2)) (descending order)
    m = orc.partials_minimal([(1, 2), (2, 1)], 2)
    assert m.tolist() == [[2, 1], [1, 2]]
    # nested specification collapses: ∂(1) ∪ ∂(1,1) == ∂(1,1)
    assert orc.partials_minimal([(1,), (1, 1)], 2).tolist() == [[1, 1]]
    # ∂(1,1): CartesianIndices order, first index fastest
    assert orc.deriv_spec([(1, 1)], 2).tolist() == [[0, 0], [1, 0], [0, 1], [1, 1]]
