"""Reference-run golden vectors (tests/golden/julia/out/, produced by tests/golden/make_golden.jl with the real
SpectralKit.jl).  Julia is not available where this repository was built, so the outputs are not committed and these
tests are SKIPPED until someone runs

    julia --project=/path/to/SpectralKit.jl tests/golden/make_golden.jl

after which they pin (i) the CPU oracle and (ii) the CUDA library against the reference itself on identical inputs:
bit for bit for index sets, grids, basis values and the reference-order linear combinations, within 1e-13·Σ|θ_k b_k|
for the fast kernels.  Inputs: tests/golden/julia/in (committed; tests/golden/make_julia_inputs.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "golden"))
IN = os.path.join(HERE, "golden", "julia", "in")
OUT = os.path.join(HERE, "golden", "julia", "out")
HAVE = os.path.exists(os.path.join(OUT, "MANIFEST"))
needs_julia = pytest.mark.skipif(not HAVE, reason="no reference-run vectors: run tests/golden/make_golden.jl with Julia + SpectralKit.jl")

MIXED6 = [(1, 0.0, 4.0), (1, -2.0, 2.0), (2, 1.0, 2.0), (3, 0.0, 4.0), (1, -1.0, 2.0), (2, -3.0, 3.0)]


def inp(name, *shape):
    a = np.fromfile(os.path.join(IN, name + ".f64"), dtype="<f8")
    return a.reshape(shape) if shape else a


def manifest():
    m = {}
    for line in open(os.path.join(OUT, "MANIFEST")):
        f = line.split()
        if len(f) >= 3:
            m[f[0]] = (f[1], tuple(int(v) for v in f[2:]))
    return m


def ref(name):
    kind, shape = manifest()[name]
    a = np.fromfile(os.path.join(OUT, f"{name}.{kind}"), dtype="<f8" if kind == "f64" else "<i8")
    return a.reshape(shape)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_inputs_are_committed():
    """The inputs the Julia script reads are part of the repository (and regenerate deterministically)."""
    for name in ("c1_theta", "c1_x", "c2_theta", "c2_y", "c3_theta", "c3s_y", "c3i_y", "c4_theta", "c4a_x", "c4b_y", "c5_x", "b3_x"):
        assert os.path.getsize(os.path.join(IN, name + ".f64")) > 0
    assert os.path.exists(os.path.join(HERE, "golden", "make_golden.jl"))
    import make_julia_inputs as mj
    th = mj.c5_theta()
    assert th.shape == (77505, 4) and th[0, 0] == ((7919 + 104729) % 2003 - 1001) / 1024.0


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle as o
    return o


def otrs(orc):
    ctor = {1: orc.BoundedLinear, 2: orc.SemiInfRational, 3: orc.InfRational}
    return [ctor[k](a, b) for k, a, b in MIXED6]


# ------------------------------------------------------------------ the oracle against the reference (CPU)
@needs_julia
def test_oracle_grids_and_indices_match_julia(orc):
    for code in (0, 1, 2):
        for N in (20, 31, 63, 128, 243):
            assert same(orc.cheb_grid(code, N), ref(f"grid_g{code}_N{N}")), f"grid kind {code}, N={N}"
        for d, B, M in ((2, 3, 3), (3, 3, 2), (4, 2, 2), (3, 4, 4)):
            assert same(orc.smolyak_indices(code, d, B, M), ref(f"idx_g{code}_d{d}_B{B}_M{M}"))
            assert same(orc.smolyak_grid(code, d, B, M), ref(f"sgrid_g{code}_d{d}_B{B}_M{M}"))
    assert same(orc.smolyak_indices(2, 6, 4, 4), ref("c4_indices"))
    assert same(orc.smolyak_grid(2, 6, 4, 4), ref("c4_grid"))


@needs_julia
def test_oracle_univariate_matches_julia(orc):
    assert same(orc.uni_lincomb(20, inp("c1_theta"), inp("c1_x"))[:, 0], ref("c1_val"))
    assert same(orc.uni_basis_at(20, orc.cheb_grid(0, 20))[:, :, 0].T, ref("c1_colloc"))
    assert same(orc.uni_lincomb(64, inp("c2_theta"), inp("c2_y"), tr=orc.BoundedLinear(-2, 3), D=1), ref("c2_out"))
    for tag, t in (("c3s", orc.SemiInfRational(0.7, 0.3)), ("c3i", orc.InfRational(0.4, 0.9))):
        assert same(orc.uni_lincomb(128, inp("c3_theta"), inp(tag + "_y"), tr=t, D=1), ref(tag + "_out"))
        assert same(orc.uni_lincomb(128, inp("c3_theta"), inp(tag + "_y"), tr=t)[:, 0], ref(tag + "_val"))


@needs_julia
def test_oracle_smolyak_matches_julia(orc):
    th = inp("c4_theta")
    spec = orc.gradient_spec(6)
    for tag, pts, trs in (("c4a", inp("c4a_x", 1500, 6), None), ("c4b", inp("c4b_y", 1500, 6), otrs(orc))):
        assert same(orc.smolyak_lincomb(2, 6, 4, 4, th, pts, trs=trs)[:, 0], ref(tag + "_val"))
        assert same(orc.smolyak_lincomb(2, 6, 4, 4, th, pts, trs=trs, spec=spec), ref(tag + "_grad"))
    import make_julia_inputs as mj
    assert same(orc.smolyak_lincomb(2, 10, 5, 5, mj.c5_theta(), inp("c5_x", 64, 10)), ref("c5_out"))
    x = inp("b3_x", 40, 3)
    assert same(orc.smolyak_basis_at(2, 3, 3, 3, x)[:, :, 0].T, ref("b3_val"))
    assert same(orc.smolyak_basis_at(2, 3, 3, 3, x, spec=orc.gradient_spec(3)).transpose(1, 0, 2), ref("b3_grad"))


# ------------------------------------------------------------------ the CUDA library against the reference
@needs_julia
@pytest.mark.gpu
def test_gpu_matches_julia(orc):
    import __graft_entry__ as ge
    sk = ge.load_package()
    G = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}
    for code in (0, 1, 2):
        for N in (20, 31, 63, 128, 243):
            assert same(sk.grid(sk.Chebyshev(G[code], N)), ref(f"grid_g{code}_N{N}"))
        for d, B, M in ((2, 3, 3), (3, 3, 2), (4, 2, 2), (3, 4, 4)):
            b = sk.smolyak_basis(sk.Chebyshev, G[code], sk.SmolyakParameters(B, M), d)
            assert same(sk.smolyak_indices(b), ref(f"idx_g{code}_d{d}_B{B}_M{M}"))
            assert same(sk.grid(b), ref(f"sgrid_g{code}_d{d}_B{B}_M{M}"))
    # univariate: exact mode bit for bit, fast mode within tolerance
    b1 = sk.Chebyshev(sk.InteriorGrid(), 20)
    assert same(sk.linear_combination(b1, inp("c1_theta"), inp("c1_x"), exact=True), ref("c1_val"))
    assert same(sk.collocation_matrix(b1), ref("c1_colloc"))
    b2 = sk.Chebyshev(sk.InteriorGrid(), 64) @ sk.BoundedLinear(-2, 3)
    assert same(sk.linear_combination(b2, inp("c2_theta"), sk.d(inp("c2_y")), exact=True), ref("c2_out"))
    for tag, t in (("c3s", sk.SemiInfRational(0.7, 0.3)), ("c3i", sk.InfRational(0.4, 0.9))):
        b3 = sk.Chebyshev(sk.InteriorGrid(), 128) @ t
        assert same(sk.linear_combination(b3, inp("c3_theta"), sk.d(inp(tag + "_y")), exact=True), ref(tag + "_out"))
    # Smolyak cfg 4a / 4b
    ctor = {1: sk.BoundedLinear, 2: sk.SemiInfRational, 3: sk.InfRational}
    ct = sk.coordinate_transformations(*[ctor[k](a, b) for k, a, b in MIXED6])
    b6 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
    th = inp("c4_theta")
    spec = orc.gradient_spec(6)
    for tag, basis, pts, trs in (("c4a", b6, inp("c4a_x", 1500, 6), None), ("c4b", b6 @ ct, inp("c4b_y", 1500, 6), otrs(orc))):
        want = ref(tag + "_grad")
        assert same(sk.linear_combination(basis, th, sk.gradient(6)(pts), exact=True), want)
        scale = orc.smolyak_lincomb(2, 6, 4, 4, th, pts, trs=trs, spec=spec, absmode=True)
        fast = sk.linear_combination(basis, th, sk.gradient(6)(pts))
        assert np.max(np.abs(fast - want) / scale) <= 1e-13          # tolerance of north_star, per derivative order
        assert same(sk.linear_combination(basis, th, pts, exact=True), ref(tag + "_val"))
    # cfg 5 reduced (SVector coefficients) and basis_at
    import make_julia_inputs as mj
    b10 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(5), 10)
    th5, x5 = mj.c5_theta(), inp("c5_x", 64, 10)
    want = ref("c5_out")
    scale = orc.smolyak_lincomb(2, 10, 5, 5, th5, x5, absmode=True)
    assert np.max(np.abs(sk.linear_combination(b10, th5, x5) - want) / scale) <= 1e-13
    bb = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(3), 3)
    x = inp("b3_x", 40, 3)
    assert same(np.asarray(sk.basis_at(bb, x)), ref("b3_val"))
    assert same(np.asarray(sk.basis_at(bb, sk.gradient(3)(x))), ref("b3_grad"))
