"""Writes the INPUTS of the reference-run golden vectors (tests/golden/julia/in/*.f64, raw little-endian Float64).

The reference (tpapp/SpectralKit.jl) is Julia and cannot run in the build container, so its outputs are not
committed.  Anyone with Julia closes the gap in one command:

    julia --project=/path/to/SpectralKit.jl tests/golden/make_golden.jl

which reads these inputs, evaluates them with the reference itself and writes tests/golden/julia/out/.
tests/test_julia_golden.py then compares the oracle (CPU) and the CUDA library against those outputs; it is
reported as skipped while out/ is absent.  Run from the repo root: python tests/golden/make_julia_inputs.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IN = os.path.join(HERE, "julia", "in")
MIXED6 = [(1, 0.0, 4.0), (1, -2.0, 2.0), (2, 1.0, 2.0), (3, 0.0, 4.0), (1, -1.0, 2.0), (2, -3.0, 3.0)]   # cfg 4b


def from_pm1(kind, a, b, x):
    """transform_from (src/transformations.jl:178-181, 244, 309) — only used to place inputs inside the domains."""
    if kind == 1:
        s, m = (b - a) / 2, (a + b) / 2
        return x * s + m
    if kind == 2:
        return a + (b * (1 + x)) / (1 - x)
    return a + (b * x) / np.sqrt(1 - x * x)


def rand_pm1(rng, shape):
    """test/utilities.jl:28: about 10 % probability mass on each endpoint"""
    return np.clip((rng.random(shape) - 0.5) * 2.5, -1, 1)


def c5_theta(K=77505, nvec=4):
    """θ[k, v] = (((7919·k + 104729·v) mod 2003) − 1001) / 1024 with 1-based k, v (make_golden.jl uses the same formula)."""
    k = np.arange(1, K + 1, dtype=np.int64)[:, None]
    v = np.arange(1, nvec + 1, dtype=np.int64)[None, :]
    return (((7919 * k + 104729 * v) % 2003) - 1001) / 1024.0


def write(name, a):
    np.ascontiguousarray(a, dtype="<f8").tofile(os.path.join(IN, name + ".f64"))


def main():
    os.makedirs(IN, exist_ok=True)
    rng = np.random.default_rng(20261020)
    # c1: Chebyshev(InteriorGrid(), 20), values at 1000 points
    write("c1_theta", rng.standard_normal(20)); write("c1_x", rand_pm1(rng, 1000))
    # c2: Chebyshev(InteriorGrid(), 64) ∘ BoundedLinear(-2, 3), value + first derivative
    write("c2_theta", rng.standard_normal(64)); write("c2_y", from_pm1(1, -2.0, 3.0, rand_pm1(rng, 2000)))
    # c3: N = 128 ∘ SemiInfRational(0.7, 0.3) / InfRational(0.4, 0.9), value + first derivative; ±Inf included
    write("c3_theta", rng.standard_normal(128))
    u = rng.uniform(-0.99, 0.99, 2000)
    ys = from_pm1(2, 0.7, 0.3, u); ys[:2] = [np.inf, -np.inf]
    yi = from_pm1(3, 0.4, 0.9, u); yi[:2] = [np.inf, -np.inf]
    write("c3s_y", ys); write("c3i_y", yi)
    # c4: smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(4), 6), K = 2561, value + gradient, 1500 points
    write("c4_theta", rng.standard_normal(2561))
    u = rng.uniform(-0.99, 0.99, (1500, 6))
    write("c4a_x", u)                                                     # untransformed basis: points in [-1, 1]^6
    y = np.stack([from_pm1(k, a, b, u[:, j]) for j, (k, a, b) in enumerate(MIXED6)], axis=1)
    write("c4b_y", y)
    # c5: smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(5), 10), K = 77505, SVector{4} coefficients, 64 points
    # (its 77505 × 4 coefficients are the closed form c5_theta(k, v) below — exact in binary64, the same in both languages)
    write("c5_x", rng.uniform(-1, 1, (64, 10)))
    # basis_at: d = 3, B = 3 (InteriorGrid2, K = 111... whatever dimension(basis) is), 40 points, values and gradient
    write("b3_x", rand_pm1(rng, (40, 3)))
    print("wrote", sorted(os.listdir(IN)))


if __name__ == "__main__":
    main()
