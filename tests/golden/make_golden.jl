# Reference-run golden vectors.  Run with the reference package itself:
#
#     julia --project=/path/to/SpectralKit.jl tests/golden/make_golden.jl
#
# Reads tests/golden/julia/in/*.f64 (written by tests/golden/make_julia_inputs.py, committed), evaluates every case
# with tpapp/SpectralKit.jl (v0.16.x) and writes tests/golden/julia/out/*.f64 / *.i64 (raw little-endian) plus a
# MANIFEST.  tests/test_julia_golden.py picks the outputs up.  This file could not be executed where it was written
# (no Julia in the build container); it uses only the exported API and StaticArrays.
#
# Layout convention: an output "n × E" is written point-major (the E numbers of point 1, then point 2, …), i.e. what
# numpy reads back with reshape(n, E).
using SpectralKit
using StaticArrays

const HERE = @__DIR__
const IN = joinpath(HERE, "julia", "in")
const OUT = joinpath(HERE, "julia", "out")
mkpath(OUT)
const MANIFEST = String[]

readf64(name) = reinterpret(Float64, read(joinpath(IN, name * ".f64")))
function writef64(name, a, shape)
    open(io -> write(io, htol.(collect(Float64, vec(a)))), joinpath(OUT, name * ".f64"), "w")
    push!(MANIFEST, "$(name) f64 $(join(shape, ' '))")
end
function writei64(name, a, shape)
    open(io -> write(io, htol.(collect(Int64, vec(a)))), joinpath(OUT, name * ".i64"), "w")
    push!(MANIFEST, "$(name) i64 $(join(shape, ' '))")
end
# point-major copy of a vector of 𝑑Expansion / ∂Expansion / SVector results
rows(v) = reduce(hcat, [collect(Float64, r) for r in v])          # E × n, column j = point j
coeffs(r) = r.coefficients

# ---- c1: Chebyshev(InteriorGrid(), 20): linear_combination at 1000 points + collocation matrix at its own grid
let basis = Chebyshev(InteriorGrid(), 20), θ = collect(readf64("c1_theta")), x = readf64("c1_x")
    writef64("c1_val", [linear_combination(basis, θ, xi) for xi in x], (length(x),))
    C = collocation_matrix(basis)                                  # 20 × 20, C[i, j] = basis function j at grid point i
    writef64("c1_colloc", permutedims(C), size(C))                 # written row by row
end

# ---- c2: N = 64 ∘ BoundedLinear(-2, 3), value + first derivative
let basis = Chebyshev(InteriorGrid(), 64) ∘ BoundedLinear(-2.0, 3.0), θ = collect(readf64("c2_theta")), y = readf64("c2_y")
    writef64("c2_out", rows([coeffs(linear_combination(basis, θ, 𝑑(yi))) for yi in y]), (length(y), 2))
end

# ---- c3: N = 128 ∘ SemiInfRational(0.7, 0.3) / InfRational(0.4, 0.9), value + first derivative (±Inf are the first two points)
let θ = collect(readf64("c3_theta"))
    for (tag, t) in (("c3s", SemiInfRational(0.7, 0.3)), ("c3i", InfRational(0.4, 0.9)))
        basis = Chebyshev(InteriorGrid(), 128) ∘ t
        y = readf64(tag * "_y")
        writef64(tag * "_out", rows([coeffs(linear_combination(basis, θ, 𝑑(yi))) for yi in y]), (length(y), 2))
        writef64(tag * "_val", [linear_combination(basis, θ, yi) for yi in y], (length(y),))
    end
end

# ---- c4: Smolyak d = 6, B = 4, InteriorGrid2: values and value + gradient, untransformed (4a) and mixed transformations (4b)
let basis = smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(4), 6), θ = collect(readf64("c4_theta"))
    @assert dimension(basis) == 2561
    D = ∂(1) ∪ ∂(0, 1) ∪ ∂(0, 0, 1) ∪ ∂(0, 0, 0, 1) ∪ ∂(0, 0, 0, 0, 1) ∪ ∂(0, 0, 0, 0, 0, 1)
    ct = coordinate_transformations(BoundedLinear(0.0, 4.0), BoundedLinear(-2.0, 2.0), SemiInfRational(1.0, 2.0),
                                    InfRational(0.0, 4.0), BoundedLinear(-1.0, 2.0), SemiInfRational(-3.0, 3.0))
    for (tag, b, file) in (("c4a", basis, "c4a_x"), ("c4b", basis ∘ ct, "c4b_y"))
        X = reshape(readf64(file), 6, :)                            # column j = point j (the file is point-major)
        pts = [Tuple(X[:, j]) for j in axes(X, 2)]
        writef64(tag * "_val", [linear_combination(b, θ, p) for p in pts], (length(pts),))
        writef64(tag * "_grad", rows([coeffs(linear_combination(b, θ, D(p))) for p in pts]), (length(pts), 7))
    end
    writei64("c4_indices", reduce(hcat, [collect(Int64, i) for i in basis.smolyak_indices]), (2561, 6))
    writef64("c4_grid", reduce(hcat, [collect(Float64, g) for g in grid(basis)]), (2561, 6))
end

# ---- c5: Smolyak d = 10, B = 5, InteriorGrid2 with SVector{4} coefficients (test/test_generic_api.jl:95-102)
let basis = smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(5), 10)
    K = dimension(basis)
    @assert K == 77505
    θ = [SVector{4,Float64}(ntuple(v -> (mod(7919 * k + 104729 * v, 2003) - 1001) / 1024, 4)) for k in 1:K]
    X = reshape(readf64("c5_x"), 10, :)
    writef64("c5_out", rows([linear_combination(basis, θ, Tuple(X[:, j])) for j in axes(X, 2)]), (size(X, 2), 4))
end

# ---- basis_at: d = 3, B = 3, InteriorGrid2, 40 points: values (n × K) and gradient-valued entries (n × K × 4)
let basis = smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(3), 3)
    K = dimension(basis)
    X = reshape(readf64("b3_x"), 3, :)
    D = ∂(1) ∪ ∂(0, 1) ∪ ∂(0, 0, 1)
    vals = reduce(hcat, [collect(Float64, basis_at(basis, Tuple(X[:, j]))) for j in axes(X, 2)])              # K × n
    writef64("b3_val", vals, (size(X, 2), K))
    grads = reduce(hcat, [reduce(vcat, [collect(Float64, coeffs(e)) for e in basis_at(basis, D(Tuple(X[:, j])))]) for j in axes(X, 2)])   # 4K × n
    writef64("b3_grad", grads, (size(X, 2), K, 4))
end

# ---- grids (chebyshev.jl:88-108: Julia Base sinpi/cospi) and Smolyak index sets / grids of the reference's test matrix
for (code, kind) in ((0, InteriorGrid()), (1, EndpointGrid()), (2, InteriorGrid2()))
    for N in (20, 31, 63, 128, 243)
        writef64("grid_g$(code)_N$(N)", collect(grid(Chebyshev(kind, N))), (N,))
    end
    for (d, B, M) in ((2, 3, 3), (3, 3, 2), (4, 2, 2), (3, 4, 4))
        basis = smolyak_basis(Chebyshev, kind, SmolyakParameters(B, M), d)
        idx = reduce(hcat, [collect(Int64, i) for i in basis.smolyak_indices])                              # d × K
        writei64("idx_g$(code)_d$(d)_B$(B)_M$(M)", idx, (size(idx, 2), d))
        writef64("sgrid_g$(code)_d$(d)_B$(B)_M$(M)", reduce(hcat, [collect(Float64, g) for g in grid(basis)]), (size(idx, 2), d))
    end
end

open(io -> foreach(l -> println(io, l), MANIFEST), joinpath(OUT, "MANIFEST"), "w")
println("wrote ", length(MANIFEST), " arrays to ", OUT, " with SpectralKit ", pkgversion(SpectralKit))
