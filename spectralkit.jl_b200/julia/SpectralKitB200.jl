# SpectralKitB200.jl — Julia `ccall` shim over libspectralkit_b200.so (include/spectralkit_b200.h).
#
# STATUS: written against the C ABI but NEVER EXECUTED — Julia is not installed in the build image
# or on the GPU box (SURVEY.md §0 fact 2).  The executable harness with the same semantics is the
# Python mirror (spectralkit.jl_b200/api.py); every call below has a one-to-one counterpart there.
#
# What it does: keeps SpectralKit.jl's own types (bases, transformations, 𝑑/∂ requests are built by
# the reference's constructors, so every DomainError/ArgumentError of a constructor is the
# reference's own) and adds BATCHED methods that hand whole point collections to the GPU:
#
#     using SpectralKit, SpectralKitB200
#     basis = smolyak_basis(Chebyshev, InteriorGrid2(), SmolyakParameters(4), 6) ∘ ct
#     θ  = randn(dimension(basis))
#     xs = [ntuple(_ -> rand(), 6) for _ in 1:10^6]          # Vector{NTuple{6,Float64}}
#     D  = ∂gradient(6)
#     ys = linear_combination.(basis, Ref(θ), D.(xs))         # the reference's own idiom (test/test_generic_api.jl:22):
#                                                             # the broadcast is intercepted and runs as ONE GPU batch
#     ys = linear_combination(basis, θ, D, xs)                # the same, explicit batched method
#     C  = collocation_matrix(basis, xs)                      # Matrix{Float64} (generic_api.jl:217), on the GPU
#
# Memory: Julia's `Vector{NTuple{d,Float64}}`, `Vector{SVector{M,Float64}}`, `d×n Matrix{Float64}`
# and `Matrix{Float64}` already have the layouts the C ABI takes, so pointers are passed through
# (GC.@preserve) without copies on the Julia side.
module SpectralKitB200

using SpectralKit
using SpectralKit: Chebyshev, SmolyakBasis, SmolyakIndices, TransformedBasis, BoundedLinear,
    SemiInfRational, InfRational, CoordinateTransformations, InteriorGrid, EndpointGrid, InteriorGrid2,
    𝑑Derivatives, ∂Derivatives, Partials, _partials_canonical_expansion
using StaticArrays

export ∂gradient, collocation_matrix_gpu, basis_at_gpu, collocation_solve_gpu, SKPlan

const libsk = get(ENV, "SPECTRALKIT_B200_LIB", joinpath(@__DIR__, "..", "lib", "libspectralkit_b200.so"))

# ---- C structs (include/spectralkit_b200.h) ----------------------------------------------------
struct SkBasisDesc
    kind::Int32; grid_kind::Int32; N::Int32; d::Int32; B::Int32; M::Int32
end
struct SkTransform
    kind::Int32; reserved::Int32; p0::Float64; p1::Float64
end
struct SkDerivSpec
    order::Int32; ncomp::Int32; orders::Ptr{Int32}; allow_rational_d2::Int32
end

const SK_MEM_HOST, SK_MEM_DEVICE, SK_MODE_FAST, SK_MODE_EXACT = UInt32(0), UInt32(1), UInt32(0), UInt32(2)

# status codes → the reference's exception types (SURVEY.md §8(b) "Errors")
function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:sk_last_error, libsk), Cstring, ()))
    rc == -1 && throw(ArgumentError(msg))
    rc == -2 && throw(DomainError(msg))
    rc == -3 && error(msg)                       # "… not implemented yet", MethodError-like cases
    error("spectralkit_b200: CUDA failure or no device (status $rc): $msg")   # no CPU fallback
end

grid_code(::InteriorGrid) = Int32(0)
grid_code(::EndpointGrid) = Int32(1)
grid_code(::InteriorGrid2) = Int32(2)

desc(b::Chebyshev) = SkBasisDesc(0, grid_code(b.grid_kind), b.N, 1, 0, 0)
function desc(b::SmolyakBasis{<:SmolyakIndices{N,H,B,M}}) where {N,H,B,M}
    SkBasisDesc(1, grid_code(b.univariate_parent.grid_kind), 0, N, B, M)
end

# the reference's stored fields travel unchanged (src/transformations.jl:150-153, 207-210, 281-284)
sk(t::BoundedLinear) = SkTransform(1, 0, t.m, t.s)
sk(t::SemiInfRational) = SkTransform(2, 0, t.A, t.L)
sk(t::InfRational) = SkTransform(3, 0, t.A, t.L)

split(b::TransformedBasis) = (b.parent, b.transformation isa CoordinateTransformations ?
                              SkTransform[sk(t) for t in Tuple(b.transformation)] : SkTransform[sk(b.transformation)])
split(b) = (b, nothing)

"`∂(1) ∪ ∂(0,1) ∪ …`: value and all first partials of an `n`-dimensional argument."
∂gradient(n::Int) = reduce(∪, (∂(ntuple(j -> j == i ? 1 : 0, i)...) for i in 1:n))

# canonical expansion of a ∂ specification as the Int32 table the ABI takes (derivatives.jl:284-300)
function order_table(::∂Derivatives{Ps}, N::Int) where {Ps}
    exp = collect(_partials_canonical_expansion(Val(N), fieldtypes(Ps)))
    Int32[exp[c][j] for j in 1:N, c in eachindex(exp)]          # column c = component c: [ncomp][N] in C order
end

# ---- plans --------------------------------------------------------------------------------------
mutable struct SKPlan
    handle::Ptr{Cvoid}
    E::Int
    K::Int
    function SKPlan(basis; order::Int = 0, spec::Union{Nothing,∂Derivatives} = nothing, device::Int = 0,
                    allow_rational_d2::Bool = false)
        parent, trs = split(basis)
        d = desc(parent)
        table = spec === nothing ? Int32[] : vec(order_table(spec, Int(d.d)))
        ncomp = spec === nothing ? 0 : length(table) ÷ d.d
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve table begin
            ds = SkDerivSpec(order, ncomp, isempty(table) ? C_NULL : pointer(table), allow_rational_d2)
            check(ccall((:sk_plan_create, libsk), Cint,
                        (Ref{SkBasisDesc}, Ptr{SkTransform}, Ref{SkDerivSpec}, Cint, Ref{Ptr{Cvoid}}),
                        d, trs === nothing ? C_NULL : trs, ds, device, h))
        end
        E = spec === nothing ? order + 1 : max(ncomp, 1)
        p = new(h[], E, dimension(basis))
        finalizer(q -> ccall((:sk_plan_destroy, libsk), Cint, (Ptr{Cvoid},), q.handle), p)
        p
    end
end

# plan cache; the header promises re-entrancy, so the Dict is only touched under a lock (plans themselves are immutable
# and may be used from several tasks at once)
const PLANS = Dict{Any,SKPlan}()
const PLANS_LOCK = ReentrantLock()
plan_for(basis; kw...) = lock(PLANS_LOCK) do
    get!(() -> SKPlan(basis; kw...), PLANS, (basis, kw))
end

# ---- batched linear_combination (generic_api.jl:114-138) ----------------------------------------
# univariate: xs::AbstractVector{Float64}; `𝑑^D` requests derivatives (derivatives.jl:66-107)
function SpectralKit.linear_combination(basis, θ::Vector{Float64}, xs::Vector{Float64};
                                        D::𝑑Derivatives{Dn} = 𝑑Derivatives{0}(), exact::Bool = false) where {Dn}
    p = plan_for(basis; order = Dn)
    out = Matrix{Float64}(undef, Dn + 1, length(xs))          # column i = 𝑑Expansion of point i
    flags = SK_MEM_HOST | (exact ? SK_MODE_EXACT : SK_MODE_FAST)
    GC.@preserve θ xs out check(ccall((:sk_linear_combination, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Int64, Ptr{Float64}, UInt32, Ptr{Cvoid}),
        p.handle, θ, length(θ), 1, xs, length(xs), out, flags, C_NULL))
    Dn == 0 ? vec(out) : [SpectralKit.𝑑Expansion(SVector{Dn + 1}(view(out, :, i))) for i in axes(out, 2)]
end

# multivariate: xs::Vector{NTuple{N,Float64}} (or Vector{SVector{N}}); `spec` is a ∂ request or nothing
function SpectralKit.linear_combination(basis, θ::Vector{Float64}, spec::Union{Nothing,∂Derivatives},
                                        xs::Vector{<:Union{NTuple{N,Float64},SVector{N,Float64}}};
                                        exact::Bool = false) where {N}
    p = plan_for(basis; spec = spec)
    out = Matrix{Float64}(undef, p.E, length(xs))
    flags = SK_MEM_HOST | (exact ? SK_MODE_EXACT : SK_MODE_FAST)
    GC.@preserve θ xs out check(ccall((:sk_linear_combination, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Int64, Ptr{Float64}, UInt32, Ptr{Cvoid}),
        p.handle, θ, length(θ), 1, Ptr{Float64}(pointer(xs)), length(xs), out, flags, C_NULL))
    spec === nothing ? vec(out) :
        [SpectralKit.∂Expansion(spec, SVector{size(out, 1)}(view(out, :, i))) for i in axes(out, 2)]
end

# SVector coefficients (test/test_generic_api.jl:95-102): θ::Vector{SVector{M,Float64}}
function SpectralKit.linear_combination(basis, θ::Vector{SVector{M,Float64}}, xs::Vector{Float64}) where {M}
    p = plan_for(basis)
    out = Vector{SVector{M,Float64}}(undef, length(xs))
    GC.@preserve θ xs out check(ccall((:sk_linear_combination, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Int64, Ptr{Float64}, UInt32, Ptr{Cvoid}),
        p.handle, Ptr{Float64}(pointer(θ)), length(θ), M, xs, length(xs), Ptr{Float64}(pointer(out)), SK_MEM_HOST, C_NULL))
    out
end

# ---- the reference's calling convention for derivatives, batched ----------------------------------
# `(𝑑^D)(x)` (derivatives.jl:105-107) seeds one point; `(𝑑^D).(xs)` is a Vector{𝑑Expansion{D+1,Float64}} whose first
# coefficients are the points (the seeds (1, 0, …) are implied and re-created on the device).
function SpectralKit.linear_combination(basis, θ::Vector{Float64}, xs::Vector{SpectralKit.𝑑Expansion{Dp1,Float64}};
                                        exact::Bool = false) where {Dp1}
    SpectralKit.linear_combination(basis, θ, Float64[x.coefficients[1] for x in xs]; D = 𝑑Derivatives{Dp1 - 1}(), exact)
end
# `D(x)` with D::∂Derivatives (derivatives.jl:455) wraps one point; `D.(xs)` is a Vector{∂CoordinateExpansion}: the request
# travels in the element type, the points are the first coefficients of the wrapped coordinates.
function SpectralKit.linear_combination(basis, θ::Vector{Float64},
                                        xs::Vector{<:SpectralKit.∂CoordinateExpansion{D,<:NTuple{N,Any}}}; exact::Bool = false) where {D,N}
    isempty(xs) && return SpectralKit.∂Expansion[]
    pts = NTuple{N,Float64}[map(c -> c.coefficients[1], x.x) for x in xs]
    SpectralKit.linear_combination(basis, θ, first(xs).∂D, pts; exact)
end

# ---- broadcast hooks (SURVEY.md §8(b)): the reference API is pointwise and callers batch by broadcasting
# (`linear_combination.(basis, Ref(φ), y)`, test/test_generic_api.jl:22; `basis_at.(…)`, `l.(xs)` for a LinearCombination).
# These methods turn exactly those broadcasts over Vector arguments into one library call; anything else falls through
# to Base's broadcast and the reference's pointwise methods.
const _Points = Union{Vector{Float64},Vector{<:NTuple{N,Float64} where N},Vector{<:SVector{N,Float64} where N},
                      Vector{<:SpectralKit.𝑑Expansion},Vector{<:SpectralKit.∂CoordinateExpansion}}
_unref(x::Base.RefValue) = x[]
_unref(x) = x
function Base.Broadcast.broadcasted(::typeof(SpectralKit.linear_combination), basis::SpectralKit.FunctionBasis,
                                    θ::Union{Base.RefValue{Vector{Float64}},Tuple{Vector{Float64}}}, xs::_Points)
    θv = θ isa Tuple ? θ[1] : _unref(θ)
    xs isa Vector{<:Union{NTuple,SVector}} ? SpectralKit.linear_combination(basis, θv, nothing, xs) :
                                             SpectralKit.linear_combination(basis, θv, xs)
end
function Base.Broadcast.broadcasted(l::SpectralKit.LinearCombination{<:SpectralKit.FunctionBasis,Vector{Float64}}, xs::_Points)
    xs isa Vector{<:Union{NTuple,SVector}} ? SpectralKit.linear_combination(l.basis, l.θ, nothing, xs) :
                                             SpectralKit.linear_combination(l.basis, l.θ, xs)
end
# `collect.(basis_at.(basis, xs))` materialises the same numbers as a collocation matrix; the lazy per-point iterators of
# the reference have no batched counterpart, so the hook returns the rows of the matrix
function Base.Broadcast.broadcasted(::typeof(SpectralKit.basis_at), basis::SpectralKit.FunctionBasis,
                                    xs::Union{Vector{Float64},Vector{<:NTuple{N,Float64} where N},Vector{<:SVector{N,Float64} where N}})
    C = collocation_matrix_gpu(basis, xs)
    [view(C, i, :) for i in axes(C, 1)]
end

# ---- collocation_matrix / basis_at (generic_api.jl:217-227): column-major n×K, as the reference ----
function collocation_matrix_gpu(basis, xs::Vector{Float64} = collect(grid(basis)))
    p = plan_for(basis)
    C = Matrix{Float64}(undef, length(xs), p.K)
    GC.@preserve xs C check(ccall((:sk_basis_at, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, UInt32, Ptr{Cvoid}),
        p.handle, xs, length(xs), C, length(xs), SK_MEM_HOST, C_NULL))
    C
end
function collocation_matrix_gpu(basis, xs::Vector{<:Union{NTuple{N,Float64},SVector{N,Float64}}}) where {N}
    p = plan_for(basis)
    C = Matrix{Float64}(undef, length(xs), p.K)
    GC.@preserve xs C check(ccall((:sk_basis_at, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, UInt32, Ptr{Cvoid}),
        p.handle, Ptr{Float64}(pointer(xs)), length(xs), C, length(xs), SK_MEM_HOST, C_NULL))
    C
end
const basis_at_gpu = collocation_matrix_gpu

# the reference's own name (generic_api.jl:217): `collocation_matrix(basis, x)` for a Vector of points goes to the GPU,
# `collocation_matrix(basis)` evaluates at the basis' own grid.  (Deliberate method overriding for these argument types:
# this is what "drop-in" means; every other iterable of points keeps the reference's CPU method.)
SpectralKit.collocation_matrix(basis::SpectralKit.FunctionBasis, xs::Vector{Float64}) = collocation_matrix_gpu(basis, xs)
SpectralKit.collocation_matrix(basis::SpectralKit.FunctionBasis, xs::Vector{<:Union{NTuple{N,Float64},SVector{N,Float64}}}) where {N} =
    collocation_matrix_gpu(basis, xs)
function SpectralKit.collocation_matrix(basis::SpectralKit.FunctionBasis)
    g = collect(grid(basis))
    collocation_matrix_gpu(basis, g isa Vector{Float64} ? g : NTuple{length(first(g)),Float64}[Tuple(p) for p in g])
end

# ---- fitting: θ = collocation_matrix(basis) \ y on the basis' own grid (docs/src/index.md:92) ----
"`collocation_matrix(basis) \\ y` with the matrix generated and factorised on the GPU; `y` are the values at `grid(basis)`."
function collocation_solve_gpu(basis, y::Vector{Float64})
    p = plan_for(basis)
    length(y) == p.K || throw(ArgumentError("length(y) == dimension(basis) must hold."))
    θ = similar(y)
    GC.@preserve y θ check(ccall((:sk_collocation_solve, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, UInt32, Ptr{Cvoid}), p.handle, y, 1, θ, SK_MEM_HOST, C_NULL))
    θ
end
function collocation_solve_gpu(basis, y::Vector{SVector{M,Float64}}) where {M}
    p = plan_for(basis)
    length(y) == p.K || throw(ArgumentError("length(y) == dimension(basis) must hold."))
    θ = similar(y)
    GC.@preserve y θ check(ccall((:sk_collocation_solve, libsk), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, UInt32, Ptr{Cvoid}),
        p.handle, Ptr{Float64}(pointer(y)), M, Ptr{Float64}(pointer(θ)), SK_MEM_HOST, C_NULL))
    θ
end

# ---- refinement between bases (chebyshev.jl:133-142, smolyak_api.jl:141-208, generic_api.jl:314-322) ----
function is_subset_basis_library(basis1, basis2)
    p1, t1 = split(basis1); p2, t2 = split(basis2)
    out = Ref{Cint}(0)
    check(ccall((:sk_is_subset_basis, libsk), Cint,
                (Ref{SkBasisDesc}, Ptr{SkTransform}, Ref{SkBasisDesc}, Ptr{SkTransform}, Ref{Cint}),
                desc(p1), t1 === nothing ? C_NULL : t1, desc(p2), t2 === nothing ? C_NULL : t2, out))
    out[] != 0
end
function augment_coefficients_library(basis1, basis2, θ1::Vector{Float64})
    p1, t1 = split(basis1); p2, t2 = split(basis2)
    θ2 = Vector{Float64}(undef, dimension(basis2))
    GC.@preserve θ1 θ2 check(ccall((:sk_augment_coefficients, libsk), Cint,
        (Ref{SkBasisDesc}, Ptr{SkTransform}, Ref{SkBasisDesc}, Ptr{SkTransform}, Ptr{Float64}, Int64, Cint, Ptr{Float64}),
        desc(p1), t1 === nothing ? C_NULL : t1, desc(p2), t2 === nothing ? C_NULL : t2, θ1, length(θ1), 1, θ2))
    θ2
end

# ---- metadata that must be bit-exact (SURVEY.md §8(b)) -------------------------------------------
"collect(grid(basis)) computed by the library (correctly rounded sinpi/cospi); compare with `collect(grid(basis))`."
function grid_gpu_library(basis)
    parent, trs = split(basis)
    d = desc(parent)
    K = dimension(basis)
    out = d.kind == 0 ? Vector{Float64}(undef, K) : Matrix{Float64}(undef, Int(d.d), K)
    GC.@preserve out check(ccall((:sk_grid, libsk), Cint, (Ref{SkBasisDesc}, Ptr{SkTransform}, Ptr{Float64}),
                                 d, trs === nothing ? C_NULL : trs, out))
    out
end

end # module
