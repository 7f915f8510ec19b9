# Runs the REAL reference (lcauser/TensorNetworks.jl) on the seeded inputs of tests/golden/hotpath_golden.npz and writes its outputs
# next to them, so that the CPU oracle (oracle/*.py) and the CUDA path can be pinned to the reference itself:
#
#   python tests/golden/export_inputs_for_julia.py
#   julia --project=/path/to/TensorNetworks.jl tests/golden/make_golden_reference.jl
#   python -m pytest tests/test_reference_fixtures.py
#
# UNVERIFIED: the build image has no Julia.  The script only uses the reference's exported API:
#   GMPS(rank, dim, tensors, center)          structures/mps/gmps.jl:8-13
#   ProjMPS(psi, H, psi; rank=2, center)      structures/mps/projmps.jl:16-42,  product :103-145, calculate :192-216, block abstractprojmps.jl:33-46
#   svd(x, idx; cutoff, maxdim, mindim)       tensors.jl:168-227
#   dmrg(psi, H; maxdim, cutoff, maxsweeps)   algorithms/mps/dmrg.jl:128-154
# Inputs / outputs: raw little-endian column-major binaries + manifest.json (dtype c128 | f64 | i64, shape).
using TensorNetworks
using LinearAlgebra

const HERE = @__DIR__
const IN = joinpath(HERE, "julia_io", "inputs")
const OUT = joinpath(HERE, "julia_io", "outputs")
mkpath(OUT)

# a tiny JSON reader is avoided on purpose: the manifest is re-derived from a flat "name dtype d1 d2 ..." listing
function read_manifest()
    txt = read(joinpath(IN, "manifest.json"), String)
    man = Dict{String,Tuple{String,Vector{Int}}}()
    for m in eachmatch(r"\"([^\"]+)\":\s*\{\s*\"dtype\":\s*\"(\w+)\",\s*\"shape\":\s*\[([^\]]*)\]", txt)
        dims = isempty(strip(m.captures[3])) ? Int[] : parse.(Int, split(m.captures[3], ","))
        man[m.captures[1]] = (m.captures[2], dims)
    end
    man
end
const MAN = read_manifest()
function load(name)
    dtype, dims = MAN[name]
    T = dtype == "c128" ? ComplexF64 : (dtype == "f64" ? Float64 : Int64)
    a = Array{T}(undef, (isempty(dims) ? (1,) : Tuple(dims))...)
    read!(joinpath(IN, name * ".bin"), a)
    a
end
const OUTMAN = String[]
function save(name, a)
    b = a isa Number ? [a] : collect(a)
    dtype = eltype(b) <: Complex ? "c128" : (eltype(b) <: AbstractFloat ? "f64" : "i64")
    b = dtype == "c128" ? ComplexF64.(b) : (dtype == "f64" ? Float64.(b) : Int64.(b))
    write(joinpath(OUT, name * ".bin"), b)
    push!(OUTMAN, "  \"$name\": {\"dtype\": \"$dtype\", \"shape\": [$(join(size(b), ", "))]}")
end
sites(prefix, n) = Array{ComplexF64}[ComplexF64.(load("$(prefix)$(i-1)")) for i in 1:n]

# ---- H_eff matvec, environment blocks, calculate (N = 6, chi = 12, w = 4; centre 3) ----
psi = GMPS(1, 2, sites("mv_psi", 6), 3)
H = GMPS(2, 2, sites("mv_mpo", 6), 0)
P = ProjMPS(psi, H, psi; rank=2, center=3)
theta = ComplexF64.(load("mv_theta"))
save("mv_L", block(P, 2))
save("mv_R", block(P, 5))
save("mv_out", product(P, theta, false, 2))
save("mv_calculate", calculate(P))

# ---- truncated SVD: singular values under the reference's truncation rule ----
x = ComplexF64.(load("svd_x"))
for (name, kw) in (("full", ()), ("cut", (cutoff=1e-10,)), ("max", (maxdim=9,)), ("min", (cutoff=1e-2, mindim=7)))
    U, S, V = svd(x, 2; kw...)
    save("svd_S_$name", real.(diag(S)))
end

# ---- DMRG energies per sweep (TFIM N = 12, XXZ delta = 0.5 N = 10) ----
# dmrg() prints the energy per sweep; the per-sweep history is recovered by running with maxsweeps = 1, 2, ... on copies
for (name, n) in (("tfim12", 12), ("xxz10", 10))
    M = GMPS(2, 2, sites("dmrg_$(name)_mpo", n), 0)
    energies = Float64[]; bonds = Int64[]
    for nsw in 1:6
        p0 = GMPS(1, 2, sites("dmrg_$(name)_psi", n), 0)
        p, E = dmrg(p0, M; maxdim=24, cutoff=1e-13, minsweeps=nsw, maxsweeps=nsw, verbose=false)
        push!(energies, real(E)); push!(bonds, maxbonddim(p))
    end
    save("dmrg_$(name)_energy", energies)
    save("dmrg_$(name)_maxbond", bonds)
end

write(joinpath(OUT, "manifest.json"), "{\n" * join(OUTMAN, ",\n") * "\n}\n")
println("wrote ", length(OUTMAN), " arrays to ", OUT)
