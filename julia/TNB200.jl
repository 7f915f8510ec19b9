# Thin ccall shim a TensorNetworks.jl maintainer would add (UNVERIFIED: Julia is not available in the
# build image).  It keeps the reference's API and dispatches the hot path into libtnb200.so.
module TNB200
using TensorNetworks
import TensorNetworks: GMPS, AbstractProjMPS, movecenter!, product, project, calculate, buildleft!, buildright!, block, applygates!, norm, normalize!, inner

const lib = get(ENV, "TNB200_LIB", "libtnb200.so")
struct TruncT; cutoff::Cdouble; maxdim::Int64; mindim::Int64; end
struct LanczosT; krylovdim::Int32; maxiter::Int32; tol::Cdouble; end
check(st) = st == 0 || error(unsafe_string(ccall((:tn_last_error, lib), Cstring, ())))

mutable struct Ctx; h::Ptr{Cvoid}; end
function Ctx(device::Integer=0)
    r = Ref{Ptr{Cvoid}}(); check(ccall((:tn_ctx_create, lib), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r))
    c = Ctx(r[]); finalizer(x -> ccall((:tn_ctx_destroy, lib), Int32, (Ptr{Cvoid},), x.h), c); c
end

# device mirror of a GMPS (structures/mps/gmps.jl:8-13)
mutable struct CuGMPS; h::Ptr{Cvoid}; rank::Int; dim::Int; N::Int; ctx::Ctx; end
function CuGMPS(ctx::Ctx, psi::GMPS)
    N = length(psi); r = psi.rank
    dims = Int64[size(psi[i])[k] for i in 1:N for k in 1:r+2]          # N x (rank+2), row-major
    ptrs = [pointer(psi.tensors[i]) for i in 1:N]
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve psi check(ccall((:tn_mps_upload, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Int64}, Ptr{Ptr{ComplexF64}}, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, r, psi.dim, N, dims, ptrs, psi.center, h))
    m = CuGMPS(h[], r, psi.dim, N, ctx); finalizer(x -> ccall((:tn_mps_free, lib), Int32, (Ptr{Cvoid},), x.h), m); m
end
function download!(psi::GMPS, m::CuGMPS)
    dims = Matrix{Int64}(undef, m.rank + 2, m.N)
    check(ccall((:tn_mps_dims, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}), m.h, dims))
    for i in 1:m.N
        t = Array{ComplexF64}(undef, dims[:, i]...)
        check(ccall((:tn_mps_download_site, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}), m.h, i, t))
        psi.tensors[i] = t
    end
    c = Ref{Int32}(); check(ccall((:tn_mps_info, lib), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ref{Int32}), m.h, C_NULL, C_NULL, C_NULL, c))
    psi.center = c[]; psi
end

# device-backed projector: slots into dmrg(psi, Hs::ProjMPSSum) through the AbstractProjMPS protocol
mutable struct CuProjMPS <: AbstractProjMPS; h::Ptr{Cvoid}; psi::CuGMPS; H::CuGMPS; center::Int; end
function CuProjMPS(psi::CuGMPS, H::CuGMPS; coeff=1.0, center=1)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:tn_env_create, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, ComplexF64, Int32, Ref{Ptr{Cvoid}}),
                psi.ctx.h, psi.h, H.h, psi.h, ComplexF64(coeff), center, h))
    CuProjMPS(h[], psi, H, center)
end
movecenter!(p::CuProjMPS, idx::Int) = (check(ccall((:tn_env_movecenter, lib), Int32, (Ptr{Cvoid}, Int32), p.h, idx)); p.center = idx)
buildleft!(p::CuProjMPS, idx::Int) = check(ccall((:tn_env_buildleft, lib), Int32, (Ptr{Cvoid}, Int32), p.h, idx))
buildright!(p::CuProjMPS, idx::Int) = check(ccall((:tn_env_buildright, lib), Int32, (Ptr{Cvoid}, Int32), p.h, idx))
function product(p::CuProjMPS, A, direction::Bool=false, nsites::Int=2)   # projmps.jl:103-145
    out = similar(A)
    check(ccall((:tn_env_product, lib), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Ptr{ComplexF64}), p.h, A, direction, out)); out
end
function calculate(p::CuProjMPS)
    v = Ref{ComplexF64}(); check(ccall((:tn_env_calculate, lib), Int32, (Ptr{Cvoid}, Ref{ComplexF64}), p.h, v)); v[]
end

# squared MPS projection ProjMPS(V, psi; rank=2, squared=true, coeff) (dmrg.jl:144-145) and sums of projections (projmpssum.jl)
function CuProjMPS(V::CuGMPS, psi::CuGMPS, ::Val{:squared}; coeff=1.0, center=1)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:tn_env_create_squared, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, ComplexF64, Int32, Ref{Ptr{Cvoid}}),
                psi.ctx.h, V.h, psi.h, ComplexF64(coeff), center, h))
    CuProjMPS(h[], psi, V, center)
end
function project(p::CuProjMPS, A, direction::Bool=false, nsites::Int=2)   # projmps.jl:153-185 (A only fixes the shape)
    out = similar(A)
    check(ccall((:tn_env_project, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{ComplexF64}), p.h, direction, nsites, out)); out
end
mutable struct CuProjMPSSum <: AbstractProjMPS; h::Ptr{Cvoid}; projs::Vector{CuProjMPS}; center::Int; end
function CuProjMPSSum(projs::Vector{CuProjMPS}; center=1)
    h = Ref{Ptr{Cvoid}}(); hs = [p.h for p in projs]
    check(ccall((:tn_envsum_create, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Int32, Ref{Ptr{Cvoid}}), projs[1].psi.ctx.h, length(hs), hs, center, h))
    s = CuProjMPSSum(h[], projs, center); finalizer(x -> ccall((:tn_envsum_free, lib), Int32, (Ptr{Cvoid},), x.h), s); s
end
movecenter!(p::CuProjMPSSum, idx::Int) = (check(ccall((:tn_envsum_movecenter, lib), Int32, (Ptr{Cvoid}, Int32), p.h, idx)); p.center = idx)
function product(p::CuProjMPSSum, A, direction::Bool=false, nsites::Int=2)   # projmpssum.jl:63-73
    out = similar(A)
    check(ccall((:tn_envsum_product, lib), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int32, Int32, Ptr{ComplexF64}), p.h, A, direction, nsites, out)); out
end
function project(p::CuProjMPSSum, A, direction::Bool=false, nsites::Int=2)   # projmpssum.jl:81-91
    out = similar(A)
    check(ccall((:tn_envsum_project, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{ComplexF64}), p.h, direction, nsites, out)); out
end
function calculate(p::CuProjMPSSum)
    v = Ref{ComplexF64}(); check(ccall((:tn_envsum_calculate, lib), Int32, (Ptr{Cvoid}, Ref{ComplexF64}), p.h, v)); v[]
end
# fused sweeps over a sum: dmrg.jl:35-63 for dmrg(psi, H1, H2, V...; coeffs) and nsites = 1 | 2; vmps.jl:36-62
function dmrg_sweep!(psi::CuGMPS, Hs::CuProjMPSSum, direction::Bool; nsites=2, krylovdim=3, kryloviter=2, cutoff=1e-12, maxdim=1000, mindim=1)
    e = Ref{Cdouble}(); D = Ref{Int64}()
    check(ccall((:tn_dmrg_sweep_sum, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, LanczosT, TruncT, Ref{Cdouble}, Ref{Int64}),
                psi.h, Hs.h, direction, nsites, LanczosT(krylovdim, kryloviter, 1e-14), TruncT(cutoff, maxdim, mindim), e, D))
    e[], D[]
end
function vmps_sweep!(psi::CuGMPS, Vs::CuProjMPSSum, direction::Bool; nsites=2, cutoff=1e-12, maxdim=1000, mindim=1)
    D = Ref{Int64}()
    check(ccall((:tn_vmps_sweep, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, TruncT, Ref{Int64}),
                psi.h, Vs.h, direction, nsites, TruncT(cutoff, maxdim, mindim), D))
    D[]
end

# fused sweep: replaces the body of dmrg.jl:35-63
function dmrg_sweep!(psi::CuGMPS, Hs::CuProjMPS, direction::Bool; krylovdim=3, kryloviter=2, cutoff=1e-12, maxdim=1000, mindim=1)
    e = Ref{Cdouble}(); D = Ref{Int64}()
    check(ccall((:tn_dmrg_sweep, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, LanczosT, TruncT, Ref{Cdouble}, Ref{Int64}),
                psi.h, Hs.h, direction, LanczosT(krylovdim, kryloviter, 1e-14), TruncT(cutoff, maxdim, mindim), e, D))
    e[], D[]
end

# ---- TEBD / QJMC side: gatelist.jl:191-227, gmps.jl:29-51, qjmc.jl:59-164 -------------------------------------------
mutable struct CuGateList; h::Ptr{Cvoid}; ctx::Ctx; end
function CuGateList(ctx::Ctx, dim::Int, gates::GateList)          # after trotterize(): rows of (site, gate tensor)
    counts = Int32[length(r) for r in gates.sites]
    sites = Int32[s for r in gates.sites for s in r]
    tens = [ComplexF64.(g) for r in gates.gates for g in r]
    nsites = Int32[ndims(g) ÷ 2 for g in tens]
    ptrs = [pointer(g) for g in tens]
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve tens check(ccall((:tn_gates_upload, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{ComplexF64}}, Ref{Ptr{Cvoid}}),
        ctx.h, dim, length(counts), counts, sites, nsites, ptrs, h))
    g = CuGateList(h[], ctx); finalizer(x -> ccall((:tn_gates_free, lib), Int32, (Ptr{Cvoid},), x.h), g); g
end
applygates!(psi::CuGMPS, gates::CuGateList; cutoff=0.0, maxdim=0, mindim=1) =
    check(ccall((:tn_apply_gates, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, TruncT), psi.h, gates.h, TruncT(cutoff, maxdim, mindim)))
function norm(psi::CuGMPS)
    v = Ref{ComplexF64}(); check(ccall((:tn_mps_norm, lib), Int32, (Ptr{Cvoid}, Ref{ComplexF64}), psi.h, v)); v[]
end
normalize!(psi::CuGMPS) = check(ccall((:tn_mps_normalize, lib), Int32, (Ptr{Cvoid},), psi.h))
movecenter!(psi::CuGMPS, idx::Int; cutoff=0.0, maxdim=0, mindim=1) =
    check(ccall((:tn_mps_movecenter, lib), Int32, (Ptr{Cvoid}, Int32, TruncT), psi.h, idx, TruncT(cutoff, maxdim, mindim)))

# One trajectory (qjmc_simulation's loop) on the device; `uniforms` = 3 per step from Julia's RNG, or nothing for the
# library's counter-based generator keyed by (seed, trajectory, step).
function qjmc_run!(psi::CuGMPS, gates::CuGateList, jumpsites::Vector{Int32}, jumpops::Array{ComplexF64,3}, coeffs::Vector{Float64},
                   steps::Int, dt::Float64; cutoff=1e-12, maxdim=0, mindim=1, uniforms=nothing, seed=0, trajectory=0,
                   obsop=nothing, save_every=1, classical=true)
    nsave = obsop === nothing ? 0 : steps ÷ save_every
    obs = zeros(ComplexF64, psi.N, max(nsave, 1)); jumps = zeros(Int32, steps + 1); times = zeros(Float64, steps + 1); nj = Ref{Int32}()
    check(ccall((:tn_qjmc_run, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{ComplexF64}, Ptr{Float64}, Int32, Float64, TruncT, Ptr{Float64}, UInt64, UInt64,
         Ptr{ComplexF64}, Int32, Ptr{ComplexF64}, Ptr{Int32}, Ptr{Float64}, Int32, Ref{Int32}, Int32),
        psi.h, gates.h, length(jumpsites), jumpsites, jumpops, coeffs, steps, dt, TruncT(cutoff, maxdim, mindim),
        uniforms === nothing ? C_NULL : pointer(uniforms), seed, trajectory, obsop === nothing ? C_NULL : pointer(obsop), save_every,
        obs, jumps, times, steps + 1, nj, classical))
    jumps[1:nj[]], times[1:nj[]], obs[:, 1:nsave]
end
# inner(st, psi, oplist, phi) (mps.jl:87-134) with psi, phi on the device: per-term coeff * <psi| O_t |phi>
function inner(st::Sitetypes, psi::CuGMPS, oplist::OpList, phi::CuGMPS)
    nops = Int32[length(s) for s in oplist.sites]
    sites = Int32[x for s in oplist.sites for x in s]
    mats = ComplexF64[]
    for ops in oplist.ops, name in ops; append!(mats, vec(ComplexF64.(op(st, name)))); end
    coeffs = ComplexF64.(oplist.coeffs); out = zeros(ComplexF64, length(nops))
    check(ccall((:tn_inner_oplist, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                psi.h, phi.h, length(nops), nops, sites, mats, coeffs, out))
    out
end

# MPO(st, H): assemble on the host as the reference does (mpo.jl:323-440), compress the bonds on the device (mpo.jl:443-457)
compress!(O::CuGMPS; cutoff=1e-15, maxdim=0, mindim=1) =
    check(ccall((:tn_mpo_compress, lib), Int32, (Ptr{Cvoid}, TruncT), O.h, TruncT(cutoff, maxdim, mindim)))
# applygates(psi, gates; error=true): product of the two-site gates' truncation fidelities (gatelist.jl:160-168,191-223)
function applygates(psi::CuGMPS, gates::CuGateList; cutoff=0.0, maxdim=0, mindim=1)
    f = Ref{Cdouble}()
    check(ccall((:tn_apply_gates_fidelity, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, TruncT, Ref{Cdouble}), psi.h, gates.h, TruncT(cutoff, maxdim, mindim), f))
    f[]
end
# iTEBD, two-site cell (itebd.jl:71-119): upload the iGMPS once, apply the cell gate nsteps times, read singular values / tensors back
mutable struct CuiGMPS; h::Ptr{Cvoid}; dim::Int; ctx::Ctx; end
function CuiGMPS(ctx::Ctx, psi::iGMPS)
    L = length(psi.tensors)
    dims = Int64[size(psi.tensors[i])[k] for i in 1:L for k in 1:3]
    tens = [ComplexF64.(t) for t in psi.tensors]; sing = [Float64.(s) for s in psi.singulars]
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve tens sing check(ccall((:tn_imps_create, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Ptr{ComplexF64}}, Ptr{Ptr{Float64}}, Ref{Ptr{Cvoid}}),
        ctx.h, psi.dim, L, dims, [pointer(t) for t in tens], [pointer(s) for s in sing], h))
    m = CuiGMPS(h[], psi.dim, ctx); finalizer(x -> ccall((:tn_imps_free, lib), Int32, (Ptr{Cvoid},), x.h), m); m
end
itebd_apply_gates!(psi::CuiGMPS, gate::Array{ComplexF64,4}, nsteps::Int; cutoff=1e-12, maxdim=0, mindim=1) =
    check(ccall((:tn_itebd_apply_gate, lib), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Int32, TruncT), psi.h, gate, nsteps, TruncT(cutoff, maxdim, mindim)))

# Many trajectories from one initial state (the loop a user writes around qjmc_simulation): tn_qjmc_ensemble hands them out to
# `workers` host threads / CUDA streams inside the library; trajectory ids key the random numbers, so results do not depend on
# the worker count or on which GPU ran them (shard ids over processes / GPUs as `ids = rank+1:world:ntraj`).
# See include/tn_c_api.h for the argument list; the call mirrors qjmc_run! with host tensors instead of device handles.
end # module
