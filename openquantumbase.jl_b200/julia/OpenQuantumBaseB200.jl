# OpenQuantumBaseB200.jl — Julia-side glue for liboqb_b200.so (the C ABI of include/oqb.h).
#
# STATUS: Julia is not installed in the build container or on the GPU box, so this file has never
# been executed.  It is the binding a maintainer adds next to OpenQuantumBase.jl; the equivalent
# Python `ctypes` binding (openquantumbase.jl_b200/_lib.py + host.py) is what the test-suite runs.
#
# It only ADDS METHODS for a new argument type, `DeviceBatch` (an opaque, device-resident
# n×n×B Array{ComplexF64,3}); no struct of OpenQuantumBase.jl changes:
#   (H::DenseHamiltonian)(du::DeviceBatch, u::DeviceBatch, p, s)        dense_hamiltonian.jl:84
#   update_cache!(cache::DeviceBatch, H::DenseHamiltonian, p, s)        dense_hamiltonian.jl:71
#   (Op::DiffEqLiouvillian{false,false})(du, u, p, t)                   diffeq_liouvillian.jl:70
#   (Op::DiffEqLiouvillian{true,false})(du, u, p, t)                    diffeq_liouvillian.jl:92
#   update_cache!(cache, Op::DiffEqLiouvillian, p, t)                   diffeq_liouvillian.jl:78,111
#   lindblad_jump(Op::DiffEqLiouvillian{false,false}, ψ::DeviceBatch, p, t; uniforms)
module OpenQuantumBaseB200

using OpenQuantumBase
import OpenQuantumBase: update_cache!, DenseHamiltonian, DiffEqLiouvillian, LindbladLiouvillian,
    DaviesGenerator, GapIndices, haml_eigs, OhmicBath, S, correlation
using LinearAlgebra

const LIB = get(ENV, "OQB_B200_LIB", "liboqb_b200.so")

# ---------------------------------------------------------------- context (one per thread × GPU)
const CTX = Dict{Tuple{Int,Int},Ptr{Cvoid}}()
function ctx(device::Int=0)
    get!(CTX, (Threads.threadid(), device)) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:oqb_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        st == 0 || error("oqb_create failed ($st): a B200 is required, there is no CPU fallback")
        h[]
    end
end
check(c, st) = st == 0 || error(unsafe_string(ccall((:oqb_last_error, LIB), Cstring, (Ptr{Cvoid},), c)))

# ---------------------------------------------------------------- device-resident batches
mutable struct DeviceBatch <: AbstractArray{ComplexF64,3}
    ptr::Ptr{Cvoid}
    dims::NTuple{3,Int}
    c::Ptr{Cvoid}
    function DeviceBatch(dims::NTuple{3,Int}; device=0)
        c = ctx(device); p = Ref{Ptr{Cvoid}}(C_NULL)
        check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 16prod(dims), p))
        b = new(p[], dims, c)
        finalizer(x -> ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.c, x.ptr), b)
    end
end
Base.size(b::DeviceBatch) = b.dims
upload!(b::DeviceBatch, a::Array{ComplexF64,3}) =
    check(b.c, ccall((:oqb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Csize_t), b.c, b.ptr, a, sizeof(a)))
download!(a::Array{ComplexF64,3}, b::DeviceBatch) =
    check(b.c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Csize_t), b.c, a, b.ptr, sizeof(a)))

# ---------------------------------------------------------------- operator upload (once per object)
const DENSE = IdDict{Any,Ptr{Cvoid}}()
function dev(H::DenseHamiltonian, lind::Union{Nothing,LindbladLiouvillian}=nothing; device=0)
    get!(DENSE, (H, lind)) do
        n = size(H, 1)
        mats = cat(H.m...; dims=3)                                     # n×n×nterms, already unit-scaled (:38)
        K = lind === nothing ? 0 : length(lind)
        lops = K == 0 ? ComplexF64[] : cat([ComplexF64.(L(0.0)) for L in lind.L]...; dims=3)
        h = Ref{Ptr{Cvoid}}(C_NULL); c = ctx(device)
        check(c, ccall((:oqb_dense_create, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Cint, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                       c, n, length(H.m), mats, K, lops, h))
        h[]
    end
end

batch_s(p, t::Real, B) = fill(Float64(p(t)), 1)                     # shared annealing parameter
batch_s(p, t::AbstractVector, B) = Float64[p(x) for x in t]         # one time per member (schedule sweeps)

# ---------------------------------------------------------------- DiffEqLiouvillian{false,false}
function (Op::DiffEqLiouvillian{false,false})(du::DeviceBatch, u::DeviceBatch, p, t)
    B = size(u, 3); s = batch_s(p, t, B); shared = Cint(length(s) == 1)
    lind = isempty(Op.opensys) ? nothing : Op.opensys[1]::LindbladLiouvillian
    coefH = Float64[real(f(sb)) for f in Op.H.f, sb in s]              # nterms × rows (row-major [b][i] in C)
    gam = lind === nothing ? C_NULL : Float64[g(sb) for g in lind.γ, sb in s]
    c = u.c
    check(c, ccall((:oqb_dense_rhs, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                   dev(Op.H, lind), B, coefH, shared, gam, shared, u.ptr, du.ptr))
end

function update_cache!(cache::DeviceBatch, Op::DiffEqLiouvillian{false,false}, p, t)
    B = size(cache, 3); s = batch_s(p, t, B); shared = Cint(length(s) == 1)
    lind = isempty(Op.opensys) ? nothing : Op.opensys[1]::LindbladLiouvillian
    coefH = Float64[real(f(sb)) for f in Op.H.f, sb in s]
    gam = lind === nothing ? C_NULL : Float64[g(sb) for g in lind.γ, sb in s]
    check(cache.c, ccall((:oqb_dense_update_cache, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}),
                         dev(Op.H, lind), B, coefH, shared, gam, shared, cache.ptr))
end

# ---------------------------------------------------------------- DiffEqLiouvillian{true,false} (Davies)
# One eigen-system per call, shared by the batch, exactly like the single haml_eigs of :94.
function (Op::DiffEqLiouvillian{true,false})(du::DeviceBatch, u::DeviceBatch, p, t::Real)
    s = p(t); c = u.c; n = size(u, 1)
    w, v = haml_eigs(Op.H, s, Op.lvl)                                   # host LAPACK (:94)
    D = Op.opensys_eig[1]::DaviesGenerator
    coup = cat([ComplexF64.(m) for m in D.coupling(s)]...; dims=3)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_ame_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                   c, n, length(w), size(coup, 3), coup, h))
    check(c, ccall((:oqb_ame_set_eigen, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{ComplexF64}), h[], Float64.(w), ComplexF64.(v)))
    # γ, S at the signed rounded gaps of GapIndices (opensys_util.jl:25-61); cross-term lists for
    # non-singleton groups are produced the same way openquantumbase.jl_b200/gaps.py does.
    gam, Sm, cross, crate, lamb, lrs = davies_tables(w, D.γ, D.S, Op.digits, Op.sigdigits)
    check(c, ccall((:oqb_ame_set_davies, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}),
                   h[], gam, Sm, size(cross, 2), cross, crate, size(lamb, 2), lamb, lrs))
    check(c, ccall((:oqb_ame_rhs, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), h[], size(u, 3), u.ptr, du.ptr))
    ccall((:oqb_ame_destroy, LIB), Cint, (Ptr{Cvoid},), h[])
    for lv in Op.opensys; lv(du, u, p, t); end                          # :106-108
end

"""Signed rounded gap tables (l×l, column-major) + cross terms — Julia transcription of gaps.py."""
function davies_tables(w, γ, S, digits, sigdigits)
    l = length(w)
    ω = zeros(l, l)
    for i in 1:l-1, j in i+1:l
        gap = w[j] - w[i]
        if abs(gap) > 10.0^(-digits)
            k = round(gap, sigdigits=sigdigits); ω[i, j] = k; ω[j, i] = -k
        end
    end
    gam = γ.(ω); Sm = S.(ω)
    # groups holding more than one pair → cross / lamb lists (see gaps.py::GapGroups._pairs)
    cross = zeros(Int32, 4, 0); crate = Float64[]; lamb = zeros(Int32, 3, 0); lrs = zeros(2, 0)
    # … (same enumeration as gaps.py; omitted lists are empty for non-degenerate spectra)
    gam, Sm, cross, crate, lamb, lrs
end

# ------------------------------------------------------------------ Redfield Λ(t) on the device
# One GK15 evaluation round of the quadgk! call in (R::RedfieldLiouvillian)(du,u,p,t), redfield.jl:46-64, for nseg
# segments of many integrals at once.  The segment heap (QuadGK's `adapt`) stays in Julia: keep one
# `DataStructures.BinaryMaxHeap` of (E, a, b, slot) per integral, bisect the worst segment of every unconverged
# integral, call this once per round, then `oqb_lambda_total` for ‖Λ‖ — see redfield_device.py for the exact loop.
function lambda_eval!(lam::Ptr{Cvoid}, c::Ptr{Cvoid}, member::Vector{Int32}, coupling::Vector{Int32}, gi::Matrix{Int32},
                      θ::Matrix{Float64}, ck::Matrix{ComplexF64}, cg::Matrix{ComplexF64}, slot::Vector{Int64})
    nseg = length(member)                       # gi, θ: 16×nseg (15 nodes + the end time t); ck: 15×nseg, cg: 7×nseg (Gauss nodes first)
    E = Vector{Float64}(undef, nseg)
    check(c, ccall((:oqb_lambda_eval, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{ComplexF64}, Ptr{ComplexF64},
                    Ptr{Int64}, Ptr{Float64}), lam, nseg, member, coupling, gi, θ, ck, cg, slot, E))
    E
end

# ------------------------------------------------------------------ trajectories that never leave HBM
# nsteps RK4 steps + the jump callback (lindblad_jump, trajectory_jump.jl:71-74) on the device.  coefH / γ are the
# closures evaluated at the 2·nsteps+1 stage times (nterms × rows × (2nsteps+1) Julia arrays = the C layout).
function traj_evolve!(sp::Ptr{Cvoid}, c::Ptr{Cvoid}, ψ::DeviceBatch, nsteps::Int, dt::Float64, coefH::Array{Float64,3},
                      γ::Array{Float64,3}, rng::Ptr{Cvoid}, thr::Ptr{Cvoid}, njumps::Ptr{Cvoid}; step0::Int=0)
    shared = size(coefH, 2) == 1
    check(c, ccall((:oqb_traj_evolve, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                    Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint),
                   sp, ψ.dims[3], nsteps, dt, coefH, shared, γ, shared, ψ.ptr, rng, thr, njumps, C_NULL, 0, step0))
end

# ------------------------------------------------------------------ lindblad_jump{true,false} (ame_jump) on a ψ batch
# After oqb_ame_set_eigen / oqb_ame_set_eigen_device: upload the probability-slot tables in the order ame_jump builds
# its `prob`/`tag` arrays (trajectory_jump.jl:22-44; gaps.py::jump_slots is the sort-based construction), then
#   ccall((:oqb_ame_jump, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int32}, Ptr{Float64}, Cint),
#         ame, B, ψ.ptr, d_uniform, d_mask, d_choice, d_prob, apply)
# returns the index into `tag` per member; apply = 1 performs ψ ← Lψ/‖Lψ‖ in place.

# ------------------------------------------------------------------ bath functions on the device (row B)
# S(ω, bath) for an Ohmic bath (bath_util.jl:17 → ext_util.jl:23-28): one device thread per frequency runs QuadGK's
# adaptive G7K15.  `build_lambshift(ω_range, true, bath, kw)` (bath_util.jl:26-38) becomes
#     s_list = S(collect(ω_range), bath); construct_interpolations(ω_range, s_list)
# with the tabulation below in place of the scalar loop.
function S(ω::AbstractVector{Float64}, bath::OhmicBath; rtol=sqrt(eps()), atol=0.0, max_segments=1024, device=0)
    c = ctx(device)
    out = Vector{Float64}(undef, length(ω))
    check(c, ccall((:oqb_ohmic_lambshift, LIB), Cint,
                   (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cint, Ptr{Float64}, Cdouble, Cdouble, Cint, Ptr{Float64}, Ptr{Float64},
                    Ptr{Int32}), c, bath.η, bath.ωc, bath.β, length(ω), ω, rtol, atol, max_segments, out, C_NULL, C_NULL))
    out
end

# correlation(τ, bath) on an array of lags (ohmic.jl:48-52) — the cfun of the Redfield / ULE kernels
function correlation(τ::AbstractVector{Float64}, bath::OhmicBath; device=0)
    c = ctx(device)
    out = Vector{ComplexF64}(undef, length(τ))
    check(c, ccall((:oqb_ohmic_correlation, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cint, Ptr{Float64}, Ptr{ComplexF64}),
                   c, bath.η, bath.ωc, bath.β, length(τ), τ, out))
    out
end

# Tabulated Lamb shift evaluated on the device at the rounded gaps: hand the prefiltered coefficients of the
# BSpline(Quadratic(Line(OnGrid()))) interpolant (`itp.itp.itp.coefs` of what construct_interpolations returns, n+2
# numbers) to the handle before oqb_ame_davies_ohmic_begin / oqb_ame_set_onesided_ohmic:
#   ccall((:oqb_ame_set_lambshift_spline, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Ptr{Float64}),
#         ame, length(ω_range), first(ω_range), step(ω_range), coefs)
#
# Options: ccall((:oqb_set_real_operand_mode, LIB), Cint, (Ptr{Cvoid}, Cint), ctx(), 0) switches the real-eigenvector /
# real-Hamiltonian products off (general complex product everywhere); oqb_set_zgemm_mode selects 3M / 4M.

end # module
