# OpenQuantumBaseB200.jl — Julia-side glue for liboqb_b200.so (the C ABI of include/oqb.h).
#
# STATUS: Julia is not installed in the build container or on the GPU box, so this file has NEVER BEEN EXECUTED.
# Everything it calls is exercised through the same C entry points by the plain-C programs examples/c_abi_demo.c,
# examples/c_liouv_demo.c, examples/c_comm_demo.c (compiled and run by the test-suite) and by the ctypes host mirror.
# The file is deliberately thin: all orchestration the reference does inside an RHS call — H(s) assembly, haml_eigs,
# GapIndices, cross-term lists, γ/S tables, plan caching, grouping a batch by s, the NCCL reduce — lives in the library
# (csrc/liouvillian.cu, csrc/ame_terms.cu, csrc/comm.cu); Julia only evaluates the user's closures and passes pointers.
#
# It only ADDS METHODS for new argument types (no struct of OpenQuantumBase.jl changes):
#   DeviceBatch  — an opaque device-resident n×n×B (ρ batch) or n×1×B (ψ ensemble) Array{ComplexF64,3}
#   (H::DenseHamiltonian)(du::DeviceBatch, u::DeviceBatch, p, s)              dense_hamiltonian.jl:84
#   (H::SparseHamiltonian)(du::DeviceBatch, u::DeviceBatch, p, s)             sparse_hamiltonian.jl:83
#   update_cache!(cache::DeviceBatch, H, p, s), update_vectorized_cache!       dense_hamiltonian.jl:71,78
#   (Op::DiffEqLiouvillian{false,false})(du, u, p, t), update_cache!           diffeq_liouvillian.jl:70,78
#   (Op::DiffEqLiouvillian{true,false})(du, u, p, t), update_cache!            diffeq_liouvillian.jl:92,111
#   (R::RedfieldLiouvillian)(du, u, p, t) with Λ supplied on the device        redfield.jl:65-68
#   traj_matvec!(dψ, Op, ψ, p, t), lindblad_jump(Op{false,false}, ψ, p, t; uniforms)   trajectory_jump.jl:71-74
#   lindblad_jump(Op{false,false}, ψ, p, t, uniforms::Matrix) with several Liouvillians + resample   trajectory_jump.jl:71-88
#   the same methods for ConstantDenseHamiltonian / ConstantSparseHamiltonian (one term, coefficient 1)   dense_hamiltonian.jl:134-168
#   redfield_vectorized!(cache, S, Λ)                                          redfield.jl:97-102
#   lindblad_jump(Op{true,false}, ψ, p, t; uniforms) — ame_jump for a DaviesGenerator                 trajectory_jump.jl:14-51,64-69
#   norm2(Op, ψ), expect(obs, state), free!(H)
#   ensemble_expect_allreduce!(comm, out, obs, ψ), allgather_obs!               the ensemble-average reduce / per-schedule gather (SURVEY §8(e))
# tests/test_julia_glue_static.py checks every ccall below against include/oqb.h (symbol exported, argument count, argument classes).
module OpenQuantumBaseB200

using OpenQuantumBase
import OpenQuantumBase: update_cache!, update_vectorized_cache!, lindblad_jump, DenseHamiltonian, SparseHamiltonian,
    DiffEqLiouvillian, LindbladLiouvillian, DaviesGenerator, OneSidedAMELiouvillian, RedfieldLiouvillian,
    ConstantCouplings, OhmicBath
using LinearAlgebra, SparseArrays

const LIB = "liboqb_b200.so"   # resolved through the loader path (LD_LIBRARY_PATH / rpath); no environment switches

# ---------------------------------------------------------------- context: one per (thread, GPU)
const CTX = Dict{Tuple{Int,Int},Ptr{Cvoid}}()
function ctx(device::Int=0)
    get!(CTX, (Threads.threadid(), device)) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:oqb_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        st == 0 || error("oqb_create failed ($st): a B200 is required, there is no CPU fallback")
        h[]
    end
end
lasterr(c) = unsafe_string(ccall((:oqb_last_error, LIB), Cstring, (Ptr{Cvoid},), c))
check(c, st) = st == 0 ? nothing : error(lasterr(c))
sync(c=ctx()) = check(c, ccall((:oqb_sync, LIB), Cint, (Ptr{Cvoid},), c))
set_option(name::String, v::Integer; c=ctx()) =
    check(c, ccall((:oqb_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), c, name, v))

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
# element access goes through the host (display, tests); bulk transfers use upload!/download!
function Base.getindex(b::DeviceBatch, i::Int, j::Int, k::Int)
    v = Ref{ComplexF64}()
    off = 16 * ((i - 1) + b.dims[1] * ((j - 1) + b.dims[2] * (k - 1)))
    check(b.c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ref{ComplexF64}, Ptr{Cvoid}, Csize_t), b.c, v, b.ptr + off, 16))
    v[]
end
Base.IndexStyle(::Type{DeviceBatch}) = IndexCartesian()
upload!(b::DeviceBatch, a::Array{ComplexF64,3}) =
    check(b.c, ccall((:oqb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Csize_t), b.c, b.ptr, a, sizeof(a)))
download!(a::Array{ComplexF64,3}, b::DeviceBatch) =
    check(b.c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Csize_t), b.c, a, b.ptr, sizeof(a)))
DeviceBatch(a::Array{ComplexF64,3}; device=0) = (b = DeviceBatch(size(a); device=device); upload!(b, a); b)

# closure values per batch member: coefficient rows are row-major [member][term] in C = column-major term×member here
batch_s(p, t::Real) = Float64[p(t)]                                  # one annealing parameter for the whole batch
batch_s(p, t::AbstractVector) = Float64[p(x) for x in t]             # one time per member (schedule sweeps)
coef_rows(fs, s) = Float64[real(f(sb)) for f in fs, sb in s]
shared(s) = Cint(length(s) == 1)

# ---------------------------------------------------------------- operators: uploaded once per object, immutable afterwards
const HANDLES = IdDict{Any,Ptr{Cvoid}}()

# Constant Hamiltonians (dense_hamiltonian.jl:134-168, sparse_hamiltonian.jl:166-203) are the one-term case with coefficient 1:
# the same device operator and entry points (a single Pauli-structured matrix takes the stencil pass inside oqb_dense_rhs).
const AnyDense = Union{DenseHamiltonian,OpenQuantumBase.ConstantDenseHamiltonian}
const AnySparse = Union{SparseHamiltonian,OpenQuantumBase.ConstantSparseHamiltonian}
const ONE = (s -> 1.0,)
hf(H::Union{DenseHamiltonian,SparseHamiltonian}) = H.f
hf(::Union{OpenQuantumBase.ConstantDenseHamiltonian,OpenQuantumBase.ConstantSparseHamiltonian}) = ONE
hm(H::Union{DenseHamiltonian,SparseHamiltonian}) = H.m
hm(H::Union{OpenQuantumBase.ConstantDenseHamiltonian,OpenQuantumBase.ConstantSparseHamiltonian}) = [H.u_cache]   # already unit-scaled (:141-142)

function dev(H::AnyDense, lind::Union{Nothing,LindbladLiouvillian}=nothing; c=ctx())
    get!(HANDLES, (H, lind)) do
        n = size(H, 1)
        mats = cat([ComplexF64.(m) for m in hm(H)]...; dims=3)          # already unit-scaled (dense_hamiltonian.jl:38)
        K = lind === nothing ? 0 : length(lind.L)
        lops = K == 0 ? ComplexF64[] : cat([ComplexF64.(Matrix(L(0.0))) for L in lind.L]...; dims=3)  # constant L_k
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(c, ccall((:oqb_dense_create, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Cint, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                       c, n, length(hm(H)), mats, K, lops, h))
        h[]
    end
end

"concatenated CSC arrays of a list of SparseMatrixCSC{ComplexF64,Int} exactly as Julia stores them (1-based)"
function pack(ms)
    colptr = vcat([Int64.(m.colptr) for m in ms]...)
    rowval = vcat([Int64.(m.rowval) for m in ms]...)
    nzval = vcat([ComplexF64.(m.nzval) for m in ms]...)
    off = Int64[0; cumsum([nnz(m) for m in ms])]
    isempty(rowval) && (rowval = Int64[1]; nzval = ComplexF64[0])
    colptr, rowval, nzval, off
end

function dev(H::AnySparse, lind::Union{Nothing,LindbladLiouvillian}=nothing; c=ctx())
    get!(HANDLES, (H, lind)) do
        n = size(H, 1)
        cp, rv, nz, off = pack(hm(H))
        K = lind === nothing ? 0 : length(lind.L)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        if K == 0
            check(c, ccall((:oqb_sparse_create, LIB), Cint,
                           (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Int64},
                            Ptr{ComplexF64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                           c, n, length(hm(H)), cp, rv, nz, off, 0, C_NULL, C_NULL, C_NULL, C_NULL, h))
        else
            lcp, lrv, lnz, loff = pack([sparse(ComplexF64.(L(0.0))) for L in lind.L])
            check(c, ccall((:oqb_sparse_create, LIB), Cint,
                           (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Ptr{Int64}, Cint, Ptr{Int64}, Ptr{Int64},
                            Ptr{ComplexF64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                           c, n, length(hm(H)), cp, rv, nz, off, K, lcp, lrv, lnz, loff, h))
        end
        h[]
    end
end

# ---------------------------------------------------------------- Hamiltonians
function (H::AnyDense)(du::DeviceBatch, u::DeviceBatch, p, s)                  # du = −i[H(s),u]  (OVERWRITES)
    sv = s isa Real ? Float64[s] : Float64.(s)
    check(u.c, ccall((:oqb_dense_rhs, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                     dev(H), size(u, 3), coef_rows(hf(H), sv), shared(sv), C_NULL, 1, u.ptr, du.ptr))
end
function update_cache!(cache::DeviceBatch, H::AnyDense, p, s)                   # cache = −iH(s)
    sv = s isa Real ? Float64[s] : Float64.(s)
    check(cache.c, ccall((:oqb_dense_update_cache, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}),
                         dev(H), size(cache, 3), coef_rows(hf(H), sv), shared(sv), C_NULL, 1, cache.ptr))
end
function update_vectorized_cache!(cache::DeviceBatch, H::AnyDense, p, s)        # cache = i(Hᵀ⊗I − I⊗H)
    sv = s isa Real ? Float64[s] : Float64.(s)
    check(cache.c, ccall((:oqb_dense_update_vectorized_cache, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}),
                         dev(H), size(cache, 3), coef_rows(hf(H), sv), shared(sv), cache.ptr))
end
function (H::AnySparse)(du::DeviceBatch, u::DeviceBatch, p, s)
    sv = s isa Real ? Float64[s] : Float64.(s)
    check(u.c, ccall((:oqb_sparse_commutator, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                     dev(H), size(u, 3), coef_rows(hf(H), sv), shared(sv), u.ptr, du.ptr))
end
"update_cache!(cache, H::SparseHamiltonian, p, s): the values on the union pattern, as a SparseMatrixCSC (one member)"
function update_cache!(cache::SparseMatrixCSC{ComplexF64,Int}, H::AnySparse, p, s::Real; c=ctx())
    h = dev(H)
    nnzU = Ref{Int64}(0)
    check(c, ccall((:oqb_sparse_pattern_nnz, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), h, nnzU))
    d = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 16nnzU[], d))
    check(c, ccall((:oqb_sparse_update_cache, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}),
                   h, 1, coef_rows(hf(H), Float64[s]), 1, C_NULL, 1, d[]))
    length(cache.nzval) == nnzU[] || error("cache must have the union sparsity pattern of the terms (sparse_hamiltonian.jl:33-34)")
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Csize_t), c, cache.nzval, d[], 16nnzU[]))
    ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, d[])
    cache
end

function update_vectorized_cache!(cache::DeviceBatch, H::AnySparse, p, s)                 # cache = i(Hᵀ⊗I − I⊗H), dense n²×n² (small n)
    sv = s isa Real ? Float64[s] : Float64.(s)
    check(cache.c, ccall((:oqb_sparse_update_vectorized_cache, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}),
                         dev(H), size(cache, 3), coef_rows(hf(H), sv), shared(sv), cache.ptr))
end
"release the device operator of a Hamiltonian (and Lindblad set) explicitly; otherwise it lives as long as the process"
function free!(H::Union{AnyDense,AnySparse}, lind::Union{Nothing,LindbladLiouvillian}=nothing)
    h = pop!(HANDLES, (H, lind), C_NULL)
    h == C_NULL && return
    H isa AnyDense ? ccall((:oqb_dense_destroy, LIB), Cint, (Ptr{Cvoid},), h) : ccall((:oqb_sparse_destroy, LIB), Cint, (Ptr{Cvoid},), h)
    nothing
end

# ---------------------------------------------------------------- DiffEqLiouvillian{false,false}: H + Lindblad (+ others)
lindblad_of(Op) = (i = findfirst(x -> x isa LindbladLiouvillian, Op.opensys); i === nothing ? nothing : Op.opensys[i])

function (Op::DiffEqLiouvillian{false,false})(du::DeviceBatch, u::DeviceBatch, p, t)
    s = batch_s(p, t)
    lind = Op.H isa DenseHamiltonian ? lindblad_of(Op) : nothing
    if Op.H isa DenseHamiltonian
        gam = lind === nothing ? C_NULL : Float64[g(sb) for g in lind.γ, sb in s]
        check(u.c, ccall((:oqb_dense_rhs, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                         dev(Op.H, lind), size(u, 3), coef_rows(hf(Op.H), s), shared(s), gam, shared(s), u.ptr, du.ptr))
    else
        Op.H(du, u, p, s)                                                # diffeq_liouvillian.jl:72 (s to H …)
    end
    for lv in Op.opensys                                                  # :73-75 (… t to the Liouvillians), each ACCUMULATES
        lv === lind && continue
        lv(du, u, p, t)
    end
    du
end

function update_cache!(cache::DeviceBatch, Op::DiffEqLiouvillian{false,false}, p, t)
    s = batch_s(p, t)
    lind = lindblad_of(Op)
    gam = lind === nothing ? C_NULL : Float64[g(sb) for g in lind.γ, sb in s]
    check(cache.c, ccall((:oqb_dense_update_cache, LIB), Cint,
                         (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}),
                         dev(Op.H, lind), size(cache, 3), coef_rows(hf(Op.H), s), shared(s), gam, shared(s), cache.ptr))
end

"a LindbladLiouvillian on its own ACCUMULATES (lindblad.jl:94-101)"
function (lind::LindbladLiouvillian)(du::DeviceBatch, u::DeviceBatch, p, t)
    s = batch_s(p, t)
    h = get!(HANDLES, lind) do                                            # zero Hamiltonian, the L_k of this Liouvillian
        n = lind.size[1]; z = zeros(ComplexF64, n, n, 1)
        lops = cat([ComplexF64.(Matrix(L(0.0))) for L in lind.L]...; dims=3)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(u.c, ccall((:oqb_dense_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Cint, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                         u.c, n, 1, z, length(lind.L), lops, r))
        r[]
    end
    gam = Float64[g(sb) for g in lind.γ, sb in s]
    check(u.c, ccall((:oqb_lindblad_acc, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                     h, size(u, 3), gam, shared(s), u.ptr, du.ptr))
end

# ---------------------------------------------------------------- DiffEqLiouvillian{true,false}: everything inside the library
# γ(ω) / S(ω) closures reach the library as C callbacks (oqb_spectrum_fn), evaluated at the distinct gap keys only.
function spectrum_thunk(user::Ptr{Cvoid}, n::Int64, w::Ptr{Float64}, out::Ptr{Float64})::Cvoid
    f = unsafe_pointer_to_objref(user)::Base.RefValue{Any}
    for i in 1:n
        unsafe_store!(out, Float64(f[](unsafe_load(w, i))), i)
    end
    nothing
end
const KEEP = IdDict{Any,Any}()   # closures handed to C must stay rooted as long as the handle lives

function liouv(Op::DiffEqLiouvillian{true,false}; c=ctx())
    get!(HANDLES, Op) do
        Op.H isa DenseHamiltonian || error("the one-call path takes a DenseHamiltonian; sparse H: rotate on the host and use oqb_ame_* directly")
        all(lv -> lv.coupling isa ConstantCouplings, Op.opensys_eig) || error("constant couplings only")
        mats = vcat([[ComplexF64.(Matrix(m)) for m in lv.coupling.mats] for lv in Op.opensys_eig]...)
        coup = cat(mats...; dims=3)
        L = Ref{Ptr{Cvoid}}(C_NULL)
        check(c, ccall((:oqb_liouv_create, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{ComplexF64}, Ref{Ptr{Cvoid}}),
                       c, dev(Op.H), Op.lvl, Op.digits, Op.sigdigits, size(coup, 3), coup, L))
        fn = @cfunction(spectrum_thunk, Cvoid, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}))
        roots = Any[]
        k0 = 0
        for lv in Op.opensys_eig
            nk = length(lv.coupling.mats)
            g = Ref{Any}(lv.γ); S = Ref{Any}(lv.S); push!(roots, g, S)
            gp, Sp = pointer_from_objref(g), pointer_from_objref(S)
            if lv isa DaviesGenerator
                if lv.γ isa OhmicBath   # hypothetical: γ passed as the bath object ⇒ closed form on the device, S ≡ 0
                    b = lv.γ
                    check(c, ccall((:oqb_liouv_add_davies_ohmic, LIB), Cint,
                                   (Ptr{Cvoid}, Cint, Cint, Float64, Float64, Float64, Cint, Float64, Float64, Ptr{Float64}),
                                   L[], k0, nk, b.η, b.ωc, b.β, 0, 0.0, 1.0, C_NULL))
                else
                    check(c, ccall((:oqb_liouv_add_davies_fn, LIB), Cint,
                                   (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), L[], k0, nk, fn, gp, fn, Sp))
                end
            elseif lv isa OneSidedAMELiouvillian
                al = Int32[k0 + a - 1 for (a, _) in lv.inds]; be = Int32[k0 + b - 1 for (_, b) in lv.inds]
                check(c, ccall((:oqb_liouv_add_onesided_fn, LIB), Cint,
                               (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                               L[], length(al), al, be, fn, gp, fn, Sp))
            else
                error("eigenbasis term $(typeof(lv)) is not on the device path")
            end
            k0 += nk
        end
        KEEP[Op] = roots
        L[]
    end
end

function (Op::DiffEqLiouvillian{true,false})(du::DeviceBatch, u::DeviceBatch, p, t)
    s = batch_s(p, t)
    check(u.c, ccall((:oqb_liouv_rhs, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                     liouv(Op), size(u, 3), coef_rows(hf(Op.H), s), shared(s), u.ptr, du.ptr))   # :92-105 (OVERWRITES)
    for lv in Op.opensys; lv(du, u, p, t); end                                                 # :106-108
    du
end
"same with HOST arrays (n×n×B): H2D/D2H inside the call, pipelined against the kernels"
function (Op::DiffEqLiouvillian{true,false})(du::Array{ComplexF64,3}, u::Array{ComplexF64,3}, p, t)
    isempty(Op.opensys) || error("host-buffer path: eigenbasis terms only")
    s = batch_s(p, t)
    c = ctx()
    check(c, ccall((:oqb_liouv_rhs_host, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}),
                   liouv(Op), size(u, 3), coef_rows(hf(Op.H), s), shared(s), u, du))
    du
end
function update_cache!(cache::DeviceBatch, Op::DiffEqLiouvillian{true,false}, p, t::Real)
    check(cache.c, ccall((:oqb_liouv_update_cache, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                         liouv(Op), coef_rows(hf(Op.H), Float64[p(t)]), cache.ptr))              # :111-125 (OVERWRITES)
    for lv in Op.opensys; update_cache!(cache, lv, p, t); end
    cache
end

"GapIndices of H(s) from the device grouping, in the reference's own form (uniq_w, uniq_a, uniq_b, a0, b0)"
function device_gap_indices(Op::DiffEqLiouvillian{true,false}, p, t::Real)
    c = ctx(); plan = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_liouv_prepare, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Ptr{Cvoid}}), liouv(Op), coef_rows(hf(Op.H), Float64[p(t)]), plan))
    nk = Ref{Int64}(0); nz = Ref{Int64}(0)
    check(c, ccall((:oqb_ame_gaps_build, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Int64}, Ref{Int64}), plan[], Op.digits, Op.sigdigits, nk, nz))
    l = Op.lvl; np = l * (l - 1) ÷ 2 - nz[]
    keys = Vector{Float64}(undef, nk[]); gptr = Vector{Int64}(undef, nk[] + 1)
    a = Vector{Int32}(undef, np); b = similar(a); a0 = Vector{Int32}(undef, 2nz[] + l); b0 = similar(a0)
    check(c, ccall((:oqb_ame_gaps_keys, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), plan[], keys))
    check(c, ccall((:oqb_ame_gaps_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), plan[], gptr, a, b, a0, b0))
    uniq_a = [Int.(a[gptr[g]+1:gptr[g+1]]) for g in 1:nk[]]; uniq_b = [Int.(b[gptr[g]+1:gptr[g+1]]) for g in 1:nk[]]
    keys, uniq_a, uniq_b, Int.(a0), Int.(b0)
end

"lindblad_jump(Op::DiffEqLiouvillian{true,false}, ψ, p, t) on a batch — ame_jump for ONE DaviesGenerator (trajectory_jump.jl:14-51, :64-69).
The probability slots are listed in the reference's order: for every unique positive gap (ascending) and coupling i, L₊ = sparse(a, b, σ_i[a,b])
with rate γ(ω), then L₋ = sparse(b, a, σ_i[b,a]) with rate γ(−ω); then, per coupling, the zero-gap group with rate γ(0).  Returns the chosen
slot per member (0-based); apply=true performs ψ ← v L v'ψ / ‖·‖."
function lindblad_jump(Op::DiffEqLiouvillian{true,false}, ψ::DeviceBatch, p, t::Real; uniforms::Vector{Float64}, apply=false)
    length(Op.opensys_eig) == 1 && Op.opensys_eig[1] isa DaviesGenerator || error("ame_jump is defined for one DaviesGenerator (trajectory_jump.jl:14)")
    D = Op.opensys_eig[1]; K = length(D.coupling.mats); c = ψ.c; B = size(ψ, 3)
    keys, ua, ub, a0, b0 = device_gap_indices(Op, p, t)                       # 1-based, reference order
    slot_ptr = Int64[0]; term = Int32[]; rate = Float64[]
    function slot!(i, rows, cols, r)                                         # entries sorted by (row, col), 0-based (coupling, row, col) triples
        for j in sortperm(collect(zip(rows, cols)))
            push!(term, i - 1, rows[j] - 1, cols[j] - 1)
        end
        push!(slot_ptr, length(term) ÷ 3); push!(rate, r)
    end
    for (g, w) in enumerate(keys), i in 1:K
        slot!(i, ua[g], ub[g], D.γ(w)); slot!(i, ub[g], ua[g], D.γ(-w))
    end
    for i in 1:K
        slot!(i, a0, b0, D.γ(0.0))
    end
    plan = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_liouv_prepare, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Ptr{Cvoid}}), liouv(Op), coef_rows(hf(Op.H), Float64[p(t)]), plan))
    check(c, ccall((:oqb_ame_set_jump, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}), plan[], length(rate), slot_ptr, term, rate))
    du = device_copy(c, uniforms); dc = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 4B, dc))
    check(c, ccall((:oqb_ame_jump, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   plan[], B, ψ.ptr, du, C_NULL, dc[], C_NULL, apply ? 1 : 0))
    choice = Vector{Int32}(undef, B)
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Csize_t), c, choice, dc[], 4B))
    ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, du); ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, dc[])
    choice
end

# ---------------------------------------------------------------- Redfield: apply step with Λ on the device
"du −= K₂ + K₂' with K₂ = Σ_α (S_α Λ_α u − Λ_α u S_α): the ρ-dependent half of redfield.jl:65-68; Λ (n×n×(K·B)) comes from the device builder (oqb_lambda_*)"
function redfield_apply!(du::DeviceBatch, u::DeviceBatch, S::DeviceBatch, Λ::DeviceBatch; real_couplings=false)
    n, K = size(S, 1), size(S, 3)
    check(u.c, ccall((:oqb_redfield_apply_acc, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                     u.c, n, K, S.ptr, 1 | (real_couplings ? 2 : 0), Λ.ptr, size(u, 3), u.ptr, du.ptr))
end

"cache (n²×n²×B) −= I⊗SΛ + conj(SΛ)⊗I − Sᵀ⊗Λ − conj(Λ)⊗S: update_vectorized_cache!(cache, R::RedfieldLiouvillian, p, t), redfield.jl:97-102 (ACCUMULATES)"
function redfield_vectorized!(cache::DeviceBatch, S::DeviceBatch, Λ::DeviceBatch)
    n, K = size(S, 1), size(S, 3)
    check(cache.c, ccall((:oqb_redfield_vectorized_acc, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                         cache.c, n, K, S.ptr, 1, Λ.ptr, size(cache, 3), cache.ptr))
end

"raw device copy of a host array (freed by the caller with oqb_free)"
function device_copy(c, a::Array)
    d = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, sizeof(a), d))
    check(c, ccall((:oqb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), c, d[], a, sizeof(a)))
    d[]
end

"The same with the couplings given as host matrices: diagonal S_α take the elementwise form, monomial ones (σx, σ±, Pauli strings: one entry per row and column) the gather form — K products per member instead of 3K"
function redfield_apply!(du::DeviceBatch, u::DeviceBatch, S::Vector{<:AbstractMatrix}, Λ::DeviceBatch)
    n, K, c = size(S[1], 1), length(S), u.c
    if all(isdiag, S)
        sd = DeviceBatch(c, reshape(ComplexF64[S[k][i, i] for i in 1:n, k in 1:K], n, 1, K))
        return check(c, ccall((:oqb_redfield_apply_diag_acc, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                              c, n, K, sd.ptr, Λ.ptr, size(u, 3), u.ptr, du.ptr))
    end
    if all(m -> all(<=(1), count(!iszero, m; dims=1)) && all(<=(1), count(!iszero, m; dims=2)), S)
        perm = fill(Int32(-1), n, K); pinv = fill(Int32(-1), n, K); val = zeros(ComplexF64, n, 1, K)
        for k in 1:K, idx in findall(!iszero, S[k])
            perm[idx[1], k] = idx[2] - 1; pinv[idx[2], k] = idx[1] - 1; val[idx[1], 1, k] = S[k][idx]
        end
        dperm, dpinv, dval = device_copy(c, perm), device_copy(c, pinv), DeviceBatch(c, val)
        st = ccall((:oqb_redfield_apply_mono_acc, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                   c, n, K, dperm, dpinv, dval.ptr, Λ.ptr, size(u, 3), u.ptr, du.ptr)
        ccall((:oqb_sync, LIB), Cint, (Ptr{Cvoid},), c)
        ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, dperm); ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, dpinv)
        return check(c, st)
    end
    Sd = DeviceBatch(c, reshape(reduce(hcat, [ComplexF64.(vec(m)) for m in S]), n, n, K))
    redfield_apply!(du, u, Sd, Λ; real_couplings=all(m -> all(isreal, m), S))
end

# ---------------------------------------------------------------- trajectories (ψ ensembles, n×1×B)
"dψ = (−iH(s_b) − ½Σγ_k L_k'L_k)ψ: the `cache*u` of the external solver with the cache of diffeq_liouvillian.jl:78-83"
function traj_matvec!(dψ::DeviceBatch, Op::DiffEqLiouvillian{false,false}, ψ::DeviceBatch, p, t)
    s = batch_s(p, t); lind = lindblad_of(Op)
    gam = lind === nothing ? C_NULL : Float64[g(sb) for g in lind.γ, sb in s]
    check(ψ.c, ccall((:oqb_traj_matvec, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                     dev(Op.H, lind), size(ψ, 3), coef_rows(hf(Op.H), s), shared(s), gam, shared(s), ψ.ptr, dψ.ptr))
end

"lindblad_jump(Op{false,false}, ψ, p, t) on a batch (trajectory_jump.jl:2-12, :71-74): `uniforms` are the rand() values; returns the chosen operator index per member (0-based) — apply=true performs ψ ← Lψ/‖Lψ‖"
function lindblad_jump(Op::DiffEqLiouvillian{false,false}, ψ::DeviceBatch, p, t::Real; uniforms::Vector{Float64}, apply=false)
    length(Op.opensys) == 1 || error("several Liouvillians: use lindblad_jump(Op, ψ, p, t, uniforms::Matrix) — one uniform per Liouvillian and one for resample")
    lind = Op.opensys[1]::LindbladLiouvillian
    c = ψ.c; B = size(ψ, 3); s = p(t)
    gam = Float64[g(s) for g in lind.γ]
    du = Ref{Ptr{Cvoid}}(C_NULL); dc = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 8B, du))
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 4B, dc))
    check(c, ccall((:oqb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Csize_t), c, du[], uniforms, 8B))
    check(c, ccall((:oqb_traj_jump, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                   dev(Op.H, lind), B, gam, 1, ψ.ptr, du[], C_NULL, dc[], C_NULL, apply ? 1 : 0))
    choice = Vector{Int32}(undef, B)
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Csize_t), c, choice, dc[], 4B))
    ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, du[]); ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, dc[])
    choice
end

"lindblad_jump with SEVERAL LindbladLiouvillians in Op.opensys (trajectory_jump.jl:71-79): Liouvillian m proposes one operator with its own
uniform `uniforms[:, m]`; `resample` (:81-88) draws the winner with `uniforms[:, M+1]` and the weights ‖L_mψ‖ (first power).  Returns
(pick, choices): the winning Liouvillian and every proposal per member, 0-based; apply=true performs ψ ← Lψ/‖Lψ‖ with the winner"
function lindblad_jump(Op::DiffEqLiouvillian{false,false}, ψ::DeviceBatch, p, t::Real, uniforms::Matrix{Float64}; apply=false)
    M = length(Op.opensys); c = ψ.c; B = size(ψ, 3); s = p(t)
    size(uniforms) == (B, M + 1) || error("uniforms must be B × (M+1): one column per Liouvillian and one for resample")
    dchoice = Ref{Ptr{Cvoid}}(C_NULL); dw = Ref{Ptr{Cvoid}}(C_NULL); dpick = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 4B * M, dchoice))
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 8B * M, dw))
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 4B, dpick))
    for m in 1:M
        lind = Op.opensys[m]::LindbladLiouvillian
        gam = Float64[g(s) for g in lind.γ]
        du = device_copy(c, uniforms[:, m])
        check(c, ccall((:oqb_traj_jump, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                       dev(Op.H, lind), B, gam, 1, ψ.ptr, du, C_NULL, dchoice[] + 4B * (m - 1), C_NULL, 0))
        check(c, ccall((:oqb_traj_apply_choice, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                       dev(Op.H, lind), B, ψ.ptr, dchoice[] + 4B * (m - 1), C_NULL, dw[] + 8B * (m - 1), 0))      # weight row m = ‖L_mψ‖
        ccall((:oqb_sync, LIB), Cint, (Ptr{Cvoid},), c); ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, du)
    end
    dur = device_copy(c, uniforms[:, M + 1])
    check(c, ccall((:oqb_resample, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                   c, B, M, dw[], dur, C_NULL, dpick[]))
    pick = Vector{Int32}(undef, B); choices = Matrix{Int32}(undef, B, M)
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Csize_t), c, pick, dpick[], 4B))
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Csize_t), c, choices, dchoice[], 4B * M))
    if apply
        for m in 1:M
            dmask = device_copy(c, UInt8.(pick .== m - 1))
            check(c, ccall((:oqb_traj_apply_choice, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                           dev(Op.H, Op.opensys[m]), B, ψ.ptr, dchoice[] + 4B * (m - 1), dmask, C_NULL, 1))
            ccall((:oqb_sync, LIB), Cint, (Ptr{Cvoid},), c); ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, dmask)
        end
    end
    for d in (dchoice[], dw[], dpick[], dur)
        ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, d)
    end
    pick, choices
end

"‖ψ_b‖² per member — the jump condition of the external callback"
function norm2(Op::DiffEqLiouvillian{false,false}, ψ::DeviceBatch)
    c = ψ.c; B = size(ψ, 3)
    d = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 8B, d))
    check(c, ccall((:oqb_traj_norm2, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), dev(Op.H, lindblad_of(Op)), B, ψ.ptr, d[]))
    out = Vector{Float64}(undef, B)
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Csize_t), c, out, d[], 8B))
    ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, d[])
    out
end

"out[o, b] = tr(O_o ρ_b) (states n×n×B) or ψ_b'O_oψ_b (states n×1×B) for dense observables obs (n×n×nobs on the device)"
function expect(obs::DeviceBatch, state::DeviceBatch)
    c = state.c; B = size(state, 3); nobs = size(obs, 3)
    d = Ref{Ptr{Cvoid}}(C_NULL)
    check(c, ccall((:oqb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), c, 16B * nobs, d))
    check(c, ccall((:oqb_expect, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                   c, size(obs, 1), nobs, obs.ptr, B, state.ptr, size(state, 2) == 1 ? 1 : 0, d[]))
    out = Matrix{ComplexF64}(undef, nobs, B)
    check(c, ccall((:oqb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}, Csize_t), c, out, d[], 16B * nobs))
    ccall((:oqb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), c, d[])
    out
end

# ---------------------------------------------------------------- multi-GPU: the ensemble-average reduce
"one communicator per GPU of this process (ncclCommInitAll); use with Threads.@threads over devices, one ctx(device) per thread"
function comm_init_all(ctxs::Vector{Ptr{Cvoid}})
    comms = Vector{Ptr{Cvoid}}(undef, length(ctxs))
    check(ctxs[1], ccall((:oqb_comm_init_all, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}), length(ctxs), ctxs, comms))
    comms
end
"out (device, 2·nobs doubles) = Σ over ALL GPUs of Σ_b ψ_b'O_oψ_b: per-GPU partial sums and ncclAllReduce on the context stream"
function ensemble_expect_allreduce!(comm::Ptr{Cvoid}, out::Ptr{Cvoid}, obs::DeviceBatch, ψ::DeviceBatch; is_vector=true)
    c = ψ.c
    check(c, ccall((:oqb_expect_sum, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                   c, size(obs, 1), size(obs, 3), obs.ptr, size(ψ, 3), ψ.ptr, is_vector ? 1 : 0, out))
    check(c, ccall((:oqb_allreduce_obs, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), comm, out, 2size(obs, 3)))
end

"per-schedule observables of ALL GPUs, rank-major (C5: one ncclAllGather per output point); send/recv are device pointers to count / count·nranks doubles"
allgather_obs!(comm::Ptr{Cvoid}, send::Ptr{Cvoid}, recv::Ptr{Cvoid}, count::Integer, c=ctx()) =
    check(c, ccall((:oqb_allgather_obs, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), comm, send, recv, count))
comm_destroy(comm::Ptr{Cvoid}) = ccall((:oqb_comm_destroy, LIB), Cint, (Ptr{Cvoid},), comm)

end # module
