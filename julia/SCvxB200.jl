# SCvxB200.jl — Julia `ccall` shim over libscvx_b200.so (C ABI: include/scvx_b200.h).
#
# Drop-in for the linearise-and-discretise path of BenChung/SuccessiveConvexification.  Include it
# AFTER the reference's master.jl; it adds methods to the reference's own entry points
#     Dynamics.linearize_dynamics(states, tf_guess, base_dt, cache)     (reference dynamics.jl:321-334)
#     Dynamics.predict_state(x, uk, up, sigma, dt, pinfo, cache)         (reference dynamics.jl:315-317)
# that dispatch on a cache whose `sim_prob` field (typed `Any`, master.jl:113-120) holds a
# `SCvxB200.Context`.  Everything else of the reference (SOCP assembly, Mosek solve, initial guess,
# trust-region logic) is untouched.
#
# UNVERIFIED UNDER JULIA: written without a Julia toolchain in the build container (see INTEGRATION.md).  It is
# mirrored 1:1 by the ctypes binding successiveconvexification_b200/_lib.py, which the test-suite exercises, and
# tests/test_abi.py checks its structure (every symbol bound, overrides inside the module).  tools/gen_golden.jl is the
# recipe that pins the device path against the reference's own rk4 + forward Jacobian on a machine that has Julia.
module SCvxB200

using ..RocketlandDefns
import ..Dynamics

const LIB = get(ENV, "SCVX_B200_LIB", joinpath(@__DIR__, "..", "successiveconvexification_b200", "libscvx_b200.so"))

const MODE_LITERAL = Cint(0)    # reproduces dynamics.jl:126-128
const MODE_TEXTBOOK = Cint(1)
const AERO_EXO = Cint(0)
const AERO_TABLE = Cint(1)

# mirror of scvx_probinfo
struct CProbInfo
    a::Cdouble; g0::Cdouble; sos::Cdouble
    jB::NTuple{9,Cdouble}; jBi::NTuple{9,Cdouble}
    rTB::NTuple{3,Cdouble}; rFB::NTuple{3,Cdouble}
    force_scalar::Cdouble; length_scalar::Cdouble
    Tmin::Cdouble
    aero_kind::Int32; _pad::Int32
end

function CProbInfo(info::ProbInfo, Tmin::Float64)
    atm = info.aero isa AtmosphericData
    CProbInfo(info.a, info.g0, info.sos, Tuple(info.jB), Tuple(info.jBi), Tuple(info.rTB), Tuple(info.rFB),
              atm ? info.aero.force_scalar : 0.0, atm ? info.aero.length_scalar : 0.0, Tmin,
              atm ? AERO_TABLE : AERO_EXO, Int32(0))
end

last_error() = unsafe_string(ccall((:scvx_last_error, LIB), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error("scvx_b200 error $rc: $(last_error())")   # reference style: rocketland.jl:275

device_count() = Int(ccall((:scvx_device_count, LIB), Cint, ()))
version() = Int(ccall((:scvx_version, LIB), Cint, ()))

# `mode` is the stage rule of the rk4-based entry points simulate_zygote / sensitivity_zygote and of the batched calls:
# LITERAL = the reference's arithmetic (dynamics.jl:126-128).  `live_mode` is the rule used when the device stands in for
# the reference's LIVE entry points linearize_dynamics / predict_state, which integrate the continuous dynamics with an
# adaptive BS3 solve (dynamics.jl:288-305): the consistent fixed-step integrator for those is classical RK4, so the
# default there is TEXTBOOK (the LITERAL rule is not a consistent integrator and would silently change SCvx iterates).
mutable struct Context
    handle::Ptr{Cvoid}
    npts::Int
    mode::Cint
    live_mode::Cint
    function Context(device_ids::Vector{Int}=[0]; npts::Int=10, mode::Cint=MODE_LITERAL, live_mode::Cint=MODE_TEXTBOOK)
        @assert ccall((:scvx_sizeof_probinfo, LIB), Cint, ()) == sizeof(CProbInfo)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        ids = Cint.(device_ids)
        check(ccall((:scvx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), h, ids, length(ids)))
        ctx = new(h[], npts, mode, live_mode)
        finalizer(c -> (c.handle != C_NULL && ccall((:scvx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), ctx)
        return ctx
    end
end

function set_params!(ctx::Context, recs::Vector{CProbInfo})
    check(ccall((:scvx_set_params, LIB), Cint, (Ptr{Cvoid}, Ptr{CProbInfo}, Cint), ctx.handle, recs, length(recs)))
end

# `samples`: the n_cos x n_mach matrix handed to `interpolate(...)` at aerodynamics.jl:19-21
function set_aero_table!(ctx::Context, which::Integer, samples::Matrix{Float64}, aoa::AbstractRange, mach::AbstractRange;
                         prefiltered::Bool=false)
    check(ccall((:scvx_set_aero_table, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint, Cdouble, Cdouble, Cdouble, Cdouble, Cint),
                ctx.handle, which, samples, length(aoa), length(mach), first(aoa), step(aoa), first(mach), step(mach),
                prefiltered ? 1 : 0))
end

function aero_coefficients(ctx::Context, which::Integer, n_cos::Integer, n_mach::Integer)
    out = Matrix{Float64}(undef, n_cos + 2, n_mach + 2)
    check(ccall((:scvx_get_aero_coefficients, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), ctx.handle, which, out))
    return out
end

set_kernel!(ctx::Context, which::Integer) = check(ccall((:scvx_set_kernel, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.handle, which))
# C_NULL = back to the library's own stream; the legacy default stream is cudaStreamLegacy = Ptr{Cvoid}(1)
set_stream!(ctx::Context, stream::Ptr{Cvoid}) = check(ccall((:scvx_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, stream))
synchronize(ctx::Context) = check(ccall((:scvx_synchronize, LIB), Cint, (Ptr{Cvoid},), ctx.handle))
launch_count(ctx::Context) = ccall((:scvx_launch_count, LIB), Int64, (Ptr{Cvoid},), ctx.handle)

function last_kernel_ms(ctx::Context)
    ms = Ref{Cdouble}(0.0)
    check(ccall((:scvx_last_kernel_ms, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), ctx.handle, ms))
    return ms[]
end

# Batched entry points.  X: 14 x n_nodes x B, U: 3 x n_nodes x B, sigma: B  (plain Julia arrays, column-major).
function linearize_batch(ctx::Context, X::Array{Float64,3}, U::Array{Float64,3}, sigma::Vector{Float64}, base_dt::Float64;
                         lin_err::Bool=true, tlb::Bool=true, mode::Cint=ctx.mode)
    n_nodes, B = size(X, 2), size(X, 3)
    blocks = Array{Float64,4}(undef, 14, 23, n_nodes - 1, B)
    err = lin_err ? Array{Float64,3}(undef, 14, n_nodes - 1, B) : Array{Float64,3}(undef, 0, 0, 0)
    tl = tlb ? Array{Float64,3}(undef, 4, n_nodes, B) : Array{Float64,3}(undef, 0, 0, 0)
    GC.@preserve X U sigma blocks err tl begin
        check(ccall((:scvx_linearize_batch, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cint,
                     Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, X, U, sigma, base_dt, ctx.npts, mode, n_nodes, B,
                    blocks, lin_err ? pointer(err) : Ptr{Cdouble}(C_NULL), tlb ? pointer(tl) : Ptr{Cdouble}(C_NULL)))
    end
    return blocks, err, tl
end

# Compact result records (scvx_linearize_batch_compact): only the 229 data entries of every 14 x 23 block (+ a
# per-interval non-finite flag) cross PCIe.  `expand_compact` restores the dense blocks and lin_err on the host.
const COMPACT_DOUBLES = 230
const COMPACT_DATA = 229
const COMPACT_FULL = Cint(0)     # 229 data entries + status word
const COMPACT_NO_Z = Cint(1)     # without the z column (the SOCP consumes D and lin_err only): 215 + status word = 216

compact_record_doubles(layout::Integer) = Int(ccall((:scvx_compact_record_doubles, LIB), Cint, (Cint,), layout))

function linearize_batch_compact(ctx::Context, X::Array{Float64,3}, U::Array{Float64,3}, sigma::Vector{Float64},
                                 base_dt::Float64; tlb::Bool=true, mode::Cint=ctx.mode, layout::Cint=COMPACT_FULL,
                                 out::Union{Nothing,Array{Float64,3}}=nothing)
    n_nodes, B = size(X, 2), size(X, 3)
    comp = out === nothing ? Array{Float64,3}(undef, compact_record_doubles(layout), n_nodes - 1, B) : out
    tl = tlb ? Array{Float64,3}(undef, 4, n_nodes, B) : Array{Float64,3}(undef, 0, 0, 0)
    GC.@preserve X U sigma comp tl begin
        check(ccall((:scvx_linearize_batch_compact, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cint, Cint,
                     Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, X, U, sigma, base_dt, ctx.npts, mode, n_nodes, B, layout,
                    comp, tlb ? pointer(tl) : Ptr{Cdouble}(C_NULL)))
    end
    return comp, tl
end

# dense offsets (0-based, column * 14 + row) of compact slots 1..229
function compact_layout()
    idx = Vector{Int32}(undef, COMPACT_DATA)
    check(ccall((:scvx_compact_layout, LIB), Cint, (Ptr{Int32},), idx))
    return idx
end

# -> blocks 14 x 23 x K x B, lin_err 14 x K x B, number of intervals flagged non-finite
function expand_compact(compact::Array{Float64,3}, X::Array{Float64,3}, U::Array{Float64,3}, sigma::Vector{Float64};
                        n_threads::Int=0)
    K, B = size(compact, 2), size(compact, 3)
    layout = size(compact, 1) == COMPACT_DOUBLES ? COMPACT_FULL : COMPACT_NO_Z
    blocks = Array{Float64,4}(undef, 14, 23, K, B); err = Array{Float64,3}(undef, 14, K, B)
    n = GC.@preserve compact X U sigma blocks err ccall((:scvx_expand_compact, LIB), Int64,
            (Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
            compact, layout, X, U, sigma, K + 1, B, blocks, err, n_threads)
    n < 0 && check(n)
    return blocks, err, Int(n)
end

# Page-locked host memory: the H2D / kernel / D2H pipeline of host-pointer calls only overlaps for pinned buffers.
# `pin!(A)` registers a Julia array in place (keep it alive and call `unpin!(A)` before it is freed);
# `pinned_array(T, dims...)` allocates one through the library (free with `free_pinned!`).
pin!(A::Array) = (check(ccall((:scvx_host_register, LIB), Cint, (Ptr{Cvoid}, UInt64), A, sizeof(A))); A)
unpin!(A::Array) = (check(ccall((:scvx_host_unregister, LIB), Cint, (Ptr{Cvoid},), A)); A)
function pinned_array(::Type{T}, dims::Integer...) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:scvx_host_alloc, LIB), Cint, (Ref{Ptr{Cvoid}}, UInt64), p, prod(dims) * sizeof(T)))
    return unsafe_wrap(Array, Ptr{T}(p[]), dims; own=false)
end
free_pinned!(A::Array) = check(ccall((:scvx_host_free, LIB), Cint, (Ptr{Cvoid},), A))

# SURVEY §8f-4 variant: fin forces + aero torque, control_dim = 5 (the terms commented out at dynamics.jl:60-63, 66, 69).
# NO REFERENCE CONSUMER.  X 14 x n x B, U5 5 x n x B  ->  blocks 14 x 27 x (n-1) x B = [endpoint | D (25) | z], lin_err.
function linearize_batch_fins(ctx::Context, X::Array{Float64,3}, U5::Array{Float64,3}, sigma::Vector{Float64}, base_dt::Float64;
                              mode::Cint=ctx.mode)
    n_nodes, B = size(X, 2), size(X, 3)
    blocks = Array{Float64,4}(undef, 14, 27, n_nodes - 1, B); err = Array{Float64,3}(undef, 14, n_nodes - 1, B)
    GC.@preserve X U5 sigma blocks err begin
        check(ccall((:scvx_linearize_batch_fins, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, X, U5, sigma, base_dt, ctx.npts, mode, n_nodes, B, blocks, err))
    end
    return blocks, err
end

# fin-force tables of aero/fin.csv (n_mach x n_defl, Mach fastest): which = 0 lift, 1 drag
function set_fin_table!(ctx::Context, which::Integer, samples::Matrix{Float64}, mach::AbstractRange, defl::AbstractRange)
    check(ccall((:scvx_set_fin_table, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint, Cdouble, Cdouble, Cdouble, Cdouble, Cint),
                ctx.handle, which, samples, length(mach), length(defl), first(mach), step(mach), first(defl), step(defl), 0))
end

function fin_force_batch(ctx::Context, mach::Vector{Float64}, deflection::Vector{Float64})
    n = length(mach); lift = Vector{Float64}(undef, n); drag = Vector{Float64}(undef, n)
    GC.@preserve mach deflection lift drag begin
        check(ccall((:scvx_fin_force_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, mach, deflection, n, lift, drag))
    end
    return lift, drag
end

function predict_batch(ctx::Context, X::Array{Float64,3}, U::Array{Float64,3}, sigma::Vector{Float64}, base_dt::Float64;
                       mode::Cint=ctx.mode)
    n_nodes, B = size(X, 2), size(X, 3)
    out = Array{Float64,3}(undef, 14, n_nodes - 1, B)
    GC.@preserve X U sigma out begin
        check(ccall((:scvx_predict_batch, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cdouble}),
                    ctx.handle, X, U, sigma, base_dt, ctx.npts, mode, n_nodes, B, out))
    end
    return out
end

# Fused cost / defect evaluation of the ratio test (reference rocketland.jl:289-290): per trajectory
# defect = norm(x_{k+1} - endpoint_k for k) and J = -x[1,K+1] + wNu * defect, from linearize_batch's lin_err.
function defect_cost_batch(ctx::Context, X::Array{Float64,3}, lin_err::Array{Float64,3}, wNu::Float64)
    n_nodes, B = size(X, 2), size(X, 3)
    defect = Vector{Float64}(undef, B); cost = Vector{Float64}(undef, B)
    GC.@preserve X lin_err defect cost begin
        check(ccall((:scvx_defect_cost_batch, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, X, lin_err, n_nodes, B, wNu, defect, cost))
    end
    return defect, cost
end

# Batched initial guess: FirstRound.linear_points (reference initial_solve.jl:113-129) for B dispersed initial conditions.
# rIi, vIi: 3 x B; mwet: per-trajectory wet masses or nothing.
function linear_points_batch(ctx::Context, prob::DescentProblem, rIi::Matrix{Float64}, vIi::Matrix{Float64};
                             mwet::Union{Nothing,Vector{Float64}}=nothing)
    B, K = size(rIi, 2), prob.K
    X = Array{Float64,3}(undef, 14, K + 1, B); U = Array{Float64,3}(undef, 3, K + 1, B)
    rIf = Float64.(prob.rIf); vIf = Float64.(prob.vIf)
    GC.@preserve rIi vIi mwet rIf vIf X U begin
        check(ccall((:scvx_linear_points_batch, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble},
                     Cdouble, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, rIi, vIi, mwet === nothing ? Ptr{Cdouble}(C_NULL) : pointer(mwet), prob.mwet, prob.mdry,
                    rIf, vIf, prob.g, K, B, X, U))
    end
    return X, U
end

# Dispersion set-up on the device (SURVEY §8f-3): per-trajectory normalize_problem (sample_problems.jl:5-23), ProbInfo
# (master.jl:73-83) and linear_points (initial_solve.jl:113-129) for B dispersed DIMENSIONAL initial conditions.
struct CDimProblem
    g::Cdouble; mdry::Cdouble; mwet::Cdouble; Tmin::Cdouble; Tmax::Cdouble; alpha::Cdouble; sos::Cdouble; tf_guess::Cdouble
    jB::NTuple{9,Cdouble}; rTB::NTuple{3,Cdouble}; rFB::NTuple{3,Cdouble}; rIf::NTuple{3,Cdouble}
    aero_kind::Int32; K::Int32
end
CDimProblem(p::DescentProblem) = CDimProblem(p.g, p.mdry, p.mwet, p.Tmin, p.Tmax, p.alpha, p.sos, p.tf_guess,
    Tuple(Float64.(vec(Matrix(p.jB)))), Tuple(Float64.(p.rTB)), Tuple(Float64.(p.rFB)), Tuple(Float64.(p.rIf)),
    Int32(p.aero isa AtmosphericData ? 1 : 0), Int32(p.K))

function dispersed_setup!(ctx::Context, prob::DescentProblem, rIi::Matrix{Float64}, vIi::Matrix{Float64};
                          mwet::Union{Nothing,Vector{Float64}}=nothing, install::Bool=true)
    B, K = size(rIi, 2), prob.K
    X = Array{Float64,3}(undef, 14, K + 1, B); U = Array{Float64,3}(undef, 3, K + 1, B)
    sigma = Vector{Float64}(undef, B); scales = Matrix{Float64}(undef, 3, B)
    base = Ref(CDimProblem(prob))
    GC.@preserve rIi vIi mwet X U sigma scales base begin
        check(ccall((:scvx_dispersed_setup_batch, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{CDimProblem}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble},
                     Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}, Cint),
                    ctx.handle, base, rIi, vIi, mwet === nothing ? Ptr{Cdouble}(C_NULL) : pointer(mwet), B, X, U, sigma, scales,
                    C_NULL, install ? 1 : 0))
    end
    return X, U, sigma, scales
end

# Fixed-pattern sparse form of the trajectory-dependent SOCP rows (reference rocketland.jl:117-133, 194-201, 251-265):
# CSC pattern (1-based for SparseMatrixCSC) and, per trajectory, the value array + constants.  Local column j is the
# reference's variable 17(K+1) + j (dxv, duv, dsig, nuv are created consecutively, rocketland.jl:73-76).
function socp_pattern(n_nodes::Integer)
    nr = Ref{Cint}(0); nc = Ref{Cint}(0); nz = Ref{Cint}(0)
    check(ccall((:scvx_socp_dims, LIB), Cint, (Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), n_nodes, nr, nc, nz))
    colptr = Vector{Int32}(undef, nc[] + 1); rowind = Vector{Int32}(undef, nz[])
    check(ccall((:scvx_socp_pattern, LIB), Cint, (Cint, Ptr{Int32}, Ptr{Int32}), n_nodes, colptr, rowind))
    return Int(nr[]), Int(nc[]), colptr .+ Int32(1), rowind .+ Int32(1)
end

function socp_values_batch(ctx::Context, blocks::Array{Float64,4}, lin_err::Array{Float64,3}, tlb::Array{Float64,3})
    n_nodes, B = size(tlb, 2), size(tlb, 3)
    nr = Ref{Cint}(0); nc = Ref{Cint}(0); nz = Ref{Cint}(0)
    check(ccall((:scvx_socp_dims, LIB), Cint, (Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), n_nodes, nr, nc, nz))
    vals = Matrix{Float64}(undef, nz[], B); rhs = Matrix{Float64}(undef, nr[], B)
    GC.@preserve blocks lin_err tlb vals rhs begin
        check(ccall((:scvx_socp_values_batch, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                    ctx.handle, blocks, lin_err, tlb, n_nodes, B, vals, rhs))
    end
    return vals, rhs
end

# IntegratorCache(prob, info) replacement (reference dynamics.jl:258-260): context + parameters + tables.
# `aero_samples = (drag, lift, torque, aoa_range, mach_range)` are the matrices / ranges of aerodynamics.jl:17-21.
function make_cache(prob::DescentProblem, info::ProbInfo; device_ids::Vector{Int}=[0], aero_samples=nothing,
                    mode::Cint=MODE_LITERAL, live_mode::Cint=MODE_TEXTBOOK)
    ctx = Context(device_ids; mode=mode, live_mode=live_mode)
    set_params!(ctx, [CProbInfo(info, prob.Tmin)])
    if info.aero isa AtmosphericData
        aero_samples === nothing && error("AtmosphericData needs aero_samples = (drag, lift, torque, aoa, mach)")
        drag, lift, trq, aoa, mach = aero_samples
        set_aero_table!(ctx, 0, drag, aoa, mach); set_aero_table!(ctx, 1, lift, aoa, mach); set_aero_table!(ctx, 2, trq, aoa, mach)
    end
    return IntegratorCache(ctx, nothing, nothing, nothing, Any[1.0, info], info)
end



# ---- methods added to the reference's own entry points -------------------------------------------------
# Defined INSIDE this module: `using ..RocketlandDefns` brings LinPoint / LinRes / IntegratorCache into scope here
# (master.jl exports them from RocketlandDefns but never does `using .RocketlandDefns` in Main, so unqualified names at
# top level would be undefined), and `import ..Dynamics` lets `function Dynamics.f(...)` add methods to its functions.
# The signatures equal the reference's, so these definitions REPLACE its methods (Julia prints a method-overwrite
# note); the replacement only acts on caches built by `make_cache` — any other cache raises an error.
function _scvx_ctx(cache::IntegratorCache)
    cache.sim_prob isa Context || error("cache does not hold a Context")
    return cache.sim_prob::Context
end

# Replaces the method of dynamics.jl:321-334 (same signature); the cache must hold a device context.
function Dynamics.linearize_dynamics(states::Array{LinPoint,1}, tf_guess::Float64, base_dt::Float64, cache::IntegratorCache)
    ctx = _scvx_ctx(cache)
    n = length(states)
    X = Array{Float64,3}(undef, 14, n, 1); U = Array{Float64,3}(undef, 3, n, 1)
    for i = 1:n
        X[:, i, 1] .= states[i].state; U[:, i, 1] .= states[i].control
    end
    blocks, _, _ = linearize_batch(ctx, X, U, [tf_guess], base_dt; lin_err=false, tlb=false, mode=ctx.live_mode)
    return [LinRes(blocks[:, 1, i, 1], blocks[:, 2:22, i, 1]) for i = 1:n-1]
end

# Replaces dynamics.jl:315-317.
function Dynamics.predict_state(initial_state, uk, up, sigma, dt, pinfo, cache::IntegratorCache)
    ctx = _scvx_ctx(cache)
    X = zeros(14, 2, 1); U = zeros(3, 2, 1)
    X[:, 1, 1] .= initial_state; U[:, 1, 1] .= uk; U[:, 2, 1] .= up
    return predict_batch(ctx, X, U, [Float64(sigma)], Float64(dt); mode=ctx.live_mode)[:, 1, 1]
end

# Replace dynamics.jl:308-313 (same signatures): value and (y, J') of the discrete map of one interval on the device.
function _scvx_one_interval(inp::Vector{Float64})
    X = zeros(14, 2, 1); U = zeros(3, 2, 1)
    X[:, 1, 1] .= inp[1:14]; U[:, 1, 1] .= inp[15:17]; U[:, 2, 1] .= inp[18:20]
    return X, U, [inp[21]]
end

function Dynamics.simulate_zygote(inp::Vector{Float64}, dt::Float64, cache::IntegratorCache; npts=10)
    ctx = _scvx_ctx(cache)
    X, U, sig = _scvx_one_interval(inp)
    old = ctx.npts; ctx.npts = npts
    y = predict_batch(ctx, X, U, sig, dt)[:, 1, 1]
    ctx.npts = old
    return y
end

function Dynamics.sensitivity_zygote(inp::Vector{Float64}, dt::Float64, cache::IntegratorCache)
    ctx = _scvx_ctx(cache)
    X, U, sig = _scvx_one_interval(inp)
    blocks, _, _ = linearize_batch(ctx, X, U, sig, dt; lin_err=false, tlb=false)
    return blocks[:, 1, 1, 1], permutedims(blocks[:, 2:22, 1, 1])      # (y, J') with J' 21 x 14, as Zygote.forward_jacobian
end

end # module
