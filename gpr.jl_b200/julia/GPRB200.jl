# GPRB200.jl - routes the GP calls of GPR.jl's experiments to libgprb200.so WITHOUT changing their types.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  The tested contract is the C ABI
# (include/gprb200.h, exercised through ctypes by tests/); every ccall below mirrors one prototype of that header
# one-to-one (tests/test_julia_shim_signatures.py parses this file and checks every ccall tuple against the header),
# and gpr.jl_b200/lib.py + gp.py are the executable twin.  Names and signatures of GaussianProcesses.jl 0.12.4
# internals (GPE type parameters, alloc_cK, the 8-argument GPE constructor, init_precompute) are written from the
# published source of that version (Manifest.toml:409-413) and could not be compiled here.
#
# Drop-in mechanism: GaussianProcesses.jl's own extension point for the linear-algebra back end is the covariance
# strategy (`GPE(...; covstrat)`, used by its sparse approximations).  `B200Covariance <: CovarianceStrategy` keeps the
# object a real `GaussianProcesses.GPE`, so the reference's typed call sites work unchanged:
#     gps = Vector{GPE}()                                              examples/maximal_coordinates/CPnoise.jl:35
#     predictdynamics(mechanism, gps::Vector{<:GPE}, x0, steps, getvω)   examples/utils/predictdynamics.jl:7
# Nothing is exported that GaussianProcesses / Optim export (`using GaussianProcesses, GPRB200` has no clashes).
#
# Per-GP flavour - the ONLY edit in an experiment body (CPnoise.jl:40): the constructor gets the strategy
#     kernel = SEArd(log.(params[2:end]), log(params[1]))                          # :38 unchanged
#     gp = GP(xtrain_old, yi, mean, kernel, B200Covariance())                      # :40  (+ one argument)
#     GaussianProcesses.optimize!(gp, LBFGS(linesearch = BackTracking(order=2)), Optim.Options(time_limit=10.))   # :41 unchanged
#     μ = predict_y(gp, obs)[1][1]                                                 # predictdynamics.jl:13 unchanged
# Batched flavour (what replaces the `Threads.@threads for jobid` loop of examples/parallel/core.jl:28):
#     gps = [GPE(X_t, y_tk, mean_tk, SEArd(...), -2.0, B200Covariance()) for t in trials for k in outputs]   # no evaluation yet
#     GaussianProcesses.optimize!(gps, LBFGS(linesearch = BackTracking(order=2)), Optim.Options(time_limit=10.))  # ONE gprb_optimize
#     predictdynamics(mechanism, gps[(t-1)*G+1:t*G], x0, steps, getvω)            # unchanged; the G predict_y calls of a step share one device call
module GPRB200

using Libdl
import GaussianProcesses
import GaussianProcesses: GPE, Mean, MeanZero, Kernel, SEArd, Mat12Ard, Mat32Ard, Mat52Ard, CovarianceStrategy, KernelData, EmptyData,
                          get_params, set_params!, num_params
import Optim
import PDMats

export B200Covariance, GPBatch, params_to_theta, theta_to_params, gather_results

const LIB = get(ENV, "GPRB200_LIB", joinpath(@__DIR__, "..", "libgprb200.so"))

# ---- status handling (no exceptions cross the C boundary; we raise on the Julia side) --------------------------------
struct GprbError <: Exception
    code::Cint
    msg::String
end
last_error() = unsafe_string(ccall((:gprb_last_error, LIB), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(GprbError(rc, last_error()))

# ---- context: one per process and GPU --------------------------------------------------------------------------------
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context()
    if CTX[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        dev = parse(Cint, get(ENV, "GPRB200_DEVICE", get(ENV, "LOCAL_RANK", "0")))
        check(ccall((:gprb_init, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, dev))     # fails without a B200: no CPU fallback
        CTX[] = h[]
    end
    CTX[]
end

kernel_kind(::SEArd) = Cint(0)
kernel_kind(::Mat12Ard) = Cint(1)
kernel_kind(::Mat32Ard) = Cint(2)
kernel_kind(::Mat52Ard) = Cint(3)

# ---- the covariance strategy -----------------------------------------------------------------------------------------
"Covariance strategy that keeps K, its factor and alpha in HBM (libgprb200.so) instead of a host PDMat."
struct B200Covariance <: CovarianceStrategy end

# gp.cK is never read by the reference's callers; a 1-byte placeholder satisfies the field type (no n x n host allocation,
# no n x n x d distance stack: KernelData is EmptyData)
GaussianProcesses.alloc_cK(::B200Covariance, nobs::Int) = PDMats.ScalMat(nobs, 1.0)

const B200GPE = GPE{<:AbstractMatrix,<:AbstractVector,<:Mean,<:Kernel,<:B200Covariance}

"`GPE(X, y, mean, kernel, logNoise, B200Covariance())`: a GaussianProcesses.GPE whose linear algebra lives on the B200 (no evaluation yet)."
GaussianProcesses.GPE(x::AbstractMatrix, y::AbstractVector, mean::Mean, kernel::Kernel, logNoise::Real, cs::B200Covariance) =
    GPE(Matrix{Float64}(x), Vector{Float64}(y), mean, kernel, Float64(logNoise), cs, EmptyData(), GaussianProcesses.alloc_cK(cs, length(y)))

"`GP(X, y, mean, kernel, B200Covariance())`: construct + initial update_mll!, like GaussianProcesses.GP (CPnoise.jl:40)."
function GaussianProcesses.GP(x::AbstractMatrix, y::AbstractVector, mean::Mean, kernel::Kernel, cs::B200Covariance; logNoise::Real = -2.0)
    gp = GPE(x, y, mean, kernel, logNoise, cs)
    GaussianProcesses.update_mll!(gp)
    gp
end

# GaussianProcesses.get_params order: [logNoise; mean params; kernel params] - the reference's means have none (src/mDynamics.jl:29-31)
theta_of(gp::GPE) = Vector{Float64}(get_params(gp))

# ---- GPBatch: the device residence of one or many GPEs ----------------------------------------------------------------
mutable struct GPBatch
    gps::Vector{GPE}
    handle::Ptr{Cvoid}
    datasets::Vector{Ptr{Cvoid}}
    B::Int; n::Int; d::Int; P::Int
    # shared prediction cache of one rollout step: the G predict_y(gp, obs) calls of predictdynamics.jl:13 hit one device call
    cache_key::Matrix{Float64}
    cache_mu::Matrix{Float64}
    cache_var::Matrix{Float64}
end

# gp -> (batch, slot); weak, so dropping the GPEs frees the device memory through the batch finalizer
const RESIDENT = WeakKeyDict{GPE,Tuple{GPBatch,Int}}()

function GPBatch(gps::Vector{<:GPE})
    ctx = context()
    d, n = gps[1].dim, gps[1].nobs
    B = length(gps)
    # all distinct training sets of the batch in ONE device allocation, uploaded as one contiguous block
    uniq = IdDict{Any,Int}()
    for gp in gps
        get!(uniq, gp.x, length(uniq) + 1)
    end
    T = length(uniq)
    block = Array{Float64,3}(undef, d, n, T)
    for (x, t) in uniq
        block[:, :, t] = x
    end
    ptrs = [pointer(block, (t - 1) * d * n + 1) for t in 1:T]
    dsh = Vector{Ptr{Cvoid}}(undef, T)
    GC.@preserve block begin
        check(ccall((:gprb_datasets_create, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Int32, Ptr{Ptr{Float64}}, Int64, Ptr{Ptr{Cvoid}}),
                    ctx, T, n, d, ptrs, d, dsh))
    end
    handles = [dsh[uniq[gp.x]] for gp in gps]
    # m(X) does not depend on θ: evaluate once per training set, column by column across the GPs of a trial so the shared
    # MDCache (src/mDynamics.jl:6-11,42) hits for the other G-1 outputs.  Only y - m(X) goes to the device.
    ymm = Matrix{Float64}(undef, n, B)
    for (x, _) in uniq
        members = [b for b in 1:B if gps[b].x === x]
        for j in 1:n, b in members
            ymm[j, b] = gps[b].y[j] - GaussianProcesses.mean(gps[b].mean, gps[b].x[:, j])
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gprb_batch_create, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
                ctx, B, handles, ymm, kernel_kind(gps[1].kernel), h))
    batch = GPBatch(Vector{GPE}(gps), h[], dsh, B, n, d, d + 2, zeros(0, 0), zeros(0, 0), zeros(0, 0))
    for (b, gp) in enumerate(gps)
        RESIDENT[gp] = (batch, b)
    end
    finalizer(batch) do bt
        ccall((:gprb_batch_destroy, LIB), Cint, (Ptr{Cvoid},), bt.handle)
        foreach(ds -> ccall((:gprb_dataset_destroy, LIB), Cint, (Ptr{Cvoid},), ds), bt.datasets)
    end
    batch
end

"The batch a GPE lives in (a batch of one is created on demand: the reference's per-GP call pattern)."
function residence(gp::GPE)
    haskey(RESIDENT, gp) || GPBatch(GPE[gp])
    RESIDENT[gp]
end

thetas(batch::GPBatch) = reduce(hcat, theta_of.(batch.gps))          # P × B, column per GP: the layout gprb_eval takes

"One objective evaluation per active GP of the batch; writes mll / dmll / alpha back into the GPE fields the callers read."
function evaluate!(batch::GPBatch; θ::Matrix{Float64} = thetas(batch), grad::Bool = true, active::Union{Nothing,Vector{UInt8}} = nothing)
    mll = fill(NaN, batch.B)
    g = grad ? fill(NaN, batch.P, batch.B) : nothing
    info = zeros(Int32, batch.B)
    GC.@preserve θ mll g info active begin
        check(ccall((:gprb_eval, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    batch.handle, θ, active === nothing ? C_NULL : pointer(active), mll, grad ? pointer(g) : C_NULL, info))
    end
    for (b, gp) in enumerate(batch.gps)
        (active === nothing || active[b] != 0) || continue
        gp.mll = mll[b]; gp.target = mll[b]                       # no priors are set anywhere in the reference
        if grad
            gp.dmll = g[:, b]; gp.dtarget = g[:, b]
        end
        info[b] < 0 && throw(PDMats.PosDefException(Int(info[b])))   # get_optim_target's catch branch turns this into +Inf
    end
    batch.cache_key = zeros(0, 0)
    mll, g, info
end

function only_active(batch::GPBatch, slot::Int)
    a = zeros(UInt8, batch.B)
    a[slot] = 1
    a
end

# the two evaluation entry points GaussianProcesses.optimize! / GP(...) reach (GPE.jl: update_mll!, update_mll_and_dmll!)
function GaussianProcesses.update_mll!(gp::B200GPE; kwargs...)
    batch, slot = residence(gp)
    evaluate!(batch; grad = false, active = batch.B == 1 ? nothing : only_active(batch, slot))
    gp
end
function GaussianProcesses.update_mll_and_dmll!(gp::B200GPE, precomp...; kwargs...)
    batch, slot = residence(gp)
    evaluate!(batch; grad = true, active = batch.B == 1 ? nothing : only_active(batch, slot))
    gp
end
GaussianProcesses.init_precompute(gp::B200GPE) = nothing     # no n x n host work buffer

"gp.alpha on demand (the field is filled lazily: the rollout never reads it)."
function alpha!(gp::B200GPE)
    batch, slot = residence(gp)
    a = Vector{Float64}(undef, batch.n)
    check(ccall((:gprb_get_alpha, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), batch.handle, slot - 1, a))
    gp.alpha = a
end

# ---- optimize! -------------------------------------------------------------------------------------------------------
struct LbfgsOpts                 # gprb_lbfgs_opts
    m::Int32; iterations::Int32; max_evals::Int32; ls_iterations::Int32
    g_abstol::Float64; time_limit::Float64; c_1::Float64; rho_hi::Float64; rho_lo::Float64
    cost_value::Float64; cost_grad::Float64
end
struct OptResult                 # gprb_opt_result
    mll::Float64; g_norm::Float64
    iterations::Int32; f_calls::Int32; fg_calls::Int32; converged::Int32; ls_failed::Int32; info::Int32
end

"""
    GaussianProcesses.optimize!(gps::Vector{<:GPE}, method::Optim.LBFGS, options::Optim.Options; max_evals, cost_value, cost_grad)

Batched form of the reference's `GaussianProcesses.optimize!(gp, LBFGS(linesearch=BackTracking(order=2)),
Optim.Options(time_limit=10.))` (CPnoise.jl:41): all GPs advance in lock-step inside ONE gprb_optimize call.
`options.time_limit` is the reference's per-GP 10 s: with `cost_value` / `cost_grad` (seconds one value-only / value+gradient
evaluation takes on the CPU being emulated) it runs on a deterministic per-GP virtual clock; without them it is the wall
clock of the whole batch.  `max_evals` is a per-GP evaluation cap.
"""
function GaussianProcesses.optimize!(gps::Vector{<:GPE}, method::Optim.LBFGS = Optim.LBFGS(), options::Optim.Options = Optim.Options();
                                     max_evals::Integer = 0, cost_value::Real = 0.0, cost_grad::Real = 0.0)
    batch = (haskey(RESIDENT, gps[1]) && RESIDENT[gps[1]][1].gps == gps) ? RESIDENT[gps[1]][1] : GPBatch(gps)
    ls = method.linesearch!
    o = LbfgsOpts(method.m, options.iterations, max_evals, ls.iterations, options.g_abstol,
                  isfinite(options.time_limit) ? options.time_limit : 0.0, ls.c_1, ls.ρ_hi, ls.ρ_lo, cost_value, cost_grad)
    θ = thetas(batch)
    res = Vector{OptResult}(undef, batch.B)
    check(ccall((:gprb_optimize, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{LbfgsOpts}, Ptr{OptResult}), batch.handle, θ, o, res))
    for (b, gp) in enumerate(batch.gps)
        set_params!(gp, θ[:, b])
        gp.mll = res[b].mll; gp.target = res[b].mll
    end
    batch.cache_key = zeros(0, 0)
    res
end
# the reference's per-GP call (CPnoise.jl:41) on a B200-resident GPE: a batch of one through the same driver
GaussianProcesses.optimize!(gp::B200GPE, method::Optim.LBFGS = Optim.LBFGS(), options::Optim.Options = Optim.Options(); kw...) =
    GaussianProcesses.optimize!(GPE[gp], method, options; kw...)[1]

# ---- predict_y -------------------------------------------------------------------------------------------------------
"predict_y(batch, Xstar::d×m) -> (μ::m×B, σ²::m×B | nothing); `var=false` skips the variance the reference discards. GPs without state give NaN columns."
function GaussianProcesses.predict_y(batch::GPBatch, Xstar::AbstractMatrix; var::Bool = true)
    Xs = Matrix{Float64}(Xstar)
    m = size(Xs, 2)
    mstar = nothing
    if !all(gp -> gp.mean isa MeanZero, batch.gps)
        mstar = Matrix{Float64}(undef, m, batch.B)
        for j in 1:m, b in 1:batch.B            # column-major over GPs: MDCache semantics of src/mDynamics.jl:41-55
            mstar[j, b] = GaussianProcesses.mean(batch.gps[b].mean, Xs[:, j])
        end
    end
    μ = Matrix{Float64}(undef, m, batch.B)
    σ2 = var ? Matrix{Float64}(undef, m, batch.B) : nothing
    GC.@preserve Xs mstar μ σ2 begin
        check(ccall((:gprb_predict, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    batch.handle, m, Xs, 0, mstar === nothing ? C_NULL : pointer(mstar), μ, var ? pointer(σ2) : C_NULL))
    end
    μ, σ2
end

# The reference's call `predict_y(gp, obs)[1][1]` (predictdynamics.jl:13), once per GP of the trial with the SAME obs:
# the first call predicts every GP of the batch on the device, the other G-1 calls are served from the step cache
# (the device-side twin of MDCache, src/mDynamics.jl:6-11).
function GaussianProcesses.predict_y(gp::B200GPE, Xstar::AbstractMatrix)
    batch, slot = residence(gp)
    if size(batch.cache_key) != size(Xstar) || batch.cache_key != Xstar
        μ, σ2 = GaussianProcesses.predict_y(batch, Xstar; var = true)
        batch.cache_key, batch.cache_mu, batch.cache_var = Matrix{Float64}(Xstar), μ, σ2
    end
    batch.cache_mu[:, slot], batch.cache_var[:, slot]
end

# ---- asynchronous rollout step (two trial groups alternate: device predict of one under the host projectv! of the other) -----
"Enqueue the mean prediction of GPs gp0:gp1 (1-based, inclusive) at the per-trial state blocks `states` (d × m × trials) on pipeline `slot`."
function predict_async!(batch::GPBatch, slot::Integer, gp0::Integer, gp1::Integer, states::Array{Float64,3}, G::Integer; var::Bool = false)
    d, m, _ = size(states)
    GC.@preserve states begin
        check(ccall((:gprb_predict_async, LIB), Cint,
                    (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Int32),
                    batch.handle, slot, gp0 - 1, gp1, m, states, d * m, G, C_NULL, var ? 1 : 0))
    end
end
function predict_wait!(batch::GPBatch, slot::Integer, count::Integer, m::Integer; var::Bool = false)
    μ = Matrix{Float64}(undef, m, count)
    σ2 = var ? Matrix{Float64}(undef, m, count) : nothing
    check(ccall((:gprb_predict_wait, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
                batch.handle, slot, μ, var ? pointer(σ2) : C_NULL))
    μ, σ2
end

# ---- glue the experiment drivers need ---------------------------------------------------------------------------------
"config.json order `[σ_f, ℓ_1..ℓ_d]` (examples/config/README) -> GaussianProcesses order `[logNoise, ll_1..ll_d, lσ]` (CPnoise.jl:38)."
params_to_theta(params::AbstractVector; logNoise::Real = -2.0) = vcat(Float64(logNoise), log.(params[2:end]), log(params[1]))
"Inverse of params_to_theta: what parallelsearch stores in `config[\"params\"]` / the JSON checkpoints (core.jl:94-112, utils.jl:48-63)."
theta_to_params(θ::AbstractVector) = vcat(exp(θ[end]), exp.(θ[2:end-1]))

"""
    gather_results(local_rows::Dict{Int,Vector{Float64}}, ntrials, width) -> Matrix (width × ntrials)

The final gather of per-trial results across ranks (replaces the lock-guarded result callbacks of core.jl:47-56 when the
trials are sharded over GPUs): one ncclAllGather inside libgprb200.so.  Ranks join with `comm_init_rank!` (the 128-byte id
from `comm_unique_id()` on rank 0 travels through Distributed.jl / MPI / a file).  Rows are `kstep_mse`, `projectionerror`,
`params` ... in whatever layout `resultcallback!` needs; trial ids are 1-based here, 0-based in the C ABI.
"""
function gather_results(local_rows::Dict{Int,Vector{Float64}}, ntrials::Integer, width::Integer)
    ids = Int32[t - 1 for t in sort(collect(keys(local_rows)))]
    rows = isempty(ids) ? zeros(width, 0) : reduce(hcat, [local_rows[t + 1] for t in ids])     # width × count = row-major count × width
    out = Matrix{Float64}(undef, width, ntrials)
    check(ccall((:gprb_gather, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                context(), ntrials, width, length(ids), ids, rows, out))
    out
end
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:gprb_comm_unique_id, LIB), Cint, (Ptr{Cvoid},), id))
    id
end
comm_init_rank!(nranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:gprb_comm_init_rank, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}), context(), nranks, rank, id))

end # module
