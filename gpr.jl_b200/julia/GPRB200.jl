# GPRB200.jl - thin ccall shim that keeps the GP surface GPR.jl's experiments call and routes it to libgprb200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  The tested contract is the C ABI
# (include/gprb200.h, exercised through ctypes by tests/); every ccall below mirrors one prototype of that header
# one-to-one, and gpr.jl_b200/lib.py is the executable twin of this file.
#
# Drop-in use inside the reference (paths relative to the GPR.jl checkout):
#
#     # examples/maximal_coordinates/CPnoise.jl:35-43, unchanged except for the module prefix
#     using GPRB200                       # instead of `using GaussianProcesses` for the four calls below
#     kernel = SEArd(log.(params[2:end]), log(params[1]))
#     gp = GP(xtrain_old, yi, MeanZero(), kernel)                     # or MeanDynamics(...) from src/mDynamics.jl
#     GPRB200.optimize!(gp, LBFGS(linesearch = BackTracking(order=2)), Optim.Options(time_limit=10.))
#     μ = predict_y(gp, obs)[1][1]                                    # examples/utils/predictdynamics.jl:13
#
# and, batched (what replaces the `Threads.@threads for jobid` loop of examples/parallel/core.jl:28):
#
#     gps = [GPE(X_t, y_tk, mean_tk, SEArd(...)) for t in trials for k in outputs]     # no evaluation yet
#     batch = GPBatch(gps)                                            # uploads each distinct X once
#     GPRB200.optimize!(batch, LBFGS(linesearch = BackTracking(order=2)), Optim.Options(iterations=1000))
#     μ, σ² = predict_y(batch, Xstar)                                 # B × m
module GPRB200

using Libdl
import GaussianProcesses                       # only for the Mean plug-in protocol and the kernel parameter types
import GaussianProcesses: Mean, MeanZero, SEArd, Mat12Ard, Mat32Ard, Mat52Ard, get_params, set_params!, num_params
import Optim

export GPE, GP, GPBatch, optimize!, predict_y, update_mll!, update_mll_and_dmll!, SEArd, MeanZero

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

# ---- GPE: the fields the reference's callers read (gp.x, gp.y, gp.mean, gp.kernel, gp.logNoise, gp.mll, gp.dmll, gp.alpha)
mutable struct GPE
    x::Matrix{Float64}          # d × n, one CState per column (src/CState.jl:20)
    y::Vector{Float64}
    mean::Mean
    kernel
    logNoise::Float64
    dim::Int
    nobs::Int
    mll::Float64
    dmll::Vector{Float64}
    info::Int32
    batch::Any                  # owning GPBatch
    slot::Int
end
GPE(X::AbstractMatrix, y::AbstractVector, mean::Mean, kernel, logNoise::Real = -2.0) =
    GPE(Matrix{Float64}(X), Vector{Float64}(y), mean, kernel, Float64(logNoise), size(X, 1), size(X, 2), NaN, Float64[], 0, nothing, 0)

# GaussianProcesses.get_params order: [logNoise; mean params; kernel params] - the reference's means have none (src/mDynamics.jl:29-31)
params(gp::GPE) = vcat(gp.logNoise, get_params(gp.mean), get_params(gp.kernel))
function setparams!(gp::GPE, θ::AbstractVector)
    gp.logNoise = θ[1]
    nm = num_params(gp.mean)
    nm > 0 && set_params!(gp.mean, θ[2:1+nm])
    set_params!(gp.kernel, θ[2+nm:end])
end

# ---- GPBatch ---------------------------------------------------------------------------------------------------------
mutable struct GPBatch
    gps::Vector{GPE}
    handle::Ptr{Cvoid}
    datasets::Vector{Ptr{Cvoid}}
    B::Int; n::Int; d::Int; P::Int
end

function GPBatch(gps::Vector{GPE})
    ctx = context()
    d, n = gps[1].dim, gps[1].nobs
    B = length(gps)
    seen = IdDict{Any,Ptr{Cvoid}}()
    handles = Vector{Ptr{Cvoid}}(undef, B)
    for (b, gp) in enumerate(gps)
        handles[b] = get!(seen, gp.x) do                 # GPs of one trial share the same X object => one upload
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:gprb_dataset_create, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                        ctx, n, d, gp.x, d, h))
            h[]
        end
    end
    # m(X) does not depend on θ: evaluate once per training set, column by column across the GPs of a trial so the shared
    # MDCache (src/mDynamics.jl:6-11,42) hits for the other G-1 outputs.  Only y - m(X) goes to the device.
    ymm = Matrix{Float64}(undef, n, B)
    for xs in unique(objectid(gp.x) for gp in gps)
        members = [b for b in 1:B if objectid(gps[b].x) == xs]
        for j in 1:n, b in members
            ymm[j, b] = gps[b].y[j] - GaussianProcesses.mean(gps[b].mean, gps[b].x[:, j])
        end
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gprb_batch_create, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
                ctx, B, handles, ymm, kernel_kind(gps[1].kernel), h))
    batch = GPBatch(gps, h[], collect(values(seen)), B, n, d, d + 2)
    for (b, gp) in enumerate(gps)
        gp.batch, gp.slot = batch, b
    end
    finalizer(batch) do bt
        ccall((:gprb_batch_destroy, LIB), Cint, (Ptr{Cvoid},), bt.handle)
        foreach(ds -> ccall((:gprb_dataset_destroy, LIB), Cint, (Ptr{Cvoid},), ds), bt.datasets)
    end
    batch
end

thetas(batch::GPBatch) = reduce(hcat, params.(batch.gps))          # P × B, column per GP: the layout gprb_eval takes

"One objective evaluation per GP: update_mll! (grad=false) / update_mll_and_dmll! (grad=true)."
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
        gp.mll, gp.info = mll[b], info[b]
        grad && (gp.dmll = g[:, b])
    end
    mll, g, info
end
update_mll!(gp::GPE) = (evaluate!(gp.batch; grad = false); gp)
update_mll_and_dmll!(gp::GPE) = (evaluate!(gp.batch; grad = true); gp)

"GP(X, y, mean, kernel): construct + initial update_mll! (GaussianProcesses.GP), a batch of one."
function GP(X::AbstractMatrix, y::AbstractVector, mean::Mean, kernel, logNoise::Real = -2.0)
    gp = GPE(X, y, mean, kernel, logNoise)
    evaluate!(GPBatch([gp]); grad = false)
    gp
end

# ---- optimize! -------------------------------------------------------------------------------------------------------
struct LbfgsOpts                 # gprb_lbfgs_opts
    m::Int32; iterations::Int32; max_evals::Int32; ls_iterations::Int32
    g_abstol::Float64; time_limit::Float64; c_1::Float64; rho_hi::Float64; rho_lo::Float64
end
struct OptResult                 # gprb_opt_result
    mll::Float64; g_norm::Float64
    iterations::Int32; f_calls::Int32; fg_calls::Int32; converged::Int32; ls_failed::Int32; info::Int32
end

"""
    optimize!(gp_or_batch, method::Optim.LBFGS, options::Optim.Options; max_evals = 0)

Same positional signature as the reference's `GaussianProcesses.optimize!(gp, LBFGS(linesearch=BackTracking(order=2)),
Optim.Options(time_limit=10.))` (CPnoise.jl:41).  `options.time_limit` is wall-clock for the whole batch; `max_evals`
is the deterministic stopping rule used for parity runs.
"""
function optimize!(batch::GPBatch, method::Optim.LBFGS = Optim.LBFGS(), options::Optim.Options = Optim.Options(); max_evals::Integer = 0)
    ls = method.linesearch!
    o = LbfgsOpts(method.m, options.iterations, max_evals, ls.iterations, options.g_abstol,
                  isfinite(options.time_limit) ? options.time_limit : 0.0, ls.c_1, ls.ρ_hi, ls.ρ_lo)
    θ = thetas(batch)
    res = Vector{OptResult}(undef, batch.B)
    check(ccall((:gprb_optimize, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{LbfgsOpts}, Ptr{OptResult}), batch.handle, θ, o, res))
    for (b, gp) in enumerate(batch.gps)
        setparams!(gp, θ[:, b])
        gp.mll, gp.info = res[b].mll, res[b].info
    end
    res
end
optimize!(gp::GPE, args...; kw...) = optimize!(gp.batch !== nothing && gp.batch.B == 1 ? gp.batch : GPBatch([gp]), args...; kw...)[1]
optimize!(gps::Vector{GPE}, args...; kw...) = optimize!(GPBatch(gps), args...; kw...)

# ---- predict_y -------------------------------------------------------------------------------------------------------
"predict_y(batch, Xstar::d×m) -> (μ::B×m... stored m×B, σ²) ; `var=false` skips the variance the reference discards."
function predict_y(batch::GPBatch, Xstar::AbstractMatrix; var::Bool = true)
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
function predict_y(gp::GPE, Xstar::AbstractMatrix)                 # reference call pattern: (μ::Vector, σ²::Vector)
    μ, σ2 = predict_y(gp.batch, Xstar)
    μ[:, gp.slot], σ2[:, gp.slot]
end

end # module
