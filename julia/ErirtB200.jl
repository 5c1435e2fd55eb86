# ErirtB200.jl -- reference-side binding of liberirt_b200.so (include/erirt_b200.h).
#
# This file is what a maintainer of ExtendedRtIrtModeling.jl adds to the package (INTEGRATION.md): it keeps the
# package's API (setCond, InputData, the Gibbs* constructors, sample!, MCMC.Post.mean, coef, precis) and replaces
# only the BODY of the `sample!` methods (src/GibbsRtIrt.pl.jl:210-257, :278-346, :367-426,
# src/GibbsRtIrtLatent.pl.jl:168-233, :271-337) by ccalls.  Julia is not installed in the build container, so this
# shim is delivered as source and has not been executed; the same C ABI is exercised by the Python ctypes mirror
# (extendedrtirtmodeling.jl_b200/engine.py) in the test-suite.
module ErirtB200

using ExtendedRtIrtModeling
import ExtendedRtIrtModeling: sample!, InputPara, GibbsMlIrt, GibbsRtIrt, GibbsRtIrtNull, GibbsRtIrtCross, GibbsRtIrtCrossQr,
                               GibbsRtIrtLatent, GibbsRtIrtLatentQr

export GibbsRtIrtQuantile

const LIB = get(ENV, "ERIRT_B200_LIB", joinpath(@__DIR__, "..", "extendedrtirtmodeling.jl_b200", "liberirt_b200.so"))

# struct erirt_config (include/erirt_b200.h) -- field order and C alignment must match
struct ErirtConfig
    abi_version::Int32
    model::Int32
    n_subj::Int64
    n_subj_total::Int64
    subj_offset::Int64
    n_item::Int32
    n_feat::Int32
    n_iter::Int32
    n_chain::Int32
    n_burnin::Int32
    q_rt::Float64
    intercept::Int32
    itemtype_1pl::Int32
    cov2one::Int32
    dtype::Int32
    seed::UInt64
    chain::UInt32
    compat::Int32
    person_trace::Int32
    device::Int32
    use_graph::Int32
    time_kernels::Int32
    nu_cell_moments::Int32
    reserved::NTuple{6,Int32}
end

const MODEL_ID = Dict(GibbsMlIrt => 0, GibbsRtIrt => 1, GibbsRtIrtNull => 2, GibbsRtIrtCross => 3, GibbsRtIrtCrossQr => 4,
                      GibbsRtIrtLatent => 5, GibbsRtIrtLatentQr => 6)
const F_THETA, F_ZETA, F_A, F_B, F_LAMBDA, F_SIGMA2, F_BETA, F_RHO, F_SIGMA_P, F_NU = Int32.(0:9)
const T_RA, T_RT, T_QR, T_LL = Int32.(0:3)

lasterror() = unsafe_string(ccall((:erirt_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("erirt_b200: " * lasterror())

"README.md:95 names GibbsRtIrtQuantile; the package ships that sampler as GibbsRtIrtLatentQr (export commented out at src/ExtendedRtIrtModeling.jl:65)."
GibbsRtIrtQuantile(Cond; kwargs...) = GibbsRtIrtLatentQr(Cond; kwargs...)

function _set_state(h, field, v)
    isempty(v) && return
    x = Vector{Float64}(vec(v))
    GC.@preserve x check(ccall((:erirt_set_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), h, field, x, length(x)))
end

function _trace!(h, which, out::Array{Float64,3}, first_col)
    GC.@preserve out check(ccall((:erirt_get_trace, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Float64}),
                                 h, which, first_col, size(out, 2), out))
end

function _sample_gpu!(MCMC; intercept=false, itemtype="2pl", cov2one=true, dtype=1, seed=rand(UInt64), device=0)
    if !(itemtype in ["1pl", "2pl"])
        error("Invalid input: the item type must be '1pl' or '2pl'.")     # src/GibbsRtIrt.pl.jl:212-214
    end
    Cond, Data, Para, Post = MCMC.Cond, MCMC.Data, MCMC.Para, MCMC.Post
    N, J, F = Cond.nSubj, Cond.nItem, Cond.nFeat
    has_rt = !(MCMC isa GibbsMlIrt)
    cqr = MCMC isa GibbsRtIrtCrossQr
    cfg = ErirtConfig(1, MODEL_ID[typeof(MCMC)], N, N, 0, J, F, Cond.nIter, Cond.nChain, Cond.nBurnin, Cond.qRt,
                      intercept, itemtype == "1pl", cov2one, dtype, seed, 0, 0, 1, device, 1, 0, cqr, ntuple(_ -> Int32(0), 6))
    νmean = Float64[]   # CrossQr: post-burn-in mean of the N*J weights
    href = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:erirt_create, LIB), Cint, (Ref{ErirtConfig}, Ref{Ptr{Cvoid}}), cfg, href))
    h = href[]
    try
        logT = has_rt ? Matrix{Float64}(Data.logT) : zeros(0, 0)
        X = F > 0 ? Matrix{Float64}(Data.X) : zeros(0, 0)
        if Data.Y isa Matrix{Bool}   # rand.(BernoulliLogit…) of src/SimTools.jl:165: one byte per response, passed as it lies in memory
            Y8 = Data.Y
            GC.@preserve Y8 logT X check(ccall((:erirt_set_data_y8, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Bool}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
                h, Y8, N, has_rt ? pointer(logT) : C_NULL, N, F > 0 ? pointer(X) : C_NULL, N))
        else                         # Int from CSV (src/Base.pl.jl:84-95) or Float64
            Y = Matrix{Float64}(Data.Y)
            GC.@preserve Y logT X check(ccall((:erirt_set_data, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
                h, Y, N, has_rt ? pointer(logT) : C_NULL, N, F > 0 ? pointer(X) : C_NULL, N))
        end
        _set_state(h, F_THETA, Para.θ); _set_state(h, F_A, Para.a); _set_state(h, F_B, Para.b)
        if has_rt
            _set_state(h, F_ZETA, Para.ζ); _set_state(h, F_LAMBDA, Para.λ); _set_state(h, F_SIGMA2, Para.σ²t)
            _set_state(h, F_SIGMA_P, Matrix{Float64}(Para.Σp))
        end
        _set_state(h, F_BETA, Para.β)
        (MCMC isa GibbsRtIrtCross || cqr) && _set_state(h, F_RHO, Para.ρ)
        check(ccall((:erirt_sample, LIB), Cint, (Ptr{Cvoid}, Int64), h, Cond.nIter * Cond.nChain))
        # fill the pre-allocated Post arrays (same layouts, src/GibbsRtIrt.pl.jl:63-69)
        _trace!(h, T_RA, Post.ra, 0)
        has_rt && _trace!(h, T_RT, Post.rt, 0)
        if cqr
            # the reference traces all N*J weights per sweep (src/GibbsRtIrtCross.pl.jl:65, :297); the engine keeps their running
            # mean instead, so only the [ρ; vec Σp] columns of Post.qr are filled and Post.mean.ν comes from erirt_get_moments
            small = Array{Float64}(undef, Cond.nIter, J + 4, Cond.nChain)
            _trace!(h, T_QR, small, 0)
            Post.qr[:, 1:(J + 4), :] .= small
            νmean = Vector{Float64}(undef, N * J)
            GC.@preserve νmean check(ccall((:erirt_get_moments, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64),
                                           h, F_NU, νmean, C_NULL, N * J))
        else
            _trace!(h, T_QR, Post.qr, 0)
        end
        _trace!(h, T_LL, Post.logLike, 0)
    finally
        ccall((:erirt_destroy, LIB), Cint, (Ptr{Cvoid},), h)
    end
    return νmean
end

_pm(A, r, nb) = vec(ExtendedRtIrtModeling.mean(A[(nb + 1):end, r, :], dims=(1, 3)))

# ---- the sample! methods: same signatures as the reference, body replaced ----
function sample!(MCMC::GibbsMlIrt; intercept=false, itemtype::Union{String}="2pl")
    _sample_gpu!(MCMC; intercept, itemtype)
    C, P = MCMC.Cond, MCMC.Post
    P.mean = InputPara(θ=_pm(P.ra, 1:C.nSubj, C.nBurnin), a=_pm(P.ra, (C.nSubj + 1):(C.nSubj + C.nItem), C.nBurnin),
                       b=_pm(P.ra, (C.nSubj + C.nItem + 1):size(P.ra, 2), C.nBurnin), β=_pm(P.qr, 1:(C.nFeat + 1), C.nBurnin))
    return MCMC
end

function _rt_mean!(MCMC, nβ; with_ν=false)
    C, P = MCMC.Cond, MCMC.Post
    nb, N, J = C.nBurnin, C.nSubj, C.nItem
    P.mean = InputPara(β=_pm(P.qr, 1:nβ, nb), Σp=_pm(P.qr, (nβ + 1):(nβ + 4), nb),
                       ν=with_ν ? _pm(P.qr, (nβ + 5):size(P.qr, 2), nb) : Float64[],
                       θ=_pm(P.ra, 1:N, nb), a=_pm(P.ra, (N + 1):(N + J), nb), b=_pm(P.ra, (N + J + 1):(N + 2J), nb),
                       ζ=_pm(P.rt, 1:N, nb), λ=_pm(P.rt, (N + 1):(N + J), nb), σ²t=_pm(P.rt, (N + J + 1):(N + 2J), nb))
    return MCMC
end

function sample!(MCMC::GibbsRtIrt; intercept=false, itemtype::Union{String}="2pl", cov2one=true)
    _sample_gpu!(MCMC; intercept, itemtype, cov2one); _rt_mean!(MCMC, 2 * (MCMC.Cond.nFeat + 1))
end
function sample!(MCMC::GibbsRtIrtNull; itemtype::Union{String}="2pl", cov2one=true)
    _sample_gpu!(MCMC; itemtype, cov2one); _rt_mean!(MCMC, 2 * (MCMC.Cond.nFeat + 1))
end
function _cross_mean!(MCMC; ν=Float64[])
    C, P = MCMC.Cond, MCMC.Post
    nb, N, J = C.nBurnin, C.nSubj, C.nItem
    P.mean = InputPara(ρ=_pm(P.qr, 1:J, nb), Σp=_pm(P.qr, (J + 1):(J + 4), nb), ν=ν,
                       θ=_pm(P.ra, 1:N, nb), a=_pm(P.ra, (N + 1):(N + J), nb), b=_pm(P.ra, (N + J + 1):(N + 2J), nb),
                       ζ=_pm(P.rt, 1:N, nb), λ=_pm(P.rt, (N + 1):(N + J), nb), σ²t=_pm(P.rt, (N + J + 1):(N + 2J), nb))
    return MCMC
end
function sample!(MCMC::GibbsRtIrtCross; itemtype::Union{String}="2pl", cov2one=true)      # src/GibbsRtIrtCross.pl.jl:176
    _sample_gpu!(MCMC; itemtype, cov2one); _cross_mean!(MCMC)
end
function sample!(MCMC::GibbsRtIrtCrossQr; itemtype::Union{String}="2pl", cov2one=true)    # src/GibbsRtIrtCross.pl.jl:265
    ν = _sample_gpu!(MCMC; itemtype, cov2one); _cross_mean!(MCMC; ν=ν)                     # Post.mean.ν, :310
end
function sample!(MCMC::GibbsRtIrtLatent; intercept=false, itemtype::Union{String}="2pl", cov2one=false)
    _sample_gpu!(MCMC; intercept, itemtype, cov2one); _rt_mean!(MCMC, MCMC.Cond.nFeat + 2)
end
function sample!(MCMC::GibbsRtIrtLatentQr; intercept=false, itemtype::Union{String}="2pl", cov2one=false)
    _sample_gpu!(MCMC; intercept, itemtype, cov2one); _rt_mean!(MCMC, MCMC.Cond.nFeat + 2; with_ν=true)
end

# ---- simulated data generated on the device (erirt_generate_data): the N x J part of setData* (src/SimTools.jl:117-368) never exists on
#      the host.  `truePara` holds the person-level draws θ (and ζ) made by the caller as in setData*, X the covariates (or nothing);
#      errortype: 0 truncated normal (Null / RtIrt), 1 N(0,1) (Latent*), 2 / 3 / 4 = "norm" / "tail" / "skew" (Cross) ----
function generate_data!(h, truePara, X, errortype::Integer; seed::UInt64=rand(UInt64))
    ptr(v) = isempty(v) ? Ptr{Float64}(C_NULL) : pointer(v)
    θ, ζ = Vector{Float64}(vec(truePara.θ)), Vector{Float64}(vec(truePara.ζ))
    a, b, λ = Vector{Float64}(vec(truePara.a)), Vector{Float64}(vec(truePara.b)), Vector{Float64}(vec(truePara.λ))
    σ², ρ = Vector{Float64}(vec(truePara.σ²t)), Vector{Float64}(vec(truePara.ρ))
    Xm = X === nothing ? zeros(0, 0) : Matrix{Float64}(X)
    GC.@preserve θ ζ a b λ σ² ρ Xm check(ccall((:erirt_generate_data, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, UInt64),
        h, θ, ptr(ζ), a, b, ptr(λ), ptr(σ²), ptr(ρ), ptr(Xm), max(size(Xm, 1), 1), errortype, seed))
end

# ---- checkpoint / resume of a running chain (erirt_checkpoint_*): `h` is the handle of a chain driven in chunks ----
function checkpoint(h)::Vector{UInt8}
    n = ccall((:erirt_checkpoint_size, LIB), Int64, (Ptr{Cvoid},), h)
    n < 0 && error("erirt_b200: " * lasterror())
    buf = Vector{UInt8}(undef, n)
    GC.@preserve buf check(ccall((:erirt_checkpoint_save, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), h, buf, n))
    return buf      # write(io, buf) to persist it
end
restore!(h, buf::Vector{UInt8}) =
    GC.@preserve buf check(ccall((:erirt_checkpoint_load, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), h, buf, length(buf)))

end # module
