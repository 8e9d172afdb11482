# LowRankModelsB200.jl — the Julia side of the drop-in boundary.
#
# NOT EXECUTED IN THE BUILD CONTAINER (no julia binary, no network): this file is the binding a
# LowRankModels.jl maintainer adds.  It is kept deliberately thin and is mirrored 1:1 by the Python
# ctypes host (lowrankmodels.jl_b200/{encode,fit,_abi}.py), which IS executed by the test-suite against
# the same C ABI (include/glrm_b200.h).
#
# How it plugs in: exactly like SparseProxGradParams (src/algorithms/sparse_proxgrad.jl:4-24) — a new
# `T <: AbstractParams` (src/fit.jl:4) plus a method
#     fit!(glrm::GLRM, params::T; ch, verbose, kwargs...) -> (glrm.X, glrm.Y, ch)
# selected with `fit!(glrm, B200ProxGradParams())` or `fit!(glrm; params=B200ProxGradParams())`
# (src/fit.jl:9-11).  Everything else in the package (GLRM constructor, losses, regularizers,
# cross_validate, impute, ...) is untouched.
#
#     include("LowRankModelsB200.jl"); using .LowRankModelsB200
#     X, Y, ch = fit!(glrm, B200ProxGradParams(1.0; max_iter=100))
module LowRankModelsB200

using LowRankModels
import LowRankModels: fit!, AbstractParams, GLRM, ConvergenceHistory, update_ch!, get_yidxs, embedding_dim,
                      Loss, Regularizer,
                      QuadLoss, L1Loss, HuberLoss, QuantileLoss, PeriodicLoss, PoissonLoss, OrdinalHingeLoss,
                      LogisticLoss, WeightedHingeLoss, MultinomialLoss, OvALoss, BvSLoss, OrdisticLoss,
                      MultinomialOrdinalLoss,
                      ZeroReg, QuadReg, QuadConstraint, OneReg, NonNegConstraint, NonNegOneReg,
                      OneSparseConstraint, KSparseConstraint, UnitOneSparseConstraint, SimplexConstraint,
                      lastentry1, lastentry_unpenalized, OrdinalReg, MNLOrdinalReg,
                      fixed_latent_features, fixed_last_latent_features, RemQuadReg

export B200ProxGradParams, B200SparseProxGradParams, B200Handle, set_obs!, set_reg_scale!, objective_resident, error_metric_resident, impute_resident

const LIB = get(ENV, "GLRMB200_LIB", joinpath(@__DIR__, "..", "csrc", "libglrm_b200.so"))

# ---- ProxGradParams' seven fields (src/algorithms/proxgrad.jl:4-31) + where to run ----------------------
# device: CUDA ordinal of this process; (rank, nranks): this process' shard when the fit is spread over several
# processes, one per GPU (Distributed.jl workers; the NCCL id / IPC blobs travel through the host, INTEGRATION.md)
mutable struct B200ProxGradParams <: AbstractParams
    stepsize::Float64
    max_iter::Int
    inner_iter_X::Int
    inner_iter_Y::Int
    abs_tol::Float64
    rel_tol::Float64
    min_stepsize::Float64
    device::Int
    rank::Int
    nranks::Int
end
function B200ProxGradParams(stepsize::Number=1.0; max_iter::Int=100, inner_iter_X::Int=1, inner_iter_Y::Int=1,
                            inner_iter::Int=1, abs_tol::Number=0.00001, rel_tol::Number=0.0001,
                            min_stepsize::Number=0.01*stepsize, device::Int=0, rank::Int=0, nranks::Int=1)
    B200ProxGradParams(Float64(stepsize), max_iter, max(inner_iter_X, inner_iter), max(inner_iter_Y, inner_iter),
                       Float64(abs_tol), Float64(rel_tol), Float64(min_stepsize), device, rank, nranks)
end

# ---- C structs (include/glrm_b200.h) ------------------------------------------------------------------
struct CParams            # glrmb200_params
    stepsize::Cdouble
    max_iter::Int32
    inner_iter_X::Int32
    inner_iter_Y::Int32
    abs_tol::Cdouble
    rel_tol::Cdouble
    min_stepsize::Cdouble
end

struct CProblem           # glrmb200_problem
    m::Int64; n::Int64; k::Int64; d::Int64
    loss_code::Ptr{Int32}; loss_param::Ptr{Cdouble}
    rx_count::Int64; rx_code::Ptr{Int32}; rx_param::Ptr{Cdouble}
    ry_count::Int64; ry_code::Ptr{Int32}; ry_param::Ptr{Cdouble}
    obs_full::Int32; dense_A::Ptr{Cdouble}
    row_ptr::Ptr{Int64}; row_idx::Ptr{Int32}; row_val::Ptr{Cdouble}
    col_ptr::Ptr{Int64}; col_idx::Ptr{Int32}; col_val::Ptr{Cdouble}
    rx_payload_ptr::Ptr{Int64}; rx_payload::Ptr{Cdouble}
    ry_payload_ptr::Ptr{Int64}; ry_payload::Ptr{Cdouble}
end

const NLOSSP = 8
const NREGP = 4

# ---- descriptor tables: a type without a device implementation is an error (no CPU fallback) ---------
lossrow(l::QuadLoss)          = (1, (l.scale,))
lossrow(l::L1Loss)            = (2, (l.scale,))
lossrow(l::HuberLoss)         = (3, (l.scale, l.crossover))
lossrow(l::QuantileLoss)      = (4, (l.scale, l.quantile))
lossrow(l::PeriodicLoss)      = (5, (l.scale, l.T))
lossrow(l::PoissonLoss)       = (6, (l.scale,))
lossrow(l::OrdinalHingeLoss)  = (7, (l.scale, Float64(l.min), Float64(l.max)))
lossrow(l::LogisticLoss)      = (8, (l.scale,))
lossrow(l::WeightedHingeLoss) = (9, (l.scale, l.case_weight_ratio))
lossrow(l::MultinomialLoss)   = (10, (l.scale, 0.0, Float64(l.max)))
function lossrow(l::Union{OvALoss,BvSLoss})
    bc, bp = lossrow(l.bin_loss)
    (l isa OvALoss ? 11 : 12, (l.scale, 0.0, Float64(l.max), Float64(bc), bp[1], length(bp) > 1 ? bp[2] : 0.0,
                                 length(bp) > 2 ? bp[3] : 0.0))
end
lossrow(l::OrdisticLoss)           = (13, (l.scale, 0.0, Float64(l.max)))
lossrow(l::MultinomialOrdinalLoss) = (14, (l.scale, 0.0, Float64(l.max)))
lossrow(l::Loss) = throw(ArgumentError("$(typeof(l)) has no B200 device implementation (no CPU fallback)"))

# (code, first parameter, vector payload or nothing)
const NOPAY = nothing
regrow(r::ZeroReg)                 = (0, 0.0, NOPAY)
regrow(r::QuadReg)                 = (1, r.scale, NOPAY)
regrow(r::QuadConstraint)          = (2, r.max_2norm, NOPAY)
regrow(r::OneReg)                  = (3, r.scale, NOPAY)
regrow(r::NonNegConstraint)        = (4, 0.0, NOPAY)
regrow(r::NonNegOneReg)            = (5, r.scale, NOPAY)
regrow(r::OneSparseConstraint)     = (6, 0.0, NOPAY)
regrow(r::KSparseConstraint)       = (7, Float64(r.k), NOPAY)
regrow(r::UnitOneSparseConstraint) = (8, 0.0, NOPAY)
regrow(r::SimplexConstraint)       = (9, 0.0, NOPAY)
regrow(r::RemQuadReg)              = (10, r.scale, Vector{Float64}(r.m))        # regularizers.jl:412-423
wrapped(r, flag) = begin
    c, p, pay = regrow(r.r)
    (pay === NOPAY && c < 0x100) || throw(ArgumentError("nested regularizer wrappers have no B200 device implementation"))
    (c | flag, p, NOPAY)
end
regrow(r::lastentry1)              = wrapped(r, 0x100)
regrow(r::lastentry_unpenalized)   = wrapped(r, 0x200)
regrow(r::OrdinalReg)              = wrapped(r, 0x400)                           # block regularizers (ry only)
regrow(r::MNLOrdinalReg)           = wrapped(r, 0x800)
regrow(r::fixed_latent_features)      = ((c, p, _) = wrapped(r, 0x1000); (c, p, Vector{Float64}(r.y)))   # regularizers.jl:193-210
regrow(r::fixed_last_latent_features) = ((c, p, _) = wrapped(r, 0x2000); (c, p, Vector{Float64}(r.y)))   # regularizers.jl:214-231
regrow(r::Regularizer) = throw(ArgumentError("$(typeof(r)) has no B200 device implementation (no CPU fallback)"))

# -> codes, params, payload_ptr (0-based offsets, length count+1; empty when nobody carries a payload), payload
function regtable(rs)
    rows = map(regrow, rs)
    haspay = any(r -> r[3] !== NOPAY, rows)
    if !haspay && all(==(rows[1]), rows)          # the usual fillcopies case: one shared row
        rows = rows[1:1]
    end
    codes = Int32[c for (c, _, _) in rows]
    params = zeros(Cdouble, NREGP, length(rows))
    for (i, (_, p, _)) in enumerate(rows); params[1, i] = p; end
    pptr = Int64[]; pval = Cdouble[]
    if haspay
        pptr = zeros(Int64, length(rows) + 1)
        for (i, (_, _, pay)) in enumerate(rows)
            pay !== NOPAY && append!(pval, pay)
            pptr[i+1] = length(pval)
        end
        isempty(pval) && push!(pval, 0.0)
    end
    codes, params, pptr, pval
end

# label as the engine expects it: Bool -> 0/1, numbers as Float64.  Boolean / categorical domain errors are
# raised by the library (GLRMB200_E_LABEL) — same cases in which the reference throws (losses.jl:104).
labelval(a::Bool) = a ? 1.0 : 0.0
labelval(a::Number) = Float64(a)

function flatten_obs(lists::AbstractVector, A, byrow::Bool)
    nl = length(lists)
    ptr = Vector{Int64}(undef, nl + 1); ptr[1] = 0
    for i in 1:nl; ptr[i+1] = ptr[i] + length(lists[i]); end
    idx = Vector{Int32}(undef, ptr[end]); val = Vector{Cdouble}(undef, ptr[end])
    q = 0
    for i in 1:nl, j in lists[i]        # list order and duplicates preserved (modify_glrm.jl:5-18)
        q += 1
        idx[q] = Int32(j - 1)           # 0-based across the ABI
        val[q] = byrow ? labelval(A[i, j]) : labelval(A[j, i])
    end
    ptr, idx, val
end

lasterr() = unsafe_string(ccall((:glrmb200_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("glrmb200 error $rc: $(lasterr())")

# ---- the encoded problem: every array the C struct points into, kept alive together -------------------------------
struct Encoded
    m::Int; n::Int; k::Int; d::Int
    lcode::Vector{Int32}; lparam::Matrix{Cdouble}
    rxc::Vector{Int32}; rxp::Matrix{Cdouble}; rxpp::Vector{Int64}; rxpv::Vector{Cdouble}
    ryc::Vector{Int32}; ryp::Matrix{Cdouble}; rypp::Vector{Int64}; rypv::Vector{Cdouble}
    full::Bool; dense::Array{Cdouble}
    rptr::Vector{Int64}; ridx::Vector{Int32}; rval::Vector{Cdouble}
    cptr::Vector{Int64}; cidx::Vector{Int32}; cval::Vector{Cdouble}
end

function encode(glrm::GLRM)
    A = glrm.A
    m, n = size(A)
    k = glrm.k
    yidxs = get_yidxs(glrm.losses)
    d = maximum(yidxs[end])
    size(glrm.Y) == (k, d) || error("size(glrm.Y) must be (k, embedding_dim(losses)) = ($k, $d)")   # proxgrad.jl:55-63
    lrows = map(lossrow, glrm.losses)
    lcode = Int32[c for (c, _) in lrows]
    lparam = zeros(Cdouble, NLOSSP, n)
    for (f, (_, p)) in enumerate(lrows), (i, v) in enumerate(p); lparam[i, f] = v; end
    rxc, rxp, rxpp, rxpv = regtable(glrm.rx)
    ryc, ryp, rypp, rypv = regtable(glrm.ry)
    full = all(o -> o == 1:n, glrm.observed_features) && all(o -> o == 1:m, glrm.observed_examples)
    if full
        dense = Matrix{Cdouble}(map(labelval, A))
        rptr = Int64[]; ridx = Int32[]; rval = Cdouble[]; cptr = Int64[]; cidx = Int32[]; cval = Cdouble[]
    else
        dense = Cdouble[]
        rptr, ridx, rval = flatten_obs(glrm.observed_features, A, true)
        cptr, cidx, cval = flatten_obs(glrm.observed_examples, A, false)
    end
    Encoded(m, n, k, d, lcode, lparam, rxc, rxp, rxpp, rxpv, ryc, ryp, rypp, rypv, full, dense, rptr, ridx, rval, cptr, cidx, cval)
end

ptr_or_null(v) = isempty(v) ? C_NULL : pointer(v)
# call f(Ref{CProblem}) with every array of `e` rooted
function with_problem(f::Function, e::Encoded)
    GC.@preserve e begin
        prob = Ref(CProblem(e.m, e.n, e.k, e.d, pointer(e.lcode), pointer(e.lparam),
                            length(e.rxc), pointer(e.rxc), pointer(e.rxp), length(e.ryc), pointer(e.ryc), pointer(e.ryp),
                            e.full ? 1 : 0, ptr_or_null(e.dense),
                            ptr_or_null(e.rptr), ptr_or_null(e.ridx), ptr_or_null(e.rval),
                            ptr_or_null(e.cptr), ptr_or_null(e.cidx), ptr_or_null(e.cval),
                            ptr_or_null(e.rxpp), ptr_or_null(e.rxpv), ptr_or_null(e.rypp), ptr_or_null(e.rypv)))
        f(prob)
    end
end

# ---- a live engine handle: the problem stays on the device between fits -------------------------------------------
# What cross_validate / cv_by_iter / regularization_path need (src/cross_validate.jl:31-43,141-240): the same A, swapped
# observation lists per fold (set_obs!), rescaled regularizers along a path (set_reg_scale!), warm starts from glrm.X/Y.
mutable struct B200Handle
    h::Ptr{Cvoid}
    device::Int; rank::Int; nranks::Int
end
const CREATE_GATHER_ONLY = Int32(1)     # GLRMB200_CREATE_GATHER_ONLY: what glrmb200_fit_sparse needs for a fully observed A
function B200Handle(glrm::GLRM; device::Int=0, rank::Int=0, nranks::Int=1, gather_only::Bool=false)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    with_problem(encode(glrm)) do prob
        check(ccall((:glrmb200_create_ex, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{CProblem}, Int32, Int32, Int32, Int32),
                    handle, prob, device, rank, nranks, gather_only ? CREATE_GATHER_ONLY : Int32(0)))
    end
    hd = B200Handle(handle[], device, rank, nranks)
    finalizer(close, hd)
    hd
end
function Base.close(hd::B200Handle)
    hd.h == C_NULL && return
    ccall((:glrmb200_destroy, LIB), Cint, (Ptr{Cvoid},), hd.h)
    hd.h = C_NULL
    nothing
end
# the observation lists of `glrm` replace the handle's (a train / test fold of the same A): cross_validate.jl:31-33
set_obs!(hd::B200Handle, glrm::GLRM) = with_problem(encode(glrm)) do prob
    check(ccall((:glrmb200_set_obs, LIB), Cint, (Ptr{Cvoid}, Ref{CProblem}), hd.h, prob))
end
# scale_regularizer!(glrm, s) on the device tables (glrm.jl:84-88): call both, the host model stays the source of truth
set_reg_scale!(hd::B200Handle, s::Number) = check(ccall((:glrmb200_set_reg_scale, LIB), Cint, (Ptr{Cvoid}, Cdouble), hd.h, Float64(s)))
# objective(glrm, X, Y; include_regularization) (evaluate_fit.jl:4-23) evaluated on the device
function objective_resident(hd::B200Handle, X::Matrix{Float64}, Y::Matrix{Float64}; include_regularization::Bool=true)
    out = Ref{Cdouble}(0.0)
    check(ccall((:glrmb200_objective, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ref{Cdouble}),
                hd.h, X, Y, include_regularization ? 1 : 0, out))
    out[]
end
# error_metric(glrm, domains; standardize) (evaluate_fit.jl:106-148) and impute(glrm) (evaluate_fit.jl:150) on the device:
# domains default to each loss's own (l.domain); (code, (p0, p1)) rows as include/glrm_b200.h GLRMB200_DOMAIN_*
domrow(d::RealDomain) = (Int32(1), 0.0, 0.0)
domrow(d::BoolDomain) = (Int32(2), 0.0, 0.0)
domrow(d::OrdinalDomain) = (Int32(3), Float64(d.min), Float64(d.max))
domrow(d::CategoricalDomain) = (Int32(4), Float64(d.min), Float64(d.max))
domrow(d::PeriodicDomain) = (Int32(5), Float64(d.T), 0.0)
domrow(d::CountDomain) = (Int32(6), Float64(d.max_count), 0.0)
function domtable(domains)
    rows = [domrow(d) for d in domains]
    Int32[r[1] for r in rows], collect(Iterators.flatten((r[2], r[3]) for r in rows))
end
function error_metric_resident(hd::B200Handle, glrm::GLRM, domains::Array{Domain,1}=Domain[l.domain for l in glrm.losses];
                               standardize::Bool=false)
    code, par = domtable(domains)
    out = Ref{Cdouble}(0.0)
    check(ccall((:glrmb200_error_metric, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Cdouble}, Int32, Ref{Cdouble}),
                hd.h, glrm.X, glrm.Y, code, par, standardize ? 1 : 0, out))
    out[]
end
function impute_resident(hd::B200Handle, glrm::GLRM, domains::Array{Domain,1}=Domain[l.domain for l in glrm.losses])
    code, par = domtable(domains)
    m, n = size(glrm.A)
    Ahat = Array{Float64}(undef, m, n)
    check(ccall((:glrmb200_impute, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}),
                hd.h, glrm.X, glrm.Y, code, par, Ahat))
    Ahat
end
# multi-process plumbing (one Julia worker per GPU): rank 0 makes the id, everybody joins; see INTEGRATION.md
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:glrmb200_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
comm_init!(hd::B200Handle, id::Vector{UInt8}) = check(ccall((:glrmb200_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), hd.h, id))

function feed_ch!(ch::ConvergenceHistory, obj, sec, nrec, verbose)
    for i in 1:nrec                                                             # proxgrad.jl:76,207
        update_ch!(ch, sec[i], obj[i])
        if verbose && i > 1 && (i - 1) % 10 == 0
            println("Iteration $(i-1): objective value = $(obj[i])")           # proxgrad.jl:214-216
        end
    end
end

function factors!(glrm::GLRM)
    X = glrm.X isa Matrix{Float64} ? glrm.X : (glrm.X = Matrix{Float64}(glrm.X))                      # mutated in place
    Y = glrm.Y isa Matrix{Float64} ? glrm.Y : (glrm.Y = Matrix{Float64}(glrm.Y))
    X, Y
end

### FITTING — replaces the body of fit!(::GLRM, ::ProxGradParams) (src/algorithms/proxgrad.jl:34-220)
# on a live handle: warm start from glrm.X / glrm.Y, results back in place, ch appended (cross_validate.jl:174 reuses one ch)
function fit!(hd::B200Handle, glrm::GLRM, params::B200ProxGradParams;
              ch::ConvergenceHistory=ConvergenceHistory("B200ProxGradGLRM"), verbose=true, kwargs...)
    X, Y = factors!(glrm)
    cap = params.max_iter + 1
    obj = zeros(Cdouble, cap); sec = zeros(Cdouble, cap); nrec = Ref{Int32}(0)
    cp = Ref(CParams(params.stepsize, params.max_iter, params.inner_iter_X, params.inner_iter_Y,
                     params.abs_tol, params.rel_tol, params.min_stepsize))
    if verbose println("Fitting GLRM") end                                      # proxgrad.jl:75
    GC.@preserve X Y obj sec begin
        check(ccall((:glrmb200_fit, LIB), Cint,
                    (Ptr{Cvoid}, Ref{CParams}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ref{Int32}, Ptr{Cvoid}),
                    hd.h, cp, X, Y, obj, sec, Int32(cap), nrec, C_NULL))
    end
    feed_ch!(ch, obj, sec, nrec[], verbose)
    return glrm.X, glrm.Y, ch                                                   # proxgrad.jl:219
end

# the drop-in method: encode, create, fit, destroy
function fit!(glrm::GLRM, params::B200ProxGradParams;
              ch::ConvergenceHistory=ConvergenceHistory("B200ProxGradGLRM"), verbose=true, kwargs...)
    hd = B200Handle(glrm; device=params.device, rank=params.rank, nranks=params.nranks)
    try
        return fit!(hd, glrm, params; ch=ch, verbose=verbose, kwargs...)
    finally
        close(hd)
    end
end

# ---- SparseProxGradParams' five fields (src/algorithms/sparse_proxgrad.jl:4-18) + the device -------------------
mutable struct B200SparseProxGradParams <: AbstractParams
    stepsize::Float64
    max_iter::Int
    inner_iter::Int
    abs_tol::Float64
    min_stepsize::Float64
    device::Int
end
B200SparseProxGradParams(stepsize::Number=1.0; max_iter::Int=100, inner_iter::Int=1, abs_tol::Float64=0.00001,
                         min_stepsize::Float64=0.01*stepsize, device::Int=0) =
    B200SparseProxGradParams(Float64(stepsize), max_iter, inner_iter, abs_tol, min_stepsize, device)

struct CSparseParams      # glrmb200_sparse_params
    stepsize::Cdouble
    max_iter::Int32
    inner_iter::Int32
    abs_tol::Cdouble
    min_stepsize::Cdouble
end

# Same handle, glrmb200_fit_sparse instead of glrmb200_fit; the recorded series is the reference's
# (initial objective, one entry per accepted iteration, final duplicate — sparse_proxgrad.jl:50,106,126).
function fit!(hd::B200Handle, glrm::GLRM, params::B200SparseProxGradParams;
              ch::ConvergenceHistory=ConvergenceHistory("B200SparseProxGradGLRM"), verbose=true, kwargs...)
    X, Y = factors!(glrm)
    cap = params.max_iter + 2
    obj = zeros(Cdouble, cap); sec = zeros(Cdouble, cap); nrec = Ref{Int32}(0)
    cp = Ref(CSparseParams(params.stepsize, params.max_iter, params.inner_iter, params.abs_tol, params.min_stepsize))
    GC.@preserve X Y obj sec begin
        check(ccall((:glrmb200_fit_sparse, LIB), Cint,
                    (Ptr{Cvoid}, Ref{CSparseParams}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ref{Int32}, Ptr{Cvoid}),
                    hd.h, cp, X, Y, obj, sec, Int32(cap), nrec, C_NULL))
    end
    feed_ch!(ch, obj, sec, nrec[], verbose)
    return glrm.X, glrm.Y, ch
end
function fit!(glrm::GLRM, params::B200SparseProxGradParams;
              ch::ConvergenceHistory=ConvergenceHistory("B200SparseProxGradGLRM"), verbose=true, kwargs...)
    hd = B200Handle(glrm; device=params.device, gather_only=true)
    try
        return fit!(hd, glrm, params; ch=ch, verbose=verbose, kwargs...)
    finally
        close(hd)
    end
end

end # module
