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
                      lastentry1, lastentry_unpenalized, OrdinalReg, MNLOrdinalReg

export B200ProxGradParams, B200SparseProxGradParams

const LIB = get(ENV, "GLRMB200_LIB", joinpath(@__DIR__, "..", "csrc", "libglrm_b200.so"))

# ---- ProxGradParams' seven fields (src/algorithms/proxgrad.jl:4-31) + the device -------------------
mutable struct B200ProxGradParams <: AbstractParams
    stepsize::Float64
    max_iter::Int
    inner_iter_X::Int
    inner_iter_Y::Int
    abs_tol::Float64
    rel_tol::Float64
    min_stepsize::Float64
    device::Int
end
function B200ProxGradParams(stepsize::Number=1.0; max_iter::Int=100, inner_iter_X::Int=1, inner_iter_Y::Int=1,
                            inner_iter::Int=1, abs_tol::Number=0.00001, rel_tol::Number=0.0001,
                            min_stepsize::Number=0.01*stepsize, device::Int=0)
    B200ProxGradParams(Float64(stepsize), max_iter, max(inner_iter_X, inner_iter), max(inner_iter_Y, inner_iter),
                       Float64(abs_tol), Float64(rel_tol), Float64(min_stepsize), device)
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

regrow(r::ZeroReg)                 = (0, 0.0)
regrow(r::QuadReg)                 = (1, r.scale)
regrow(r::QuadConstraint)          = (2, r.max_2norm)
regrow(r::OneReg)                  = (3, r.scale)
regrow(r::NonNegConstraint)        = (4, 0.0)
regrow(r::NonNegOneReg)            = (5, r.scale)
regrow(r::OneSparseConstraint)     = (6, 0.0)
regrow(r::KSparseConstraint)       = (7, Float64(r.k))
regrow(r::UnitOneSparseConstraint) = (8, 0.0)
regrow(r::SimplexConstraint)       = (9, 0.0)
regrow(r::lastentry1)              = ((c, p) = regrow(r.r); (c | 0x100, p))
regrow(r::lastentry_unpenalized)   = ((c, p) = regrow(r.r); (c | 0x200, p))
regrow(r::OrdinalReg)              = ((c, p) = regrow(r.r); (c | 0x400, p))   # block regularizers (ry only)
regrow(r::MNLOrdinalReg)           = ((c, p) = regrow(r.r); (c | 0x800, p))
regrow(r::Regularizer) = throw(ArgumentError("$(typeof(r)) has no B200 device implementation (no CPU fallback)"))

function regtable(rs)
    rows = map(regrow, rs)
    if all(==(rows[1]), rows)          # the usual fillcopies case: one shared row
        rows = rows[1:1]
    end
    codes = Int32[c for (c, _) in rows]
    params = zeros(Cdouble, NREGP, length(rows))
    for (i, (_, p)) in enumerate(rows); params[1, i] = p; end
    codes, params
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

### FITTING — replaces the body of fit!(::GLRM, ::ProxGradParams) (src/algorithms/proxgrad.jl:34-220)
function fit!(glrm::GLRM, params::B200ProxGradParams;
              ch::ConvergenceHistory=ConvergenceHistory("B200ProxGradGLRM"),
              verbose=true, kwargs...)
    return _fit_with(glrm, params, ch, verbose) do handle, X, Y, obj, sec, cap, nrec
        cp = Ref(CParams(params.stepsize, params.max_iter, params.inner_iter_X, params.inner_iter_Y,
                         params.abs_tol, params.rel_tol, params.min_stepsize))
        ccall((:glrmb200_fit, LIB), Cint,
              (Ptr{Cvoid}, Ref{CParams}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ref{Int32}, Ptr{Cvoid}),
              handle, cp, X, Y, obj, sec, cap, nrec, C_NULL)
    end
end

# encode the GLRM, create the handle, run `call(handle, X, Y, obj, sec, cap, nrec)`, feed `ch`, destroy the handle
function _fit_with(call::Function, glrm::GLRM, params::B200ProxGradParams, ch::ConvergenceHistory, verbose)
    A = glrm.A
    m, n = size(A)
    k = glrm.k
    yidxs = get_yidxs(glrm.losses)
    d = maximum(yidxs[end])
    size(glrm.Y) == (k, d) || error("size(glrm.Y) must be (k, embedding_dim(losses)) = ($k, $d)")   # proxgrad.jl:55-63
    X = glrm.X isa Matrix{Float64} ? glrm.X : (glrm.X = Matrix{Float64}(glrm.X))                      # mutated in place
    Y = glrm.Y isa Matrix{Float64} ? glrm.Y : (glrm.Y = Matrix{Float64}(glrm.Y))

    lrows = map(lossrow, glrm.losses)
    lcode = Int32[c for (c, _) in lrows]
    lparam = zeros(Cdouble, NLOSSP, n)
    for (f, (_, p)) in enumerate(lrows), (i, v) in enumerate(p); lparam[i, f] = v; end
    rxc, rxp = regtable(glrm.rx)
    ryc, ryp = regtable(glrm.ry)

    full = all(o -> o == 1:n, glrm.observed_features) && all(o -> o == 1:m, glrm.observed_examples)
    if full
        dense = Matrix{Cdouble}(map(labelval, A))
        rptr = Int64[]; ridx = Int32[]; rval = Cdouble[]; cptr = Int64[]; cidx = Int32[]; cval = Cdouble[]
    else
        dense = Cdouble[]
        rptr, ridx, rval = flatten_obs(glrm.observed_features, A, true)
        cptr, cidx, cval = flatten_obs(glrm.observed_examples, A, false)
    end

    cap = params.max_iter + 1
    obj = zeros(Cdouble, cap); sec = zeros(Cdouble, cap); nrec = Ref{Int32}(0)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    if verbose println("Fitting GLRM") end                                      # proxgrad.jl:75
    GC.@preserve lcode lparam rxc rxp ryc ryp dense rptr ridx rval cptr cidx cval X Y obj sec begin
        prob = Ref(CProblem(m, n, k, d, pointer(lcode), pointer(lparam),
                            length(rxc), pointer(rxc), pointer(rxp), length(ryc), pointer(ryc), pointer(ryp),
                            full ? 1 : 0, full ? pointer(dense) : C_NULL,
                            full ? C_NULL : pointer(rptr), full ? C_NULL : pointer(ridx), full ? C_NULL : pointer(rval),
                            full ? C_NULL : pointer(cptr), full ? C_NULL : pointer(cidx), full ? C_NULL : pointer(cval)))
        check(ccall((:glrmb200_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{CProblem}, Int32, Int32, Int32),
                    handle, prob, params.device, 0, 1))
        try
            check(call(handle[], X, Y, obj, sec, Int32(cap), nrec))
        finally
            ccall((:glrmb200_destroy, LIB), Cint, (Ptr{Cvoid},), handle[])
        end
    end
    for i in 1:nrec[]                                                           # proxgrad.jl:76,207
        update_ch!(ch, sec[i], obj[i])
        if verbose && i > 1 && (i - 1) % 10 == 0
            println("Iteration $(i-1): objective value = $(obj[i])")           # proxgrad.jl:214-216
        end
    end
    return glrm.X, glrm.Y, ch                                                   # proxgrad.jl:219
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

# Same encoding as above, then glrmb200_fit_sparse instead of glrmb200_fit; the recorded series is the reference's
# (initial objective, one entry per accepted iteration, final duplicate — sparse_proxgrad.jl:50,106,126).
function fit!(glrm::GLRM, params::B200SparseProxGradParams;
              ch::ConvergenceHistory=ConvergenceHistory("B200SparseProxGradGLRM"), verbose=true, kwargs...)
    pg = B200ProxGradParams(params.stepsize; max_iter=params.max_iter + 1, device=params.device)
    return _fit_with(glrm, pg, ch, verbose) do handle, X, Y, obj, sec, cap, nrec
        cp = Ref(CSparseParams(params.stepsize, params.max_iter, params.inner_iter, params.abs_tol, params.min_stepsize))
        ccall((:glrmb200_fit_sparse, LIB), Cint,
              (Ptr{Cvoid}, Ref{CSparseParams}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ref{Int32}, Ptr{Cvoid}),
              handle, cp, X, Y, obj, sec, cap, nrec, C_NULL)
    end
end

end # module
