# bench_ref.jl — the TRUE reference (LowRankModels.jl's own threaded fit!) on the same generated inputs.
# SHIPPED, NOT RUN: there is no `julia` binary in the build image or on the GPU box (BASELINE.md, baseline B3).
#
#   python tools/dump_config.py C2 1 /tmp/c2            # writes rows/cols/vals/X0/Y0 as raw little-endian files
#   JULIA_NUM_THREADS=$(nproc) julia --project=/path/to/LowRankModels.jl baseline/julia/bench_ref.jl /tmp/c2
#
# Note: plain `fit!(glrm)` on a SparseMatrixCSC selects the single-threaded SparseProxGradParams (src/fit.jl:13-15);
# the threaded hot path (src/algorithms/proxgrad_multithread.jl) must be requested with ProxGradParams().
using LowRankModels, SparseArrays

dir = ARGS[1]
meta = parse.(Int, split(read(joinpath(dir, "meta.txt"), String)))      # m n k nnz
m, n, k, nnz = meta
rd(name, T, cnt) = (a = Vector{T}(undef, cnt); read!(joinpath(dir, name), a); a)
rows = rd("rows.i64", Int64, nnz) .+ 1
cols = rd("cols.i64", Int64, nnz) .+ 1
vals = rd("vals.f64", Float64, nnz)
X0 = reshape(rd("X0.f64", Float64, k * m), k, m)
Y0 = reshape(rd("Y0.f64", Float64, k * n), k, n)

A = sparse(rows, cols, vals, m, n)
glrm = GLRM(A, QuadLoss(), QuadReg(0.1), QuadReg(0.1), k; X=copy(X0), Y=copy(Y0))   # obs = findall(!iszero, A)
params = ProxGradParams(1.0, max_iter=10, abs_tol=0.0, rel_tol=0.0)
println("threads = ", Threads.nthreads())
X, Y, ch = fit!(glrm, params, verbose=false)
dt = diff(ch.times)
println("objective: ", ch.objective)
println("seconds per iteration: ", dt)
println("observed-entries/sec/iter: ", nnz / (sum(dt[2:end]) / (length(dt) - 1)))
