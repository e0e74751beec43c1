# julia/inversion_b200.jl -- scripts/inversion.jl (P-wave travel-time inversion) driven by ONE library call per
# loss/gradient evaluation instead of one TensorFlow op per station.
#
# What stays the reference's: the input files (range.txt, allsta.csv / alleve.csv, vel0_p.h5, uobs_p.h5, qua_p.h5,
# config.json), the station sharding `rank+1:nproc:numsta` (scripts/inversion.jl:36-38), the model parametrisation
# fvar = 2*sigmoid(var_change) - 1 + vel0 (:42-43), slowness 1 ./ fvar (:61), the 8-corner sources (:48-60), the
# trilinear receiver sampling and weighted misfit with `uobs == -1` skipped (:64-105), the periodic box-filter L1
# regulariser (:107-121), Optim.jl's LBFGS with InitialStatic + BackTracking (src/mpi_optimize.jl:35-39), the
# "iter k, current loss=" / "STEP k" log lines and the `iter_k.h5` ("data") checkpoints every `steps` gradient
# evaluations (:15-29).
# What changes: no TensorFlow graph, no MPI.  `adtomo_model_loss_grad` does the parametrisation, every forward solve,
# the sampling, the misfit, every adjoint solve, the chain rule and the regulariser on this rank's GPU; one NCCL
# all-reduce of the packed [gradient | loss] buffer replaces mpi_bcast's backward + mpi_sum (:44,123).  The
# regulariser is added by rank 0 only (the reference adds it on every rank before mpi_sum, i.e. nproc times).
#
# Launch: one process per GPU, e.g.  for r in 0..P-1:  ADTOMO_RANK=r ADTOMO_NPROC=P julia inversion_b200.jl
# (rank 0 writes the NCCL id to `nccl_id.bin`, the others read it).
#
# EXPERIMENTAL: not executed in the build image (no Julia there); the Python statement of the same driver
# (adtomo.jl_b200/optimize.py: DeviceVelocityModel + gpu_optimize) is what the test-suite runs.
using CSV, DataFrames, HDF5, JSON, Optim, LineSearches
include(joinpath(@__DIR__, "eikonal_op.jl"))

rank = parse(Int, get(ENV, "ADTOMO_RANK", "0"))
nproc = parse(Int, get(ENV, "ADTOMO_NPROC", "1"))

region = "demo/"
folder = "../local/" * region * "readin_data/"
config = JSON.parsefile(folder * "config.json")["inversion"]
rfile = open(folder * "range.txt", "r")
m = parse(Int, readline(rfile)); n = parse(Int, readline(rfile))
l = parse(Int, readline(rfile)); h = parse(Float64, readline(rfile))
close(rfile)

allsta = CSV.read(folder * "sta_eve/allsta.csv", DataFrame); numsta = size(allsta, 1)
alleve = CSV.read(folder * "sta_eve/alleve.csv", DataFrame); numeve = size(alleve, 1)
vel0 = h5read(folder * "velocity/vel0_p.h5", "data")
uobs = h5read(folder * "for_P/uobs_p.h5", "matrix")
qua = h5read(folder * "for_P/qua_p.h5", "matrix")

mine = rank+1:nproc:numsta                                   # scripts/inversion.jl:36-38
sta = hcat(allsta.x[mine], allsta.y[mine], allsta.z[mine])
ptr, idx, val = corner_sources(Matrix{Float64}(sta), h, vel0)
rcv = permutedims(hcat(alleve.x, alleve.y, alleve.z) .- 1.0)   # 3 x numeve, 0-based fractional node coordinates
uobs_t = permutedims(uobs[mine, :])                          # numeve x S in Julia = row-major S x E for the library
qua_t = permutedims(qua[mine, :])
vel0_rm = _rowmajor(vel0)

ctx = adtomo_context(-1)
if nproc > 1
    idfile = "nccl_id.bin"
    if rank == 0
        write(idfile * ".tmp", nccl_unique_id()); mv(idfile * ".tmp", idfile, force = true)
    else
        while !isfile(idfile); sleep(0.1); end
    end
    nccl_init!(ctx, read(idfile), rank, nproc)
end

N = m * n * l
packed = zeros(N + 1)
lambda = Float64(config["lambda_p"]); sh = Int(config["smooth_hor"]); sv = Int(config["smooth_ver"])

# x: var_change flattened row-major.  One call evaluates loss and gradient of this rank's stations.
function evaluate!(x::Vector{Float64}, want_grad::Bool)
    loss = model_loss_grad!(ctx, want_grad ? packed : nothing, x, vel0_rm, lambda, sh, sv, rank == 0, h, m, n, l, 1e-3,
                            ptr, idx, val, 1000.0, rcv, uobs_t, qua_t)
    if !want_grad
        packed[N + 1] = loss
    end
    nproc > 1 && nccl_allreduce_sum!(ctx, packed)           # [gradient | loss] summed over ranks
    packed[N + 1]
end

loc = folder * "inv_P_" * string(config["lambda_p"]) * "/intermediate/"
steps = config["steps"]
fcnt = Ref(0); gcnt = Ref(0)
function f(x)
    L = evaluate!(x, false)
    fcnt[] += 1
    rank == 0 && println("iter $(fcnt[]), current loss=", L)             # src/mpi_optimize.jl:15-17
    L
end
function g!(G, x)
    evaluate!(x, true)
    G[:] = packed[1:N]
    gcnt[] += 1
    if rank == 0
        println("================== STEP $(gcnt[]) ==================")     # :22-25
        if mod(gcnt[], steps) == 0
            isdir(loc) || mkpath(loc)
            h5write(joinpath(loc, "iter_$(gcnt[]).h5"), "data", x)             # :26-28
        end
    end
    G
end

rank == 0 && println("[ranks = $nproc] Optimization starts...")
method = LBFGS(alphaguess = InitialStatic(), linesearch = LineSearches.BackTracking())     # :35-39
result = Optim.optimize(f, g!, zeros(N), method, Optim.Options(iterations = config["iterations"]))
rank == 0 && @info result
nproc > 1 && nccl_finalize!(ctx)
