# julia/eikonal_op.jl -- drop-in replacement for src/eikonal_op.jl of ADTomo.jl.
#
# Same exported names, argument meaning and return shapes as the reference
# (src/eikonal_op.jl:3-21 `eikonal`, :24-32 `eikonal3d`); the TensorFlow custom ops
# `eikonal`/`eikonal_grad`/`eikonal_three_d`/`eikonal_three_d_grad` of libADTomo are replaced by
# `ccall`s into libadtomo_b200.so (include/adtomo_b200.h), wrapped as TensorFlow py_func ops with a
# custom gradient (`custom_gradient_op` below, defined in THIS file) so that scripts which build one op
# per station (scripts/inversion*.jl, tests/test3d.jl, ...) keep running unchanged.  That per-station form
# is the SLOW path (one host round trip and one single-source launch per station); the batched entry points
# at the end of this file (`misfit_grad!`, `model_loss_grad!`, `nccl_*`) are the fast path, and
# julia/inversion_b200.jl is scripts/inversion.jl rewritten around ONE call per loss/gradient evaluation.
#
# EXPERIMENTAL: not executed in the build image (no Julia / ADCME there).  The C symbols, their argument
# order and types are exercised by tests/abi_check.c (compiled against include/adtomo_b200.h), by the Python
# ctypes mirror and by the GPU test-suite.  Layout notes: TensorFlow flattens row-major, so Julia arrays
# are permuted to row-major before a call exactly as `tf.reshape(f, (-1,))` did.
using PyCall
export eikonal, eikonal3d, custom_gradient_op

const LIBADTOMO_B200 = get(ENV, "LIBADTOMO_B200", joinpath(@__DIR__, "..", "adtomo.jl_b200", "libadtomo_b200.so"))

_check(rc::Cint, what) = rc < 0 ? error("$what failed: " * unsafe_string(ccall((:adtomo_last_error, LIBADTOMO_B200), Cstring, ()))) : rc

# row-major flat copy of a Julia (column-major) array, and back
_rowmajor(a::AbstractArray{Float64}) = vec(permutedims(a, reverse(1:ndims(a))))
_from_rowmajor(v::Vector{Float64}, dims) = permutedims(reshape(v, reverse(dims)), reverse(1:length(dims)))

# ---- raw calls (0-based source indices, cf. Eikonal.cpp:125) ---------------------------------
function eikonal2d_forward!(u::Vector{Float64}, f::Vector{Float64}, m::Int, n::Int, h::Float64, ix::Int, jx::Int)
    _check(ccall((:adtomo_eikonal2d_forward, LIBADTOMO_B200), Cint,
                 (Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cdouble, Cint, Cint), u, f, m, n, h, ix, jx), "eikonal2d_forward")
    u
end
function eikonal2d_backward!(gf::Vector{Float64}, gu::Vector{Float64}, u::Vector{Float64}, f::Vector{Float64},
                             m::Int, n::Int, h::Float64, ix::Int, jx::Int)
    _check(ccall((:adtomo_eikonal2d_backward, LIBADTOMO_B200), Cint,
                 (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cdouble, Cint, Cint),
                 gf, gu, u, f, m, n, h, ix, jx), "eikonal2d_backward")
    gf
end
function eikonal3d_forward!(u::Vector{Float64}, u0::Vector{Float64}, f::Vector{Float64}, h::Float64,
                            m::Int, n::Int, l::Int, tol::Float64, verbose::Bool)
    _check(ccall((:adtomo_eikonal3d_forward, LIBADTOMO_B200), Cint,
                 (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cdouble, Cint),
                 u, u0, f, h, m, n, l, tol, verbose ? 1 : 0), "eikonal3d_forward")
    u
end
function eikonal3d_backward!(gu0::Vector{Float64}, gf::Vector{Float64}, gu::Vector{Float64}, u::Vector{Float64},
                             u0::Vector{Float64}, f::Vector{Float64}, h::Float64, m::Int, n::Int, l::Int)
    _check(ccall((:adtomo_eikonal3d_backward, LIBADTOMO_B200), Cint,
                 (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint),
                 gu0, gf, gu, u, u0, f, h, m, n, l), "eikonal3d_backward")
    gu0, gf
end

# ---- TensorFlow op with a custom gradient from two Julia closures (replaces load_op_and_grad) ---
# fwd(inputs...) -> y and bwd(dy, y, inputs...) -> tuple of length(inputs) gradients, all flat Float64 vectors.
# TensorFlow 1.x as shipped with ADCME: tf.py_func + tf.custom_gradient.
function __init_custom_gradient__()
    py"""
    import tensorflow as tf
    def adtomo_make_op(fwd, bwd, n):
        @tf.custom_gradient
        def op(*xs):
            y = tf.py_func(fwd, list(xs), tf.float64)
            def grad(dy):
                g = tf.py_func(bwd, [dy, y] + list(xs), [tf.float64] * n)
                return g if n > 1 else g[0]
            return y, grad
        return op
    """
end
const _custom_gradient_ready = Ref(false)
function custom_gradient_op(fwd, bwd, inputs...)
    if !_custom_gradient_ready[]
        __init_custom_gradient__()
        _custom_gradient_ready[] = true
    end
    py"adtomo_make_op"(fwd, bwd, length(inputs))(inputs...)
end

# ---- operators with the reference's names and signatures ---------------------------------------
function eikonal(f::Union{Array{Float64}, PyObject}, srcx::Int64, srcy::Int64, h::Float64)
    n_, m_ = size(f)                       # rows (y), columns (x): src/eikonal_op.jl:5
    m = m_ - 1; n = n_ - 1
    fwd(fv) = eikonal2d_forward!(zeros(length(fv)), Vector{Float64}(fv), m, n, h, srcx - 1, srcy - 1)
    function bwd(du, u, fv)                # same argument order as EikonalGrad (Eikonal.cpp:50-57)
        (eikonal2d_backward!(zeros(length(fv)), Vector{Float64}(du), Vector{Float64}(u), Vector{Float64}(fv), m, n, h, srcx - 1, srcy - 1),)
    end
    f = convert_to_tensor(f, dtype = Float64)
    fflat = tf.reshape(f, (-1,))
    u = custom_gradient_op(fwd, bwd, fflat)
    u.set_shape((n_ * m_,))
    return tf.reshape(u, (n_, m_))
end

function eikonal3d(u0, f, h, m, n, l, tol, verbose)
    h = Float64(h); tol = Float64(tol)
    fwd(u0v, fv) = eikonal3d_forward!(zeros(m * n * l), Vector{Float64}(u0v), Vector{Float64}(fv), h, m, n, l, tol, Bool(verbose))
    function bwd(du, u, u0v, fv)           # EikonalThreeDGrad input order (EikonalThreeD.cpp:48-57)
        eikonal3d_backward!(zeros(m * n * l), zeros(m * n * l), Vector{Float64}(du), Vector{Float64}(u),
                            Vector{Float64}(u0v), Vector{Float64}(fv), h, m, n, l)   # -> (grad_u0, grad_f)
    end
    u0 = tf.reshape(convert_to_tensor(u0, dtype = Float64), (-1,))
    f = tf.reshape(convert_to_tensor(f, dtype = Float64), (-1,))
    out = custom_gradient_op(fwd, bwd, u0, f)
    out.set_shape((m * n * l,))
    return tf.reshape(out, (m, n, l))
end


# ---- batched / fused entry points: the fast path -----------------------------------------------
# One context per process (= per GPU).  All arrays below are flat, row-major, Float64 / Int32 host arrays.
const ADTOMO_HOST = Cint(0)
mutable struct AdtomoContext
    handle::Ptr{Cvoid}
end
function adtomo_context(device::Integer = -1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:adtomo_create, LIBADTOMO_B200), Cint, (Ptr{Ptr{Cvoid}}, Cint), h, device), "adtomo_create")
    ctx = AdtomoContext(h[])
    finalizer(c -> (c.handle != C_NULL && ccall((:adtomo_destroy, LIBADTOMO_B200), Cint, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), ctx)
    ctx
end

# 8-corner source table of scripts/inversion.jl:48-60 in CSR form (0-based row-major node indices).
# sta: numsta x 3 fractional 1-based node coordinates (allsta.x/y/z); vel0: m x n x l.
function corner_sources(sta::AbstractMatrix{Float64}, h::Float64, vel0::Array{Float64,3})
    m, n, l = size(vel0)
    ptr = Int32[0]; idx = Int32[]; val = Float64[]
    for i in 1:size(sta, 1)
        x, y, z = sta[i, 1], sta[i, 2], sta[i, 3]
        for cx in (ceil(Int, x), floor(Int, x)), cy in (ceil(Int, y), floor(Int, y)), cz in (ceil(Int, z), floor(Int, z))
            push!(idx, Int32(((cx - 1) * n + (cy - 1)) * l + (cz - 1)))
            push!(val, sqrt((x - cx)^2 + (y - cy)^2 + (z - cz)^2) * h / vel0[cx, cy, cz])
        end
        push!(ptr, Int32(length(idx)))
    end
    ptr, idx, val
end

# adtomo_eikonal3d_misfit_grad: misfit and slowness gradient of ALL stations of this rank in one call
# (per-source work of scripts/inversion.jl:46-105).  packed: N+1 doubles, [grad_f | misfit].
# rcv: 3 x numeve (0-based fractional coordinates, column j = event j => row-major E x 3 in memory),
# uobs / qua: numeve x numsta Julia matrices hold the row-major S x E tables (uobs_jl[e, s] = uobs[s, e]).
function misfit_grad!(ctx::AdtomoContext, packed::Vector{Float64}, f::Vector{Float64}, h::Float64, m::Int, n::Int, l::Int,
                      tol::Float64, ptr::Vector{Int32}, idx::Vector{Int32}, val::Vector{Float64}, u0_fill::Float64,
                      rcv::Matrix{Float64}, uobs::Matrix{Float64}, qua::Matrix{Float64}; max_rounds::Int = 0)
    S = length(ptr) - 1; E = size(rcv, 2)
    mis = Ref{Cdouble}(0.0)
    rc = ccall((:adtomo_eikonal3d_misfit_grad, LIBADTOMO_B200), Cint,
               (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cdouble, Cint, Cint,
                Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
               ctx.handle, mis, packed, f, h, m, n, l, tol, max_rounds, S, ptr, idx, val, u0_fill, E, rcv, uobs, qua,
               C_NULL, ADTOMO_HOST)
    _check(rc, "adtomo_eikonal3d_misfit_grad")
    mis[]
end

# adtomo_model_loss_grad: scripts/inversion.jl:42-43,61,96-121 in one call -- x (var_change, row-major) in,
# packed = [d loss / d x | loss] out; parametrisation, chain rule and regulariser run on the device.
function model_loss_grad!(ctx::AdtomoContext, packed::Union{Vector{Float64},Nothing}, x::Vector{Float64}, vel0::Vector{Float64},
                          lambda::Float64, smooth_hor::Int, smooth_ver::Int, add_reg::Bool, h::Float64, m::Int, n::Int, l::Int,
                          tol::Float64, ptr::Vector{Int32}, idx::Vector{Int32}, val::Vector{Float64}, u0_fill::Float64,
                          rcv::Matrix{Float64}, uobs::Matrix{Float64}, qua::Matrix{Float64}; max_rounds::Int = 0)
    S = length(ptr) - 1; E = size(rcv, 2)
    loss = Ref{Cdouble}(0.0)
    rc = ccall((:adtomo_model_loss_grad, LIBADTOMO_B200), Cint,
               (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cdouble,
                Cint, Cint, Cint, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cdouble},
                Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
               ctx.handle, loss, packed === nothing ? C_NULL : packed, x, vel0, lambda, smooth_hor, smooth_ver, add_reg ? 1 : 0, h,
               m, n, l, tol, max_rounds, S, ptr, idx, val, u0_fill, E, rcv, uobs, qua, C_NULL, ADTOMO_HOST)
    _check(rc, "adtomo_model_loss_grad")
    loss[]
end

# joint P+S driver (scripts/inversion_joint.jl:49-51,80,140-166): begin, one add_phase per phase, finish
model_begin!(ctx::AdtomoContext, x::Vector{Float64}, vel0::Vector{Float64}, m::Int, n::Int, l::Int) =
    _check(ccall((:adtomo_model_begin, LIBADTOMO_B200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint, Cint),
                 ctx.handle, x, vel0, m, n, l, ADTOMO_HOST), "adtomo_model_begin")
function model_add_phase!(ctx::AdtomoContext, scale::Float64, h::Float64, tol::Float64, ptr::Vector{Int32}, idx::Vector{Int32},
                          val::Vector{Float64}, u0_fill::Float64, rcv::Matrix{Float64}, uobs::Matrix{Float64},
                          qua::Matrix{Float64}; want_grad::Bool = true, max_rounds::Int = 0)
    S = length(ptr) - 1; E = size(rcv, 2)
    mis = Ref{Cdouble}(0.0); gs = Ref{Cdouble}(0.0)
    rc = ccall((:adtomo_model_add_phase, LIBADTOMO_B200), Cint,
               (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}, Cdouble, Cint,
                Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint),
               ctx.handle, scale, h, tol, max_rounds, S, ptr, idx, val, u0_fill, E, rcv, uobs, qua, C_NULL, mis, gs,
               want_grad ? 1 : 0, ADTOMO_HOST)
    _check(rc, "adtomo_model_add_phase")
    mis[], gs[]            # misfit of the phase, d misfit / d scale (the gradient of `pvs`)
end
function model_finish!(ctx::AdtomoContext, packed::Union{Vector{Float64},Nothing}, lambda::Float64, smooth_hor::Int, smooth_ver::Int,
                       add_reg::Bool)
    loss = Ref{Cdouble}(0.0)
    rc = ccall((:adtomo_model_finish, LIBADTOMO_B200), Cint,
               (Ptr{Cvoid}, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
               ctx.handle, lambda, smooth_hor, smooth_ver, add_reg ? 1 : 0, packed === nothing ? 0 : 1, loss,
               packed === nothing ? C_NULL : packed, ADTOMO_HOST)
    _check(rc, "adtomo_model_finish")
    loss[]
end

# ---- multi-GPU: replaces mpi_bcast (backward) + mpi_sum of scripts/inversion.jl:44,123 ------------
# Rank 0 creates the id and hands its 128 bytes to the other ranks (MPI.Bcast!, a file, ...).
function nccl_unique_id()
    id = zeros(UInt8, 128)
    _check(ccall((:adtomo_nccl_unique_id, LIBADTOMO_B200), Cint, (Ptr{UInt8},), id), "adtomo_nccl_unique_id")
    id
end
nccl_init!(ctx::AdtomoContext, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    _check(ccall((:adtomo_nccl_init, LIBADTOMO_B200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), ctx.handle, id, rank, nranks), "adtomo_nccl_init")
nccl_allreduce_sum!(ctx::AdtomoContext, buf::Vector{Float64}) =
    _check(ccall((:adtomo_nccl_allreduce_sum, LIBADTOMO_B200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clonglong, Cint),
                 ctx.handle, buf, length(buf), ADTOMO_HOST), "adtomo_nccl_allreduce_sum")
nccl_finalize!(ctx::AdtomoContext) = ccall((:adtomo_nccl_finalize, LIBADTOMO_B200), Cint, (Ptr{Cvoid},), ctx.handle)
