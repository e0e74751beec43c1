# julia/eikonal_op.jl -- drop-in replacement for src/eikonal_op.jl of ADTomo.jl.
#
# Same exported names, argument meaning and return shapes as the reference
# (src/eikonal_op.jl:3-21 `eikonal`, :24-32 `eikonal3d`); the TensorFlow custom ops
# `eikonal`/`eikonal_grad`/`eikonal_three_d`/`eikonal_three_d_grad` of libADTomo are replaced by
# `ccall`s into libadtomo_b200.so (include/adtomo_b200.h), wrapped as ADCME py_func custom-gradient
# ops so that the inversion scripts (scripts/inversion*.jl, tests/test3d.jl, ...) run unchanged.
#
# NOT EXECUTED in the build image (no Julia / ADCME there); the same C symbols are exercised by
# the Python ctypes mirror and the GPU test-suite.  Layout notes: TensorFlow flattens row-major, so
# Julia arrays are permuted to row-major before the call exactly as `tf.reshape(f, (-1,))` did.
export eikonal, eikonal3d

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

# ---- ADCME operators with custom gradients (replaces load_op_and_grad) ------------------------
# ADCME's `py_func`-style custom op: forward and gradient are Julia closures on flat Float64 vectors.
function eikonal(f::Union{Array{Float64}, PyObject}, srcx::Int64, srcy::Int64, h::Float64)
    n_, m_ = size(f)                       # rows (y), columns (x): src/eikonal_op.jl:5
    m = m_ - 1; n = n_ - 1
    fwd(fv) = eikonal2d_forward!(zeros(length(fv)), Vector{Float64}(fv), m, n, h, srcx - 1, srcy - 1)
    function bwd(du, u, fv)                # same argument order as EikonalGrad (Eikonal.cpp:50-57)
        eikonal2d_backward!(zeros(length(fv)), Vector{Float64}(du), Vector{Float64}(u), Vector{Float64}(fv), m, n, h, srcx - 1, srcy - 1)
    end
    f = convert_to_tensor(f, dtype = Float64)
    fflat = tf.reshape(f, (-1,))
    u = ADCME.custom_gradient_op(fwd, bwd, fflat)        # see INTEGRATION.md for the 10-line helper
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
    out = ADCME.custom_gradient_op(fwd, bwd, u0, f)
    return tf.reshape(out, (m, n, l))
end
