# WaveletsExtB200.jl -- the reference-side binding a WaveletsExt.jl maintainer would add: a minimal device-array type and
# METHODS OF THE REFERENCE'S OWN FUNCTIONS (import + extend, so reference code dispatches here for device arrays) that `ccall`
# libwx_b200.so (C ABI: include/wx_b200.h).
#
# STATUS: UNTESTED UNDER JULIA.  Julia is not available in the build image or on the GPU box, so this file has never been
# loaded.  What IS checked (tests/test_julia_shim.py, CPU): every `ccall` names a symbol declared in include/wx_b200.h with
# the declared number and kind of arguments, every reference name imported below has at least one method defined here, and
# symbols are resolved with Libdl (a `ccall((f, LIB), ...)` with a computed `f` would not lower).  The same ABI is exercised
# 1:1 by the Python host mirror (waveletsext.jl_b200/) in tests/.  See INTEGRATION.md.
module WaveletsExtB200

using Libdl
using Wavelets
import Wavelets: WT
import Wavelets.Util: maxtransformlevels, maketree, isvalidtree
import Wavelets.Transforms: wpt, wpt!, iwpt, iwpt!
import Wavelets.Threshold: bestbasistree
import WaveletsExt
import WaveletsExt.DWT: wpd, wpd!, iwpd, iwpd!, wpdall, iwpdall, wptall, iwptall, dwtall, idwtall, dwt_step!, idwt_step!
import WaveletsExt.SWT: sdwt_step!, isdwt_step!, sdwt, sdwt!, swpt, swpt!, swpd, swpd!, isdwt, isdwt!, iswpt, iswpt!, iswpd, iswpd!,
                        sdwtall, swptall, swpdall, isdwtall, iswptall, iswpdall
import WaveletsExt.ACWT: acdwt_step!, iacdwt_step!, acdwt, acdwt!, acwpt, acwpt!, acwpd, acwpd!, iacdwt, iacdwt!, iacwpt, iacwpt!,
                         iacwpd, iacwpd!, acdwtall, acwptall, acwpdall, iacdwtall, iacwptall, iacwpdall, make_acreverseqmfpair
import WaveletsExt.BestBasis: tree_costs, bestbasistreeall, JBB, LSDB, BB, LoglpCost, NormCost, ShannonEntropyCost, LogEnergyEntropyCost
import WaveletsExt.Utils: getbasiscoef, getbasiscoefall, nodelength, gettreelength
import WaveletsExt.SIWT: sidwt_step!, isidwt_step!
import WaveletsExt.WaveMult: ns_dwt, ns_idwt

export B200Array, B200Comm, wpdall_host, wpd_bestbasis_host, comm_init_all, comm_unique_id, comm_destroy, allreduce!

# ---- library handle: symbols are looked up once with Libdl and ccall'ed through the pointer ---------------------------------
const LIBPATH = get(ENV, "WX_B200_LIB", joinpath(@__DIR__, "..", "waveletsext.jl_b200", "libwx_b200.so"))
const LIBH = Ref{Ptr{Cvoid}}(C_NULL)
const SYMS = Dict{Symbol,Ptr{Cvoid}}()
function sym(name::Symbol)
    get!(SYMS, name) do
        LIBH[] == C_NULL && (LIBH[] = Libdl.dlopen(LIBPATH))
        Libdl.dlsym(LIBH[], name)
    end
end
sfx(::Type{Float64}) = "f64"
sfx(::Type{Float32}) = "f32"
fsym(stem::Symbol, ::Type{T}) where T = sym(Symbol(stem, :_, sfx(T)))      # wx_<stem>_<f64|f32>

struct WxError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall(sym(:wx_last_error), Cstring, ()))
    rc == 1 && throw(AssertionError(msg))          # WX_EINVAL: what the reference raises with @assert
    rc == 4 && throw(OutOfMemoryError())
    throw(WxError(rc, msg))
end

const WxFloat = Union{Float32,Float64}
const NOSTREAM = C_NULL                            # cudaStream_t 0

# ---- minimal device array -----------------------------------------------------------------------------------------------------
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{T}
    dims::NTuple{N,Int}
    dev::Int
    parent::Any                                    # aliases keep their owner alive; owners hold `nothing`
    function B200Array{T,N}(dims::NTuple{N,Int}; dev::Int=0) where {T,N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:wx_set_device), Cint, (Cint,), dev))
        check(ccall(sym(:wx_malloc), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), p, prod(dims) * sizeof(T)))
        a = new{T,N}(Ptr{T}(p[]), dims, dev, nothing)
        finalizer(x -> ccall(sym(:wx_free), Cint, (Ptr{Cvoid},), x.ptr), a)
        return a
    end
    # same memory, other shape (no finalizer: the parent owns the allocation)
    function B200Array{T,N}(parent::B200Array{T}, dims::NTuple{N,Int}) where {T,N}
        @assert prod(dims) == length(parent)
        return new{T,N}(parent.ptr, dims, parent.dev, parent)
    end
end
Base.size(a::B200Array) = a.dims
Base.similar(a::B200Array{T}, ::Type{S}, dims::Dims{N}) where {T,S,N} = B200Array{S,N}(dims; dev=a.dev)
Base.getindex(::B200Array, i...) = error("B200Array: scalar indexing is not supported; copy to the host with Array(a)")
Base.setindex!(::B200Array, v, i...) = error("B200Array: scalar indexing is not supported; use copyto!(a, host_array)")
usedev(a::B200Array) = check(ccall(sym(:wx_set_device), Cint, (Cint,), a.dev))

function Base.copyto!(dst::B200Array{T}, src::Array{T}) where T
    @assert length(dst) == length(src)
    usedev(dst)
    check(ccall(sym(:wx_h2d), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), dst.ptr, src, sizeof(src), NOSTREAM))
    check(ccall(sym(:wx_stream_sync), Cint, (Ptr{Cvoid},), NOSTREAM))
    return dst
end
function Base.copyto!(dst::Array{T}, src::B200Array{T}) where T
    @assert length(dst) == length(src)
    usedev(src)
    check(ccall(sym(:wx_d2h), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), dst, src.ptr, sizeof(dst), NOSTREAM))
    check(ccall(sym(:wx_stream_sync), Cint, (Ptr{Cvoid},), NOSTREAM))
    return dst
end
B200Array(x::Array{T,N}; dev::Int=0) where {T,N} = copyto!(B200Array{T,N}(size(x); dev=dev), x)
Base.Array(a::B200Array{T,N}) where {T,N} = copyto!(Array{T,N}(undef, a.dims), a)

treebytes(t::BitVector) = UInt8.(t)
qmfpair(wt::OrthoFilter) = WT.makereverseqmfpair(wt, true)      # (g, h) = (scaling, detail), Float64 (DWT.jl:141)
function acpair(wt::OrthoFilter)                                 # (h, g) = (Qmf, Pmf) as acdwt_step! consumes them (ACWT.jl:120-131)
    Pmf, Qmf = make_acreverseqmfpair(wt)
    return Vector{Float64}(Qmf), Vector{Float64}(Pmf)
end
maxlev(a::B200Array) = maxtransformlevels(minimum(size(a)))
sigdims(xw::B200Array, batch::Bool) = size(xw)[1:end-(batch ? 2 : 1)]

# ---- single steps ---------------------------------------------------------------------------------------------------------------
# dwt_step!(w1, w2, v, h, g)   dwt/dwt_one_level.jl:79-107
function dwt_step!(w₁::B200Array{T,1}, w₂::B200Array{T,1}, v::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    @assert length(w₁) == length(w₂) == length(v) ÷ 2
    @assert length(h) == length(g)
    usedev(v)
    check(ccall(fsym(:wx_dwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                w₁.ptr, w₂.ptr, v.ptr, length(v), h, g, length(h), NOSTREAM))
    return w₁, w₂
end
# idwt_step!(v, w1, w2, h, g)   dwt/dwt_one_level.jl:192-223
function idwt_step!(v::B200Array{T,1}, w₁::B200Array{T,1}, w₂::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    @assert length(w₁) == length(w₂) == length(v) ÷ 2
    @assert length(h) == length(g)
    usedev(v)
    check(ccall(fsym(:wx_idwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                v.ptr, w₁.ptr, w₂.ptr, length(v), h, g, length(h), NOSTREAM))
    return v
end
# 2-D dwt_step!(w1..w4, v, h, g, temp)   dwt/dwt_one_level.jl:319-354 (temp is not needed on the device)
function dwt_step!(w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2}, v::B200Array{T,2},
                   h::Vector{Float64}, g::Vector{Float64}, temp=nothing) where T<:WxFloat
    @assert size(w₁) == size(w₂) == size(w₃) == size(w₄)
    @assert size(v) == 2 .* size(w₁)
    usedev(v)
    check(ccall(fsym(:wx_dwt_step2, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Clong, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                w₁.ptr, w₂.ptr, w₃.ptr, w₄.ptr, v.ptr, size(w₁, 1), size(w₁, 2), h, g, length(h), NOSTREAM))
    return w₁, w₂, w₃, w₄
end
# 2-D idwt_step!(v, w1..w4, h, g, temp)   dwt/dwt_one_level.jl:401-436
function idwt_step!(v::B200Array{T,2}, w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2},
                    h::Vector{Float64}, g::Vector{Float64}, temp=nothing) where T<:WxFloat
    @assert size(w₁) == size(w₂) == size(w₃) == size(w₄)
    @assert size(v) == 2 .* size(w₁)
    usedev(v)
    check(ccall(fsym(:wx_idwt_step2, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Clong, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                v.ptr, w₁.ptr, w₂.ptr, w₃.ptr, w₄.ptr, size(w₁, 1), size(w₁, 2), h, g, length(h), NOSTREAM))
    return v
end
# sdwt_step!(w1, w2, v, d, h, g)   swt/swt_one_level.jl:99-127
function sdwt_step!(w₁::B200Array{T,1}, w₂::B200Array{T,1}, v::B200Array{T,1}, d::Integer, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    @assert length(w₁) == length(w₂) == length(v)
    @assert length(h) == length(g)
    usedev(v)
    check(ccall(fsym(:wx_sdwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                w₁.ptr, w₂.ptr, v.ptr, length(v), d, h, g, length(h), NOSTREAM))
    return w₁, w₂
end
# isdwt_step!(v, w1, w2, d, h, g)   swt/swt_one_level.jl:257-277 (average based; returns nothing like the reference)
function isdwt_step!(v::B200Array{T,1}, w₁::B200Array{T,1}, w₂::B200Array{T,1}, d::Integer, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    usedev(v)
    check(ccall(fsym(:wx_isdwt_step_avg, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                v.ptr, w₁.ptr, w₂.ptr, length(v), d, h, g, length(h), NOSTREAM))
    return nothing
end
# isdwt_step!(v, w1, w2, d, sv, sw, h, g; add2out)   swt/swt_one_level.jl:279-318 (shift based)
function isdwt_step!(v::B200Array{T,1}, w₁::B200Array{T,1}, w₂::B200Array{T,1}, d::Integer, sv::Integer, sw::Integer,
                     h::Vector{Float64}, g::Vector{Float64}; add2out::Bool=false) where T<:WxFloat
    usedev(v)
    check(ccall(fsym(:wx_isdwt_step_shift, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Cint, Clong, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}),
                v.ptr, w₁.ptr, w₂.ptr, length(v), d, sv, sw, h, g, length(h), add2out, NOSTREAM))
    return v
end
# acdwt_step!(w1, w2, v, d, h, g)   acwt/acwt_one_level.jl:101-128
function acdwt_step!(w₁::B200Array{T,1}, w₂::B200Array{T,1}, v::B200Array{T,1}, d::Integer, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    @assert length(w₁) == length(w₂) == length(v)
    @assert length(h) == length(g)
    usedev(v)
    check(ccall(fsym(:wx_acdwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                w₁.ptr, w₂.ptr, v.ptr, length(v), d, h, g, length(h), NOSTREAM))
    return w₁, w₂
end
# iacdwt_step!(v, w1, w2)   acwt/acwt_one_level.jl:217-224
function iacdwt_step!(v::B200Array{T,1}, w₁::B200Array{T,1}, w₂::B200Array{T,1}) where T<:WxFloat
    @assert length(v) == length(w₁) == length(w₂)
    usedev(v)
    check(ccall(fsym(:wx_iacdwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cvoid}), v.ptr, w₁.ptr, w₂.ptr, length(v), NOSTREAM))
    return v
end
# 2-D redundant steps   swt/swt_one_level.jl:334-469, acwt/acwt_one_level.jl:240-322
function rdwt_step2!(ac::Bool, w₁::B200Array{T,2}, w₂, w₃, w₄, v::B200Array{T,2}, d::Integer, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    @assert size(v) == size(w₁) == size(w₂) == size(w₃) == size(w₄)
    usedev(v)
    check(ccall(fsym(:wx_rdwt_step2, T), Cint, (Cint, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Clong, Clong, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                ac, w₁.ptr, w₂.ptr, w₃.ptr, w₄.ptr, v.ptr, size(v, 1), size(v, 2), d, h, g, length(h), NOSTREAM))
    return w₁, w₂, w₃, w₄
end
sdwt_step!(w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2}, v::B200Array{T,2}, d::Integer,
           h::Vector{Float64}, g::Vector{Float64}, temp=nothing) where T<:WxFloat = rdwt_step2!(false, w₁, w₂, w₃, w₄, v, d, h, g)
acdwt_step!(w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2}, v::B200Array{T,2}, d::Integer,
            h::Vector{Float64}, g::Vector{Float64}, temp=nothing) where T<:WxFloat = rdwt_step2!(true, w₁, w₂, w₃, w₄, v, d, h, g)
function irdwt_step2!(mode::Integer, v::B200Array{T,2}, w₁, w₂, w₃, w₄, d::Integer, sv::Integer, sw::Integer, h::Vector{Float64}, g::Vector{Float64}) where T<:WxFloat
    @assert size(v) == size(w₁) == size(w₂) == size(w₃) == size(w₄)
    usedev(v)
    check(ccall(fsym(:wx_irdwt_step2, T), Cint, (Cint, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                mode, v.ptr, w₁.ptr, w₂.ptr, w₃.ptr, w₄.ptr, size(v, 1), size(v, 2), d, sv, sw, h, g, length(h), NOSTREAM))
    return v
end
isdwt_step!(v::B200Array{T,2}, w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2}, d::Integer,
            h::Vector{Float64}, g::Vector{Float64}, temp=nothing) where T<:WxFloat = irdwt_step2!(0, v, w₁, w₂, w₃, w₄, d, 0, 0, h, g)
isdwt_step!(v::B200Array{T,2}, w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2}, d::Integer,
            sv::Integer, sw::Integer, h::Vector{Float64}, g::Vector{Float64}, temp=nothing) where T<:WxFloat =
    irdwt_step2!(1, v, w₁, w₂, w₃, w₄, d, sv, sw, h, g)
iacdwt_step!(v::B200Array{T,2}, w₁::B200Array{T,2}, w₂::B200Array{T,2}, w₃::B200Array{T,2}, w₄::B200Array{T,2}, temp=nothing) where T<:WxFloat =
    irdwt_step2!(2, v, w₁, w₂, w₃, w₄, 0, 0, 0, Float64[0.0], Float64[0.0])
# sidwt_step!(w1, w2, v, h, g, s)   siwt/siwt_one_level.jl:71-98 and isidwt_step!(v, w1, w2, h, g, s) :153-184
function sidwt_step!(w₁::B200Array{T,1}, w₂::B200Array{T,1}, v::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}, s::Bool) where T<:WxFloat
    @assert length(w₁) == length(w₂) == length(v) ÷ 2
    @assert length(h) == length(g)
    usedev(v)
    check(ccall(fsym(:wx_sidwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}),
                w₁.ptr, w₂.ptr, v.ptr, length(v), h, g, length(h), s, NOSTREAM))
    return w₁, w₂
end
function isidwt_step!(v::B200Array{T,1}, w₁::B200Array{T,1}, w₂::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}, s::Bool) where T<:WxFloat
    @assert length(w₁) == length(w₂) == length(v) ÷ 2
    @assert length(h) == length(g)
    usedev(v)
    check(ccall(fsym(:wx_isidwt_step, T), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}),
                v.ptr, w₁.ptr, w₂.ptr, length(v), h, g, length(h), s, NOSTREAM))
    return v
end

# ---- decimated trees: low-level wrappers (N = number of signals; m = 0 for 1-D) ----------------------------------------------------
function _wpd!(y::B200Array{T}, x::B200Array{T}, m::Integer, n::Integer, L::Integer, N::Integer, wt::OrthoFilter) where T<:WxFloat
    g, h = qmfpair(wt)
    usedev(x)
    if m == 0
        check(ccall(fsym(:wx_wpd1d, T), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                    y.ptr, x.ptr, n, L, N, h, g, length(h), NOSTREAM))
    else
        check(ccall(fsym(:wx_wpd2d, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                    y.ptr, x.ptr, m, n, L, N, h, g, length(h), NOSTREAM))
    end
    return y
end
function _wpt!(y::B200Array{T}, x::B200Array{T}, m::Integer, n::Integer, N::Integer, wt::OrthoFilter, tree::BitVector, inverse::Bool) where T<:WxFloat
    g, h = qmfpair(wt)
    t = treebytes(tree)
    usedev(x)
    if m == 0 && !inverse
        check(ccall(fsym(:wx_wpt1d, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                    y.ptr, x.ptr, n, N, t, length(t), h, g, length(h), NOSTREAM))
    elseif m == 0
        check(ccall(fsym(:wx_iwpt1d, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                    y.ptr, x.ptr, n, N, t, length(t), h, g, length(h), NOSTREAM))
    elseif !inverse
        check(ccall(fsym(:wx_wpt2d, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Clong, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                    y.ptr, x.ptr, m, n, N, t, length(t), h, g, length(h), NOSTREAM))
    else
        check(ccall(fsym(:wx_iwpt2d, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Clong, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                    y.ptr, x.ptr, m, n, N, t, length(t), h, g, length(h), NOSTREAM))
    end
    return y
end
function _iwpd!(x::B200Array{T}, Xw::B200Array{T}, m::Integer, n::Integer, K::Integer, N::Integer, wt::OrthoFilter, tree::BitVector) where T<:WxFloat
    g, h = qmfpair(wt)
    t = treebytes(tree)
    usedev(Xw)
    check(ccall(fsym(:wx_iwpd, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                x.ptr, Xw.ptr, m, n, K, N, t, length(t), h, g, length(h), NOSTREAM))
    return x
end
function _gather!(out::B200Array{T}, Xw::B200Array{T}, m::Integer, n::Integer, K::Integer, N::Integer, tree::BitVector) where T<:WxFloat
    t = treebytes(tree)
    usedev(Xw)
    check(ccall(fsym(:wx_gather_basis, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{UInt8}, Clong, Ptr{Cvoid}),
                out.ptr, Xw.ptr, m, n, K, N, t, length(t), NOSTREAM))
    return out
end
fulltree(sz::Tuple, L::Integer, s::Symbol=:full) = maketree(sz..., L, s)
mn(sz::NTuple{1,Int}) = (0, sz[1])              # (m, n) as the C ABI wants them: m = 0 selects the 1-D kernels
mn(sz::NTuple{2,Int}) = (sz[1], sz[2])

# wpd / wpd!   DWT.jl:60-209 (one signal / one image)
function wpd!(y::B200Array{T,2}, x::B200Array{T,1}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat
    @assert 0 ≤ L ≤ maxlev(x)
    @assert size(y) == (length(x), L + 1)
    return _wpd!(y, x, 0, length(x), L, 1, wt)
end
function wpd!(y::B200Array{T,3}, x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat
    @assert 0 ≤ L ≤ maxlev(x)
    @assert size(y) == (size(x)..., L + 1)
    return _wpd!(y, x, size(x, 1), size(x, 2), L, 1, wt)
end
wpd(x::B200Array{T,1}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat = wpd!(similar(x, T, (length(x), L + 1)), x, wt, L)
wpd(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat = wpd!(similar(x, T, (size(x)..., L + 1)), x, wt, L)
# wpdall(x, wt, L)   dwt/dwt_all.jl:260-282: x (n, N) -> (n, L+1, N); images x (m, n, N) -> (m, n, L+1, N)
function wpdall(x::B200Array{T,D}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(x)[1:end-1]))) where {T<:WxFloat,D}
    @assert D > 1
    sz = size(x)[1:end-1]
    @assert 0 ≤ L ≤ maxtransformlevels(minimum(sz))
    N = size(x)[end]
    y = similar(x, T, (sz..., L + 1, N))
    m, n = mn(sz)
    return _wpd!(y, x, m, n, L, N, wt)
end
# iwpd / iwpd!   DWT.jl:257-401
function iwpd!(x̂::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat
    @assert size(x̂) == size(xw)[1:end-1]
    @assert isvalidtree(zeros(T, size(x̂)), tree)
    m, n = mn(size(x̂))
    return _iwpd!(x̂, xw, m, n, size(xw)[end], 1, wt, tree)
end
iwpd!(x̂::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxlev(x̂)) where T<:WxFloat = iwpd!(x̂, xw, wt, fulltree(size(x̂), L))
iwpd(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat = iwpd!(similar(xw, T, size(xw)[1:end-1]), xw, wt, tree)
iwpd(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(xw)[1:end-1]))) where T<:WxFloat =
    iwpd(xw, wt, fulltree(size(xw)[1:end-1], L))
# iwpdall(xw, wt[, L | tree])   dwt/dwt_all.jl:324-342
function iwpdall(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat
    @assert ndims(xw) > 2
    sz = size(xw)[1:end-2]
    N = size(xw)[end]
    m, n = mn(sz)
    return _iwpd!(similar(xw, T, (sz..., N)), xw, m, n, size(xw)[end-1], N, wt, tree)
end
iwpdall(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(xw)[1:end-2]))) where T<:WxFloat =
    iwpdall(xw, wt, fulltree(size(xw)[1:end-2], L))
# wpt / wpt! / iwpt / iwpt! by tree (1-D: Wavelets.jl's functions; 2-D: DWT.jl:440-710)
function wpt!(y::B200Array{T,D}, x::B200Array{T,D}, wt::OrthoFilter, tree::BitVector) where {T<:WxFloat,D}
    @assert size(y) == size(x)
    @assert isvalidtree(zeros(T, size(x)), tree)
    m, n = mn(size(x))
    return _wpt!(y, x, m, n, 1, wt, tree, false)
end
wpt!(y::B200Array{T,D}, x::B200Array{T,D}, wt::OrthoFilter, L::Integer=maxlev(x)) where {T<:WxFloat,D} = wpt!(y, x, wt, fulltree(size(x), L))
wpt(x::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat = wpt!(similar(x), x, wt, tree)
wpt(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat = wpt(x, wt, fulltree(size(x), L))
function iwpt!(x̂::B200Array{T,D}, xw::B200Array{T,D}, wt::OrthoFilter, tree::BitVector) where {T<:WxFloat,D}
    @assert size(x̂) == size(xw)
    @assert isvalidtree(zeros(T, size(xw)), tree)
    m, n = mn(size(xw))
    return _wpt!(x̂, xw, m, n, 1, wt, tree, true)
end
iwpt!(x̂::B200Array{T,D}, xw::B200Array{T,D}, wt::OrthoFilter, L::Integer=maxlev(xw)) where {T<:WxFloat,D} = iwpt!(x̂, xw, wt, fulltree(size(xw), L))
iwpt(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat = iwpt!(similar(xw), xw, wt, tree)
iwpt(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxlev(xw)) where T<:WxFloat = iwpt(xw, wt, fulltree(size(xw), L))
# wptall / iwptall / dwtall / idwtall   dwt/dwt_all.jl:39-225 (batch = last dimension)
function wptall(x::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat
    @assert ndims(x) > 1
    m, n = mn(size(x)[1:end-1])
    return _wpt!(similar(x), x, m, n, size(x)[end], wt, tree, false)
end
wptall(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(x)[1:end-1]))) where T<:WxFloat =
    wptall(x, wt, fulltree(size(x)[1:end-1], L))
function iwptall(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat
    @assert ndims(xw) > 1
    m, n = mn(size(xw)[1:end-1])
    return _wpt!(similar(xw), xw, m, n, size(xw)[end], wt, tree, true)
end
iwptall(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(xw)[1:end-1]))) where T<:WxFloat =
    iwptall(xw, wt, fulltree(size(xw)[1:end-1], L))
dwtall(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(x)[1:end-1]))) where T<:WxFloat =
    wptall(x, wt, fulltree(size(x)[1:end-1], L, :dwt))
idwtall(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(xw)[1:end-1]))) where T<:WxFloat =
    iwptall(xw, wt, fulltree(size(xw)[1:end-1], L, :dwt))
# getbasiscoef / getbasiscoefall   Utils.jl:101-225
function getbasiscoef(Xw::B200Array{T}, tree::BitVector) where T<:WxFloat
    @assert 2 ≤ ndims(Xw) ≤ 3
    sz = size(Xw)[1:end-1]
    @assert isvalidtree(zeros(T, sz), tree)
    m, n = mn(sz)
    return _gather!(similar(Xw, T, sz), Xw, m, n, size(Xw)[end], 1, tree)
end
function getbasiscoefall(Xw::B200Array{T}, tree::BitVector) where T<:WxFloat
    @assert 3 ≤ ndims(Xw) ≤ 4
    sz = size(Xw)[1:end-2]
    @assert isvalidtree(zeros(T, sz), tree)
    @assert length(tree) == gettreelength(sz...)
    m, n = mn(sz)
    N = size(Xw)[end]
    return _gather!(similar(Xw, T, (sz..., N)), Xw, m, n, size(Xw)[end-1], N, tree)
end
function getbasiscoefall(Xw::B200Array{T}, tree::BitArray{2}) where T<:WxFloat      # one tree per signal (Utils.jl:199-225)
    @assert 3 ≤ ndims(Xw) ≤ 4
    sz = size(Xw)[1:end-2]
    N = size(Xw)[end]
    nₜ, mₜ = size(tree)
    @assert N == mₜ
    @assert nₜ == gettreelength(sz...)
    @assert all(mapslices(tᵢ -> isvalidtree(zeros(T, sz), BitVector(tᵢ)), tree, dims=1))
    m, n = mn(sz)
    td = B200Array(UInt8.(tree); dev=Xw.dev)
    out = similar(Xw, T, (sz..., N))
    check(ccall(fsym(:wx_gather_basis_multi, T), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{UInt8}, Clong, Ptr{Cvoid}),
                out.ptr, Xw.ptr, m, n, size(Xw)[end-1], N, td.ptr, nₜ, NOSTREAM))
    return out
end

# ---- redundant trees: stationary (ac = false) and autocorrelation (ac = true) ---------------------------------------------------------
const MODE_DWT, MODE_WPT, MODE_WPD = 0, 1, 2
ncolumns(mode::Integer, L::Integer, two::Bool) = mode == MODE_DWT ? (two ? 3L + 1 : L + 1) : mode == MODE_WPT ? (two ? 4^L : 1 << L) :
                                                  (two ? (4^(L + 1) - 1) ÷ 3 : (1 << (L + 1)) - 1)
function check_levels(sz::Tuple, L::Integer)                     # SWT.jl:114-116 (ArgumentError stays a Julia-side check)
    L ≤ maxtransformlevels(minimum(sz)) || throw(ArgumentError("Too many transform levels (length(x) < 2^L"))
    L ≥ 1 || throw(ArgumentError("L must be ≥ 1"))
end
function _rwt!(ac::Bool, mode::Integer, xw::B200Array{T}, x::B200Array{T}, sz::Tuple, L::Integer, N::Integer, wt::OrthoFilter) where T<:WxFloat
    check_levels(sz, L)
    @assert size(xw)[1:length(sz)+1] == (sz..., ncolumns(mode, L, length(sz) == 2))
    h, g = ac ? acpair(wt) : reverse(qmfpair(wt))
    m, n = mn(sz)
    usedev(x)
    check(ccall(fsym(:wx_rwt, T), Cint, (Cint, Cint, Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                ac, mode, xw.ptr, x.ptr, m, n, L, N, h, g, length(h), NOSTREAM))
    return xw
end
# sm < 0: average based (stationary) ; tree only for MODE_WPD
function _irwt!(ac::Bool, mode::Integer, x::B200Array{T}, xw::B200Array{T}, sz::Tuple, N::Integer, wt::Union{OrthoFilter,Nothing}, tree::BitVector, sm::Integer) where T<:WxFloat
    h, g = ac ? (Float64[0.0], Float64[0.0]) : reverse(qmfpair(wt))
    t = treebytes(tree)
    m, n = mn(sz)
    usedev(xw)
    check(ccall(fsym(:wx_irwt, T), Cint, (Cint, Cint, Ptr{T}, Ptr{T}, Clong, Clong, Clong, Cint, Clong, Ptr{UInt8}, Clong, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                ac, mode, x.ptr, xw.ptr, m, n, size(xw)[length(sz)+1], 0, N, t, length(t), sm, h, g, length(h), NOSTREAM))
    return x
end
const NOTREE = BitVector()
# forward, one signal   SWT.jl:60-902, ACWT.jl:60-793
for (fn, fn!, ac, mode) in ((:sdwt, :sdwt!, false, MODE_DWT), (:swpt, :swpt!, false, MODE_WPT), (:swpd, :swpd!, false, MODE_WPD),
                            (:acdwt, :acdwt!, true, MODE_DWT), (:acwpt, :acwpt!, true, MODE_WPT), (:acwpd, :acwpd!, true, MODE_WPD))
    @eval begin
        $fn!(xw::B200Array{T}, x::B200Array{T}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat = _rwt!($ac, $mode, xw, x, size(x), L, 1, wt)
        function $fn(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat
            @assert 1 ≤ ndims(x) ≤ 2
            check_levels(size(x), L)
            return $fn!(similar(x, T, (size(x)..., ncolumns($mode, L, ndims(x) == 2))), x, wt, L)
        end
    end
end
# forward, batch   swt/swt_all.jl:33-296, acwt/acwt_all.jl:33-256
for (fn, ac, mode) in ((:sdwtall, false, MODE_DWT), (:swptall, false, MODE_WPT), (:swpdall, false, MODE_WPD),
                       (:acdwtall, true, MODE_DWT), (:acwptall, true, MODE_WPT), (:acwpdall, true, MODE_WPD))
    @eval function $fn(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(size(x)[1:end-1]))) where T<:WxFloat
        @assert 2 ≤ ndims(x) ≤ 3
        sz = size(x)[1:end-1]
        check_levels(sz, L)
        N = size(x)[end]
        return _rwt!($ac, $mode, similar(x, T, (sz..., ncolumns($mode, L, length(sz) == 2), N)), x, sz, L, N, wt)
    end
end
# stationary inverses, one signal   SWT.jl:197-358, 551-758, 952-1199 (sm: shift based; without: average based)
isdwt!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, sm::Integer) where T<:WxFloat = _irwt!(false, MODE_DWT, x, xw, size(x), 1, wt, NOTREE, sm)
isdwt!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter) where T<:WxFloat = _irwt!(false, MODE_DWT, x, xw, size(x), 1, wt, NOTREE, -1)
isdwt(xw::B200Array{T}, wt::OrthoFilter, sm::Integer) where T<:WxFloat = isdwt!(similar(xw, T, sigdims(xw, false)), xw, wt, sm)
isdwt(xw::B200Array{T}, wt::OrthoFilter) where T<:WxFloat = isdwt!(similar(xw, T, sigdims(xw, false)), xw, wt)
iswpt!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, sm::Integer) where T<:WxFloat = _irwt!(false, MODE_WPT, x, xw, size(x), 1, wt, NOTREE, sm)
iswpt!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter) where T<:WxFloat = _irwt!(false, MODE_WPT, x, xw, size(x), 1, wt, NOTREE, -1)
iswpt(xw::B200Array{T}, wt::OrthoFilter, sm::Integer) where T<:WxFloat = iswpt!(similar(xw, T, sigdims(xw, false)), xw, wt, sm)
iswpt(xw::B200Array{T}, wt::OrthoFilter) where T<:WxFloat = iswpt!(similar(xw, T, sigdims(xw, false)), xw, wt)
function iswpd!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, tree::BitVector, sm::Integer) where T<:WxFloat
    @assert isvalidtree(zeros(T, size(x)), tree)
    return _irwt!(false, MODE_WPD, x, xw, size(x), 1, wt, tree, sm)
end
function iswpd!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat
    @assert isvalidtree(zeros(T, size(x)), tree)
    return _irwt!(false, MODE_WPD, x, xw, size(x), 1, wt, tree, -1)
end
function iswpd!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, L::Integer, sm::Integer) where T<:WxFloat
    check_levels(size(x), L)
    return iswpd!(x, xw, wt, fulltree(size(x), L), sm)
end
function iswpd!(x::B200Array{T}, xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxlev(x)) where T<:WxFloat
    check_levels(size(x), L)
    return iswpd!(x, xw, wt, fulltree(size(x), L))
end
iswpd(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector, sm::Integer) where T<:WxFloat = iswpd!(similar(xw, T, sigdims(xw, false)), xw, wt, tree, sm)
iswpd(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat = iswpd!(similar(xw, T, sigdims(xw, false)), xw, wt, tree)
iswpd(xw::B200Array{T}, wt::OrthoFilter, L::Integer, sm::Integer) where T<:WxFloat = iswpd(xw, wt, fulltree(sigdims(xw, false), L), sm)
iswpd(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(sigdims(xw, false)))) where T<:WxFloat =
    iswpd(xw, wt, fulltree(sigdims(xw, false), L))
# stationary inverses, batch   swt/swt_all.jl:89-390
function _irwtall(ac::Bool, mode::Integer, xw::B200Array{T}, wt, tree::BitVector, sm::Integer) where T<:WxFloat
    @assert 3 ≤ ndims(xw) ≤ 4
    sz = sigdims(xw, true)
    N = size(xw)[end]
    return _irwt!(ac, mode, similar(xw, T, (sz..., N)), xw, sz, N, wt, tree, sm)
end
isdwtall(xw::B200Array{T}, wt::OrthoFilter) where T<:WxFloat = _irwtall(false, MODE_DWT, xw, wt, NOTREE, -1)
isdwtall(xw::B200Array{T}, wt::OrthoFilter, sm::Integer) where T<:WxFloat = _irwtall(false, MODE_DWT, xw, wt, NOTREE, sm)
iswptall(xw::B200Array{T}, wt::OrthoFilter) where T<:WxFloat = _irwtall(false, MODE_WPT, xw, wt, NOTREE, -1)
iswptall(xw::B200Array{T}, wt::OrthoFilter, sm::Integer) where T<:WxFloat = _irwtall(false, MODE_WPT, xw, wt, NOTREE, sm)
iswpdall(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector) where T<:WxFloat = _irwtall(false, MODE_WPD, xw, wt, tree, -1)
iswpdall(xw::B200Array{T}, wt::OrthoFilter, tree::BitVector, sm::Integer) where T<:WxFloat = _irwtall(false, MODE_WPD, xw, wt, tree, sm)
iswpdall(xw::B200Array{T}, wt::OrthoFilter, L::Integer, sm::Integer) where T<:WxFloat = iswpdall(xw, wt, fulltree(sigdims(xw, true), L), sm)
iswpdall(xw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(minimum(sigdims(xw, true)))) where T<:WxFloat =
    iswpdall(xw, wt, fulltree(sigdims(xw, true), L))
# autocorrelation inverses (plain pairwise sums, no filter)   ACWT.jl:244-1000, acwt/acwt_all.jl:86-335
iacdwt!(x::B200Array{T}, xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing) where T<:WxFloat = _irwt!(true, MODE_DWT, x, xw, size(x), 1, nothing, NOTREE, -1)
iacdwt(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing) where T<:WxFloat = iacdwt!(similar(xw, T, sigdims(xw, false)), xw)
iacwpt!(x::B200Array{T}, xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing) where T<:WxFloat = _irwt!(true, MODE_WPT, x, xw, size(x), 1, nothing, NOTREE, -1)
iacwpt(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing) where T<:WxFloat = iacwpt!(similar(xw, T, sigdims(xw, false)), xw)
function iacwpd!(x::B200Array{T}, xw::B200Array{T}, tree::BitVector) where T<:WxFloat
    @assert isvalidtree(zeros(T, size(x)), tree)
    return _irwt!(true, MODE_WPD, x, xw, size(x), 1, nothing, tree, -1)
end
function iacwpd!(x::B200Array{T}, xw::B200Array{T}, L::Integer) where T<:WxFloat
    check_levels(size(x), L)
    return iacwpd!(x, xw, fulltree(size(x), L))
end
iacwpd!(x::B200Array{T}, xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}, tree::BitVector) where T<:WxFloat = iacwpd!(x, xw, tree)
iacwpd!(x::B200Array{T}, xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing, L::Integer=maxlev(x)) where T<:WxFloat = iacwpd!(x, xw, L)
iacwpd(xw::B200Array{T}, tree::BitVector) where T<:WxFloat = iacwpd!(similar(xw, T, sigdims(xw, false)), xw, tree)
iacwpd(xw::B200Array{T}, L::Integer) where T<:WxFloat = iacwpd!(similar(xw, T, sigdims(xw, false)), xw, L)
iacwpd(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}, tree::BitVector) where T<:WxFloat = iacwpd(xw, tree)
iacwpd(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing, L::Integer=maxtransformlevels(size(xw, 1))) where T<:WxFloat = iacwpd(xw, L)
iacdwtall(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing) where T<:WxFloat = _irwtall(true, MODE_DWT, xw, nothing, NOTREE, -1)
iacwptall(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing) where T<:WxFloat = _irwtall(true, MODE_WPT, xw, nothing, NOTREE, -1)
iacwpdall(xw::B200Array{T}, tree::BitVector) where T<:WxFloat = _irwtall(true, MODE_WPD, xw, nothing, tree, -1)
iacwpdall(xw::B200Array{T}, L::Integer) where T<:WxFloat = iacwpdall(xw, fulltree(sigdims(xw, true), L))
iacwpdall(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}, tree::BitVector) where T<:WxFloat = iacwpdall(xw, tree)
iacwpdall(xw::B200Array{T}, wt::Union{OrthoFilter,Nothing}=nothing, L::Integer=maxtransformlevels(minimum(sigdims(xw, true)))) where T<:WxFloat = iacwpdall(xw, L)

# ---- the collective: NCCL communicators inside libwx_b200 (wx_comm.cu) -----------------------------------------------------------------
# One process (or task) per GPU: rank 0 calls comm_unique_id(), ships the 128 bytes to the others (Distributed.jl, MPI.jl, a
# file), every rank builds B200Comm(id, rank, world) on its device.  One Julia thread driving all GPUs (the reference is single
# threaded): comms = comm_init_all(ndev) and the multi-shard methods of tree_costs / bestbasistree below.
mutable struct B200Comm
    h::Ptr{Cvoid}
end
const NOCOMM = B200Comm(C_NULL)
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall(sym(:wx_comm_unique_id), Cint, (Ptr{UInt8},), id))
    return id
end
function B200Comm(id::Vector{UInt8}, rank::Integer, world::Integer; dev::Integer=rank)
    @assert length(id) == 128
    check(ccall(sym(:wx_set_device), Cint, (Cint,), dev))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(sym(:wx_comm_init_rank), Cint, (Ptr{Ptr{Cvoid}}, Ptr{UInt8}, Cint, Cint), h, id, rank, world))
    return B200Comm(h[])
end
function comm_init_all(ndev::Integer)
    hs = Vector{Ptr{Cvoid}}(undef, ndev)
    check(ccall(sym(:wx_comm_init_all), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}), hs, ndev, C_NULL))
    return [B200Comm(h) for h in hs]
end
function comm_destroy(c::B200Comm)
    check(ccall(sym(:wx_comm_destroy), Cint, (Ptr{Cvoid},), c.h))
    c.h = C_NULL
    return nothing
end
# in-place all-reduce of a device buffer over the communicator (op: :sum, :min, :max)
function allreduce!(c::B200Comm, a::B200Array{T}, op::Symbol=:sum) where T<:WxFloat
    usedev(a)
    check(ccall(sym(:wx_allreduce), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Clong, Cint, Cint, Ptr{Cvoid}),
                c.h, a.ptr, length(a), T === Float64 ? 0 : 1, op === :sum ? 0 : op === :min ? 1 : 2, NOSTREAM))
    check(ccall(sym(:wx_stream_sync), Cint, (Ptr{Cvoid},), NOSTREAM))
    return a
end

# ---- best basis ------------------------------------------------------------------------------------------------------------------------
ncosts(sz::Tuple, K::Integer, redundant::Bool) = redundant ? K : (length(sz) == 2 ? (4^K - 1) ÷ 3 : (1 << K) - 1)
jbbkind(method::JBB) = method.cost isa LoglpCost ? 0 : 1
# tree_costs(X, ::JBB | ::LSDB)   bestbasis/bestbasis_tree.jl:104-207.  X is the LOCAL shard (…, K, N_local); with a communicator
# the costs are those of the global batch (the reductions over the batch dimension run through NCCL inside the library).
function tree_costs(X::B200Array{T,D}, method::JBB; comm::B200Comm=NOCOMM) where {T<:WxFloat,D}
    @assert 3 ≤ D ≤ 4
    sz = size(X)[1:end-2]
    K, N = size(X)[end-1], size(X)[end]
    m, n = mn(sz)
    costs = Vector{Float64}(undef, ncosts(sz, K, method.redundant))
    usedev(X)
    check(ccall(fsym(:wx_tree_costs_jbb, T), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{T}, Clong, Clong, Cint, Clong, Cint, Cint, Cdouble, Ptr{Cvoid}),
                comm.h, costs, X.ptr, m, n, K, N, method.redundant, jbbkind(method), Float64(method.cost.p), NOSTREAM))
    return T.(costs)
end
function tree_costs(X::B200Array{T,D}, method::LSDB; comm::B200Comm=NOCOMM) where {T<:WxFloat,D}
    @assert 3 ≤ D ≤ 4
    sz = size(X)[1:end-2]
    K, N = size(X)[end-1], size(X)[end]
    m, n = mn(sz)
    costs = Vector{Float64}(undef, ncosts(sz, K, method.redundant))
    usedev(X)
    check(ccall(fsym(:wx_tree_costs_lsdb, T), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{T}, Clong, Clong, Cint, Clong, Cint, Ptr{Cvoid}),
                comm.h, costs, X.ptr, m, n, K, N, method.redundant, NOSTREAM))
    return T.(costs)
end
# tree_costs(X, ::BB)   bestbasis/bestbasis_tree.jl:210-256 for ONE decomposed signal (…, K)
function bbcosts(X::B200Array{T}, sz::Tuple, K::Integer, N::Integer, method::BB) where T<:WxFloat
    m, n = mn(sz)
    costs = B200Array{Float64,2}((ncosts(sz, K, method.redundant), N); dev=X.dev)
    usedev(X)
    check(ccall(fsym(:wx_bb_costs, T), Cint, (Ptr{Cdouble}, Ptr{T}, Clong, Clong, Cint, Clong, Cint, Cint, Ptr{Cvoid}),
                costs.ptr, X.ptr, m, n, K, N, method.redundant, method.cost isa ShannonEntropyCost ? 0 : 1, NOSTREAM))
    return costs
end
function tree_costs(X::B200Array{T,D}, method::BB) where {T<:WxFloat,D}
    @assert 2 ≤ D ≤ 3
    return T.(vec(Array(bbcosts(X, size(X)[1:end-1], size(X)[end], 1, method))))
end
# bestbasistree(X, ::JBB | ::LSDB)   BestBasis.jl:185-201 in ONE library call (reduction kernels, NCCL exchange, costs, selection)
function bestbasistree_b200(X::B200Array{T,D}, methodcode::Integer, redundant::Bool, kind::Integer, p::Real, comm::B200Comm) where {T<:WxFloat,D}
    @assert 3 ≤ D ≤ 4
    sz = size(X)[1:end-2]
    K, N = size(X)[end-1], size(X)[end]
    m, n = mn(sz)
    tree = Vector{UInt8}(undef, gettreelength(sz...))
    usedev(X)
    check(ccall(fsym(:wx_bestbasistree, T), Cint,
                (Ptr{Cvoid}, Cint, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{T}, Clong, Clong, Cint, Clong, Cint, Cint, Cdouble, Ptr{Cvoid}),
                comm.h, methodcode, tree, length(tree), C_NULL, X.ptr, m, n, K, N, redundant, kind, Float64(p), NOSTREAM))
    return BitVector(tree .!= 0)
end
bestbasistree(X::B200Array{T}, method::JBB; comm::B200Comm=NOCOMM) where T<:WxFloat =
    bestbasistree_b200(X, 0, method.redundant, jbbkind(method), method.cost.p, comm)
bestbasistree(X::B200Array{T}, method::LSDB; comm::B200Comm=NOCOMM) where T<:WxFloat = bestbasistree_b200(X, 1, method.redundant, 0, 0.0, comm)
bestbasistree(X::B200Array{T}, method::BB) where T<:WxFloat = vec(bestbasistreeall(reshape_batch1(X), method))
reshape_batch1(X::B200Array{T,D}) where {T,D} = B200Array{T,D + 1}(X, (size(X)..., 1))
# the same from ONE Julia thread driving several GPUs: shards[i] lives on comms[i]'s device (comm_init_all)
function bestbasistree(shards::Vector{B200Array{T,D}}, method::Union{JBB,LSDB}, comms::Vector{B200Comm}) where {T<:WxFloat,D}
    @assert length(shards) == length(comms) ≥ 1
    sz = size(shards[1])[1:end-2]
    K = size(shards[1])[end-1]
    m, n = mn(sz)
    tree = Vector{UInt8}(undef, gettreelength(sz...))
    hs = [c.h for c in comms]
    ptrs = [Ptr{Cvoid}(s.ptr) for s in shards]
    Ns = Clong[size(s)[end] for s in shards]
    isj = method isa JBB
    check(ccall(fsym(:wx_bestbasistree_multi, T), Cint,
                (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}, Ptr{Clong}, Clong, Clong, Cint, Cint, Cint, Cdouble),
                hs, length(hs), isj ? 0 : 1, tree, length(tree), C_NULL, ptrs, Ns, m, n, K, method.redundant, isj ? jbbkind(method) : 0,
                isj ? Float64(method.cost.p) : 0.0))
    return BitVector(tree .!= 0)
end
# bestbasistreeall(X, ::BB)   BestBasis.jl:253-262 : per-signal costs and the bottom-up selection both on the device;
# returns the BitMatrix (nₜ, N) of the reference
function bestbasistreeall(X::B200Array{T,D}, method::BB) where {T<:WxFloat,D}
    @assert 3 ≤ D ≤ 4
    sz = size(X)[1:end-2]
    K, N = size(X)[end-1], size(X)[end]
    m, n = mn(sz)
    costs = bbcosts(X, sz, K, N, method)
    nₜ = gettreelength(sz...)
    trees = B200Array{UInt8,2}((nₜ, N); dev=X.dev)
    check(ccall(sym(:wx_bb_select), Cint, (Ptr{UInt8}, Ptr{Cdouble}, Clong, Clong, Clong, Clong, Cint, Ptr{Cvoid}),
                trees.ptr, costs.ptr, size(costs, 1), m, n, N, sizeof(T), NOSTREAM))
    return BitMatrix(Array(trees) .!= 0)
end

# ---- host arrays stay host arrays ---------------------------------------------------------------------------------------------------------
# wpdall for Array inputs: the library streams the batch through the GPU with overlapped copies   (wx_wpdall_host_*)
function wpdall_host(x::Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T<:WxFloat
    n, N = size(x)
    g, h = qmfpair(wt)
    y = Array{T,3}(undef, (n, L + 1, N))
    check(ccall(fsym(:wx_wpdall_host, T), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Clong), y, x, n, L, N, h, g, length(h), 0))
    return y
end
# the paper's pipeline (paper/paper.md:60-118) in one call: wpdall -> bestbasistree -> getbasiscoefall; the packet table stays in HBM
function wpd_bestbasis_host(x::Array{T,2}, wt::OrthoFilter, method::Union{JBB,LSDB}=JBB(), L::Integer=maxtransformlevels(size(x, 1));
                            comm::B200Comm=NOCOMM) where T<:WxFloat
    n, N = size(x)
    g, h = qmfpair(wt)
    coef = Array{T,2}(undef, (n, N))
    tree = Vector{UInt8}(undef, n - 1)
    isj = method isa JBB
    check(ccall(fsym(:wx_wpd_bestbasis_host, T), Cint,
                (Ptr{Cvoid}, Ptr{T}, Ptr{UInt8}, Clong, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint, Cdouble, Clong),
                comm.h, coef, tree, n - 1, x, n, L, N, h, g, length(h), isj ? 0 : 1, isj ? jbbkind(method) : 0, isj ? Float64(method.cost.p) : 0.0, 0))
    return coef, BitVector(tree .!= 0)
end

# ---- "next" rows: nonstandard-form transform and the denoising estimators ------------------------------------------------------------------
# ns_dwt(x, wt, L)   wavemult/transforms.jl:52-74 for a vector (n,) or a batch (n, N) of vectors -> (2n[, N])
function ns_dwt(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T<:WxFloat
    n = size(x, 1); N = ndims(x) == 1 ? 1 : size(x, 2)
    @assert 1 ≤ L ≤ maxtransformlevels(n)
    @assert ispow2(n)
    g, h = qmfpair(wt)
    nxw = similar(x, T, ndims(x) == 1 ? (2n,) : (2n, N))
    usedev(x)
    check(ccall(fsym(:wx_ns_dwt, T), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                nxw.ptr, x.ptr, n, L, N, h, g, length(h), NOSTREAM))
    return nxw
end
# ns_idwt(nxw, wt, L)   wavemult/transforms.jl:120-139
function ns_idwt(nxw::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(nxw, 1)) - 1) where T<:WxFloat
    n2 = size(nxw, 1); N = ndims(nxw) == 1 ? 1 : size(nxw, 2)
    @assert 1 ≤ L ≤ maxtransformlevels(n2) - 1
    @assert ispow2(n2 ÷ 2)
    g, h = qmfpair(wt)
    x = similar(nxw, T, ndims(nxw) == 1 ? (n2 ÷ 2,) : (n2 ÷ 2, N))
    usedev(nxw)
    check(ccall(fsym(:wx_ns_idwt, T), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                x.ptr, nxw.ptr, n2, L, N, h, g, length(h), NOSTREAM))
    return x
end
# noisest(x, false)   Denoising.jl:214-232 for every column of a batch of dwt coefficients (n, N): N noise levels on the device
function noisestall(X::B200Array{T,2}) where T<:WxFloat
    n, N = size(X)
    sigma = B200Array{Float64,1}((N,); dev=X.dev)
    usedev(X)
    check(ccall(fsym(:wx_noisest, T), Cint, (Ptr{Cdouble}, Ptr{T}, Clong, Clong, Clong, Clong, Ptr{Cvoid}),
                sigma.ptr, X.ptr, n, n ÷ 2, n - n ÷ 2, N, NOSTREAM))
    return sigma
end
# denoiseall(x, :dwt, wt; L, dnt, smooth)   Denoising.jl:651-713 with estnoise = noisest, bestTH = nothing: per-signal thresholds
# sigma_k * dnt.t applied on the device, then idwtall
function denoiseall_dwt(X::B200Array{T,2}, wt::OrthoFilter; L::Integer=maxtransformlevels(size(X, 1)),
                        dnt=Wavelets.Threshold.VisuShrink(size(X, 1)), smooth::Symbol=:regular) where T<:WxFloat
    n, N = size(X)
    sigma = noisestall(X)
    th = dnt.th isa Wavelets.Threshold.HardTH ? 0 : dnt.th isa Wavelets.Threshold.SoftTH ? 1 :
         dnt.th isa Wavelets.Threshold.SemiSoftTH ? 2 : 3
    keep_hi = smooth == :undersmooth ? nodelength(n, L) : 0
    Xt = similar(X)
    check(ccall(fsym(:wx_threshold, T), Cint,
                (Ptr{T}, Ptr{T}, Clong, Clong, Ptr{UInt8}, Clong, Clong, Cint, Ptr{Cdouble}, Cdouble, Clong, Ptr{Cvoid}),
                Xt.ptr, X.ptr, n, 1, C_NULL, 0, keep_hi, th, sigma.ptr, Float64(dnt.t), N, NOSTREAM))
    return idwtall(Xt, wt, L)
end

end # module
