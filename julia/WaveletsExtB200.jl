# WaveletsExtB200.jl -- the reference-side binding a WaveletsExt.jl maintainer would add: a minimal device-array type and
# methods with the reference's own names that `ccall` libwx_b200.so (C ABI: include/wx_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not available in the build image or on the GPU box.  The same ABI is
# exercised 1:1 by the Python host mirror (waveletsext.jl_b200/) in tests/.  See INTEGRATION.md.
module WaveletsExtB200

using Wavelets
import Wavelets: WT
import WaveletsExt
import WaveletsExt.DWT: wpd, wpd!, wpdall, iwpdall, wptall, iwptall, dwt_step!, idwt_step!
import WaveletsExt.SWT: sdwt_step!, swpd, swpd!, swpdall, iswpdall, sdwtall, swptall, isdwtall, iswptall
import WaveletsExt.ACWT: acdwt_step!, acwpd!, acwpdall, iacwpdall, acdwtall, acwptall, iacdwtall, iacwptall, make_acreverseqmfpair
import WaveletsExt.BestBasis: tree_costs, bestbasis_treeselection, JBB, LSDB, LoglpCost, NormCost
import WaveletsExt.Utils: getbasiscoefall

const LIB = get(ENV, "WX_B200_LIB", joinpath(@__DIR__, "..", "waveletsext.jl_b200", "libwx_b200.so"))

struct WxError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:wx_last_error, LIB), Cstring, ()))
    rc == 1 && throw(AssertionError(msg))          # WX_EINVAL: what the reference raises with @assert
    rc == 4 && throw(OutOfMemoryError())
    throw(WxError(rc, msg))
end

# ---- minimal device array ---------------------------------------------------------------------------------------------
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{T}
    dims::NTuple{N,Int}
    dev::Int
    function B200Array{T,N}(dims::NTuple{N,Int}; dev::Int=0) where {T,N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:wx_set_device, LIB), Cint, (Cint,), dev))
        check(ccall((:wx_malloc, LIB), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), p, prod(dims) * sizeof(T)))
        a = new{T,N}(Ptr{T}(p[]), dims, dev)
        finalizer(x -> ccall((:wx_free, LIB), Cint, (Ptr{Cvoid},), x.ptr), a)
        return a
    end
end
Base.size(a::B200Array) = a.dims
Base.similar(a::B200Array{T}, ::Type{T}, dims::Dims{N}) where {T,N} = B200Array{T,N}(dims; dev=a.dev)
Base.getindex(::B200Array, i...) = error("B200Array: scalar indexing is not supported; copy to the host with Array(a)")

function B200Array(x::Array{T,N}; dev::Int=0) where {T<:Union{Float32,Float64},N}
    a = B200Array{T,N}(size(x); dev=dev)
    check(ccall((:wx_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), a.ptr, x, sizeof(x), C_NULL))
    check(ccall((:wx_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return a
end
function Base.Array(a::B200Array{T,N}) where {T,N}
    x = Array{T,N}(undef, a.dims)
    check(ccall((:wx_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), x, a.ptr, sizeof(x), C_NULL))
    check(ccall((:wx_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    return x
end

sfx(::Type{Float64}) = "f64"
sfx(::Type{Float32}) = "f32"
treebytes(t::BitVector) = UInt8.(t)

# ---- decimated -----------------------------------------------------------------------------------------------------------
# wpdall(x, wt, L)   dwt/dwt_all.jl:260-282
function wpdall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T
    @assert 0 ≤ L ≤ maxtransformlevels(size(x, 1))
    n, N = size(x)
    g, h = WT.makereverseqmfpair(wt, true)                 # g = scaling, h = detail, Float64 (DWT.jl:141)
    y = B200Array{T,3}((n, L + 1, N); dev=x.dev)
    f = T === Float64 ? :wx_wpd1d_f64 : :wx_wpd1d_f32
    check(ccall((f, LIB), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                y.ptr, x.ptr, n, L, N, h, g, length(h), C_NULL))
    return y
end
# 2-D images x(m,n,N)   DWT.jl:164-209
function wpdall(x::B200Array{T,3}, wt::OrthoFilter, L::Integer=maxtransformlevels(min(size(x, 1), size(x, 2)))) where T
    m, n, N = size(x)
    g, h = WT.makereverseqmfpair(wt, true)
    y = B200Array{T,4}((m, n, L + 1, N); dev=x.dev)
    f = T === Float64 ? :wx_wpd2d_f64 : :wx_wpd2d_f32
    check(ccall((f, LIB), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                y.ptr, x.ptr, m, n, L, N, h, g, length(h), C_NULL))
    return y
end
# host arrays stay host arrays: the library streams the batch through the GPU   (wx_wpdall_host_*)
function wpdall_b200(x::Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T<:Union{Float32,Float64}
    n, N = size(x)
    g, h = WT.makereverseqmfpair(wt, true)
    y = Array{T,3}(undef, (n, L + 1, N))
    f = T === Float64 ? :wx_wpdall_host_f64 : :wx_wpdall_host_f32
    check(ccall((f, LIB), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Clong), y, x, n, L, N, h, g, length(h), 0))
    return y
end

# iwptall(xw, wt, tree)   dwt/dwt_all.jl:210-225
function iwptall(xw::B200Array{T,2}, wt::OrthoFilter, tree::BitVector=maketree(size(xw, 1), maxtransformlevels(size(xw, 1)), :full)) where T
    n, N = size(xw)
    g, h = WT.makereverseqmfpair(wt, true)
    t = treebytes(tree)
    y = similar(xw)
    f = T === Float64 ? :wx_iwpt1d_f64 : :wx_iwpt1d_f32
    check(ccall((f, LIB), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Ptr{UInt8}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                y.ptr, xw.ptr, n, N, t, length(t), h, g, length(h), C_NULL))
    return y
end
iwptall(xw::B200Array{T,2}, wt::OrthoFilter, L::Integer) where T = iwptall(xw, wt, maketree(size(xw, 1), L, :full))

# getbasiscoefall(Xw, tree)   Utils.jl:169-197
function getbasiscoefall(Xw::B200Array{T,3}, tree::BitVector) where T
    n, K, N = size(Xw)
    @assert isvalidtree(zeros(n), tree)
    t = treebytes(tree)
    out = B200Array{T,2}((n, N); dev=Xw.dev)
    f = T === Float64 ? :wx_gather_basis_f64 : :wx_gather_basis_f32
    check(ccall((f, LIB), Cint, (Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{UInt8}, Clong, Ptr{Cvoid}), out.ptr, Xw.ptr, 0, n, K, N, t, length(t), C_NULL))
    return out
end

# dwt_step!(w1, w2, v, h, g)   dwt/dwt_one_level.jl:79-107
function dwt_step!(w₁::B200Array{T,1}, w₂::B200Array{T,1}, v::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}) where T
    @assert length(w₁) == length(w₂) == length(v) ÷ 2
    @assert length(h) == length(g)
    f = T === Float64 ? :wx_dwt_step_f64 : :wx_dwt_step_f32
    check(ccall((f, LIB), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}), w₁.ptr, w₂.ptr, v.ptr, length(v), h, g, length(h), C_NULL))
    return w₁, w₂
end

# ---- redundant ------------------------------------------------------------------------------------------------------------
# swpdall / acwpdall   swt/swt_all.jl:279-296, acwt/acwt_all.jl:239-256     (mode 2 = wpd, 1 = wpt, 0 = dwt)
function rwtall(ac::Bool, mode::Integer, x::B200Array{T,2}, wt::OrthoFilter, L::Integer) where T
    n, N = size(x)
    L ≤ maxtransformlevels(n) || throw(ArgumentError("Too many transform levels (length(x) < 2^L"))
    L ≥ 1 || throw(ArgumentError("L must be >= 1"))
    if ac
        Pmf, Qmf = make_acreverseqmfpair(wt); h, g = Qmf, Pmf           # acdwt_step!(w1, w2, v, d, Qmf, Pmf)  ACWT.jl:756
    else
        g, h = WT.makereverseqmfpair(wt, true)
    end
    ncol = mode == 2 ? (1 << (L + 1)) - 1 : (mode == 1 ? 1 << L : L + 1)
    xw = B200Array{T,3}((n, ncol, N); dev=x.dev)
    f = T === Float64 ? :wx_rwt_f64 : :wx_rwt_f32
    check(ccall((f, LIB), Cint, (Cint, Cint, Ptr{T}, Ptr{T}, Clong, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                ac, mode, xw.ptr, x.ptr, 0, n, L, N, h, g, length(h), C_NULL))
    return xw
end
swpdall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T = rwtall(false, 2, x, wt, L)
swptall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T = rwtall(false, 1, x, wt, L)
sdwtall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T = rwtall(false, 0, x, wt, L)
acwpdall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T = rwtall(true, 2, x, wt, L)
acwptall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T = rwtall(true, 1, x, wt, L)
acdwtall(x::B200Array{T,2}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T = rwtall(true, 0, x, wt, L)

# iswpdall(xw, wt, tree[, sm])   swt/swt_all.jl:343-390
function iswpdall(xw::B200Array{T,3}, wt::OrthoFilter, tree::BitVector, sm::Integer=-1) where T
    n, ncol, N = size(xw)
    g, h = WT.makereverseqmfpair(wt, true)
    t = treebytes(tree)
    x = B200Array{T,2}((n, N); dev=xw.dev)
    f = T === Float64 ? :wx_irwt_f64 : :wx_irwt_f32
    check(ccall((f, LIB), Cint, (Cint, Cint, Ptr{T}, Ptr{T}, Clong, Clong, Clong, Cint, Clong, Ptr{UInt8}, Clong, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                0, 2, x.ptr, xw.ptr, 0, n, ncol, 0, N, t, length(t), sm, h, g, length(h), C_NULL))
    return x
end

# ---- best basis -------------------------------------------------------------------------------------------------------------
# tree_costs(X, ::JBB)   bestbasis/bestbasis_tree.jl:150-180.  `allreduce!` is the hook for the multi-GPU driver: it must sum
# the 2*n*K moment buffer over ranks (NCCL); single GPU: identity.
function tree_costs(X::B200Array{T,3}, method::JBB; allreduce!::Function=identity, Ntotal::Integer=size(X, 3)) where T
    n, K, N = size(X)
    mom = B200Array{Float64,2}((n * K, 2); dev=X.dev)
    f = T === Float64 ? :wx_jbb_moments_f64 : :wx_jbb_moments_f32
    check(ccall((f, LIB), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{T}, Clong, Clong, Ptr{Cvoid}), mom.ptr, mom.ptr + 8 * n * K, X.ptr, n * K, N, C_NULL))
    allreduce!(mom)
    ncost = method.redundant ? K : (1 << K) - 1
    costs = Vector{Float64}(undef, ncost)
    kind = method.cost isa LoglpCost ? 0 : 1
    check(ccall((:wx_jbb_costs, LIB), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Clong, Clong, Clong, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cvoid}),
                costs, mom.ptr, mom.ptr + 8 * n * K, Ntotal, 0, n, K, method.redundant, kind, Float64(method.cost.p), sizeof(T), C_NULL))
    return T.(costs)
end

# bestbasistree(X, ::JBB)   BestBasis.jl:194-201 : costs on the device, O(n) selection on the host (the reference's own code)
Wavelets.Threshold.bestbasistree(X::B200Array{T,3}, method::JBB; kw...) where T =
    bestbasis_treeselection(tree_costs(X, method; kw...), size(X, 1))

# tree_costs(X, ::LSDB)   bestbasis/bestbasis_tree.jl:104-124.  Single GPU shown; across ranks all-gather rows 1-4 of `stats` and
# the two rows of `logsum` and combine them with wx_dd_sum, all-reduce rows 5-6 (min / max) and `counts` (INTEGRATION.md section 4).
function tree_costs(X::B200Array{T,3}, method::LSDB) where T
    n, K, N = size(X)
    szK = n * K
    stats = B200Array{Float64,2}((szK, 7); dev=X.dev)
    x0 = Vector{T}(undef, szK)                                  # row 0 of stats = the first signal of the (global) batch, as Float64
    check(ccall((:wx_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), x0, X.ptr, szK * sizeof(T), C_NULL))
    check(ccall((:wx_stream_sync, LIB), Cint, (Ptr{Cvoid},), C_NULL))
    check(ccall((:wx_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), stats.ptr, Float64.(x0), szK * 8, C_NULL))
    sfx_ = sfx(T)
    check(ccall((Symbol("wx_lsdb_pass1_", sfx_), LIB), Cint, (Ptr{Cdouble}, Ptr{T}, Clong, Clong, Ptr{Cvoid}), stats.ptr, X.ptr, szK, N, C_NULL))
    npts = Ref{Clong}(0)
    check(ccall((:wx_lsdb_grid, LIB), Cint, (Clong, Ptr{Clong}, Ptr{Clong}, Ptr{Clong}), N, C_NULL, C_NULL, npts))
    counts = B200Array{Float64,2}((szK, npts[]); dev=X.dev)
    check(ccall((Symbol("wx_lsdb_pass2_", sfx_), LIB), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{T}, Clong, Clong, Clong, Ptr{Cvoid}),
                counts.ptr, stats.ptr, X.ptr, szK, N, N, C_NULL))
    logsum = B200Array{Float64,2}((szK, 2); dev=X.dev)
    check(ccall((Symbol("wx_lsdb_pass3_", sfx_), LIB), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{T}, Clong, Clong, Clong, Ptr{Cvoid}),
                logsum.ptr, counts.ptr, stats.ptr, X.ptr, szK, N, N, C_NULL))
    costs = Vector{Float64}(undef, method.redundant ? K : (1 << K) - 1)
    check(ccall((:wx_lsdb_costs, LIB), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Clong, Clong, Clong, Cint, Cint, Ptr{Cvoid}),
                costs, logsum.ptr, N, 0, n, K, method.redundant, C_NULL))
    return T.(costs)
end

# bestbasistreeall(X, ::BB)   BestBasis.jl:253-262 : per-signal costs and the bottom-up selection both on the device;
# returns the BitMatrix (n-1, N) of the reference
function bestbasistreeall(X::B200Array{T,3}, method::WaveletsExt.BestBasis.BB) where T
    n, K, N = size(X)
    nn = method.redundant ? K : (1 << K) - 1
    costs = B200Array{Float64,2}((nn, N); dev=X.dev)
    kind = method.cost isa WaveletsExt.BestBasis.ShannonEntropyCost ? 0 : 1
    check(ccall((Symbol("wx_bb_costs_", sfx(T)), LIB), Cint, (Ptr{Cdouble}, Ptr{T}, Clong, Clong, Cint, Clong, Cint, Cint, Ptr{Cvoid}),
                costs.ptr, X.ptr, 0, n, K, N, method.redundant, kind, C_NULL))
    trees = B200Array{UInt8,2}((n - 1, N); dev=X.dev)
    check(ccall((:wx_bb_select, LIB), Cint, (Ptr{UInt8}, Ptr{Cdouble}, Clong, Clong, Clong, Clong, Cint, Ptr{Cvoid}),
                trees.ptr, costs.ptr, nn, 0, n, N, sizeof(T), C_NULL))
    return BitMatrix(Array(trees) .!= 0)
end

# noisest(x, false)   Denoising.jl:214-232 for every column of a batch of dwt coefficients (n, N): N noise levels on the device
function noisestall(X::B200Array{T,2}) where T
    n, N = size(X)
    sigma = B200Array{Float64,1}((N,); dev=X.dev)
    check(ccall((Symbol("wx_noisest_", sfx(T)), LIB), Cint, (Ptr{Cdouble}, Ptr{T}, Clong, Clong, Clong, Clong, Ptr{Cvoid}),
                sigma.ptr, X.ptr, n, n ÷ 2, n - n ÷ 2, N, C_NULL))
    return sigma
end

# denoiseall(x, :dwt, wt; L, dnt, smooth)   Denoising.jl:651-713 with estnoise = noisest, bestTH = nothing: per-signal thresholds
# sigma_k * dnt.t applied on the device, then idwtall (the :dwt tree through wx_iwpt1d)
function denoiseall_dwt(X::B200Array{T,2}, wt::OrthoFilter; L::Integer=maxtransformlevels(size(X,1)),
                        dnt=VisuShrink(size(X,1)), smooth::Symbol=:regular) where T
    n, N = size(X)
    sigma = noisestall(X)
    th = dnt.th isa Wavelets.Threshold.HardTH ? 0 : dnt.th isa Wavelets.Threshold.SoftTH ? 1 :
         dnt.th isa Wavelets.Threshold.SemiSoftTH ? 2 : 3
    keep_hi = smooth == :undersmooth ? nodelength(n, L) : 0
    Xt = similar(X)
    check(ccall((Symbol("wx_threshold_", sfx(T)), LIB), Cint,
                (Ptr{T}, Ptr{T}, Clong, Clong, Ptr{UInt8}, Clong, Clong, Cint, Ptr{Cdouble}, Cdouble, Clong, Ptr{Cvoid}),
                Xt.ptr, X.ptr, n, 1, C_NULL, 0, keep_hi, th, sigma.ptr, dnt.t, N, C_NULL))
    return iwptall(Xt, wt, maketree(n, L, :dwt))
end

# sidwt_step!(w1, w2, v, h, g, s)   siwt/siwt_one_level.jl:71-98 and isidwt_step!(v, w1, w2, h, g, s) :153-184
function sidwt_step!(w1::B200Array{T,1}, w2::B200Array{T,1}, v::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}, s::Bool) where T
    @assert length(w1) == length(w2) == length(v) ÷ 2
    @assert length(h) == length(g)
    check(ccall((Symbol("wx_sidwt_step_", sfx(T)), LIB), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}),
                w1.ptr, w2.ptr, v.ptr, length(v), h, g, length(h), s, C_NULL))
    return w1, w2
end
function isidwt_step!(v::B200Array{T,1}, w1::B200Array{T,1}, w2::B200Array{T,1}, h::Vector{Float64}, g::Vector{Float64}, s::Bool) where T
    @assert length(w1) == length(w2) == length(v) ÷ 2
    @assert length(h) == length(g)
    check(ccall((Symbol("wx_isidwt_step_", sfx(T)), LIB), Cint, (Ptr{T}, Ptr{T}, Ptr{T}, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}),
                v.ptr, w1.ptr, w2.ptr, length(v), h, g, length(h), s, C_NULL))
    return v
end

# ns_dwt(x, wt, L)   wavemult/transforms.jl:52-74 for a vector (n,) or a batch (n, N) of vectors -> (2n[, N])
function ns_dwt(x::B200Array{T}, wt::OrthoFilter, L::Integer=maxtransformlevels(size(x, 1))) where T
    n = size(x, 1); N = ndims(x) == 1 ? 1 : size(x, 2)
    @assert 1 ≤ L ≤ maxtransformlevels(n)
    @assert ispow2(n)
    g, h = WT.makereverseqmfpair(wt, true)
    nxw = B200Array{T,ndims(x)}(ndims(x) == 1 ? (2n,) : (2n, N); dev=x.dev)
    check(ccall((Symbol("wx_ns_dwt_", sfx(T)), LIB), Cint, (Ptr{T}, Ptr{T}, Clong, Cint, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
                nxw.ptr, x.ptr, n, L, N, h, g, length(h), C_NULL))
    return nxw
end

end # module
