# DiffPointRasterisationB200Ext.jl
#
# Drop-in replacement for the reference's ext/DiffPointRasterisationCUDAExt.jl: the same method signatures, but
# the kernels live in libdpr.so (hand-written CUDA for sm_100a, C ABI in include/dpr.h) and are reached by `ccall`.
# Everything above the canonical-form methods (src/interface.jl, the ChainRules rrule) is untouched.
#
# Wiring (see INTEGRATION.md): put this file in ext/, point the `DiffPointRasterisationCUDAExt = "CUDA"` entry of
# Project.toml [extensions] at it (or rename the module), and set ENV["DPR_B200_LIB"] to the path of libdpr.so.
#
# NOTE: Julia is not installed in the build image, so this file has not been executed there; the identical symbols
# are exercised through Python `ctypes` by tests/ (tests/test_gpu_parity.py), argument for argument.
module DiffPointRasterisationB200Ext

using DiffPointRasterisation, CUDA
using ArgCheck
using FillArrays
using StaticArrays

const libdpr = get(ENV, "DPR_B200_LIB", "libdpr.so")

# reference: ext/DiffPointRasterisationCUDAExt.jl:15-17
const CuOrFillArray{T,N} = Union{CuArray{T,N},FillArrays.AbstractFill{T,N}}
const CuOrFillVector{T} = CuOrFillArray{T,1}

const DPR_OP_FORWARD = Cint(0)
const DPR_OP_PULLBACK = Cint(1)

function _check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dpr_status_string, libdpr), Cstring, (Cint,), rc))
    detail = unsafe_string(ccall((:dpr_last_error_message, libdpr), Cstring, ()))
    return error("libdpr: $msg" * (isempty(detail) ? "" : " [$detail]"))
end

# FillArrays defaults (src/interface.jl:368-394) map to NULL - but only the default of THAT argument: the C ABI reads
# NULL as background = 0 and as weight = 1, so `background = Ones(..)` or `out_weight = Zeros(..)` must not become NULL.
# Every other Fill value is materialised on the device (rare; keeps the semantics of the reference, which honours the
# actual Fill value: ext/DiffPointRasterisationCUDAExt.jl:15-17).
_null(::Type{T}) where {T} = reinterpret(CuPtr{T}, CUDA.CU_NULL)
_fill_on_device(a::FillArrays.AbstractFill{T}) where {T} = CUDA.fill(T(FillArrays.getindex_value(a)), size(a)...)
# background: Zeros -> NULL
_bg(a::CuArray) = a
_bg(a::FillArrays.Zeros) = a
_bg(a::FillArrays.AbstractFill) = _fill_on_device(a)
_bgptr(::Type{T}, a::CuArray{T}) where {T} = pointer(a)
_bgptr(::Type{T}, a::FillArrays.Zeros{T}) where {T} = _null(T)
# out_weight / point_weight: Ones -> NULL
_wt(a::CuArray) = a
_wt(a::FillArrays.Ones) = a
_wt(a::FillArrays.AbstractFill) = _fill_on_device(a)
_wptr(::Type{T}, a::CuArray{T}) where {T} = pointer(a)
_wptr(::Type{T}, a::FillArrays.Ones{T}) where {T} = _null(T)

# One scratch buffer per task, shared by raster! and raster_pullback! and kept between calls: with DPR_OPT_BINNING_CACHE
# (option 8) the pullback that follows a forward on the same inputs - the rrule's call order - reuses the point bins the
# forward left in it.  The option's contract: the first 256 bytes are zero when the buffer is allocated (CUDA.zeros),
# and nobody else writes into it.
const _BINNING_CACHE_ON = Ref(false)
function _workspace(op, n_in, n_out, grid::Vector{Int64}, P, B, ::Type{T}) where {T}
    if !_BINNING_CACHE_ON[]
        ccall((:dpr_set_option, libdpr), Cint, (Cint, Int64), 8, 1)
        _BINNING_CACHE_ON[] = true
    end
    nbytes = ccall((:dpr_workspace_bytes, libdpr), Csize_t,
                   (Cint, Cint, Cint, Ptr{Int64}, Int64, Int64, Cint),
                   op, n_in, n_out, grid, P, B, sizeof(T))
    need = max(Int(nbytes), 256)
    tls = task_local_storage()
    ws = get(tls, :dpr_b200_workspace, nothing)
    if ws === nothing || length(ws) < need || CUDA.device(ws) != CUDA.device()
        ws = CUDA.zeros(UInt8, need)
        tls[:dpr_b200_workspace] = ws
    end
    return ws::CuVector{UInt8}
end

# --------------------------------------------------------------------------------------------------------------
# forward: more specific than the canonical method src/raster.jl:5-13 (same argument list, CuArray storage)
# --------------------------------------------------------------------------------------------------------------
function DiffPointRasterisation.raster!(
    out::CuArray{T,N_out_p1},
    points::CuVector{<:StaticVector{N_in,T}},
    rotation::CuVector{<:StaticMatrix{N_out,N_in,T}},
    translation::CuVector{<:StaticVector{N_out,T}},
    background::CuOrFillVector{T},
    out_weight::CuOrFillVector{T},
    point_weight::CuOrFillVector{T},
) where {T<:Union{Float32,Float64},N_in,N_out,N_out_p1}
    # argument checks of src/raster.jl:14-23, unchanged
    @argcheck N_out == N_out_p1 - 1 DimensionMismatch
    batch_size = size(out, N_out_p1)
    @argcheck batch_size == length(rotation) == length(translation) == length(background) == length(out_weight) DimensionMismatch
    n_points = length(points)
    @argcheck length(point_weight) == n_points

    background, out_weight, point_weight = _bg(background), _wt(out_weight), _wt(point_weight)
    grid = collect(Int64, size(out)[1:N_out])
    ws = _workspace(DPR_OP_FORWARD, N_in, N_out, grid, n_points, batch_size, T)
    GC.@preserve out points rotation translation background out_weight point_weight ws begin
        rc = if T === Float32
            ccall((:dpr_raster_forward_f32, libdpr), Cint,
                  (Cint, Cint, Ptr{Int64}, Int64, Int64, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T},
                   CuPtr{T}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                  N_in, N_out, grid, n_points, batch_size,
                  reinterpret(CuPtr{T}, pointer(points)), reinterpret(CuPtr{T}, pointer(rotation)),
                  reinterpret(CuPtr{T}, pointer(translation)), _bgptr(T, background), _wptr(T, out_weight),
                  _wptr(T, point_weight), pointer(out), reinterpret(CuPtr{Cvoid}, pointer(ws)), sizeof(ws),
                  CUDA.stream().handle)
        else
            ccall((:dpr_raster_forward_f64, libdpr), Cint,
                  (Cint, Cint, Ptr{Int64}, Int64, Int64, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T},
                   CuPtr{T}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                  N_in, N_out, grid, n_points, batch_size,
                  reinterpret(CuPtr{T}, pointer(points)), reinterpret(CuPtr{T}, pointer(rotation)),
                  reinterpret(CuPtr{T}, pointer(translation)), _bgptr(T, background), _wptr(T, out_weight),
                  _wptr(T, point_weight), pointer(out), reinterpret(CuPtr{Cvoid}, pointer(ws)), sizeof(ws),
                  CUDA.stream().handle)
        end
        _check(rc)
    end
    return out
end

# --------------------------------------------------------------------------------------------------------------
# pullback, single image: the reference's extension only has an error stub for it
# (ext/DiffPointRasterisationCUDAExt.jl:213-228).  libdpr handles B = 1 with the batch kernels (the points are split
# over the SMs), so the single-image method forwards to the batch method with a singleton batch dimension and
# unwraps the per-pose results, returning the same NamedTuple as the CPU method (src/raster_pullback.jl:74-81).
# --------------------------------------------------------------------------------------------------------------
function DiffPointRasterisation.raster_pullback!(
    ds_dout::CuArray{T,N_out},
    points::CuVector{<:StaticVector{N_in,T}},
    rotation::StaticMatrix{N_out,N_in,T},
    translation::StaticVector{N_out,T},
    background::Number,
    out_weight::Number,
    point_weight::CuOrFillVector{T},
    ds_dpoints::CuMatrix{T},
    ds_dpoint_weight::CuVector{T};
    kwargs...,
) where {T<:Union{Float32,Float64},N_in,N_out}
    res = DiffPointRasterisation.raster_pullback!(
        reshape(ds_dout, size(ds_dout)..., 1),
        points,
        CuArray([SMatrix{N_out,N_in,T}(rotation)]),
        CuArray([SVector{N_out,T}(translation)]),
        Zeros(T, 1),                    # background does not enter the gradients
        CUDA.fill(T(out_weight), 1),
        point_weight,
        ds_dpoints,
        CuArray{T}(undef, N_out, N_in, 1),
        CuArray{T}(undef, N_out, 1),
        CuArray{T}(undef, 1),
        CuArray{T}(undef, 1),
        ds_dpoint_weight,
    )
    return (;
        points=res.points,
        rotation=SMatrix{N_out,N_in,T}(Array(res.rotation)[:, :, 1]),
        translation=SVector{N_out,T}(Array(res.translation)[:, 1]),
        background=Array(res.background)[1],
        out_weight=Array(res.out_weight)[1],
        point_weight=res.point_weight,
    )
end

# --------------------------------------------------------------------------------------------------------------
# pullback, batch of images: same 13-argument signature as ext/DiffPointRasterisationCUDAExt.jl:231-245
# --------------------------------------------------------------------------------------------------------------
function DiffPointRasterisation.raster_pullback!(
    ds_dout::CuArray{T,N_out_p1},
    points::CuVector{<:StaticVector{N_in,T}},
    rotation::CuVector{<:StaticMatrix{N_out,N_in,T}},
    translation::CuVector{<:StaticVector{N_out,T}},
    background::CuOrFillVector{T},
    out_weight::CuOrFillVector{T},
    point_weight::CuOrFillVector{T},
    ds_dpoints::CuMatrix{T},
    ds_drotation::CuArray{T,3},
    ds_dtranslation::CuMatrix{T},
    ds_dbackground::CuVector{T},
    ds_dout_weight::CuVector{T},
    ds_dpoint_weight::CuVector{T},
) where {T<:Union{Float32,Float64},N_in,N_out,N_out_p1}
    # argument checks of ext/DiffPointRasterisationCUDAExt.jl:247-262, unchanged
    batch_axis = axes(ds_dout, N_out_p1)
    @argcheck N_out == N_out_p1 - 1
    @argcheck batch_axis == axes(rotation, 1) == axes(translation, 1) == axes(background, 1) == axes(out_weight, 1)
    @argcheck batch_axis == axes(ds_drotation, 3) == axes(ds_dtranslation, 2) == axes(ds_dbackground, 1) == axes(ds_dout_weight, 1)
    n_points = length(points)
    @argcheck length(ds_dpoint_weight) == n_points
    @argcheck size(ds_dpoints) == (N_in, n_points)
    batch_size = length(batch_axis)

    out_weight, point_weight = _wt(out_weight), _wt(point_weight)
    grid = collect(Int64, size(ds_dout)[1:N_out])
    ws = _workspace(DPR_OP_PULLBACK, N_in, N_out, grid, n_points, batch_size, T)
    # `ccall` needs its argument types as a literal tuple and takes no splatted arguments, so both element types are
    # written out (the same 20 arguments as dpr_raster_pullback_f32 / _f64 in include/dpr.h)
    GC.@preserve ds_dout points rotation translation out_weight point_weight ds_dpoints ds_drotation ds_dtranslation ds_dbackground ds_dout_weight ds_dpoint_weight ws begin
        rc = if T === Float32
            ccall((:dpr_raster_pullback_f32, libdpr), Cint,
                  (Cint, Cint, Ptr{Int64}, Int64, Int64, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T},
                   CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                  N_in, N_out, grid, n_points, batch_size, pointer(ds_dout),
                  reinterpret(CuPtr{T}, pointer(points)), reinterpret(CuPtr{T}, pointer(rotation)),
                  reinterpret(CuPtr{T}, pointer(translation)), _wptr(T, out_weight), _wptr(T, point_weight),
                  pointer(ds_dpoints), pointer(ds_drotation), pointer(ds_dtranslation), pointer(ds_dbackground),
                  pointer(ds_dout_weight), pointer(ds_dpoint_weight), reinterpret(CuPtr{Cvoid}, pointer(ws)), sizeof(ws),
                  CUDA.stream().handle)
        else
            ccall((:dpr_raster_pullback_f64, libdpr), Cint,
                  (Cint, Cint, Ptr{Int64}, Int64, Int64, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T},
                   CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{T}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                  N_in, N_out, grid, n_points, batch_size, pointer(ds_dout),
                  reinterpret(CuPtr{T}, pointer(points)), reinterpret(CuPtr{T}, pointer(rotation)),
                  reinterpret(CuPtr{T}, pointer(translation)), _wptr(T, out_weight), _wptr(T, point_weight),
                  pointer(ds_dpoints), pointer(ds_drotation), pointer(ds_dtranslation), pointer(ds_dbackground),
                  pointer(ds_dout_weight), pointer(ds_dpoint_weight), reinterpret(CuPtr{Cvoid}, pointer(ws)), sizeof(ws),
                  CUDA.stream().handle)
        end
        _check(rc)
    end
    # same NamedTuple, same order, as ext/DiffPointRasterisationCUDAExt.jl:313-320 (the rrule indexes it positionally)
    return (;
        points=ds_dpoints,
        rotation=ds_drotation,
        translation=ds_dtranslation,
        background=ds_dbackground,
        out_weight=ds_dout_weight,
        point_weight=ds_dpoint_weight,
    )
end

# mixed element types: the reference promotes (src/interface.jl:63-64, ext :246); there is no CPU fallback here, so
# convert once on the device and re-dispatch.
function DiffPointRasterisation.raster_pullback!(
    ds_dout::CuArray{<:Number,N_out_p1},
    points::CuVector{<:StaticVector{N_in,<:Number}},
    rotation::CuVector{<:StaticMatrix{N_out,N_in,<:Number}},
    translation::CuVector{<:StaticVector{N_out,<:Number}},
    background::CuOrFillVector{<:Number},
    out_weight::CuOrFillVector{<:Number},
    point_weight::CuOrFillVector{<:Number},
    ds_dpoints::CuMatrix{TP},
    ds_drotation::CuArray{TR,3},
    ds_dtranslation::CuMatrix{TT},
    ds_dbackground::CuVector{<:Number},
    ds_dout_weight::CuVector{OW},
    ds_dpoint_weight::CuVector{PW},
) where {N_in,N_out,N_out_p1,TP<:Number,TR<:Number,TT<:Number,OW<:Number,PW<:Number}
    T = promote_type(eltype(ds_dout), TP, TR, TT, OW, PW)
    T <: Union{Float32,Float64} || error("libdpr computes in Float32 or Float64; got $T")
    cv(a::CuArray) = CuArray{T}(a)
    cv(a::CuVector{<:StaticVector{N}}) where {N} = CuVector{SVector{N,T}}(a)
    cv(a::CuVector{<:StaticMatrix{M,N}}) where {M,N} = CuVector{SMatrix{M,N,T,M * N}}(a)
    cv(a::FillArrays.Zeros) = Zeros(T, size(a)...)
    cv(a::FillArrays.Ones) = Ones(T, size(a)...)
    cv(a::FillArrays.AbstractFill) = CUDA.fill(T(FillArrays.getindex_value(a)), size(a)...)
    res = DiffPointRasterisation.raster_pullback!(
        cv(ds_dout), cv(points), cv(rotation), cv(translation), cv(background), cv(out_weight), cv(point_weight),
        cv(ds_dpoints), cv(ds_drotation), cv(ds_dtranslation), cv(ds_dbackground), cv(ds_dout_weight), cv(ds_dpoint_weight))
    copyto!(ds_dpoints, res.points); copyto!(ds_drotation, res.rotation); copyto!(ds_dtranslation, res.translation)
    copyto!(ds_dbackground, res.background); copyto!(ds_dout_weight, res.out_weight); copyto!(ds_dpoint_weight, res.point_weight)
    return (; points=ds_dpoints, rotation=ds_drotation, translation=ds_dtranslation, background=ds_dbackground,
            out_weight=ds_dout_weight, point_weight=ds_dpoint_weight)
end

# CuArray allocator overrides, kept exactly as ext/DiffPointRasterisationCUDAExt.jl:323-333: on the GPU d_points is
# (N_in, P) and d_point_weight is (P) - there are no per-thread slabs.
function DiffPointRasterisation.default_ds_dpoints_batched(
    points::CuVector{<:AbstractVector{TP}}, N_in, batch_size
) where {TP<:Number}
    return similar(points, TP, (N_in, length(points)))
end

function DiffPointRasterisation.default_ds_dpoint_weight_batched(
    points::CuVector{<:AbstractVector{<:Number}}, T, batch_size
)
    return similar(points, T)
end

end  # module
