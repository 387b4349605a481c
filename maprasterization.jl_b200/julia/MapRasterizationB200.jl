# MapRasterizationB200.jl -- Julia host side of the B200 categorise + clean hot path.
#
# Drop-in for the per-pixel path of rafaqz/MapRasterization.jl: same names (minus the leading
# underscore), keyword arguments and return types as
#   _categorize(pixels, points; tolerances, match, segmented, scan_threshold)   src/categories.jl:49-74
#   _component_stats(A, pointvecs)                                              src/categories.jl:2-30
#   blur(A; hood=Window{1}(), repeat=1)  (kwarg shape donor for `clean`)        src/blur.jl:6
#   _clean(A, known_categories; categories, line_threshold, prune_threshold)    src/clean.jl:59-93
#     (exposed as `clean_segments`: the name `clean` is the north_star's Window majority filter)
# Every label is computed by libmaprast.so (CUDA, sm_100a) through `ccall`; there is NO CPU
# fallback: unsupported options throw ArgumentError, a missing GPU makes `Context()` throw.
#
#   fill_gaps(src; categories, stats, known, original, mask, ...)                src/clean.jl:2-54
#
# NOTE: Julia is not installed in the build image, so this file has NEVER BEEN EXECUTED; the same C
# ABI is exercised by the Python mirror (maprasterization.jl_b200/api.py) that the test-suite runs, and
# tests/test_julia_binding.py checks every `ccall` below (symbol, return type, argument count and
# widths) against the prototypes of include/maprast.h.  Keep the two in sync (the header is the contract).
module MapRasterizationB200

using ColorTypes, FixedPointNumbers, Statistics
import Neighborhoods: Window

export categorize, clean, categorize_clean, categorize_u8, clean_u8, categorize_clean_u8, clean_segments,
       clean_segments_u8, fill_gaps, fill_gaps_u8, blur, balance, stds, stripes, polymask, paint_polygons, shrink_grow, erode, dilate, recolor, rasterize, MapSelection, Context, MultiContext, Palette

const libmaprast = get(ENV, "LIBMAPRAST", joinpath(@__DIR__, "..", "libmaprast.so"))
const DEFAULT_CATEGORY_TOLERANCE = 2.0            # src/selectcolors.jl:1

struct MRCounters
    n_pixels::UInt64
    n_nomatch::UInt64
    n_fine::UInt64
    n_changed::UInt64
end

mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(device::Integer=0)
        err = Ref{Cint}(0)
        p = ccall((:mr_create, libmaprast), Ptr{Cvoid}, (Cint, Ref{Cint}), device, err)
        p == C_NULL && error(unsafe_string(ccall((:mr_last_error, libmaprast), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(p)
        finalizer(c -> ccall((:mr_destroy, libmaprast), Cvoid, (Ptr{Cvoid},), c.ptr), ctx)
        return ctx
    end
end

const _default_ctx = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(_default_ctx[], (_default_ctx[] = Context(); _default_ctx[]))

function _check(ctx::Context, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:mr_last_error, libmaprast), Cstring, (Ptr{Cvoid},), ctx.ptr))
    (rc == -1 || rc == -5) ? throw(ArgumentError(msg)) : error("libmaprast ($rc): $msg")
end

# ---- palette = `category_stats` of src/categories.jl:52 ----------------------------------------
struct Palette
    mean::Matrix{Float64}   # C x K (column = category), C = 3 (RGB / Lab) or 4 (RGBA)
    lo::Matrix{Float64}
    hi::Matrix{Float64}
    colourspace::Cint       # 0 RGB, 1 Lab
end

_channels(::Type{<:RGB}) = (:r, :g, :b)      # RGB{N0f8} and RGB{Float64} (the output of `blur`)
_channels(::Type{<:RGBA}) = (:r, :g, :b, :alpha)

"`_component_stats` for one category (src/categories.jl:5-9, 18-30) on Float64 channels."
function _stats(A::AbstractMatrix{T}, pointvec) where {T<:Colorant}
    length(pointvec) == 0 && return missing
    px = map(P -> A[map(p -> round(Int, p), P)...], pointvec)            # :155-157
    map(_channels(T)) do k
        v = [Float64(getfield(c, k)) for c in px]                        # Float64(::N0f8) = i/255
        (mean=mean(v), min=minimum(v), max=maximum(v), sd=std(v))        # :29
    end
end

function Palette(A::AbstractMatrix{T}, points; tolerances=fill(DEFAULT_CATEGORY_TOLERANCE, length(points))) where {T<:Colorant}
    C, K = length(_channels(T)), length(points)
    mean_ = fill(Inf, C, K); lo = fill(NaN, C, K); hi = fill(NaN, C, K)   # `missing` category: mean = +Inf
    for (k, pv) in enumerate(points)
        st = _stats(A, pv)
        ismissing(st) && continue
        for (c, s) in enumerate(st)
            mean_[c, k] = s.mean
            lo[c, k] = s.min - s.sd * tolerances[k]                      # src/categories.jl:131-133
            hi[c, k] = s.max + s.sd * tolerances[k]
        end
    end
    all(isinf, mean_) && throw(ArgumentError("no category has any points"))
    Palette(mean_, lo, hi, 0)
end

function _upload(ctx::Context, pal::Palette)
    C, K = size(pal.mean)
    _check(ctx, ccall((:mr_set_palette, libmaprast), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint),
        ctx.ptr, K, pal.mean, pal.lo, pal.hi, pal.colourspace, C))
end

# -> match bits (1 color | 2 std | 4 stripe); `match` must be a non-empty subset of (:color, :std, :stripe) in that
# order (what `_match_from_bools` produces, src/selectcolors.jl:661-668: the error is a left-folded mean over it)
_check_match(match, segmented) = begin
    m = match isa Val ? typeof(match).parameters[1] : Tuple(match)
    ok = !isempty(m) && all(x -> x in (:color, :std, :stripe), m) && collect(m) == [x for x in (:color, :std, :stripe) if x in m]
    ok || throw(ArgumentError("match=$m: must be a non-empty subset of (:color, :std, :stripe) in that order"))
    segmented && throw(ArgumentError("segmented=true (fast_scanning: a sequential scan) is not supported on the GPU path"))
    return Cint((:color in m) + 2 * (:std in m) + 4 * (:stripe in m))
end

# ---- known-category plane (`_set_categories!`, src/selectcolors.jl:670-681) ---------------------------
# The plane lives in the context until it is replaced: EVERY entry point below sets or clears it before
# its call and clears it afterwards, so that no call ever sees the points of an earlier one.
_clear_known(ctx::Context) =
    _check(ctx, ccall((:mr_set_known, libmaprast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}), ctx.ptr, 0, C_NULL, C_NULL))

function _set_known(ctx::Context, known::Union{Nothing,AbstractMatrix}, sz)
    known === nothing && return _clear_known(ctx)
    size(known) == sz || throw(ArgumentError("the known-category plane must have the size of the image"))
    idx = Int64[i - 1 for i in eachindex(IndexLinear(), known) if known[i] != 0]   # i + j*n1, 0-based
    cat = UInt8[known[i + 1] for i in idx]
    GC.@preserve idx cat begin
        _check(ctx, ccall((:mr_set_known, libmaprast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}),
            ctx.ptr, length(idx), pointer(idx), pointer(cat)))
    end
end

# ---- the hot path ---------------------------------------------------------------------------------
"""
Fused clean(categorize(img)): UInt8 labels, 0 = no match.  `known`: optional plane of known categories
(0 = none) for the override of src/categories.jl:110-112; `nothing` = no clicked points.
"""
function categorize_clean_u8(img::AbstractMatrix{T}, points_or_palette; tolerances=nothing,
        hood=Window{1}(), match=(:color,), segmented=false, scan_threshold=0.1, known=nothing,
        ctx::Context=default_context()) where {T<:Union{RGB{N0f8},RGBA{N0f8}}}
    _check_match(match, segmented) == 1 || throw(ArgumentError("std / stripe matching takes the blurred RGB{Float64} image"))
    A = img isa Matrix ? img : Matrix(img)                       # dense, column-major, n1 = size(A, 1)
    pal = points_or_palette isa Palette ? points_or_palette :
          Palette(A, points_or_palette; tolerances=something(tolerances, fill(DEFAULT_CATEGORY_TOLERANCE, length(points_or_palette))))
    _upload(ctx, pal)
    n1, n2 = size(A)
    R = _radius(hood)
    out = Matrix{UInt8}(undef, n1, n2)
    _set_known(ctx, known, size(A))
    try
        GC.@preserve A out begin
            _check(ctx, ccall((:mr_categorize_clean_host, libmaprast), Cint,
                (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Cint, Ptr{UInt8}, Int64, Ptr{MRCounters}),
                ctx.ptr, pointer(A), n1, n2, n1 * sizeof(T), sizeof(T), R, pointer(out), n1, C_NULL))
        end
    finally
        known === nothing || _clear_known(ctx)
    end
    return out
end

_radius(::Window{R}) where {R} = R

"""
`blur(A; hood=Window{1}(), repeat=1)` (src/blur.jl:6-14): `repeat` passes of the Window mean in RGB{Float64}.
`A`: `Matrix{RGB{N0f8}}` or `Matrix{RGB{Float64}}`; returns `Matrix{RGB{Float64}}` (the input's wrapper type is kept).
"""
function blur(img::AbstractMatrix{T}; hood=Window{1}(), repeat::Int=1, ctx::Context=default_context()) where {T<:Union{RGB{N0f8},RGB{Float64}}}
    A = img isa Matrix ? img : Matrix(img)
    n1, n2 = size(A)
    out = Matrix{RGB{Float64}}(undef, n1, n2)
    GC.@preserve A out begin
        _check(ctx, ccall((:mr_blur_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Float64}, Int64),
            ctx.ptr, pointer(A), n1, n2, n1 * sizeof(T), sizeof(T), _radius(hood), repeat,
            Ptr{Float64}(pointer(out)), n1 * 24))
    end
    return _rebuild(img, out)
end

"""
The host part of `_balance(A, points; category_bitindex)` (src/balance.jl:1-24): colour deviations of the
clicked points from their category's mean colour, least-squares fit of `c ~ i^3 + i^2 + i + j^3 + j^2 + j`
per channel (`X \\ Y` stands in for GLM's `lm`: same minimiser).  Returns the 7 x 3 coefficients
(intercept first) or `nothing` with fewer than 8 points (the reference then returns `A .* 1.0`, :16-18).
"""
function balance_fit(A::AbstractMatrix{<:Colorant}, points; category_bitindex=[c == 1 for c in 1:length(points)])
    rows = Vector{NTuple{7,Float64}}(); resp = Vector{NTuple{3,Float64}}()
    for (c, pv) in enumerate(points)
        (category_bitindex[c] && length(pv) > 0) || continue
        cols = [begin px = A[round(Int, p[1]), round(Int, p[2])]; (Float64(px.r), Float64(px.g), Float64(px.b)) end for p in pv]
        m = ntuple(k -> mean(getindex.(cols, k)), 3)
        for (p, col) in zip(pv, cols)
            i, j = Float64(p[1]), Float64(p[2])
            push!(rows, (1.0, i^3, i^2, i, j^3, j^2, j)); push!(resp, col .- m)
        end
    end
    length(rows) < 8 && return nothing
    X = [r[k] for r in rows, k in 1:7]; Y = [r[k] for r in resp, k in 1:3]
    return X \ Y                                   # 7 x 3, column = channel: the layout mr_balance_* reads (3 x 7 row-major)
end

"`_balance(A, points; category_bitindex)` (src/balance.jl:1-42): RGB{Float64} out; fit on the host, pixels on the GPU."
function balance(img::AbstractMatrix{RGB{N0f8}}, points; category_bitindex=[c == 1 for c in 1:length(points)],
        ctx::Context=default_context())
    A = img isa Matrix ? img : Matrix(img)
    coef = balance_fit(A, points; category_bitindex)
    coef === nothing && return blur(img; repeat=0, ctx)          # A .* 1.0
    n1, n2 = size(A)
    out = Matrix{RGB{Float64}}(undef, n1, n2)
    C = Matrix{Float64}(coef)                                       # 7 x 3 column-major = 3 x 7 row-major
    GC.@preserve A out C begin
        _check(ctx, ccall((:mr_balance_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Int64),
            ctx.ptr, Ptr{UInt8}(pointer(A)), n1, n2, n1 * 3, 3, pointer(C), Ptr{Float64}(pointer(out)), n1 * 24))
    end
    return _rebuild(img, out)
end

"""
`_stds(A)` (src/selectcolors.jl:706-715): the "texture" plane of the blurred RGB{Float64} image
(5 x 5 per-channel sample standard deviations, summed, divided by the image maximum).
"""
function stds(img::AbstractMatrix{RGB{Float64}}; ctx::Context=default_context())
    A = img isa Matrix ? img : Matrix(img)
    n1, n2 = size(A)
    out = Matrix{Float64}(undef, n1, n2)
    GC.@preserve A out begin
        _check(ctx, ccall((:mr_stds_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64),
            ctx.ptr, Ptr{Float64}(pointer(A)), n1, n2, n1 * 24, pointer(out), n1 * 8))
    end
    return _rebuild(img, out)
end

"`_stripes(img; radius=3)` (src/stripes.jl:20-35): `Matrix{NTuple{2,Float64}}` of (vert - horz, angle45 - angle135)."
function stripes(img::AbstractMatrix{RGB{Float64}}; radius::Int=3, ctx::Context=default_context())
    A = img isa Matrix ? img : Matrix(img)
    n1, n2 = size(A)
    out = Matrix{NTuple{2,Float64}}(undef, n1, n2)
    GC.@preserve A out begin
        _check(ctx, ccall((:mr_stripes_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Ptr{Float64}, Int64),
            ctx.ptr, Ptr{Float64}(pointer(A)), n1, n2, n1 * 24, radius, Ptr{Float64}(pointer(out)), n1 * 16))
    end
    return _rebuild(img, out)
end

# `std` / `stripe` of _component_stats(::Vector{<:PixelProp}) (src/categories.jl:10-17) at the clicked points,
# bounds as in :131-133; rows of a `missing` category stay NaN (its colour mean is +Inf)
function _upload_texture(ctx::Context, std, stripe, points, tolerances)
    K = length(points)
    sm = fill(NaN, K); slo = fill(NaN, K); shi = fill(NaN, K)
    tm = fill(NaN, 2, K); tlo = fill(NaN, 2, K); thi = fill(NaN, 2, K)
    stat(v, t) = (mean(v), minimum(v) - std_(v) * t, maximum(v) + std_(v) * t)
    std_(v) = Statistics.std(v)
    for (k, pv) in enumerate(points)
        length(pv) == 0 && continue
        idx = [CartesianIndex(round(Int, p[1]), round(Int, p[2])) for p in pv]
        sm[k], slo[k], shi[k] = stat([std === nothing ? 0.0 : std[i] for i in idx], tolerances[k])
        for e in 1:2
            tm[e, k], tlo[e, k], thi[e, k] = stat([stripe === nothing ? 0.0 : stripe[i][e] for i in idx], tolerances[k])
        end
    end
    _check(ctx, ccall((:mr_set_texture_stats, libmaprast), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.ptr, K, sm, slo, shi, tm, tlo, thi))
end

"""
Fused clean(categorize(img)) on RGB{Float64} pixels (what `blur` / `_balance` hand to `_categorize`): direct
Float64 evaluation.  With `:std` / `:stripe` in `match` the pixels are PixelProps (src/categories.jl:45-48):
`std` / `stripe` are the planes of `stds` / `stripes` (`nothing` = the reference's initial all-zero plane) and
`points_or_palette` must be the clicked points (the texture statistics are sampled at them).
"""
function categorize_clean_u8(img::AbstractMatrix{RGB{Float64}}, points_or_palette; tolerances=nothing,
        hood=Window{1}(), match=(:color,), segmented=false, scan_threshold=0.1, known=nothing,
        std=nothing, stripe=nothing, ctx::Context=default_context())
    bits = _check_match(match, segmented)
    A = img isa Matrix ? img : Matrix(img)
    tol = points_or_palette isa Palette ? nothing :
          something(tolerances, fill(DEFAULT_CATEGORY_TOLERANCE, length(points_or_palette)))
    pal = points_or_palette isa Palette ? points_or_palette : Palette(A, points_or_palette; tolerances=tol)
    _upload(ctx, pal)
    n1, n2 = size(A)
    out = Matrix{UInt8}(undef, n1, n2)
    S = (bits & 2) != 0 ? Matrix{Float64}(std === nothing ? zeros(n1, n2) : std) : nothing
    P = (bits & 4) != 0 ? Matrix{NTuple{2,Float64}}(stripe === nothing ? fill((0.0, 0.0), n1, n2) : stripe) : nothing
    if bits != 1
        points_or_palette isa Palette && throw(ArgumentError("std / stripe matching needs the clicked points, not a Palette"))
        _upload_texture(ctx, S, P, points_or_palette, tol)
    end
    _set_known(ctx, known, size(A))
    try
        GC.@preserve A out S P begin
            if bits == 1
                _check(ctx, ccall((:mr_categorize_clean_f64_host, libmaprast), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Ptr{UInt8}, Int64),
                    ctx.ptr, Ptr{Float64}(pointer(A)), n1, n2, n1 * 24, _radius(hood), pointer(out), n1))
            else
                _check(ctx, ccall((:mr_categorize_clean_props_host, libmaprast), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Cint, Cint,
                     Ptr{UInt8}, Int64),
                    ctx.ptr, Ptr{Float64}(pointer(A)), n1 * 24, S === nothing ? Ptr{Float64}(C_NULL) : pointer(S), n1 * 8,
                    P === nothing ? Ptr{Float64}(C_NULL) : Ptr{Float64}(pointer(P)), n1 * 16, n1, n2, bits, _radius(hood),
                    pointer(out), n1))
            end
        end
    finally
        known === nothing || _clear_known(ctx)
    end
    return out
end

# eltype Int and the input's container type, like `map(pixels) do ... end` (src/categories.jl:69-73)
_rewrap(img, labels::Matrix{UInt8}) = _rebuild(img, Int.(labels))
_rebuild(img::Matrix, data) = data
# Raster / DimArray in => the same wrapper out: `rebuild` is looked up in the module that owns the input's
# type (Rasters and DimensionalData both bind it), so this file needs no dependency on either package.
function _rebuild(img, data)
    m = parentmodule(typeof(img))
    return isdefined(m, :rebuild) ? getfield(m, :rebuild)(img, data) : data
end

categorize_clean(img, p; kw...) = _rewrap(img, categorize_clean_u8(img, p; kw...))
categorize_u8(img, p; kw...) = categorize_clean_u8(img, p; hood=Window{0}(), kw...)
"`_categorize(pixels, points; tolerances, match, segmented, scan_threshold)` -> labels (eltype Int)."
categorize(img, p; kw...) = _rewrap(img, categorize_u8(img, p; kw...))

"`clean(labels; hood=Window{1}(), repeat=1)`: Window majority filter, ties -> lowest label."
function clean_u8(labels::AbstractMatrix{<:Integer}; hood=Window{1}(), repeat::Int=1, ctx::Context=default_context())
    (isempty(labels) || (0 <= minimum(labels) && maximum(labels) <= 254)) || throw(ArgumentError("labels must be in 0..254"))
    L = Matrix{UInt8}(labels)
    n1, n2 = size(L)
    out = similar(L)
    GC.@preserve L out begin
        _check(ctx, ccall((:mr_clean_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Cint, Ptr{UInt8}, Int64),
            ctx.ptr, pointer(L), n1, n2, n1, _radius(hood), repeat, pointer(out), n1))
    end
    return out
end
clean(labels; kw...) = _rewrap(labels, clean_u8(labels; kw...))

"""
`_clean(A, known_categories; categories, line_threshold=0.01, prune_threshold=20)` (src/clean.jl:59-93):
4-connected segments of equal labels (`fast_scanning!`, src/scan.jl:56-157); segments of a category that
is not in `categories`, with fewer than `prune_threshold` pixels, or with
`count / (max(h, w)^2 / 2) < line_threshold` become 0 -- or the known category clicked inside them.
Throws if one segment holds two different known categories (the reference result then depends on
its scan order).  UInt8 labels out (the reference returns the Float64 segment colour).
"""
function clean_segments_u8(A::AbstractMatrix{<:Integer}, known_categories=nothing; categories,
        line_threshold=0.01, prune_threshold=20, ctx::Context=default_context())
    (isempty(A) || (0 <= minimum(A) && maximum(A) <= 254)) || throw(ArgumentError("labels must be in 0..254"))
    L = Matrix{UInt8}(A)
    n1, n2 = size(L)
    keep = zeros(UInt8, 256)
    for c in categories
        keep[c + 1] = 1
    end
    out = similar(L)
    nseg = Ref{Int64}(0)
    _set_known(ctx, known_categories, size(A))
    try
        GC.@preserve L out keep begin
            _check(ctx, ccall((:mr_clean_segments_host, libmaprast), Cint,
                (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Ptr{UInt8}, Float64, Float64, Ptr{UInt8}, Int64, Ref{Int64}),
                ctx.ptr, pointer(L), n1, n2, n1, pointer(keep), Float64(line_threshold), Float64(prune_threshold),
                pointer(out), n1, nseg))
        end
    finally
        known_categories === nothing || _clear_known(ctx)
    end
    return out
end
clean_segments(A, known=nothing; kw...) = _rewrap(A, clean_segments_u8(A, known; kw...))

"""
`fill_gaps(src; categories, neighborhood=VonNeumann{2}(), missingval=0, stats, match=(:color,), known, original,
mask)` (src/clean.jl:2-54), applied `repeat` times as `_fill` does (src/selectcolors.jl:683-699).
`stats`: the `Palette` of the categoriser (`stats[i]` of clean.jl:29-30 = its column `categories[i]`);
`original`: the image the categoriser saw; `known` / `mask`: planes of the size of `src` or `nothing`.
Only `missingval = 0`, the VonNeumann{2} neighbourhood and `match = (:color,)` exist on the GPU path.
"""
function fill_gaps_u8(src::AbstractMatrix{<:Integer}; categories, stats::Palette, original::AbstractMatrix{T},
        known=nothing, mask=nothing, match=(:color,), missingval=0, repeat::Int=1,
        ctx::Context=default_context()) where {T<:Union{RGB{N0f8},RGBA{N0f8}}}
    missingval == 0 || throw(ArgumentError("fill_gaps: only missingval = 0 is supported"))
    _check_match(match, false)
    size(original) == size(src) || throw(ArgumentError("original must have the size of src"))
    (isempty(src) || (0 <= minimum(src) && maximum(src) <= 254)) || throw(ArgumentError("labels must be in 0..254"))
    L = Matrix{UInt8}(src)
    O = original isa Matrix ? original : Matrix(original)
    n1, n2 = size(L)
    cats = UInt8[c for c in categories]
    M = mask === nothing ? UInt8[] : Matrix{UInt8}(mask .!= 0)
    mask === nothing || size(mask) == size(src) || throw(ArgumentError("mask must have the size of src"))
    out = similar(L)
    _upload(ctx, stats)
    _set_known(ctx, known, size(src))
    try
        GC.@preserve L O M cats out begin
            _check(ctx, ccall((:mr_fill_gaps_host, libmaprast), Cint,
                (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Ptr{UInt8}, Int64, Cint, Ptr{UInt8}, Int64,
                 Ptr{UInt8}, Cint, Cint, Ptr{UInt8}, Int64),
                ctx.ptr, pointer(L), n1, n2, n1, pointer(O), n1 * sizeof(T), sizeof(T),
                mask === nothing ? Ptr{UInt8}(C_NULL) : pointer(M), n1, pointer(cats), length(cats), repeat,
                pointer(out), n1))
        end
    finally
        known === nothing || _clear_known(ctx)
    end
    return out
end
fill_gaps(src; kw...) = _rewrap(src, fill_gaps_u8(src; kw...))

# ---- display layers (src/selectcolors.jl:650-655, 701-704) ------------------------------------------------
"`_mean_color(A, pointvec)` (src/selectcolors.jl:701-704) as (r, g, b) Float64; the zero colour without points."
function _mean_rgb(A::AbstractMatrix{<:Colorant}, pointvec)
    st = _stats(A, pointvec)
    return ismissing(st) ? (0.0, 0.0, 0.0) : (st[1].mean, st[2].mean, st[3].mean)
end

"`_recolor(categories, source, points)` (src/selectcolors.jl:650-655): `Matrix{RGBA{Float64}}`, transparent black for 0."
function recolor(categories::AbstractMatrix{<:Integer}, source::AbstractMatrix{<:Colorant}, points; ctx::Context=default_context())
    L = Matrix{UInt8}(categories)
    n1, n2 = size(L)
    cols = Float64[c for ps in points for c in _mean_rgb(source, ps)]          # K x 3, row = category
    out = Matrix{RGBA{Float64}}(undef, n1, n2)
    GC.@preserve L cols out begin
        _check(ctx, ccall((:mr_recolor_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Int64),
            ctx.ptr, pointer(L), n1, n2, n1, length(points), pointer(cols), Ptr{Float64}(pointer(out)), n1 * 32))
    end
    return _rebuild(categories, out)
end

# ---- categorical erosion / dilation (src/clean.jl:96-140) ------------------------------------------------
"`_shrink_grow(A, i, j)` (src/clean.jl:132-140): `i` passes of `_erode` (:121-129), then `j` passes of `_dilate` (:96-118)."
function shrink_grow_u8(A::AbstractMatrix{<:Integer}, i::Int, j::Int; ctx::Context=default_context())
    L = Matrix{UInt8}(A)
    n1, n2 = size(L)
    out = similar(L)
    GC.@preserve L out begin
        _check(ctx, ccall((:mr_shrink_grow_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Cint, Ptr{UInt8}, Int64),
            ctx.ptr, pointer(L), n1, n2, n1, i, j, pointer(out), n1))
    end
    return out
end
shrink_grow(A, i::Int, j::Int; kw...) = _rewrap(A, shrink_grow_u8(A, i, j; kw...))
erode(A; kw...) = shrink_grow(A, 1, 0; kw...)
dilate(A; kw...) = shrink_grow(A, 0, 1; kw...)

# ---- polygons (src/selectcolors.jl:480-481, 491-497, 559-563) ------------------------------------------
# `polygons`: per category a vector of rings, a ring = vector of points (the reference's
# Vector{Vector{Vector{Point2{Float32}}}}, src/selectcolors.jl:35, :60).  -> offsets, coordinates, category per ring
function _rings(polygons)
    off = Int32[0]; xy = Float64[]; cats = UInt8[]
    for (c, polys) in enumerate(polygons), r in polys
        for p in r
            push!(xy, Float64(Float32(p[1])), Float64(Float32(p[2])))
        end
        push!(off, Int32(length(xy) ÷ 2)); push!(cats, UInt8(c))
    end
    return off, xy, cats
end

function _polygons!(plane::Matrix{UInt8}, polygons, fill; ctx::Context=default_context())
    off, xy, cats = _rings(polygons)
    f = fill === nothing ? nothing : (fill === :category ? cats : Base.fill(UInt8(fill), length(cats)))
    n1, n2 = size(plane)
    GC.@preserve plane off xy f begin
        _check(ctx, ccall((:mr_polygons_host, libmaprast), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}, Int64, Int64, Ptr{UInt8}, Int64),
            ctx.ptr, length(cats), pointer(off), isempty(xy) ? Ptr{Float64}(C_NULL) : pointer(xy),
            f === nothing ? Ptr{UInt8}(C_NULL) : pointer(f), n1, n2, pointer(plane), n1))
    end
    return plane
end

"`_polymask(polygons, A)` (src/selectcolors.jl:559-563): `Matrix{Bool}` of the pixels inside any polygon."
polymask(polygons, sz::Tuple{Int,Int}; ctx::Context=default_context()) =
    _polygons!(Matrix{UInt8}(undef, sz), polygons, nothing; ctx) .!= 0
"`rasterize!(raster, p; fill = i, shape = :polygon)` for the polygons of every category i (:494-496); `fill = 0`: `(1 .- mask) .* A` (:481)."
paint_polygons(labels::AbstractMatrix{<:Integer}, polygons; fill=:category, ctx::Context=default_context()) =
    _polygons!(Matrix{UInt8}(labels), polygons, fill; ctx)

# ---- MapSelection: the non-interactive entry (src/selectcolors.jl:3-8, 21-24, 111-127) ------------
"Same fields as the reference's `MapSelection`; `settings` is the NamedTuple built at src/selectcolors.jl:111-125."
struct MapSelection{S<:NamedTuple,O}
    settings::S
    removals::Vector{Vector{NTuple{2,Float32}}}
    points::Vector{Vector{NTuple{2,Float32}}}
    output::O
end

_kept_categories(bitindex::AbstractVector{Bool}) = (1:length(bitindex))[bitindex]     # src/selectcolors.jl:189

"`_set_categories!` (src/selectcolors.jl:670-681): later categories overwrite earlier ones."
function _known_plane(sz, points)
    A = zeros(UInt8, sz)
    for (i, pv) in enumerate(points), p in pv
        A[round(Int, p[1]), round(Int, p[2])] = i
    end
    return A
end

categorize_u8(img, sel::MapSelection; kw...) =
    categorize_clean_u8(img, sel.points; tolerances=collect(Float64, sel.settings.category_tolerances),
                        hood=Window{0}(), match=get(sel.settings, :match, (:color,)),
                        segmented=get(sel.settings, :segment, false), kw...)
categorize(img, sel::MapSelection; kw...) = _rewrap(img, categorize_u8(img, sel; kw...))

"""
The button sequence of `selectcolors` without the GUI: "categorize" (src/selectcolors.jl:466-478, with the
known-category override of the clicked points), optional Window majority filter (`hood`), "clean"
(:479-489), "fill" (:490-501, `fill_repeat` passes of `fill_gaps`; `mask` = the caller's polygon mask).
Returns `(; categorized, cleaned, filled)`.
"""
function rasterize(img::AbstractMatrix{T}, sel::MapSelection; hood=Window{0}(), mask=nothing, ctx::Context=default_context()) where {T}
    A = img isa Matrix ? img : Matrix(img)
    # balanced = _balance(raw, points; category_bitindex = category_balance) (src/selectcolors.jl:222; default:
    # category 1), then the "blur" button (:455-465).  Fewer than 8 balance points: A .* 1.0 = the N0f8 path.
    blur_repeat = get(sel.settings, :blur_repeat, 0)
    bal = get(sel.settings, :category_balance, [c == 1 for c in 1:length(sel.points)])
    detrend = T <: RGB{N0f8} && balance_fit(A, sel.points; category_bitindex=collect(Bool, bal)) !== nothing
    pixels = detrend ? balance(A, sel.points; category_bitindex=collect(Bool, bal), ctx) : A
    pixels = blur_repeat > 0 ? blur(pixels; repeat=blur_repeat, ctx) : pixels
    tol = collect(Float64, sel.settings.category_tolerances)
    pal = Palette(pixels, sel.points; tolerances=tol)
    known = _known_plane(size(A), sel.points)
    match = get(sel.settings, :match, (:color,))
    bits = _check_match(match, get(sel.settings, :segment, false))
    if bits == 1
        categorized = categorize_clean_u8(pixels, pal; hood, match, known, ctx)
    else
        # the "blur" button refreshes the texture planes of the matched properties (src/selectcolors.jl:457-462)
        eltype(pixels) <: RGB{Float64} || (pixels = blur(pixels; repeat=0, ctx))      # A .* 1.0
        S = (bits & 2) != 0 ? stds(pixels; ctx) : nothing
        P = (bits & 4) != 0 ? stripes(pixels; radius=get(sel.settings, :stripe_radius, 2), ctx) : nothing
        categorized = categorize_clean_u8(pixels, sel.points; tolerances=tol, hood, match, known, std=S, stripe=P, ctx)
    end
    keep = get(sel.settings, :category_keep, fill(true, length(sel.points)))
    kept = _kept_categories(collect(Bool, keep))
    # "clean" (src/selectcolors.jl:479-489): masked = (1 .- mask) .* categorized, mask = the drawn polygons
    polys = get(sel.settings, :polygons, nothing)
    drawn = polys !== nothing && any(r -> length(r) >= 3, (r for ps in polys for r in ps))
    mask === nothing && drawn && (mask = polymask(polys, size(A); ctx))
    masked = drawn ? paint_polygons(categorized, polys; fill=0, ctx) : categorized
    cleaned = clean_segments_u8(masked, known; categories=kept,
        line_threshold=get(sel.settings, :line_threshold, 0.01),
        prune_threshold=get(sel.settings, :prune_threshold, 20), ctx)
    filled = fill_gaps_u8(cleaned; categories=kept, stats=pal, original=A, known, mask,
        repeat=get(sel.settings, :fill_repeat, 1), ctx)
    drawn && (filled = paint_polygons(filled, polys; ctx))      # rasterize!(fill_raster, p; fill = i) (:494-496)
    return (; categorized=_rewrap(img, categorized), cleaned=_rewrap(img, cleaned), filled=_rewrap(img, filled))
end

# ---- several GPUs of one box (mr_multi_*): one call, the library splits the sheet into row bands ------
mutable struct MultiContext
    ptr::Ptr{Cvoid}
    function MultiContext(devices::AbstractVector{<:Integer}=Int[])
        err = Ref{Cint}(0)
        ids = Cint[d for d in devices]
        p = GC.@preserve ids ccall((:mr_multi_create, libmaprast), Ptr{Cvoid}, (Ptr{Cint}, Cint, Ref{Cint}),
            isempty(ids) ? Ptr{Cint}(C_NULL) : pointer(ids), length(ids), err)
        p == C_NULL && error(unsafe_string(ccall((:mr_multi_last_error, libmaprast), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(p)
        finalizer(x -> ccall((:mr_multi_destroy, libmaprast), Cvoid, (Ptr{Cvoid},), x.ptr), m)
        return m
    end
end

function _check(m::MultiContext, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:mr_multi_last_error, libmaprast), Cstring, (Ptr{Cvoid},), m.ptr))
    (rc == -1 || rc == -5) ? throw(ArgumentError(msg)) : error("libmaprast ($rc): $msg")
end

"Fused clean(categorize(img)) over every GPU of the handle: byte-identical to the single-GPU call."
function categorize_clean_u8(img::AbstractMatrix{T}, pal::Palette, m::MultiContext; hood=Window{1}(),
        known=nothing) where {T<:Union{RGB{N0f8},RGBA{N0f8}}}
    A = img isa Matrix ? img : Matrix(img)
    n1, n2 = size(A)
    C, K = size(pal.mean)
    _check(m, ccall((:mr_multi_set_palette, libmaprast), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint),
        m.ptr, K, pal.mean, pal.lo, pal.hi, pal.colourspace, C))
    idx = known === nothing ? Int64[] : Int64[i - 1 for i in eachindex(IndexLinear(), known) if known[i] != 0]
    cat = UInt8[known[i + 1] for i in idx]
    out = Matrix{UInt8}(undef, n1, n2)
    GC.@preserve A out idx cat begin
        _check(m, ccall((:mr_multi_set_known, libmaprast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}),
            m.ptr, length(idx), pointer(idx), pointer(cat)))
        _check(m, ccall((:mr_multi_categorize_clean_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Cint, Ptr{UInt8}, Int64, Ptr{MRCounters}),
            m.ptr, pointer(A), n1, n2, n1 * sizeof(T), sizeof(T), _radius(hood), pointer(out), n1, C_NULL))
        _check(m, ccall((:mr_multi_set_known, libmaprast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}),
            m.ptr, 0, C_NULL, C_NULL))
    end
    return out
end

end # module
