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
# NOTE: Julia is not installed in the build image, so this file is reviewed but not executed
# there; the same C ABI is exercised by the Python mirror (maprasterization.jl_b200/api.py) that
# the test-suite runs.  Keep the two in sync (include/maprast.h is the contract).
module MapRasterizationB200

using ColorTypes, FixedPointNumbers, Statistics
import Neighborhoods: Window

export categorize, clean, categorize_clean, categorize_u8, clean_u8, categorize_clean_u8, clean_segments,
       clean_segments_u8, rasterize, MapSelection, Context, Palette

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

_channels(::Type{<:RGB}) = (:r, :g, :b)
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

_check_match(match, segmented) = begin
    m = match isa Val ? typeof(match).parameters[1] : Tuple(match)
    m == (:color,) || throw(ArgumentError("match=$m is not supported by the B200 hot path (only (:color,))"))
    segmented && throw(ArgumentError("segmented=true is not supported by the B200 hot path"))
end

# ---- the hot path ---------------------------------------------------------------------------------
"Fused clean(categorize(img)): UInt8 labels, 0 = no match."
function categorize_clean_u8(img::AbstractMatrix{T}, points_or_palette; tolerances=nothing,
        hood=Window{1}(), match=(:color,), segmented=false, scan_threshold=0.1,
        ctx::Context=default_context()) where {T<:Union{RGB{N0f8},RGBA{N0f8}}}
    _check_match(match, segmented)
    A = img isa Matrix ? img : Matrix(img)                       # dense, column-major, n1 = size(A, 1)
    pal = points_or_palette isa Palette ? points_or_palette :
          Palette(A, points_or_palette; tolerances=something(tolerances, fill(DEFAULT_CATEGORY_TOLERANCE, length(points_or_palette))))
    _upload(ctx, pal)
    n1, n2 = size(A)
    R = _radius(hood)
    out = Matrix{UInt8}(undef, n1, n2)
    GC.@preserve A out begin
        _check(ctx, ccall((:mr_categorize_clean_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Cint, Cint, Ptr{UInt8}, Int64, Ptr{MRCounters}),
            ctx.ptr, pointer(A), n1, n2, n1 * sizeof(T), sizeof(T), R, pointer(out), n1, C_NULL))
    end
    return out
end

_radius(::Window{R}) where {R} = R

# eltype Int and the input's container type, like `map(pixels) do ... end` (src/categories.jl:69-73)
_rewrap(img, labels::Matrix{UInt8}) = _rebuild(img, Int.(labels))
_rebuild(img::Matrix, data) = data
_rebuild(img, data) = applicable(rebuild, img, data) ? rebuild(img, data) : data   # Rasters.Raster in => Raster out

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
    if known_categories === nothing
        _check(ctx, ccall((:mr_set_known, libmaprast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}), ctx.ptr, 0, C_NULL, C_NULL))
    else
        size(known_categories) == size(A) || throw(ArgumentError("known_categories must have the size of A"))
        idx = Int64[i - 1 for i in eachindex(known_categories) if known_categories[i] != 0]   # i + j*n1, 0-based
        cat = UInt8[known_categories[i + 1] for i in idx]
        _check(ctx, ccall((:mr_set_known, libmaprast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{UInt8}), ctx.ptr, length(idx), idx, cat))
    end
    out = similar(L)
    nseg = Ref{Int64}(0)
    GC.@preserve L out keep begin
        _check(ctx, ccall((:mr_clean_segments_host, libmaprast), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int64, Ptr{UInt8}, Float64, Float64, Ptr{UInt8}, Int64, Ref{Int64}),
            ctx.ptr, pointer(L), n1, n2, n1, pointer(keep), line_threshold, prune_threshold, pointer(out), n1, nseg))
    end
    return out
end
clean_segments(A, known=nothing; kw...) = _rewrap(A, clean_segments_u8(A, known; kw...))

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
The button sequence of `selectcolors` without the GUI: "categorize" (src/selectcolors.jl:466-478),
optional Window majority filter (`hood`), "clean" (:479-489).  Returns `(; categorized, cleaned)`.
"""
function rasterize(img, sel::MapSelection; hood=Window{0}(), ctx::Context=default_context())
    get(sel.settings, :blur_repeat, 0) == 0 ||
        throw(ArgumentError("blur_repeat > 0 is not supported by the B200 hot path"))
    categorized = categorize_clean_u8(img, sel.points;
        tolerances=collect(Float64, sel.settings.category_tolerances), hood,
        match=get(sel.settings, :match, (:color,)), segmented=get(sel.settings, :segment, false), ctx)
    known = _known_plane(size(categorized), sel.points)
    keep = get(sel.settings, :category_keep, fill(true, length(sel.points)))
    cleaned = clean_segments_u8(categorized, known; categories=_kept_categories(collect(Bool, keep)),
        line_threshold=get(sel.settings, :line_threshold, 0.01),
        prune_threshold=get(sel.settings, :prune_threshold, 20), ctx)
    return (; categorized=_rewrap(img, categorized), cleaned=_rewrap(img, cleaned))
end

end # module
