"""Host-side mirror of the reference's interface for the categorise + clean path.

The reference is Julia (toolchain absent in this image), so this Python layer plays the role
of the Julia wrapper in maprasterization.jl_b200/julia/: same names (minus the leading
underscore), argument meaning and error behaviour as

  * `_categorize(pixels, points; tolerances, match, segmented, scan_threshold)`
       /root/reference/src/categories.jl:45-74  (call site src/selectcolors.jl:466-478)
  * `_component_stats`   src/categories.jl:2-30, `_point_values` :155-157
  * `blur(A; hood=Window{1}(), repeat=1)`  src/blur.jl:6  (kwarg shape of `clean`)

All labels are produced by libmaprast.so (CUDA); nothing here computes labels on the CPU.
Array convention: a Julia `Matrix{RGB{N0f8}}` of size n1 x n2 (column-major) is the numpy
array of shape (n2, n1, 3) (C order) -- the same bytes.  Labels: (n2, n1).
"""
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Context, MapRastError  # noqa: F401

DEFAULT_CATEGORY_TOLERANCE = 2.0   # src/selectcolors.jl:1
DEFAULT_MATCH = ("color", "std", "stripe")  # src/scan.jl:1; ("color",) alone takes the table-driven N0f8 hot path
MATCH_BITS = {"color": 1, "std": 2, "stripe": 4}   # `match` argument of mr_categorize_clean_props_*
DEFAULT_STRIPE_RADIUS = 2                   # src/selectcolors.jl:38


class Window:
    """`Window{R}` of Neighborhoods.jl: square of side 2R+1, centre included (src/blur.jl:6)."""

    def __init__(self, R=1):
        R = int(R)
        if R < 0:
            raise ValueError("Window radius must be >= 0")
        self.R = R

    def __repr__(self):
        return f"Window({self.R})"


def component_stats(values):
    """(mean, min, max, sd) per channel of an (n, C) array -- src/categories.jl:18-30.
    sd is the sample standard deviation (n-1); one sample gives NaN, as in Julia."""
    v = np.asarray(values, dtype=np.float64)
    if v.ndim == 1:
        v = v[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        sd = v.std(axis=0, ddof=1) if v.shape[0] > 1 else np.full(v.shape[1], np.nan)
    return {"mean": v.mean(axis=0), "min": v.min(axis=0), "max": v.max(axis=0), "sd": sd}


_SRGB_LIN = None


def rgb8_to_lab(rgb):
    """sRGB (D65) bytes -> Lab for HOST-SIDE palette statistics only.
    The per-pixel conversion used for labelling happens inside the CUDA table build."""
    global _SRGB_LIN
    if _SRGB_LIN is None:
        v = np.arange(256, dtype=np.float64) / 255.0
        _SRGB_LIN = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    rgb = np.asarray(rgb, dtype=np.uint8)
    lin = _SRGB_LIN[rgb]
    r, g, b = lin[..., 0], lin[..., 1], lin[..., 2]
    X = (0.4124564390896921 * r + 0.357576077643909 * g) + 0.18043748326639894 * b
    Y = (0.21267285140562248 * r + 0.715152155287818 * g) + 0.07217499330655958 * b
    Z = (0.019333895582329317 * r + 0.119192025881303 * g) + 0.9503040785363677 * b

    def f(t):
        return np.where(t > 216.0 / 24389.0, np.cbrt(t), (24389.0 / 27.0 * t + 16.0) / 116.0)

    fx, fy, fz = f(X / 0.95047), f(Y), f(Z / 1.08883)
    return np.stack([116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz)], axis=-1)


@dataclass
class Palette:
    """Per-category statistics = `category_stats` of src/categories.jl:52.

    mean/min/max/sd: (K, C) Float64; a category without points (`missing`, :6) has mean=+Inf.
    tolerances: (K,) -- `tolerances[best]` of src/categories.jl:81."""
    mean: np.ndarray
    min: np.ndarray
    max: np.ndarray
    sd: np.ndarray
    tolerances: np.ndarray = None
    colourspace: str = "rgb"
    lo: np.ndarray = field(init=False)
    hi: np.ndarray = field(init=False)

    def __post_init__(self):
        self.mean = np.atleast_2d(np.asarray(self.mean, dtype=np.float64))
        self.min = np.atleast_2d(np.asarray(self.min, dtype=np.float64))
        self.max = np.atleast_2d(np.asarray(self.max, dtype=np.float64))
        self.sd = np.atleast_2d(np.asarray(self.sd, dtype=np.float64))
        K = self.mean.shape[0]
        if self.tolerances is None:
            self.tolerances = np.full(K, DEFAULT_CATEGORY_TOLERANCE)
        self.tolerances = np.asarray(self.tolerances, dtype=np.float64).reshape(-1)
        if self.tolerances.size != K:
            raise ValueError("tolerances must have one entry per category")
        if self.colourspace not in ("rgb", "lab"):
            raise ValueError("colourspace must be 'rgb' or 'lab'")
        # src/categories.jl:131-133: (s.min - s.sd * tol), (s.max + s.sd * tol): two rounded ops each
        with np.errstate(invalid="ignore"):
            w = self.sd * self.tolerances[:, None]
            self.lo = self.min - w
            self.hi = self.max + w

    @property
    def K(self):
        return self.mean.shape[0]

    @property
    def channels(self):
        return self.mean.shape[1]

    @classmethod
    def from_points(cls, img, points, tolerances=None, colourspace="rgb"):
        """`_component_stats(A, pointvecs)` (src/categories.jl:2-9): sample img at the clicked
        points (1-based (i, j), rounded like `round(Int, p)`, :155-157) per category."""
        img = np.asarray(img)
        n2, n1, bpp = img.shape
        K = len(points)
        C = 3 if colourspace == "lab" else bpp
        mean = np.full((K, C), np.inf)
        mn = np.full((K, C), np.nan)
        mx = np.full((K, C), np.nan)
        sd = np.full((K, C), np.nan)
        for c, pts in enumerate(points):
            pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
            if pts.shape[0] == 0:
                continue   # `missing`
            ij = np.rint(pts).astype(np.int64)   # Julia round(Int, x): ties to even
            i, j = ij[:, 0] - 1, ij[:, 1] - 1
            if (i < 0).any() or (i >= n1).any() or (j < 0).any() or (j >= n2).any():
                raise IndexError("point outside the image (BoundsError in the reference)")
            samples = img[j, i]
            if img.dtype == np.float64:      # RGB{Float64} image (blurred / balanced): the values themselves
                if colourspace == "lab":
                    raise ValueError("Lab statistics need an N0f8 image")
                vals = samples
            else:
                vals = rgb8_to_lab(samples[:, :3]) if colourspace == "lab" else samples.astype(np.float64) / 255.0
            st = component_stats(vals)
            mean[c], mn[c], mx[c], sd[c] = st["mean"], st["min"], st["max"], st["sd"]
        return cls(mean, mn, mx, sd, tolerances, colourspace)


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _check_match(match, segmented):
    """-> match bits.  `match` must be a non-empty subset of (color, std, stripe) in that order (the order
    `_match_from_bools` produces, src/selectcolors.jl:661-668; the error is a left-folded mean over it)."""
    m = tuple(match)
    if not m or any(x not in MATCH_BITS for x in m) or list(m) != [x for x in DEFAULT_MATCH if x in m]:
        raise ValueError(f"match={m!r}: must be a non-empty subset of ('color', 'std', 'stripe') in that order")
    if segmented:
        raise ValueError("segmented=True (fast_scanning, src/categories.jl:53-67: a sequential scan with running "
                         "segment means) is not supported on the GPU path")
    return sum(MATCH_BITS[x] for x in m)


@dataclass
class TextureStats:
    """`std` and `stripe` entries of `_component_stats(::Vector{<:PixelProp})` (src/categories.jl:10-17) per category:
    mean / min / max / sd of the property at the clicked points; bounds as in src/categories.jl:131-133."""
    std: dict       # mean, min, max, sd: (K,)
    stripe: dict    # mean, min, max, sd: (K, 2)
    tolerances: np.ndarray

    def bounds(self):
        t = np.asarray(self.tolerances, dtype=np.float64).reshape(-1)
        with np.errstate(invalid="ignore"):
            w1, w2 = self.std["sd"] * t, self.stripe["sd"] * t[:, None]
            return (self.std["min"] - w1, self.std["max"] + w1), (self.stripe["min"] - w2, self.stripe["max"] + w2)

    @classmethod
    def from_points(cls, std_plane, stripe_plane, points, tolerances=None):
        """sample the planes at the clicked points (1-based (i, j), `round(Int, p)`); a plane that is None counts
        as the reference's initial all-zero plane (src/selectcolors.jl:229-231)"""
        K = len(points)
        tol = np.full(K, DEFAULT_CATEGORY_TOLERANCE) if tolerances is None else np.asarray(tolerances, dtype=np.float64)
        std = {k: np.full(K, np.nan) for k in ("mean", "min", "max", "sd")}
        stripe = {k: np.full((K, 2), np.nan) for k in ("mean", "min", "max", "sd")}
        for c, pts in enumerate(points):
            pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
            if pts.shape[0] == 0:
                continue
            ij = np.rint(pts).astype(np.int64)
            i, j = ij[:, 0] - 1, ij[:, 1] - 1
            a = component_stats(np.zeros(len(i)) if std_plane is None else std_plane[j, i])
            b = component_stats(np.zeros((len(i), 2)) if stripe_plane is None else stripe_plane[j, i])
            for k in std:
                std[k][c] = a[k][0]
                stripe[k][c] = b[k]
        return cls(std, stripe, tol)


def _upload_texture(ctx, tex):
    (slo, shi), (tlo, thi) = tex.bounds()
    ctx.set_texture_stats(tex.std["mean"], slo, shi, tex.stripe["mean"], tlo, thi)


def stds(A, *, ctx=None):
    """`_stds(A)` (src/selectcolors.jl:706-715): the "texture" plane of the blurred RGB{Float64} image."""
    if not _is_f64_img(A):
        raise ValueError("stds takes a (n2, n1, 3) float64 image (RGB{Float64}, the output of blur)")
    return (ctx or default_context()).stds_host(A)


def stripes(img, *, radius=3, ctx=None):
    """`_stripes(img; radius=3)` (src/stripes.jl:20-35): (n2, n1, 2) float64."""
    if not _is_f64_img(img):
        raise ValueError("stripes takes a (n2, n1, 3) float64 image (RGB{Float64}, the output of blur)")
    return (ctx or default_context()).stripes_host(img, radius)


def _as_palette(img, points_or_palette, tolerances, colourspace):
    if isinstance(points_or_palette, Palette):
        return points_or_palette
    return Palette.from_points(img, points_or_palette, tolerances, colourspace)


def _upload_palette(ctx, pal):
    if ctx.palette is not pal:
        ctx.set_palette(pal.mean, pal.lo, pal.hi, _lib.LAB if pal.colourspace == "lab" else _lib.RGB)
        ctx.palette = pal


def _known_from_points(points, n1):
    """`_set_categories!` (src/selectcolors.jl:670-681): later categories overwrite earlier ones."""
    plane = {}
    for c, pts in enumerate(points):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        for p in np.rint(pts).astype(np.int64):
            plane[(int(p[0]) - 1) + (int(p[1]) - 1) * n1] = c + 1
    idx = np.fromiter(plane.keys(), dtype=np.int64, count=len(plane))
    cat = np.fromiter(plane.values(), dtype=np.uint8, count=len(plane))
    return idx, cat


def _is_f64_img(img):
    return isinstance(img, np.ndarray) and img.dtype == np.float64 and img.ndim == 3 and img.shape[2] == 3


def _check_img(img, f64_ok=False):
    if f64_ok and _is_f64_img(img):
        return
    if not isinstance(img, np.ndarray) or img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4):
        raise ValueError("img must be a (n2, n1, 3|4) uint8 array (RGB{N0f8} / RGBA{N0f8}) or a (n2, n1, 3) float64 "
                         "array (RGB{Float64}, e.g. the output of blur); other element types are not supported")


def categorize_clean_u8(img, points_or_palette, *, tolerances=None, hood=Window(1), match=("color",),
                        segmented=False, scan_threshold=0.1, colourspace="rgb", use_known=False,
                        ctx=None, counters=False, std=None, stripe=None, texture=None):
    """Fused clean(categorize(img)) returning uint8 labels (n2, n1).
    A `match` with "std" / "stripe" categorises PixelProp pixels (src/categories.jl:45-48): img must be the
    RGB{Float64} image, `std` / `stripe` the planes of `stds` / `stripes` (None = the all-zero initial plane of
    the reference); `texture`: TextureStats (default: from the planes at the clicked points)."""
    del scan_threshold   # only used by the unsupported segmented branch
    bits = _check_match(match, segmented)
    _check_img(img, f64_ok=True)
    if bits != 1 and (not _is_f64_img(img) or colourspace != "rgb"):
        raise ValueError(f"match={tuple(match)!r}: std / stripe matching takes RGB{{Float64}} pixels (the blurred "
                         "image) and RGB statistics")
    ctx = ctx or default_context()
    pal = _as_palette(img, points_or_palette, tolerances, colourspace)
    _upload_palette(ctx, pal)
    if bits != 1:
        if counters:
            raise ValueError("counters are not available for RGB{Float64} pixels")
        n2, n1 = img.shape[:2]
        if (bits & 2) and std is None:
            std = np.zeros((n2, n1))
        if (bits & 4) and stripe is None:
            stripe = np.zeros((n2, n1, 2))
        if texture is None:
            if isinstance(points_or_palette, Palette):
                raise ValueError("a Palette needs `texture` (TextureStats) for std / stripe matching")
            texture = TextureStats.from_points(std if bits & 2 else None, stripe if bits & 4 else None,
                                               points_or_palette, pal.tolerances)
        _upload_texture(ctx, texture)
        if use_known and not isinstance(points_or_palette, Palette):
            ctx.set_known(*_known_from_points(points_or_palette, n1))
        else:
            ctx.set_known(None)
        R = hood.R if isinstance(hood, Window) else int(hood)
        return ctx.categorize_clean_props_host(img, std if bits & 2 else None, stripe if bits & 4 else None, bits, R)
    if use_known and not isinstance(points_or_palette, Palette):
        ctx.set_known(*_known_from_points(points_or_palette, img.shape[1]))
    else:
        ctx.set_known(None)
    R = hood.R if isinstance(hood, Window) else int(hood)
    if _is_f64_img(img):     # RGB{Float64} pixels (blurred / balanced): direct Float64 evaluation
        if pal.colourspace != "rgb":
            raise ValueError("RGB{Float64} pixels take RGB statistics")
        if counters:
            raise ValueError("counters are not available for RGB{Float64} pixels")
        return ctx.categorize_clean_f64_host(img, R)
    return ctx.categorize_clean_host(img, R, counters=counters)


def blur(A, *, hood=Window(1), repeat=1, ctx=None):
    """`blur(A; hood=Window{1}(), repeat=1)` (src/blur.jl:6-14): `repeat` passes of the Window mean in
    RGB{Float64}.  A: (n2, n1, 3) uint8 (RGB{N0f8}) or float64; returns (n2, n1, 3) float64."""
    _check_img(A, f64_ok=True)
    if A.shape[2] != 3:
        raise ValueError("blur takes RGB pixels")
    ctx = ctx or default_context()
    R = hood.R if isinstance(hood, Window) else int(hood)
    return ctx.blur_host(A, R, repeat)


def balance_fit(img, points, category_bitindex=None):
    """The host part of `_balance(A, points; category_bitindex)` (src/balance.jl:1-24): the table of colour
    deviations of the clicked points from their category's mean colour and the least-squares fit of
    `c ~ i^3 + i^2 + i + j^3 + j^2 + j` per channel.  Returns the 3 x 7 coefficients (intercept first), or
    None with fewer than 8 points (the reference then returns `A .* 1.0`, :16-18).
    (numpy's lstsq stands in for GLM's `lm`: same minimiser, last bits may differ.)"""
    img = np.asarray(img)
    n2, n1 = img.shape[:2]
    K = len(points)
    if category_bitindex is None:
        category_bitindex = [c == 0 for c in range(K)]       # `category_balance` default: category 1 only (selectcolors.jl:46)
    rows, resp = [], []
    for c, pts in enumerate(points):
        if not category_bitindex[c]:
            continue
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        if pts.shape[0] == 0:
            continue
        ij = np.rint(pts).astype(np.int64)
        cols = img[ij[:, 1] - 1, ij[:, 0] - 1, :3].astype(np.float64) / 255.0
        dev = cols - cols.mean(axis=0)
        for (i, j), d in zip(pts, dev):
            rows.append([1.0, i ** 3, i ** 2, i, j ** 3, j ** 2, j])
            resp.append(d)
    if len(rows) < 8:
        return None
    X, Y = np.asarray(rows), np.asarray(resp)
    beta, *_ = np.linalg.lstsq(X, Y, rcond=None)
    return np.ascontiguousarray(beta.T)                       # (3, 7)


def balance(img, points, *, category_bitindex=None, ctx=None):
    """`_balance(A, points; category_bitindex)` (src/balance.jl:1-42): detrend the image with the cubic
    polynomial fitted to the clicked points' colour deviations; RGB{Float64} out.  The fit is host code, the
    per-pixel prediction / mean / clamp runs on the GPU."""
    _check_img(img)
    if img.shape[2] != 3:
        raise ValueError("balance takes RGB{N0f8} pixels")
    ctx = ctx or default_context()
    coef = balance_fit(img, points, category_bitindex)
    if coef is None:
        return ctx.blur_host(img, 0, 0)          # A .* 1.0
    return ctx.balance_host(img, coef)


def categorize_clean(img, points_or_palette, **kw):
    """Same, eltype Int (Julia `Matrix{Int}`, src/categories.jl:69-73)."""
    out = categorize_clean_u8(img, points_or_palette, **kw)
    if isinstance(out, tuple):
        return out[0].astype(np.int64), out[1]
    return out.astype(np.int64)


def categorize_u8(img, points_or_palette, **kw):
    kw["hood"] = Window(0)
    return categorize_clean_u8(img, points_or_palette, **kw)


def categorize(img, points_or_palette, *, tolerances=None, match=("color",), segmented=False,
               scan_threshold=0.1, **kw):
    """`_categorize(pixels, points; tolerances, match, segmented, scan_threshold)`
    (src/categories.jl:49-74) -> labels, eltype Int, same shape as the image."""
    out = categorize_u8(img, points_or_palette, tolerances=tolerances, match=match, segmented=segmented,
                        scan_threshold=scan_threshold, **kw)
    if isinstance(out, tuple):
        return out[0].astype(np.int64), out[1]
    return out.astype(np.int64)


def clean_u8(labels, *, hood=Window(1), repeat=1, ctx=None):
    labels = np.asarray(labels)
    if labels.ndim != 2 or not np.issubdtype(labels.dtype, np.integer):
        raise ValueError("labels must be a 2-D integer array")
    if labels.size and (labels.min() < 0 or labels.max() > 254):
        raise ValueError("labels must be in 0..254")
    ctx = ctx or default_context()
    R = hood.R if isinstance(hood, Window) else int(hood)
    return ctx.clean_host(labels.astype(np.uint8), R, repeat)


def clean(labels, *, hood=Window(1), repeat=1, ctx=None):
    """`clean(labels; hood=Window{1}(), repeat=1)`: Window majority filter, SURVEY Appendix A.3."""
    return clean_u8(labels, hood=hood, repeat=repeat, ctx=ctx).astype(np.int64)


def clean_segments_u8(labels, known_categories=None, *, categories, line_threshold=0.01, prune_threshold=20,
                      ctx=None):
    """`_clean(A, known_categories; categories, line_threshold, prune_threshold)`
    (src/clean.jl:59-93; the reference's "clean" button, call site src/selectcolors.jl:479-489):
    4-connected segments of equal labels (fast_scanning!, src/scan.jl:56-157); segments of a category
    that is not kept, segments smaller than `prune_threshold` pixels and line-like segments
    (count / (max(h, w)^2 / 2) < line_threshold) become 0 -- or the known category clicked inside them.
    labels / known_categories: (n2, n1) integer arrays.  Returns uint8 labels.
    (The un-underscored name `clean` is taken by the north_star's Window majority filter.)"""
    labels = np.asarray(labels)
    if labels.ndim != 2 or not np.issubdtype(labels.dtype, np.integer):
        raise ValueError("labels must be a 2-D integer array")
    if labels.size and (labels.min() < 0 or labels.max() > 254):
        raise ValueError("labels must be in 0..254")
    keep = np.zeros(256, dtype=np.uint8)
    cats = np.asarray(list(categories), dtype=np.int64)
    if cats.size and (cats.min() < 0 or cats.max() > 254):
        raise ValueError("categories must be in 0..254")
    keep[cats] = 1
    ctx = ctx or default_context()
    if known_categories is None:
        ctx.set_known(None)
    else:
        kn = np.asarray(known_categories)
        if kn.shape != labels.shape:
            raise ValueError("known_categories must have the shape of labels")
        idx = np.flatnonzero(kn)   # lin_idx = i + j*n1 of the (n2, n1) line-major view
        ctx.set_known(idx, kn.reshape(-1)[idx].astype(np.uint8))
    try:
        out, _ = ctx.clean_segments_host(labels.astype(np.uint8), keep, line_threshold, prune_threshold)
    finally:
        if known_categories is not None:
            ctx.set_known(None)
    return out


def clean_segments(labels, known_categories=None, **kw):
    """Same, eltype Int."""
    return clean_segments_u8(labels, known_categories, **kw).astype(np.int64)


def fill_gaps_u8(src, *, categories, stats, original, known=None, mask=None, match=("color",), missingval=0,
                 neighborhood="VonNeumann{2}", repeat=1, ctx=None):
    """`fill_gaps(src; categories, neighborhood = VonNeumann{2}(), missingval, stats, match, known, original,
    mask)` (src/clean.jl:2-54), applied `repeat` times as `_fill` does (src/selectcolors.jl:683-699).
    src / known / mask: (n2, n1) arrays; original: the (n2, n1, 3|4) uint8 image the categoriser saw;
    stats: the `Palette` (`stats[i]` of clean.jl:29-30 is the row of `categories[i]`); categories: the kept
    categories in order.  Only `missingval = 0`, `VonNeumann{2}` and `match = ("color",)` exist on the GPU
    path (what `_fill` passes, minus the texture planes)."""
    if missingval != 0:
        raise ValueError("fill_gaps: only missingval = 0 (what _fill passes) is supported")
    if str(neighborhood).replace(" ", "") not in ("VonNeumann{2}", "VonNeumann{2}()"):
        raise ValueError("fill_gaps: only the VonNeumann{2} neighbourhood is supported")
    if tuple(match) != ("color",):
        raise ValueError("fill_gaps: only match = ('color',) is supported")
    if not isinstance(stats, Palette):
        raise ValueError("stats must be a Palette")
    src = np.asarray(src)
    if src.ndim != 2 or not np.issubdtype(src.dtype, np.integer):
        raise ValueError("src must be a 2-D integer array")
    if src.size and (src.min() < 0 or src.max() > 254):
        raise ValueError("labels must be in 0..254")
    _check_img(original)
    if original.shape[:2] != src.shape:
        raise ValueError("original must have the shape of src")
    cats = np.asarray(list(categories), dtype=np.int64)
    if cats.size and (cats.min() < 1 or cats.max() > stats.K):
        raise ValueError("categories must be in 1..K")
    ctx = ctx or default_context()
    _upload_palette(ctx, stats)
    if known is None:
        ctx.set_known(None)
    else:
        kn = np.asarray(known)
        if kn.shape != src.shape:
            raise ValueError("known must have the shape of src")
        idx = np.flatnonzero(kn)
        ctx.set_known(idx, kn.reshape(-1)[idx].astype(np.uint8))
    try:
        return ctx.fill_gaps_host(src.astype(np.uint8), original, cats.astype(np.uint8),
                                  None if mask is None else (np.asarray(mask) != 0).astype(np.uint8), repeat)
    finally:
        if known is not None:
            ctx.set_known(None)


def fill_gaps(src, **kw):
    """Same, eltype Int."""
    return fill_gaps_u8(src, **kw).astype(np.int64)


def _labels_u8(labels):
    labels = np.asarray(labels)
    if labels.ndim != 2 or not np.issubdtype(labels.dtype, np.integer):
        raise ValueError("labels must be a 2-D integer array")
    if labels.size and (labels.min() < 0 or labels.max() > 254):
        raise ValueError("labels must be in 0..254")
    return labels.astype(np.uint8)


def mean_color(A, pointvec):
    """`_mean_color(A, pointvec)` (src/selectcolors.jl:701-704): mean colour of the clicked points; zero colour when the
    category has no points"""
    pal = Palette.from_points(np.asarray(A), [pointvec])
    return np.zeros(3) if np.isinf(pal.mean[0, 0]) else pal.mean[0, :3].copy()


def recolor(categories, source, points, *, ctx=None):
    """`_recolor(categories, source, points)` (src/selectcolors.jl:650-655): the RGBA{Float64} display layer of a label
    image: the category's mean colour at its clicked points (alpha 1), transparent black for label 0.
    Returns (n2, n1, 4) float64."""
    colors = np.stack([mean_color(source, ps) for ps in points]) if len(points) else np.zeros((0, 3))
    if colors.shape[0] == 0:
        raise ValueError("recolor needs at least one category")
    return (ctx or default_context()).recolor_host(_labels_u8(categories), colors)


def shrink_grow(A, i, j, *, missingval=0, ctx=None):
    """`_shrink_grow(A, i, j; missingval=0)` (src/clean.jl:132-140): `i` categorical erosions (`_erode`, :121-129), then
    `j` categorical dilations (`_dilate`, :96-118).  Returns labels, eltype Int."""
    if missingval != 0:
        raise ValueError("shrink_grow: only missingval = 0 is supported")
    if i < 0 or j < 0:
        raise ValueError("pass counts must be >= 0")
    return (ctx or default_context()).shrink_grow_host(_labels_u8(A), i, j).astype(np.int64)


def erode(A, *, missingval=0, ctx=None):
    """`_erode(A; missingval=0)` (src/clean.jl:121-129)"""
    return shrink_grow(A, 1, 0, missingval=missingval, ctx=ctx)


def dilate(A, *, missingval=0, ctx=None):
    """`_dilate(A; missingval=0)` (src/clean.jl:96-118)"""
    return shrink_grow(A, 0, 1, missingval=missingval, ctx=ctx)


def _rings_of(polygons):
    """`polygons`: per category a list of rings (src/selectcolors.jl:35: Vector{Vector{Vector{Point2}}}) ->
    (rings, category of every ring)"""
    rings, cats = [], []
    for c, polys in enumerate(polygons or []):
        for r in polys:
            rings.append(r)
            cats.append(c + 1)
    return rings, cats


def polymask(polygons, shape, *, ctx=None):
    """`_polymask(polygons, A)` (src/selectcolors.jl:559-563): boolean mask (n2, n1) of the pixels inside any polygon."""
    rings, _ = _rings_of(polygons)
    return (ctx or default_context()).polygons_host(rings, tuple(shape)) != 0


def paint_polygons(labels, polygons, *, fill=None, ctx=None):
    """`rasterize!(raster, p; fill = i, shape = :polygon)` for the polygons of every category i
    (src/selectcolors.jl:494-496); `fill=0` paints zeros: `(1 .- mask) .* labels` (:481).  Returns uint8 labels."""
    rings, cats = _rings_of(polygons)
    labels = np.asarray(labels)
    if labels.ndim != 2 or not np.issubdtype(labels.dtype, np.integer):
        raise ValueError("labels must be a 2-D integer array")
    if labels.size and (labels.min() < 0 or labels.max() > 254):
        raise ValueError("labels must be in 0..254")
    f = np.asarray(cats, dtype=np.uint8) if fill is None else np.full(len(rings), fill, dtype=np.uint8)
    return (ctx or default_context()).polygons_host(rings, labels.shape, fill=f, plane=labels.astype(np.uint8))


@dataclass
class MapSelection:
    """`MapSelection` (src/selectcolors.jl:3-8): what `selectcolors` returns -- the clicked points per
    category and the widget settings (src/selectcolors.jl:111-127: `category_tolerances`,
    `category_keep`, `prune_threshold`, `line_threshold`, `match`, `segment`, ...)."""
    settings: dict
    removals: list = field(default_factory=list)
    points: list = field(default_factory=list)
    output: object = None
    polygons: list = field(default_factory=list)   # per category a list of rings (src/selectcolors.jl:35, :124)

    def tolerances(self):
        t = self.settings.get("category_tolerances")
        return None if t is None else np.asarray(t, dtype=np.float64)

    def kept_categories(self):
        """`_kept_categories` (src/selectcolors.jl:189): 1-based indices of the keep toggles that are on."""
        keep = self.settings.get("category_keep")
        if keep is None:
            return tuple(range(1, len(self.points) + 1))
        return tuple(i + 1 for i, k in enumerate(keep) if k)


def rasterize(img, sel, *, hood=None, colourspace="rgb", ctx=None):
    """The button sequence of `selectcolors` without the GUI, on the GPU path:
    "categorize" (src/selectcolors.jl:466-478: `_categorize(pixels, points; tolerances, match,
    segmented)`, known-category points from `_set_categories!`, :670-681) -> optional Window
    majority filter (`hood`; north_star's `clean`) -> "clean" (:479-489: `_clean(A, known;
    categories = kept, line_threshold, prune_threshold)`).
    -> "fill" (:490-501: `_fill` = `fill_gaps` `fill_repeat` times, :683-699; the polygon mask is the caller's).
    Returns a dict with the uint8 label images `categorized`, `cleaned` and `filled`.
    `blur_repeat > 0` runs `blur` first and categorises its RGB{Float64} output (statistics from the
    blurred image, as the reference does); the tie-break of `fill_gaps` then still reads the unblurred pixel.
    A `match` with "std" / "stripe" computes the texture planes from the blurred image (`stds`, `stripes` with
    `stripe_radius`) and categorises PixelProp pixels (src/categories.jl:45-48).
    Settings that select unsupported stages raise ValueError (no CPU fallback): `segment=True`."""
    if not isinstance(sel, MapSelection):
        raise ValueError("sel must be a MapSelection")
    st = sel.settings
    match = tuple(st.get("match", ("color",)))
    bits = _check_match(match, bool(st.get("segment", False)))
    _check_img(img)
    ctx = ctx or default_context()
    n2, n1 = img.shape[:2]
    tol = sel.tolerances()
    R = 0 if hood is None else (hood.R if isinstance(hood, Window) else int(hood))
    # `balanced = _balance(raw, points; category_bitindex = category_balance)` (src/selectcolors.jl:222; default:
    # the points of category 1), then the "blur" button (:455-465): blur(balanced; repeat = blur_repeat).
    # With fewer than 8 balance points `_balance` is A .* 1.0 (src/balance.jl:16-18): the N0f8 path.
    blur_repeat = int(st.get("blur_repeat", 0))
    pixels = img
    coef = None
    if img.shape[2] == 3 and colourspace == "rgb":
        coef = balance_fit(img, sel.points, st.get("category_balance"))
    if coef is not None:
        pixels = ctx.balance_host(img, coef)
    if blur_repeat > 0:
        if img.shape[2] != 3 or colourspace != "rgb":
            raise ValueError("blur_repeat > 0 needs RGB{N0f8} pixels and RGB matching")
        pixels = blur(pixels, repeat=blur_repeat, ctx=ctx)
    std_plane = stripe_plane = None
    if bits != 1:
        # the "blur" button also refreshes the texture planes of the matched properties (src/selectcolors.jl:457-462)
        if img.shape[2] != 3 or colourspace != "rgb":
            raise ValueError("std / stripe matching needs RGB{N0f8} pixels and RGB matching")
        if not _is_f64_img(pixels):
            pixels = ctx.blur_host(img, 0, 0)          # A .* 1.0 (src/balance.jl:16-18)
        if bits & 2:
            std_plane = stds(pixels, ctx=ctx)
        if bits & 4:
            stripe_plane = stripes(pixels, radius=int(st.get("stripe_radius", DEFAULT_STRIPE_RADIUS)), ctx=ctx)
    categorized = categorize_clean_u8(pixels, sel.points, tolerances=tol, hood=Window(R), colourspace=colourspace,
                                      use_known=True, ctx=ctx, match=match, std=std_plane, stripe=stripe_plane)
    idx, cat = _known_from_points(sel.points, n1)
    known = np.zeros((n2, n1), dtype=np.uint8)
    known.reshape(-1)[idx] = cat
    # "clean" (src/selectcolors.jl:479-489): masked = (1 .- mask) .* categorized, mask = the drawn polygons
    polygons = sel.polygons or st.get("polygons")        # the reference keeps them in `settings` (:124)
    have_polys = any(len(r) >= 3 for polys in (polygons or []) for r in polys)
    mask = polymask(polygons, (n2, n1), ctx=ctx) if have_polys else None
    masked = paint_polygons(categorized, polygons, fill=0, ctx=ctx) if have_polys else categorized
    cleaned = clean_segments_u8(masked, known, categories=sel.kept_categories(),
                                line_threshold=float(st.get("line_threshold", 0.01)),
                                prune_threshold=float(st.get("prune_threshold", 20)), ctx=ctx)
    # `_component_stats(original, points[categories])` (selectcolors.jl:689): the rows of the kept categories
    pal = ctx.palette if isinstance(ctx.palette, Palette) else _as_palette(img, sel.points, tol, colourspace)
    # "fill" (:490-501): fill_gaps with the polygon mask, then every polygon painted with its category
    filled = fill_gaps_u8(cleaned, categories=sel.kept_categories(), stats=pal, original=img, known=known,
                          mask=mask, repeat=int(st.get("fill_repeat", 1)), ctx=ctx)
    if have_polys:
        filled = paint_polygons(filled, polygons, ctx=ctx)
    return {"categorized": categorized, "cleaned": cleaned, "filled": filled}
