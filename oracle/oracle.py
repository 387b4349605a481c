"""ctypes loader for the CPU ORACLE (test infrastructure, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It wraps oracle/libmaprast_oracle.so (built by
oracle/Makefile from oracle/maprast_oracle.c, the plain-C restatement of
/root/reference/src/categories.jl:76-133 + SURVEY.md Appendix A.3/A.4).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmaprast_oracle.so")

RGB, LAB = 0, 1
TIE_LOWEST, TIE_CENTRE_IF_MODAL = 0, 1

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = C.c_void_p
_i64 = C.c_int64


def build(force=False):
    src = os.path.join(_HERE, "maprast_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.mro_bounds.argtypes = [C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.mro_component_stats.argtypes = [_f64p, C.c_int, _f64p]
        L.mro_categorize_f64.argtypes = [_f64p, C.c_int, C.c_int, _f64p, _f64p, _f64p, C.c_int]
        L.mro_categorize_f64.restype = C.c_int
        L.mro_n0f8.argtypes = [C.c_int]
        L.mro_n0f8.restype = C.c_double
        L.mro_cbrt.argtypes = [C.c_double]
        L.mro_cbrt.restype = C.c_double
        L.mro_rgb8_to_lab.argtypes = [C.c_int, C.c_int, C.c_int, _f64p]
        L.mro_categorize_u8.argtypes = [_u8p, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                        _f64p, _f64p, _f64p, _i64, C.c_void_p, C.c_void_p,
                                        _u8p, _i64, C.c_int]
        L.mro_build_lut.argtypes = [C.c_int, C.c_int, _f64p, _f64p, _f64p, _u8p, C.c_int]
        L.mro_majority.argtypes = [_u8p, _i64, _i64, _i64, C.c_int, C.c_int, _i64, _i64,
                                   _u8p, _i64, C.c_int]
        L.mro_categorize_clean.argtypes = [_u8p, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_int,
                                           _u8p, _i64, C.c_int]
        L.mro_philox4x32.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        L.mro_synth_scan.argtypes = [C.c_uint32, C.c_uint32, C.c_int, _u8p, _i64, _i64, _i64, _i64,
                                     C.c_int, _u8p, _i64, C.c_int]
        L.mro_clean_segments.argtypes = [_u8p, _i64, _i64, _i64, C.c_void_p, _i64, _u8p, C.c_double,
                                         C.c_double, _u8p, _i64]
        L.mro_clean_segments.restype = _i64
        L.mro_fill_gaps.argtypes = [_u8p, _i64, _i64, _i64, _u8p, _i64, C.c_int, C.c_int, C.c_int, _f64p,
                                    C.c_void_p, _i64, C.c_void_p, _i64, _u8p, C.c_int, _u8p, _i64]
        L.mro_blur.argtypes = [C.c_void_p, C.c_int, _i64, _i64, _i64, C.c_int, C.c_void_p, _i64]
        L.mro_categorize_f64_image.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_int, _f64p, _f64p, _f64p, _i64,
                                               C.c_void_p, C.c_void_p, _u8p, _i64]
        L.mro_balance.argtypes = [_u8p, _i64, _i64, _i64, _f64p, C.c_void_p, _i64, C.c_void_p]
        L.mro_recolor.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_int, _f64p, C.c_void_p, _i64]
        L.mro_erode.argtypes = [_u8p, _i64, _i64, _i64, _u8p, _i64]
        L.mro_dilate.argtypes = [_u8p, _i64, _i64, _i64, _u8p, _i64]
        L.mro_polygons.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _i64, _i64, C.c_void_p, _i64]
        L.mro_stds.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_void_p, _i64]
        L.mro_stripes.argtypes = [C.c_void_p, _i64, _i64, _i64, C.c_int, C.c_void_p, _i64]
        L.mro_categorize_props_image.argtypes = [C.c_void_p, _i64, C.c_void_p, _i64, C.c_void_p, _i64, _i64, _i64,
                                                 C.c_int, C.c_int] + [_f64p] * 9 + [_i64, C.c_void_p, C.c_void_p,
                                                                                    C.c_void_p, _i64]
        L.mro_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(lib().mro_max_threads())


def bounds(mn, mx, sd, tol):
    """lo = min - sd*tol, hi = max + sd*tol (categories.jl:131-133); arrays broadcast to K x C."""
    mn, mx, sd = _f64(mn), _f64(mx), _f64(sd)
    tol = _f64(np.broadcast_to(np.asarray(tol, dtype=np.float64).reshape(-1, *([1] * (mn.ndim - 1))), mn.shape))
    lo, hi = np.empty_like(mn), np.empty_like(mn)
    lib().mro_bounds(mn.size, mn.ravel(), mx.ravel(), sd.ravel(), tol.ravel(), lo.ravel(), hi.ravel())
    return lo, hi


def component_stats(v):
    v = _f64(v).ravel()
    out = np.empty(4, dtype=np.float64)
    lib().mro_component_stats(v, v.size, out)
    return out


def categorize_f64(x, mean, lo, hi, known=0):
    x, mean, lo, hi = _f64(x), _f64(mean), _f64(lo), _f64(hi)
    K, Cn = mean.shape
    return int(lib().mro_categorize_f64(x, Cn, K, mean, lo, hi, known))


def rgb8_to_lab(r, g, b):
    out = np.empty(3, dtype=np.float64)
    lib().mro_rgb8_to_lab(int(r), int(g), int(b), out)
    return out


def _img(px):
    """px: (n2, n1, bpp) uint8 C-contiguous (line-major view of Julia's n1 x n2 column-major)."""
    px = np.ascontiguousarray(px, dtype=np.uint8)
    assert px.ndim == 3 and px.shape[2] in (3, 4)
    return px


def categorize_u8(px, mean, lo, hi, colourspace=RGB, channels=3, known_idx=None, known_cat=None,
                  nthreads=0):
    px = _img(px)
    n2, n1, bpp = px.shape
    mean, lo, hi = _f64(mean), _f64(lo), _f64(hi)
    out = np.empty((n2, n1), dtype=np.uint8)
    nk = 0
    ki = kc = None
    if known_idx is not None and len(known_idx):
        ki = np.ascontiguousarray(known_idx, dtype=np.int64)
        kc = np.ascontiguousarray(known_cat, dtype=np.uint8)
        nk = ki.size
    lib().mro_categorize_u8(px.ctypes.data, n1, n2, n1 * bpp, bpp, channels, colourspace,
                            mean.shape[0], mean, lo, hi, nk,
                            ki.ctypes.data if nk else None, kc.ctypes.data if nk else None,
                            out.ctypes.data, n1, nthreads)
    return out


def build_lut(mean, lo, hi, colourspace=RGB, nthreads=0):
    mean, lo, hi = _f64(mean), _f64(lo), _f64(hi)
    lut = np.empty(1 << 24, dtype=np.uint8)
    lib().mro_build_lut(colourspace, mean.shape[0], mean, lo, hi, lut.ctypes.data, nthreads)
    return lut


def majority(labels, R, tie_rule=TIE_LOWEST, halo_lo=0, halo_hi=0, nthreads=0):
    """labels: (n2 + halo_lo + halo_hi, n1) uint8; returns the (n2, n1) filtered core."""
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    nt, n1 = labels.shape
    n2 = nt - halo_lo - halo_hi
    out = np.empty((n2, n1), dtype=np.uint8)
    base = labels.ctypes.data + halo_lo * n1
    lib().mro_majority(base, n1, n2, n1, R, tie_rule, halo_lo, halo_hi, out.ctypes.data, n1, nthreads)
    return out


def categorize_clean(px, mean, lo, hi, R, colourspace=RGB, channels=3, tie_rule=TIE_LOWEST, nthreads=0):
    px = _img(px)
    n2, n1, bpp = px.shape
    mean, lo, hi = _f64(mean), _f64(lo), _f64(hi)
    out = np.empty((n2, n1), dtype=np.uint8)
    lib().mro_categorize_clean(px.ctypes.data, n1, n2, n1 * bpp, bpp, channels, colourspace,
                               mean.shape[0], mean, lo, hi, R, tie_rule, out.ctypes.data, n1, nthreads)
    return out


def keep_table(categories):
    """256-entry table: 1 where the category is kept (`categories` of _clean, src/clean.jl:59)."""
    t = np.zeros(256, dtype=np.uint8)
    t[np.asarray(list(categories), dtype=np.int64)] = 1
    return t


def clean_segments(labels, known=None, categories=(), line_threshold=0.01, prune_threshold=20):
    """`_clean(A, known; categories, line_threshold, prune_threshold)` (src/clean.jl:59-93),
    labels / known: (n2, n1) uint8 (line-major view).  Returns (out, n_segments)."""
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    n2, n1 = labels.shape
    out = np.empty_like(labels)
    kn = None
    if known is not None:
        kn = np.ascontiguousarray(known, dtype=np.uint8)
        assert kn.shape == labels.shape
    keep = keep_table(categories)
    n = lib().mro_clean_segments(labels.ctypes.data, n1, n2, n1, kn.ctypes.data if kn is not None else None, n1,
                                 keep.ctypes.data, float(line_threshold), float(prune_threshold),
                                 out.ctypes.data, n1)
    return out, int(n)


def fill_gaps(labels, px, mean, categories, known=None, mask=None, repeat=1, colourspace=RGB):
    """`fill_gaps` (src/clean.jl:2-54) applied `repeat` times as `_fill` does (src/selectcolors.jl:692-696).
    labels / known / mask: (n2, n1) uint8, px: (n2, n1, bpp) uint8, mean: K x C, categories: ordered kept list."""
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    px = _img(px)
    n2, n1 = labels.shape
    assert px.shape[:2] == (n2, n1)
    mean = _f64(mean)
    cats = np.ascontiguousarray(list(categories), dtype=np.uint8)
    kn = np.ascontiguousarray(known, dtype=np.uint8) if known is not None else None
    mk = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    cur = labels
    for _ in range(repeat):
        out = np.empty_like(cur)
        lib().mro_fill_gaps(cur.ctypes.data, n1, n2, n1, px.ctypes.data, n1 * px.shape[2], px.shape[2],
                            mean.shape[1], colourspace, mean, kn.ctypes.data if kn is not None else None, n1,
                            mk.ctypes.data if mk is not None else None, n1, cats.ctypes.data, cats.size,
                            out.ctypes.data, n1)
        cur = out
    return cur if repeat > 0 else labels.copy()


def blur(img, R=1, repeat=1):
    """`blur(A; hood=Window{R}(), repeat)` (src/blur.jl:6-14): img (n2, n1, 3) uint8 (N0f8) or float64 ->
    (n2, n1, 3) float64."""
    cur = np.ascontiguousarray(img)
    assert cur.ndim == 3 and cur.shape[2] == 3 and cur.dtype in (np.uint8, np.float64)
    n2, n1, _ = cur.shape
    if repeat == 0:
        return cur.astype(np.float64) / 255.0 if cur.dtype == np.uint8 else cur.copy()
    for _ in range(repeat):
        out = np.empty((n2, n1, 3), dtype=np.float64)
        u8 = cur.dtype == np.uint8
        lib().mro_blur(cur.ctypes.data, 1 if u8 else 0, n1 * (3 if u8 else 24), n1, n2, int(R), out.ctypes.data, n1 * 24)
        cur = out
    return cur


def categorize_f64_image(px, mean, lo, hi, known_idx=None, known_cat=None):
    """`_categorize` over an RGB{Float64} image: px (n2, n1, 3) float64 -> (n2, n1) uint8 labels."""
    px = np.ascontiguousarray(px, dtype=np.float64)
    n2, n1, _ = px.shape
    mean, lo, hi = _f64(mean), _f64(lo), _f64(hi)
    out = np.empty((n2, n1), dtype=np.uint8)
    nk = 0
    ki = kc = None
    if known_idx is not None and len(known_idx):
        ki = np.ascontiguousarray(known_idx, dtype=np.int64)
        kc = np.ascontiguousarray(known_cat, dtype=np.uint8)
        nk = ki.size
    lib().mro_categorize_f64_image(px.ctypes.data, n1, n2, n1 * 24, mean.shape[0], mean, lo, hi, nk,
                                   ki.ctypes.data if nk else None, kc.ctypes.data if nk else None, out.ctypes.data, n1)
    return out


def recolor(labels, colors):
    """`_recolor` (src/selectcolors.jl:650-655): labels (n2, n1) uint8, colors (K, 3) -> (n2, n1, 4) float64 RGBA"""
    labels = np.ascontiguousarray(labels, dtype=np.uint8)
    colors = _f64(colors).reshape(-1, 3)
    n2, n1 = labels.shape
    out = np.empty((n2, n1, 4), dtype=np.float64)
    lib().mro_recolor(labels.ctypes.data, n1, n2, n1, colors.shape[0], colors, out.ctypes.data, n1 * 32)
    return out


def shrink_grow(labels, n_erode, n_dilate):
    """`_shrink_grow(A, i, j)` (src/clean.jl:132-140): i passes of `_erode`, then j passes of `_dilate`."""
    cur = np.ascontiguousarray(labels, dtype=np.uint8)
    n2, n1 = cur.shape
    for k in range(n_erode + n_dilate):
        out = np.empty_like(cur)
        (lib().mro_erode if k < n_erode else lib().mro_dilate)(cur.ctypes.data, n1, n2, n1, out.ctypes.data, n1)
        cur = out
    return cur.copy()


def flatten_rings(rings):
    """list of rings (each a sequence of (d1, d2) vertices) -> (offsets int32 (n + 1), xy float64 (nv, 2))"""
    off = np.zeros(len(rings) + 1, dtype=np.int32)
    pts = []
    for k, r in enumerate(rings):
        r = np.asarray(r, dtype=np.float32).astype(np.float64).reshape(-1, 2)     # Point2{Float32} vertices
        pts.append(r)
        off[k + 1] = off[k] + r.shape[0]
    xy = np.ascontiguousarray(np.concatenate(pts) if pts else np.zeros((0, 2)))
    return off, xy


def polygons(rings, shape, fill=None, plane=None):
    """mask (fill None) or painted label plane of the rings; shape = (n2, n1)"""
    off, xy = flatten_rings(rings)
    n2, n1 = shape
    out = np.zeros((n2, n1), dtype=np.uint8) if plane is None else np.ascontiguousarray(plane, dtype=np.uint8).copy()
    f = None if fill is None else np.ascontiguousarray(fill, dtype=np.uint8)
    lib().mro_polygons(len(rings), off.ctypes.data, xy.ctypes.data, None if f is None else f.ctypes.data, n1, n2,
                       out.ctypes.data, n1)
    return out


def stds(px):
    """`_stds(A)` (src/selectcolors.jl:706-715): px (n2, n1, 3) float64 -> (n2, n1) float64."""
    px = np.ascontiguousarray(px, dtype=np.float64)
    n2, n1, _ = px.shape
    out = np.empty((n2, n1), dtype=np.float64)
    lib().mro_stds(px.ctypes.data, n1, n2, n1 * 24, out.ctypes.data, n1 * 8)
    return out


def stripes(px, radius=3):
    """`_stripes(img; radius)` (src/stripes.jl:20-35): px (n2, n1, 3) float64 -> (n2, n1, 2) float64."""
    px = np.ascontiguousarray(px, dtype=np.float64)
    n2, n1, _ = px.shape
    out = np.empty((n2, n1, 2), dtype=np.float64)
    lib().mro_stripes(px.ctypes.data, n1, n2, n1 * 24, int(radius), out.ctypes.data, n1 * 16)
    return out


MATCH_BITS = {"color": 1, "std": 2, "stripe": 4}


def categorize_props_image(px, sd, stripe, match, mean, lo, hi, sd_stats, st_stats, known_idx=None, known_cat=None):
    """`_categorize` over PixelProp pixels with a match tuple (src/categories.jl:76-133).  sd_stats / st_stats:
    (mean, lo, hi) of shape (K,) / (K, 2).  -> (n2, n1) uint8 labels."""
    px = np.ascontiguousarray(px, dtype=np.float64)
    n2, n1, _ = px.shape
    bits = sum(MATCH_BITS[m] for m in match)
    sd = None if sd is None else np.ascontiguousarray(sd, dtype=np.float64)
    stripe = None if stripe is None else np.ascontiguousarray(stripe, dtype=np.float64)
    mean, lo, hi = _f64(mean), _f64(lo), _f64(hi)
    K = mean.shape[0]
    sdm, sdl, sdh = (_f64(np.asarray(a, dtype=np.float64).reshape(K)) for a in sd_stats)
    stm, stl, sth = (_f64(np.asarray(a, dtype=np.float64).reshape(K, 2)) for a in st_stats)
    out = np.empty((n2, n1), dtype=np.uint8)
    nk = 0
    ki = kc = None
    if known_idx is not None and len(known_idx):
        ki = np.ascontiguousarray(known_idx, dtype=np.int64)
        kc = np.ascontiguousarray(known_cat, dtype=np.uint8)
        nk = ki.size
    lib().mro_categorize_props_image(px.ctypes.data, n1 * 24, None if sd is None else sd.ctypes.data, n1 * 8,
                                     None if stripe is None else stripe.ctypes.data, n1 * 16, n1, n2, bits, K,
                                     mean, lo, hi, sdm, sdl, sdh, stm, stl, sth, nk,
                                     ki.ctypes.data if nk else None, kc.ctypes.data if nk else None,
                                     out.ctypes.data, n1)
    return out


def balance(img, coef):
    """per-pixel part of `_balance` (src/balance.jl:25-42): img (n2, n1, 3) uint8, coef (3, 7) -> (n2, n1, 3) float64."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    n2, n1, _ = img.shape
    coef = _f64(coef).reshape(3, 7)
    out = np.empty((n2, n1, 3), dtype=np.float64)
    lib().mro_balance(img.ctypes.data, n1, n2, n1 * 3, coef, out.ctypes.data, n1 * 24, None)
    return out


def philox4x32(c, k):
    out = (C.c_uint32 * 4)()
    lib().mro_philox4x32(c[0], c[1], c[2], c[3], k[0], k[1], out)
    return [int(v) for v in out]


def synth_scan(seed, sheet_id, palette_rgb, n1, n2, line0=0, nlines=None, mode=0, nthreads=0):
    pal = np.ascontiguousarray(palette_rgb, dtype=np.uint8)
    if nlines is None:
        nlines = n2 - line0
    out = np.empty((nlines, n1, 3), dtype=np.uint8)
    lib().mro_synth_scan(seed, sheet_id, pal.shape[0], pal.ctypes.data, n1, n2, line0, nlines,
                         mode, out.ctypes.data, n1 * 3, nthreads)
    return out
