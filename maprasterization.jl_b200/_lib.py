"""ctypes binding of libmaprast.so (the C ABI in include/maprast.h).

This is the same surface a Julia `ccall` wrapper binds.  There is no CPU
fallback: if the shared library is missing or no CUDA device is usable, loading / creating a
context raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MAPRAST_LIB selects another build of the same library (kernel experiments: csrc/Makefile VARIANT=...)
LIB_PATH = os.environ.get("MAPRAST_LIB") or os.path.join(_HERE, "libmaprast.so")

MR_OK, MR_ERR_ARG, MR_ERR_CUDA, MR_ERR_STATE, MR_ERR_NOMEM = 0, -1, -2, -3, -4
RGB, LAB = 0, 1
LUT_SIZE = 1 << 24

# every symbol include/maprast.h declares (tests check the library exports all of them)
SYMBOLS = [
    "mr_version", "mr_create", "mr_destroy", "mr_last_error", "mr_set_palette", "mr_set_known",
    "mr_get_lut", "mr_palette_build_ms", "mr_enable_counters", "mr_read_counters", "mr_launch_count",
    "mr_categorize_clean_host", "mr_categorize_clean_host_band", "mr_categorize_clean_dev", "mr_batch_categorize_clean_dev",
    "mr_categorize_clean_dev_peer", "mr_ipc_export", "mr_ipc_open", "mr_ipc_close",
    "mr_clean_host", "mr_clean_dev", "mr_clean_dev_bounded", "mr_clean_segments_dev", "mr_clean_segments_host",
    "mr_fill_gaps_dev", "mr_fill_gaps_host", "mr_set_known_band",
    "mr_blur_host", "mr_blur_dev", "mr_categorize_clean_f64_host", "mr_categorize_clean_f64_dev",
    "mr_balance_host", "mr_balance_dev",
    "mr_stds_host", "mr_stds_dev", "mr_stripes_host", "mr_stripes_dev", "mr_set_texture_stats",
    "mr_categorize_clean_props_host", "mr_categorize_clean_props_dev", "mr_polygons_host", "mr_polygons_dev",
    "mr_shrink_grow_host", "mr_shrink_grow_dev", "mr_recolor_host", "mr_recolor_dev",
    "mr_multi_create", "mr_multi_destroy", "mr_multi_last_error", "mr_multi_device_count", "mr_multi_ctx",
    "mr_multi_band_range", "mr_multi_set_palette", "mr_multi_set_known", "mr_multi_categorize_clean_host",
    "mr_multi_batch_categorize_clean_host", "mr_multi_categorize_clean_dev",
    "mr_host_alloc", "mr_host_free", "mr_synth_scan_dev",
]


class Counters(C.Structure):
    _fields_ = [("n_pixels", C.c_uint64), ("n_nomatch", C.c_uint64), ("n_fine", C.c_uint64),
                ("n_changed", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class MapRastError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmaprast error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libmaprast.so or raise (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  maprast-b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u8p = C.c_void_p, C.c_int64, C.c_int, C.c_void_p
    L.mr_version.restype = i32
    L.mr_create.argtypes = [i32, C.POINTER(i32)]
    L.mr_create.restype = vp
    L.mr_destroy.argtypes = [vp]
    L.mr_destroy.restype = None
    L.mr_last_error.argtypes = [vp]
    L.mr_last_error.restype = C.c_char_p
    L.mr_set_palette.argtypes = [vp, i32, vp, vp, vp, i32, i32]
    L.mr_set_known.argtypes = [vp, i64, vp, vp]
    L.mr_get_lut.argtypes = [vp, u8p]
    L.mr_palette_build_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.mr_enable_counters.argtypes = [vp, i32]
    L.mr_read_counters.argtypes = [vp, C.POINTER(Counters), i32]
    L.mr_launch_count.argtypes = [vp]
    L.mr_launch_count.restype = i64
    L.mr_categorize_clean_host.argtypes = [vp, u8p, i64, i64, i64, i32, i32, u8p, i64, C.POINTER(Counters)]
    L.mr_categorize_clean_host_band.argtypes = [vp, u8p, i64, i64, i64, i32, i32, i32, i32, u8p, i64,
                                                C.POINTER(Counters)]
    L.mr_categorize_clean_dev.argtypes = [vp, u8p, i64, i64, i64, i32, i32, i32, i32, u8p, i64, vp]
    L.mr_batch_categorize_clean_dev.argtypes = [vp, i32, u8p, i64, i64, i64, i64, i32, i32, u8p, i64, i64, vp]
    L.mr_categorize_clean_dev_peer.argtypes = [vp, u8p, i64, i64, i64, i32, i32, u8p, i32, u8p, i32, u8p, i64, vp]
    L.mr_ipc_export.argtypes = [vp, vp, C.POINTER(i64)]
    L.mr_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.mr_ipc_close.argtypes = [vp, vp]
    L.mr_clean_host.argtypes = [vp, u8p, i64, i64, i64, i32, i32, u8p, i64]
    L.mr_clean_dev.argtypes = [vp, u8p, i64, i64, i64, i32, i32, i32, u8p, i64, vp]
    L.mr_clean_dev_bounded.argtypes = [vp, u8p, i64, i64, i64, i32, i32, i32, i32, u8p, i64, vp]
    L.mr_clean_segments_dev.argtypes = [vp, u8p, i64, i64, i64, u8p, C.c_double, C.c_double, u8p, i64,
                                        C.POINTER(i64), vp]
    L.mr_clean_segments_host.argtypes = [vp, u8p, i64, i64, i64, u8p, C.c_double, C.c_double, u8p, i64,
                                         C.POINTER(i64)]
    L.mr_fill_gaps_dev.argtypes = [vp, u8p, i64, i64, i64, u8p, i64, i32, u8p, i64, u8p, i32, i32, u8p, i64, vp]
    L.mr_fill_gaps_host.argtypes = [vp, u8p, i64, i64, i64, u8p, i64, i32, u8p, i64, u8p, i32, i32, u8p, i64]
    L.mr_set_known_band.argtypes = [vp, i64, vp, vp, i64]
    L.mr_balance_host.argtypes = [vp, u8p, i64, i64, i64, i32, vp, vp, i64]
    L.mr_balance_dev.argtypes = [vp, u8p, i64, i64, i64, i32, vp, vp, i64, vp]
    L.mr_blur_host.argtypes = [vp, vp, i64, i64, i64, i32, i32, i32, vp, i64]
    L.mr_blur_dev.argtypes = [vp, vp, i64, i64, i64, i32, i32, i32, vp, i64, vp]
    L.mr_categorize_clean_f64_host.argtypes = [vp, vp, i64, i64, i64, i32, u8p, i64]
    L.mr_categorize_clean_f64_dev.argtypes = [vp, vp, i64, i64, i64, i32, u8p, i64, vp]
    L.mr_stds_host.argtypes = [vp, vp, i64, i64, i64, vp, i64]
    L.mr_stds_dev.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp]
    L.mr_stripes_host.argtypes = [vp, vp, i64, i64, i64, i32, vp, i64]
    L.mr_stripes_dev.argtypes = [vp, vp, i64, i64, i64, i32, vp, i64, vp]
    L.mr_set_texture_stats.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.mr_categorize_clean_props_host.argtypes = [vp, vp, i64, vp, i64, vp, i64, i64, i64, i32, i32, u8p, i64]
    L.mr_categorize_clean_props_dev.argtypes = [vp, vp, i64, vp, i64, vp, i64, i64, i64, i32, i32, u8p, i64, vp]
    L.mr_recolor_host.argtypes = [vp, u8p, i64, i64, i64, i32, vp, vp, i64]
    L.mr_recolor_dev.argtypes = [vp, u8p, i64, i64, i64, i32, vp, vp, i64, vp]
    L.mr_shrink_grow_host.argtypes = [vp, u8p, i64, i64, i64, i32, i32, u8p, i64]
    L.mr_shrink_grow_dev.argtypes = [vp, u8p, i64, i64, i64, i32, i32, u8p, i64, vp]
    L.mr_polygons_host.argtypes = [vp, i32, vp, vp, vp, i64, i64, u8p, i64]
    L.mr_polygons_dev.argtypes = [vp, i32, vp, vp, vp, i64, i64, u8p, i64, vp]
    L.mr_multi_create.argtypes = [vp, i32, C.POINTER(i32)]
    L.mr_multi_create.restype = vp
    L.mr_multi_destroy.argtypes = [vp]
    L.mr_multi_destroy.restype = None
    L.mr_multi_last_error.argtypes = [vp]
    L.mr_multi_last_error.restype = C.c_char_p
    L.mr_multi_device_count.argtypes = [vp]
    L.mr_multi_ctx.argtypes = [vp, i32]
    L.mr_multi_ctx.restype = vp
    L.mr_multi_band_range.argtypes = [vp, i64, i32, C.POINTER(i64), C.POINTER(i64)]
    L.mr_multi_set_palette.argtypes = [vp, i32, vp, vp, vp, i32, i32]
    L.mr_multi_set_known.argtypes = [vp, i64, vp, vp]
    L.mr_multi_categorize_clean_host.argtypes = [vp, u8p, i64, i64, i64, i32, i32, u8p, i64, C.POINTER(Counters)]
    L.mr_multi_batch_categorize_clean_host.argtypes = [vp, i32, u8p, i64, i64, i64, i64, i32, i32, u8p, i64, i64]
    L.mr_multi_categorize_clean_dev.argtypes = [vp, vp, i64, i64, i64, i32, i32, vp, i64]
    L.mr_host_alloc.argtypes = [C.c_size_t]
    L.mr_host_alloc.restype = vp
    L.mr_host_free.argtypes = [vp]
    L.mr_host_free.restype = None
    L.mr_synth_scan_dev.argtypes = [vp, C.c_uint32, C.c_uint32, i32, u8p, i64, i64, i64, i64, i32, u8p, i64, vp]
    for name in ("mr_set_palette", "mr_set_known", "mr_get_lut", "mr_palette_build_ms", "mr_enable_counters",
                 "mr_read_counters", "mr_categorize_clean_host", "mr_categorize_clean_host_band",
                 "mr_categorize_clean_dev",
                 "mr_categorize_clean_dev_peer", "mr_ipc_export", "mr_ipc_open", "mr_ipc_close",
                 "mr_batch_categorize_clean_dev", "mr_clean_host", "mr_clean_dev", "mr_clean_segments_dev",
                 "mr_clean_segments_host", "mr_fill_gaps_dev", "mr_fill_gaps_host", "mr_synth_scan_dev",
                 "mr_blur_host", "mr_blur_dev", "mr_categorize_clean_f64_host", "mr_categorize_clean_f64_dev",
                 "mr_balance_host", "mr_balance_dev", "mr_stds_host", "mr_stds_dev", "mr_stripes_host", "mr_stripes_dev",
                 "mr_set_texture_stats", "mr_categorize_clean_props_host", "mr_categorize_clean_props_dev",
                 "mr_polygons_host", "mr_polygons_dev", "mr_shrink_grow_host", "mr_shrink_grow_dev", "mr_recolor_host", "mr_recolor_dev",
                 "mr_set_known_band", "mr_clean_dev_bounded", "mr_multi_device_count", "mr_multi_band_range", "mr_multi_set_palette",
                 "mr_multi_set_known", "mr_multi_categorize_clean_host", "mr_multi_batch_categorize_clean_host",
                 "mr_multi_categorize_clean_dev"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def ipc_export(dev_ptr):
    """(64-byte CUDA IPC handle of the allocation holding dev_ptr, byte offset of dev_ptr in it)."""
    L = load()
    hb = C.create_string_buffer(64)
    off = C.c_int64(0)
    rc = L.mr_ipc_export(C.c_void_p(_ptr(dev_ptr)), hb, C.byref(off))
    if rc != 0:
        raise MapRastError(rc, L.mr_last_error(None).decode())
    return hb.raw, int(off.value)


def _ptr(a):
    """Raw address of a numpy array / torch tensor / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()   # torch.Tensor


class Context:
    """One GPU.  Wraps mr_ctx; raises MapRastError on any non-zero return."""

    def __init__(self, device=0):
        L = load()
        err = C.c_int(0)
        self._h = L.mr_create(int(device), C.byref(err))
        if not self._h:
            raise MapRastError(err.value, L.mr_last_error(None).decode())
        self._L = L
        self.device = int(device)
        self.palette = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.mr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MapRastError(rc, self._L.mr_last_error(self._h).decode())

    # -- palette -----------------------------------------------------------------------
    def set_palette(self, mean, lo, hi, colourspace=RGB, channels=None):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        if mean.ndim != 2 or lo.shape != mean.shape or hi.shape != mean.shape:
            raise ValueError("mean/lo/hi must be K x channels arrays of identical shape")
        K, Cn = mean.shape
        if channels is None:
            channels = Cn
        self._ck(self._L.mr_set_palette(self._h, K, _ptr(mean), _ptr(lo), _ptr(hi), int(colourspace), int(channels)))

    def set_known(self, lin_idx=None, cat=None):
        if lin_idx is None or len(lin_idx) == 0:
            self._ck(self._L.mr_set_known(self._h, 0, None, None))
            return
        idx = np.ascontiguousarray(lin_idx, dtype=np.int64)
        ct = np.ascontiguousarray(cat, dtype=np.uint8)
        self._ck(self._L.mr_set_known(self._h, idx.size, _ptr(idx), _ptr(ct)))

    def set_known_band(self, lin_idx, cat, first_line):
        """Whole-sheet indices; first_line = sheet line of line 0 of the band passed to the next calls."""
        idx = np.ascontiguousarray(lin_idx, dtype=np.int64)
        ct = np.ascontiguousarray(cat, dtype=np.uint8)
        self._ck(self._L.mr_set_known_band(self._h, idx.size, _ptr(idx), _ptr(ct), int(first_line)))

    def get_lut(self):
        lut = np.empty(LUT_SIZE, dtype=np.uint8)
        self._ck(self._L.mr_get_lut(self._h, _ptr(lut)))
        return lut

    def palette_build_ms(self):
        ms = C.c_float(0)
        self._ck(self._L.mr_palette_build_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def enable_counters(self, on=True):
        self._ck(self._L.mr_enable_counters(self._h, 1 if on else 0))

    def read_counters(self, reset=True):
        c = Counters()
        self._ck(self._L.mr_read_counters(self._h, C.byref(c), 1 if reset else 0))
        return c.as_dict()

    def launch_count(self):
        return int(self._L.mr_launch_count(self._h))

    # -- host buffers (numpy) ------------------------------------------------------------
    def categorize_clean_host(self, px, R, out=None, counters=False, halo_lo=0, halo_hi=0):
        """px: (halo_lo + n2 + halo_hi, n1, bpp) uint8 numpy (memory-identical to Julia's n1 x n2
        Matrix{RGB{N0f8}}); with halos, the first/last lines are the neighbouring bands' lines."""
        if px.dtype != np.uint8 or px.ndim != 3 or px.shape[2] not in (3, 4):
            raise ValueError("px must be a (n2, n1, 3|4) uint8 array")
        if px.strides[2] != 1 or px.strides[1] != px.shape[2]:
            px = np.ascontiguousarray(px)
        nt, n1, bpp = px.shape
        n2 = nt - halo_lo - halo_hi
        stride2 = px.strides[0] if nt > 1 else n1 * bpp
        if out is None:
            out = np.empty((n2, n1), dtype=np.uint8)
        cnt = Counters() if counters else None
        self._ck(self._L.mr_categorize_clean_host_band(
            self._h, _ptr(px) + halo_lo * stride2, n1, n2, stride2, bpp, int(R), int(halo_lo), int(halo_hi),
            _ptr(out), out.strides[0] if n2 > 1 else n1, C.byref(cnt) if counters else None))
        return (out, cnt.as_dict()) if counters else out

    def clean_host(self, labels, R, repeat=1):
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        n2, n1 = labels.shape
        out = np.empty((n2, n1), dtype=np.uint8)
        self._ck(self._L.mr_clean_host(self._h, _ptr(labels), n1, n2, n1, int(R), int(repeat), _ptr(out), n1))
        return out

    def clean_segments_host(self, labels, keep256, line_threshold, prune_threshold):
        """`_clean` (src/clean.jl:59-93) on host labels; the known plane is the one of set_known.
        Returns (labels, n_segments)."""
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        keep = np.ascontiguousarray(keep256, dtype=np.uint8)
        if keep.size != 256:
            raise ValueError("keep256 must have 256 entries")
        n2, n1 = labels.shape
        out = np.empty((n2, n1), dtype=np.uint8)
        nseg = C.c_int64(0)
        self._ck(self._L.mr_clean_segments_host(self._h, _ptr(labels), n1, n2, n1, _ptr(keep), float(line_threshold),
                                                float(prune_threshold), _ptr(out), n1, C.byref(nseg)))
        return out, int(nseg.value)

    def balance_host(self, img, coef):
        """per-pixel part of `_balance` (src/balance.jl:25-42): img (n2, n1, 3) uint8, coef (3, 7) float64
        (intercept, i^3, i^2, i, j^3, j^2, j per channel) -> (n2, n1, 3) float64."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("img must be a (n2, n1, 3) uint8 array")
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        if coef.shape != (3, 7):
            raise ValueError("coef must be 3 x 7")
        n2, n1, _ = img.shape
        out = np.empty((n2, n1, 3), dtype=np.float64)
        self._ck(self._L.mr_balance_host(self._h, _ptr(img), n1, n2, n1 * 3, 3, _ptr(coef), _ptr(out), n1 * 24))
        return out

    def blur_host(self, img, R=1, repeat=1):
        """`blur(A; hood=Window{R}(), repeat)` (src/blur.jl:6-14): (n2, n1, 3) uint8 or float64 -> float64."""
        img = np.ascontiguousarray(img)
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype not in (np.uint8, np.float64):
            raise ValueError("img must be a (n2, n1, 3) uint8 (RGB{N0f8}) or float64 (RGB{Float64}) array")
        n2, n1, _ = img.shape
        bpp = 3 if img.dtype == np.uint8 else 24
        out = np.empty((n2, n1, 3), dtype=np.float64)
        self._ck(self._L.mr_blur_host(self._h, _ptr(img), n1, n2, n1 * bpp, bpp, int(R), int(repeat), _ptr(out), n1 * 24))
        return out

    def categorize_clean_f64_host(self, px, R):
        """px: (n2, n1, 3) float64 (RGB{Float64}); labels (n2, n1) uint8."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        if px.ndim != 3 or px.shape[2] != 3:
            raise ValueError("px must be a (n2, n1, 3) float64 array")
        n2, n1, _ = px.shape
        out = np.empty((n2, n1), dtype=np.uint8)
        self._ck(self._L.mr_categorize_clean_f64_host(self._h, _ptr(px), n1, n2, n1 * 24, int(R), _ptr(out), n1))
        return out

    def stds_host(self, px):
        """`_stds(A)` (src/selectcolors.jl:706-715): (n2, n1, 3) float64 -> (n2, n1) float64."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        if px.ndim != 3 or px.shape[2] != 3:
            raise ValueError("px must be a (n2, n1, 3) float64 array")
        n2, n1, _ = px.shape
        out = np.empty((n2, n1), dtype=np.float64)
        self._ck(self._L.mr_stds_host(self._h, _ptr(px), n1, n2, n1 * 24, _ptr(out), n1 * 8))
        return out

    def stripes_host(self, px, radius=3):
        """`_stripes(img; radius)` (src/stripes.jl:20-35): (n2, n1, 3) float64 -> (n2, n1, 2) float64."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        if px.ndim != 3 or px.shape[2] != 3:
            raise ValueError("px must be a (n2, n1, 3) float64 array")
        n2, n1, _ = px.shape
        out = np.empty((n2, n1, 2), dtype=np.float64)
        self._ck(self._L.mr_stripes_host(self._h, _ptr(px), n1, n2, n1 * 24, int(radius), _ptr(out), n1 * 16))
        return out

    def set_texture_stats(self, std_mean, std_lo, std_hi, stripe_mean, stripe_lo, stripe_hi):
        """statistics of the `std` / `stripe` properties per category (src/categories.jl:10-17): (K,) and (K, 2)"""
        a = [np.ascontiguousarray(v, dtype=np.float64).reshape(-1) for v in (std_mean, std_lo, std_hi)]
        b = [np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 2) for v in (stripe_mean, stripe_lo, stripe_hi)]
        K = a[0].size
        if any(v.size != K for v in a) or any(v.shape[0] != K for v in b):
            raise ValueError("texture statistics must have one row per category")
        self._ck(self._L.mr_set_texture_stats(self._h, K, *[_ptr(v) for v in a + b]))

    def categorize_clean_props_host(self, px, std, stripe, match_bits, R):
        """PixelProp categoriser (src/categories.jl:76-133): px (n2, n1, 3), std (n2, n1), stripe (n2, n1, 2) float64
        (std / stripe may be None when not matched); match_bits: 1 color | 2 std | 4 stripe; labels (n2, n1) uint8."""
        px = np.ascontiguousarray(px, dtype=np.float64)
        if px.ndim != 3 or px.shape[2] != 3:
            raise ValueError("px must be a (n2, n1, 3) float64 array")
        n2, n1, _ = px.shape
        std = None if std is None else np.ascontiguousarray(std, dtype=np.float64)
        stripe = None if stripe is None else np.ascontiguousarray(stripe, dtype=np.float64)
        if std is not None and std.shape != (n2, n1):
            raise ValueError("std must be a (n2, n1) float64 array")
        if stripe is not None and stripe.shape != (n2, n1, 2):
            raise ValueError("stripe must be a (n2, n1, 2) float64 array")
        out = np.empty((n2, n1), dtype=np.uint8)
        self._ck(self._L.mr_categorize_clean_props_host(
            self._h, _ptr(px), n1 * 24, None if std is None else _ptr(std), n1 * 8,
            None if stripe is None else _ptr(stripe), n1 * 16, n1, n2, int(match_bits), int(R), _ptr(out), n1))
        return out

    def recolor_host(self, labels, colors):
        """`_recolor` (src/selectcolors.jl:650-655): labels (n2, n1) uint8, colors (K, 3) float64 -> (n2, n1, 4) float64"""
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        colors = np.ascontiguousarray(colors, dtype=np.float64).reshape(-1, 3)
        n2, n1 = labels.shape
        out = np.empty((n2, n1, 4), dtype=np.float64)
        self._ck(self._L.mr_recolor_host(self._h, _ptr(labels), n1, n2, n1, colors.shape[0], _ptr(colors), _ptr(out), n1 * 32))
        return out

    def shrink_grow_host(self, labels, n_erode, n_dilate):
        """`_shrink_grow(A, i, j)` (src/clean.jl:132-140): i passes of `_erode`, j passes of `_dilate`; (n2, n1) uint8."""
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        n2, n1 = labels.shape
        out = np.empty((n2, n1), dtype=np.uint8)
        self._ck(self._L.mr_shrink_grow_host(self._h, _ptr(labels), n1, n2, n1, int(n_erode), int(n_dilate), _ptr(out), n1))
        return out

    def polygons_host(self, rings, shape, fill=None, plane=None):
        """rings: list of vertex lists ((d1, d2), 1-based GUI coordinates).  fill None: the 0 / 1 mask of shape
        (n2, n1) (`_polymask`, src/selectcolors.jl:559-563); else `plane` with fill[p] painted inside ring p."""
        off = np.zeros(len(rings) + 1, dtype=np.int32)
        pts = []
        for k, r in enumerate(rings):
            r = np.asarray(r, dtype=np.float32).astype(np.float64).reshape(-1, 2)     # Point2{Float32}
            pts.append(r)
            off[k + 1] = off[k] + r.shape[0]
        xy = np.ascontiguousarray(np.concatenate(pts) if pts else np.zeros((0, 2)))
        n2, n1 = shape
        if fill is None:
            out = np.empty((n2, n1), dtype=np.uint8)
            f = None
        else:
            out = np.ascontiguousarray(plane, dtype=np.uint8).copy()
            if out.shape != (n2, n1):
                raise ValueError("plane must have the given shape")
            f = np.ascontiguousarray(fill, dtype=np.uint8)
            if f.size != len(rings):
                raise ValueError("one fill value per ring")
        self._ck(self._L.mr_polygons_host(self._h, len(rings), _ptr(off), _ptr(xy) if xy.size else None,
                                          None if f is None else _ptr(f), n1, n2, _ptr(out), n1))
        return out

    def fill_gaps_host(self, labels, px, categories, mask=None, repeat=1):
        """`fill_gaps` (src/clean.jl:2-54) `repeat` times on host buffers; palette = `stats`, the known plane
        is the one of set_known.  labels / mask: (n2, n1) uint8, px: (n2, n1, bpp) uint8."""
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        px = np.ascontiguousarray(px, dtype=np.uint8)
        n2, n1 = labels.shape
        if px.ndim != 3 or px.shape[:2] != (n2, n1) or px.shape[2] not in (3, 4):
            raise ValueError("px must be a (n2, n1, 3|4) uint8 array of the shape of labels")
        cats = np.ascontiguousarray(categories, dtype=np.uint8)
        mk = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, dtype=np.uint8)
            if mk.shape != labels.shape:
                raise ValueError("mask must have the shape of labels")
        out = np.empty((n2, n1), dtype=np.uint8)
        self._ck(self._L.mr_fill_gaps_host(self._h, _ptr(labels), n1, n2, n1, _ptr(px), n1 * px.shape[2], px.shape[2],
                                           _ptr(mk), n1, _ptr(cats), cats.size, int(repeat), _ptr(out), n1))
        return out

    # -- device buffers (raw pointers / torch tensors) -------------------------------------
    def fill_gaps_dev(self, labels, n1, n2, stride2, px, px_stride2, bpp, categories, out, out_stride2, mask=None,
                      mask_stride2=0, repeat=1, stream=0):
        cats = np.ascontiguousarray(categories, dtype=np.uint8)
        self._ck(self._L.mr_fill_gaps_dev(self._h, _ptr(labels), n1, n2, stride2, _ptr(px), px_stride2, int(bpp),
                                          _ptr(mask), mask_stride2, _ptr(cats), cats.size, int(repeat), _ptr(out),
                                          out_stride2, stream or None))

    def clean_segments_dev(self, labels, n1, n2, stride2, keep256, line_threshold, prune_threshold, out,
                           out_stride2, stream=0):
        keep = np.ascontiguousarray(keep256, dtype=np.uint8)
        if keep.size != 256:
            raise ValueError("keep256 must have 256 entries")
        nseg = C.c_int64(0)
        self._ck(self._L.mr_clean_segments_dev(self._h, _ptr(labels), n1, n2, stride2, _ptr(keep),
                                               float(line_threshold), float(prune_threshold), _ptr(out),
                                               out_stride2, C.byref(nseg), stream or None))
        return int(nseg.value)

    def categorize_clean_dev(self, px, n1, n2, stride2, bpp, R, out, out_stride2, halo_lo=0, halo_hi=0, stream=0):
        self._ck(self._L.mr_categorize_clean_dev(self._h, _ptr(px), n1, n2, stride2, bpp, int(R), int(halo_lo),
                                                 int(halo_hi), _ptr(out), out_stride2, stream or None))

    def categorize_clean_dev_peer(self, px, n1, n2, stride2, bpp, R, halo_lo_px, halo_lo, halo_hi_px, halo_hi,
                                  out, out_stride2, stream=0):
        """Band whose halo lines are read in place from other buffers (peer GPU memory)."""
        self._ck(self._L.mr_categorize_clean_dev_peer(self._h, _ptr(px), n1, n2, stride2, bpp, int(R),
                                                      _ptr(halo_lo_px), int(halo_lo), _ptr(halo_hi_px),
                                                      int(halo_hi), _ptr(out), out_stride2, stream or None))

    def ipc_open(self, handle):
        """Map another process's device allocation (64-byte CUDA IPC handle); returns its base address."""
        base = C.c_void_p(0)
        hb = C.create_string_buffer(bytes(handle), 64)
        self._ck(self._L.mr_ipc_open(self._h, hb, C.byref(base)))
        return int(base.value)

    def ipc_close(self, base):
        self._ck(self._L.mr_ipc_close(self._h, C.c_void_p(base)))

    def batch_categorize_clean_dev(self, n_sheets, px, sheet_stride, n1, n2, stride2, bpp, R, out,
                                   out_sheet_stride, out_stride2, stream=0):
        self._ck(self._L.mr_batch_categorize_clean_dev(self._h, int(n_sheets), _ptr(px), sheet_stride, n1, n2,
                                                       stride2, bpp, int(R), _ptr(out), out_sheet_stride,
                                                       out_stride2, stream or None))

    def clean_dev(self, labels, n1, n2, stride2, R, out, out_stride2, halo_lo=0, halo_hi=0, stream=0, max_label=None):
        """max_label: upper bound of the labels (e.g. K): no label-maximum pass, no synchronisation."""
        if max_label is not None:
            self._ck(self._L.mr_clean_dev_bounded(self._h, _ptr(labels), n1, n2, stride2, int(R), int(halo_lo),
                                                  int(halo_hi), int(max_label), _ptr(out), out_stride2, stream or None))
            return
        self._ck(self._L.mr_clean_dev(self._h, _ptr(labels), n1, n2, stride2, int(R), int(halo_lo), int(halo_hi),
                                      _ptr(out), out_stride2, stream or None))

    def synth_scan_dev(self, seed, sheet_id, palette_rgb, n1, n2, line0, nlines, out, out_stride2, mode=0, stream=0):
        pal = np.ascontiguousarray(palette_rgb, dtype=np.uint8)
        self._ck(self._L.mr_synth_scan_dev(self._h, seed, sheet_id, pal.shape[0], _ptr(pal), n1, n2, line0, nlines,
                                           mode, _ptr(out), out_stride2, stream or None))


class MultiContext:
    """Several GPUs of one box behind one handle (mr_multi_*): the library owns one context and one host
    thread per device; a sheet is split into row bands over n2, a batch into groups of sheets."""

    def __init__(self, devices=None):
        L = load()
        err = C.c_int(0)
        if devices is None:
            self._h = L.mr_multi_create(None, 0, C.byref(err))
        else:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._h = L.mr_multi_create(ids, len(devices), C.byref(err))
        if not self._h:
            raise MapRastError(err.value, L.mr_multi_last_error(None).decode())
        self._L = L
        self.n_dev = int(L.mr_multi_device_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.mr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MapRastError(rc, self._L.mr_multi_last_error(self._h).decode())

    def band_range(self, n2, i):
        a, n = C.c_int64(0), C.c_int64(0)
        self._ck(self._L.mr_multi_band_range(self._h, int(n2), int(i), C.byref(a), C.byref(n)))
        return int(a.value), int(n.value)

    def set_palette(self, mean, lo, hi, colourspace=RGB, channels=None):
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        K, Cn = mean.shape
        self._ck(self._L.mr_multi_set_palette(self._h, K, _ptr(mean), _ptr(lo), _ptr(hi), int(colourspace),
                                              int(channels or Cn)))

    def set_known(self, lin_idx=None, cat=None):
        if lin_idx is None or len(lin_idx) == 0:
            self._ck(self._L.mr_multi_set_known(self._h, 0, None, None))
            return
        idx = np.ascontiguousarray(lin_idx, dtype=np.int64)
        ct = np.ascontiguousarray(cat, dtype=np.uint8)
        self._ck(self._L.mr_multi_set_known(self._h, idx.size, _ptr(idx), _ptr(ct)))

    def categorize_clean_host(self, px, R, out=None, counters=False):
        """px: (n2, n1, bpp) uint8 numpy (or a raw pinned buffer wrapped as such); labels (n2, n1)."""
        if px.dtype != np.uint8 or px.ndim != 3 or px.shape[2] not in (3, 4):
            raise ValueError("px must be a (n2, n1, 3|4) uint8 array")
        if px.strides[2] != 1 or px.strides[1] != px.shape[2]:
            px = np.ascontiguousarray(px)
        n2, n1, bpp = px.shape
        stride2 = px.strides[0] if n2 > 1 else n1 * bpp
        if out is None:
            out = np.empty((n2, n1), dtype=np.uint8)
        cnt = Counters() if counters else None
        self._ck(self._L.mr_multi_categorize_clean_host(self._h, _ptr(px), n1, n2, stride2, bpp, int(R), _ptr(out),
                                                        out.strides[0] if n2 > 1 else n1,
                                                        C.byref(cnt) if counters else None))
        return (out, cnt.as_dict()) if counters else out

    def batch_categorize_clean_host(self, px, R, out=None):
        """px: (n_sheets, n2, n1, bpp) uint8, C-contiguous."""
        px = np.ascontiguousarray(px, dtype=np.uint8)
        ns, n2, n1, bpp = px.shape
        if out is None:
            out = np.empty((ns, n2, n1), dtype=np.uint8)
        self._ck(self._L.mr_multi_batch_categorize_clean_host(self._h, ns, _ptr(px), n2 * n1 * bpp, n1, n2, n1 * bpp, bpp,
                                                              int(R), _ptr(out), n2 * n1, n1))
        return out

    def categorize_clean_dev(self, band_px, n1, n2, stride2, bpp, R, band_out, out_stride2):
        """band_px / band_out: one device buffer (torch tensor or address) per band, on that band's device."""
        pin = (C.c_void_p * self.n_dev)(*[_ptr(b) if b is not None else None for b in band_px])
        pout = (C.c_void_p * self.n_dev)(*[_ptr(b) if b is not None else None for b in band_out])
        self._ck(self._L.mr_multi_categorize_clean_dev(self._h, pin, n1, n2, stride2, bpp, int(R), pout, out_stride2))
