"""maprast-b200: B200-native categorise + clean hot path of rafaqz/MapRasterization.jl.

The directory name contains a dot, so import it through `maprast_b200.py` at the repo root
(`import maprast_b200 as mr`), which registers this package under that name.
"""
from . import _lib, api, distributed, synth  # noqa: F401
from ._lib import Context, Counters, MapRastError, MultiContext, LIB_PATH  # noqa: F401
from .api import (DEFAULT_CATEGORY_TOLERANCE, MapSelection, Palette, Window, balance, balance_fit, blur, rasterize, categorize, categorize_clean,  # noqa: F401
                  categorize_clean_u8, categorize_u8, clean, clean_segments, clean_segments_u8, clean_u8,
                  component_stats, default_context, fill_gaps, fill_gaps_u8, stds, stripes, TextureStats, polymask, paint_polygons, shrink_grow, erode, dilate, recolor, mean_color)
