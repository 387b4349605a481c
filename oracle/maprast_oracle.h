/*
 * maprast_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the hot path of rafaqz/MapRasterization.jl
 * (per-pixel categorise + the north_star-defined Window{R} majority filter).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product library
 * (libmaprast.so) never links, loads or calls anything in oracle/.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - categorise (RGB, Float64):   PINNED by the reference's own four
 *     known-answer vectors, test/runtests.jl:6-20 (tests/golden/).
 *   - N0f8 -> Float64 conversion:  parity unpinned (FixedPointNumbers is not
 *     in /root/reference); adopted: correctly rounded i/255.0.
 *   - majority ("clean") filter:   parity unpinned (north_star-defined; the
 *     reference has no window-majority body), SURVEY.md Appendix A.3.
 *   - Lab variant:                 parity unpinned (extension, Appendix A.4).
 *   - segment prune, fill_gaps:    parity unpinned (the reference has no fixtures for
 *     _clean / fill_gaps; fill_gaps also depends on Neighborhoods.jl 0.1, not in the tree).
 *
 * Conventions: image is column-major n1 x n2, n1 is the memory-fast axis
 * ("line" = n1 contiguous pixels), stride2 = bytes between consecutive lines.
 * Labels: 0 = no match, 1..K = category.  K <= 255.
 */
#ifndef MAPRAST_ORACLE_H
#define MAPRAST_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { MRO_RGB = 0, MRO_LAB = 1 };
enum { MRO_TIE_LOWEST = 0, MRO_TIE_CENTRE_IF_MODAL = 1 };

/* categories.jl:131-133 -- lo = min - sd*tol, hi = max + sd*tol (two rounded ops) */
void mro_bounds(int n, const double* mn, const double* mx, const double* sd,
                const double* tol_per_entry, double* lo, double* hi);

/* categories.jl:28-30 -- (mean, min, max, sd[n-1]) of n values; n==1 -> sd = NaN */
void mro_component_stats(const double* v, int n, double out4[4]);

/* categories.jl:76-82 with Float64 colour input (C = 3 or 4 channels).
 * known != 0 selects the known-category override of categories.jl:110-112. */
int mro_categorize_f64(const double* x, int C, int K, const double* mean,
                       const double* lo, const double* hi, int known);

/* correctly rounded i/255.0 table (Appendix A.0) and the sRGB->Lab map (A.4) */
double mro_n0f8(int i);
void mro_rgb8_to_lab(int r, int g, int b, double lab[3]);
double mro_cbrt(double t);

/* one pixel given as bytes; colourspace MRO_RGB|MRO_LAB; bpp 3|4.
 * For bpp==4 and channels==4 the alpha term joins the distance (categories.jl:92),
 * alpha==0 forces 0 (categories.jl:116).  known as above. */
int mro_categorize_px(const uint8_t* px, int bpp, int channels, int colourspace, int K,
                      const double* mean, const double* lo, const double* hi, int known);

/* whole image (OpenMP over lines).  out_stride2 in bytes. */
void mro_categorize_u8(const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2,
                       int bpp, int channels, int colourspace, int K,
                       const double* mean, const double* lo, const double* hi,
                       int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                       uint8_t* out, int64_t out_stride2, int nthreads);

/* full 2^24 label table, index = r | g<<8 | b<<16 */
void mro_build_lut(int colourspace, int K, const double* mean, const double* lo,
                   const double* hi, uint8_t* lut, int nthreads);

/* Appendix A.3 majority filter, windows clipped at the image border.
 * halo_lo/halo_hi: number of real lines available before line 0 / after line
 * n2-1 of `labels` (used for band-seam tests; 0 for a whole image). */
void mro_majority(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                  int R, int tie_rule, int64_t halo_lo, int64_t halo_hi,
                  uint8_t* out, int64_t out_stride2, int nthreads);

/* A.5: clean(categorize(img)) computed as two whole-image passes. */
void mro_categorize_clean(const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2,
                          int bpp, int channels, int colourspace, int K,
                          const double* mean, const double* lo, const double* hi,
                          int R, int tie_rule, uint8_t* out, int64_t out_stride2,
                          int nthreads);

/* ---- deterministic synthetic scans; integer-only so that the
 * GPU generator in the product library produces identical bytes. ---- */
void mro_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]);
/* lines [line0, line0+nlines) of the n1 x n2 sheet `sheet_id` into out (3 B/px) */
void mro_synth_scan(uint32_t seed, uint32_t sheet_id, int K, const uint8_t* palette_rgb,
                    int64_t n1, int64_t n2, int64_t line0, int64_t nlines,
                    int mode, uint8_t* out, int64_t out_stride2, int nthreads);

/* `_clean` (src/clean.jl:59-93) over `fast_scanning!` (src/scan.jl:56-157) on a label image:
 * sequential, literal restatement (scan order and neighbour order of the reference).
 * known: known-category plane (NULL = all 0); keep256[c] != 0: category c is kept
 * (`categories`); returns the number of segments. */
int64_t mro_clean_segments(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                           const uint8_t* known, int64_t known_stride2, const uint8_t* keep256,
                           double line_threshold, double prune_threshold,
                           uint8_t* out, int64_t out_stride2);

/* fill_gaps (src/clean.jl:2-54), one pass; PARITY UNPINNED (see the .c file for the decisions
 * about Neighborhoods.jl's VonNeumann{2}).  categories: kept categories in the caller's order;
 * mean: K x C palette means; known / mask: full planes or NULL.  Not in place. */
void mro_fill_gaps(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                   const uint8_t* px, int64_t px_stride2, int bpp, int C, int colourspace,
                   const double* mean, const uint8_t* known, int64_t known_stride2,
                   const uint8_t* mask, int64_t mask_stride2,
                   const uint8_t* categories, int n_categories, uint8_t* out, int64_t out_stride2);

/* blur (src/blur.jl:6-14), one pass, RGB{Float64} out; PARITY UNPINNED (decisions in the .c file).
 * in_u8 != 0: packed N0f8 RGB in, else 3 doubles per pixel.  Strides in bytes. */
void mro_blur(const void* in, int in_u8, int64_t in_stride2, int64_t n1, int64_t n2, int R,
              double* out, int64_t out_stride2);
/* _categorize over an RGB{Float64} image (src/categories.jl:68-73), known points as in mro_categorize_u8 */
void mro_categorize_f64_image(const double* px, int64_t n1, int64_t n2, int64_t stride2, int K,
                              const double* mean, const double* lo, const double* hi,
                              int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                              uint8_t* out, int64_t out_stride2);

/* per-pixel part of `_balance` (src/balance.jl:25-42): predictions of the cubic trend coef[3][7], their
 * means, clamp(x - p + mean, 0, 1); PARITY UNPINNED (decisions in the .c file).  means_out may be NULL. */
void mro_balance(const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2, const double* coef,
                 double* out, int64_t out_stride2, double means_out[3]);

int mro_max_threads(void);

/* Texture features of match = (:color, :std, :stripe); PARITY UNPINNED (decisions in
 * the .c file).  px: RGB{Float64}, 3 doubles per pixel; strides in bytes.
 * mro_stds: src/selectcolors.jl:706-715, one double per pixel out.
 * mro_stripes: src/stripes.jl:20-35, two doubles per pixel out. */
void mro_stds(const double* px, int64_t n1, int64_t n2, int64_t stride2, double* out, int64_t out_stride2);
void mro_stripes(const double* px, int64_t n1, int64_t n2, int64_t stride2, int radius, double* out,
                 int64_t out_stride2);
/* src/categories.jl:76-133 on a PixelProp; match: bit 0 color, bit 1 std, bit 2 stripe */
int mro_categorize_props(const double* rgb, double sd, const double* stripe, int match, int K,
                         const double* mean, const double* lo, const double* hi,
                         const double* sd_mean, const double* sd_lo, const double* sd_hi,
                         const double* st_mean, const double* st_lo, const double* st_hi, int known);
void mro_categorize_props_image(const double* px, int64_t stride2, const double* sd, int64_t sd_stride2,
                                const double* stripe, int64_t st_stride2, int64_t n1, int64_t n2, int match, int K,
                                const double* mean, const double* lo, const double* hi,
                                const double* sd_mean, const double* sd_lo, const double* sd_hi,
                                const double* st_mean, const double* st_lo, const double* st_hi,
                                int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                                uint8_t* out, int64_t out_stride2);

/* polygon masks / polygon painting of the "clean" and "fill" buttons (src/selectcolors.jl:480-481, 491-497, 559-563);
 * PARITY UNPINNED (Rasters.jl's rule; decisions in the .c file).  fill == NULL: 0 / 1 mask; else paint. */
void mro_polygons(int n_rings, const int32_t* off, const double* xy, const uint8_t* fill, int64_t n1, int64_t n2,
                  uint8_t* plane, int64_t stride2);

/* categorical erosion / dilation over Moore(1), one pass (src/clean.jl:96-129); PARITY UNPINNED (decisions in the .c file) */
void mro_erode(const uint8_t* a, int64_t n1, int64_t n2, int64_t stride2, uint8_t* out, int64_t out_stride2);
void mro_dilate(const uint8_t* a, int64_t n1, int64_t n2, int64_t stride2, uint8_t* out, int64_t out_stride2);

/* _recolor (src/selectcolors.jl:650-655): labels -> RGBA{Float64} display layer (4 doubles per pixel) */
void mro_recolor(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int K, const double* colors,
                 double* out, int64_t out_stride2);

#ifdef __cplusplus
}
#endif
#endif
