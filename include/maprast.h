/*
 * maprast.h -- C ABI of libmaprast.so: the B200 (sm_100a) implementation of the
 * categorise + clean hot path of rafaqz/MapRasterization.jl.
 *
 * The reference has no FFI of its own (it is pure Julia); the entry points below
 * are what a Julia `ccall` binding for this path needs.  Each one names the
 * reference interface it replaces (paths under the reference checkout).
 * INTEGRATION.md shows the Julia-side stubs.
 *
 * Conventions:
 *   - image = Julia column-major Matrix, n1 = size(A,1) is the memory-fast axis;
 *     a "line" is n1 consecutive pixels; stride2 = bytes between lines.
 *   - pixels are packed RGB{N0f8} (bpp = 3) or RGBA{N0f8} (bpp = 4).
 *   - labels are uint8: 0 = no match, 1..K = category (K <= 254; 255 is reserved and rejected).
 *   - every function returns 0 (MR_OK) or a negative error code, never throws
 *     or aborts; mr_last_error() gives the message.  There is NO CPU fallback:
 *     without a usable CUDA device mr_create fails.
 *   - one ctx = one GPU; calls on one ctx must be serialised by the caller, and asynchronous
 *     (device-pointer) calls on one ctx must all use the same stream (the ctx owns scratch
 *     buffers that those launches share); different ctxs are independent, also on one device
 *     and from different host threads (no process-global state).
 */
#ifndef MAPRAST_H
#define MAPRAST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MR_OK          0
#define MR_ERR_ARG    (-1)   /* bad argument (Julia side: ArgumentError) */
#define MR_ERR_CUDA   (-2)   /* CUDA runtime failure */
#define MR_ERR_STATE  (-3)   /* call order (e.g. no palette set) */
#define MR_ERR_NOMEM  (-4)
#define MR_ERR_UNSUPPORTED (-5) /* input the reference handles in a scan-order dependent way */

#define MR_COLOURSPACE_RGB 0 /* match on RGB{Float64}: src/balance.jl:13-15 + src/blur.jl:8 */
#define MR_COLOURSPACE_LAB 1 /* extension */

typedef struct mr_ctx mr_ctx;

/* Diagnostics the north_star asks to be "counted and reported". */
typedef struct mr_counters {
    uint64_t n_pixels;   /* pixels labelled */
    uint64_t n_nomatch;  /* output labels equal to 0 */
    uint64_t n_fine;     /* pixels whose colour fell in a mixed coarse cell and was resolved
                            through the exact per-colour table (the "near-tie" path) */
    uint64_t n_changed;  /* pixels whose label the clean filter changed */
} mr_counters;

int mr_version(void);

/* Library/context.  Replaces nothing in the reference (it has no device state). */
mr_ctx* mr_create(int device, int* err);
void mr_destroy(mr_ctx* ctx);
const char* mr_last_error(const mr_ctx* ctx); /* ctx may be NULL: last error of this thread */

/* Palette = per-category statistics.  Replaces the `category_stats` value built at
 * src/categories.jl:52 by _component_stats (:2-30) and consumed at :76-82.
 * mean/lo/hi are K x channels Float64, row-major (category-major);
 * lo = min - sd*tol, hi = max + sd*tol are precomputed by the host exactly as
 * src/categories.jl:131-133 evaluates them.  A missing category (:6, :78) is
 * passed as mean = +Inf.  NaN bounds are allowed (never accept).  channels = 3
 * (RGB) or 4 (RGBA stats, alpha joins the distance, :92).  Builds the exact
 * 2^24-entry label table on the GPU (channels == 3). */
int mr_set_palette(mr_ctx* ctx, int K, const double* mean, const double* lo, const double* hi,
                   int colourspace, int channels);

/* Known-category plane, replaces _set_categories! (src/selectcolors.jl:670-681) feeding
 * the override at src/categories.jl:110-112.  lin_idx = i + j*n1 (0-based). n = 0 clears. */
int mr_set_known(mr_ctx* ctx, int64_t n, const int64_t* lin_idx, const uint8_t* cat);

/* Same for one row band of a larger sheet: lin_idx still count from SHEET line 0 and first_line is the
 * sheet line of line 0 of the px / px_dev passed to the band entry points below.  Points on the band's
 * halo lines matter too (their labels enter the windows of the band's edge lines); points outside
 * [first_line - halo_lo, first_line + n2 + halo_hi) are ignored.  mr_set_known resets first_line to 0. */
int mr_set_known_band(mr_ctx* ctx, int64_t n, const int64_t* lin_idx, const uint8_t* cat,
                      int64_t first_line);

/* Introspection used by the parity tests (exhaustive 2^24 sweep) and reports. */
int mr_get_lut(mr_ctx* ctx, uint8_t* lut_host /* 2^24 bytes, index r | g<<8 | b<<16 */);
int mr_palette_build_ms(mr_ctx* ctx, float* ms);
int mr_enable_counters(mr_ctx* ctx, int on);
int mr_read_counters(mr_ctx* ctx, mr_counters* out, int reset); /* synchronises */
int64_t mr_launch_count(const mr_ctx* ctx); /* kernels launched by this ctx so far */

/* The fused hot path: labels = clean(categorize(px)).  Replaces
 * _categorize(pixels, points; ...) non-segmented branch, src/categories.jl:49-52,68-73
 * (call site src/selectcolors.jl:466-478), followed by the Window{R} majority filter
 * (window kwarg shape of blur, src/blur.jl:6).  R = 0: categorise only.
 * Host variant: caller-owned host buffers, copies inside. */
int mr_categorize_clean_host(mr_ctx* ctx, const uint8_t* px, int64_t n1, int64_t n2,
                             int64_t stride2_bytes, int bytes_per_px, int R,
                             uint8_t* labels_out, int64_t out_stride2, mr_counters* counters);

/* Host variant for one row band of a larger host-resident sheet (single-process multi-GPU or
 * one rank of a multi-process job): px points at line 0 of the band and the halo_lo / halo_hi
 * (0..R) lines before / after it are read straight from the caller's sheet -- no exchange. */
int mr_categorize_clean_host_band(mr_ctx* ctx, const uint8_t* px, int64_t n1, int64_t n2,
                                  int64_t stride2_bytes, int bytes_per_px, int R,
                                  int halo_lo, int halo_hi,
                                  uint8_t* labels_out, int64_t out_stride2, mr_counters* counters);

/* Device variant (CuArray callers, band-sharded multi-GPU, benchmarking resident data).
 * px_dev points at line 0 of this band.  halo_lo / halo_hi (0..R) say how many REAL image
 * lines are readable before line 0 / after line n2-1 (the neighbour band's lines after a
 * halo exchange); where they are 0 the window is clipped as at the image border.
 * stream is a cudaStream_t (NULL = default stream).  Asynchronous.  Lines that are not 16-byte
 * aligned are repacked into aligned pitches (scratch owned by ctx) so that they still take the
 * fast kernel. */
int mr_categorize_clean_dev(mr_ctx* ctx, const uint8_t* px_dev, int64_t n1, int64_t n2,
                            int64_t stride2_bytes, int bytes_per_px, int R,
                            int halo_lo, int halo_hi,
                            uint8_t* labels_dev, int64_t out_stride2, void* stream);

/* Same, but the halo lines are read IN PLACE from another buffer -- the neighbour band on a peer
 * GPU of the same box (NVLink / NVSwitch peer memory; mr_ipc_* below maps it into this process).
 * The kernel's TMA line copies fetch the 2R halo lines straight from the neighbour's HBM, so a
 * band-sharded sheet needs no halo exchange step at all.  halo_lo_px = first of the halo_lo lines
 * that precede line 0 (stride2_bytes apart), halo_hi_px = first of the halo_hi lines that follow
 * line n2-1.  The caller guarantees that the neighbour's lines are complete before the launch
 * (e.g. a barrier after upload).  Fast path only: packed RGB, 16-byte aligned pointers and
 * strides, n1 % 16 == 0, R <= 3, K <= 126; no known-category plane.  Replaces nothing in the
 * reference (which is single-process); SURVEY 8(e). */
int mr_categorize_clean_dev_peer(mr_ctx* ctx, const uint8_t* px_dev, int64_t n1, int64_t n2,
                                 int64_t stride2_bytes, int bytes_per_px, int R,
                                 const uint8_t* halo_lo_px, int halo_lo,
                                 const uint8_t* halo_hi_px, int halo_hi,
                                 uint8_t* labels_dev, int64_t out_stride2, void* stream);

/* CUDA IPC plumbing for one-process-per-GPU jobs: export a device allocation (64-byte handle of
 * the allocation that contains dev_ptr + byte offset of dev_ptr in it), map a neighbour's
 * allocation into this process with peer access enabled, unmap it. */
int mr_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset);
int mr_ipc_open(mr_ctx* ctx, const void* handle64, void** base_out);
int mr_ipc_close(mr_ctx* ctx, void* base);

/* Batch of independent sheets of identical shape (BASELINE config 5): one launch. */
int mr_batch_categorize_clean_dev(mr_ctx* ctx, int n_sheets, const uint8_t* px_dev,
                                  int64_t sheet_stride_bytes, int64_t n1, int64_t n2,
                                  int64_t stride2_bytes, int bytes_per_px, int R,
                                  uint8_t* labels_dev, int64_t out_sheet_stride,
                                  int64_t out_stride2, void* stream);

/* clean(labels; hood = Window{R}(), repeat): labels in -> labels out (Appendix A.3).
 * Kwarg shape follows blur(A; hood, repeat), src/blur.jl:6-14.  Not in place.  16-byte aligned
 * lines with n1 % 16 == 0, R <= 3 and labels <= 126 take the bit-sliced fast kernel (the device
 * variant then synchronises `stream` once: the largest label picks the number of bit planes);
 * anything else takes the generic kernel. */
int mr_clean_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                  int R, int repeat, uint8_t* out, int64_t out_stride2);
int mr_clean_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                 int R, int halo_lo, int halo_hi, uint8_t* out_dev, int64_t out_stride2,
                 void* stream);

/* Same as mr_clean_dev with an upper bound of the label values supplied by the caller (typically the K of
 * the palette that produced them; a label above the bound is undefined behaviour of the vote, not of
 * memory): skips the label-maximum pass and its stream synchronisation -- fully asynchronous. */
int mr_clean_dev_bounded(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                         int R, int halo_lo, int halo_hi, int max_label, uint8_t* out_dev,
                         int64_t out_stride2, void* stream);

/* Segment prune = the reference's "clean" button: _clean(A, known_categories; categories,
 * line_threshold, prune_threshold), src/clean.jl:59-93 (call site src/selectcolors.jl:479-489),
 * over the segmentation fast_scanning!(..., 1e-9; match = (:color,)), src/scan.jl:56-157:
 * 4-connected segments of equal labels; a segment whose label is not kept becomes 0; a segment
 * with fewer than prune_threshold pixels, or count / (max(h, w)^2 / 2) < line_threshold, becomes
 * 0 -- or its known category, if one was clicked inside it (clean.jl:79-85).
 * keep256[c] != 0 <=> category c is in `categories`.  The known-category plane is the one set by
 * mr_set_known.  If one segment with a KEPT label holds two DIFFERENT known categories the
 * reference result depends on its scan order (scan.jl:140-146): the call then fails with
 * MR_ERR_UNSUPPORTED (segments whose label is not kept become 0 however they are split).
 * n1 * n2 < 2^31.  Labels are uint8 (the reference returns the Float64 segment colour; the Julia
 * wrapper converts).  The device variant synchronises `stream` (the number of segments sizes a
 * scratch buffer). */
int mr_clean_segments_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2,
                          int64_t stride2, const uint8_t* keep256, double line_threshold,
                          double prune_threshold, uint8_t* out_dev, int64_t out_stride2,
                          int64_t* n_segments, void* stream);
int mr_clean_segments_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2,
                           int64_t stride2, const uint8_t* keep256, double line_threshold,
                           double prune_threshold, uint8_t* out, int64_t out_stride2,
                           int64_t* n_segments);

/* blur(A; hood = Window{R}(), repeat), src/blur.jl:6-14 (the "blur" button, src/selectcolors.jl:455-465):
 * every pass replaces a pixel by mean(neighbors(h)) over its Window{R} (centre included) in
 * RGB{Float64}.  Input: packed RGB{N0f8} (bytes_per_px = 3; Float64(::N0f8) = i/255) or RGB{Float64}
 * (bytes_per_px = 24, three doubles per pixel); output: RGB{Float64}, 24 bytes per pixel, same geometry.
 * repeat = 0 is the conversion A .* 1.0.  Sum in neighbourhood order (memory-fast axis fastest), times
 * 1 / (2R+1)^2; neighbours outside the image are the zero colour (decisions about
 * Neighborhoods.jl / ColorVectorSpace, parity unpinned).  Strides in bytes, multiples of 8 for Float64. */
int mr_blur_host(mr_ctx* ctx, const void* px, int64_t n1, int64_t n2, int64_t stride2_bytes,
                 int bytes_per_px, int R, int repeat, double* out, int64_t out_stride2_bytes);
int mr_blur_dev(mr_ctx* ctx, const void* px_dev, int64_t n1, int64_t n2, int64_t stride2_bytes,
                int bytes_per_px, int R, int repeat, double* out_dev, int64_t out_stride2_bytes,
                void* stream);

/* _balance(A, points; category_bitindex), src/balance.jl:1-42 (the "balance" button,
 * src/selectcolors.jl:448-454), per-pixel part: predictions of the fitted cubic trend (`_predict`,
 * :36-42), their means (:25) and `_final_balance` (:28-35): out = clamp(x - p + mean(p), 0, 1) per channel
 * in RGB{Float64}.  The least-squares fit (lm(@formula(r ~ i^3 + i^2 + i + j^3 + j^2 + j), ...), :19-23)
 * stays host code of the caller; coef21 = HOST array, 3 channels x (intercept, i^3, i^2, i, j^3, j^2, j);
 * (i, j) are the 1-based pixel indices, i the memory-fast axis.  With fewer than 8 points the reference
 * returns A .* 1.0 (:16-18): that is mr_blur_* with repeat = 0.  Evaluation order and the order of the sum
 * behind mean(p) are decisions (parity unpinned).  Packed RGB{N0f8} in, RGB{Float64}
 * out (24 bytes per pixel). */
int mr_balance_host(mr_ctx* ctx, const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2_bytes,
                    int bytes_per_px, const double* coef21, double* out, int64_t out_stride2_bytes);
int mr_balance_dev(mr_ctx* ctx, const uint8_t* px_dev, int64_t n1, int64_t n2, int64_t stride2_bytes,
                   int bytes_per_px, const double* coef21, double* out_dev, int64_t out_stride2_bytes,
                   void* stream);

/* The categoriser on RGB{Float64} pixels -- what `blur` / `_balance` hand to
 * _categorize(pixels, points; ...) (src/categories.jl:49-52,68-73; the reference's own known-answer
 * vectors are Float64 colours) -- followed by the Window{R} majority filter.  Direct Float64
 * evaluation per pixel (no table).  3-channel RGB statistics only; the known plane of mr_set_known
 * applies.  px: three doubles per pixel, stride2 in bytes (multiple of 8). */
int mr_categorize_clean_f64_host(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2,
                                 int64_t stride2_bytes, int R, uint8_t* labels_out,
                                 int64_t out_stride2);
int mr_categorize_clean_f64_dev(mr_ctx* ctx, const double* px_dev, int64_t n1, int64_t n2,
                                int64_t stride2_bytes, int R, uint8_t* labels_dev,
                                int64_t out_stride2, void* stream);

/* Texture features behind the reference's default match = (:color, :std, :stripe), computed
 * from the blurred RGB{Float64} image by the "blur" button (src/selectcolors.jl:455-465):
 *   mr_stds_*     _stds(A), src/selectcolors.jl:706-715: per pixel the sum over r, g, b of the sample standard
 *                 deviation of the channel over Window(2) (5 x 5, centre included), the image divided by its maximum;
 *                 one double per pixel out;
 *   mr_stripes_*  _stripes(img; radius), src/stripes.jl:20-35: per pixel (vert - horz, angle45 - angle135) of the
 *                 directional HSL contrasts (mean over `radius` neighbours per side of (ds^2 + dl^2), the smaller
 *                 side), the image divided by mean((mean(abs o first), mean(abs o last))); two doubles per pixel out.
 * px: three doubles per pixel; strides in bytes, multiples of 8.  Neighbours outside the image are the zero
 * colour; summation orders, the two-pass std and the RGB -> HSL formula are decisions (DESIGN.md section 14:
 * Neighborhoods.jl, Statistics and Colors.jl are not in the reference tree; parity unpinned). */
int mr_stds_host(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2_bytes,
                 double* out, int64_t out_stride2_bytes);
int mr_stds_dev(mr_ctx* ctx, const double* px_dev, int64_t n1, int64_t n2, int64_t stride2_bytes,
                double* out_dev, int64_t out_stride2_bytes, void* stream);
int mr_stripes_host(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2_bytes,
                    int radius, double* out, int64_t out_stride2_bytes);
int mr_stripes_dev(mr_ctx* ctx, const double* px_dev, int64_t n1, int64_t n2, int64_t stride2_bytes,
                   int radius, double* out_dev, int64_t out_stride2_bytes, void* stream);

/* Statistics of the texture properties per category: `std` and `stripe` of _component_stats(::Vector{<:PixelProp}),
 * src/categories.jl:10-17 -- mean, and the bounds min - sd * tol / max + sd * tol of src/categories.jl:131-133,
 * computed by the host like the colour bounds of mr_set_palette.  std_*: K doubles; stripe_*: K x 2 doubles.
 * Call after mr_set_palette (a new palette drops them). */
int mr_set_texture_stats(mr_ctx* ctx, int K, const double* std_mean, const double* std_lo, const double* std_hi,
                         const double* stripe_mean, const double* stripe_lo, const double* stripe_hi);

/* _categorize(pixels::Array{PixelProp}, points; match) (src/categories.jl:45-52,68-82) followed by the Window{R}
 * majority filter: per pixel the error of a category is the mean over the MATCHED properties, in the order
 * color, std, stripe, of the colour error (:91-95), (std - mean)^2 (:101) and mean((stripe - mean)^2) (:96-100);
 * first minimum; accepted when every matched property lies in the winner's box (:106-133) or the pixel is a
 * clicked point of that category (mr_set_known).  match: bit 0 color, bit 1 std, bit 2 stripe (1 = what
 * mr_categorize_clean_f64_* does).  px: RGB{Float64}; std plane: one double, stripe plane: two doubles per
 * pixel (NULL when not matched); strides in bytes, multiples of 8. */
int mr_categorize_clean_props_host(mr_ctx* ctx, const double* px, int64_t stride2_bytes,
                                   const double* std_plane, int64_t std_stride2_bytes,
                                   const double* stripe_plane, int64_t stripe_stride2_bytes,
                                   int64_t n1, int64_t n2, int match, int R, uint8_t* labels_out,
                                   int64_t out_stride2);
int mr_categorize_clean_props_dev(mr_ctx* ctx, const double* px_dev, int64_t stride2_bytes,
                                  const double* std_dev, int64_t std_stride2_bytes,
                                  const double* stripe_dev, int64_t stripe_stride2_bytes,
                                  int64_t n1, int64_t n2, int match, int R, uint8_t* labels_dev,
                                  int64_t out_stride2, void* stream);

/* _recolor(categories, source, points) (src/selectcolors.jl:650-655), the display layers of the GUI (:378-388):
 * label c >= 1 -> RGBA(colors[c]), colors[c] = _mean_color(source, points[c]) (:701-704: the mean r, g, b of the
 * category's clicked points, computed by the host) with alpha 1; label 0 -> RGBA(RGB(0), 0).  colors: HOST array,
 * K x 3 doubles; out: RGBA{Float64}, 4 doubles per pixel, stride in bytes (multiple of 8).  A label above K (a
 * BoundsError in the reference) gives the transparent pixel. */
int mr_recolor_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int K,
                    const double* colors, double* out, int64_t out_stride2_bytes);
int mr_recolor_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2, int K,
                   const double* colors, double* out_dev, int64_t out_stride2_bytes, void* stream);

/* _shrink_grow(A, i, j) (src/clean.jl:132-140): n_erode passes of the categorical erosion `_erode` (:121-129: a pixel
 * whose 8 Moore neighbours all equal it becomes 0) followed by n_dilate passes of the categorical dilation `_dilate`
 * (:96-118: a 0 pixel takes the neighbour that occurs most often among its 8 neighbours -- first maximum in
 * neighbourhood order --, or, if that is 0, the first neighbour that is not 0).  Neighbours outside the image are 0;
 * Moore(1) in CartesianIndices order (decisions; parity unpinned).  Not in place. */
int mr_shrink_grow_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                        int n_erode, int n_dilate, uint8_t* out, int64_t out_stride2);
int mr_shrink_grow_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                       int n_erode, int n_dilate, uint8_t* out_dev, int64_t out_stride2, void* stream);

/* Polygon masks and polygon painting of the "clean" / "fill" buttons (src/selectcolors.jl:480-481, 491-497):
 *   fill == NULL   plane = _polymask(polygons, A) (:559-563: Rasters.boolmask(...; shape = :polygon)): 1 inside any
 *                  ring, 0 elsewhere;
 *   fill != NULL   plane[inside ring p] = fill[p], ring after ring (later rings win), other pixels unchanged:
 *                  rasterize!(fill_raster, p; fill = i, shape = :polygon) with fill = the category of the ring;
 *                  fill = 0 for every ring is `(1 .- mask) .* A`.
 * Rings: HOST arrays; ring p = vertices offsets[p] .. offsets[p + 1] of xy (two doubles each: the 1-based (d1, d2)
 * coordinates of the GUI, Point2{Float32} widened); also for the device variant (packed and uploaded per call).
 * Pixel (i, j) is the point (i, j); even-odd rule over the closed ring, an edge (a, b) is crossed when
 * (a2 > j) != (b2 > j) and i < a1 + (b1 - a1) * ((j - a2) / (b2 - a2)); rings with fewer than three vertices are
 * empty (the GUI's placeholder polygon).  Rasters.jl is not in the reference tree: the rule is a decision
 * (parity unpinned).  In place on `plane`. */
int mr_polygons_host(mr_ctx* ctx, int n_rings, const int32_t* offsets, const double* xy, const uint8_t* fill,
                     int64_t n1, int64_t n2, uint8_t* plane, int64_t stride2);
int mr_polygons_dev(mr_ctx* ctx, int n_rings, const int32_t* offsets, const double* xy, const uint8_t* fill,
                    int64_t n1, int64_t n2, uint8_t* plane_dev, int64_t stride2, void* stream);

/* fill_gaps(src; categories, neighborhood = VonNeumann{2}(), missingval = 0, stats, match = (:color,),
 * known, original, mask), src/clean.jl:2-54, applied `repeat` times as _fill does
 * (src/selectcolors.jl:683-699; call site :490-501).  Per pixel and pass: a masked pixel stays
 * (clean.jl:13); a clicked point whose category is kept takes it (:14, the plane of mr_set_known);
 * a kept label stays, a label that is not kept becomes 0 (:16-17, :35-36); a gap (0) whose 12
 * VonNeumann{2} neighbours are not "mostly missing" (:20) takes the kept category with the largest
 * inverse-distance weight if that exceeds 1 (:21-23), and when the runner-up also exceeds 1 the one
 * of the two whose mean colour is closer to the ORIGINAL pixel (:24-31, the colour error of
 * categories.jl:91-95 against the palette of mr_set_palette; stats[i] = palette row categories[i]-1).
 * categories: HOST array of the kept categories, 1..K, in the caller's order (equal weights go to the
 * earlier one, findmax); also for the device variant (it travels in the kernel's parameter block).  px: the pixels the categoriser saw (packed RGB / RGBA, same geometry as labels).
 * mask: optional plane, nonzero = masked (NULL: none).  Neighbours outside the image read as 0;
 * output has the size of the input; not in place.  Asynchronous on `stream` (device variant). */
int mr_fill_gaps_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                     const uint8_t* px_dev, int64_t px_stride2, int bytes_per_px,
                     const uint8_t* mask_dev, int64_t mask_stride2,
                     const uint8_t* categories, int n_categories, int repeat,
                     uint8_t* out_dev, int64_t out_stride2, void* stream);
int mr_fill_gaps_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                      const uint8_t* px, int64_t px_stride2, int bytes_per_px,
                      const uint8_t* mask, int64_t mask_stride2,
                      const uint8_t* categories, int n_categories, int repeat,
                      uint8_t* out, int64_t out_stride2);

/* ---- several GPUs of one box behind one object (mr_create(device_ids, n_dev)) ----------
 * The reference call site (src/selectcolors.jl:466-478) stays a single call: the library owns one
 * context and one host thread per device, splits the sheet into contiguous row bands over n2
 * (mr_multi_band_range; multiples of 16 lines) and runs the fused path on every band at once.
 * device_ids = NULL: devices 0..n_dev-1; n_dev <= 0: every visible device.  A device may be listed more
 * than once (one context per entry).  Results are byte-identical to the single-GPU call.
 * mr_multi_set_known takes WHOLE-SHEET indices; every band receives the points its windows can see. */
typedef struct mr_multi mr_multi;
mr_multi* mr_multi_create(const int* device_ids, int n_dev, int* err);
void mr_multi_destroy(mr_multi* m);
const char* mr_multi_last_error(const mr_multi* m); /* m may be NULL: last error of this thread */
int mr_multi_device_count(const mr_multi* m);
mr_ctx* mr_multi_ctx(mr_multi* m, int i);            /* the context of band i (owned by m) */
int mr_multi_band_range(const mr_multi* m, int64_t n2, int i, int64_t* line0, int64_t* nlines);
int mr_multi_set_palette(mr_multi* m, int K, const double* mean, const double* lo, const double* hi,
                         int colourspace, int channels);
int mr_multi_set_known(mr_multi* m, int64_t n, const int64_t* lin_idx, const uint8_t* cat);
/* Host-resident sheet: every band reads its halo lines from the caller's sheet; no exchange between
 * GPUs.  counters (may be NULL) = sums over the bands. */
int mr_multi_categorize_clean_host(mr_multi* m, const uint8_t* px, int64_t n1, int64_t n2,
                                   int64_t stride2_bytes, int bytes_per_px, int R,
                                   uint8_t* labels_out, int64_t out_stride2, mr_counters* counters);
/* Batch of independent host-resident sheets (BASELINE config 5): contiguous groups of sheets per device. */
int mr_multi_batch_categorize_clean_host(mr_multi* m, int n_sheets, const uint8_t* px,
                                         int64_t sheet_stride_bytes, int64_t n1, int64_t n2,
                                         int64_t stride2_bytes, int bytes_per_px, int R,
                                         uint8_t* labels_out, int64_t out_sheet_stride,
                                         int64_t out_stride2);
/* Device-resident bands: band_px[i] / band_out[i] point at line 0 of band i in the memory of device i
 * (lines of mr_multi_band_range(i), stride2_bytes apart; no halo slots).  The fused kernel reads the
 * 2R halo lines of a band in place from the neighbour bands over NVLink peer access, which
 * mr_multi_create enabled in-process (MR_ERR_UNSUPPORTED when the devices cannot reach each other).
 * Fast-path geometry only (packed RGB, 16-byte aligned pointers / strides, n1 % 16 == 0, R <= 3,
 * K <= 126).  Runs on the default stream of every device and returns when all bands are done. */
int mr_multi_categorize_clean_dev(mr_multi* m, const uint8_t* const* band_px, int64_t n1, int64_t n2,
                                  int64_t stride2_bytes, int bytes_per_px, int R,
                                  uint8_t* const* band_out, int64_t out_stride2);

/* Pinned host staging (optional; pageable caller buffers also work, slower). */
void* mr_host_alloc(size_t bytes);
void mr_host_free(void* p);

/* Bench/test utility (not part of the reference surface): deterministic synthetic scan,
 * SURVEY 8(d).  Writes lines [line0, line0+nlines) of sheet `sheet_id` (3 B/px). */
int mr_synth_scan_dev(mr_ctx* ctx, uint32_t seed, uint32_t sheet_id, int K,
                      const uint8_t* palette_rgb_host, int64_t n1, int64_t n2, int64_t line0,
                      int64_t nlines, int mode, uint8_t* out_dev, int64_t out_stride2,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif
