// mr_internal.h -- shared declarations between the kernel and API translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MR_LUT_SIZE (1u << 24)      // exact per-colour label table, index r | g<<8 | b<<16
// Coarse table: 32^3 cells of 8^3 colours, 0xFF = mixed cell.  Cell (r5, g5, b5) sits at
// r5 + 32*g5 + 1032*b5: the b stride is 1024 + 8 so that neighbouring b5 values fall into different
// shared-memory banks (with a stride of 1024 the colours of one map region -- b5 +- 1 -- would all
// hit the same bank with different words).  Size rounded up to a multiple of 128.
#define MR_COARSE_BSTRIDE 1032u
#define MR_COARSE_SIZE 33152u
#define MR_COARSE_IDX(r5, g5, b5) ((r5) + 32u * (g5) + MR_COARSE_BSTRIDE * (b5))
#define MR_MIXED 0xFFu

// Device-resident palette (global memory; read uniformly -> broadcast loads).
struct MrPalette {
    const double* mean;   // K x C
    const double* lo;     // K x C
    const double* hi;     // K x C
    const double* lin;    // 256-entry sRGB inverse companding table (Lab only)
    int K;
    int C;                // 3 | 4
    int colourspace;      // 0 RGB | 1 Lab
};

struct MrCountersDev {
    unsigned long long n_pixels, n_nomatch, n_fine, n_changed;
};

// One fused launch = n_sheets independent sheets (a band is one sheet with halos).
struct MrJob {
    const uint8_t* px;        // line 0 of sheet 0
    uint8_t* out;             // line 0 of sheet 0
    int64_t n1, n2;
    int64_t stride2;          // bytes between input lines
    int64_t out_stride2;
    int64_t sheet_stride;     // bytes between sheets (input)
    int64_t out_sheet_stride;
    int n_sheets;
    int bpp;                  // 3 | 4 (fused) ; 1 for label input (clean only)
    int R;
    int halo_lo, halo_hi;     // real lines available outside [0, n2)
    // optional: the halo lines live elsewhere (the neighbour band's buffer, e.g. peer GPU memory over
    // NVLink): line -k (k = 1..halo_lo) is halo_lo_px + (halo_lo - k) * stride2, line n2 + k is
    // halo_hi_px + k * stride2.  nullptr: the lines are in place around px (fast path only).
    const uint8_t* halo_lo_px;
    const uint8_t* halo_hi_px;
    int padded;               // in/out lines are readable/writable up to the next multiple of 16 bytes
    const uint8_t* lut;       // 2^24 exact table (nullptr when direct evaluation is used)
    const uint8_t* coarse;    // 2^15 coarse table
    MrPalette pal;
    MrCountersDev* counters;  // nullptr = do not count
};

// qtab: scratch of MR_QTAB_BYTES (squared channel differences, RGB colour space); nullptr = one-step build
#define MR_QTAB_BYTES ((size_t)3 * 254 * 256 * sizeof(double))
cudaError_t mr_launch_build_lut(const MrPalette& pal, uint8_t* lut, uint8_t* coarse, double* qtab, cudaStream_t s,
                                int* n_launches);
cudaError_t mr_launch_generic(const MrJob& job, bool labels_in, int sm_count, cudaStream_t s,
                              int* n_launches);
// Fast path (packed RGB, 16-byte aligned lines, R in 0..3).  Returns cudaErrorNotSupported
// when the job does not qualify (caller then uses the generic kernel).
// Item-queue counters of the fast kernel, owned by a context (mr_ctx): MR_WORK_SLOTS zero-initialised
// uint32 on the device, handed out round robin (see mr_launch_fast).
#define MR_WORK_SLOTS 4096u
struct MrWorkRing {
    unsigned int* slots = nullptr;   // device, MR_WORK_SLOTS entries, zeroed at creation
    unsigned int next = 0;           // launches so far (the context is caller-serialised)
};
cudaError_t mr_launch_fast(const MrJob& job, int sm_count, cudaStream_t s, int* n_launches, MrWorkRing* ring);
bool mr_fast_supported(const MrJob& job);
cudaError_t mr_launch_rgba_strip(const uint8_t* px, int64_t stride2, int64_t n1, int64_t lines, uint8_t* rgb,
                                 int64_t rgb_stride2, unsigned int* d_any_special, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_alpha_fix(const MrPalette& pal, const uint8_t* px, int64_t stride2, int64_t n1, int64_t lines,
                                uint8_t* labels, int64_t labels_stride2, const unsigned int* d_any_special,
                                cudaStream_t s, int* n_launches);
cudaError_t mr_launch_max_label(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, unsigned int* d_out,
                                int sm_count, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_known_fixup(const MrJob& job, int64_t n_known, const int64_t* idx,
                                  const uint8_t* cat, uint8_t* labels, int64_t labels_stride2, int64_t origin,
                                  int64_t j_first, int64_t n_lines, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_synth(uint32_t seed, uint32_t sheet, int K, const uint8_t* pal_dev, int64_t n1,
                            int64_t line0, int64_t nlines, int mode, uint8_t* out, int64_t out_stride2,
                            cudaStream_t s, int* n_launches);

// segment prune (`_clean`, src/clean.jl:59-93): run-based connected components + per-component rule
// (mr_ccl.cu).  Three phases with a host read-back between them (run count, component count).
size_t mr_prune_cbase_bytes(int64_t n1, int64_t n2);
size_t mr_prune_stage_bytes(int64_t n1, int64_t n2);
size_t mr_prune_run_bytes(unsigned int n_runs);
size_t mr_prune_rec_bytes(unsigned int n_segments);
cudaError_t mr_launch_prune_count(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, uint32_t* cbase,
                                  void* stage_mem, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_prune_components(int64_t n1, int64_t n2, const uint32_t* cbase, const void* stage_mem, void* run_mem,
                                       unsigned int n_runs, unsigned int* d_n_roots, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_prune_apply(int64_t n1, int64_t n2, const uint32_t* cbase, const void* stage_mem, void* run_mem,
                                  unsigned int n_runs, void* rec, unsigned int n_segments, int64_t n_known,
                                  const int64_t* known_idx, const uint8_t* known_cat, const uint8_t* keep_dev,
                                  double line_threshold, double prune_threshold, uint8_t* out, int64_t out_stride2,
                                  unsigned int* d_conflict, cudaStream_t s, int* n_launches);

// fill_gaps (src/clean.jl:2-54), one pass; the known-category override (clean.jl:14) runs as a second,
// sparse launch (mr_fill.cu)
cudaError_t mr_launch_fill_gaps(const MrPalette& pal, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                                const uint8_t* px, int64_t px_stride2, int bpp, const uint8_t* mask,
                                int64_t mask_stride2, const uint8_t* categories, int n_categories, int64_t n_known,
                                const int64_t* known_idx, const uint8_t* known_cat, uint8_t* out, int64_t out_stride2,
                                int sm_count, cudaStream_t s, int* n_launches);

// blur (src/blur.jl:6-14) (mr_blur.cu)
cudaError_t mr_launch_blur(const void* in, bool in_is_u8, int64_t in_stride2, int64_t n1, int64_t n2, int R, double* out,
                           int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches);

// per-pixel part of `_balance` (src/balance.jl:25-42): cubic trend coef21 = 3 x 7 doubles (host memory)
size_t mr_balance_scratch_bytes(int64_t n1, int64_t n2);
cudaError_t mr_launch_balance(const double* coef21, const uint8_t* px, int64_t stride2, int64_t n1, int64_t n2,
                              double* scratch, double* out, int64_t out_stride2, int sm_count, cudaStream_t s,
                              int* n_launches);

// texture features of match = (:color, :std, :stripe) and the categoriser on PixelProp pixels (mr_texture.cu)
cudaError_t mr_launch_stds(const double* px, int64_t stride2, int64_t n1, int64_t n2, double* out, int64_t out_stride2,
                           double* scratch, int sm_count, cudaStream_t s, int* n_launches);
size_t mr_stripes_scratch_bytes(int64_t n1, int64_t n2);
cudaError_t mr_launch_stripes(const double* px, int64_t stride2, int64_t n1, int64_t n2, int radius, double* out,
                              int64_t out_stride2, double* scratch, int sm_count, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_categorize_props(const MrPalette& pal, const double* tex, int match, const double* px, int64_t stride2,
                                       const double* sd, int64_t sd_stride2, const double* st, int64_t st_stride2,
                                       int64_t n1, int64_t n2, int64_t n_known, const int64_t* known_idx,
                                       const uint8_t* known_cat, uint8_t* out, int64_t out_stride2,
                                       MrCountersDev* counters, int sm_count, cudaStream_t s, int* n_launches);

// polygon masks / polygon painting (mr_polygon.cu); all pointers device memory
cudaError_t mr_launch_polygons(int n_rings, int n_edges, const int32_t* off, const double* xy, const double* ext,
                               const int32_t* ring_of, const uint8_t* fill, int64_t n1, int64_t n2, uint8_t* plane,
                               int64_t stride2, int sm_count, cudaStream_t s, int* n_launches);

// categorical erosion / dilation, one pass (mr_morph.cu)
cudaError_t mr_launch_morph(bool dilate, const uint8_t* a, int64_t stride2, int64_t n1, int64_t n2, uint8_t* out,
                            int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches);
cudaError_t mr_launch_recolor(const uint8_t* lab, int64_t stride2, int64_t n1, int64_t n2, int K, const double* colors_dev,
                              double* out, int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches);
