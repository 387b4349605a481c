/*
 * maprast_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See maprast_oracle.h for scope and parity status.  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fopenmp): every Float64 operation below is an
 * individually rounded IEEE-754 binary64 operation, no FMA, no reassociation,
 * exactly as Julia executes the reference.
 *
 * Each function cites the reference file:line (under /root/reference) it follows.
 */
#include "maprast_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int mro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* nthreads <= 0: the OpenMP default (OMP_NUM_THREADS); an explicit request is honoured up to the
 * number of processors, even when a launcher exported OMP_NUM_THREADS=1 (torch.distributed.run). */
static int clamp_threads(int nthreads) {
    if (nthreads <= 0) return mro_max_threads();
#ifdef _OPENMP
    if (nthreads > omp_get_num_procs()) nthreads = omp_get_num_procs();
#else
    nthreads = 1;
#endif
    return nthreads;
}

/* ---- src/categories.jl:131-133: (s.min - s.sd * tol), (s.max + s.sd * tol) ---- */
void mro_bounds(int n, const double* mn, const double* mx, const double* sd,
                const double* tol, double* lo, double* hi) {
    for (int i = 0; i < n; ++i) {
        volatile double w = sd[i] * tol[i];   /* rounded product first */
        lo[i] = mn[i] - w;
        hi[i] = mx[i] + w;
    }
}

/* ---- src/categories.jl:28-30: (mean, minimum, maximum, std) ; std = sample sd (n-1).
 * Summation is a plain left-to-right loop (Julia's pairwise/@simd order is not
 * reproducible here: SURVEY Appendix C.5 -- stats are host-side inputs anyway). ---- */
void mro_component_stats(const double* v, int n, double out4[4]) {
    double s = 0.0, mn = v[0], mx = v[0];
    for (int i = 0; i < n; ++i) {
        s += v[i];
        if (v[i] < mn) mn = v[i];
        if (v[i] > mx) mx = v[i];
    }
    double mean = s / (double)n;
    double ss = 0.0;
    for (int i = 0; i < n; ++i) {
        double d = v[i] - mean;
        ss += d * d;
    }
    out4[0] = mean;
    out4[1] = mn;
    out4[2] = mx;
    out4[3] = (n > 1) ? sqrt(ss / (double)(n - 1)) : NAN;
}

/* ---- Appendix A.0: Float64(::N0f8) adopted as the correctly rounded i/255.0 ---- */
double mro_n0f8(int i) { return (double)i / 255.0; }

/* ---- Appendix A.4 (extension, no reference body): hand-written FMA-free cbrt ----
 * Bit-hack seed on the high word, then five Newton steps y <- (2y + t/y^2)/3,
 * each step a fixed sequence of IEEE mul/div/add.  Valid for t > 0 normal. */
double mro_cbrt(double t) {
    uint64_t bits;
    memcpy(&bits, &t, 8);
    uint32_t hi = (uint32_t)(bits >> 32);
    hi = hi / 3u + 715094163u;
    bits = (uint64_t)hi << 32;
    double y;
    memcpy(&y, &bits, 8);
    for (int it = 0; it < 5; ++it) {
        double yy = y * y;
        double q = t / yy;
        double s = y + y;
        s = s + q;
        y = s / 3.0;
    }
    return y;
}

static double g_lin[256];
static int g_lin_ready = 0;
static void init_lin(void) {
    if (g_lin_ready) return;
#pragma omp critical(mro_lin)
    {
        if (!g_lin_ready) {
            for (int i = 0; i < 256; ++i) {
                double v = (double)i / 255.0;
                g_lin[i] = (v <= 0.04045) ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4);
            }
            g_lin_ready = 1;
        }
    }
}

static double lab_f(double v) {
    const double eps = 216.0 / 24389.0;
    const double kappa = 24389.0 / 27.0;
    if (v > eps) return mro_cbrt(v);
    double t = kappa * v;
    t = t + 16.0;
    return t / 116.0;
}

/* sRGB (D65) -> CIE L*a*b*, constants as remembered from Colors.jl
 * ("unverified against Colors.jl"). */
void mro_rgb8_to_lab(int r8, int g8, int b8, double lab[3]) {
    init_lin();
    double r = g_lin[r8], g = g_lin[g8], b = g_lin[b8];
    double X = (0.4124564390896921 * r + 0.357576077643909 * g) + 0.18043748326639894 * b;
    double Y = (0.21267285140562248 * r + 0.715152155287818 * g) + 0.07217499330655958 * b;
    double Z = (0.019333895582329317 * r + 0.119192025881303 * g) + 0.9503040785363677 * b;
    double fx = lab_f(X / 0.95047);
    double fy = lab_f(Y / 1.0);
    double fz = lab_f(Z / 1.08883);
    lab[0] = 116.0 * fy - 16.0;
    lab[1] = 500.0 * (fx - fy);
    lab[2] = 200.0 * (fy - fz);
}

/* ---- src/categories.jl:76-82 (argmin + accept), :91-95 (distance),
 *      :106-121,131-133 (box test / known override / alpha rule) ---- */
int mro_categorize_f64(const double* x, int C, int K, const double* mean,
                       const double* lo, const double* hi, int known) {
    int best = 0;
    double beste = 0.0;
    for (int c = 0; c < K; ++c) {
        const double* m = mean + (int64_t)c * C;
        /* categories.jl:92-94: sum(map((v - mean)^2)) over (r,g,b[,alpha]), left fold */
        double d0 = x[0] - m[0], d1 = x[1] - m[1], d2 = x[2] - m[2];
        double q0 = d0 * d0, q1 = d1 * d1, q2 = d2 * d2;
        double e = q0 + q1;
        e = e + q2;
        if (C == 4) {
            double d3 = x[3] - m[3];
            double q3 = d3 * d3;
            e = e + q3;
        }
        /* categories.jl:86: mean over the 1-tuple of matched properties = e/1 (exact).
         * categories.jl:80: findmin -> first minimum (strict <). */
        if (c == 0 || e < beste) { beste = e; best = c; }
    }
    /* categories.jl:110-112: known-category override skips the box test */
    if (known != 0) return (best + 1 == known) ? best + 1 : 0;
    /* categories.jl:116: transparent pixel never matches */
    if (C == 4 && x[3] == 0.0) return 0;
    /* categories.jl:119-121,131-133: box on r,g,b of the winner only; NaN -> false */
    const double* l = lo + (int64_t)best * C;
    const double* h = hi + (int64_t)best * C;
    for (int ch = 0; ch < 3; ++ch)
        if (!(x[ch] >= l[ch] && x[ch] <= h[ch])) return 0;
    return best + 1;
}

int mro_categorize_px(const uint8_t* px, int bpp, int channels, int colourspace, int K,
                      const double* mean, const double* lo, const double* hi, int known) {
    double x[4];
    if (colourspace == MRO_LAB) {
        mro_rgb8_to_lab(px[0], px[1], px[2], x);
        /* transparent pixel rule still applies to RGBA storage */
        if (bpp == 4 && px[3] == 0 && known == 0) return 0;
        return mro_categorize_f64(x, 3, K, mean, lo, hi, known);
    }
    x[0] = mro_n0f8(px[0]);
    x[1] = mro_n0f8(px[1]);
    x[2] = mro_n0f8(px[2]);
    if (bpp == 4) {
        x[3] = mro_n0f8(px[3]);
        if (channels == 4) return mro_categorize_f64(x, 4, K, mean, lo, hi, known);
        if (px[3] == 0 && known == 0) return 0;
    }
    return mro_categorize_f64(x, 3, K, mean, lo, hi, known);
}

/* ---- src/categories.jl:68-72: map over all pixels (non-segmented branch) ---- */
void mro_categorize_u8(const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2,
                       int bpp, int channels, int colourspace, int K,
                       const double* mean, const double* lo, const double* hi,
                       int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                       uint8_t* out, int64_t out_stride2, int nthreads) {
    nthreads = clamp_threads(nthreads);
    init_lin();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t j = 0; j < n2; ++j) {
        const uint8_t* line = px + j * stride2;
        uint8_t* o = out + j * out_stride2;
        for (int64_t i = 0; i < n1; ++i)
            o[i] = (uint8_t)mro_categorize_px(line + i * bpp, bpp, channels, colourspace, K,
                                              mean, lo, hi, 0);
    }
    /* src/selectcolors.jl:670-681 known plane: linear index = i + j*n1 (0-based) */
    for (int64_t k = 0; k < n_known; ++k) {
        int64_t i = known_idx[k] % n1, j = known_idx[k] / n1;
        if (known_cat[k] == 0) continue;
        out[j * out_stride2 + i] = (uint8_t)mro_categorize_px(
            px + j * stride2 + i * bpp, bpp, channels, colourspace, K, mean, lo, hi, known_cat[k]);
    }
}

void mro_build_lut(int colourspace, int K, const double* mean, const double* lo,
                   const double* hi, uint8_t* lut, int nthreads) {
    nthreads = clamp_threads(nthreads);
    init_lin();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int b = 0; b < 256; ++b)
        for (int g = 0; g < 256; ++g)
            for (int r = 0; r < 256; ++r) {
                uint8_t p[3] = {(uint8_t)r, (uint8_t)g, (uint8_t)b};
                lut[(size_t)r | ((size_t)g << 8) | ((size_t)b << 16)] =
                    (uint8_t)mro_categorize_px(p, 3, 3, colourspace, K, mean, lo, hi, 0);
            }
}

/* ---- SURVEY Appendix A.3 (north_star-defined; conventions from src/clean.jl:21-22
 * (findmax => first/lowest on ties), :96,121 (0 is an ordinary label here),
 * src/blur.jl:6 (Window{R} includes the centre, side 2R+1)).  Naive histogram. ---- */
void mro_majority(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                  int R, int tie_rule, int64_t halo_lo, int64_t halo_hi,
                  uint8_t* out, int64_t out_stride2, int nthreads) {
    nthreads = clamp_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t j = 0; j < n2; ++j) {
        for (int64_t i = 0; i < n1; ++i) {
            int cnt[256];
            int maxl = 0;
            memset(cnt, 0, sizeof(cnt));
            for (int64_t dj = -R; dj <= R; ++dj) {
                int64_t jj = j + dj;
                if (jj < -halo_lo || jj >= n2 + halo_hi) continue;
                const uint8_t* line = labels + jj * stride2;
                for (int64_t di = -R; di <= R; ++di) {
                    int64_t ii = i + di;
                    if (ii < 0 || ii >= n1) continue;
                    int l = line[ii];
                    cnt[l]++;
                    if (l > maxl) maxl = l;
                }
            }
            /* first maximum in label order = lowest label on ties (findmax, clean.jl:21-22) */
            int best = 0, bc = cnt[0];
            for (int l = 1; l <= maxl; ++l)
                if (cnt[l] > bc) { bc = cnt[l]; best = l; }
            if (tie_rule == MRO_TIE_CENTRE_IF_MODAL) {
                int c = labels[j * stride2 + i];
                if (cnt[c] == bc) best = c;
            }
            out[j * out_stride2 + i] = (uint8_t)best;
        }
    }
}

void mro_categorize_clean(const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2,
                          int bpp, int channels, int colourspace, int K,
                          const double* mean, const double* lo, const double* hi,
                          int R, int tie_rule, uint8_t* out, int64_t out_stride2,
                          int nthreads) {
    if (R <= 0) {
        mro_categorize_u8(px, n1, n2, stride2, bpp, channels, colourspace, K, mean, lo, hi,
                          0, 0, 0, out, out_stride2, nthreads);
        return;
    }
    uint8_t* tmp = (uint8_t*)malloc((size_t)n1 * (size_t)n2);
    mro_categorize_u8(px, n1, n2, stride2, bpp, channels, colourspace, K, mean, lo, hi,
                      0, 0, 0, tmp, n1, nthreads);
    mro_majority(tmp, n1, n2, n1, R, tie_rule, 0, 0, out, out_stride2, nthreads);
    free(tmp);
}

/* ------------------------------------------------------------------------- *
 * Synthetic scans, SURVEY.md 8(d).  Integer-only (Philox4x32-10 + Irwin-Hall
 * noise) so the CUDA generator in the product library yields identical bytes.
 * ------------------------------------------------------------------------- */
void mro_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline int synth_noise(uint32_t w) {
    int s = (int)(w & 255u) + (int)((w >> 8) & 255u) + (int)((w >> 16) & 255u) + (int)(w >> 24) - 510;
    /* sigma(sum of 4 U[0,255]) = 147.8; * 333/8192 -> sigma ~ 6 (N0f8 units), rounded */
    return ((s * 333 + 4096 + (1 << 20)) >> 13) - 128;
}

static inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

static void synth_pixel(uint32_t seed, uint32_t sheet, int K, const uint8_t* pal,
                        int64_t i, int64_t j, int mode, uint8_t* o) {
    uint32_t w[4];
    mro_philox4x32((uint32_t)i, (uint32_t)j, 0x50495845u, 0u, seed, sheet, w);
    if (mode == 1) { /* uniform random RGB: worst case for any colour-table cache */
        o[0] = (uint8_t)(w[0] & 255u); o[1] = (uint8_t)(w[1] & 255u); o[2] = (uint8_t)(w[2] & 255u);
        return;
    }
    /* Voronoi label field: one jittered site per 64x64 cell, nearest of the 3x3 cells */
    int64_t ci = i >> 6, cj = j >> 6;
    int64_t bestd = -1;
    int idx = 0;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            uint32_t s[4];
            int64_t a = ci + di, b = cj + dj;
            mro_philox4x32((uint32_t)a, (uint32_t)b, 0x53495445u, 0u, seed, sheet, s);
            int64_t sx = a * 64 + (int64_t)(s[0] & 63u), sy = b * 64 + (int64_t)(s[1] & 63u);
            int64_t d = (i - sx) * (i - sx) + (j - sy) * (j - sy);
            if (bestd < 0 || d < bestd) { bestd = d; idx = (int)(s[2] % (uint32_t)K); }
        }
    uint32_t sel = w[3] & 0xFFFFu;
    if (K > 1 && sel < 1311u) {            /* 2 % speckle: a different palette colour */
        idx = (int)(((uint32_t)idx + 1u + ((w[3] >> 16) % (uint32_t)(K - 1))) % (uint32_t)K);
    } else if (sel >= 1311u && sel < 1639u) { /* 0.5 % stray: uniform RGB */
        o[0] = (uint8_t)(w[0] & 255u); o[1] = (uint8_t)(w[1] & 255u); o[2] = (uint8_t)(w[2] & 255u);
        return;
    }
    for (int ch = 0; ch < 3; ++ch)
        o[ch] = (uint8_t)clamp255((int)pal[idx * 3 + ch] + synth_noise(w[ch]));
}

void mro_synth_scan(uint32_t seed, uint32_t sheet_id, int K, const uint8_t* palette_rgb,
                    int64_t n1, int64_t n2, int64_t line0, int64_t nlines,
                    int mode, uint8_t* out, int64_t out_stride2, int nthreads) {
    (void)n2;
    nthreads = clamp_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t jj = 0; jj < nlines; ++jj) {
        uint8_t* o = out + jj * out_stride2;
        for (int64_t i = 0; i < n1; ++i)
            synth_pixel(seed, sheet_id, K, palette_rgb, i, line0 + jj, mode, o + 3 * i);
    }
}

/* ====================================================================================
 * Segment prune: the reference's real "clean" button, `_clean` (src/clean.jl:59-93), on an
 * Int label image, through the sequential union-find segmentation `fast_scanning!`
 * (src/scan.jl:56-157) with threshold 1e-9 and match = (:color,) (src/clean.jl:62-64).
 * The "colour" of a pixel is its label; PixelProp.category is the known-category plane.
 *
 * Restated literally, in the reference's pixel order (CartesianIndices of a column-major
 * matrix: first index fastest) and neighbour order (x - e1, then x - e2; scan.jl:61-63):
 *   - a backward neighbour is valid when (mean.color - label)^2 < 1e-9 (scan.jl:81);
 *     the running mean of equal labels stays exactly the label (scan.jl:106,118,124);
 *   - no valid neighbour -> new segment (:95-101); all valid roots equal and categories
 *     mergeable -> join (:104-108); all categories mergeable -> union (:110-133); otherwise
 *     a new segment starts at this pixel (:140-146).  _canmerge: scan.jl:169.
 *   - per segment (clean.jl:65-88): size, bounding box, rectangle_area = max(h, w)^2 / 2,
 *     ratio = count / rectangle_area; label not kept -> 0; count < prune_threshold or
 *     ratio < line_threshold: no known category -> 0, label not in (0, category) -> category.
 * Returns the number of segments.  Output: the segment's (new) colour per pixel (clean.jl:91-93).
 * ==================================================================================== */
static int64_t ds_find(int64_t* parent, int64_t x) {
    while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
    }
    return x;
}

static int canmerge(int a, int b) { return (a == 0) | (b == 0) | (a == b); }
static int mergecat(int a, int b) { return a == 0 ? b : a; }   /* scan.jl:22-31 (conflicts are guarded) */

int64_t mro_clean_segments(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                           const uint8_t* known, int64_t known_stride2, const uint8_t* keep256,
                           double line_threshold, double prune_threshold,
                           uint8_t* out, int64_t out_stride2) {
    const int64_t n = n1 * n2;
    if (n == 0) return 0;
    int64_t* result = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);   /* label per pixel */
    int64_t* parent = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);   /* temp_labels (at most n) */
    double* mean = (double*)malloc(sizeof(double) * (size_t)n);        /* region_means[...].color */
    int* cat = (int*)malloc(sizeof(int) * (size_t)n);                  /* region_means[...].category */
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);      /* region_pix_count (0 = deleted) */
    int64_t* bb = (int64_t*)malloc(sizeof(int64_t) * 4 * (size_t)n);   /* region_bbox: i0, j0, i1, j1 */
    int64_t n_labels = 0;
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            const double x = (double)labels[j * stride2 + i];
            const int ctg = known ? known[j * known_stride2 + i] : 0;
            int sz = 0, same_label = 1;
            int64_t prev_label = -1, v_neigh[2];
            int ctg_neigh[2];
            for (int d = 0; d < 2; ++d) {   /* neighbourhood(point) = (point - e1, point - e2) */
                const int64_t pi = d == 0 ? i - 1 : i, pj = d == 0 ? j : j - 1;
                if (pi < 0 || pj < 0) continue;
                const int64_t root_p = ds_find(parent, result[pj * n1 + pi]);
                const double diff = mean[root_p] - x;
                if (diff * diff < 0.000000001) {
                    if (prev_label < 0) prev_label = root_p;
                    else if (prev_label != root_p) same_label = 0;
                    v_neigh[sz] = root_p;
                    ctg_neigh[sz] = cat[root_p];
                    ++sz;
                }
            }
            int64_t lab;
            int fresh = 0;
            if (sz == 0) {
                fresh = 1;
            } else if (same_label && canmerge(ctg, cat[prev_label])) {
                lab = prev_label;
                cnt[lab] += 1;
                mean[lab] += (x - mean[lab]) / (double)cnt[lab];
                cat[lab] = mergecat(cat[lab], ctg);
                if (i < bb[4 * lab + 0]) bb[4 * lab + 0] = i;
                if (j < bb[4 * lab + 1]) bb[4 * lab + 1] = j;
                if (i > bb[4 * lab + 2]) bb[4 * lab + 2] = i;
                if (j > bb[4 * lab + 3]) bb[4 * lab + 3] = j;
            } else {
                int ok = 1;
                for (int k = 0; k < sz; ++k) ok = ok && canmerge(ctg, ctg_neigh[k]);
                ok = ok && (sz < 2 || canmerge(ctg_neigh[0], ctg_neigh[1]));
                if (ok) {
                    int64_t u = v_neigh[0];
                    for (int k = 0; k < sz; ++k) {   /* union!: the smaller index becomes the root here;
                                                        which root survives does not change any statistic */
                        const int64_t a = ds_find(parent, u), b = ds_find(parent, v_neigh[k]);
                        if (a != b) {
                            if (a < b) { parent[b] = a; u = a; } else { parent[a] = b; u = b; }
                        } else u = a;
                    }
                    lab = u;
                    /* scan.jl:117-119: the pixel joins the union label first ... */
                    cnt[u] += 1;
                    mean[u] += (x - mean[u]) / (double)cnt[u];
                    cat[u] = mergecat(cat[u], ctg);
                    if (i < bb[4 * u + 0]) bb[4 * u + 0] = i;
                    if (j < bb[4 * u + 1]) bb[4 * u + 1] = j;
                    if (i > bb[4 * u + 2]) bb[4 * u + 2] = i;
                    if (j > bb[4 * u + 3]) bb[4 * u + 3] = j;
                    /* ... then the other segments are folded in (:121-132) */
                    for (int k = 0; k < sz; ++k) {
                        const int64_t v = v_neigh[k];
                        if (v != u && cnt[v] > 0) {
                            cnt[u] += cnt[v];
                            mean[u] += (mean[v] - mean[u]) * (double)cnt[v] / (double)cnt[u];
                            cat[u] = mergecat(cat[u], cat[v]);
                            if (bb[4 * v + 0] < bb[4 * u + 0]) bb[4 * u + 0] = bb[4 * v + 0];
                            if (bb[4 * v + 1] < bb[4 * u + 1]) bb[4 * u + 1] = bb[4 * v + 1];
                            if (bb[4 * v + 2] > bb[4 * u + 2]) bb[4 * u + 2] = bb[4 * v + 2];
                            if (bb[4 * v + 3] > bb[4 * u + 3]) bb[4 * u + 3] = bb[4 * v + 3];
                            cnt[v] = 0;   /* delete!(region_pix_count, v) */
                        }
                    }
                } else {
                    fresh = 1;
                }
            }
            if (fresh) {
                lab = n_labels++;
                parent[lab] = lab;
                mean[lab] = x;
                cat[lab] = ctg;
                cnt[lab] = 1;
                bb[4 * lab + 0] = bb[4 * lab + 2] = i;
                bb[4 * lab + 1] = bb[4 * lab + 3] = j;
            }
            result[j * n1 + i] = lab;
        }
    /* clean.jl:65-88 on the surviving segments */
    int64_t n_seg = 0;
    for (int64_t l = 0; l < n_labels; ++l) {
        if (cnt[l] == 0 || parent[l] != l) continue;
        ++n_seg;
        const int64_t h = bb[4 * l + 2] - bb[4 * l + 0] + 1, w = bb[4 * l + 3] - bb[4 * l + 1] + 1;
        const int64_t m = h > w ? h : w;
        const double rectangle_area = (double)(m * m) / 2.0;
        const double ratio = (double)cnt[l] / rectangle_area;
        const double colour = mean[l];
        int kept = 0;
        for (int c = 0; c < 256; ++c)
            if (keep256[c] && colour == (double)c) kept = 1;
        if (!kept) {
            mean[l] = 0.0;
        } else if (((double)cnt[l] < prune_threshold) || (ratio < line_threshold)) {
            if (cat[l] == 0) mean[l] = 0.0;
            else if (!(colour == 0.0 || colour == (double)cat[l])) mean[l] = (double)cat[l];
        }
    }
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i)
            out[j * out_stride2 + i] = (uint8_t)mean[ds_find(parent, result[j * n1 + i])];
    free(result); free(parent); free(mean); free(cat); free(cnt); free(bb);
    return n_seg;
}

/* =====================================================================================
 * fill_gaps (src/clean.jl:2-54) and its repeat loop `_fill` (src/selectcolors.jl:683-699).
 * PARITY UNPINNED: the reference has no test or fixture for fill_gaps, and the neighbourhood
 * type lives in Neighborhoods.jl 0.1, which is not in the reference tree.  Decisions made here
 * (stated in DESIGN.md section 11):
 *   - VonNeumann{2} = the 12 offsets with 1 <= |d1| + |d2| <= 2, visited in CartesianIndices
 *     order (d1, the memory-fast axis, runs fastest).  Whether the centre belongs to the
 *     neighbourhood does not matter: the stencil only runs for a centre equal to missingval,
 *     which shifts both sides of the "mostly missing" test (clean.jl:20) by one and never
 *     matches a category.  The visiting order only matters for the rounding of the Float64 sums.
 *   - distances = sqrt(d1^2 + d2^2) in Float64: weights 1/1, 1/sqrt(2), 1/2 (clean.jl:44-52).
 *   - neighbours outside the image read as missingval; the result has the size of the input.
 *   - missingval = 0 (as _fill passes it); `original` is the pixel the categoriser saw and the
 *     colour error is the categoriser's own (categories.jl:91-95, match = (:color,)).
 * categories: the kept categories in the caller's order (findmax / the reduce of clean.jl:24-27
 * return the FIRST of equal counts, i.e. the one that comes first in this list).
 * mean: K x C palette means (stats[i] of clean.jl:29-30 = row categories[i]-1).
 * known / mask may be NULL.  Not in place. */
static double fill_colour_error(const uint8_t* px, int bpp, int C, int colourspace, const double* m) {
    double x[4] = {0.0, 0.0, 0.0, 0.0};
    if (colourspace == MRO_LAB) {
        mro_rgb8_to_lab(px[0], px[1], px[2], x);
    } else {
        x[0] = mro_n0f8(px[0]); x[1] = mro_n0f8(px[1]); x[2] = mro_n0f8(px[2]);
        if (bpp == 4) x[3] = mro_n0f8(px[3]);
    }
    double d0 = x[0] - m[0], d1 = x[1] - m[1], d2 = x[2] - m[2];
    double e = d0 * d0 + d1 * d1;
    e = e + d2 * d2;
    if (C == 4) { double d3 = x[3] - m[3]; e = e + d3 * d3; }
    return e;
}

void mro_fill_gaps(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                   const uint8_t* px, int64_t px_stride2, int bpp, int C, int colourspace,
                   const double* mean, const uint8_t* known, int64_t known_stride2,
                   const uint8_t* mask, int64_t mask_stride2,
                   const uint8_t* categories, int n_categories, uint8_t* out, int64_t out_stride2) {
    static const int D1[12] = {0, -1, 0, 1, -2, -1, 1, 2, -1, 0, 1, 0};
    static const int D2[12] = {-2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2};
    double invd[12];
    for (int k = 0; k < 12; ++k) invd[k] = 1.0 / sqrt((double)(D1[k] * D1[k] + D2[k] * D2[k]));
    int kept[256];
    for (int c = 0; c < 256; ++c) kept[c] = 0;
    for (int i = 0; i < n_categories; ++i) kept[categories[i]] = 1;
    double* cnt = (double*)malloc(sizeof(double) * (size_t)(n_categories > 0 ? n_categories : 1));
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            const int v = labels[j * stride2 + i];
            uint8_t* o = out + j * out_stride2 + i;
            if (mask && mask[j * mask_stride2 + i]) { *o = (uint8_t)v; continue; }              /* clean.jl:13 */
            const int kn = known ? known[j * known_stride2 + i] : 0;
            if (kn != 0 && kept[kn]) { *o = (uint8_t)kn; continue; }                            /* clean.jl:14 */
            if (v != 0) { *o = kept[v] ? (uint8_t)v : 0; continue; }                            /* clean.jl:16-17, 35-36 */
            int hood[12], missing = 0;
            for (int k = 0; k < 12; ++k) {
                const int64_t a = i + D1[k], b = j + D2[k];
                hood[k] = (a >= 0 && a < n1 && b >= 0 && b < n2) ? labels[b * stride2 + a] : 0;
                missing += hood[k] == 0;                                                        /* clean.jl:15 */
            }
            if (missing > 12 - 12 / 4) { *o = 0; continue; }                                    /* clean.jl:20 */
            for (int c = 0; c < n_categories; ++c) {                                            /* clean.jl:43-54 */
                double acc = 0.0;
                for (int k = 0; k < 12; ++k)
                    if (hood[k] != 0 && hood[k] == categories[c]) acc += invd[k];
                cnt[c] = acc;
            }
            int im = 0;                                                                         /* findmax: first maximum */
            for (int c = 1; c < n_categories; ++c)
                if (cnt[c] > cnt[im]) im = c;
            if (n_categories == 0 || !(cnt[im] > 1.0)) { *o = 0; continue; }                    /* clean.jl:23 */
            double second = 0.0;                                                                /* clean.jl:24-27 */
            int is = -1;
            for (int c = 0; c < n_categories; ++c) {
                if (c == im) continue;
                if (cnt[c] > second) { second = cnt[c]; is = c; }
            }
            int pick = im;
            if (second > 1.0) {                                                                 /* clean.jl:28-31 */
                const uint8_t* p = px + j * px_stride2 + i * bpp;
                const double ea = fill_colour_error(p, bpp, C, colourspace, mean + (int64_t)(categories[im] - 1) * C);
                const double eb = fill_colour_error(p, bpp, C, colourspace, mean + (int64_t)(categories[is] - 1) * C);
                pick = ea <= eb ? im : is;
            }
            *o = categories[pick];
        }
    free(cnt);
}

/* =====================================================================================
 * blur(A; hood = Window{R}(), repeat) (src/blur.jl:6-14) on RGB{Float64}, and the image-level
 * categoriser on RGB{Float64} pixels (src/categories.jl:68-73 over :76-95).
 * PARITY UNPINNED for blur (Neighborhoods.jl 0.1 and ColorVectorSpace are not in the reference tree).
 * Decisions: neighbours in CartesianIndices order (d1 fastest), left-folded
 * sum, times (1 / count) -- ColorVectorSpace's RGB / Integer --, count = (2R+1)^2 always, neighbours
 * outside the image are the zero colour (adding 0.0 is exact: they are skipped), same size out.
 * in_u8 != 0: `in` holds packed N0f8 RGB (3 B per pixel), else 3 doubles per pixel.  Strides in bytes. */
void mro_blur(const void* in, int in_u8, int64_t in_stride2, int64_t n1, int64_t n2, int R,
              double* out, int64_t out_stride2) {
    const double rcnt = 1.0 / (double)((2 * R + 1) * (2 * R + 1));
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            double s[3] = {0.0, 0.0, 0.0};
            for (int d2 = -R; d2 <= R; ++d2) {
                const int64_t b = j + d2;
                if (b < 0 || b >= n2) continue;
                const uint8_t* line = (const uint8_t*)in + b * in_stride2;
                for (int d1 = -R; d1 <= R; ++d1) {
                    const int64_t a = i + d1;
                    if (a < 0 || a >= n1) continue;
                    for (int c = 0; c < 3; ++c) {
                        const double x = in_u8 ? mro_n0f8(line[a * 3 + c]) : ((const double*)line)[a * 3 + c];
                        s[c] = s[c] + x;
                    }
                }
            }
            double* o = (double*)((uint8_t*)out + j * out_stride2) + i * 3;
            for (int c = 0; c < 3; ++c) o[c] = rcnt * s[c];
        }
}

void mro_categorize_f64_image(const double* px, int64_t n1, int64_t n2, int64_t stride2, int K,
                              const double* mean, const double* lo, const double* hi,
                              int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                              uint8_t* out, int64_t out_stride2) {
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            const double* p = (const double*)((const uint8_t*)px + j * stride2) + i * 3;
            out[j * out_stride2 + i] = (uint8_t)mro_categorize_f64(p, 3, K, mean, lo, hi, 0);
        }
    for (int64_t k = 0; k < n_known; ++k) {
        const int64_t i = known_idx[k] % n1, j = known_idx[k] / n1;
        if (j >= n2 || known_cat[k] == 0) continue;
        const double* p = (const double*)((const uint8_t*)px + j * stride2) + i * 3;
        out[j * out_stride2 + i] = (uint8_t)mro_categorize_f64(p, 3, K, mean, lo, hi, known_cat[k]);
    }
}

/* =====================================================================================
 * `_balance` (src/balance.jl:1-42), the per-pixel part: predictions of the fitted cubic trend
 * (`_predict`, :36-42), their means (:25), and `_final_balance` (:28-42):
 *   out = clamp(x - p + mean(p), 0, 1) per channel, x = Float64(N0f8), RGB{Float64} out.
 * The least-squares fit itself (`lm(@formula(r ~ i^3 + i^2 + i + j^3 + j^2 + j), ...)`, :19-23) is host
 * code of the caller (GLM in Julia, numpy in the Python mirror) and hands over coef[3][7] =
 * (intercept, i^3, i^2, i, j^3, j^2, j) per channel; (i, j) are the 1-based indices, i the memory-fast axis.
 * PARITY UNPINNED (GLM's `predict` is a BLAS matrix-vector product and `mean` a pairwise sum: neither
 * order is visible from the reference tree).  Decisions: p = ((((((b0 + b1 i^3) + b2 i^2) + b3 i) + b4 j^3)
 * + b5 j^2) + b6 j), every product and sum rounded once, no FMA; mean(p) = sum / (n1 n2) with the sum taken
 * sequentially over chunks of 256 pixels of a line, then over the chunk sums of the line, then over the lines
 * (a fixed order that a GPU can reproduce). */
static double balance_pred(const double* b, int64_t i1, int64_t j1) {
    const double fi = (double)i1, fj = (double)j1;
    const double i2 = fi * fi, i3 = i2 * fi, j2 = fj * fj, j3 = j2 * fj;
    double p = b[0] + b[1] * i3;
    p = p + b[2] * i2;
    p = p + b[3] * fi;
    p = p + b[4] * j3;
    p = p + b[5] * j2;
    p = p + b[6] * fj;
    return p;
}

void mro_balance(const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2, const double* coef,
                 double* out, int64_t out_stride2, double means_out[3]) {
    double mean[3];
    for (int c = 0; c < 3; ++c) {
        double total = 0.0;
        for (int64_t j = 0; j < n2; ++j) {
            double line = 0.0;
            for (int64_t i0 = 0; i0 < n1; i0 += 256) {
                double chunk = 0.0;
                for (int64_t i = i0; i < n1 && i < i0 + 256; ++i) chunk = chunk + balance_pred(coef + 7 * c, i + 1, j + 1);
                line = line + chunk;
            }
            total = total + line;
        }
        mean[c] = total / (double)(n1 * n2);
        if (means_out) means_out[c] = mean[c];
    }
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            const uint8_t* p = px + j * stride2 + i * 3;
            double* o = (double*)((uint8_t*)out + j * out_stride2) + i * 3;
            for (int c = 0; c < 3; ++c) {
                double v = mro_n0f8(p[c]) - balance_pred(coef + 7 * c, i + 1, j + 1);
                v = v + mean[c];
                v = v > 0.0 ? v : 0.0;          /* max(c, 0) */
                o[c] = v < 1.0 ? v : 1.0;       /* min(1, .) */
            }
        }
}


/* =====================================================================================
 * Texture features of the reference's default match = (:color, :std, :stripe):
 *   _stds(A)            src/selectcolors.jl:706-715  (computed from the blurred RGB{Float64} image, :458)
 *   _stripes(img; radius) src/stripes.jl:1-35        (:461)
 *   _categorize on PixelProp with a match tuple   src/categories.jl:76-133 (:85-87, :96-101, :106-114, :122-133)
 * PARITY UNPINNED: Neighborhoods.jl 0.1, Statistics' reduction order over a neighbourhood and Colors.jl's
 * RGB -> HSL are not in the reference tree.  Decisions:
 *  - neighbours outside the image are the zero colour (as for blur), their number is fixed;
 *  - _stds: Window(2) = 5 x 5 incl. the centre in CartesianIndices order (d1 fastest); per channel
 *    std = sqrt(sum((x - m)^2) / 24) with m = sum(x) / 25, both sums left-folded (the two-pass `varm` of an
 *    array); pixel value (rs + gs) + bs; the image is divided by its maximum (0 / 0 = NaN as in Julia);
 *  - HSL (only s and l are used): mx = max(r, g, b), mn = min(r, g, b), l = (mx + mn) / 2,
 *    s = 0 if mx == mn, (mx - mn) / (mx + mn) if l < 0.5, else (mx - mn) / ((2 - mx) - mn);
 *  - _stripes: direction (di, dj) in vert (1, 0), horz (0, 1), angle45 (1, 1), angle135 (1, -1) (d1, d2);
 *    side a = offsets x (di, dj), x = -radius .. -1, side b = x = 1 .. radius; mean over a side =
 *    left-folded sum of ((s - s_n)^2 + (l - l_n)^2) divided by radius; direction value = min(a, b);
 *    pixel value (vert - horz, angle45 - angle135); the image is divided by
 *    m = (mean(abs first) + mean(abs last)) / 2, each mean = sum / (n1 n2) with the sum taken in the fixed
 *    order of mro_balance (chunks of 256 pixels of a line, the chunk sums of a line, the lines). */
static void hsl_sl(const double* rgb, double* s, double* l) {
    const double r = rgb[0], g = rgb[1], b = rgb[2];
    double mx = r > g ? r : g, mn = r < g ? r : g;
    mx = mx > b ? mx : b;
    mn = mn < b ? mn : b;
    *l = (mx + mn) / 2.0;
    if (mx == mn) *s = 0.0;
    else if (*l < 0.5) *s = (mx - mn) / (mx + mn);
    else *s = (mx - mn) / ((2.0 - mx) - mn);
}

static const double ZERO3[3] = {0.0, 0.0, 0.0};
static const double* px_or_zero(const double* px, int64_t n1, int64_t n2, int64_t stride2, int64_t a, int64_t b) {
    if (a < 0 || a >= n1 || b < 0 || b >= n2) return ZERO3;
    return (const double*)((const uint8_t*)px + b * stride2) + a * 3;
}

/* raw: the un-normalised plane (may be NULL); returns the maximum */
void mro_stds(const double* px, int64_t n1, int64_t n2, int64_t stride2, double* out, int64_t out_stride2) {
    double mxv = 0.0;
    int have = 0;
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            double sd[3];
            for (int c = 0; c < 3; ++c) {
                double sum = 0.0;
                for (int d2 = -2; d2 <= 2; ++d2)
                    for (int d1 = -2; d1 <= 2; ++d1) sum = sum + px_or_zero(px, n1, n2, stride2, i + d1, j + d2)[c];
                const double m = sum / 25.0;
                double ss = 0.0;
                for (int d2 = -2; d2 <= 2; ++d2)
                    for (int d1 = -2; d1 <= 2; ++d1) {
                        const double d = px_or_zero(px, n1, n2, stride2, i + d1, j + d2)[c] - m;
                        ss = ss + d * d;
                    }
                sd[c] = sqrt(ss / 24.0);
            }
            const double v = (sd[0] + sd[1]) + sd[2];
            ((double*)((uint8_t*)out + j * out_stride2))[i] = v;
            if (!have || v > mxv) { mxv = v; have = 1; }       /* maximum (no NaN can occur: finite colours) */
        }
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            double* o = (double*)((uint8_t*)out + j * out_stride2) + i;
            *o = *o / mxv;
        }
}

void mro_stripes(const double* px, int64_t n1, int64_t n2, int64_t stride2, int radius, double* out,
                 int64_t out_stride2) {
    static const int DI[4] = {1, 0, 1, 1}, DJ[4] = {0, 1, 1, -1};   /* vert, horz, angle45, angle135 */
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            double s0, l0, dir[4];
            hsl_sl(px_or_zero(px, n1, n2, stride2, i, j), &s0, &l0);
            for (int k = 0; k < 4; ++k) {
                double side[2];
                for (int sd = 0; sd < 2; ++sd) {
                    double acc = 0.0;
                    for (int t = 0; t < radius; ++t) {
                        const int x = sd == 0 ? -radius + t : 1 + t;
                        double sn, ln;
                        hsl_sl(px_or_zero(px, n1, n2, stride2, i + DI[k] * x, j + DJ[k] * x), &sn, &ln);
                        const double a = s0 - sn, b = l0 - ln;
                        acc = acc + (a * a + b * b);
                    }
                    side[sd] = acc / (double)radius;
                }
                dir[k] = side[1] < side[0] ? side[1] : side[0];
            }
            double* o = (double*)((uint8_t*)out + j * out_stride2) + 2 * i;
            o[0] = dir[0] - dir[1];
            o[1] = dir[2] - dir[3];
        }
    double mean2[2];
    for (int e = 0; e < 2; ++e) {
        double total = 0.0;
        for (int64_t j = 0; j < n2; ++j) {
            double line = 0.0;
            for (int64_t i0 = 0; i0 < n1; i0 += 256) {
                double chunk = 0.0;
                for (int64_t i = i0; i < n1 && i < i0 + 256; ++i)
                    chunk = chunk + fabs(((const double*)((const uint8_t*)out + j * out_stride2))[2 * i + e]);
                line = line + chunk;
            }
            total = total + line;
        }
        mean2[e] = total / (double)(n1 * n2);
    }
    const double m = (mean2[0] + mean2[1]) / 2.0;
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < 2 * n1; ++i) {
            double* o = (double*)((uint8_t*)out + j * out_stride2) + i;
            *o = *o / m;
        }
}

/* _categorize(x::PixelProp, categories, tolerances, match) (src/categories.jl:76-82):
 * error = mean over the matched properties, in the order (:color, :std, :stripe), of
 *   color  ((r - mr)^2 + (g - mg)^2) + (b - mb)^2                        (:91-95)
 *   std    (v - m)^2                                                      (:101)
 *   stripe ((v1 - m1)^2 + (v2 - m2)^2) / 2                                (:96-100)
 * (mean of a tuple = left-folded sum / count); first minimum; accepted when the pixel carries the
 * winner's category (:110-112) or every matched property lies in the winner's box (:113, :122-133).
 * match: bit 0 color, bit 1 std, bit 2 stripe.  A `missing` category has a colour mean of +Inf. */
int mro_categorize_props(const double* rgb, double sd, const double* stripe, int match, int K,
                         const double* mean, const double* lo, const double* hi,
                         const double* sd_mean, const double* sd_lo, const double* sd_hi,
                         const double* st_mean, const double* st_lo, const double* st_hi, int known) {
    const int cnt = (match & 1) + ((match >> 1) & 1) + ((match >> 2) & 1);
    int best = 0;
    double beste = 0.0;
    for (int c = 0; c < K; ++c) {
        double e = 0.0;
        int first = 1;
        if (mean[3 * c] == INFINITY) e = INFINITY;                       /* ismissing(cat) ? typemax(Float64) */
        else {
            if (match & 1) {
                const double d0 = rgb[0] - mean[3 * c], d1 = rgb[1] - mean[3 * c + 1], d2 = rgb[2] - mean[3 * c + 2];
                const double t = (d0 * d0 + d1 * d1) + d2 * d2;
                e = t; first = 0;
            }
            if (match & 2) {
                const double d = sd - sd_mean[c];
                const double t = d * d;
                e = first ? t : e + t; first = 0;
            }
            if (match & 4) {
                const double d0 = stripe[0] - st_mean[2 * c], d1 = stripe[1] - st_mean[2 * c + 1];
                const double t = (d0 * d0 + d1 * d1) / 2.0;
                e = first ? t : e + t; first = 0;
            }
            e = e / (double)cnt;
        }
        if (c == 0 || e < beste) { beste = e; best = c; }
    }
    if (known != 0) return best + 1 == known ? best + 1 : 0;
    int ok = 1;
    if (match & 1)
        for (int k = 0; k < 3; ++k) ok = ok && rgb[k] >= lo[3 * best + k] && rgb[k] <= hi[3 * best + k];
    if (match & 2) ok = ok && sd >= sd_lo[best] && sd <= sd_hi[best];
    if (match & 4)
        for (int k = 0; k < 2; ++k) ok = ok && stripe[k] >= st_lo[2 * best + k] && stripe[k] <= st_hi[2 * best + k];
    return ok ? best + 1 : 0;
}

void mro_categorize_props_image(const double* px, int64_t stride2, const double* sd, int64_t sd_stride2,
                                const double* stripe, int64_t st_stride2, int64_t n1, int64_t n2, int match, int K,
                                const double* mean, const double* lo, const double* hi,
                                const double* sd_mean, const double* sd_lo, const double* sd_hi,
                                const double* st_mean, const double* st_lo, const double* st_hi,
                                int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                                uint8_t* out, int64_t out_stride2) {
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t n = pass == 0 ? n1 * n2 : n_known;
        for (int64_t k = 0; k < n; ++k) {
            const int64_t lin = pass == 0 ? k : known_idx[k];
            const int kn = pass == 0 ? 0 : known_cat[k];
            const int64_t i = lin % n1, j = lin / n1;
            if (j >= n2 || (pass == 1 && kn == 0)) continue;
            const double* p = (const double*)((const uint8_t*)px + j * stride2) + i * 3;
            const double v = sd ? ((const double*)((const uint8_t*)sd + j * sd_stride2))[i] : 0.0;
            const double zero2[2] = {0.0, 0.0};
            const double* s = stripe ? (const double*)((const uint8_t*)stripe + j * st_stride2) + 2 * i : zero2;
            out[j * out_stride2 + i] = (uint8_t)mro_categorize_props(p, v, s, match, K, mean, lo, hi, sd_mean, sd_lo,
                                                                      sd_hi, st_mean, st_lo, st_hi, kn);
        }
    }
}

/* =====================================================================================
 * Polygon masks of the "clean" / "fill" buttons (src/selectcolors.jl:480-481, 491-497):
 *   mask = _polymask(polygons, A)  (:559-563: Rasters.boolmask(polygons; to = raster, shape = :polygon))
 *   masked = (1 .- mask) .* categorized;   rasterize!(fill_raster, p; fill = i, shape = :polygon)
 * PARITY UNPINNED: the rule lives in Rasters.jl 0.4 (PolygonInbounds), not in the reference tree.  Decisions
 * (DESIGN.md section 15): the raster's lookups are the index ranges (`X(ax[1])`, `Y(ax[2])`, points), so pixel (i, j)
 * (1-based) is the POINT (i, j) in polygon coordinates; a pixel belongs to a polygon when the crossing number of
 * the ray towards +d1 is odd (even-odd rule over the closed ring, the last vertex joined to the first); an edge
 * (a, b) is crossed when (a2 > j) != (b2 > j) and i < a1 + (b1 - a1) * ((j - a2) / (b2 - a2)), every operation one
 * rounded Float64 operation in that order; a ring with fewer than three vertices (the GUI's placeholder
 * [[Point2(100, 100)]], :35) is empty.  Vertices are Float32 points converted to Float64.
 * off[p] .. off[p + 1]: the vertices of ring p in xy (2 doubles each, (d1, d2)). */
static int in_ring(const double* xy, int nv, double pi, double pj) {
    if (nv < 3) return 0;
    int in = 0;
    for (int a = 0, b = nv - 1; a < nv; b = a++) {
        const double a1 = xy[2 * a], a2 = xy[2 * a + 1], b1 = xy[2 * b], b2 = xy[2 * b + 1];
        if ((a2 > pj) != (b2 > pj)) {
            const double t = (pj - a2) / (b2 - a2);
            const double x = a1 + (b1 - a1) * t;
            if (pi < x) in = !in;
        }
    }
    return in;
}

/* fill == NULL: plane[j][i] = 1 inside any ring, else 0 (the mask).
 * fill != NULL: plane[j][i] = fill[p] for the LAST ring p that contains the pixel; other pixels keep their value
 * (rasterize!(…; fill = i) ring after ring; fill = 0 everywhere is `(1 .- mask) .* A`). */
void mro_polygons(int n_rings, const int32_t* off, const double* xy, const uint8_t* fill, int64_t n1, int64_t n2,
                  uint8_t* plane, int64_t stride2) {
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            int hit = -1;
            for (int p = 0; p < n_rings; ++p)
                if (in_ring(xy + 2 * (int64_t)off[p], off[p + 1] - off[p], (double)(i + 1), (double)(j + 1))) hit = p;
            uint8_t* o = plane + j * stride2 + i;
            if (!fill) *o = hit >= 0;
            else if (hit >= 0) *o = fill[hit];
        }
}

/* =====================================================================================
 * Categorical erosion / dilation, `_shrink_grow(A, i, j)` (src/clean.jl:96-140; its call sites in the reference are
 * commented out, src/selectcolors.jl:699, :703).  PARITY UNPINNED (Neighborhoods.jl 0.1 is not in the tree).
 * Decisions: Moore(1) = the 8 neighbours without the centre in CartesianIndices order (d1 fastest); neighbours
 * outside the image are missingval = 0; output of the size of the input.
 *   _erode  (:121-129): all 8 neighbours equal the pixel -> 0, else the pixel (as written: it empties interiors);
 *   _dilate (:96-118):  a pixel != 0 stays; for a 0 pixel counts[k] = number of neighbours equal to neighbour k,
 *                       v = the neighbour at the FIRST maximum of counts; v != 0 -> v, else the first neighbour
 *                       that is not 0 (0 when there is none). */
static const int MO1[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, MO2[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static void moore8(const uint8_t* a, int64_t n1, int64_t n2, int64_t stride2, int64_t i, int64_t j, int h[8]) {
    for (int k = 0; k < 8; ++k) {
        const int64_t x = i + MO1[k], y = j + MO2[k];
        h[k] = (x >= 0 && x < n1 && y >= 0 && y < n2) ? a[y * stride2 + x] : 0;
    }
}
void mro_erode(const uint8_t* a, int64_t n1, int64_t n2, int64_t stride2, uint8_t* out, int64_t out_stride2) {
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            int h[8], all = 1;
            const int x = a[j * stride2 + i];
            moore8(a, n1, n2, stride2, i, j, h);
            for (int k = 0; k < 8; ++k) all = all && h[k] == x;
            out[j * out_stride2 + i] = all ? 0 : (uint8_t)x;
        }
}
void mro_dilate(const uint8_t* a, int64_t n1, int64_t n2, int64_t stride2, uint8_t* out, int64_t out_stride2) {
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            const int x = a[j * stride2 + i];
            if (x != 0) { out[j * out_stride2 + i] = (uint8_t)x; continue; }
            int h[8], best = 0, bc = -1;
            moore8(a, n1, n2, stride2, i, j, h);
            for (int k = 0; k < 8; ++k) {
                int c = 0;
                for (int q = 0; q < 8; ++q) c += h[q] == h[k];
                if (c > bc) { bc = c; best = k; }                 /* findmax: first maximum */
            }
            int v = h[best];
            if (v == 0)
                for (int k = 0; k < 8; ++k)
                    if (h[k] != 0) { v = h[k]; break; }           /* findfirst(!=(missingval), hood) */
            out[j * out_stride2 + i] = (uint8_t)v;
        }
}

/* `_recolor(categories, source, points)` (src/selectcolors.jl:650-655), the display layers of the GUI (:378-388):
 * label c >= 1 -> RGBA(colors[c]) = (mean r, g, b of the category's clicked points, `_mean_color` :701-704, alpha 1),
 * label 0 -> RGBA(RGB(0), 0).  colors: K x 3 doubles; out: 4 doubles per pixel.  Labels above K (a BoundsError in
 * the reference) give the transparent pixel.  No arithmetic: the means are copied. */
void mro_recolor(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int K, const double* colors,
                 double* out, int64_t out_stride2) {
    for (int64_t j = 0; j < n2; ++j)
        for (int64_t i = 0; i < n1; ++i) {
            const int c = labels[j * stride2 + i];
            double* o = (double*)((uint8_t*)out + j * out_stride2) + 4 * i;
            if (c >= 1 && c <= K) { o[0] = colors[3 * (c - 1)]; o[1] = colors[3 * (c - 1) + 1]; o[2] = colors[3 * (c - 1) + 2]; o[3] = 1.0; }
            else o[0] = o[1] = o[2] = o[3] = 0.0;
        }
}
