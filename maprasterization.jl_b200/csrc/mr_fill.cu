// mr_fill.cu -- fill_gaps (src/clean.jl:2-54): label gaps (missingval = 0) take the category that
// dominates their VonNeumann{2} neighbourhood, weighted by inverse distance; ties between the two
// strongest categories go to the one whose mean colour is closer to the original pixel.
//
// Decisions about what the reference tree does not contain (Neighborhoods.jl 0.1), see DESIGN.md
// section 11: the 12 offsets with 1 <= |d1| + |d2| <= 2 in CartesianIndices order (d1 = memory-fast
// axis fastest), Float64 weights 1/sqrt(d1^2 + d2^2) added in that order, neighbours outside the
// image read as missingval, output of the size of the input.
//
// Bound: HBM, 2 B per pixel (label in, label out) -- the neighbour labels of a gap pixel come from
// lines the CTA's neighbours have just read (L1 / L2), the original pixel is only read where the two
// strongest categories both weigh more than 1 (clean.jl:28-31).
// One thread = 16 consecutive pixels of a line (one 16-byte load / store); pixels that are not gaps
// (the bulk) are mapped through the kept-category table in registers, gap pixels run the stencil.
#include "mr_device.cuh"
#include "mr_internal.h"

namespace {

struct FillParams {
    const uint8_t* labels;
    int64_t n1, n2, stride2;
    const uint8_t* px;
    int64_t px_stride2;
    int bpp;
    const uint8_t* mask;      // nullptr: nothing masked
    int64_t mask_stride2;
    uint8_t* out;
    int64_t out_stride2;
    MrPalette pal;
    double w_diag;            // 1 / sqrt(2), computed by the host in Float64
    uint8_t rank[256];        // position of a category in `categories` (0xFF: not kept)
    int all_kept_upto;        // every category 1 .. all_kept_upto is kept (0: none): four-pixel fast path
    int vec16;                // lines can be read / written as whole 16-byte chunks
};

// colour error of the original pixel against category c (1-based): categories.jl:91-95, the
// operations of mr_eval_px (one rounded binary64 operation each, no FMA)
__device__ __forceinline__ double fill_colour_error(const FillParams& P, const uint8_t* p, int c) {
    const MrPalette& pal = P.pal;
    const int C = pal.C;
    double x0, x1, x2, x3 = 0.0;
    if (pal.colourspace == 1) {
        mr_rgb8_to_lab(pal.lin, p[0], p[1], p[2], x0, x1, x2);
    } else {
        x0 = __ddiv_rn((double)p[0], 255.0);
        x1 = __ddiv_rn((double)p[1], 255.0);
        x2 = __ddiv_rn((double)p[2], 255.0);
        if (P.bpp == 4) x3 = __ddiv_rn((double)p[3], 255.0);
    }
    const double* m = pal.mean + (size_t)(c - 1) * C;
    const double d0 = __dsub_rn(x0, __ldg(m + 0)), d1 = __dsub_rn(x1, __ldg(m + 1)), d2 = __dsub_rn(x2, __ldg(m + 2));
    double e = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
    if (C == 4) {
        const double d3 = __dsub_rn(x3, __ldg(m + 3));
        e = __dadd_rn(e, __dmul_rn(d3, d3));
    }
    return e;
}

// the stencil of one gap pixel (clean.jl:15-34)
__device__ __noinline__ unsigned fill_gap_pixel(const FillParams& P, int64_t i, int64_t j) {
    constexpr int D1[12] = {0, -1, 0, 1, -2, -1, 1, 2, -1, 0, 1, 0};
    constexpr int D2[12] = {-2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2};
    // distinct kept categories among the neighbours, with their weights (at most 12)
    unsigned cat[12];
    double acc[12];
    int nd = 0, missing = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const int64_t a = i + D1[k], b = j + D2[k];
        unsigned l = 0;
        if (a >= 0 && a < P.n1 && b >= 0 && b < P.n2) l = __ldg(P.labels + b * P.stride2 + a);
        if (l == 0) { ++missing; continue; }
        if (P.rank[l] == 0xFFu) continue;   // not a kept category: neither missing nor counted
        const int sq = D1[k] * D1[k] + D2[k] * D2[k];
        const double w = sq == 1 ? 1.0 : (sq == 2 ? P.w_diag : 0.5);
        int s = 0;
        while (s < nd && cat[s] != l) ++s;
        if (s == nd) { cat[nd] = l; acc[nd] = 0.0; ++nd; }
        acc[s] = __dadd_rn(acc[s], w);      // clean.jl:49, in neighbourhood order
    }
    if (missing > 12 - 12 / 4) return 0u;   // clean.jl:20
    // findmax = first maximum in the order of `categories`; second = the same among the others
    int im = -1, is = -1;
    for (int s = 0; s < nd; ++s)
        if (im < 0 || acc[s] > acc[im] || (acc[s] == acc[im] && P.rank[cat[s]] < P.rank[cat[im]])) im = s;
    if (im < 0 || !(acc[im] > 1.0)) return 0u;   // clean.jl:23
    for (int s = 0; s < nd; ++s) {
        if (s == im) continue;
        if (is < 0 || acc[s] > acc[is] || (acc[s] == acc[is] && P.rank[cat[s]] < P.rank[cat[is]])) is = s;
    }
    if (is >= 0 && acc[is] > 1.0) {              // clean.jl:28-31
        const uint8_t* p = P.px + j * P.px_stride2 + i * P.bpp;
        const double ea = fill_colour_error(P, p, (int)cat[im]);
        const double eb = fill_colour_error(P, p, (int)cat[is]);
        return ea <= eb ? cat[im] : cat[is];
    }
    return cat[im];
}

__device__ __forceinline__ unsigned fill_pixel(const FillParams& P, int64_t i, int64_t j, unsigned v, unsigned m) {
    if (m) return v;                                  // clean.jl:13
    if (v != 0) return P.rank[v] != 0xFFu ? v : 0u;   // clean.jl:16-17, 35-36
    return fill_gap_pixel(P, i, j);
}

// Few registers (the stencil of a gap pixel is a call and may spill; the pass-through path must not limit
// the occupancy) and the load of the next line's chunk in flight while this one is processed: the kernel
// is a stream with very little work per byte, so bytes in flight are what it lives on.
__global__ void __launch_bounds__(256) k_fill_gaps(const __grid_constant__ FillParams P) {
    // grid.x covers one line in 16-pixel chunks, grid.y strides over the lines (no 64-bit division per chunk)
    // (narrow images: blockDim.y lines per block)
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= P.n1) return;
    const int64_t jstep = (int64_t)gridDim.y * blockDim.y;
    int64_t j = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
    uint4 nextq = make_uint4(0u, 0u, 0u, 0u);
    if (P.vec16 && j < P.n2) nextq = __ldg(reinterpret_cast<const uint4*>(P.labels + j * P.stride2 + i0));
    for (; j < P.n2; j += jstep) {
        const uint8_t* src = P.labels + j * P.stride2 + i0;
        uint8_t* dst = P.out + j * P.out_stride2 + i0;
        const int n = (int)min((int64_t)16, P.n1 - i0);
        uint32_t w[4] = {0u, 0u, 0u, 0u}, mk[4] = {0u, 0u, 0u, 0u};
        if (P.vec16) {
            const uint4 q = nextq;
            if (j + jstep < P.n2) nextq = __ldg(reinterpret_cast<const uint4*>(src + jstep * P.stride2));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            if (P.mask) {
                const uint4 mq = __ldg(reinterpret_cast<const uint4*>(P.mask + j * P.mask_stride2 + i0));
                mk[0] = mq.x; mk[1] = mq.y; mk[2] = mq.z; mk[3] = mq.w;
            }
        } else {
            for (int k = 0; k < n; ++k) {
                w[k >> 2] |= (uint32_t)__ldg(src + k) << (8 * (k & 3));
                if (P.mask) mk[k >> 2] |= (uint32_t)__ldg(P.mask + j * P.mask_stride2 + i0 + k) << (8 * (k & 3));
            }
        }
        uint32_t o[4];
        // four pixels at a time: unmasked, no gap (no zero byte) and every label in 1 .. all_kept_upto
        // (no byte above it; bytes >= 128 take the slow path) -> the word passes through unchanged
        const uint32_t above = (uint32_t)(127 - P.all_kept_upto) * 0x01010101u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t w4 = w[g];
            const uint32_t zero = ~(((w4 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w4 | 0x7F7F7F7Fu);   // 0x80 in every zero byte
            const uint32_t big = (((w4 & 0x7F7F7F7Fu) + above) | w4) & 0x80808080u;          // 0x80 where label > all_kept_upto
            if (4 * g + 4 <= n && (zero | big | mk[g]) == 0u) {
                o[g] = w4;
                continue;
            }
            uint32_t r = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int k = 4 * g + b;
                unsigned v = (w[g] >> (8 * b)) & 0xFFu;
                if (k < n) v = fill_pixel(P, i0 + k, j, v, (mk[g] >> (8 * b)) & 0xFFu);
                r |= v << (8 * b);
            }
            o[g] = r;
        }
        if (P.vec16) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            for (int k = 0; k < n; ++k) dst[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
        }
    }
}

// clean.jl:14: a clicked point whose category is kept keeps that category (unless masked)
__global__ void k_fill_known(const FillParams P, int64_t n_known, const int64_t* __restrict__ idx,
                             const uint8_t* __restrict__ cat) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_known) return;
    const int64_t p = idx[t];
    if (p < 0 || p >= P.n1 * P.n2) return;
    const int64_t j = p / P.n1, i = p - j * P.n1;
    const unsigned kn = cat[t];
    if (kn == 0 || P.rank[kn] == 0xFFu) return;
    if (P.mask && P.mask[j * P.mask_stride2 + i]) return;
    P.out[j * P.out_stride2 + i] = (uint8_t)kn;
}

}  // namespace

cudaError_t mr_launch_fill_gaps(const MrPalette& pal, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                                const uint8_t* px, int64_t px_stride2, int bpp, const uint8_t* mask,
                                int64_t mask_stride2, const uint8_t* categories, int n_categories, int64_t n_known,
                                const int64_t* known_idx, const uint8_t* known_cat, uint8_t* out, int64_t out_stride2,
                                int sm_count, cudaStream_t s, int* n_launches) {
    FillParams P;
    P.labels = labels; P.n1 = n1; P.n2 = n2; P.stride2 = stride2;
    P.px = px; P.px_stride2 = px_stride2; P.bpp = bpp;
    P.mask = mask; P.mask_stride2 = mask_stride2;
    P.out = out; P.out_stride2 = out_stride2;
    P.pal = pal;
    P.w_diag = 1.0 / sqrt(2.0);
    for (int c = 0; c < 256; ++c) P.rank[c] = 0xFF;
    for (int k = n_categories - 1; k >= 0; --k) P.rank[categories[k]] = (uint8_t)k;   // first occurrence wins
    P.all_kept_upto = 0;
    while (P.all_kept_upto < 127 && P.rank[P.all_kept_upto + 1] != 0xFF) ++P.all_kept_upto;
    P.vec16 = ((uintptr_t)labels % 16 == 0) && ((uintptr_t)out % 16 == 0) && (stride2 % 16 == 0) &&
              (out_stride2 % 16 == 0) && (n1 % 16 == 0) &&
              (!mask || (((uintptr_t)mask % 16 == 0) && (mask_stride2 % 16 == 0)));
    const int64_t chunks = (n1 + 15) / 16;
    int bx = 256;
    while (bx > 32 && bx / 2 >= chunks) bx /= 2;
    const int by = 256 / bx;
    const int64_t gx = (chunks + bx - 1) / bx;
    int64_t gy = ((int64_t)sm_count * 32 + gx - 1) / gx;   // about 32 blocks per SM in total
    if (gy > (n2 + by - 1) / by) gy = (n2 + by - 1) / by;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    k_fill_gaps<<<dim3((unsigned)gx, (unsigned)gy), dim3(bx, by), 0, s>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    *n_launches += 1;
    if (n_known > 0) {
        k_fill_known<<<(unsigned)((n_known + 255) / 256), 256, 0, s>>>(P, n_known, known_idx, known_cat);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        *n_launches += 1;
    }
    return cudaSuccess;
}
