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

constexpr int FILL_GROUP = 16;   // consecutive lines a block processes before it jumps (multiple of every blockDim.y)

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
// s_n0f8[v] = Float64(N0f8(v)) = v / 255 (correctly rounded), tabulated once per block
__device__ __forceinline__ double fill_colour_error(const FillParams& P, const double* __restrict__ s_n0f8,
                                                    const uint8_t* p, int c) {
    const MrPalette& pal = P.pal;
    const int C = pal.C;
    double x0, x1, x2, x3 = 0.0;
    if (pal.colourspace == 1) {
        mr_rgb8_to_lab(pal.lin, p[0], p[1], p[2], x0, x1, x2);
    } else {
        x0 = s_n0f8[p[0]];
        x1 = s_n0f8[p[1]];
        x2 = s_n0f8[p[2]];
        if (P.bpp == 4) x3 = s_n0f8[p[3]];
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

// the stencil of one gap pixel (clean.jl:15-34).  The 12 neighbour labels sit in three registers (one byte
// each); the distinct labels are taken one after the other (first uncounted neighbour), the neighbours that
// carry the label are found by a byte compare, and their weights are added IN NEIGHBOURHOOD ORDER
// (clean.jl:43-54: `acc += 1/ds[i]`); best / second best (weight, position in `categories`) are kept
// (findmax and the reduce of clean.jl:24-27 = first of equal weights).
__device__ __forceinline__ unsigned fill_gap_pixel(const FillParams& P, const uint8_t* __restrict__ s_rank,
                                                   const double* __restrict__ s_n0f8, int64_t i, int64_t j) {
    constexpr int D1[12] = {0, -1, 0, 1, -2, -1, 1, 2, -1, 0, 1, 0};
    constexpr int D2[12] = {-2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2};
    // weight class of neighbour m, two bits each: 0 -> 1 (distance 1), 1 -> 1/sqrt(2), 2 -> 1/2 (distance 2)
    constexpr unsigned WCLASS = 2u | (1u << 2) | (0u << 4) | (1u << 6) | (2u << 8) | (0u << 10) | (0u << 12) |
                                (2u << 14) | (1u << 16) | (0u << 18) | (1u << 20) | (2u << 22);
    uint32_t lw[3] = {0u, 0u, 0u};
    if (i >= 2 && i + 2 < P.n1 && j >= 2 && j + 2 < P.n2) {   // interior: no bounds tests
        const uint8_t* c0 = P.labels + j * P.stride2 + i;
#pragma unroll
        for (int k = 0; k < 12; ++k)
            lw[k >> 2] |= (uint32_t)__ldg(c0 + D2[k] * P.stride2 + D1[k]) << (8 * (k & 3));
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const int64_t a = i + D1[k], b = j + D2[k];
            unsigned l = 0u;
            if (a >= 0 && a < P.n1 && b >= 0 && b < P.n2) l = __ldg(P.labels + b * P.stride2 + a);
            lw[k >> 2] |= l << (8 * (k & 3));
        }
    }
    auto eq_mask = [&](unsigned c) {      // 12-bit mask of the neighbours whose label is c
        uint32_t m = 0u;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const uint32_t x = lw[g] ^ (c * 0x01010101u);
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every zero byte
            m |= ((((z >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * g);                    // bits 0, 8, 16, 24 -> 0..3
        }
        return m;
    };
    uint32_t rem = 0xFFFu & ~eq_mask(0u);                 // neighbours that carry a label, not yet counted
    if (12 - __popc(rem) > 12 - 12 / 4) return 0u;        // clean.jl:20: mostly missing
    double a1 = 0.0, a2 = 0.0;                            // best / second weight (0: none)
    unsigned r1 = 0x100u, r2 = 0x100u, c1 = 0u, c2 = 0u;
    while (rem) {
        const int k = __ffs(rem) - 1;
        const uint32_t word = k < 4 ? lw[0] : (k < 8 ? lw[1] : lw[2]);
        const unsigned c = (word >> (8 * (k & 3))) & 0xFFu;
        uint32_t mm = eq_mask(c);
        rem &= ~mm;
        const unsigned r = s_rank[c];
        if (r == 0xFFu) continue;                         // not a kept category: neither missing nor counted
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < 12; ++m) {                    // ascending m = neighbourhood order; predicated additions
            const unsigned wc = (WCLASS >> (2 * m)) & 3u;
            if ((mm >> m) & 1u) acc = __dadd_rn(acc, wc == 0u ? 1.0 : (wc == 1u ? P.w_diag : 0.5));
        }
        if (acc > a1 || (acc == a1 && r < r1)) {
            a2 = a1; r2 = r1; c2 = c1;
            a1 = acc; r1 = r; c1 = c;
        } else if (acc > a2 || (acc == a2 && r < r2)) {
            a2 = acc; r2 = r; c2 = c;
        }
    }
    if (!(a1 > 1.0)) return 0u;                           // clean.jl:23
    if (a2 > 1.0) {                                       // clean.jl:28-31
        const uint8_t* p = P.px + j * P.px_stride2 + i * P.bpp;
        const double ea = fill_colour_error(P, s_n0f8, p, (int)c1);
        const double eb = fill_colour_error(P, s_n0f8, p, (int)c2);
        return ea <= eb ? c1 : c2;
    }
    return c1;
}

// One thread = 16 consecutive pixels of a line; a warp = 512 consecutive pixels.  Words of four pixels that
// are unmasked, without a gap and entirely inside 1 .. all_kept_upto pass through; the other non-gap pixels
// map through the rank table; gap pixels are only MARKED, the 16 bytes are stored, and then the warp
// spreads its gap pixels evenly over its lanes (inclusive scan of the per-lane gap counts, owner by binary
// search) and patches single bytes -- a sheet with 5 % gaps would otherwise run the stencil on one lane in
// sixteen.  grid.x covers one line, grid.y strides over the lines; narrow images: blockDim.y lines per block.
__global__ void __launch_bounds__(256) k_fill_gaps(const __grid_constant__ FillParams P) {
    __shared__ uint8_t s_rank[256];
    __shared__ double s_n0f8[256];
    for (int t = threadIdx.y * blockDim.x + threadIdx.x; t < 256; t += blockDim.x * blockDim.y) {
        s_rank[t] = P.rank[t];
        s_n0f8[t] = __ddiv_rn((double)t, 255.0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int n = i0 >= P.n1 ? 0 : (int)min((int64_t)16, P.n1 - i0);
    // a block walks down GROUPS of FILL_GROUP consecutive lines (the neighbour lines of a gap pixel are then lines
    // this SM has just read: L1 hits instead of a 32-byte sector from L2 per neighbour), groups stride over grid.y
    auto advance = [&](int64_t y) {       // the line after y in this thread's sequence
        const int64_t g0 = y / FILL_GROUP * FILL_GROUP;
        const int64_t yn = y + blockDim.y;
        if (yn < g0 + FILL_GROUP) return yn;
        return g0 + (int64_t)gridDim.y * FILL_GROUP + threadIdx.y;
    };
    int64_t j = (int64_t)blockIdx.y * FILL_GROUP + threadIdx.y;
    uint4 nextq = make_uint4(0u, 0u, 0u, 0u);
    if (P.vec16 && n && j < P.n2) nextq = __ldg(reinterpret_cast<const uint4*>(P.labels + j * P.stride2 + i0));
    const uint32_t above = (uint32_t)(127 - P.all_kept_upto) * 0x01010101u;
    const int64_t n2r = (P.n2 + FILL_GROUP - 1) / FILL_GROUP * FILL_GROUP;   // every warp of a block sees the same groups
    for (; j < n2r; j = advance(j)) {     // warp-uniform: a warp holds consecutive chunks of ONE line
        const int64_t jn = advance(j);
        if (j >= P.n2) continue;
        const uint8_t* src = P.labels + j * P.stride2 + i0;
        uint8_t* dst = P.out + j * P.out_stride2 + i0;
        uint32_t w[4] = {0u, 0u, 0u, 0u}, mk[4] = {0u, 0u, 0u, 0u};
        if (P.vec16) {
            const uint4 q = nextq;
            if (n && jn < P.n2) nextq = __ldg(reinterpret_cast<const uint4*>(P.labels + jn * P.stride2 + i0));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            if (P.mask && n) {
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
        uint32_t gaps = 0u;               // bit k: pixel i0 + k is an unmasked gap
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t w4 = w[g];
            const uint32_t zero = ~(((w4 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w4 | 0x7F7F7F7Fu);   // 0x80 in every zero byte
            const uint32_t big = (((w4 & 0x7F7F7F7Fu) + above) | w4) & 0x80808080u;          // 0x80 where label > all_kept_upto
            if (4 * g + 4 <= n && big == 0u) {
                // every label of the word is 0 or kept: the bytes pass through; unmasked zero bytes are gaps
                const uint32_t mz = ~(((mk[g] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | mk[g] | 0x7F7F7F7Fu);   // 0x80: pixel not masked
                const uint32_t gz = zero & mz;
                gaps |= ((((gz >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * g);
                o[g] = w4;
                continue;
            }
            uint32_t r = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int k = 4 * g + b;
                unsigned v = (w4 >> (8 * b)) & 0xFFu;
                const unsigned m = (mk[g] >> (8 * b)) & 0xFFu;
                if (k < n && !m) {                                     // clean.jl:13: a masked pixel stays
                    if (v != 0u) v = s_rank[v] != 0xFFu ? v : 0u;      // clean.jl:16-17, 35-36
                    else gaps |= 1u << k;
                }
                r |= v << (8 * b);
            }
            o[g] = r;
        }
        if (n) {
            if (P.vec16) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                for (int k = 0; k < n; ++k) dst[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
            }
        }
        // ---- the warp's gap pixels, spread evenly over its lanes
        if (!__any_sync(0xffffffffu, gaps != 0u)) continue;
        __syncwarp();                     // orders the 16-byte stores above before the byte patches below
        const int c = __popc(gaps);
        int S = c;                        // inclusive scan of the gap counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, S, d);
            if (lane >= d) S += v;
        }
        const int T = __shfl_sync(0xffffffffu, S, 31);
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            int owner = 0;                // number of lanes whose inclusive scan is <= t
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, S, (owner + step - 1) & 31);
                if (v <= t) owner += step;
            }
            owner &= 31;
            const int before = __shfl_sync(0xffffffffu, S - c, owner);
            uint32_t m = __shfl_sync(0xffffffffu, gaps, owner);
            if (t < T) {
                for (int r = t - before; r > 0; --r) m &= m - 1;       // skip r set bits
                const int k = __ffs(m) - 1;
                const int64_t i = i0 + (int64_t)(owner - lane) * 16 + k;
                P.out[j * P.out_stride2 + i] = (uint8_t)fill_gap_pixel(P, s_rank, s_n0f8, i, j);
            }
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
    const int64_t groups = (n2 + FILL_GROUP - 1) / FILL_GROUP;
    if (gy > groups) gy = groups;
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
