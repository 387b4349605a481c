// mr_generic.cu -- table build + the general-purpose fused kernel.
//
//  * k_build_lut / k_build_coarse: per palette, evaluate the reference arithmetic
//    (mr_device.cuh) once for each of the 2^24 N0f8 colours -> exact 16 MiB label table,
//    then reduce it to a 32 KiB table of 8x8x8-colour cells (label, or MIXED).
//    With match = (:color,) the label is a pure function of the 24-bit colour
//    (src/categories.jl:76-95), so a table lookup is bit-exact by construction.
//  * k_generic: any stride/alignment, RGB or RGBA (direct Float64 evaluation), any R <= 7,
//    label input (clean only).  It is the on-GPU general path; mr_fast.cu holds the
//    tuned packed-RGB path.  There is no CPU path.
#include "mr_device.cuh"

// ------------------------------------------------------------------ table build
__global__ void __launch_bounds__(256) k_build_lut(MrPalette pal, uint8_t* __restrict__ lut) {
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // r | g<<8 | b<<16
    if (idx >= MR_LUT_SIZE) return;
    int r = idx & 255, g = (idx >> 8) & 255, b = (idx >> 16) & 255;
    lut[idx] = (uint8_t)mr_eval_px(pal, r, g, b, 255, false, 0);
}

// One thread per coarse cell (r>>3, g>>3, b>>3): uniform label or MIXED (layout: MR_COARSE_IDX).
__global__ void __launch_bounds__(128) k_build_coarse(const uint8_t* __restrict__ lut,
                                                      uint8_t* __restrict__ coarse) {
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (1u << 15)) return;
    unsigned r0 = (c & 31u) << 3, g0 = ((c >> 5) & 31u) << 3, b0 = ((c >> 10) & 31u) << 3;
    unsigned first = lut[r0 | (g0 << 8) | (b0 << 16)];
    unsigned long long want = 0x0101010101010101ull * first;
    bool uniform = true;
    for (unsigned b = 0; b < 8; ++b)
        for (unsigned g = 0; g < 8; ++g) {
            unsigned long long v = *reinterpret_cast<const unsigned long long*>(
                lut + (r0 | ((g0 + g) << 8) | ((b0 + b) << 16)));
            uniform = uniform && (v == want);
        }
    coarse[MR_COARSE_IDX(c & 31u, (c >> 5) & 31u, c >> 10)] =
        (uniform && first != MR_MIXED) ? (uint8_t)first : (uint8_t)MR_MIXED;
}

// RGB colour space: one block per (g, b) pair, thread = r.  The squared g / b (and alpha = 255)
// differences of every category are computed once per block; per colour and category the
// reference's left fold ((q_r + q_g) + q_b) [+ q_a] is then two additions -- the same individually
// rounded operations as mr_eval_px, in the same order (src/categories.jl:91-95), first minimum,
// box test on the winner (:119-121, :131-133).
__global__ void __launch_bounds__(256) k_build_lut_rgb(MrPalette pal, uint8_t* __restrict__ lut) {
    __shared__ double s_mr[254], s_qg[254], s_qb[254], s_qa[254];
    const int g = blockIdx.x & 255, b = blockIdx.x >> 8, r = threadIdx.x;
    const int C = pal.C, K = pal.K;
    const double xg = __ddiv_rn((double)g, 255.0), xb = __ddiv_rn((double)b, 255.0), xr = __ddiv_rn((double)r, 255.0);
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        const double* m = pal.mean + (size_t)c * C;
        s_mr[c] = __ldg(m + 0);
        const double dg = __dsub_rn(xg, __ldg(m + 1)), db = __dsub_rn(xb, __ldg(m + 2));
        s_qg[c] = __dmul_rn(dg, dg);
        s_qb[c] = __dmul_rn(db, db);
        if (C == 4) {
            const double da = __dsub_rn(1.0, __ldg(m + 3));   // Float64(N0f8(255)) = 1.0
            s_qa[c] = __dmul_rn(da, da);
        }
    }
    __syncthreads();
    int best = 0;
    double beste = 0.0;
    for (int c = 0; c < K; ++c) {
        const double dr = __dsub_rn(xr, s_mr[c]);
        double e = __dadd_rn(__dadd_rn(__dmul_rn(dr, dr), s_qg[c]), s_qb[c]);
        if (C == 4) e = __dadd_rn(e, s_qa[c]);
        if (c == 0 || e < beste) { beste = e; best = c; }   // first minimum
    }
    const double* l = pal.lo + (size_t)best * C;
    const double* h = pal.hi + (size_t)best * C;
    const bool ok = (xr >= __ldg(l + 0)) && (xr <= __ldg(h + 0)) && (xg >= __ldg(l + 1)) && (xg <= __ldg(h + 1)) &&
                    (xb >= __ldg(l + 2)) && (xb <= __ldg(h + 2));
    lut[(unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16)] = ok ? (uint8_t)(best + 1) : (uint8_t)0;
}

// ---- the same in two steps (RGB colour space), four times faster: the squared channel differences of the
// 256 N0f8 values against every category are tabulated once (3 x K x 256 doubles), then the error of a
// colour is two additions per category, (q_r + q_g) + q_b [+ q_a] -- the very operations above, in the
// same order.  Block = (b, 32 values of g), thread = r: q_r and q_b of the thread sit in registers for
// K <= 16 (template), q_g of the block's 32 g values in shared memory (broadcast reads).
#define MR_QTAB_DOUBLES (3 * 254 * 256)
__global__ void k_sq_tables(MrPalette pal, double* __restrict__ qtab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;   // (ch, c, v)
    const int K = pal.K;
    if (t >= 3 * K * 256) return;
    const int v = t & 255, c = (t >> 8) % K, ch = (t >> 8) / K;
    const double d = __dsub_rn(__ddiv_rn((double)v, 255.0), __ldg(pal.mean + (size_t)c * pal.C + ch));
    qtab[((size_t)ch * K + c) * 256 + v] = __dmul_rn(d, d);
}

template <int KR>   // KR > 0: K <= KR, per-thread tables in registers; KR == 0: any K, tables read per category
__global__ void __launch_bounds__(256) k_build_lut_rgb2(MrPalette pal, const double* __restrict__ qtab,
                                                        uint8_t* __restrict__ lut) {
    constexpr int GB = KR > 0 ? 32 : 16;         // g values per block
    constexpr int KS = KR > 0 ? KR : 254;
    __shared__ double s_qg[GB * KS];             // [gi][c]
    __shared__ double s_qa[254];
    const int b = blockIdx.x / (256 / GB), g0 = (blockIdx.x % (256 / GB)) * GB, r = threadIdx.x;
    const int C = pal.C, K = pal.K;
    const double* qr = qtab;                     // [c][v]
    const double* qg = qtab + (size_t)K * 256;
    const double* qb = qtab + (size_t)2 * K * 256;
    for (int t = threadIdx.x; t < GB * K; t += blockDim.x) {
        const int gi = t / K, c = t - gi * K;
        s_qg[gi * K + c] = __ldg(qg + (size_t)c * 256 + g0 + gi);
    }
    if (C == 4)
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            const double da = __dsub_rn(1.0, __ldg(pal.mean + (size_t)c * C + 3));   // Float64(N0f8(255)) = 1.0
            s_qa[c] = __dmul_rn(da, da);
        }
    double rq[KR > 0 ? KR : 1], bq[KR > 0 ? KR : 1];
    if (KR > 0) {
#pragma unroll
        for (int c = 0; c < KR; ++c) {
            rq[c] = c < K ? __ldg(qr + (size_t)c * 256 + r) : 0.0;
            bq[c] = c < K ? __ldg(qb + (size_t)c * 256 + b) : 0.0;
        }
    }
    __syncthreads();
    const double xr = __ddiv_rn((double)r, 255.0), xb = __ddiv_rn((double)b, 255.0);
    for (int gi = 0; gi < GB; ++gi) {
        const int g = g0 + gi;
        const double* sg = s_qg + gi * K;
        int best = 0;
        double beste = 0.0;
        if (KR > 0) {
#pragma unroll
            for (int c = 0; c < KR; ++c) {
                if (c < K) {
                    double e = __dadd_rn(__dadd_rn(rq[c], sg[c]), bq[c]);
                    if (C == 4) e = __dadd_rn(e, s_qa[c]);
                    if (c == 0 || e < beste) { beste = e; best = c; }   // first minimum
                }
            }
        } else {
            for (int c = 0; c < K; ++c) {
                double e = __dadd_rn(__dadd_rn(__ldg(qr + (size_t)c * 256 + r), sg[c]), __ldg(qb + (size_t)c * 256 + b));
                if (C == 4) e = __dadd_rn(e, s_qa[c]);
                if (c == 0 || e < beste) { beste = e; best = c; }
            }
        }
        const double xg = __ddiv_rn((double)g, 255.0);
        const double* l = pal.lo + (size_t)best * C;
        const double* h = pal.hi + (size_t)best * C;
        const bool ok = (xr >= __ldg(l + 0)) && (xr <= __ldg(h + 0)) && (xg >= __ldg(l + 1)) && (xg <= __ldg(h + 1)) &&
                        (xb >= __ldg(l + 2)) && (xb <= __ldg(h + 2));
        lut[(unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16)] = ok ? (uint8_t)(best + 1) : (uint8_t)0;
    }
}

cudaError_t mr_launch_build_lut(const MrPalette& pal, uint8_t* lut, uint8_t* coarse, double* qtab, cudaStream_t s,
                                int* n_launches) {
    if (pal.colourspace == 0 && pal.K <= 254 && qtab) {
        k_sq_tables<<<(3 * pal.K * 256 + 255) / 256, 256, 0, s>>>(pal, qtab);
        if (pal.K <= 16) k_build_lut_rgb2<16><<<256 * 8, 256, 0, s>>>(pal, qtab, lut);
        else k_build_lut_rgb2<0><<<256 * 16, 256, 0, s>>>(pal, qtab, lut);
        *n_launches += 1;
    } else if (pal.colourspace == 0 && pal.K <= 254) {
        k_build_lut_rgb<<<65536, 256, 0, s>>>(pal, lut);
    } else {
        k_build_lut<<<MR_LUT_SIZE / 256, 256, 0, s>>>(pal, lut);
    }
    k_build_coarse<<<(1u << 15) / 128, 128, 0, s>>>(lut, coarse);
    *n_launches += 2;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ generic fused kernel
#define G_TW 128
#define G_TH 16
#define G_MAXR 7
#define G_VOID 0xFFu   // outside the image (window clipping); never a real label in smem

// Label of one input element.  labels_in: the input already is a label plane (bpp == 1).
__device__ __forceinline__ unsigned g_label(const MrJob& job, const uint8_t* __restrict__ p,
                                            const uint8_t* __restrict__ s_coarse, bool labels_in,
                                            bool direct, unsigned& fine) {
    if (labels_in) return p[0];
    if (direct) return (unsigned)mr_eval_px(job.pal, p[0], p[1], p[2], job.bpp == 4 ? p[3] : 255,
                                            job.bpp == 4, 0);
    unsigned r = p[0], g = p[1], b = p[2];
    unsigned v = s_coarse[MR_COARSE_IDX(r >> 3, g >> 3, b >> 3)];
    if (v == MR_MIXED) {
        v = __ldg(job.lut + (r | (g << 8) | (b << 16)));
        ++fine;
    }
    return v;
}

__global__ void __launch_bounds__(256) k_generic(MrJob job, bool labels_in, bool direct) {
    __shared__ uint8_t s_coarse[MR_COARSE_SIZE];
    __shared__ uint8_t s_lab[(G_TH + 2 * G_MAXR) * (G_TW + 2 * G_MAXR)];
    const int R = job.R;
    const int SW = G_TW + 2 * R;
    const bool use_coarse = !labels_in && !direct;
    if (use_coarse) {
        const uint4* src = reinterpret_cast<const uint4*>(job.coarse);
        uint4* dst = reinterpret_cast<uint4*>(s_coarse);
        for (int i = threadIdx.x; i < (int)(MR_COARSE_SIZE / 16); i += blockDim.x) dst[i] = src[i];
    }
    const int64_t tiles_x = (job.n1 + G_TW - 1) / G_TW;
    const int64_t tiles_y = (job.n2 + G_TH - 1) / G_TH;
    const int64_t tiles = tiles_x * tiles_y * job.n_sheets;
    unsigned c_nomatch = 0, c_fine = 0, c_changed = 0, c_pix = 0;

    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t sheet = t / (tiles_x * tiles_y);
        const int64_t rem = t - sheet * tiles_x * tiles_y;
        const int64_t ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int64_t i0 = tx * G_TW, j0 = ty * G_TH;
        const uint8_t* px = job.px + sheet * job.sheet_stride;
        uint8_t* out = job.out + sheet * job.out_sheet_stride;
        __syncthreads();   // previous tile's votes are done; also publishes s_coarse
        // phase 1: labels of the tile + halo
        const int region = SW * (G_TH + 2 * R);
        for (int e = threadIdx.x; e < region; e += blockDim.x) {
            const int ly = e / SW, lx = e - ly * SW;
            const int64_t i = i0 + lx - R, j = j0 + ly - R;
            unsigned v = G_VOID;
            if (i >= 0 && i < job.n1 && j >= -(int64_t)job.halo_lo && j < job.n2 + job.halo_hi) {
                unsigned fine = 0;
                v = g_label(job, px + j * job.stride2 + i * job.bpp, s_coarse, labels_in, direct, fine);
                // count the near-tie path once per owned pixel (not for halo re-reads)
                if (lx >= R && lx < R + G_TW && ly >= R && ly < R + G_TH && j >= 0 && j < job.n2)
                    c_fine += fine;
            }
            s_lab[e] = (uint8_t)v;
        }
        __syncthreads();
        // phase 2: Window{R} majority, ties -> lowest label (Appendix A.3)
        for (int e = threadIdx.x; e < G_TW * G_TH; e += blockDim.x) {
            const int oy = e / G_TW, ox = e - oy * G_TW;
            const int64_t i = i0 + ox, j = j0 + oy;
            if (i >= job.n1 || j >= job.n2) continue;
            const uint8_t* w = s_lab + oy * SW + ox;   // top-left of the window
            const unsigned centre = w[R * SW + R];
            unsigned best = centre;
            if (R > 0) {
                int bc = 0;
                unsigned prev = G_VOID;
                best = G_VOID;
                const int side = 2 * R + 1;
                for (int a = 0; a < side * side; ++a) {
                    const unsigned c = w[(a / side) * SW + (a % side)];
                    if (c == G_VOID || c == prev) continue;
                    prev = c;
                    if (c == best) continue;
                    int cnt = 0;
                    for (int y = 0; y < side; ++y)
                        for (int x = 0; x < side; ++x) cnt += (w[y * SW + x] == c);
                    if (cnt > bc || (cnt == bc && c < best)) { bc = cnt; best = c; }
                }
            }
            out[j * job.out_stride2 + i] = (uint8_t)best;
            ++c_pix;
            c_nomatch += (best == 0);
            c_changed += (best != centre);
        }
    }
    if (job.counters) {
        // warp-aggregate, one atomic per warp and counter
        for (int o = 16; o > 0; o >>= 1) {
            c_pix += __shfl_xor_sync(0xffffffffu, c_pix, o);
            c_nomatch += __shfl_xor_sync(0xffffffffu, c_nomatch, o);
            c_fine += __shfl_xor_sync(0xffffffffu, c_fine, o);
            c_changed += __shfl_xor_sync(0xffffffffu, c_changed, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&job.counters->n_pixels, (unsigned long long)c_pix);
            atomicAdd(&job.counters->n_nomatch, (unsigned long long)c_nomatch);
            atomicAdd(&job.counters->n_fine, (unsigned long long)c_fine);
            atomicAdd(&job.counters->n_changed, (unsigned long long)c_changed);
        }
    }
}

cudaError_t mr_launch_generic(const MrJob& job, bool labels_in, int sm_count, cudaStream_t s,
                              int* n_launches) {
    if (job.R > G_MAXR) return cudaErrorInvalidValue;
    const bool direct = !labels_in && job.lut == nullptr;
    const int64_t tiles = ((job.n1 + G_TW - 1) / G_TW) * ((job.n2 + G_TH - 1) / G_TH) * job.n_sheets;
    if (tiles == 0) return cudaSuccess;
    int64_t grid = (int64_t)sm_count * 4;
    if (grid > tiles) grid = tiles;
    k_generic<<<(unsigned)grid, 256, 0, s>>>(job, labels_in, direct);
    *n_launches += 1;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ known-category fix-up
// src/categories.jl:110-112 + src/selectcolors.jl:670-681: a clicked pixel with known
// category k is labelled (best == k) ? best : 0, skipping the box test.  Sparse (<= a few
// hundred pixels), so it runs as its own tiny kernel on the categorised plane.
// The indices count from sheet line 0; the plane `labels` holds the band-local lines
// [j_first, j_first + n_lines) of a band whose line 0 is sheet line `origin` (whole image: 0, 0, n2).
// Points outside the plane are skipped (never a write outside it).
__global__ void k_known_fixup(MrJob job, int64_t n_known, const int64_t* __restrict__ idx,
                              const uint8_t* __restrict__ cat, uint8_t* __restrict__ labels,
                              int64_t labels_stride2, int64_t origin, int64_t j_first, int64_t n_lines) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_known) return;
    if (idx[k] < 0 || cat[k] == 0) return;
    const int64_t i = idx[k] % job.n1, j = idx[k] / job.n1 - origin;   // band-local line
    if (j < j_first || j >= j_first + n_lines) return;
    const uint8_t* line = job.px + j * job.stride2;
    if (j < 0 && job.halo_lo_px) line = job.halo_lo_px + (j + job.halo_lo) * job.stride2;
    else if (j >= job.n2 && job.halo_hi_px) line = job.halo_hi_px + (j - job.n2) * job.stride2;
    const uint8_t* p = line + i * job.bpp;
    labels[(j - j_first) * labels_stride2 + i] =
        (uint8_t)mr_eval_px(job.pal, p[0], p[1], p[2], job.bpp == 4 ? p[3] : 255, job.bpp == 4, cat[k]);
}

cudaError_t mr_launch_known_fixup(const MrJob& job, int64_t n_known, const int64_t* idx,
                                  const uint8_t* cat, uint8_t* labels, int64_t labels_stride2, int64_t origin,
                                  int64_t j_first, int64_t n_lines, cudaStream_t s, int* n_launches) {
    if (n_known <= 0) return cudaSuccess;
    k_known_fixup<<<(unsigned)((n_known + 127) / 128), 128, 0, s>>>(job, n_known, idx, cat, labels,
                                                                     labels_stride2, origin, j_first, n_lines);
    *n_launches += 1;
    return cudaGetLastError();
}


// largest label of a label image (picks the plane count of the fast standalone clean).
// 16 bytes per thread and step when the lines are 16-byte aligned.
__global__ void k_max_label(const uint8_t* __restrict__ lab, int64_t n1, int64_t n2, int64_t stride2, bool al16,
                            unsigned int* __restrict__ out) {
    unsigned m4 = 0;   // four byte-wise maxima
    const int64_t per_line = (n1 + 15) / 16;
    const int64_t n = per_line * n2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = i / per_line, x = (i - y * per_line) * 16;
        const uint8_t* p = lab + y * stride2 + x;
        if (al16 && x + 16 <= n1) {
            const uint4 v = *reinterpret_cast<const uint4*>(p);
            m4 = __vmaxu4(m4, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
        } else {
            for (int k = 0; k < 16 && x + k < n1; ++k) m4 = __vmaxu4(m4, (unsigned)p[k]);
        }
    }
    unsigned m = __vmaxu4(m4, m4 >> 16);
    m = __vmaxu4(m, m >> 8) & 0xFFu;
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

cudaError_t mr_launch_max_label(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, unsigned int* d_out,
                                int sm_count, cudaStream_t s, int* n_launches) {
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    const bool al16 = ((uintptr_t)labels % 16 == 0) && (stride2 % 16 == 0);
    k_max_label<<<sm_count * 8, 256, 0, s>>>(labels, n1, n2, stride2, al16, d_out);
    *n_launches += 1;
    return cudaGetLastError();
}


// ------------------------------------------------------------------ RGBA{N0f8} front end
// RGBA sheets take the fast kernels through packed RGB: k_rgba_strip drops the alpha byte (and notes
// whether any pixel is not opaque), the label table -- built with alpha = 255 in the distance
// (src/categories.jl:92) -- serves every opaque pixel, and k_alpha_fix re-evaluates the others
// directly (alpha term, alpha == 0 rule of src/categories.jl:115-118) before the majority filter.
__global__ void k_rgba_strip(const uint8_t* __restrict__ px, int64_t stride2, int64_t n1, int64_t lines,
                             uint8_t* __restrict__ rgb, int64_t rgb_stride2, unsigned int* __restrict__ any_special) {
    const int64_t groups = (n1 + 3) / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool special = false;
    if (i < groups * lines) {
        const int64_t y = i / groups, x = (i - y * groups) * 4;
        const uint8_t* p = px + y * stride2 + x * 4;
        uint8_t* q = rgb + y * rgb_stride2 + x * 3;
        if (x + 4 <= n1 && ((uintptr_t)p % 4 == 0)) {
            const uint32_t a = *reinterpret_cast<const uint32_t*>(p), b = *reinterpret_cast<const uint32_t*>(p + 4),
                           c = *reinterpret_cast<const uint32_t*>(p + 8), d = *reinterpret_cast<const uint32_t*>(p + 12);
            special = ((a & b & c & d) >> 24) != 0xFFu;
            uint32_t* o = reinterpret_cast<uint32_t*>(q);   // rgb lines are 16-byte aligned, x * 3 is a multiple of 4
            o[0] = (a & 0x00FFFFFFu) | (b << 24);
            o[1] = ((b >> 8) & 0x0000FFFFu) | (c << 16);
            o[2] = ((c >> 16) & 0x000000FFu) | (d << 8);
        } else {
            for (int k = 0; k < 4 && x + k < n1; ++k) {
                q[3 * k] = p[4 * k]; q[3 * k + 1] = p[4 * k + 1]; q[3 * k + 2] = p[4 * k + 2];
                special = special || p[4 * k + 3] != 0xFFu;
            }
        }
    }
    if (__any_sync(0xffffffffu, special) && (threadIdx.x & 31) == 0) *any_special = 1u;
}

__global__ void k_alpha_fix(MrPalette pal, const uint8_t* __restrict__ px, int64_t stride2, int64_t n1, int64_t lines,
                            uint8_t* __restrict__ labels, int64_t labels_stride2,
                            const unsigned int* __restrict__ any_special) {
    if (*any_special == 0u) return;   // every pixel is opaque: nothing to do
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1 * lines) return;
    const int64_t y = i / n1, x = i - y * n1;
    const uint8_t* p = px + y * stride2 + x * 4;
    if (p[3] != 0xFFu) labels[y * labels_stride2 + x] = (uint8_t)mr_eval_px(pal, p[0], p[1], p[2], p[3], true, 0);
}

cudaError_t mr_launch_rgba_strip(const uint8_t* px, int64_t stride2, int64_t n1, int64_t lines, uint8_t* rgb,
                                 int64_t rgb_stride2, unsigned int* d_any_special, cudaStream_t s, int* n_launches) {
    cudaError_t e = cudaMemsetAsync(d_any_special, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    const int64_t n = (n1 + 3) / 4 * lines;
    k_rgba_strip<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(px, stride2, n1, lines, rgb, rgb_stride2, d_any_special);
    *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t mr_launch_alpha_fix(const MrPalette& pal, const uint8_t* px, int64_t stride2, int64_t n1, int64_t lines,
                                uint8_t* labels, int64_t labels_stride2, const unsigned int* d_any_special,
                                cudaStream_t s, int* n_launches) {
    const int64_t n = n1 * lines;
    k_alpha_fix<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pal, px, stride2, n1, lines, labels, labels_stride2,
                                                            d_any_special);
    *n_launches += 1;
    return cudaGetLastError();
}
