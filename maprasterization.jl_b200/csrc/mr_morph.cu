// mr_morph.cu -- label-plane helpers around the path: `_recolor` (display layers, at the end of the file) and the
// categorical erosion / dilation of a label plane: `_erode` (src/clean.jl:121-129), `_dilate` (:96-118)
// and their loop `_shrink_grow(A, i, j)` (:132-140; the reference's call sites are commented out,
// src/selectcolors.jl:699, :703).  Moore(1) = the 8 neighbours without the centre in CartesianIndices order (d1
// fastest); neighbours outside the image are missingval = 0 (decisions: Neighborhoods.jl 0.1 is not in the reference
// tree, parity unpinned).
//   erode:  all 8 neighbours equal the pixel -> 0, else the pixel (as written in the reference: it empties interiors)
//   dilate: a pixel != 0 stays; a 0 pixel takes the neighbour at the FIRST maximum of "how many neighbours equal
//           neighbour k"; if that is 0, the first neighbour that is not 0 (0 when there is none)
// Bound: HBM, 2 B per pixel and pass.
#include "mr_internal.h"

namespace {

// bytes of a 16-pixel row segment shifted by one pixel: pixel x-1 (left) / x+1 (right) at the position of pixel x;
// lb / rb = the pixels just outside the segment
__device__ __forceinline__ void shift_px(const uint32_t (&w)[4], uint32_t lb, uint32_t rb, uint32_t (&l)[4], uint32_t (&r)[4]) {
    l[0] = (w[0] << 8) | lb;
    l[1] = __funnelshift_l(w[0], w[1], 8);
    l[2] = __funnelshift_l(w[1], w[2], 8);
    l[3] = __funnelshift_l(w[2], w[3], 8);
    r[0] = __funnelshift_r(w[0], w[1], 8);
    r[1] = __funnelshift_r(w[1], w[2], 8);
    r[2] = __funnelshift_r(w[2], w[3], 8);
    r[3] = (w[3] >> 8) | (rb << 24);
}
// 0xFF in every byte of t that is not zero
__device__ __forceinline__ uint32_t nz_bytes(uint32_t t) {
    return (((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) >> 7) & 0x01010101u) * 0xFFu;
}

// One thread = 16 consecutive pixels of a line (16-byte loads of the three rows when the lines are 16-byte aligned,
// the two pixels beside the segment as byte loads).  Erode is byte-parallel: xor of the pixel word with its eight
// shifted neighbour words, the bytes that stay zero are the pixels to clear.  Dilate copies words without a 0 pixel
// and runs the counting rule only for the 0 pixels.
template <bool DILATE>
__global__ void __launch_bounds__(256) k_morph(const uint8_t* __restrict__ a, int64_t stride2, int64_t n1, int64_t n2,
                                               uint8_t* __restrict__ out, int64_t out_stride2, bool al16, bool out_al16) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n1) return;
    const int nv = n1 - i0 < 16 ? (int)(n1 - i0) : 16;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        uint32_t w[3][4], lb[3], rb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int64_t y = j + r - 1;
            w[r][0] = w[r][1] = w[r][2] = w[r][3] = 0u;
            lb[r] = rb[r] = 0u;
            if (y < 0 || y >= n2) continue;                    // outside: 0
            const uint8_t* line = a + y * stride2;
            if (nv == 16 && al16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(line + i0));
                w[r][0] = v.x; w[r][1] = v.y; w[r][2] = v.z; w[r][3] = v.w;
            } else {   // ragged end of the line / unaligned lines.  A real loop: unrolled, the compiler predicated these 16
                       // byte loads into the aligned path as well (a fifth of the kernel's instructions, all switched off)
                unsigned long long lo = 0ull, hi = 0ull;
#pragma unroll 1
                for (int k = 0; k < nv; ++k) {
                    const unsigned long long b = __ldg(line + i0 + k);
                    if (k < 8) lo |= b << (8 * k); else hi |= b << (8 * (k - 8));
                }
                w[r][0] = (uint32_t)lo; w[r][1] = (uint32_t)(lo >> 32); w[r][2] = (uint32_t)hi; w[r][3] = (uint32_t)(hi >> 32);
            }
            if (i0 > 0) lb[r] = __ldg(line + i0 - 1);
            if (i0 + 16 < n1) rb[r] = __ldg(line + i0 + 16);
        }
        uint32_t res[4];
        if (!DILATE) {
            uint32_t l[3][4], rr[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r) shift_px(w[r], lb[r], rb[r], l[r], rr[r]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t m = w[1][q];
                const uint32_t d = (l[0][q] ^ m) | (w[0][q] ^ m) | (rr[0][q] ^ m) | (l[1][q] ^ m) | (rr[1][q] ^ m) | (l[2][q] ^ m) |
                                   (w[2][q] ^ m) | (rr[2][q] ^ m);
                res[q] = m & nz_bytes(d);                      // a pixel whose eight neighbours all equal it becomes 0
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) res[q] = w[1][q];
            uint32_t zero = 0u;                                // bit k: pixel k is 0
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t z = ~nz_bytes(w[1][q]) & 0x01010101u;
                zero |= ((z * 0x10204080u) >> 28) << (4 * q);
            }
            zero &= nv == 16 ? 0xFFFFu : ((1u << nv) - 1u);
            if (zero) {   // a 0 pixel whose eight neighbours are all 0 stays 0: drop those first (byte-parallel)
                uint32_t l[3][4], rr[3][4];
#pragma unroll
                for (int r = 0; r < 3; ++r) shift_px(w[r], lb[r], rb[r], l[r], rr[r]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t any = l[0][q] | w[0][q] | rr[0][q] | l[1][q] | rr[1][q] | l[2][q] | w[2][q] | rr[2][q];
                    const uint32_t nzb = nz_bytes(any) & 0x01010101u;       // pixels with a neighbour that is not 0
                    zero &= ~(0xFu << (4 * q)) | (((nzb * 0x10204080u) >> 28) << (4 * q));
                }
            }
            while (zero) {
                const int k = __ffs(zero) - 1;
                zero &= zero - 1;
                // the byte at segment position p (-1 .. 16) of row r
                auto px = [&](int r, int p) -> uint32_t {
                    if (p < 0) return lb[r];
                    if (p > 15) return rb[r];
                    const uint32_t ww = p < 8 ? (p < 4 ? w[r][0] : w[r][1]) : (p < 12 ? w[r][2] : w[r][3]);
                    return (ww >> (8 * (p & 3))) & 0xFFu;
                };
                // neighbours in CartesianIndices order: (-1,-1) (0,-1) (1,-1) (-1,0) (1,0) (-1,1) (0,1) (1,1)
                const uint32_t h[8] = {px(0, k - 1), px(0, k), px(0, k + 1), px(1, k - 1), px(1, k + 1), px(2, k - 1), px(2, k), px(2, k + 1)};
                uint32_t first = 0;
#pragma unroll
                for (int t = 7; t >= 0; --t)
                    if (h[t] != 0) first = h[t];                 // ends as the first neighbour that is not 0
                bool one = true;                                 // every neighbour is 0 or `first`
#pragma unroll
                for (int t = 0; t < 8; ++t) one = one && (h[t] == 0 || h[t] == first);
                uint32_t val = first;
                // One label besides 0 around the pixel (the usual case at a rim): the rule gives that label whatever the
                // counts are -- if it is the first maximum it is taken, if 0 is, the first neighbour that is not 0 is.
                if (!one) {
                    int best = 0, bc = -1;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        int c = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) c += h[q] == h[t];
                        if (c > bc) { bc = c; best = t; }       // findmax: first maximum
                    }
                    uint32_t v = 0;
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (t == best) v = h[t];
                    if (v != 0) val = v;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((k >> 2) == q) res[q] |= val << (8 * (k & 3));
            }
        }
        uint8_t* o = out + j * out_stride2 + i0;
        if (nv == 16 && out_al16) {
            *reinterpret_cast<uint4*>(o) = make_uint4(res[0], res[1], res[2], res[3]);
        } else {
            for (int k = 0; k < nv; ++k) o[k] = (uint8_t)(res[k >> 2] >> (8 * (k & 3)));
        }
    }
}

// `_recolor` (src/selectcolors.jl:650-655): label -> RGBA{Float64} display pixel (the category's mean colour, alpha 1;
// label 0 or a label above K: the transparent pixel).  One thread = one pixel, two 16-byte stores; the colour table
// in shared memory.  Bound: HBM, 33 B per pixel (32 of them written).
__global__ void __launch_bounds__(256) k_recolor(const uint8_t* __restrict__ lab, int64_t stride2, int64_t n1, int64_t n2, int K,
                                                 const double* __restrict__ colors, double* __restrict__ out, int64_t out_stride2) {
    __shared__ double s_c[255 * 4];
    for (int t = threadIdx.x; t < 4 * (K + 1); t += blockDim.x) {
        const int c = t >> 2, ch = t & 3;
        s_c[t] = c == 0 ? 0.0 : (ch == 3 ? 1.0 : __ldg(colors + 3 * (c - 1) + ch));
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        int c = __ldg(lab + j * stride2 + i);
        if (c > K) c = 0;
        double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out) + j * out_stride2) + 4 * i;
        if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
            reinterpret_cast<double2*>(o)[0] = make_double2(s_c[4 * c], s_c[4 * c + 1]);
            reinterpret_cast<double2*>(o)[1] = make_double2(s_c[4 * c + 2], s_c[4 * c + 3]);
        } else {
            o[0] = s_c[4 * c]; o[1] = s_c[4 * c + 1]; o[2] = s_c[4 * c + 2]; o[3] = s_c[4 * c + 3];
        }
    }
}

}  // namespace

cudaError_t mr_launch_recolor(const uint8_t* lab, int64_t stride2, int64_t n1, int64_t n2, int K, const double* colors_dev,
                              double* out, int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const int64_t gx = (n1 + 255) / 256;
    int64_t gy = ((int64_t)sm_count * 8 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    k_recolor<<<dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy)), 256, 0, s>>>(lab, stride2, n1, n2, K, colors_dev, out, out_stride2);
    *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t mr_launch_morph(bool dilate, const uint8_t* a, int64_t stride2, int64_t n1, int64_t n2, uint8_t* out,
                            int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const int64_t gx = ((n1 + 15) / 16 + 255) / 256;
    int64_t gy = ((int64_t)sm_count * 8 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    const dim3 g((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy));
    const bool al = ((uintptr_t)a % 16 == 0) && (stride2 % 16 == 0), oal = ((uintptr_t)out % 16 == 0) && (out_stride2 % 16 == 0);
    if (dilate) k_morph<true><<<g, 256, 0, s>>>(a, stride2, n1, n2, out, out_stride2, al, oal);
    else k_morph<false><<<g, 256, 0, s>>>(a, stride2, n1, n2, out, out_stride2, al, oal);
    *n_launches += 1;
    return cudaGetLastError();
}
