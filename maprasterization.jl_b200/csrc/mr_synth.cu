// mr_synth.cu -- deterministic synthetic scans, bench/test utility.
// Integer-only (Philox4x32-10 counter RNG + Irwin-Hall noise + jittered-grid Voronoi label
// field) so that a CPU implementation of the same specification yields identical bytes.
// Specification (DESIGN.md "Synthetic scans"):
//   pixel (i, j) of sheet s:  w = philox(ctr = (i, j, 'PIXE', 0), key = (seed, s))
//   mode 1: rgb = low bytes of w0, w1, w2 (uniform random colours)
//   mode 0: label field = nearest jittered site among the 3x3 neighbouring 64x64 cells
//           (site of cell (a, b): philox(ctr = (a, b, 'SITE', 0)) -> x = 64a + (w0 & 63),
//           y = 64b + (w1 & 63), palette index = w2 % K; ties keep the first in (da, db) order);
//           w3 & 0xFFFF < 1311 -> speckle (another palette colour); < 1639 -> stray uniform RGB;
//           channel = clamp(palette + ((sum of the 4 bytes of w_ch - 510) * 333 + 4096 >> 13)).
#include "mr_internal.h"

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t* o) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ int synth_noise(uint32_t w) {
    const int s = (int)(w & 255u) + (int)((w >> 8) & 255u) + (int)((w >> 16) & 255u) + (int)(w >> 24) - 510;
    return ((s * 333 + 4096 + (1 << 20)) >> 13) - 128;
}

__global__ void __launch_bounds__(256) k_synth(uint32_t seed, uint32_t sheet, int K,
                                               const uint8_t* __restrict__ pal, int64_t n1, int64_t line0,
                                               int64_t nlines, int mode, uint8_t* __restrict__ out,
                                               int64_t out_stride2) {
    const int64_t total = n1 * nlines;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t jj = e / n1, i = e - jj * n1, j = line0 + jj;
        uint8_t* o = out + jj * out_stride2 + 3 * i;
        uint32_t w[4];
        philox4x32((uint32_t)i, (uint32_t)j, 0x50495845u, 0u, seed, sheet, w);
        if (mode == 1) {
            o[0] = (uint8_t)(w[0] & 255u); o[1] = (uint8_t)(w[1] & 255u); o[2] = (uint8_t)(w[2] & 255u);
            continue;
        }
        const int64_t ci = i >> 6, cj = j >> 6;
        int64_t bestd = -1;
        int idx = 0;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                uint32_t s[4];
                const int64_t a = ci + di, b = cj + dj;
                philox4x32((uint32_t)a, (uint32_t)b, 0x53495445u, 0u, seed, sheet, s);
                const int64_t sx = a * 64 + (int64_t)(s[0] & 63u), sy = b * 64 + (int64_t)(s[1] & 63u);
                const int64_t d = (i - sx) * (i - sx) + (j - sy) * (j - sy);
                if (bestd < 0 || d < bestd) { bestd = d; idx = (int)(s[2] % (uint32_t)K); }
            }
        const uint32_t sel = w[3] & 0xFFFFu;
        if (K > 1 && sel < 1311u) {
            idx = (int)(((uint32_t)idx + 1u + ((w[3] >> 16) % (uint32_t)(K - 1))) % (uint32_t)K);
        } else if (sel >= 1311u && sel < 1639u) {
            o[0] = (uint8_t)(w[0] & 255u); o[1] = (uint8_t)(w[1] & 255u); o[2] = (uint8_t)(w[2] & 255u);
            continue;
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            int v = (int)pal[idx * 3 + ch] + synth_noise(w[ch]);
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            o[ch] = (uint8_t)v;
        }
    }
}

cudaError_t mr_launch_synth(uint32_t seed, uint32_t sheet, int K, const uint8_t* pal_dev, int64_t n1,
                            int64_t line0, int64_t nlines, int mode, uint8_t* out, int64_t out_stride2,
                            cudaStream_t s, int* n_launches) {
    const int64_t total = n1 * nlines;
    int64_t grid = (total + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    k_synth<<<(unsigned)grid, 256, 0, s>>>(seed, sheet, K, pal_dev, n1, line0, nlines, mode, out, out_stride2);
    *n_launches += 1;
    return cudaGetLastError();
}
