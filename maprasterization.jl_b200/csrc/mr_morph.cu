// mr_morph.cu -- categorical erosion / dilation of a label plane: `_erode` (src/clean.jl:121-129), `_dilate` (:96-118)
// and their loop `_shrink_grow(A, i, j)` (:132-140; the reference's call sites are commented out,
// src/selectcolors.jl:699, :703).  Moore(1) = the 8 neighbours without the centre in CartesianIndices order (d1
// fastest); neighbours outside the image are missingval = 0 (decisions: Neighborhoods.jl 0.1 is not in the reference
// tree, parity unpinned).
//   erode:  all 8 neighbours equal the pixel -> 0, else the pixel (as written in the reference: it empties interiors)
//   dilate: a pixel != 0 stays; a 0 pixel takes the neighbour at the FIRST maximum of "how many neighbours equal
//           neighbour k"; if that is 0, the first neighbour that is not 0 (0 when there is none)
// One thread = four consecutive pixels of a line (one 32-bit store); the 3 x 6 neighbourhood bytes come through L1.
// Bound: HBM, 2 B per pixel and pass.
#include "mr_internal.h"

namespace {

template <bool DILATE>
__global__ void __launch_bounds__(256) k_morph(const uint8_t* __restrict__ a, int64_t stride2, int64_t n1, int64_t n2,
                                               uint8_t* __restrict__ out, int64_t out_stride2) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n1) return;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        // rows j-1, j, j+1, columns i0-1 .. i0+4 (0 outside the image)
        uint8_t w[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int64_t y = j + r - 1;
            const bool oky = y >= 0 && y < n2;
            const uint8_t* line = a + (oky ? y : 0) * stride2;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int64_t x = i0 - 1 + q;
                w[r][q] = (oky && x >= 0 && x < n1) ? __ldg(line + x) : (uint8_t)0;
            }
        }
        uint8_t res[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            // neighbours in CartesianIndices order: (-1,-1) (0,-1) (1,-1) (-1,0) (1,0) (-1,1) (0,1) (1,1)
            const uint8_t h[8] = {w[0][p], w[0][p + 1], w[0][p + 2], w[1][p], w[1][p + 2], w[2][p], w[2][p + 1], w[2][p + 2]};
            const uint8_t x = w[1][p + 1];
            if (!DILATE) {
                bool all = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) all = all && h[k] == x;
                res[p] = all ? (uint8_t)0 : x;
            } else if (x != 0) {
                res[p] = x;
            } else {
                int best = 0, bc = -1;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    int c = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) c += h[q] == h[k];
                    if (c > bc) { bc = c; best = k; }           // findmax: first maximum
                }
                uint8_t v = 0, first = 0;
#pragma unroll
                for (int k = 7; k >= 0; --k) {
                    if (k == best) v = h[k];
                    if (h[k] != 0) first = h[k];                 // ends as the first neighbour that is not 0
                }
                res[p] = v != 0 ? v : first;
            }
        }
        uint8_t* o = out + j * out_stride2 + i0;
        if (i0 + 4 <= n1 && (reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
            *reinterpret_cast<uint32_t*>(o) = (uint32_t)res[0] | ((uint32_t)res[1] << 8) | ((uint32_t)res[2] << 16) | ((uint32_t)res[3] << 24);
        } else {
            for (int p = 0; p < 4 && i0 + p < n1; ++p) o[p] = res[p];
        }
    }
}

}  // namespace

cudaError_t mr_launch_morph(bool dilate, const uint8_t* a, int64_t stride2, int64_t n1, int64_t n2, uint8_t* out,
                            int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const int64_t gx = ((n1 + 3) / 4 + 255) / 256;
    int64_t gy = ((int64_t)sm_count * 8 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    const dim3 g((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy));
    if (dilate) k_morph<true><<<g, 256, 0, s>>>(a, stride2, n1, n2, out, out_stride2);
    else k_morph<false><<<g, 256, 0, s>>>(a, stride2, n1, n2, out, out_stride2);
    *n_launches += 1;
    return cudaGetLastError();
}
