// mr_polygon.cu -- polygon masks and polygon painting of the reference's "clean" / "fill" buttons:
//   mask = _polymask(polygons, A)              src/selectcolors.jl:559-563 (Rasters.boolmask(...; shape = :polygon))
//   masked = (1 .- mask) .* categorized        :480-481      = painting every ring with 0
//   rasterize!(fill_raster, p; fill = i)       :494-496      = painting the rings of category i with i
// Rule (Rasters.jl 0.4 is not in the reference tree: a decision, parity unpinned): pixel
// (i, j) (1-based) is the point (i, j); it belongs to a ring when the ray towards +d1 crosses an odd number of
// edges; edge (a, b) is crossed when (a2 > j) != (b2 > j) and i < a1 + (b1 - a1) * ((j - a2) / (b2 - a2)), one
// rounded Float64 operation per step (the oracle's order); rings with fewer than three vertices are empty.
// One thread = one pixel column position, lines strided over grid.y; rings whose d2 extent misses the line are
// skipped (extents computed by the host).  Bound: the per-pixel edge loop (a GUI polygon has tens of vertices).
#include "mr_internal.h"

namespace {

// ring p: vertices xy[2 * off[p]] .. xy[2 * off[p + 1]); ext[2 p], ext[2 p + 1] = min / max of its d2 coordinates
__global__ void __launch_bounds__(256) k_polygons(int n_rings, const int32_t* __restrict__ off, const double* __restrict__ xy,
                                                  const double* __restrict__ ext, const uint8_t* __restrict__ fill, int64_t n1,
                                                  int64_t n2, uint8_t* __restrict__ plane, int64_t stride2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const double pi = (double)(i + 1);
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        const double pj = (double)(j + 1);
        int hit = -1;
        for (int p = 0; p < n_rings; ++p) {
            const int v0 = __ldg(off + p), nv = __ldg(off + p + 1) - v0;
            if (nv < 3 || pj < __ldg(ext + 2 * p) || pj > __ldg(ext + 2 * p + 1)) continue;   // warp-uniform
            const double* r = xy + 2 * (int64_t)v0;
            bool in = false;
            double b1 = __ldg(r + 2 * (nv - 1)), b2 = __ldg(r + 2 * (nv - 1) + 1);
            for (int a = 0; a < nv; ++a) {
                const double a1 = __ldg(r + 2 * a), a2 = __ldg(r + 2 * a + 1);
                if ((a2 > pj) != (b2 > pj)) {
                    const double t = __ddiv_rn(__dsub_rn(pj, a2), __dsub_rn(b2, a2));
                    const double x = __dadd_rn(a1, __dmul_rn(__dsub_rn(b1, a1), t));
                    if (pi < x) in = !in;
                }
                b1 = a1;
                b2 = a2;
            }
            if (in) hit = p;
        }
        uint8_t* o = plane + j * stride2 + i;
        if (!fill) *o = hit >= 0;
        else if (hit >= 0) *o = __ldg(fill + hit);
    }
}

// The crossings of a line are the same for all of its pixels: per (block, line) the edges are evaluated once
// (one thread per edge; the x of every crossed edge is appended to its ring's segment in shared memory -- the
// order inside a segment does not matter, parity is commutative), then a pixel only compares its d1 coordinate
// with the few crossings of each ring the line crosses.
constexpr int PXL = 8;          // pixels per thread: two groups of four consecutive pixels (one 32-bit store each)
__global__ void __launch_bounds__(256) k_polygons_lines(int n_rings, int n_edges, const int32_t* __restrict__ off,
                                                        const double* __restrict__ xy, const int32_t* __restrict__ ring_of,
                                                        const uint8_t* __restrict__ fill, int64_t n1, int64_t n2,
                                                        uint8_t* __restrict__ plane, int64_t stride2) {
    extern __shared__ int s_x[];                           // [n_edges] crossing thresholds, ring p in [off[p], off[p] + s_cnt[p])
    int* s_cnt = s_x + n_edges;                            // [n_rings]
    int* s_off = s_cnt + n_rings;                          // [n_rings]
    int* s_act = s_off + n_rings;                          // [n_rings] rings with crossings on this line, ascending
    __shared__ int s_nact;
    for (int p = threadIdx.x; p < n_rings; p += blockDim.x) s_off[p] = __ldg(off + p);
    const int64_t i0 = ((int64_t)blockIdx.x * (256 * (PXL / 4)) + threadIdx.x) * 4;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        const double pj = (double)(j + 1);
        for (int p = threadIdx.x; p < n_rings; p += blockDim.x) s_cnt[p] = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
            const int p = __ldg(ring_of + e);
            if (p < 0) continue;                           // ring with fewer than three vertices
            const int v0 = __ldg(off + p), v1 = __ldg(off + p + 1);
            const int eb = e == v0 ? v1 - 1 : e - 1;       // edge (a, b): vertex e and the vertex before it in the ring
            const double a1 = __ldg(xy + 2 * e), a2 = __ldg(xy + 2 * e + 1), b1 = __ldg(xy + 2 * eb), b2 = __ldg(xy + 2 * eb + 1);
            if ((a2 > pj) != (b2 > pj)) {
                const double t = __ddiv_rn(__dsub_rn(pj, a2), __dsub_rn(b2, a2));
                const double x = __dadd_rn(a1, __dmul_rn(__dsub_rn(b1, a1), t));
                // pixel i (0-based) is left of the crossing iff (double)(i + 1) < x  <=>  i < ceil(x) - 1: an integer
                // threshold, clamped to the line (the pixel phase then needs no Float64 at all)
                const double th = ceil(x) - 1.0;
                s_x[v0 + atomicAdd(&s_cnt[p], 1)] = th < 0.0 ? 0 : (th > (double)n1 ? (int)n1 : (int)th);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {                            // the rings this line crosses, in ring order (later rings win)
            int n = 0;
            for (int p = 0; p < n_rings; ++p)
                if (s_cnt[p]) s_act[n++] = p;
            s_nact = n;
        }
        __syncthreads();
        const int nact = s_nact;
        uint8_t* row = plane + j * stride2;
#pragma unroll 1
        for (int g = 0; g < PXL / 4; ++g) {
            const int64_t i = i0 + (int64_t)g * 1024;      // 256 threads x 4 pixels per group
            if (i >= n1) break;
            int hit[4] = {-1, -1, -1, -1};
            for (int a = 0; a < nact; ++a) {
                const int p = s_act[a], c = s_cnt[p];
                const int* x = s_x + s_off[p];
                unsigned in = 0u;
                for (int q = 0; q < c; ++q) {
                    const int left = x[q] - (int)i;             // how many of the four pixels lie left of the crossing
                    in ^= (1u << (left < 0 ? 0 : (left > 4 ? 4 : left))) - 1u;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((in >> k) & 1u) hit[k] = p;
            }
            uint8_t* o = row + i;
            if (i + 4 <= n1 && (reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
                uint32_t w = fill ? *reinterpret_cast<const uint32_t*>(o) : 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!fill) w |= (uint32_t)(hit[k] >= 0) << (8 * k);
                    else if (hit[k] >= 0) w = (w & ~(0xFFu << (8 * k))) | ((uint32_t)__ldg(fill + hit[k]) << (8 * k));
                }
                *reinterpret_cast<uint32_t*>(o) = w;
            } else {
                for (int k = 0; k < 4 && i + k < n1; ++k) {
                    if (!fill) o[k] = hit[k] >= 0;
                    else if (hit[k] >= 0) o[k] = __ldg(fill + hit[k]);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

// all pointers device memory; ring_of[e] = ring of vertex / edge e (-1: its ring has fewer than three vertices)
cudaError_t mr_launch_polygons(int n_rings, int n_edges, const int32_t* off, const double* xy, const double* ext,
                               const int32_t* ring_of, const uint8_t* fill, int64_t n1, int64_t n2, uint8_t* plane,
                               int64_t stride2, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const size_t smem = (size_t)n_edges * 4 + (size_t)n_rings * 12 + 16;
    if (smem <= 40 * 1024) {          // the usual case: a handful of hand-drawn polygons
        const int64_t gx = (n1 + 256 * PXL - 1) / (256 * PXL);
        int64_t gy = ((int64_t)sm_count * 8 + gx - 1) / gx;
        if (gy > n2) gy = n2;
        if (gy > 65535) gy = 65535;
        k_polygons_lines<<<dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy)), 256, smem, s>>>(n_rings, n_edges, off, xy, ring_of, fill,
                                                                                            n1, n2, plane, stride2);
    } else {                          // thousands of vertices: every pixel walks the rings itself
        const int64_t gx = (n1 + 255) / 256;
        int64_t gy = ((int64_t)sm_count * 8 + gx - 1) / gx;
        if (gy > n2) gy = n2;
        if (gy > 65535) gy = 65535;
        k_polygons<<<dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy)), 256, 0, s>>>(n_rings, off, xy, ext, fill, n1, n2, plane, stride2);
    }
    *n_launches += 1;
    return cudaGetLastError();
}
