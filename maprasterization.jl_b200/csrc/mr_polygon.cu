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

}  // namespace

// dev: off (n_rings + 1 int32) | pad to 8 | xy (2 nv doubles) | ext (2 n_rings doubles) | fill (n_rings bytes), see
// mr_polygon_pack_bytes / the packing in mr_api.cu
cudaError_t mr_launch_polygons(int n_rings, const int32_t* off, const double* xy, const double* ext, const uint8_t* fill,
                               int64_t n1, int64_t n2, uint8_t* plane, int64_t stride2, int sm_count, cudaStream_t s,
                               int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const int64_t gx = (n1 + 255) / 256;
    int64_t gy = ((int64_t)sm_count * 8 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    k_polygons<<<dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy)), 256, 0, s>>>(n_rings, off, xy, ext, fill, n1, n2, plane, stride2);
    *n_launches += 1;
    return cudaGetLastError();
}
