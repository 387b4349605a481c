// mr_blur.cu -- the colour pre-smoothing of the reference and the categoriser on its output:
//   blur(A; hood = Window{R}(), repeat)   src/blur.jl:6-14: every pass replaces a pixel by
//       mean(neighbors(h)) of its Window{R} (centre included), in RGB{Float64};
//   _categorize on RGB{Float64} pixels       src/categories.jl:68-73,76-95 (what `blur` / `_balance` hand it;
//       the reference's own known-answer vectors are Float64 colours).
// The exact-table trick of the N0f8 path does not apply (the colour is no longer one of 2^24 values):
// every pixel is evaluated directly in Float64 with the operations of mr_device.cuh.
//
// Decisions about Neighborhoods.jl 0.1 (not in the reference tree; parity unpinned):
// neighbours in CartesianIndices order (d1 = memory-fast axis fastest), `mean` = left-folded sum of the
// colours times the reciprocal of their number (2R+1)^2 (ColorVectorSpace divides an RGB by a count as
// (1 / count) * colour), neighbours outside the image contribute the zero colour
// (boundary = Remove(zero)), output of the size of the input.  Adding 0.0 is exact, so skipping the
// outside neighbours IS adding them.
// Bound: FP64 issue (9 .. 49 additions per channel and pixel) for the blur, FP64 issue (about 9 K
// operations per pixel) for the categoriser; both stream 24 B per pixel of Float64 colour.
#include "mr_device.cuh"
#include "mr_internal.h"

namespace {

// in: u8 (bpp = 3, N0f8) or f64 (3 doubles per pixel); out: 3 doubles per pixel
template <bool U8IN>
__global__ void __launch_bounds__(256) k_blur(const void* __restrict__ in, int64_t in_stride2, int64_t n1, int64_t n2,
                                              int R, double* __restrict__ out, int64_t out_stride2) {
    const int64_t total = n1 * n2;
    // total / count on an RGB is (one(T) / count) * colour in ColorVectorSpace: reciprocal, then one product per channel
    const double rcnt = __ddiv_rn(1.0, (double)((2 * R + 1) * (2 * R + 1)));
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / n1, i = t - j * n1;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int d2 = -R; d2 <= R; ++d2) {
            const int64_t b = j + d2;
            if (b < 0 || b >= n2) continue;
            const uint8_t* line = reinterpret_cast<const uint8_t*>(in) + b * in_stride2;
            for (int d1 = -R; d1 <= R; ++d1) {
                const int64_t a = i + d1;
                if (a < 0 || a >= n1) continue;
                double x0, x1, x2;
                if (U8IN) {
                    const uint8_t* p = line + a * 3;
                    x0 = __ddiv_rn((double)p[0], 255.0);
                    x1 = __ddiv_rn((double)p[1], 255.0);
                    x2 = __ddiv_rn((double)p[2], 255.0);
                } else {
                    const double* p = reinterpret_cast<const double*>(line) + a * 3;
                    x0 = p[0]; x1 = p[1]; x2 = p[2];
                }
                s0 = __dadd_rn(s0, x0);
                s1 = __dadd_rn(s1, x1);
                s2 = __dadd_rn(s2, x2);
            }
        }
        double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out) + j * out_stride2) + i * 3;
        o[0] = __dmul_rn(rcnt, s0);
        o[1] = __dmul_rn(rcnt, s1);
        o[2] = __dmul_rn(rcnt, s2);
    }
}

// palette means / bounds in shared memory (K <= 254, C = 3): the K-loop reads them for every pixel
__global__ void __launch_bounds__(256) k_categorize_f64(MrPalette pal, const double* __restrict__ px, int64_t stride2,
                                                        int64_t n1, int64_t n2, uint8_t* __restrict__ out,
                                                        int64_t out_stride2, MrCountersDev* counters) {
    const int64_t total = n1 * n2;
    unsigned nomatch = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / n1, i = t - j * n1;
        const double* p = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(px) + j * stride2) + i * 3;
        const int l = mr_eval_f64(pal, p[0], p[1], p[2], 0.0, false, 0);
        out[j * out_stride2 + i] = (uint8_t)l;
        nomatch += l == 0;
    }
    if (counters) {
        for (int o = 16; o > 0; o >>= 1) nomatch += __shfl_xor_sync(0xffffffffu, nomatch, o);
        if ((threadIdx.x & 31) == 0 && nomatch) atomicAdd(&counters->n_nomatch, (unsigned long long)nomatch);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters->n_pixels, (unsigned long long)total);
    }
}

__global__ void k_known_fixup_f64(MrPalette pal, const double* __restrict__ px, int64_t stride2, int64_t n1, int64_t n2,
                                  int64_t n_known, const int64_t* __restrict__ idx, const uint8_t* __restrict__ cat,
                                  uint8_t* __restrict__ labels, int64_t labels_stride2) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_known) return;
    if (idx[k] < 0 || idx[k] >= n1 * n2 || cat[k] == 0) return;
    const int64_t i = idx[k] % n1, j = idx[k] / n1;
    const double* p = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(px) + j * stride2) + i * 3;
    labels[j * labels_stride2 + i] = (uint8_t)mr_eval_f64(pal, p[0], p[1], p[2], 0.0, false, cat[k]);
}

unsigned grid_for(int64_t total, int sm_count) {
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 32;
    if (g > cap) g = cap;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

cudaError_t mr_launch_blur(const void* in, bool in_is_u8, int64_t in_stride2, int64_t n1, int64_t n2, int R, double* out,
                           int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const unsigned g = grid_for(n1 * n2, sm_count);
    if (in_is_u8) k_blur<true><<<g, 256, 0, s>>>(in, in_stride2, n1, n2, R, out, out_stride2);
    else k_blur<false><<<g, 256, 0, s>>>(in, in_stride2, n1, n2, R, out, out_stride2);
    *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t mr_launch_categorize_f64(const MrPalette& pal, const double* px, int64_t stride2, int64_t n1, int64_t n2,
                                     int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat, uint8_t* out,
                                     int64_t out_stride2, MrCountersDev* counters, int sm_count, cudaStream_t s,
                                     int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    k_categorize_f64<<<grid_for(n1 * n2, sm_count), 256, 0, s>>>(pal, px, stride2, n1, n2, out, out_stride2, counters);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    *n_launches += 1;
    if (n_known > 0) {
        k_known_fixup_f64<<<(unsigned)((n_known + 127) / 128), 128, 0, s>>>(pal, px, stride2, n1, n2, n_known, known_idx,
                                                                             known_cat, out, out_stride2);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        *n_launches += 1;
    }
    return cudaSuccess;
}
