// mr_blur.cu -- the colour pre-smoothing of the reference and the categoriser on its output:
//   blur(A; hood = Window{R}(), repeat)   src/blur.jl:6-14: every pass replaces a pixel by
//       mean(neighbors(h)) of its Window{R} (centre included), in RGB{Float64};
// (The categoriser on RGB{Float64} pixels, src/categories.jl:68-73,76-95, is the PixelProp categoriser of
//  mr_texture.cu with match = color.)
// The exact-table trick of the N0f8 path does not apply (the colour is no longer one of 2^24 values):
// every pixel is evaluated directly in Float64 with the operations of mr_device.cuh.
//
// Decisions about Neighborhoods.jl 0.1 (not in the reference tree; parity unpinned):
// neighbours in CartesianIndices order (d1 = memory-fast axis fastest), `mean` = left-folded sum of the
// colours times the reciprocal of their number (2R+1)^2 (ColorVectorSpace divides an RGB by a count as
// (1 / count) * colour), neighbours outside the image contribute the zero colour
// (boundary = Remove(zero)), output of the size of the input.  Adding 0.0 is exact, so skipping the
// outside neighbours IS adding them.
// Bound: FP64 issue (9 .. 49 additions per channel and pixel); streams 24 B per pixel of Float64 colour.
#include "mr_device.cuh"
#include "mr_internal.h"

namespace {

// in: u8 (bpp = 3, N0f8) or f64 (3 doubles per pixel); out: 3 doubles per pixel.
// One thread = PXT consecutive pixels of a line: the (PXT + 2R) x (2R + 1) neighbourhood colours are fetched
// once per line into registers (N0f8 -> Float64 through a 256-entry table in shared memory: i / 255 correctly
// rounded, tabulated once per block), the sums of the PXT pixels then run over registers in the reference's
// order.  grid.x covers a line, grid.y strides over the lines (no 64-bit division per pixel).
constexpr int PXT = 4;
template <bool U8IN, int R>
__global__ void __launch_bounds__(128) k_blur(const void* __restrict__ in, int64_t in_stride2, int64_t n1, int64_t n2,
                                              double* __restrict__ out, int64_t out_stride2) {
    __shared__ double s_n0f8[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) s_n0f8[t] = __ddiv_rn((double)t, 255.0);
    __syncthreads();
    constexpr int W = PXT + 2 * R;
    // total / count on an RGB is (one(T) / count) * colour in ColorVectorSpace: reciprocal, then one product per channel
    const double rcnt = __ddiv_rn(1.0, (double)((2 * R + 1) * (2 * R + 1)));
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PXT;
    if (i0 >= n1) return;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        double s[PXT][3];
#pragma unroll
        for (int p = 0; p < PXT; ++p) s[p][0] = s[p][1] = s[p][2] = 0.0;
#pragma unroll
        for (int d2 = -R; d2 <= R; ++d2) {
            const int64_t b = j + d2;
            if (b < 0 || b >= n2) continue;                    // outside: the zero colour (adding 0.0 is exact)
            const uint8_t* line = reinterpret_cast<const uint8_t*>(in) + b * in_stride2;
            double x[W][3];
#pragma unroll
            for (int q = 0; q < W; ++q) {
                const int64_t a = i0 - R + q;
                const bool ok = a >= 0 && a < n1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double v = 0.0;
                    if (ok) v = U8IN ? s_n0f8[__ldg(line + a * 3 + c)] : __ldg(reinterpret_cast<const double*>(line) + a * 3 + c);
                    x[q][c] = v;
                }
            }
#pragma unroll
            for (int p = 0; p < PXT; ++p)
#pragma unroll
                for (int d1 = 0; d1 <= 2 * R; ++d1)            // d1 ascending = neighbourhood order within the line
#pragma unroll
                    for (int c = 0; c < 3; ++c) s[p][c] = __dadd_rn(s[p][c], x[p + d1][c]);
        }
        double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out) + j * out_stride2) + i0 * 3;
        double r[PXT * 3];
#pragma unroll
        for (int p = 0; p < PXT; ++p)
#pragma unroll
            for (int c = 0; c < 3; ++c) r[3 * p + c] = __dmul_rn(rcnt, s[p][c]);
        if (i0 + PXT <= n1 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {   // 96 contiguous bytes: six 16-byte stores
#pragma unroll
            for (int v = 0; v < PXT * 3 / 2; ++v) reinterpret_cast<double2*>(o)[v] = make_double2(r[2 * v], r[2 * v + 1]);
        } else {
#pragma unroll
            for (int p = 0; p < PXT; ++p)
                if (i0 + p < n1) {
                    o[3 * p + 0] = r[3 * p + 0];
                    o[3 * p + 1] = r[3 * p + 1];
                    o[3 * p + 2] = r[3 * p + 2];
                }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// `_balance` (src/balance.jl:25-42), per-pixel part: predictions of the fitted cubic trend, their means,
// out = clamp(x - p + mean(p), 0, 1).  The sum of the predictions runs in a FIXED order that the oracle
// repeats: chunks of 256 pixels of a line sequentially, then the chunk sums of a line, then the lines.
struct BalanceCoef { double b[3][7]; };

__device__ __forceinline__ double balance_pred(const double* b, int64_t i1, int64_t j1) {
    const double fi = (double)i1, fj = (double)j1;
    const double i2 = __dmul_rn(fi, fi), i3 = __dmul_rn(i2, fi), j2 = __dmul_rn(fj, fj), j3 = __dmul_rn(j2, fj);
    double p = __dadd_rn(b[0], __dmul_rn(b[1], i3));
    p = __dadd_rn(p, __dmul_rn(b[2], i2));
    p = __dadd_rn(p, __dmul_rn(b[3], fi));
    p = __dadd_rn(p, __dmul_rn(b[4], j3));
    p = __dadd_rn(p, __dmul_rn(b[5], j2));
    p = __dadd_rn(p, __dmul_rn(b[6], fj));
    return p;
}

// one thread = one chunk of 256 pixels of one line: partial[(j * chunks + c) * 3 + ch]
__global__ void k_balance_chunks(BalanceCoef K, int64_t n1, int64_t n2, int64_t chunks, double* __restrict__ partial) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * n2) return;
    const int64_t j = t / chunks, c = t - j * chunks;
    const int64_t i0 = c * 256, i1 = min(i0 + 256, n1);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t i = i0; i < i1; ++i) {
        s0 = __dadd_rn(s0, balance_pred(K.b[0], i + 1, j + 1));
        s1 = __dadd_rn(s1, balance_pred(K.b[1], i + 1, j + 1));
        s2 = __dadd_rn(s2, balance_pred(K.b[2], i + 1, j + 1));
    }
    partial[t * 3 + 0] = s0;
    partial[t * 3 + 1] = s1;
    partial[t * 3 + 2] = s2;
}
// one thread = one line: the chunk sums of the line, in order
__global__ void k_balance_lines(int64_t n2, int64_t chunks, const double* __restrict__ partial, double* __restrict__ lines) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t c = 0; c < chunks; ++c) {
        const double* q = partial + (j * chunks + c) * 3;
        s0 = __dadd_rn(s0, q[0]);
        s1 = __dadd_rn(s1, q[1]);
        s2 = __dadd_rn(s2, q[2]);
    }
    lines[j * 3 + 0] = s0; lines[j * 3 + 1] = s1; lines[j * 3 + 2] = s2;
}
// three threads add the line sums in order (the block stages them in shared memory, 256 lines at a time, so that
// the serial chain is additions only), divided by the pixel count
__global__ void __launch_bounds__(256) k_balance_total(int64_t n1, int64_t n2, const double* __restrict__ lines, double* __restrict__ means) {
    __shared__ double s_v[768];
    double s = 0.0;
    for (int64_t j0 = 0; j0 < n2; j0 += 256) {
        const int n = (int)min((int64_t)256, n2 - j0);
        for (int k = threadIdx.x; k < 3 * n; k += blockDim.x) s_v[k] = lines[j0 * 3 + k];
        __syncthreads();
        if (threadIdx.x < 3)
            for (int k = 0; k < n; ++k) s = __dadd_rn(s, s_v[3 * k + threadIdx.x]);
        __syncthreads();
    }
    if (threadIdx.x < 3) means[threadIdx.x] = __ddiv_rn(s, (double)(n1 * n2));
}
__global__ void __launch_bounds__(256) k_balance_apply(BalanceCoef K, const uint8_t* __restrict__ px, int64_t stride2,
                                                       int64_t n1, int64_t n2, const double* __restrict__ means,
                                                       double* __restrict__ out, int64_t out_stride2) {
    __shared__ double s_n0f8[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) s_n0f8[t] = __ddiv_rn((double)t, 255.0);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const double m0 = means[0], m1 = means[1], m2 = means[2];
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        const uint8_t* p = px + j * stride2 + i * 3;
        double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out) + j * out_stride2) + i * 3;
        const double mm[3] = {m0, m1, m2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = __dsub_rn(s_n0f8[p[c]], balance_pred(K.b[c], i + 1, j + 1));
            v = __dadd_rn(v, mm[c]);
            v = v > 0.0 ? v : 0.0;
            o[c] = v < 1.0 ? v : 1.0;
        }
    }
}

static dim3 grid_lines(int64_t n1, int64_t n2, int px_per_thread, int threads, int sm_count) {
    const int64_t gx = ((n1 + px_per_thread - 1) / px_per_thread + threads - 1) / threads;
    int64_t gy = ((int64_t)sm_count * 32 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy);
}

template <bool U8IN>
static void launch_blur_r(const void* in, int64_t in_stride2, int64_t n1, int64_t n2, int R, double* out,
                          int64_t out_stride2, dim3 g, cudaStream_t s) {
    switch (R) {
        case 0: k_blur<U8IN, 0><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        case 1: k_blur<U8IN, 1><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        case 2: k_blur<U8IN, 2><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        case 3: k_blur<U8IN, 3><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        case 4: k_blur<U8IN, 4><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        case 5: k_blur<U8IN, 5><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        case 6: k_blur<U8IN, 6><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
        default: k_blur<U8IN, 7><<<g, 128, 0, s>>>(in, in_stride2, n1, n2, out, out_stride2); break;
    }
}

cudaError_t mr_launch_blur(const void* in, bool in_is_u8, int64_t in_stride2, int64_t n1, int64_t n2, int R, double* out,
                           int64_t out_stride2, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const dim3 g = grid_lines(n1, n2, PXT, 128, sm_count);
    if (in_is_u8) launch_blur_r<true>(in, in_stride2, n1, n2, R, out, out_stride2, g, s);
    else launch_blur_r<false>(in, in_stride2, n1, n2, R, out, out_stride2, g, s);
    *n_launches += 1;
    return cudaGetLastError();
}

// scratch: (ceil(n1 / 256) * n2 + n2 + 1) * 3 doubles
size_t mr_balance_scratch_bytes(int64_t n1, int64_t n2) {
    const int64_t chunks = (n1 + 255) / 256;
    return (size_t)(chunks * n2 + n2 + 1) * 3 * sizeof(double);
}

cudaError_t mr_launch_balance(const double* coef21, const uint8_t* px, int64_t stride2, int64_t n1, int64_t n2,
                              double* scratch, double* out, int64_t out_stride2, int sm_count, cudaStream_t s,
                              int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    BalanceCoef K;
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 7; ++k) K.b[c][k] = coef21[c * 7 + k];
    const int64_t chunks = (n1 + 255) / 256;
    double* partial = scratch;
    double* lines = scratch + chunks * n2 * 3;
    double* means = lines + n2 * 3;
    k_balance_chunks<<<(unsigned)((chunks * n2 + 127) / 128), 128, 0, s>>>(K, n1, n2, chunks, partial);
    k_balance_lines<<<(unsigned)((n2 + 127) / 128), 128, 0, s>>>(n2, chunks, partial, lines);
    k_balance_total<<<1, 256, 0, s>>>(n1, n2, lines, means);
    const int64_t gx = (n1 + 255) / 256;
    int64_t gy = ((int64_t)sm_count * 16 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    k_balance_apply<<<dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy)), 256, 0, s>>>(K, px, stride2, n1, n2, means, out, out_stride2);
    *n_launches += 4;
    return cudaGetLastError();
}
