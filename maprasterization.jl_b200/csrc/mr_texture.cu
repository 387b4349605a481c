// mr_texture.cu -- the texture features behind the reference's default match = (:color, :std, :stripe)
//  and the categoriser on PixelProp pixels:
//   _stds(A)                 src/selectcolors.jl:706-715   5 x 5 per-channel standard deviations, summed, / maximum
//   _stripes(img; radius)    src/stripes.jl:1-35           directional HSL contrasts, / mean absolute value
//   _categorize(x::PixelProp, categories, tolerances, match)   src/categories.jl:76-133
// All of it on the blurred RGB{Float64} image (src/selectcolors.jl:455-465), in Float64, every operation
// one rounded __d*_rn operation in the order the oracle states (oracle/maprast_oracle.c, DESIGN.md
// section 14: the decisions about Neighborhoods.jl / Statistics / Colors.jl, parity unpinned).
// Bound: FP64 issue (stds: about 100 operations per channel and pixel); the planes stream 8 / 16 B per pixel.
#include <math_constants.h>

#include "mr_internal.h"

namespace {

__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }

// ---- _stds: one thread = PXT consecutive pixels of a line, one channel at a time: the (PXT + 4) x 5
// neighbourhood values of the channel sit in registers, sums in CartesianIndices order (d1 fastest).
// (ncu: L1 throughput 80 %, FP64 pipe 48 % -- the 8-byte loads of one channel use a quarter of every sector.  Putting
// the three channels on neighbouring lanes for denser accesses was SLOWER, 3.8 -> 4.7 ms: three times the threads
// repeat the address and bounds arithmetic.  Eight pixels per thread (229 registers): 5.7 ms; capped at 128 registers
// for four blocks per SM (spills): 4.2 ms.)
constexpr int SPX = 4;
__global__ void __launch_bounds__(128) k_stds_raw(const double* __restrict__ px, int64_t stride2, int64_t n1, int64_t n2,
                                                  double* __restrict__ out, int64_t out_stride2,
                                                  unsigned long long* __restrict__ d_max) {
    constexpr int W = SPX + 4;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SPX;
    double vmax = 0.0;
    if (i0 < n1) {
        for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
            double tot[SPX];
#pragma unroll
            for (int p = 0; p < SPX; ++p) tot[p] = 0.0;
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                double x[5][W];
#pragma unroll
                for (int d2 = 0; d2 < 5; ++d2) {
                    const int64_t b = j + d2 - 2;
                    const bool okb = b >= 0 && b < n2;
                    const double* line = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(px) + (okb ? b : 0) * stride2);
#pragma unroll
                    for (int q = 0; q < W; ++q) {
                        const int64_t a = i0 - 2 + q;
                        x[d2][q] = (okb && a >= 0 && a < n1) ? __ldg(line + a * 3 + c) : 0.0;   // outside: the zero colour
                    }
                }
#pragma unroll
                for (int p = 0; p < SPX; ++p) {
                    double s = 0.0;
#pragma unroll
                    for (int d2 = 0; d2 < 5; ++d2)
#pragma unroll
                        for (int d1 = 0; d1 < 5; ++d1) s = __dadd_rn(s, x[d2][p + d1]);
                    const double m = __ddiv_rn(s, 25.0);
                    double ss = 0.0;
#pragma unroll
                    for (int d2 = 0; d2 < 5; ++d2)
#pragma unroll
                        for (int d1 = 0; d1 < 5; ++d1) {
                            const double d = __dsub_rn(x[d2][p + d1], m);
                            ss = __dadd_rn(ss, __dmul_rn(d, d));
                        }
                    const double sd = __dsqrt_rn(__ddiv_rn(ss, 24.0));
                    tot[p] = c == 0 ? sd : __dadd_rn(tot[p], sd);     // (rs + gs) + bs
                }
            }
            double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out) + j * out_stride2) + i0;
#pragma unroll
            for (int p = 0; p < SPX; ++p)
                if (i0 + p < n1) {
                    o[p] = tot[p];
                    vmax = tot[p] > vmax ? tot[p] : vmax;
                }
        }
    }
    // non-negative doubles order like their bit patterns
    unsigned long long b = dbits(vmax);
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = t > b ? t : b;
    }
    if ((threadIdx.x & 31) == 0 && b) atomicMax(d_max, b);
}

// plane[j][i] /= *scalar for `width` doubles per line
__global__ void __launch_bounds__(256) k_div_plane(double* __restrict__ plane, int64_t stride2, int64_t width, int64_t n2,
                                                   const double* __restrict__ scalar) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    const double m = *scalar;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(plane) + j * stride2) + i;
        *o = __ddiv_rn(*o, m);
    }
}

// ---- HSL saturation and lightness of an RGB{Float64} colour (the classical conversion; only s and l are used)
__device__ __forceinline__ void hsl_sl(double r, double g, double b, double& s, double& l) {
    double mx = r > g ? r : g, mn = r < g ? r : g;   // the oracle's comparisons (signed zeros, NaN: same picks)
    mx = mx > b ? mx : b;
    mn = mn < b ? mn : b;
    l = __ddiv_rn(__dadd_rn(mx, mn), 2.0);
    if (mx == mn) s = 0.0;
    else if (l < 0.5) s = __ddiv_rn(__dsub_rn(mx, mn), __dadd_rn(mx, mn));
    else s = __ddiv_rn(__dsub_rn(mx, mn), __dsub_rn(__dsub_rn(2.0, mx), mn));
}

__global__ void __launch_bounds__(256) k_sl_plane(const double* __restrict__ px, int64_t stride2, int64_t n1, int64_t n2,
                                                  double2* __restrict__ sl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        const double* p = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(px) + j * stride2) + i * 3;
        double s, l;
        hsl_sl(__ldg(p), __ldg(p + 1), __ldg(p + 2), s, l);
        sl[j * n1 + i] = make_double2(s, l);
    }
}

// ---- _stripes, un-normalised: one thread per pixel over the (s, l) plane; outside = the zero colour (s = l = 0).
// RT > 0: the radius as a template parameter (unrolled); INNER: every neighbour of the line's pixels this thread
// can reach lies inside the image (no bounds tests, plain pointer offsets).
template <int RT, bool INNER>
__device__ __forceinline__ void stripes_px(const double2* __restrict__ sl, int64_t n1, int64_t n2, int radius, int64_t i, int64_t j,
                                           bool pow2, double rinv, double* __restrict__ o) {
    const int rad = RT > 0 ? RT : radius;
    const double2* ctr = sl + j * n1 + i;
    const double2 c = __ldg(ctr);
    double dir[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // vert (1, 0), horz (0, 1), angle45 (1, 1), angle135 (1, -1)
        const int di = k == 1 ? 0 : 1, dj = k == 0 ? 0 : (k == 3 ? -1 : 1);
        const int64_t step = (int64_t)dj * n1 + di;          // element offset of one step along the direction
        double side[2];
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < rad; ++t) {
                const int x = sd == 0 ? -rad + t : 1 + t;
                double2 n = make_double2(0.0, 0.0);
                if (INNER) n = __ldg(ctr + x * step);
                else {
                    const int64_t a = i + di * x, b = j + dj * x;
                    if (a >= 0 && a < n1 && b >= 0 && b < n2) n = __ldg(ctr + x * step);
                }
                const double ds = __dsub_rn(c.x, n.x), dl = __dsub_rn(c.y, n.y);
                acc = __dadd_rn(acc, __dadd_rn(__dmul_rn(ds, ds), __dmul_rn(dl, dl)));
            }
            side[sd] = pow2 ? __dmul_rn(acc, rinv) : __ddiv_rn(acc, (double)rad);   // a power of two: the product is the quotient
        }
        dir[k] = side[1] < side[0] ? side[1] : side[0];
    }
    o[0] = __dsub_rn(dir[0], dir[1]);
    o[1] = __dsub_rn(dir[2], dir[3]);
}

template <int RT>
__global__ void __launch_bounds__(256) k_stripes_raw(const double2* __restrict__ sl, int64_t n1, int64_t n2, int radius,
                                                     double* __restrict__ out, int64_t out_stride2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int rad = RT > 0 ? RT : radius;
    const bool pow2 = (rad & (rad - 1)) == 0;
    const double rinv = 1.0 / (double)rad;
    const bool inner_i = i >= rad && i + rad < n1;
    for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
        double* o = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out) + j * out_stride2) + 2 * i;
        if (inner_i && j >= rad && j + rad < n2) stripes_px<RT, true>(sl, n1, n2, radius, i, j, pow2, rinv, o);
        else stripes_px<RT, false>(sl, n1, n2, radius, i, j, pow2, rinv, o);
    }
}

// ---- mean(abs o first), mean(abs o last) in the fixed order of _balance's mean: chunks of 256 pixels of a
// line sequentially, then the chunk sums of a line, then the lines.
// One warp = one chunk: the 256 values of a component are fetched coalesced into shared memory, lane 0 / 1 add
// them in order.
__global__ void __launch_bounds__(256) k_abs_chunks(const double* __restrict__ raw, int64_t stride2, int64_t n1, int64_t n2,
                                                    int64_t chunks, double* __restrict__ partial) {
    __shared__ double s_v[8][512];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * 8 + w;
    if (t >= chunks * n2) return;
    const int64_t j = t / chunks, c = t - j * chunks;
    const int64_t i0 = c * 256;
    const int n = (int)min((int64_t)256, n1 - i0);
    const double* line = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(raw) + j * stride2) + 2 * i0;
    for (int k = lane; k < 2 * n; k += 32) s_v[w][k] = fabs(__ldg(line + k));
    __syncwarp();
    if (lane < 2) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s = __dadd_rn(s, s_v[w][2 * k + lane]);
        partial[t * 2 + lane] = s;
    }
}
__global__ void k_abs_lines(int64_t n2, int64_t chunks, const double* __restrict__ partial, double* __restrict__ lines) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    double s0 = 0.0, s1 = 0.0;
    for (int64_t c = 0; c < chunks; ++c) {
        s0 = __dadd_rn(s0, partial[(j * chunks + c) * 2]);
        s1 = __dadd_rn(s1, partial[(j * chunks + c) * 2 + 1]);
    }
    lines[j * 2] = s0;
    lines[j * 2 + 1] = s1;
}
// m = mean((mean(abs o first), mean(abs o last))) = (M0 + M1) / 2.  The line sums are added in order by one
// thread per component; the block fetches them 512 lines at a time into shared memory (coalesced) so that the
// serial chain is additions only, not dependent global loads.
__global__ void __launch_bounds__(256) k_abs_total(int64_t n1, int64_t n2, const double* __restrict__ lines, double* __restrict__ m_out) {
    __shared__ double s_v[1024];
    __shared__ double s_m[2];
    double s = 0.0;
    for (int64_t j0 = 0; j0 < n2; j0 += 512) {
        const int n = (int)min((int64_t)512, n2 - j0);
        for (int k = threadIdx.x; k < 2 * n; k += blockDim.x) s_v[k] = lines[j0 * 2 + k];
        __syncthreads();
        if (threadIdx.x < 2)
            for (int k = 0; k < n; ++k) s = __dadd_rn(s, s_v[2 * k + threadIdx.x]);
        __syncthreads();
    }
    if (threadIdx.x < 2) s_m[threadIdx.x] = __ddiv_rn(s, (double)(n1 * n2));
    __syncthreads();
    if (threadIdx.x == 0) *m_out = __ddiv_rn(__dadd_rn(s_m[0], s_m[1]), 2.0);
}

// ---- the categoriser on PixelProp pixels.  Statistics in shared memory:
// colour mean | lo | hi (3 K each), std mean | lo | hi (K each), stripe mean | lo | hi (2 K each).
struct PropPx {
    double r, g, b, sd, s0, s1;
};
// undivided error sum of one pixel against category c: the matched terms in the order color, std, stripe
template <int match>
__device__ __forceinline__ double props_sum(const double* sp, int K, int c, const PropPx& x) {
    const double* cm = sp + 3 * c;
    if (cm[0] == CUDART_INF) return CUDART_INF;       // `missing` category: typemax(Float64)
    double e = 0.0;
    if (match & 1) {
        const double d0 = __dsub_rn(x.r, cm[0]), d1 = __dsub_rn(x.g, cm[1]), d2 = __dsub_rn(x.b, cm[2]);
        e = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
    }
    if (match & 2) {
        const double d = __dsub_rn(x.sd, sp[9 * K + c]);
        const double t = __dmul_rn(d, d);
        e = (match & 1) ? __dadd_rn(e, t) : t;
    }
    if (match & 4) {
        const double d0 = __dsub_rn(x.s0, sp[12 * K + 2 * c]), d1 = __dsub_rn(x.s1, sp[12 * K + 2 * c + 1]);
        const double t = __dmul_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), 0.5);   // / 2: exact either way
        e = (match & 3) ? __dadd_rn(e, t) : t;
    }
    return e;
}

// P pixels at once: the statistics of a category (shared memory, broadcast reads) are fetched once for all of them.
// match (compile time): bit 0 color, bit 1 std, bit 2 stripe.
//
// The error of a category is sum / count (count = number of matched properties) and the reference takes the
// FIRST minimum of the rounded quotients.  The quotient is monotone in the sum, so the winner is the first
// minimum of the undivided sums -- unless an EARLIER category has a sum one ulp above the minimum whose quotient
// rounds to the same double (dividing by 3 can merge adjacent doubles, never doubles two apart; dividing by 2
// is exact).  The loop therefore runs without the (slow, FP64) division and remembers that one possible rival;
// two divisions at the end settle it.  Sums so small that their quotients are subnormal (where spacing
// arguments fail) take a plain loop with one division per category.
template <int P, int match>
__device__ __forceinline__ void eval_props(const double* sp, int K, const PropPx (&x)[P], const int (&known)[P],
                                           int (&lab)[P]) {
    constexpr int CNT = (match & 1) + ((match >> 1) & 1) + ((match >> 2) & 1);
    int c1[P], c2[P];            // first minimum of the sums; earlier category one ulp above it (-1: none)
    double s1[P], s2[P];
#pragma unroll
    for (int q = 0; q < P; ++q) { c1[q] = 0; c2[q] = -1; s1[q] = s2[q] = 0.0; }
    for (int c = 0; c < K; ++c) {
        // the category's statistics once for the P pixels
        const double m0 = sp[3 * c], m1 = sp[3 * c + 1], m2 = sp[3 * c + 2];
        const double ms = (match & 2) ? sp[9 * K + c] : 0.0;
        const double t0 = (match & 4) ? sp[12 * K + 2 * c] : 0.0, t1 = (match & 4) ? sp[12 * K + 2 * c + 1] : 0.0;
        const bool missing = m0 == CUDART_INF;          // `missing` category: typemax(Float64)
#pragma unroll
        for (int q = 0; q < P; ++q) {
            double e = 0.0;
            if (match & 1) {
                const double d0 = __dsub_rn(x[q].r, m0), d1 = __dsub_rn(x[q].g, m1), d2 = __dsub_rn(x[q].b, m2);
                e = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            }
            if (match & 2) {
                const double d = __dsub_rn(x[q].sd, ms);
                const double t = __dmul_rn(d, d);
                e = (match & 1) ? __dadd_rn(e, t) : t;
            }
            if (match & 4) {
                const double d0 = __dsub_rn(x[q].s0, t0), d1 = __dsub_rn(x[q].s1, t1);
                const double t = __dmul_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), 0.5);   // / 2: exact either way
                e = (match & 3) ? __dadd_rn(e, t) : t;
            }
            // colour alone: (x - Inf)^2 = Inf by itself; with texture terms their statistics may be NaN -> explicit
            if ((match & 6) && missing) e = CUDART_INF;
            if (c == 0) s1[q] = e;
            else if (e < s1[q]) {
                if (CNT > 1) {
                    const double up = __longlong_as_double(__double_as_longlong(e) + 1);   // next double above e >= 0
                    if (s1[q] <= up) { s2[q] = s1[q]; c2[q] = c1[q]; } else c2[q] = -1;
                }
                s1[q] = e;
                c1[q] = c;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
        int bq = c1[q];
        if (CNT > 1) {
            if (s1[q] < 1e-290) {          // quotients near the subnormal range: the literal loop
                double beste = 0.0;
                for (int c = 0; c < K; ++c) {
                    const double e = __ddiv_rn(props_sum<match>(sp, K, c, x[q]), (double)CNT);
                    if (c == 0 || e < beste) { beste = e; bq = c; }
                }
            } else if (c2[q] >= 0 && __ddiv_rn(s2[q], (double)CNT) == __ddiv_rn(s1[q], (double)CNT)) {
                bq = c2[q];
            }
        }
        if (known[q] != 0) { lab[q] = bq + 1 == known[q] ? bq + 1 : 0; continue; }
        bool ok = true;
        if (match & 1) {
            const double* l = sp + 3 * K + 3 * bq;
            const double* h = sp + 6 * K + 3 * bq;
            ok = ok && x[q].r >= l[0] && x[q].r <= h[0] && x[q].g >= l[1] && x[q].g <= h[1] && x[q].b >= l[2] && x[q].b <= h[2];
        }
        if (match & 2) ok = ok && x[q].sd >= sp[10 * K + bq] && x[q].sd <= sp[11 * K + bq];
        if (match & 4) {
            const double* l = sp + 14 * K + 2 * bq;
            const double* h = sp + 16 * K + 2 * bq;
            ok = ok && x[q].s0 >= l[0] && x[q].s0 <= h[0] && x[q].s1 >= l[1] && x[q].s1 <= h[1];
        }
        lab[q] = ok ? bq + 1 : 0;
    }
}

__device__ __forceinline__ PropPx load_props(const double* px, int64_t stride2, const double* sd, int64_t sd_stride2,
                                             const double* st, int64_t st_stride2, int64_t i, int64_t j) {
    PropPx x;
    const double* p = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(px) + j * stride2) + i * 3;
    x.r = __ldg(p); x.g = __ldg(p + 1); x.b = __ldg(p + 2);
    x.sd = sd ? __ldg(reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(sd) + j * sd_stride2) + i) : 0.0;
    x.s0 = x.s1 = 0.0;
    if (st) {
        const double* q = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(st) + j * st_stride2) + 2 * i;
        x.s0 = __ldg(q); x.s1 = __ldg(q + 1);
    }
    return x;
}

__device__ __forceinline__ void load_stats(double* sp, const MrPalette& pal, const double* tex) {
    const int K = pal.K;
    for (int t = threadIdx.x; t < 3 * K; t += blockDim.x) {
        sp[t] = __ldg(pal.mean + t);
        sp[3 * K + t] = __ldg(pal.lo + t);
        sp[6 * K + t] = __ldg(pal.hi + t);
    }
    if (tex)   // nullptr: colour-only matching, no texture statistics
        for (int t = threadIdx.x; t < 9 * K; t += blockDim.x) sp[9 * K + t] = __ldg(tex + t);
    __syncthreads();
}

template <int match>
__global__ void __launch_bounds__(128) k_categorize_props(MrPalette pal, const double* __restrict__ tex,
                                                          const double* __restrict__ px, int64_t stride2,
                                                          const double* __restrict__ sd, int64_t sd_stride2,
                                                          const double* __restrict__ st, int64_t st_stride2, int64_t n1,
                                                          int64_t n2, uint8_t* __restrict__ out, int64_t out_stride2,
                                                          MrCountersDev* counters) {
    extern __shared__ double sp[];
    load_stats(sp, pal, tex);
    constexpr int P = 4;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * P;
    unsigned nomatch = 0;
    if (i0 < n1) {
        const int n = (int)min((int64_t)P, n1 - i0);
        for (int64_t j = blockIdx.y; j < n2; j += gridDim.y) {
            PropPx x[P];
            int lab[P];
            const int known[P] = {0, 0, 0, 0};
            const double* pc = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(px) + j * stride2) + i0 * 3;
            const double* ps = sd ? reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(sd) + j * sd_stride2) + i0 : nullptr;
            const double* pt = st ? reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(st) + j * st_stride2) + 2 * i0 : nullptr;
            const uintptr_t al = reinterpret_cast<uintptr_t>(pc) | reinterpret_cast<uintptr_t>(ps) | reinterpret_cast<uintptr_t>(pt);
            if (n == P && (al & 15u) == 0) {      // whole 16-byte loads: 96 + 32 + 64 contiguous bytes per thread
                double v[12];
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const double2 d = __ldg(reinterpret_cast<const double2*>(pc) + t);
                    v[2 * t] = d.x; v[2 * t + 1] = d.y;
                }
#pragma unroll
                for (int q = 0; q < P; ++q) { x[q].r = v[3 * q]; x[q].g = v[3 * q + 1]; x[q].b = v[3 * q + 2]; x[q].sd = x[q].s0 = x[q].s1 = 0.0; }
                if (ps) {
                    const double2 a = __ldg(reinterpret_cast<const double2*>(ps)), b = __ldg(reinterpret_cast<const double2*>(ps) + 1);
                    x[0].sd = a.x; x[1].sd = a.y; x[2].sd = b.x; x[3].sd = b.y;
                }
                if (pt) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const double2 d = __ldg(reinterpret_cast<const double2*>(pt) + q);
                        x[q].s0 = d.x; x[q].s1 = d.y;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < P; ++q) x[q] = load_props(px, stride2, sd, sd_stride2, st, st_stride2, q < n ? i0 + q : i0, j);
            }
            eval_props<P, match>(sp, pal.K, x, known, lab);
            uint8_t* o = out + j * out_stride2 + i0;
            const uint32_t word = (uint32_t)lab[0] | ((uint32_t)lab[1] << 8) | ((uint32_t)lab[2] << 16) | ((uint32_t)lab[3] << 24);
            if (n == P && (((uintptr_t)o) & 3u) == 0) *reinterpret_cast<uint32_t*>(o) = word;
            else
                for (int q = 0; q < n; ++q) o[q] = (uint8_t)lab[q];
            for (int q = 0; q < n; ++q) nomatch += lab[q] == 0;
        }
    }
    if (counters) {
        for (int o = 16; o > 0; o >>= 1) nomatch += __shfl_xor_sync(0xffffffffu, nomatch, o);
        if ((threadIdx.x & 31) == 0 && nomatch) atomicAdd(&counters->n_nomatch, (unsigned long long)nomatch);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&counters->n_pixels, (unsigned long long)(n1 * n2));
    }
}

// clicked points: the known-category override of src/categories.jl:110-112
template <int match>
__global__ void __launch_bounds__(256) k_known_props(MrPalette pal, const double* __restrict__ tex,
                                                     const double* __restrict__ px, int64_t stride2,
                                                     const double* __restrict__ sd, int64_t sd_stride2,
                                                     const double* __restrict__ st, int64_t st_stride2, int64_t n1, int64_t n2,
                                                     int64_t n_known, const int64_t* __restrict__ idx,
                                                     const uint8_t* __restrict__ cat, uint8_t* __restrict__ out,
                                                     int64_t out_stride2) {
    extern __shared__ double sp[];
    load_stats(sp, pal, tex);
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_known) return;
    if (idx[k] < 0 || idx[k] >= n1 * n2 || cat[k] == 0) return;
    const int64_t i = idx[k] % n1, j = idx[k] / n1;
    const PropPx x[1] = {load_props(px, stride2, sd, sd_stride2, st, st_stride2, i, j)};
    const int known[1] = {cat[k]};
    int lab[1];
    eval_props<1, match>(sp, pal.K, x, known, lab);
    out[j * out_stride2 + i] = (uint8_t)lab[0];
}

dim3 grid_cols(int64_t width, int64_t n2, int threads, int sm_count) {
    const int64_t gx = (width + threads - 1) / threads;
    int64_t gy = ((int64_t)sm_count * 16 + gx - 1) / gx;
    if (gy > n2) gy = n2;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy);
}

}  // namespace

// scratch: one 8-byte scalar (the maximum)
cudaError_t mr_launch_stds(const double* px, int64_t stride2, int64_t n1, int64_t n2, double* out, int64_t out_stride2,
                           double* scratch, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(scratch, 0, 8, s);
    if (e != cudaSuccess) return e;
    k_stds_raw<<<grid_cols((n1 + SPX - 1) / SPX, n2, 128, sm_count * 2), 128, 0, s>>>(
        px, stride2, n1, n2, out, out_stride2, reinterpret_cast<unsigned long long*>(scratch));
    k_div_plane<<<grid_cols(n1, n2, 256, sm_count), 256, 0, s>>>(out, out_stride2, n1, n2, scratch);
    *n_launches += 2;
    return cudaGetLastError();
}

// scratch: the (s, l) plane, the chunk and line sums and the scalar m
size_t mr_stripes_scratch_bytes(int64_t n1, int64_t n2) {
    const int64_t chunks = (n1 + 255) / 256;
    return (size_t)(n1 * n2 * 2 + chunks * n2 * 2 + n2 * 2 + 2) * sizeof(double);
}

cudaError_t mr_launch_stripes(const double* px, int64_t stride2, int64_t n1, int64_t n2, int radius, double* out,
                              int64_t out_stride2, double* scratch, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const int64_t chunks = (n1 + 255) / 256;
    double2* sl = reinterpret_cast<double2*>(scratch);
    double* partial = scratch + n1 * n2 * 2;
    double* lines = partial + chunks * n2 * 2;
    double* m = lines + n2 * 2;
    const dim3 g = grid_cols(n1, n2, 256, sm_count);
    k_sl_plane<<<g, 256, 0, s>>>(px, stride2, n1, n2, sl);
    switch (radius) {   // the usual radii unrolled (default 3, the GUI's slider starts at 2)
        case 1: k_stripes_raw<1><<<g, 256, 0, s>>>(sl, n1, n2, radius, out, out_stride2); break;
        case 2: k_stripes_raw<2><<<g, 256, 0, s>>>(sl, n1, n2, radius, out, out_stride2); break;
        case 3: k_stripes_raw<3><<<g, 256, 0, s>>>(sl, n1, n2, radius, out, out_stride2); break;
        case 4: k_stripes_raw<4><<<g, 256, 0, s>>>(sl, n1, n2, radius, out, out_stride2); break;
        default: k_stripes_raw<0><<<g, 256, 0, s>>>(sl, n1, n2, radius, out, out_stride2); break;
    }
    k_abs_chunks<<<(unsigned)((chunks * n2 + 7) / 8), 256, 0, s>>>(out, out_stride2, n1, n2, chunks, partial);
    k_abs_lines<<<(unsigned)((n2 + 127) / 128), 128, 0, s>>>(n2, chunks, partial, lines);
    k_abs_total<<<1, 256, 0, s>>>(n1, n2, lines, m);
    k_div_plane<<<grid_cols(2 * n1, n2, 256, sm_count), 256, 0, s>>>(out, out_stride2, 2 * n1, n2, m);
    *n_launches += 6;
    return cudaGetLastError();
}

// tex: device array of 9 K doubles (std mean | lo | hi, stripe mean | lo | hi); sd / st may be nullptr when unmatched
cudaError_t mr_launch_categorize_props(const MrPalette& pal, const double* tex, int match, const double* px, int64_t stride2,
                                       const double* sd, int64_t sd_stride2, const double* st, int64_t st_stride2,
                                       int64_t n1, int64_t n2, int64_t n_known, const int64_t* known_idx,
                                       const uint8_t* known_cat, uint8_t* out, int64_t out_stride2,
                                       MrCountersDev* counters, int sm_count, cudaStream_t s, int* n_launches) {
    if (n1 <= 0 || n2 <= 0) return cudaSuccess;
    const size_t smem = (size_t)pal.K * 18 * sizeof(double);   // <= 254 * 144 B = 36.6 KB
    const dim3 g = grid_cols((n1 + 3) / 4, n2, 128, sm_count * 2);
    const unsigned gk = (unsigned)((n_known + 255) / 256);
#define MR_PROPS_CASE(M)                                                                                              \
    case M:                                                                                                           \
        k_categorize_props<M><<<g, 128, smem, s>>>(pal, tex, px, stride2, sd, sd_stride2, st, st_stride2, n1, n2, out, \
                                                   out_stride2, counters);                                            \
        if (n_known > 0)                                                                                              \
            k_known_props<M><<<gk, 256, smem, s>>>(pal, tex, px, stride2, sd, sd_stride2, st, st_stride2, n1, n2,      \
                                                   n_known, known_idx, known_cat, out, out_stride2);                  \
        break;
    switch (match) {
        MR_PROPS_CASE(1) MR_PROPS_CASE(2) MR_PROPS_CASE(3) MR_PROPS_CASE(4) MR_PROPS_CASE(5) MR_PROPS_CASE(6) MR_PROPS_CASE(7)
        default: return cudaErrorInvalidValue;
    }
#undef MR_PROPS_CASE
    *n_launches += n_known > 0 ? 2 : 1;
    return cudaGetLastError();
}
