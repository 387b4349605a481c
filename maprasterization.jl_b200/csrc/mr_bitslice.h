// mr_bitslice.h -- bit-sliced ("vertical") integer arithmetic on 32 pixels at once.
//
// A W-bit unsigned number per pixel is held as W 32-bit words, word k carrying bit k of all
// 32 pixels (bit i of every word belongs to pixel i).  Additions and comparisons then cost a
// handful of 3-input logic ops (LOP3) per bit for 32 pixels, which is what lets the window
// majority vote of SURVEY.md Appendix A.3 run inside ~10 instructions per pixel.
// Host+device: plain C++ apart from the funnel shifts, which have a host fallback.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MR_HD __host__ __device__ __forceinline__
#else
#define MR_HD inline
#endif

template <int W>
struct Bs {
    uint32_t b[W];   // b[0] = least significant bit plane
};

MR_HD uint32_t mr_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

// out = a + b, WO bits kept (caller guarantees the true sum fits).
template <int WA, int WB, int WO>
MR_HD Bs<WO> bs_add(const Bs<WA>& a, const Bs<WB>& b) {
    Bs<WO> o;
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < WO; ++k) {
        const uint32_t x = k < WA ? a.b[k] : 0u;
        const uint32_t y = k < WB ? b.b[k] : 0u;
        o.b[k] = x ^ y ^ carry;
        carry = mr_maj(x, y, carry);
    }
    return o;
}

// a > b (per pixel), equal widths; eq_out (optional) = a == b
template <int W>
MR_HD uint32_t bs_gt(const Bs<W>& a, const Bs<W>& b, uint32_t* eq_out) {
    uint32_t gt = 0u, eq = ~0u;
#pragma unroll
    for (int k = W - 1; k >= 0; --k) {
        gt |= eq & a.b[k] & ~b.b[k];
        eq &= ~(a.b[k] ^ b.b[k]);
    }
    if (eq_out) *eq_out = eq;
    return gt;
}

// a > constant c (per pixel)
template <int W>
MR_HD uint32_t bs_gt_const(const Bs<W>& a, unsigned c) {
    uint32_t gt = 0u, eq = ~0u;
#pragma unroll
    for (int k = W - 1; k >= 0; --k) {
        if ((c >> k) & 1u) {
            eq &= a.b[k];
        } else {
            gt |= eq & a.b[k];
            eq &= ~a.b[k];
        }
    }
    return gt;
}

// per-pixel select: m ? a : b
template <int W>
MR_HD Bs<W> bs_sel(uint32_t m, const Bs<W>& a, const Bs<W>& b) {
    Bs<W> o;
#pragma unroll
    for (int k = 0; k < W; ++k) o.b[k] = (a.b[k] & m) | (b.b[k] & ~m);
    return o;
}

// ---- sums of 2R+1 one-bit planes -> small counters (the vertical pass of the box sum)
MR_HD Bs<2> bs_sum3(uint32_t m0, uint32_t m1, uint32_t m2) {
    Bs<2> o;
    o.b[0] = m0 ^ m1 ^ m2;
    o.b[1] = mr_maj(m0, m1, m2);
    return o;
}
MR_HD Bs<3> bs_sum5(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t m4) {
    const uint32_t s1 = m0 ^ m1 ^ m2, c1 = mr_maj(m0, m1, m2);
    const uint32_t s2 = s1 ^ m3 ^ m4, c2 = mr_maj(s1, m3, m4);
    Bs<3> o;
    o.b[0] = s2;
    o.b[1] = c1 ^ c2;
    o.b[2] = c1 & c2;
    return o;
}
MR_HD Bs<3> bs_sum7(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t m4, uint32_t m5,
                    uint32_t m6) {
    const uint32_t s1 = m0 ^ m1 ^ m2, c1 = mr_maj(m0, m1, m2);
    const uint32_t s2 = m3 ^ m4 ^ m5, c2 = mr_maj(m3, m4, m5);
    const uint32_t s3 = s1 ^ s2 ^ m6, c3 = mr_maj(s1, s2, m6);
    Bs<3> o;
    o.b[0] = s3;
    o.b[1] = c1 ^ c2 ^ c3;
    o.b[2] = mr_maj(c1, c2, c3);
    return o;
}

// out = a - b (per pixel), WA >= WB, caller guarantees a >= b.
template <int WA, int WB>
MR_HD Bs<WA> bs_sub(const Bs<WA>& a, const Bs<WB>& b) {
    Bs<WA> o;
    uint32_t borrow = 0;
#pragma unroll
    for (int k = 0; k < WA; ++k) {
        const uint32_t x = a.b[k];
        const uint32_t y = k < WB ? b.b[k] : 0u;
        o.b[k] = x ^ y ^ borrow;
        borrow = (~x & (y | borrow)) | (y & borrow);
    }
    return o;
}

// Horizontal window sum of a ONE-bit plane: out(x) = sum over dx in [-R, R] of m(x + dx); bits that
// fall off the word count as 0 (the vote words carry their own R-pixel halo, see mr_fast.cu).
// R=1: 0..3 (2 bits), R=2: 0..5 (3 bits), R=3: 0..7 (3 bits).
template <int R>
MR_HD Bs<(R == 1 ? 2 : 3)> bs_hsum1(uint32_t m) {
    if constexpr (R == 1) {
        return bs_sum3(m << 1, m, m >> 1);
    } else if constexpr (R == 2) {
        return bs_sum5(m << 2, m << 1, m, m >> 1, m >> 2);
    } else {
        return bs_sum7(m << 3, m << 2, m << 1, m, m >> 1, m >> 2, m >> 3);
    }
}

// funnel shifts over the (left | centre | right) triple of neighbouring words:
// value of pixel i+d (d > 0: towards the right neighbour word) delivered at bit i.
MR_HD uint32_t mr_take_right(uint32_t centre, uint32_t right, int d) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(centre, right, d);
#else
    return d == 0 ? centre : (centre >> d) | (right << (32 - d));
#endif
}
MR_HD uint32_t mr_take_left(uint32_t left, uint32_t centre, int d) {   // pixel i-d at bit i
#ifdef __CUDA_ARCH__
    return __funnelshift_l(left, centre, d);
#else
    return d == 0 ? centre : (centre << d) | (left >> (32 - d));
#endif
}

// Horizontal pass of the box sum: sum over dx in [-R, R] of v(x + dx), v is WV bits wide,
// result WO bits (R=1: 9 -> 4 bits, R=2: 25 -> 5 bits, R=3: 49 -> 6 bits).
template <int R, int WV, int WO>
MR_HD Bs<WO> bs_hsum(const Bs<WV>& l, const Bs<WV>& c, const Bs<WV>& r) {
    Bs<WO> acc;
#pragma unroll
    for (int k = 0; k < WO; ++k) acc.b[k] = k < WV ? c.b[k] : 0u;
#pragma unroll
    for (int d = 1; d <= R; ++d) {
        Bs<WV> a, b;
#pragma unroll
        for (int k = 0; k < WV; ++k) {
            a.b[k] = mr_take_right(c.b[k], r.b[k], d);
            b.b[k] = mr_take_left(l.b[k], c.b[k], d);
        }
        const Bs<WV + 1> p = bs_add<WV, WV, WV + 1>(a, b);
        acc = bs_add<WO, WV + 1, WO>(acc, p);
    }
    return acc;
}
