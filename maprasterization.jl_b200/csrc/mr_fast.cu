// mr_fast.cu -- the tuned fused categorise + clean kernel for packed RGB{N0f8} sheets (sm_100a).
//
// Data flow per CTA (persistent, dynamic work queue of  strip x line-chunk  items):
//   phase A   warp = one line, lane = one aligned 32-pixel word: six 16-byte loads of packed RGB
//             per lane, colours -> labels through the 32 KiB coarse table in shared memory, labels
//             packed four per register in a bit-transpose-friendly order and stored into a rolling
//             ring of TH+2R lines (a) as bytes and (b) as NB bit planes.  The plane words are stored
//             OVERLAPPED: vote lane l owns the 32 pixels starting R before its U = 32-2R output
//             pixels, so a vote window never leaves the lane's own words.
//   phase A'  the rare colours that fall into a MIXED coarse cell (~1 %) are resolved from the exact
//             2^24-entry table (L2-resident) by a dense work list, not by divergent lanes.
//   phase B   Window{R} majority vote, bit-sliced: per word two candidate labels (A, B); their
//             (2R+1)^2 box counts come from plane-wise match masks, a vertical carry-save sum and a
//             horizontal shift-and-add sum.  A pixel is decided when max(cA,cB) > N - cA - cB (no
//             other label can reach or tie it); the winner between A and B is the larger count,
//             ties -> lower label (Appendix A.3).  Lines leave through a per-warp staging line with
//             16-byte stores.
//   phase C   undecided pixels (3-label junctions, clipped corners) get the exact histogram vote
//             from the pristine byte ring, again through a dense work list.
// Every path is exact: the candidate choice is only a performance heuristic.
//
// Reference semantics: src/categories.jl:76-133 (via the table built in mr_generic.cu) and the
// north_star-defined majority filter, SURVEY.md Appendix A.3.
#include "mr_bitslice.h"
#include "mr_internal.h"

namespace {

constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;
constexpr int TH = 16;              // lines voted per step
constexpr int CHUNK = 128;          // lines per work item
constexpr int SENT_CAP = 2048;
constexpr int FB_CAP = 2048;
constexpr unsigned VOIDB = 0xFFu;   // byte label of "outside the image" (all planes 1)

// Strip geometry for window radius R: U = 32-2R useful pixels per vote word, 32 vote words per
// warp -> SW = 32*U output pixels per strip = U aligned 32-pixel words, plus one aligned halo word
// on each side (R > 0).  Aligned word = lane in phase A; vote word = lane in phase B.
// Window coordinate wx = pixel x - xw0, xw0 = strip*SW - 32*HW.
template <int R> struct Geo {
    static constexpr int U = 32 - 2 * R;
    static constexpr int HW = R > 0 ? 1 : 0;          // halo words per side
    static constexpr int SW = 32 * U;                 // output pixels per strip
    static constexpr int AW = U + 2 * HW;             // aligned words loaded per line
    static constexpr int O0 = 32 * HW - R;            // wx of bit 0 of vote word 0
    static constexpr int WV = R == 1 ? 2 : 3;         // bits of a vertical sum (<= 2R+1)
    static constexpr int WC = R == 1 ? 4 : (R == 2 ? 5 : 6);   // bits of a box count (<= (2R+1)^2)
};

struct FastParams {
    MrJob job;
    int n_strips;
    int n_chunks;
    long long n_items;
    unsigned int* work_counter;
    int padded;   // lines are readable / writable up to the next multiple of 16 bytes
};

struct Smem {
    uint8_t* coarse;
    uint32_t* planes;   // [RING][NB][32]  overlapped vote words
    uint8_t* bytes;     // [RING][1024]    natural order, index wx
    uint8_t* stage;     // [NWARPS][1024]
    uint16_t* sent;     // codes slot<<10 | wx
    uint16_t* fb;
    uint16_t* w2code;   // round-2 word list: slot<<5 | vote lane
    uint32_t* w2mask;   //                    undecided pixels of that word
    int* misc;          // [0] n_sent [1] n_fb [2] item [3] fb overflow [4] n_w2 [8 .. 8+RING) line of slot
};

template <int R, int NB>
__host__ __device__ constexpr size_t smem_bytes() {
    return MR_COARSE_SIZE + (size_t)(TH + 2 * R) * (R > 0 ? NB : 0) * 128 + (size_t)(TH + 2 * R) * 1024 +
           (R > 0 ? NWARPS * 1024 : 0) + SENT_CAP * 2 + FB_CAP * 2 + TH * 32 * 6 + (8 + TH + 2 * R) * 4 + 16;
}

__device__ __forceinline__ uint4 ldg16(const uint8_t* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}

// ---------------------------------------------------------------------------------------------
// exact resolution of one MIXED-cell pixel (phase A' / overflow path)
template <int R, int NB>
__device__ __forceinline__ void fix_sentinel(const FastParams& P, const Smem& S, const uint8_t* px_sheet,
                                             long long xw0, int code, int y0, int y1, unsigned& c_fine) {
    using G = Geo<R>;
    const int slot = code >> 10, wx = code & 1023;
    const int y = S.misc[8 + slot];
    const long long x = xw0 + wx;
    const uint8_t* p = px_sheet + (long long)y * P.job.stride2 + x * 3;
    const unsigned rgb = (unsigned)__ldg(p) | ((unsigned)__ldg(p + 1) << 8) | ((unsigned)__ldg(p + 2) << 16);
    const unsigned l = __ldg(P.job.lut + rgb);
    S.bytes[slot * 1024 + wx] = (uint8_t)l;
    if (R > 0 && wx >= G::O0) {
        // the pixel lives in up to two overlapped vote words
        const int lh = (wx - G::O0) / G::U;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int vl = lh - t;
            const int q = wx - G::O0 - vl * G::U;
            if (vl < 0 || vl > 31 || q < 0 || q > 31) continue;
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (!((l >> b) & 1u)) atomicAnd(&S.planes[(slot * NB + b) * 32 + vl], ~(1u << q));
        }
    }
    if (wx >= 32 * G::HW && wx < 32 * G::HW + G::SW && y >= y0 && y < y1) ++c_fine;
}

// ---------------------------------------------------------------------------------------------
// phase A: one warp categorises one line of the strip window into ring slot `slot`
template <int R, int NB>
__device__ __forceinline__ void categorize_line(const FastParams& P, const Smem& S, const uint8_t* px_sheet,
                                                long long xw0, int y, int slot, int lane, unsigned cmask,
                                                unsigned vmask, unsigned smask, bool edge_strip, int y0,
                                                int y1, unsigned& c_fine) {
    using G = Geo<R>;
    uint32_t* pl = S.planes + slot * NB * 32;
    uint8_t* by = S.bytes + slot * 1024;
    if (lane == 0) S.misc[8 + slot] = y;
    const bool line_ok = (y >= -P.job.halo_lo) && (y < P.job.n2 + P.job.halo_hi);
    if (!line_ok) {   // outside the sheet: VOID everywhere (warp-uniform)
        const uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
        *reinterpret_cast<uint4*>(by + lane * 32) = v;
        *reinterpret_cast<uint4*>(by + lane * 32 + 16) = v;
        if (R > 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) pl[b * 32 + lane] = ~0u;
        }
        return;
    }
    const uint8_t* src = px_sheet + (long long)y * P.job.stride2 + (xw0 + (long long)lane * 32) * 3;
    uint32_t w[24];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
        if ((cmask >> c) & 1u) v = ldg16(src + 16 * c);
        w[4 * c + 0] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
    // labels, packed so that register j holds pixels {j, 8+j, 16+j, 24+j} in bytes 0..3: the bit
    // transpose below then needs no final gather
    uint32_t Wp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Wp[j] = 0u;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
        uint32_t p[4];
        p[0] = w0 & 0xFFFFFFu;
        p[1] = __funnelshift_r(w0, w1, 24) & 0xFFFFFFu;
        p[2] = __funnelshift_r(w1, w2, 16) & 0xFFFFFFu;
        p[3] = w2 >> 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t a = p[i] >> 3, b = p[i] >> 6, c = p[i] >> 9;
            const uint32_t t = (a & 0x1Fu) | (b & ~0x1Fu);
            const uint32_t idx = (t & 0x3FFu) | (c & ~0x3FFu);     // r5 | g5<<5 | b5<<10
            const uint32_t l = S.coarse[idx];
            Wp[4 * (g & 1) + i] |= l << (8 * (g >> 1));
        }
    }
    if (edge_strip) {   // warp-uniform: force VOID outside [0, n1)
#pragma unroll
        for (int j = 0; j < 8; ++j) Wp[j] |= ((~vmask >> j) & 0x01010101u) * 0xFFu;
    }
    // bytes in natural order (pixel 4n+k in byte k of register n)
    {
        uint32_t nat[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int k = n >> 1, base = 4 * (n & 1);
            const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
            const uint32_t lo = __byte_perm(Wp[base + 0], Wp[base + 1], sel);
            const uint32_t hi = __byte_perm(Wp[base + 2], Wp[base + 3], sel);
            nat[n] = __byte_perm(lo, hi, 0x5410);
        }
        *reinterpret_cast<uint4*>(by + lane * 32) = make_uint4(nat[0], nat[1], nat[2], nat[3]);
        *reinterpret_cast<uint4*>(by + lane * 32 + 16) = make_uint4(nat[4], nat[5], nat[6], nat[7]);
    }
    if (R > 0) {
        // vote word `lane` starts at wx = O0 + lane*U: funnel of aligned words k, k+1
        const int o = G::O0 + lane * G::U;
        const int k = o >> 5, sh = o & 31;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            uint32_t t = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t s = (j >= b) ? (Wp[j] << (j - b)) : (Wp[j] >> (b - j));
                t |= s & (0x01010101u << j);
            }
            // t: bit q = bit b of the label of pixel q of aligned word `lane`
            const uint32_t lo = __shfl_sync(0xffffffffu, t, k);
            const uint32_t hi = __shfl_sync(0xffffffffu, t, (k + 1) & 31);
            pl[b * 32 + lane] = __funnelshift_r(lo, hi, sh);
        }
    }
    // MIXED-cell pixels (label byte 0xFF on a real, loaded pixel) -> dense work list
    const uint32_t any = (Wp[0] | Wp[1] | Wp[2] | Wp[3] | Wp[4] | Wp[5] | Wp[6] | Wp[7]) & 0x80808080u;
    if (any) {
        uint32_t s7 = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) s7 |= ((Wp[j] >> 7) & 0x01010101u) << j;
        s7 &= smask;
        while (s7) {
            const int q = __ffs(s7) - 1;
            s7 &= s7 - 1;
            const int code = (slot << 10) | (lane * 32 + q);
            const int pos = atomicAdd(&S.misc[0], 1);
            if (pos < SENT_CAP) S.sent[pos] = (uint16_t)code;
            else fix_sentinel<R, NB>(P, S, px_sheet, xw0, code, y0, y1, c_fine);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// exact vote of one pixel (phase C): plane words of the pixel's vote word, one popcount pass per
// distinct label in the window; stops as soon as no unseen label can still win or tie.
template <int R, int NB>
__device__ __forceinline__ unsigned exact_vote(const Smem& S, int slot_of_y, int wx, int ring) {
    using G = Geo<R>;
    constexpr int SIDE = 2 * R + 1;
    const int vl = (wx - 32 * G::HW) / G::U;            // vote word that owns wx as an output pixel
    const int q = wx - G::O0 - vl * G::U;               // bit of the pixel, in [R, R+U)
    const uint32_t wm = ((1u << SIDE) - 1u) << (q - R);
    uint32_t Pl[SIDE][NB], rem[SIDE];
    int left = 0;
#pragma unroll
    for (int d = 0; d < SIDE; ++d) {
        int sl = slot_of_y + d - R; sl += (sl < 0) ? ring : 0; sl -= (sl >= ring) ? ring : 0;
        uint32_t vd = ~0u;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            Pl[d][b] = S.planes[(sl * NB + b) * 32 + vl];
            vd &= Pl[d][b];
        }
        rem[d] = wm & ~vd;                               // real (non-VOID) window pixels not yet counted
        left += __popc(rem[d]);
    }
    int best = 0x100, bc = 0;
#pragma unroll
    for (int d = 0; d < SIDE; ++d) {
        while (rem[d] && left >= bc) {
            const int i = __ffs(rem[d]) - 1;
            unsigned c = 0;
            uint32_t L[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t bit = (Pl[d][b] >> i) & 1u;
                c |= bit << b;
                L[b] = 0u - bit;
            }
            int cnt = 0;
#pragma unroll
            for (int e = 0; e < SIDE; ++e) {
                uint32_t m = rem[e];
#pragma unroll
                for (int b = 0; b < NB; ++b) m &= ~(Pl[e][b] ^ L[b]);
                cnt += __popc(m);
                rem[e] &= ~m;
            }
            left -= cnt;
            if (cnt > bc || (cnt == bc && (int)c < best)) { bc = cnt; best = (int)c; }
        }
    }
    return (unsigned)best;
}

template <int R, int NB, bool COUNT>
__device__ __forceinline__ void fallback_pixel(const FastParams& P, const Smem& S, uint8_t* out_sheet,
                                               long long xw0, int code, int ring, unsigned& c_nomatch,
                                               unsigned& c_changed) {
    const int slot = code >> 10, wx = code & 1023;
    const int y = S.misc[8 + slot];
    const unsigned best = exact_vote<R, NB>(S, slot, wx, ring);
    out_sheet[(long long)y * P.job.out_stride2 + xw0 + wx] = (uint8_t)best;
    if (COUNT) {
        c_nomatch += (best == 0);
        c_changed += (best != S.bytes[slot * 1024 + wx]);
    }
}

// store the 32 labels of one aligned word (natural order) to global memory
__device__ __forceinline__ void store_word(const FastParams& P, uint8_t* dst, long long x0, uint4 v0, uint4 v1) {
    const long long n1 = P.job.n1;
    if (x0 + 32 <= n1) {
        *reinterpret_cast<uint4*>(dst) = v0;
        *reinterpret_cast<uint4*>(dst + 16) = v1;
        return;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long xs = x0 + 16 * h;
        const uint4 v = h ? v1 : v0;
        if (xs >= n1) break;
        if (xs + 16 <= n1 || P.padded) {
            *reinterpret_cast<uint4*>(dst + 16 * h) = v;
        } else {
            const uint32_t a[4] = {v.x, v.y, v.z, v.w};
            for (int i = 0; i < 16 && xs + i < n1; ++i) dst[16 * h + i] = (uint8_t)(a[i >> 2] >> (8 * (i & 3)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bit-sliced two-candidate vote of one 32-pixel vote word.  Pl = plane words of lines y-R..y+R.
// decided: pixels whose mode is certainly the better of {A, B}; selA: that winner is A.
template <int R, int NB>
__device__ __forceinline__ void vote_core(const uint32_t (&Pl)[2 * R + 1][NB], unsigned a_lab, unsigned b_lab,
                                          uint32_t& decided, uint32_t& selA, uint32_t& diff, uint32_t& zero) {
    using G = Geo<R>;
    constexpr int SIDE = 2 * R + 1;
    constexpr int N = SIDE * SIDE;
    constexpr int WV = G::WV, WC = G::WC;
    constexpr unsigned VOIDL = (1u << NB) - 1u;
    const uint32_t nbm = (b_lab != a_lab && b_lab != VOIDL) ? ~0u : 0u;   // B == A -> B's masks are empty
    const uint32_t avm = (a_lab != VOIDL) ? ~0u : 0u;     // a bitwise majority can alias the VOID code
    uint32_t LA[NB], LB[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        LA[b] = 0u - ((a_lab >> b) & 1u);
        LB[b] = 0u - ((b_lab >> b) & 1u);
    }
    // ---- match masks + vertical sums
    uint32_t mA[SIDE], mB[SIDE];
#pragma unroll
    for (int d = 0; d < SIDE; ++d) {
        uint32_t ma = avm, mb = nbm;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            ma &= ~(Pl[d][b] ^ LA[b]);
            mb &= ~(Pl[d][b] ^ LB[b]);
        }
        mA[d] = ma;
        mB[d] = mb;
    }
    Bs<WV> VA, VB;
    if constexpr (R == 1) { VA = bs_sum3(mA[0], mA[1], mA[2]); VB = bs_sum3(mB[0], mB[1], mB[2]); }
    if constexpr (R == 2) {
        VA = bs_sum5(mA[0], mA[1], mA[2], mA[3], mA[4]);
        VB = bs_sum5(mB[0], mB[1], mB[2], mB[3], mB[4]);
    }
    if constexpr (R == 3) {
        VA = bs_sum7(mA[0], mA[1], mA[2], mA[3], mA[4], mA[5], mA[6]);
        VB = bs_sum7(mB[0], mB[1], mB[2], mB[3], mB[4], mB[5], mB[6]);
    }
    // ---- horizontal sums: the word carries its own R-pixel halo on both sides
    Bs<WV> zz;
#pragma unroll
    for (int k = 0; k < WV; ++k) zz.b[k] = 0u;
    const Bs<WC> cA = bs_hsum<R, WV, WC>(zz, VA, zz);
    const Bs<WC> cB = bs_hsum<R, WV, WC>(zz, VB, zz);
    // ---- decide: winner of {A, B} is the mode iff max(cA, cB) > N - cA - cB
    uint32_t eq;
    const uint32_t bgt = bs_gt<WC>(cB, cA, &eq);
    const uint32_t a_lower = (a_lab < b_lab) ? ~0u : 0u;
    selA = ~bgt & (~eq | a_lower);
    const Bs<WC> mx = bs_sel<WC>(bgt, cB, cA), mn = bs_sel<WC>(bgt, cA, cB);
    Bs<WC + 1> mx2;
    mx2.b[0] = 0u;
#pragma unroll
    for (int k = 0; k < WC; ++k) mx2.b[k + 1] = mx.b[k];
    const Bs<WC + 2> u = bs_add<WC + 1, WC, WC + 2>(mx2, mn);
    decided = bs_gt_const<WC + 2>(u, (unsigned)N);
    diff = 0u;
    zero = ~0u;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t o = (LA[b] & selA) | (LB[b] & ~selA);
        diff |= o ^ Pl[R][b];
        zero &= ~o;
    }
}

template <int R, int NB>
__device__ __forceinline__ void load_planes(const Smem& S, int slot, int ring, int vl, uint32_t (&Pl)[2 * R + 1][NB]) {
#pragma unroll
    for (int d = 0; d < 2 * R + 1; ++d) {
        int sl = slot + d - R; sl += (sl < 0) ? ring : 0; sl -= (sl >= ring) ? ring : 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) Pl[d][b] = S.planes[(sl * NB + b) * 32 + vl];
    }
}

// label at bit `pos`: vertical bitwise majority of the three centre lines (robust to isolated
// speckle) unless the line above / below is outside the sheet
template <int R, int NB>
__device__ __forceinline__ unsigned label_maj3(const uint32_t (&Pl)[2 * R + 1][NB], bool vedge, int pos) {
    unsigned l = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t m = vedge ? Pl[R][b] : mr_maj(Pl[R - 1][b], Pl[R][b], Pl[R + 1][b]);
        l |= ((m >> pos) & 1u) << b;
    }
    return l;
}

// phase B, round 1: one warp votes one line.  vvalid: output pixels owned by this lane's vote
// word; avalid: output pixels of this lane's aligned word (for the final store).
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void vote_line(const FastParams& P, const Smem& S, uint8_t* out_sheet, long long xw0,
                                          int y, int slot, int ring, int lane, int warp, unsigned vvalid,
                                          unsigned avalid, unsigned& c_nomatch, unsigned& c_changed) {
    using G = Geo<R>;
    constexpr int SIDE = 2 * R + 1;
    constexpr unsigned VOIDL = (1u << NB) - 1u;
    const uint8_t* by = S.bytes + slot * 1024;
    uint8_t* st = S.stage + warp * 1024;
    // pristine bytes of this lane's aligned word -> staging (patched below, then stored)
    {
        const uint4 v0 = *reinterpret_cast<const uint4*>(by + lane * 32);
        const uint4 v1 = *reinterpret_cast<const uint4*>(by + lane * 32 + 16);
        *reinterpret_cast<uint4*>(st + lane * 32) = v0;
        *reinterpret_cast<uint4*>(st + lane * 32 + 16) = v1;
    }
    __syncwarp();
    if (vvalid) {
        uint32_t Pl[SIDE][NB];
        load_planes<R, NB>(S, slot, ring, lane, Pl);
        // ---- candidate labels (heuristic only).  A / B: last / first output pixel of the word; if
        // they agree, B comes from the top / bottom line of the window instead.
        const bool vedge = (y - 1 < -P.job.halo_lo) || (y + 1 >= P.job.n2 + P.job.halo_hi);   // warp-uniform
        const int posA = 31 - __clz(vvalid), posB = __ffs(vvalid) - 1;
        const int posM = ((vvalid >> 16) & 1u) ? 16 : posA;
        const unsigned a_lab = label_maj3<R, NB>(Pl, vedge, posA);
        unsigned b_lab = label_maj3<R, NB>(Pl, vedge, posB);
        if (b_lab == a_lab) {
            unsigned t_lab = 0, u_lab = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                t_lab |= ((Pl[0][b] >> posM) & 1u) << b;
                u_lab |= ((Pl[SIDE - 1][b] >> posM) & 1u) << b;
            }
            b_lab = (t_lab != a_lab && t_lab != VOIDL) ? t_lab : u_lab;
        }
        uint32_t decided, selA, diff, zero;
        vote_core<R, NB>(Pl, a_lab, b_lab, decided, selA, diff, zero);
        uint32_t changed = decided & vvalid & diff;
        const uint32_t undec = ~decided & vvalid;
        if (COUNT) {
            c_changed += __popc(changed);
            c_nomatch += __popc(decided & vvalid & zero);
        }
        const int o0 = G::O0 + lane * G::U;   // wx of bit 0 of this vote word
        while (changed) {
            const int q = __ffs(changed) - 1;
            changed &= changed - 1;
            st[o0 + q] = (uint8_t)(((selA >> q) & 1u) ? a_lab : b_lab);
        }
        if (undec) {   // second round with candidates taken from the undecided pixels (dense list)
            const int pos = atomicAdd(&S.misc[4], 1);
            S.w2code[pos] = (uint16_t)((slot << 5) | lane);
            S.w2mask[pos] = undec;
        }
    }
    __syncwarp();
    if (avalid) {   // this lane's aligned word holds output pixels
        const uint4 v0 = *reinterpret_cast<const uint4*>(st + lane * 32);
        const uint4 v1 = *reinterpret_cast<const uint4*>(st + lane * 32 + 16);
        const long long x0 = xw0 + (long long)lane * 32;
        store_word(P, out_sheet + (long long)y * P.job.out_stride2 + x0, x0, v0, v1);
    }
    __syncwarp();   // staging is reused by this warp's next line
}

// phase B, round 2: one thread re-votes one word whose round-1 candidates left pixels undecided.
// Candidates are the labels at the last / first undecided pixel.  Newly decided pixels go
// straight to global memory; the rest is queued for the exact per-pixel vote.
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void vote_word2(const FastParams& P, const Smem& S, uint8_t* out_sheet, long long xw0,
                                           int code, uint32_t undec, int ring, unsigned& c_nomatch,
                                           unsigned& c_changed) {
    using G = Geo<R>;
    constexpr int SIDE = 2 * R + 1;
    const int slot = code >> 5, vl = code & 31;
    const int y = S.misc[8 + slot];
    uint32_t Pl[SIDE][NB];
    load_planes<R, NB>(S, slot, ring, vl, Pl);
    const bool vedge = (y - 1 < -P.job.halo_lo) || (y + 1 >= P.job.n2 + P.job.halo_hi);
    const int posA = 31 - __clz(undec), posB = __ffs(undec) - 1;
    const unsigned a_lab = label_maj3<R, NB>(Pl, vedge, posA);
    const unsigned b_lab = label_maj3<R, NB>(Pl, vedge, posB);
    uint32_t decided, selA, diff, zero;
    vote_core<R, NB>(Pl, a_lab, b_lab, decided, selA, diff, zero);
    const int o0 = G::O0 + vl * G::U;
    uint8_t* orow = out_sheet + (long long)y * P.job.out_stride2 + xw0 + o0;
    uint32_t now = undec & decided;
    if (COUNT) {
        c_changed += __popc(now & diff);
        c_nomatch += __popc(now & zero);
    }
    now &= diff;   // unchanged pixels already hold the right (pristine) label
    while (now) {
        const int q = __ffs(now) - 1;
        now &= now - 1;
        orow[q] = (uint8_t)(((selA >> q) & 1u) ? a_lab : b_lab);
    }
    uint32_t still = undec & ~decided;
    while (still) {
        const int q = __ffs(still) - 1;
        still &= still - 1;
        const int pos = atomicAdd(&S.misc[1], 1);
        if (pos < FB_CAP) S.fb[pos] = (uint16_t)((slot << 10) | (o0 + q));
        else S.misc[3] = 1;   // list full: phase C re-votes the whole step exactly
    }
}

// R == 0: categorise only -- emit the byte ring line
__device__ __forceinline__ void emit_line(const FastParams& P, const Smem& S, uint8_t* out_sheet, long long xw0,
                                          int y, int slot, int lane, unsigned vmask, bool count,
                                          unsigned& c_nomatch) {
    if (vmask == 0u) return;
    const uint8_t* by = S.bytes + slot * 1024;
    const uint4 v0 = *reinterpret_cast<const uint4*>(by + lane * 32);
    const uint4 v1 = *reinterpret_cast<const uint4*>(by + lane * 32 + 16);
    if (count) {
        const uint32_t W[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const uint32_t nz = (W[n] | (W[n] >> 4)) & 0x0F0F0F0Fu;
            const uint32_t nz2 = (nz | (nz >> 1) | (nz >> 2) | (nz >> 3)) & 0x01010101u;   // 1 where byte != 0
            // valid pixels 4n .. 4n+3 -> bits 0, 8, 16, 24
            const uint32_t v4 = (vmask >> (4 * n)) & 0xFu;
            const uint32_t vm = (v4 & 1u) | ((v4 & 2u) << 7) | ((v4 & 4u) << 14) | ((v4 & 8u) << 21);
            c_nomatch += __popc(vm & ~nz2);
        }
    }
    const long long x0 = xw0 + (long long)lane * 32;
    store_word(P, out_sheet + (long long)y * P.job.out_stride2 + x0, x0, v0, v1);
}

// ---------------------------------------------------------------------------------------------
template <int R, int NB, bool COUNT>
__global__ void __launch_bounds__(NTHREADS, 2) k_fast(const FastParams P) {
    using G = Geo<R>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int RING = TH + 2 * R;
    constexpr int NBP = R > 0 ? NB : 0;
    Smem S;
    S.coarse = smem_raw;
    S.planes = reinterpret_cast<uint32_t*>(smem_raw + MR_COARSE_SIZE);
    S.bytes = reinterpret_cast<uint8_t*>(S.planes + RING * NBP * 32);
    S.stage = S.bytes + RING * 1024;
    S.sent = reinterpret_cast<uint16_t*>(S.stage + (R > 0 ? NWARPS * 1024 : 0));
    S.fb = S.sent + SENT_CAP;
    S.w2mask = reinterpret_cast<uint32_t*>(S.fb + FB_CAP);
    S.w2code = reinterpret_cast<uint16_t*>(S.w2mask + TH * 32);
    S.misc = reinterpret_cast<int*>(S.w2code + TH * 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.job.coarse);
        uint4* dst = reinterpret_cast<uint4*>(S.coarse);
        for (int i = tid; i < (int)(MR_COARSE_SIZE / 16); i += NTHREADS) dst[i] = __ldg(src + i);
    }
    if (tid < 8) S.misc[tid] = 0;
    unsigned c_fine = 0, c_nomatch = 0, c_changed = 0;
    const long long n1 = P.job.n1;
    const int n2 = (int)P.job.n2;

    for (;;) {
        __syncthreads();
        if (tid == 0) S.misc[2] = (int)atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const long long item = (unsigned)S.misc[2];
        if (item >= P.n_items) break;
        const int per_sheet = P.n_strips * P.n_chunks;
        const int sheet = (int)(item / per_sheet);
        const int rem = (int)(item - (long long)sheet * per_sheet);
        const int chunk = rem / P.n_strips, strip = rem - chunk * P.n_strips;   // neighbouring strips first
        const uint8_t* px_sheet = P.job.px + (long long)sheet * P.job.sheet_stride;
        uint8_t* out_sheet = P.job.out + (long long)sheet * P.job.out_sheet_stride;
        const long long xw0 = (long long)strip * G::SW - 32 * G::HW;     // pixel x of wx = 0
        const int y0 = chunk * CHUNK, y1 = min(y0 + CHUNK, n2);
        const int ybase = y0 - R;
        // ---- per-lane masks for this strip
        // aligned word `lane` (phase A, stores): pixels inside [0, n1)
        const long long x0 = xw0 + (long long)lane * 32;
        unsigned vmask = 0u;
        if (lane < G::AW && x0 >= 0 && x0 < n1) vmask = (x0 + 32 <= n1) ? ~0u : ((1u << (int)(n1 - x0)) - 1u);
        const bool edge_strip = (strip == 0) || (xw0 + 32 * G::AW > n1);
        const long long limit = P.padded ? ((n1 * 3 + 15) / 16) * 16 : n1 * 3;
        unsigned cmask = 0u;
#pragma unroll
        for (int c = 0; c < 6; ++c)
            if (lane < G::AW && x0 >= 0 && x0 * 3 + 16 * c + 16 <= limit) cmask |= 1u << c;
        unsigned smask = vmask;   // pixels whose label is real (loaded and inside the sheet)
        if (R > 0) {              // halo words only need the pixels next to the interior
            if (lane == 0) { cmask &= 1u << 5; smask &= 0xF8000000u; }
            if (lane == G::AW - 1) { cmask &= 1u << 0; smask &= 0x0000001Fu; }
        }
        const unsigned avalid = (lane >= G::HW && lane < G::HW + G::U) ? vmask : 0u;   // output words
        // vote word `lane` (phase B): output pixels are bits [R, R+U) that lie inside [0, n1)
        unsigned vvalid = 0u;
        if (R > 0) {
            const long long xv = xw0 + G::O0 + (long long)lane * G::U + R;   // x of the first output bit
            long long cnt = n1 - xv;
            cnt = cnt < 0 ? 0 : (cnt > G::U ? G::U : cnt);
            if (cnt > 0) vvalid = ((1u << (int)cnt) - 1u) << R;
        }

        for (int a = y0; a < y1; a += TH) {
            const int vote_end = min(a + TH, y1);
            const int cat_begin = (a == y0) ? a - R : a + R;
            const int cat_end = vote_end + R;
            // ---- phase A
            for (int y = cat_begin + warp; y < cat_end; y += NWARPS)
                categorize_line<R, NB>(P, S, px_sheet, xw0, y, (y - ybase) % RING, lane, cmask, vmask, smask,
                                       edge_strip, y0, y1, c_fine);
            __syncthreads();
            // ---- phase A': resolve MIXED-cell pixels from the exact table
            {
                const int n = min(S.misc[0], SENT_CAP);
                for (int e = tid; e < n; e += NTHREADS)
                    fix_sentinel<R, NB>(P, S, px_sheet, xw0, S.sent[e], y0, y1, c_fine);
                if (tid == 0) { S.misc[1] = 0; S.misc[3] = 0; S.misc[4] = 0; }
            }
            __syncthreads();
            // ---- phase B
            if (tid == 0) S.misc[0] = 0;
            if constexpr (R > 0) {
                unsigned t_nomatch = 0, t_changed = 0;   // this step's counts of the decided pixels
                for (int y = a + warp; y < vote_end; y += NWARPS)
                    vote_line<R, NB, COUNT>(P, S, out_sheet, xw0, y, (y - ybase) % RING, RING, lane, warp, vvalid,
                                            avalid, t_nomatch, t_changed);
                __syncthreads();
                // ---- phase B2: words with undecided pixels, new candidates (dense)
                {
                    const int n = S.misc[4];
                    for (int e = tid; e < n; e += NTHREADS)
                        vote_word2<R, NB, COUNT>(P, S, out_sheet, xw0, S.w2code[e], S.w2mask[e], RING, t_nomatch,
                                                 t_changed);
                }
                __syncthreads();
                // ---- phase C: exact vote of the still undecided pixels
                if (S.misc[3] == 0) {
                    c_nomatch += t_nomatch;
                    c_changed += t_changed;
                    const int n = S.misc[1];
                    for (int e = tid; e < n; e += NTHREADS)
                        fallback_pixel<R, NB, COUNT>(P, S, out_sheet, xw0, S.fb[e], RING, c_nomatch, c_changed);
                } else {
                    // work list overflowed (adversarial input): all pixels of the step are re-voted
                    // and re-counted exactly (the step's bit-sliced counts are dropped)
                    const int npx = (vote_end - a) * G::SW;
                    for (int e = tid; e < npx; e += NTHREADS) {
                        const int ly = e / G::SW, wx = 32 * G::HW + (e - ly * G::SW);
                        if (xw0 + wx >= n1) continue;
                        const int yy = a + ly, sl = (yy - ybase) % RING;
                        const unsigned best = exact_vote<R, NB>(S, sl, wx, RING);
                        uint8_t* o = out_sheet + (long long)yy * P.job.out_stride2 + xw0 + wx;
                        if (COUNT) {
                            c_nomatch += (best == 0);
                            c_changed += (best != S.bytes[sl * 1024 + wx]);
                        }
                        *o = (uint8_t)best;
                    }
                }
            } else {
                for (int y = a + warp; y < vote_end; y += NWARPS)
                    emit_line(P, S, out_sheet, xw0, y, (y - ybase) % RING, lane, avalid, COUNT, c_nomatch);
            }
            __syncthreads();
        }
    }
    if (COUNT && P.job.counters) {
        for (int o = 16; o > 0; o >>= 1) {
            c_fine += __shfl_xor_sync(0xffffffffu, c_fine, o);
            c_nomatch += __shfl_xor_sync(0xffffffffu, c_nomatch, o);
            c_changed += __shfl_xor_sync(0xffffffffu, c_changed, o);
        }
        if (lane == 0) {
            atomicAdd(&P.job.counters->n_fine, (unsigned long long)c_fine);
            atomicAdd(&P.job.counters->n_nomatch, (unsigned long long)c_nomatch);
            atomicAdd(&P.job.counters->n_changed, (unsigned long long)c_changed);
        }
        if (blockIdx.x == 0 && tid == 0)
            atomicAdd(&P.job.counters->n_pixels, (unsigned long long)(P.job.n1 * P.job.n2 * P.job.n_sheets));
    }
}

// ---------------------------------------------------------------------------------------------
unsigned int* g_work[16] = {nullptr};
int g_work_next[16] = {0};
constexpr int WORK_SLOTS = 1024;

template <int R, int NB, bool COUNT>
cudaError_t launch_t(const FastParams& P, int sm_count, cudaStream_t s) {
    auto kern = k_fast<R, NB, COUNT>;
    constexpr size_t smem = smem_bytes<R, NB>();
    static bool configured[16] = {false};
    static int occ[16] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 16 && !configured[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[dev], kern, NTHREADS, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int per_sm = (dev < 16 && occ[dev] > 0) ? occ[dev] : 1;
    long long grid = (long long)sm_count * per_sm;
    if (grid > P.n_items) grid = P.n_items;
    kern<<<(unsigned)grid, NTHREADS, smem, s>>>(P);
    return cudaGetLastError();
}

template <int R, int NB>
cudaError_t launch_c(const FastParams& P, int sm_count, cudaStream_t s) {
    return P.job.counters ? launch_t<R, NB, true>(P, sm_count, s) : launch_t<R, NB, false>(P, sm_count, s);
}

template <int NB>
cudaError_t launch_r(const FastParams& P, int sm_count, cudaStream_t s) {
    switch (P.job.R) {
        case 1: return launch_c<1, NB>(P, sm_count, s);
        case 2: return launch_c<2, NB>(P, sm_count, s);
        case 3: return launch_c<3, NB>(P, sm_count, s);
        default: return cudaErrorNotSupported;
    }
}

template <int R>
int strips_for(long long n1) { return (int)((n1 + Geo<R>::SW - 1) / Geo<R>::SW); }

}  // namespace

bool mr_fast_supported(const MrJob& job) {
    if (job.bpp != 3 || job.lut == nullptr || job.coarse == nullptr) return false;
    if (job.pal.K > 126 || job.R < 0 || job.R > 3) return false;
    if (job.n1 <= 0 || job.n2 <= 0 || job.n2 > (1 << 30) || job.n_sheets <= 0) return false;
    const bool a16 = ((uintptr_t)job.px % 16 == 0) && ((uintptr_t)job.out % 16 == 0) && (job.stride2 % 16 == 0) &&
                     (job.out_stride2 % 16 == 0) && (job.sheet_stride % 16 == 0) && (job.out_sheet_stride % 16 == 0);
    if (!a16) return false;
    if (!job.padded && (job.n1 % 16 != 0)) return false;
    return true;
}

cudaError_t mr_launch_fast(const MrJob& job, int sm_count, cudaStream_t s, int* n_launches) {
    if (!mr_fast_supported(job)) return cudaErrorNotSupported;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 16) return cudaErrorNotSupported;
    if (!g_work[dev]) {
        e = cudaMalloc(&g_work[dev], WORK_SLOTS * sizeof(unsigned int));
        if (e != cudaSuccess) return e;
    }
    unsigned int* ctr = g_work[dev] + (g_work_next[dev]++ % WORK_SLOTS);
    e = cudaMemsetAsync(ctr, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    FastParams P;
    P.job = job;
    P.n_strips = job.R == 0 ? strips_for<0>(job.n1)
                            : (job.R == 1 ? strips_for<1>(job.n1)
                                          : (job.R == 2 ? strips_for<2>(job.n1) : strips_for<3>(job.n1)));
    P.n_chunks = (int)((job.n2 + CHUNK - 1) / CHUNK);
    P.n_items = (long long)P.n_strips * P.n_chunks * job.n_sheets;
    if (P.n_items >= (1ll << 31)) return cudaErrorNotSupported;
    P.work_counter = ctr;
    P.padded = job.padded;
    const int K = job.pal.K;
    if (job.R == 0) e = launch_c<0, 4>(P, sm_count, s);
    else if (K <= 14) e = launch_r<4>(P, sm_count, s);
    else if (K <= 30) e = launch_r<5>(P, sm_count, s);
    else if (K <= 62) e = launch_r<6>(P, sm_count, s);
    else e = launch_r<7>(P, sm_count, s);
    if (e == cudaSuccess) *n_launches += 1;
    return e;
}
