// mr_fast.cu -- the tuned fused categorise + clean kernel for packed RGB{N0f8} sheets (sm_100a).
//
// Persistent CTAs pull  strip x line-chunk  work items from a global queue and run them through a
// four-stage SOFTWARE PIPELINE, one block barrier per tick, all stages in flight every tick
// (stage k of tick t works on step t-k, a step = TH lines of one item):
//   SA  categorise: warp = one line, lane = one aligned 32-pixel word.  Six 16-byte loads of
//       packed RGB per lane, colours -> labels through the 32 KiB coarse table in shared memory.
//       The labels go (a) straight to global memory as the provisional output (16-byte stores) and
//       (b) as NB bit planes into a shared-memory ring.  Plane words are stored OVERLAPPED: vote
//       lane l owns the 32 pixels starting R before its U = 32-2R output pixels, so a vote window
//       never leaves the lane's own words.
//   SF  the rare colours that fall into a MIXED coarse cell (~1 %) are resolved from the exact
//       2^24-entry table (L2 resident) through a dense work list, not by divergent lanes.
//   SB  Window{R} majority vote, bit-sliced: per word two candidate labels (A, B); their (2R+1)^2
//       box counts come from plane-wise match masks, a vertical carry-save sum and a horizontal
//       shift-and-add sum.  A pixel is decided when max(cA,cB) > N - cA - cB (no other label can
//       reach or tie it); the winner of {A, B} is the larger count, ties -> lower label (Appendix
//       A.3).  Only pixels whose label changes are written (byte stores into the L2-hot line).
//   SC  words with undecided pixels are re-voted with candidates taken from those pixels; what is
//       still undecided (3-label junctions, clipped corners) gets an exact popcount vote.
// Every path is exact: the candidate choice is only a performance heuristic.
//
// Reference semantics: src/categories.jl:76-133 (via the table built in mr_generic.cu) and the
// north_star-defined majority filter, SURVEY.md Appendix A.3.
#include "mr_bitslice.h"
#include "mr_internal.h"

namespace {

constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;
constexpr int TH = 8;               // lines per step (one per warp)
constexpr int CHUNK = 128;          // lines per work item
constexpr int RINGL = 56;           // plane ring, lines: 5 steps in flight + halos (5*TH + 4R <= 52)
constexpr int SENT_CAP = 1024;
constexpr int W2_CAP = TH * 32;
constexpr int FB_CAP = 1024;        // undecided pixels per step queued for the exact vote
constexpr int NDESC = 8;
constexpr int STAGE_BYTES = 32 * 96;   // one line of the strip window, packed RGB

// Strip geometry for window radius R: U = 32-2R useful pixels per vote word, 32 vote words per
// warp -> SW = 32*U output pixels per strip = U aligned 32-pixel words, plus one aligned halo word
// on each side (R > 0).  Aligned word = lane in SA; vote word = lane in SB.
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

// one step of one item, as seen by the later pipeline stages
struct StepDesc {
    long long xw0;            // pixel x of wx = 0
    const uint8_t* px;        // sheet base (input)
    uint8_t* out;             // sheet base (output)
    int y0, y1;               // the item's own lines
    int a, vote_end;          // lines voted in this step
    int cat_begin, cat_end;   // lines categorised in this step
    int pos0;                 // ring position of line y0 - R
    int item;                 // item id, -1 = empty step
    int edge;                 // strip touches the left / right sheet border
    int pad_;
};

// shared-memory layout (byte offsets from the dynamic shared memory base)
template <int R, int NB> struct Lay {
    static constexpr int NBP = R > 0 ? NB : 0;
    static constexpr int COARSE = 0;
    static constexpr int PLANES = COARSE + (int)MR_COARSE_SIZE;          // uint32 [RINGL][NB][32]
    static constexpr int SENT = PLANES + RINGL * NBP * 128;              // uint16 [3][SENT_CAP]
    static constexpr int W2MASK = SENT + 3 * SENT_CAP * 2;               // uint32 [3][W2_CAP]
    static constexpr int W2CODE = W2MASK + 3 * W2_CAP * 4;               // uint16 [3][W2_CAP]
    static constexpr int FB = W2CODE + 3 * W2_CAP * 2;                   // uint16 [3][FB_CAP]
    static constexpr int DESC = FB + 3 * FB_CAP * 2;                     // StepDesc [NDESC]
    static constexpr int LINEOF = DESC + NDESC * (int)sizeof(StepDesc);  // int [RINGL]
    static constexpr int MISC = LINEOF + RINGL * 4;                      // int [16]
    static constexpr int MBAR = MISC + 16 * 4;                           // uint64 [NWARPS] TMA barriers
    static constexpr int STAGE = (MBAR + NWARPS * 8 + 127) / 128 * 128;  // uint8 [NWARPS][STAGE_BYTES]
    static constexpr int TOTAL = STAGE + NWARPS * STAGE_BYTES;
};
// misc: [0..2] n_sent, [4..6] n_w2, [8..9] prefetched item ids, [12..14] n_fb

__device__ __forceinline__ uint4 ldg16(const uint8_t* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ int ring_slot(int pos) { return pos % RINGL; }

// ---- TMA (1-D bulk async copy, global -> shared) with mbarrier completion
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// per-lane validity masks of a strip
struct LaneMasks {
    unsigned vmask;    // aligned word `lane`: pixels inside [0, n1)
    unsigned cmask;    // 16-byte chunks of the aligned word to load
    unsigned smask;    // pixels whose label is real (loaded and inside the sheet)
    unsigned avalid;   // aligned word holds output pixels (vmask if interior word)
    unsigned vvalid;   // vote word `lane`: output pixels inside [0, n1)
};

template <int R>
__device__ __forceinline__ LaneMasks lane_masks(const FastParams& P, long long xw0, int lane) {
    using G = Geo<R>;
    const long long n1 = P.job.n1;
    LaneMasks m;
    const long long x0 = xw0 + (long long)lane * 32;
    m.vmask = 0u;
    if (lane < G::AW && x0 >= 0 && x0 < n1) m.vmask = (x0 + 32 <= n1) ? ~0u : ((1u << (int)(n1 - x0)) - 1u);
    const long long limit = P.padded ? ((n1 * 3 + 15) / 16) * 16 : n1 * 3;
    m.cmask = 0u;
#pragma unroll
    for (int c = 0; c < 6; ++c)
        if (lane < G::AW && x0 >= 0 && x0 * 3 + 16 * c + 16 <= limit) m.cmask |= 1u << c;
    m.smask = m.vmask;
    if (R > 0) {   // halo words only need the pixels next to the interior
        if (lane == 0) { m.cmask &= 1u << 5; m.smask &= 0xF8000000u; }
        if (lane == G::AW - 1) { m.cmask &= 1u << 0; m.smask &= 0x0000001Fu; }
    }
    m.avalid = (lane >= G::HW && lane < G::HW + G::U) ? m.vmask : 0u;
    m.vvalid = 0u;
    if (R > 0) {
        const long long xv = xw0 + G::O0 + (long long)lane * G::U + R;   // x of the first output bit
        long long cnt = n1 - xv;
        cnt = cnt < 0 ? 0 : (cnt > G::U ? G::U : cnt);
        if (cnt > 0) m.vvalid = ((1u << (int)cnt) - 1u) << R;
    }
    return m;
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
// SF: exact resolution of one MIXED-cell pixel
template <int R, int NB>
__device__ __forceinline__ void fix_sentinel(const FastParams& P, uint8_t* sm, const StepDesc& d, int code,
                                             unsigned& c_fine, unsigned& c_nomatch) {
    using G = Geo<R>;
    using L = Lay<R, NB>;
    const int slot = code >> 10, wx = code & 1023;
    const int y = reinterpret_cast<const int*>(sm + L::LINEOF)[slot];
    const long long x = d.xw0 + wx;
    const uint8_t* p = d.px + (long long)y * P.job.stride2 + x * 3;
    const unsigned rgb = (unsigned)__ldg(p) | ((unsigned)__ldg(p + 1) << 8) | ((unsigned)__ldg(p + 2) << 16);
    const unsigned l = __ldg(P.job.lut + rgb);
    if (R > 0 && wx >= G::O0) {
        // the pixel lives in up to two overlapped vote words
        uint32_t* planes = reinterpret_cast<uint32_t*>(sm + L::PLANES);
        const int lh = (wx - G::O0) / G::U;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int vl = lh - t;
            const int q = wx - G::O0 - vl * G::U;
            if (vl < 0 || vl > 31 || q < 0 || q > 31) continue;
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (!((l >> b) & 1u)) atomicAnd(&planes[(slot * NB + b) * 32 + vl], ~(1u << q));
        }
    }
    if (wx >= 32 * G::HW && wx < 32 * G::HW + G::SW && y >= d.y0 && y < d.y1) {   // owned pixel
        d.out[(long long)y * P.job.out_stride2 + x] = (uint8_t)l;                  // provisional label
        ++c_fine;
        if (R == 0) c_nomatch += (l == 0);
    }
}

// ---------------------------------------------------------------------------------------------
// SA: one warp categorises one line of the strip window
// staged != nullptr: the packed RGB of this line was brought to shared memory by TMA
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void categorize_line(const FastParams& P, uint8_t* sm, const StepDesc& d, int y,
                                                int lane, const LaneMasks& M, int sent_buf,
                                                const uint8_t* staged, unsigned& c_fine, unsigned& c_nomatch) {
    using G = Geo<R>;
    using L = Lay<R, NB>;
    const int slot = ring_slot(d.pos0 + (y - (d.y0 - R)));
    uint32_t* pl = reinterpret_cast<uint32_t*>(sm + L::PLANES) + slot * NB * 32;
    if (lane == 0) reinterpret_cast<int*>(sm + L::LINEOF)[slot] = y;
    const bool line_ok = (y >= -P.job.halo_lo) && (y < P.job.n2 + P.job.halo_hi);
    if (!line_ok) {   // outside the sheet: VOID everywhere (warp-uniform)
        if (R > 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) pl[b * 32 + lane] = ~0u;
        }
        return;
    }
    const uint8_t* src = d.px + (long long)y * P.job.stride2 + (d.xw0 + (long long)lane * 32) * 3;
    uint32_t w[24];
    if (staged) {   // warp-uniform
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
            if ((M.cmask >> c) & 1u) v = *reinterpret_cast<const uint4*>(staged + lane * 96 + 16 * c);
            w[4 * c + 0] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
            if ((M.cmask >> c) & 1u) v = ldg16(src + 16 * c);
            w[4 * c + 0] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
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
            const uint32_t l = sm[L::COARSE + idx];
            Wp[4 * (g & 1) + i] |= l << (8 * (g >> 1));
        }
    }
    if (d.edge) {   // warp-uniform: force VOID outside [0, n1)
#pragma unroll
        for (int j = 0; j < 8; ++j) Wp[j] |= ((~M.vmask >> j) & 0x01010101u) * 0xFFu;
    }
    // provisional output: the categorised labels of the words / lines this item owns
    const bool own_line = (y >= d.y0) && (y < d.y1);
    if (own_line && M.avalid) {
        uint32_t nat[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {   // natural order: pixel 4n+k in byte k of register n
            const int k = n >> 1, base = 4 * (n & 1);
            const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
            const uint32_t lo = __byte_perm(Wp[base + 0], Wp[base + 1], sel);
            const uint32_t hi = __byte_perm(Wp[base + 2], Wp[base + 3], sel);
            nat[n] = __byte_perm(lo, hi, 0x5410);
        }
        if (COUNT && R == 0) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const uint32_t nz = (nat[n] | (nat[n] >> 4)) & 0x0F0F0F0Fu;
                const uint32_t nz2 = (nz | (nz >> 1) | (nz >> 2) | (nz >> 3)) & 0x01010101u;   // byte != 0
                const uint32_t v4 = (M.avalid >> (4 * n)) & 0xFu;
                const uint32_t vm = (v4 & 1u) | ((v4 & 2u) << 7) | ((v4 & 4u) << 14) | ((v4 & 8u) << 21);
                c_nomatch += __popc(vm & ~nz2);
            }
        }
        const long long x0 = d.xw0 + (long long)lane * 32;
        store_word(P, d.out + (long long)y * P.job.out_stride2 + x0, x0, make_uint4(nat[0], nat[1], nat[2], nat[3]),
                   make_uint4(nat[4], nat[5], nat[6], nat[7]));
    }
    uint32_t s_all = ~0u;   // pixels whose label byte has all NB low bits set: MIXED (0xFF) or VOID
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
            s_all &= t;
            const uint32_t lo = __shfl_sync(0xffffffffu, t, k);
            const uint32_t hi = __shfl_sync(0xffffffffu, t, (k + 1) & 31);
            pl[b * 32 + lane] = __funnelshift_r(lo, hi, sh);
        }
    } else {
        s_all = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) s_all |= ((Wp[j] >> 7) & 0x01010101u) << j;
    }
    // MIXED-cell pixels (real, loaded pixels only) -> dense work list for SF
    uint32_t s7 = s_all & M.smask;
    if (s7) {
        int* n_sent = reinterpret_cast<int*>(sm + L::MISC) + sent_buf;
        uint16_t* sent = reinterpret_cast<uint16_t*>(sm + L::SENT) + sent_buf * SENT_CAP;
        int pos = atomicAdd(n_sent, __popc(s7));
        while (s7) {
            const int q = __ffs(s7) - 1;
            s7 &= s7 - 1;
            const int code = (slot << 10) | (lane * 32 + q);
            if (pos < SENT_CAP) sent[pos] = (uint16_t)code;
            else fix_sentinel<R, NB>(P, sm, d, code, c_fine, c_nomatch);   // list full: resolve inline
            ++pos;
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

// plane words of lines y-R .. y+R (ring positions pos-R .. pos+R) of vote word vl
template <int R, int NB>
__device__ __forceinline__ void load_planes(const uint8_t* sm, int pos, int vl, uint32_t (&Pl)[2 * R + 1][NB]) {
    using L = Lay<R, NB>;
    const uint32_t* planes = reinterpret_cast<const uint32_t*>(sm + L::PLANES);
    int sl = ring_slot(pos - R);
#pragma unroll
    for (int d = 0; d < 2 * R + 1; ++d) {
#pragma unroll
        for (int b = 0; b < NB; ++b) Pl[d][b] = planes[(sl * NB + b) * 32 + vl];
        sl = (sl + 1 == RINGL) ? 0 : sl + 1;
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

// exact vote of the pixel at bit q (in [R, R+U)) of a vote word: one popcount pass per distinct
// label in the window; stops as soon as no unseen label can still win or tie.
template <int R, int NB>
__device__ __forceinline__ unsigned exact_vote(const uint32_t (&Pl)[2 * R + 1][NB], int q) {
    constexpr int SIDE = 2 * R + 1;
    const uint32_t wm = ((1u << SIDE) - 1u) << (q - R);
    uint32_t rem[SIDE];
    int left = 0;
#pragma unroll
    for (int d = 0; d < SIDE; ++d) {
        uint32_t vd = ~0u;
#pragma unroll
        for (int b = 0; b < NB; ++b) vd &= Pl[d][b];
        rem[d] = wm & ~vd;                               // real (non-VOID) window pixels not yet counted
        left += __popc(rem[d]);
    }
    int best = 0x100, bc = 0;
#pragma unroll
    for (int d = 0; d < SIDE; ++d) {
        while (rem[d] && left >= bc) {
            const int i = __ffs(rem[d]) - 1;
            unsigned c = 0;
            uint32_t Lm[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t bit = (Pl[d][b] >> i) & 1u;
                c |= bit << b;
                Lm[b] = 0u - bit;
            }
            int cnt = 0;
#pragma unroll
            for (int e = 0; e < SIDE; ++e) {
                uint32_t m = rem[e];
#pragma unroll
                for (int b = 0; b < NB; ++b) m &= ~(Pl[e][b] ^ Lm[b]);
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
__device__ __forceinline__ void exact_pixel(const FastParams&, const uint32_t (&Pl)[2 * R + 1][NB], uint8_t* orow, int q,
                                            unsigned& c_nomatch, unsigned& c_changed) {
    const unsigned best = exact_vote<R, NB>(Pl, q);
    unsigned centre = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) centre |= ((Pl[R][b] >> q) & 1u) << b;
    if (best != centre) orow[q] = (uint8_t)best;   // the provisional label is the centre label
    if (COUNT) {
        c_nomatch += (best == 0);
        c_changed += (best != centre);
    }
}

// ---------------------------------------------------------------------------------------------
// SB: one warp votes one line (round 1)
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void vote_line(const FastParams& P, uint8_t* sm, const StepDesc& d, int y, int lane,
                                          unsigned vvalid, int w2_buf, unsigned& c_nomatch, unsigned& c_changed) {
    using G = Geo<R>;
    using L = Lay<R, NB>;
    constexpr int SIDE = 2 * R + 1;
    constexpr unsigned VOIDL = (1u << NB) - 1u;
    if (!vvalid) return;
    const int pos = d.pos0 + (y - (d.y0 - R));
    uint32_t Pl[SIDE][NB];
    load_planes<R, NB>(sm, pos, lane, Pl);
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
    uint8_t* orow = d.out + (long long)y * P.job.out_stride2 + d.xw0 + o0;
    while (changed) {   // the provisional labels are already in global memory: patch only changes
        const int q = __ffs(changed) - 1;
        changed &= changed - 1;
        orow[q] = (uint8_t)(((selA >> q) & 1u) ? a_lab : b_lab);
    }
    if (undec) {   // second round with candidates taken from the undecided pixels (dense list, SC)
        int* n_w2 = reinterpret_cast<int*>(sm + L::MISC) + 4 + w2_buf;
        const int e = atomicAdd(n_w2, 1);
        reinterpret_cast<uint16_t*>(sm + L::W2CODE)[w2_buf * W2_CAP + e] = (uint16_t)(((y - d.a) << 5) | lane);
        reinterpret_cast<uint32_t*>(sm + L::W2MASK)[w2_buf * W2_CAP + e] = undec;
    }
}

// SC: one thread re-votes one word whose round-1 candidates left pixels undecided.  Candidates are
// the labels at the last / first undecided pixel; what is still undecided is voted exactly.
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void vote_word2(const FastParams& P, uint8_t* sm, const StepDesc& d, int code,
                                           uint32_t undec, int fb_buf, unsigned& c_nomatch, unsigned& c_changed) {
    using G = Geo<R>;
    constexpr int SIDE = 2 * R + 1;
    const int y = d.a + (code >> 5), vl = code & 31;
    const int pos = d.pos0 + (y - (d.y0 - R));
    uint32_t Pl[SIDE][NB];
    load_planes<R, NB>(sm, pos, vl, Pl);
    const bool vedge = (y - 1 < -P.job.halo_lo) || (y + 1 >= P.job.n2 + P.job.halo_hi);
    const int posA = 31 - __clz(undec), posB = __ffs(undec) - 1;
    const unsigned a_lab = label_maj3<R, NB>(Pl, vedge, posA);
    const unsigned b_lab = label_maj3<R, NB>(Pl, vedge, posB);
    uint32_t decided, selA, diff, zero;
    vote_core<R, NB>(Pl, a_lab, b_lab, decided, selA, diff, zero);
    const int o0 = G::O0 + vl * G::U;
    uint8_t* orow = d.out + (long long)y * P.job.out_stride2 + d.xw0 + o0;
    uint32_t now = undec & decided;
    if (COUNT) {
        c_changed += __popc(now & diff);
        c_nomatch += __popc(now & zero);
    }
    now &= diff;   // unchanged pixels already hold the right (provisional) label
    while (now) {
        const int q = __ffs(now) - 1;
        now &= now - 1;
        orow[q] = (uint8_t)(((selA >> q) & 1u) ? a_lab : b_lab);
    }
    uint32_t still = undec & ~decided;
    if (still) {   // exact votes are spread over all threads in the next stage (SD)
        using L = Lay<R, NB>;
        int* n_fb = reinterpret_cast<int*>(sm + L::MISC) + 12 + fb_buf;
        uint16_t* fb = reinterpret_cast<uint16_t*>(sm + L::FB) + fb_buf * FB_CAP;
        int pos = atomicAdd(n_fb, __popc(still));
        while (still) {
            const int q = __ffs(still) - 1;
            still &= still - 1;
            if (pos < FB_CAP) fb[pos] = (uint16_t)((code << 5) | q);
            else exact_pixel<R, NB, COUNT>(P, Pl, orow, q, c_nomatch, c_changed);   // list full
            ++pos;
        }
    }
}

// SD: exact vote of one still undecided pixel (code = line-in-step << 10 | vote word << 5 | bit)
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void vote_pixel(const FastParams& P, uint8_t* sm, const StepDesc& d, int code,
                                           unsigned& c_nomatch, unsigned& c_changed) {
    using G = Geo<R>;
    const int y = d.a + (code >> 10), vl = (code >> 5) & 31, q = code & 31;
    const int pos = d.pos0 + (y - (d.y0 - R));
    uint32_t Pl[2 * R + 1][NB];
    load_planes<R, NB>(sm, pos, vl, Pl);
    uint8_t* orow = d.out + (long long)y * P.job.out_stride2 + d.xw0 + G::O0 + vl * G::U;
    exact_pixel<R, NB, COUNT>(P, Pl, orow, q, c_nomatch, c_changed);
}

// ---------------------------------------------------------------------------------------------
// step generator: identical state in all threads; yields the steps of successive work items
struct StepGen {
    StepDesc cur;
    int next_a, cursor;
};

// produce the next step; `prefetched` is the item id thread 0 fetched one tick earlier.
// took = the prefetched id was consumed (thread 0 must fetch another one).
template <int R>
__device__ __forceinline__ StepDesc next_step(const FastParams& P, StepGen& g, int prefetched, bool& took) {
    using G = Geo<R>;
    took = false;
    if (g.cur.item < 0 || g.next_a >= g.cur.y1) {
        const unsigned item = (unsigned)prefetched;
        if (item < (unsigned)P.n_items) {
            took = true;
            const unsigned per_sheet = (unsigned)(P.n_strips * P.n_chunks);
            const int sheet = (int)(item / per_sheet);
            const int rem = (int)(item - (unsigned)sheet * per_sheet);
            const int chunk = rem / P.n_strips, strip = rem - chunk * P.n_strips;   // neighbouring strips first
            g.cur.px = P.job.px + (long long)sheet * P.job.sheet_stride;
            g.cur.out = P.job.out + (long long)sheet * P.job.out_sheet_stride;
            g.cur.xw0 = (long long)strip * G::SW - 32 * G::HW;
            g.cur.y0 = chunk * CHUNK;
            g.cur.y1 = min(g.cur.y0 + CHUNK, (int)P.job.n2);
            g.cur.pos0 = g.cursor;
            g.cursor += (g.cur.y1 - g.cur.y0) + 2 * R;
            if (g.cursor >= RINGL * 1024) g.cursor -= RINGL * 1024;
            g.cur.item = (int)item;
            g.cur.edge = (strip == 0) || (g.cur.xw0 + 32 * G::AW > P.job.n1);
            g.next_a = g.cur.y0;
        } else {
            g.cur.item = -1;   // queue exhausted
        }
    }
    StepDesc d = g.cur;
    if (g.cur.item >= 0) {
        d.a = g.next_a;
        d.vote_end = min(g.next_a + TH, g.cur.y1);
        d.cat_begin = (g.next_a == g.cur.y0) ? g.next_a - R : g.next_a + R;
        d.cat_end = d.vote_end + R;
        g.next_a += TH;
    }
    return d;
}

// TMA: bring the packed RGB of line `y` of step d (this warp's first SA line) into the warp's
// staging buffer.  Returns false if the line needs no load (outside the sheet / no such line).
template <int R>
__device__ __forceinline__ bool issue_line_copy(const FastParams& P, const StepDesc& d, int y, uint32_t stage,
                                                uint32_t bar, int lane) {
    using G = Geo<R>;
    if (d.item < 0 || y >= d.cat_end) return false;
    if (y < -P.job.halo_lo || y >= P.job.n2 + P.job.halo_hi) return false;
    const long long limit = P.padded ? ((P.job.n1 * 3 + 15) / 16) * 16 : P.job.n1 * 3;
    const long long w0 = d.xw0 * 3;                       // byte offset of the window in the line
    long long lo = w0 + (R > 0 ? 80 : 0);                 // halo words: only the chunk next to the interior
    long long hi = w0 + (long long)G::AW * 96 - (R > 0 ? 80 : 0);
    lo = lo < 0 ? 0 : lo;
    hi = hi > limit ? limit : hi;
    if (hi <= lo) return false;
    if (lane == 0) {
        const uint32_t bytes = (uint32_t)(hi - lo);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(stage + (uint32_t)(lo - w0), d.px + (long long)y * P.job.stride2 + lo, bytes, bar);
    }
    return true;
}

template <int R, int NB, bool COUNT>
__global__ void __launch_bounds__(NTHREADS, 2) k_fast(const FastParams P) {
    using L = Lay<R, NB>;
    extern __shared__ __align__(128) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* misc = reinterpret_cast<int*>(sm + L::MISC);
    StepDesc* descs = reinterpret_cast<StepDesc*>(sm + L::DESC);
    const uint32_t my_bar = smem_u32(sm + L::MBAR + warp * 8);
    uint8_t* my_stage = sm + L::STAGE + warp * STAGE_BYTES;
    const uint32_t my_stage_u32 = smem_u32(my_stage);
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.job.coarse);
        uint4* dst = reinterpret_cast<uint4*>(sm + L::COARSE);
        for (int i = tid; i < (int)(MR_COARSE_SIZE / 16); i += NTHREADS) dst[i] = __ldg(src + i);
    }
    if (tid < 8) misc[tid] = 0;
    if (tid >= 12 && tid < 16) misc[tid] = 0;
    if (tid < NDESC) descs[tid].item = -1;
    if (tid == 0) {
        misc[8] = (int)atomicAdd(P.work_counter, 1u);
        misc[9] = 0;
        for (int w = 0; w < NWARPS; ++w) mbar_init(smem_u32(sm + L::MBAR + w * 8), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned c_fine = 0, c_nomatch = 0, c_changed = 0;

    StepGen gen;
    gen.cur.item = -1;
    gen.next_a = 0;
    gen.cursor = 0;
    // cached per-lane masks of the strips SA / SB are working on
    LaneMasks MA, MB;
    long long MA_x = -(1ll << 60), MB_x = -(1ll << 60);
    MA.vmask = MA.cmask = MA.smask = MA.avalid = MA.vvalid = 0u;
    MB = MA;

    // ---- prologue: step 0 and its TMA loads
    __syncthreads();
    bool took;
    StepDesc dN = next_step<R>(P, gen, misc[8], took);        // the step SA works on next
    if (tid == 0) misc[9] = took ? (int)atomicAdd(P.work_counter, 1u) : misc[8];
    uint32_t parity = 0;                                       // phase of this warp's TMA barrier
    bool pend = issue_line_copy<R>(P, dN, dN.cat_begin + warp, my_stage_u32, my_bar, lane);

    int idle = 0;
    for (int t = 0;; ++t) {
        __syncthreads();
        const StepDesc dA = dN;                                // step t
        // ---- look one step ahead (its loads are issued during this tick)
        const int prefetched = misc[8 + ((t + 1) & 1)];
        dN = next_step<R>(P, gen, prefetched, took);           // step t+1
        if (dA.item >= 0) {
            idle = 0;
        } else {
            ++idle;
            if (idle > 4) break;   // pipeline drained
        }
        if (tid == 0) {
            descs[t & (NDESC - 1)] = dA;
            misc[8 + (t & 1)] = took ? (int)atomicAdd(P.work_counter, 1u) : prefetched;
            misc[(t + 1) % 3] = 0;             // sent list SA fills next tick
            misc[4 + (t + 2) % 3] = 0;         // w2 list SB fills next tick (step t-1)
            misc[12 + (t + 1) % 3] = 0;        // fb list SC fills next tick (step t-2)
        }
        // ---- the stages; half of the warps start with the vote so that ALU-heavy and
        // latency-heavy stages of different warps overlap
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
            const bool front = ((warp & 1) == ph);
            if (front) {
                // ---- SA: categorise step t; the warp's first line comes from its TMA staging buffer
                if (dA.item >= 0) {
                    if (MA_x != dA.xw0) { MA = lane_masks<R>(P, dA.xw0, lane); MA_x = dA.xw0; }
                    int y = dA.cat_begin + warp;
                    if (y < dA.cat_end) {
                        if (pend) {
                            mbar_wait(my_bar, parity);
                            parity ^= 1u;
                        }
                        categorize_line<R, NB, COUNT>(P, sm, dA, y, lane, MA, t % 3, pend ? my_stage : nullptr, c_fine,
                                                      c_nomatch);
                        for (y += NWARPS; y < dA.cat_end; y += NWARPS)
                            categorize_line<R, NB, COUNT>(P, sm, dA, y, lane, MA, t % 3, nullptr, c_fine, c_nomatch);
                    }
                }
                __syncwarp();   // every lane is done reading the staging buffer
                pend = issue_line_copy<R>(P, dN, dN.cat_begin + warp, my_stage_u32, my_bar, lane);
                // ---- SF: resolve the MIXED-cell pixels of step t-1
                if (t >= 1) {
                    const StepDesc& dF = descs[(t - 1) & (NDESC - 1)];
                    if (dF.item >= 0) {
                        const int buf = (t - 1) % 3;
                        const int n = min(misc[buf], SENT_CAP);
                        const uint16_t* sent = reinterpret_cast<const uint16_t*>(sm + L::SENT) + buf * SENT_CAP;
                        for (int e = lane * NWARPS + warp; e < n; e += NTHREADS)   // spread over warps first
                            fix_sentinel<R, NB>(P, sm, dF, sent[e], c_fine, c_nomatch);
                    }
                }
            } else if constexpr (R > 0) {
                // ---- SB: vote step t-2
                if (t >= 2) {
                    const StepDesc& dB = descs[(t - 2) & (NDESC - 1)];
                    if (dB.item >= 0) {
                        if (MB_x != dB.xw0) { MB = lane_masks<R>(P, dB.xw0, lane); MB_x = dB.xw0; }
                        for (int y = dB.a + warp; y < dB.vote_end; y += NWARPS)
                            vote_line<R, NB, COUNT>(P, sm, dB, y, lane, MB.vvalid, (t - 2) % 3, c_nomatch, c_changed);
                    }
                }
                // ---- SC: second round of step t-3
                if (t >= 3) {
                    const StepDesc& dC = descs[(t - 3) & (NDESC - 1)];
                    if (dC.item >= 0) {
                        const int buf = (t - 3) % 3;
                        const int n = misc[4 + buf];
                        const uint16_t* code = reinterpret_cast<const uint16_t*>(sm + L::W2CODE) + buf * W2_CAP;
                        const uint32_t* mask = reinterpret_cast<const uint32_t*>(sm + L::W2MASK) + buf * W2_CAP;
                        for (int e = lane * NWARPS + warp; e < n; e += NTHREADS)
                            vote_word2<R, NB, COUNT>(P, sm, dC, code[e], mask[e], buf, c_nomatch, c_changed);
                    }
                }
                // ---- SD: exact votes of step t-4
                if (t >= 4) {
                    const StepDesc& dD = descs[(t - 4) & (NDESC - 1)];
                    if (dD.item >= 0) {
                        const int buf = (t - 4) % 3;
                        const int n = min(misc[12 + buf], FB_CAP);
                        const uint16_t* fb = reinterpret_cast<const uint16_t*>(sm + L::FB) + buf * FB_CAP;
                        for (int e = lane * NWARPS + warp; e < n; e += NTHREADS)
                            vote_pixel<R, NB, COUNT>(P, sm, dD, fb[e], c_nomatch, c_changed);
                    }
                }
            }
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
    constexpr size_t smem = Lay<R, NB>::TOTAL;
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
    if (P.n_items >= (1ll << 31) - 65536) return cudaErrorNotSupported;
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
