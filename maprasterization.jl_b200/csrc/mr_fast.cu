// mr_fast.cu -- the tuned fused categorise + clean kernel for packed RGB{N0f8} sheets (sm_100a),
// and, with a labels-in front end, the standalone Window majority filter.
//
// WARP-AUTONOMOUS design: one persistent CTA per SM, every warp is an independent worker that
// pulls  strip x line-chunk  items from a global queue and owns a private slice of shared memory
// (plane ring, TMA staging line, output line, round-2 list).  Only the coarse colour table is
// shared.  There is no block barrier after start-up: a warp that waits (TMA, exact-table gathers)
// is hidden by the other 15 warps, not by a barrier-synchronised phase structure.
//
// Per warp, for an item of lines [y0, y1) of a strip of 32*U output pixels (U = 32-2R), in blocks
// of B output lines (B = the largest block whose ring fits the shared memory):
//   load        the packed RGB of a line arrives by a 1-D TMA bulk copy (cp.async.bulk + mbarrier)
//               issued one line ahead (halo lines of a band possibly from a peer GPU's memory);
//               the line three further down is prefetched into L2.
//   categorise  lane = one aligned 32-pixel word.  Bytes quantised to x>>3, cell address by byte dot
//               products (IDP.4A), one LDS.U8 gather per pixel from the coarse table, labels
//               transposed into NB bit planes (delta-swap network) and stored to the warp's ring as
//               OVERLAPPED vote words: vote lane l owns the 32 pixels starting R before its U output
//               pixels, so a vote window never leaves the lane's own words.
//   MIXED       colours in a MIXED coarse cell (~1 %) take the exact 2^24-entry table (L2 resident):
//               colours captured from the staging line, loads in flight during the transpose, bits
//               patched in registers before the planes are stored.
//   vote        Window{R} majority, bit-sliced, two candidate labels (A, B) per (vote word, block):
//               per line match masks and their horizontal window sums once, box counts sliding down
//               the block.  A pixel is decided when max(cA,cB) > N - cA - cB (no other label can
//               reach or tie it); winner = larger count, ties -> lower label (Appendix A.3).
//   store       every vote word writes its U label bytes into the warp's line buffer; one TMA bulk
//               store per line.  Undecided pixels get a guess that the next stage overwrites.
//   round 2     words with undecided pixels are re-voted, densely over the lanes, with candidates
//               taken from those pixels; what is still undecided (3-label junctions, clipped
//               corners) gets an exact popcount vote, again densely.
// Every path is exact: the candidate choice is only a performance heuristic.
//
// Reference semantics: src/categories.jl:76-133 (via the table built in mr_generic.cu) and the
// north_star-defined majority filter, SURVEY.md Appendix A.3.
#include <stdlib.h>

#include <mutex>

#include "mr_bitslice.h"
#include "mr_internal.h"

namespace {

// staging line: the AW aligned words of the strip window, packed RGB (lanes >= AW read past it into
// the lists; their labels are never used)
constexpr int stage_bytes(int R) { return ((R > 0 ? 34 - 2 * R : 32) * 96 + 127) / 128 * 128; }
constexpr int LIST_CAP = 96;           // round-2 word list entries; a block that would overflow is cut short

// Strip geometry for window radius R: U = 32-2R useful pixels per vote word, 32 vote words per
// warp -> SW = 32*U output pixels per strip = U aligned 32-pixel words, plus one aligned halo word
// on each side (R > 0).  Aligned word = lane while categorising; vote word = lane while voting.
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

// kernel configuration per (R, NB): warps per CTA and lines per block, sized so that one CTA
// (coarse table + per-warp slices) fits the 227 KiB of shared memory of an SM.
// The round-2 list holds one entry per 32-pixel WORD (code + pixel mask) and is capped (LIST_CAP).
// One ring line = NB planes of 32 words, plus one word of skew: the words of one vote word on
// consecutive lines then sit in consecutive banks (round 2 reads (line, word) pairs that mostly
// share the word and differ in the line).
__host__ __device__ constexpr int ring_line(int NB) { return NB * 128 + 4; }
constexpr int ring_bytes(int R, int NB, int B) { return R > 0 ? ((B + 2 * R) * ring_line(NB) + 127) / 128 * 128 : 0; }
constexpr int warp_bytes(int R, int NB, int B) {   // per-warp slice for B lines per block
    const int raw = ring_bytes(R, NB, B) + stage_bytes(R) + LIST_CAP * 6 + (R > 0 ? 32 * (32 - 2 * R) : 0) + 24;
    return (raw + 127) / 128 * 128;
}
constexpr int best_block(int R, int NB, int warps) {   // largest block that fits (at most 16 lines)
    int b = 1;
    while (b < 16 && (int)MR_COARSE_SIZE + warps * warp_bytes(R, NB, b + 1) <= 227 * 1024) ++b;
    return b;
}
template <int R, int NB> struct Cfg {
    static constexpr int NBP = R > 0 ? NB : 0;
#ifdef MR_WARPS   // experiment: warps per CTA for small palettes
    static constexpr int WARPS = (R > 0 && NB >= 6) ? 12 : MR_WARPS;
#else
    static constexpr int WARPS = (R > 0 && NB >= 6) ? 12 : 16;
#endif
    static constexpr int B = R == 0 ? 8 : best_block(R, NB, WARPS);            // lines voted per block
    static constexpr int RL = B + 2 * R;                                       // plane ring, lines
    static constexpr int N_W2 = LIST_CAP;                                      // list capacity
    // per-warp slice (byte offsets)
    static constexpr int W_PLANES = 0;                                 // uint32 [RL][NB * 32 + 1]
    static constexpr int W_STAGE = W_PLANES + ring_bytes(R, NB, B);    // uint8  [stage_bytes(R)]
    static constexpr int W_LISTS = W_STAGE + stage_bytes(R);              // round-2 word list; the exact-vote list
    static constexpr int W_W2MASK = W_LISTS;                           // uint32 [N_W2]   overwrites it in place
    static constexpr int W_W2CODE = W_W2MASK + N_W2 * 4;               // uint16 [N_W2]
    static constexpr int W_OUT = W_W2CODE + N_W2 * 2;                  // uint8 [32 * U]: final labels of one line
    static constexpr int W_CNT = W_OUT + (R > 0 ? 32 * (32 - 2 * R) : 0);   // int [4]: -, n_w2, n_fb
    static constexpr int W_MBAR = W_CNT + 16;                          // uint64
    static constexpr int W_BYTES = (W_MBAR + 8 + 127) / 128 * 128;
    static constexpr int SLICES = (int)MR_COARSE_SIZE;                 // CTA layout: coarse | slices
    static constexpr int TOTAL = SLICES + WARPS * W_BYTES;
    static_assert(TOTAL <= 227 * 1024 && B >= 2, "shared memory budget");
};

// Work items = (column, line chunk); column = (sheet, strip).  The lines of a sheet are split into
// up to four classes with their own chunk height, handed out tallest first (guided scheduling: big
// items keep the 2R categorise-only halo lines cheap, small items at the end level the finish).
struct SchedClass {
    int line0, line1;      // lines [line0, line1) of every column
    int chunk;             // lines per work item
    unsigned first_item;   // index of the class's first item
};
struct FastParams {
    MrJob job;
    int n_strips;
    int cols;              // n_strips * n_sheets
    int n_classes;
    SchedClass cls[4];
    long long n_items;
    unsigned int* work_counter;   // this launch's item queue (zero when the launch starts)
    unsigned int* work_reset;     // a slot far ahead in the ring: zeroed here for a later launch (no memset per launch)
    int padded;   // lines are readable / writable up to the next multiple of 16 bytes
#ifdef MR_TIMELINE
    unsigned long long* timeline;   // experiment: per warp (start, first vote, end, lines) in ns
#endif
};

// one work item as seen by a warp
struct Item {
    long long xw0;            // pixel x of wx = 0
    const uint8_t* px;        // sheet base (input)
    uint8_t* out;             // sheet base (output)
    int y0, y1;               // the item's own lines
    int edge;                 // strip touches the left / right sheet border
};

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

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
#ifdef MR_KO_GATHER   // timing experiment: conflict-free gathers (all lanes read one of four bytes)
    addr = (addr & 3u) | (MR_KO_GATHER & ~3u);
#endif
#ifdef MR_KO_NOLDS   // timing experiment: no gather at all, label from the address
    return (addr >> 2) & 15u;
#endif
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));   // the table is read-only once the CTA runs
#ifdef MR_KO_NOMIXED   // timing experiment: MIXED cells never seen
    v &= 0x0Fu;
#endif
    return v;
}

// per-lane validity masks of a strip
struct LaneMasks {
    unsigned vmask;    // aligned word `lane`: pixels inside [0, n1)
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
    m.smask = m.vmask;
    if (R > 0) {   // halo words: only the 16-byte chunk next to the interior is loaded
        if (lane == 0) m.smask &= 0xF8000000u;
        if (lane == G::AW - 1) m.smask &= 0x0000001Fu;
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

// TMA extents of a strip window: which bytes of a line are copied, and where they land
struct CopyPlan {
    const uint8_t* src0;   // sheet base + first byte of the window part that is loaded
    uint32_t dst;          // shared address the first loaded byte goes to
    uint32_t bytes;        // 0: nothing to load
    long long col_off;     // byte offset of the first loaded byte in its line
};

template <int R, int BPP>
__device__ __forceinline__ CopyPlan copy_plan(const FastParams& P, const Item& it, uint32_t stage) {
    using G = Geo<R>;
    constexpr int WB = 32 * BPP;                          // bytes of an aligned 32-pixel word
    const long long limit = P.padded ? ((P.job.n1 * BPP + 15) / 16) * 16 : P.job.n1 * BPP;
    const long long w0 = it.xw0 * BPP;                    // byte offset of the window in the line
    long long lo = w0 + (R > 0 ? WB - 16 : 0);            // halo words: only the chunk next to the interior
    long long hi = w0 + (long long)G::AW * WB - (R > 0 ? WB - 16 : 0);
    lo = lo < 0 ? 0 : lo;
    hi = hi > limit ? limit : hi;
    CopyPlan c;
    c.src0 = it.px + lo;
    c.col_off = lo;
    c.dst = stage + (uint32_t)(lo - w0);
    c.bytes = hi > lo ? (uint32_t)(hi - lo) : 0u;
    return c;
}

// TMA: bring the packed RGB of line y into the warp's staging buffer, and pull the line PF lines
// further down into L2 (the next copy then completes at L2 latency instead of DRAM latency).
// Returns false if the line needs no load (outside the sheet).  ylo / yhi: loadable lines,
// pflo / pfhi: lines whose prefetch target lies in this GPU's memory (halo lines that are read in
// place from a peer GPU bypass the local L2).  Every argument is warp-uniform: the address
// arithmetic runs on the uniform datapath and the copy needs no operand broadcast.
constexpr int PF_LINES = 3;
// global address of the loaded part of line y (halo lines of a band may live in the neighbour band's
// buffer, i.e. in a peer GPU's memory: they are read in place)
__device__ __forceinline__ const uint8_t* line_address(const FastParams& P, const CopyPlan& c, int y) {
    const uint8_t* base = c.src0;
    long long row = y;
    if (y < 0 && P.job.halo_lo_px) {
        base = P.job.halo_lo_px + c.col_off;
        row = y + P.job.halo_lo;
    } else if (y >= P.job.n2 && P.job.halo_hi_px) {
        base = P.job.halo_hi_px + c.col_off;
        row = y - P.job.n2;
    }
    return base + row * P.job.stride2;
}
// `cur` walks down the lines of the item: the copies are issued for consecutive y, so the address is the
// previous one plus the line stride except at the first loadable line and where a band meets its halo
// lines (the full 64-bit address arithmetic once per line was 5 % of the kernel's instructions).
__device__ __forceinline__ bool issue_line_copy(const FastParams& P, const CopyPlan& c, int y, int ylo, int yhi,
                                                int pflo, int pfhi, uint32_t bar, int lane, const uint8_t*& cur) {
    if (y < ylo || y >= yhi) return false;
#ifdef MR_KO_TMA   // timing experiment: no input traffic
    return false;
#endif
    if (y == ylo || y == 0 || y == (int)P.job.n2) cur = line_address(P, c, y);
    else cur += P.job.stride2;
    const bool pf = y >= pflo && y + PF_LINES < pfhi;
    if (lane == 0) {
        mbar_expect_tx(bar, c.bytes);
        tma_load_1d(c.dst, cur, c.bytes, bar);
        if (pf)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(cur + (long long)PF_LINES * P.job.stride2), "r"(c.bytes)
                         : "memory");
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Work lists hold one entry per 32-pixel word (code + pixel mask).  The rare per-pixel jobs are
// spread evenly over the lanes: 32 entries at a time, inclusive scan of their pixel counts, every
// lane then finds "its" pixel (entry by binary search over the scan, bit by skipping set bits).
// f(code, bit) is called once per (entry, set bit).  Whole warp, converged.
template <typename F>
__device__ __forceinline__ void for_each_pixel(const uint16_t* codes, const uint32_t* masks, int n, int lane, F f) {
#pragma unroll 1
    for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        const uint32_t my_mask = e < n ? masks[e] : 0u;
        const int my_code = e < n ? codes[e] : 0;
        const int c = __popc(my_mask);
        int S = c;   // inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, S, o);
            if (lane >= o) S += v;
        }
        const int T = __shfl_sync(0xffffffffu, S, 31);
#pragma unroll 1
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            int idx = 0;   // number of entries whose inclusive scan is <= t  (= entry that owns task t)
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, S, (idx + step - 1) & 31);
                if (v <= t) idx += step;
            }
            idx &= 31;
            const int before = __shfl_sync(0xffffffffu, S - c, idx);
            uint32_t m = __shfl_sync(0xffffffffu, my_mask, idx);
            const int code = __shfl_sync(0xffffffffu, my_code, idx);
            if (t < T) {
                for (int r = t - before; r > 0; --r) m &= m - 1;   // skip r set bits
                f(code, __ffs(m) - 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Labels in (standalone clean, `clean(labels; hood)`): the staging line holds label bytes instead of
// packed RGB; lane = one aligned 32-pixel word = 32 bytes.  The bytes are brought into the layout
// of the bit transpose (pixel 8k+j in byte k of register j), transposed into NB planes and stored
// to the ring exactly like the categorised labels of the fused path.
template <int R, int NB, typename F>
__device__ __forceinline__ void labels_line(uint8_t* ws, const Item& it, int y, int lane, const LaneMasks& M,
                                            const uint8_t* staged, F after_read) {
    using G = Geo<R>;
    using C = Cfg<R, NB>;
    const int slot = (y - (it.y0 - R)) % C::RL;
    uint32_t* pl = reinterpret_cast<uint32_t*>(ws + C::W_PLANES) + slot * (ring_line(NB) / 4);
    if (!staged) {   // outside the sheet: VOID everywhere (warp-uniform)
        after_read();
#pragma unroll
        for (int b = 0; b < NB; ++b) pl[b * 32 + lane] = ~0u;
        return;
    }
    const uint4 v0 = *reinterpret_cast<const uint4*>(staged + lane * 32);
    const uint4 v1 = *reinterpret_cast<const uint4*>(staged + lane * 32 + 16);
    after_read();
    const uint32_t nat[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};   // pixel 4n+k in byte k of nat[n]
    uint32_t T[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // pixel 8k+j = byte (j & 3) of nat[2k + (j >> 2)]
        const uint32_t sel = (uint32_t)(j & 3) | ((uint32_t)(4 + (j & 3)) << 4);
        const uint32_t lo = __byte_perm(nat[j >> 2], nat[2 + (j >> 2)], sel);
        const uint32_t hi = __byte_perm(nat[4 + (j >> 2)], nat[6 + (j >> 2)], sel);
        T[j] = __byte_perm(lo, hi, 0x5410);
    }
    if (it.edge) {   // warp-uniform: force VOID outside [0, n1)
#pragma unroll
        for (int j = 0; j < 8; ++j) T[j] |= ((~M.vmask >> j) & 0x01010101u) * 0xFFu;
    }
#pragma unroll
    for (int st = 0; st < 3; ++st) {   // bit transpose, see categorize_line
        const int d = 1 << st;
        const uint32_t m = st == 0 ? 0x55555555u : (st == 1 ? 0x33333333u : 0x0F0F0F0Fu);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & d) continue;
            const uint32_t x = T[j], y2 = T[j + d];
            T[j] = (x & m) | ((y2 << d) & ~m);
            T[j + d] = ((x >> d) & m) | (y2 & ~m);
        }
    }
    const int o = G::O0 + lane * G::U;
    const int k = o >> 5, sh = o & 31;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t lo = __shfl_sync(0xffffffffu, T[b], k);
        const uint32_t hi = __shfl_sync(0xffffffffu, T[b], (k + 1) & 31);
        pl[b * 32 + lane] = __funnelshift_r(lo, hi, sh);
    }
}

// ---------------------------------------------------------------------------------------------
// categorise one line of the strip window (whole warp) from the TMA staging buffer
// after_read() is called as soon as the staging buffer has been consumed (the caller uses it to
// issue the TMA copy of the next line, so that the copy overlaps the rest of this line's work).
template <int R, int NB, bool COUNT, typename F>
__device__ __forceinline__ void categorize_line(const FastParams& P, const uint8_t* coarse, uint8_t* ws,
                                                const Item& it, int y, int lane, const LaneMasks& M,
                                                const uint8_t* staged, F after_read, unsigned& c_fine,
                                                unsigned& c_nomatch) {
    using G = Geo<R>;
    using C = Cfg<R, NB>;
    const int slot = (y - (it.y0 - R)) % C::RL;
    uint32_t* pl = reinterpret_cast<uint32_t*>(ws + C::W_PLANES) + slot * (ring_line(NB) / 4);
    if (!staged) {   // outside the sheet: VOID everywhere (warp-uniform)
        after_read();
        if (R > 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) pl[b * 32 + lane] = ~0u;
        }
        return;
    }
    // labels, packed so that register j holds pixels {j, 8+j, 16+j, 24+j} in bytes 0..3 (the layout
    // the bit transpose below wants).  Two rounds of 16 pixels (48 bytes = three 16-byte reads);
    // unloaded parts of halo words give garbage that is never used.
    // Cell index r5 | g5<<5 | b5<<10 of a pixel = byte dot product (IDP.4A, off the ALU pipe) of the
    // quantised words with the weights (1, 32, 128): r and g bytes are quantised to x>>3, b bytes to
    // x & 0xF8 (= 8*b5), the byte masks depend only on the position of the word in its 12-byte group.
    const uint32_t tab = smem_u32(coarse);   // the coarse table (offset 0 of the dynamic shared memory)
    uint32_t Wp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Wp[j] = 0u;
#ifdef MR_KO_STAGECONF   // timing experiment: conflict-free staging reads (wrong bytes)
    const uint8_t* mine = staged + lane * 16;
#else
    const uint8_t* mine = staged + lane * 96;
#endif
    // rounds run from the last 16 pixels to the first so that "Wp = Wp * 256 + label" (one IMAD on
    // the FMA pipe) leaves pixels 8k+j in byte k of register j.
#ifdef MR_KO_CAT   // timing experiment: no categorise arithmetic (labels from the line number)
#pragma unroll
    for (int j = 0; j < 8; ++j) Wp[j] = 0x01010101u * (uint32_t)(((y + lane) >> 4) & 7);
    for (int h = 1; h >= 2; --h) {
#else
    // 16-byte reads at a lane stride of 96 bytes: lanes l and l + 4 of a quarter warp would hit the
    // same banks.  Lanes with bit 2 set therefore read the two 48-byte halves of their word in the
    // opposite order (bank offset 3 chunks = odd: all eight lanes of a quarter warp fall into
    // different bank groups) and swap the halves of their label registers back below.
    const int hswap = (lane >> 2) & 1;
#pragma unroll
    for (int h = 1; h >= 0; --h) {
#endif
        const uint8_t* half = mine + 48 * (h ^ hswap);
        const uint4 v0 = *reinterpret_cast<const uint4*>(half);
        const uint4 v1 = *reinterpret_cast<const uint4*>(half + 16);
        const uint4 v2 = *reinterpret_cast<const uint4*>(half + 32);
        const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int g = 3; g >= 0; --g) {   // 4 pixels = 12 bytes: [r0 g0 b0 r1] [g1 b1 r2 g2] [b2 r3 g3 b3]
            const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
            // every byte quantised to x >> 3 (two ALU operations per word); the r / g part of the
            // index is one byte dot product starting at the table address, the b part (weight 129)
            // a second one that an IMAD scales by 8: cell = r5 + 32 g5 + 1032 b5 (MR_COARSE_IDX)
            const uint32_t q0 = (w0 >> 3) & 0x1F1F1F1Fu, q1 = (w1 >> 3) & 0x1F1F1F1Fu, q2 = (w2 >> 3) & 0x1F1F1F1Fu;
            const uint32_t i0 = __dp4a(q0, 0x00810000u, 0u) * 8u + __dp4a(q0, 0x00002001u, tab);
            const uint32_t i1 = __dp4a(q1, 0x00008100u, 0u) * 8u + __dp4a(q1, 0x00000020u, __dp4a(q0, 0x01000000u, tab));
            const uint32_t i2 = __dp4a(q2, 0x00000081u, 0u) * 8u + __dp4a(q1, 0x20010000u, tab);
            const uint32_t i3 = __dp4a(q2, 0x81000000u, 0u) * 8u + __dp4a(q2, 0x00200100u, tab);
            // pixel 16h + 4g + t -> register (4g + t) & 7, byte order by descending pixel index
            const int j = (4 * g) & 7;
            Wp[j + 0] = Wp[j + 0] * 256u + lds_u8(i0);
            Wp[j + 1] = Wp[j + 1] * 256u + lds_u8(i1);
            Wp[j + 2] = Wp[j + 2] * 256u + lds_u8(i2);
            Wp[j + 3] = Wp[j + 3] * 256u + lds_u8(i3);
        }
    }
    {   // lanes that read the halves in the opposite order: pixels 0..15 sit in bytes 2, 3 -> swap
        const uint32_t sel = hswap ? 0x1032u : 0x3210u;
#pragma unroll
        for (int j = 0; j < 8; ++j) Wp[j] = __byte_perm(Wp[j], 0u, sel);
    }
    if (it.edge) {   // warp-uniform: force VOID outside [0, n1)
        uint32_t nv = ~M.vmask;
        asm volatile("" : "+r"(nv));   // keeps the eight masks out of registers and the branch a real one
#pragma unroll
        for (int j = 0; j < 8; ++j) Wp[j] |= ((nv >> j) & 0x01010101u) * 0xFFu;
    }
    const bool own_line = (y >= it.y0) && (y < it.y1);
    const uint32_t ownmask = own_line ? M.avalid : 0u;
    // ---- MIXED-cell pixels (real, loaded pixels only; ~1 %) get their exact label from the
    // 2^24-entry table (L2 resident).  Bit 7 of a label byte is set for MIXED (and VOID) only.
    // The colours of the first two MIXED pixels of every word are captured from the staging line
    // right away, so that the line can be released (next TMA copy) and the table loads fly while
    // the labels are transposed; words with more than two (rare) keep the line a little longer.
    uint32_t s = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) s |= (Wp[j] >> (7 - j)) & (0x01010101u << j);
    s &= M.smask;
    auto colour = [&](int q) {
        const uint8_t* c3 = mine + 3 * q;
        return (unsigned)c3[0] | ((unsigned)c3[1] << 8) | ((unsigned)c3[2] << 16);
    };
    // bq = the isolated bit of the pixel (0: the word has no such pixel).  Only the lanes that have
    // one read its colour: these byte reads sit 96 bytes apart between lanes, i.e. up to eight lanes
    // per bank -- issued by all 32 lanes they were a quarter of the kernel's shared-memory wavefronts.
    // MIXED pixels per word that are looked up without holding the line: three for small palettes (a word with
    // more occurs on 8 % of the lines of config 2, with two it was 37 %); big palettes have so many MIXED
    // pixels that the wide loop below runs on most lines anyway (three: 4 % slower on config 4)
#ifdef MR_NQ
    constexpr int NQ = MR_NQ;
#else
    constexpr int NQ = NB >= 6 ? 2 : 3;
#endif
    uint32_t bq[NQ];
    int qi[NQ];
    unsigned rgbq[NQ], lq[NQ];
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
        bq[t] = s & (0u - s);
        s ^= bq[t];
        qi[t] = 31 - __clz(bq[t]);
        rgbq[t] = 0u;
        if (bq[t]) rgbq[t] = colour(qi[t]);
    }
    const bool more = __any_sync(0xffffffffu, s != 0u);
    if (!more) after_read();
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
        lq[t] = 0u;
        if (bq[t]) lq[t] = __ldg(P.job.lut + rgbq[t]);
    }
    if constexpr (R == 0) {
        // ---- no clean: the categorised labels are the output
        uint8_t* orow = it.out + (long long)y * P.job.out_stride2 + it.xw0 + (long long)lane * 32;
        if (ownmask) {
            uint32_t nat[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) {   // natural order: pixel 4n+k in byte k of register n
                const int k = n >> 1, base = 4 * (n & 1);
                const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
                const uint32_t lo = __byte_perm(Wp[base + 0], Wp[base + 1], sel);
                const uint32_t hi = __byte_perm(Wp[base + 2], Wp[base + 3], sel);
                nat[n] = __byte_perm(lo, hi, 0x5410);
            }
            if (COUNT) {
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const uint32_t nz = (nat[n] | (nat[n] >> 4)) & 0x0F0F0F0Fu;
                    const uint32_t nz2 = (nz | (nz >> 1) | (nz >> 2) | (nz >> 3)) & 0x01010101u;   // byte != 0
                    const uint32_t v4 = (ownmask >> (4 * n)) & 0xFu;
                    const uint32_t vm = (v4 & 1u) | ((v4 & 2u) << 7) | ((v4 & 4u) << 14) | ((v4 & 8u) << 21);
                    c_nomatch += __popc(vm & ~nz2);
                }
            }
            // mr_fast_supported guarantees n1 % 16 == 0 or padded lines: whole 16-byte chunks only
            *reinterpret_cast<uint4*>(orow) = make_uint4(nat[0], nat[1], nat[2], nat[3]);
            if (it.xw0 + (long long)lane * 32 + 16 < P.job.n1)
                *reinterpret_cast<uint4*>(orow + 16) = make_uint4(nat[4], nat[5], nat[6], nat[7]);
        }
        auto patch = [&](int q, unsigned l) {
            if ((ownmask >> q) & 1u) {
                orow[q] = (uint8_t)l;
                ++c_fine;
                if (COUNT) c_nomatch += (l == 0);
            }
        };
        if (more) {
#pragma unroll 1
            while (__any_sync(0xffffffffu, s != 0u)) {
                if (s) {
                    const int q = __ffs(s) - 1;
                    s &= s - 1;
                    patch(q, __ldg(P.job.lut + colour(q)));
                }
            }
            after_read();
        }
#pragma unroll
        for (int t = 0; t < NQ; ++t)
            if (bq[t]) patch(qi[t], lq[t]);
    } else {
        // ---- bit transpose (three delta-swap stages between register pairs): T[b] bit q = bit b of
        // the label of pixel q of aligned word `lane`
        uint32_t T[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) T[j] = Wp[j];
#ifndef MR_KO_TRANSPOSE
#pragma unroll
        for (int st = 0; st < 3; ++st) {
            const int d = 1 << st;
            const uint32_t m = st == 0 ? 0x55555555u : (st == 1 ? 0x33333333u : 0x0F0F0F0Fu);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j & d) continue;
                const uint32_t x = T[j], y2 = T[j + d];
                T[j] = (x & m) | ((y2 << d) & ~m);
                T[j + d] = ((x >> d) & m) | (y2 & ~m);
            }
        }
#endif
        auto apply = [&](uint32_t bitq, unsigned l) {   // MIXED = all ones: clear the zero bits of the label
            const uint32_t keep = ~bitq;                 // bitq == 0 (no such pixel): nothing changes
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                    "and.b32 t, %2, %3;\n\t"
                    "setp.eq.u32 p, t, 0;\n\t"
                    "@p and.b32 %0, %0, %1;\n\t}"
                    : "+r"(T[b]) : "r"(keep), "r"(l), "r"(1u << b));
            }
            if (COUNT) c_fine += (ownmask & bitq) != 0u;
        };
        if (more) {
#pragma unroll 1
            // big palettes (K > 30) have many MIXED cells (5.7 % of the pixels on config 4): four pixels
            // per lane and pass, their table loads in flight together; small palettes keep the code short
            constexpr int WIDE = NB >= 6 ? 4 : 1;
            while (__any_sync(0xffffffffu, s != 0u)) {
                int qq[WIDE];
                unsigned ll[WIDE];
                bool hh[WIDE];
#pragma unroll
                for (int t = 0; t < WIDE; ++t) {
                    hh[t] = s != 0u;
                    qq[t] = 0;
                    ll[t] = 0;
                    if (hh[t]) {
                        qq[t] = __ffs(s) - 1;
                        s &= s - 1;
                        ll[t] = __ldg(P.job.lut + colour(qq[t]));
                    }
                }
#pragma unroll
                for (int t = 0; t < WIDE; ++t)
                    if (hh[t]) apply(1u << qq[t], ll[t]);
            }
            after_read();
        }
#pragma unroll
        for (int t = 0; t < NQ; ++t) apply(bq[t], lq[t]);
        // vote word `lane` starts at wx = O0 + lane*U: funnel of aligned words k, k+1
        const int o = G::O0 + lane * G::U;
        const int k = o >> 5, sh = o & 31;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t lo = __shfl_sync(0xffffffffu, T[b], k);
            const uint32_t hi = __shfl_sync(0xffffffffu, T[b], (k + 1) & 31);
            pl[b * 32 + lane] = __funnelshift_r(lo, hi, sh);
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

// plane words of lines y-R .. y+R of vote word vl from the warp's ring
template <int R, int NB>
__device__ __forceinline__ void load_planes(const uint8_t* ws, const Item& it, int y, int vl,
                                            uint32_t (&Pl)[2 * R + 1][NB]) {
    using C = Cfg<R, NB>;
    constexpr int LINE = ring_line(NB), RING = C::RL * LINE;
    const uint8_t* base = ws + C::W_PLANES + vl * 4;
    int off = ((y - it.y0) % C::RL) * LINE;   // slot of line y - R
#pragma unroll
    for (int d = 0; d < 2 * R + 1; ++d) {
#pragma unroll
        for (int b = 0; b < NB; ++b) Pl[d][b] = *reinterpret_cast<const uint32_t*>(base + off + b * 128);
        off += LINE;
        if (off == RING) off = 0;
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
// label in the window; stops as soon as no unseen label can still win or tie.  Compact code (rare
// path): the line of the next candidate is picked by an unrolled select chain.
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
#pragma unroll 1
    while (left > 0 && left >= bc) {
        // first line that still has uncounted pixels, and its plane words
        uint32_t r = 0u, Ls[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) Ls[b] = 0u;
#pragma unroll
        for (int d = SIDE - 1; d >= 0; --d) {
            if (rem[d]) {
                r = rem[d];
#pragma unroll
                for (int b = 0; b < NB; ++b) Ls[b] = Pl[d][b];
            }
        }
        const int i = __ffs(r) - 1;
        unsigned c = 0;
        uint32_t Lm[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t bit = (Ls[b] >> i) & 1u;
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
    return (unsigned)best;
}

// ---------------------------------------------------------------------------------------------
// vote one word.  mask: the pixels to decide (round 1: all output pixels of the word; round 2: the
// pixels round 1 left undecided).  Candidates are the labels at the last / first pixel of `mask`.
// Pixels still undecided are appended (word entry: code + mask) to the next list.
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void vote_word(const FastParams& P, uint8_t* ws, const Item& it, int y, int blk_a,
                                          int vl, uint32_t mask, unsigned& c_nomatch, unsigned& c_changed) {
    using G = Geo<R>;
    using C = Cfg<R, NB>;
    constexpr int SIDE = 2 * R + 1;
    constexpr unsigned VOIDL = (1u << NB) - 1u;
    uint32_t Pl[SIDE][NB];
    load_planes<R, NB>(ws, it, y, vl, Pl);
    // ---- candidate labels (heuristic only): last / first pixel of the mask; if they agree, B comes
    // from the top / bottom line of the window instead
    const bool vedge = (y - 1 < -P.job.halo_lo) || (y + 1 >= P.job.n2 + P.job.halo_hi);   // warp-uniform
    const int posA = 31 - __clz(mask), posB = __ffs(mask) - 1;
    const int posM = ((mask >> 16) & 1u) ? 16 : posA;
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
    uint32_t now = decided & mask;
    const uint32_t undec = ~decided & mask;
#ifndef MR_STATS
    if (COUNT) {
        c_changed += __popc(now & diff);
        c_nomatch += __popc(now & zero);
    }
#endif
    uint8_t* orow = it.out + (long long)y * P.job.out_stride2 + it.xw0 + G::O0 + vl * G::U;
#pragma unroll 1
    while (now) {   // round 1 stored a guess for these pixels: overwrite it with the decision
        const int q = __ffs(now) - 1;
        now &= now - 1;
        orow[q] = (uint8_t)(((selA >> q) & 1u) ? a_lab : b_lab);
    }
    if (undec) {   // -> exact-vote list (in place: at most one entry per round-2 entry already read)
        int* cnt = reinterpret_cast<int*>(ws + C::W_CNT);
        const int e = atomicAdd(cnt + 2, 1);
        reinterpret_cast<uint16_t*>(ws + C::W_W2CODE)[e] = (uint16_t)(((y - blk_a) << 5) | vl);
        reinterpret_cast<uint32_t*>(ws + C::W_W2MASK)[e] = undec;
    }
}

// ---------------------------------------------------------------------------------------------
// Round 1 of the vote for a whole block of nl output lines [a, a + nl) of vote word `lane`.
// The two candidate labels are chosen ONCE per (vote word, block), so every line's match masks
// and their horizontal window sums are computed once and shared by the 2R+1 output lines that
// see the line: the box counts slide down the block (add the line entering the window, subtract
// the line leaving it).  Decision test and tie rule as in vote_core.  Words with undecided pixels
// go to the round-2 list (candidates per word there).
//
// Output: the labels of every output line are final here except for the undecided pixels (which get
// a guess that round 2 / the exact vote overwrite): each vote word writes its U bytes into the warp's
// line buffer and one TMA bulk store (shared -> global, 32*U contiguous bytes) sends the line out.
// Returns the number of lines completed: the block is cut short if the round-2 list could overflow
// (the caller votes the rest after emptying the list).
template <int R, int NB, bool COUNT>
__device__ __forceinline__ int block_vote(const FastParams& P, uint8_t* ws, const Item& it, int a,
                                           int nl, int vl, uint32_t mask, uint32_t obytes, unsigned& hint,
                                           unsigned& c_nomatch, unsigned& c_changed, int& n_w2) {
    using G = Geo<R>;
    using C = Cfg<R, NB>;
    constexpr int SIDE = 2 * R + 1;
    constexpr int N = SIDE * SIDE;
    constexpr int WH = R == 1 ? 2 : 3, WC = G::WC;
    constexpr unsigned VOIDL = (1u << NB) - 1u;
    constexpr int LINE = ring_line(NB), RING = C::RL * LINE;
    const uint8_t* base = ws + C::W_PLANES + vl * 4;
    // ---- candidates (heuristic only): A = label at the centre of the block, B = the label farthest
    // from the centre among the pixels of the centre line that differ from A; if the centre line
    // shows only A, the other label this vote word saw in the blocks above (`hint`: the previous B,
    // or the previous A when A changed).  Labels are read through a vertical bitwise majority of
    // three lines, which hides isolated speckle.
    unsigned a_lab = 0, b_lab = 0;
    {
        const int yy = a + (nl >> 1);
        int off = ((yy - 1 - (it.y0 - R)) % C::RL) * LINE;
        uint32_t Q[NB];
        {
            uint32_t X[3][NB];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
#pragma unroll
                for (int b = 0; b < NB; ++b) X[d][b] = *reinterpret_cast<const uint32_t*>(base + off + b * 128);
                off += LINE;
                if (off == RING) off = 0;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) Q[b] = mr_maj(X[0][b], X[1][b], X[2][b]);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) a_lab |= ((Q[b] >> 16) & 1u) << b;
        uint32_t nota = 0u, isvoid = ~0u;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            nota |= Q[b] ^ (0u - ((a_lab >> b) & 1u));
            isvoid &= Q[b];
        }
        nota &= mask & ~isvoid;
        const unsigned prev_a = hint >> 8, prev_b = hint & 0xFFu;
        b_lab = (prev_a != a_lab) ? prev_a : prev_b;     // no second label on the centre line: the one from above
        if (nota) {
            const int pf = __ffs(nota) - 1, pl = 31 - __clz(nota);
            const int pos = (pl - 16 > 16 - pf) ? pl : pf;
            b_lab = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) b_lab |= ((Q[b] >> pos) & 1u) << b;
        }
        hint = (a_lab << 8) | b_lab;
    }
    const uint32_t nbm = (b_lab != a_lab && b_lab != VOIDL) ? ~0u : 0u;   // B == A -> B's masks are empty
    const uint32_t avm = (a_lab != VOIDL) ? ~0u : 0u;
    const uint32_t a_lower = (a_lab < b_lab) ? ~0u : 0u;
    uint32_t LA[NB], LB[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        LA[b] = 0u - ((a_lab >> b) & 1u);
        LB[b] = 0u - ((b_lab >> b) & 1u);
    }
    // ---- slide down the block.  Counted per pixel: cA = window pixels equal to A, cU = window pixels
    // equal to A or B (the masks are disjoint).  The better of {A, B} is certainly the mode when
    // cA >= N/2 + 1 (strict majority), or when cU >= 2N/3 + 1: then max(cA, cB) >= ceil(cU / 2) exceeds
    // the N - cU pixels every other label could have.  Winner: A iff 2 cA > cU, ties -> lower label.
    constexpr unsigned TA = N / 2 + 1, TU = 2 * N / 3 + 1;
    Bs<WC> SA, SU;
#pragma unroll
    for (int k = 0; k < WC; ++k) SA.b[k] = SU.b[k] = 0u;
    Bs<WH> HA[SIDE], HU[SIDE];        // horizontal window sums of the last 2R+1 lines (oldest first)
    uint32_t hA[R + 1], hB[R + 1];    // COUNT: match masks of the last R+1 lines
#pragma unroll
    for (int d = 0; d < SIDE; ++d) {
#pragma unroll
        for (int k = 0; k < WH; ++k) HA[d].b[k] = HU[d].b[k] = 0u;
    }
#pragma unroll
    for (int d = 0; d <= R; ++d) hA[d] = hB[d] = 0u;
    int off = ((a - it.y0) % C::RL) * LINE;   // slot of line a - R
    int n_listed = 0;   // entries this call has put on the round-2 list (warp-uniform)
    // first byte of output line a of this strip (warp-uniform: the bulk store takes uniform operands)
    const uint8_t* gline = it.out + (long long)a * P.job.out_stride2 + it.xw0 + 32 * G::HW;
    uint8_t* obuf = ws + C::W_OUT + vl * G::U;
    const uint32_t obuf0 = smem_u32(ws + C::W_OUT);
    const uint32_t B4 = b_lab * 0x01010101u;
    const int n_it = nl + 2 * R;
#pragma unroll 1
    for (int i = 0; i < n_it; ++i) {
        // ring of horizontal sums: line i goes to slot DN, line i - 2R (leaving the window) sits in slot DO.
        // (Unrolling the loop by SIDE so that the slots become compile-time indices without this rotation
        // was measured 20 % slower: five copies of the body no longer fit the instruction cache.)
        constexpr int DN = SIDE - 1, DO = 0;
#pragma unroll
        for (int d = 0; d < SIDE - 1; ++d) {
            HA[d] = HA[d + 1];
            HU[d] = HU[d + 1];
        }
        uint32_t ma = avm, mb = nbm;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t p = *reinterpret_cast<const uint32_t*>(base + off + b * 128);
            ma &= ~(p ^ LA[b]);
            mb &= ~(p ^ LB[b]);
        }
        off += LINE;
        if (off == RING) off = 0;
        HA[DN] = bs_hsum1<R>(ma);
        HU[DN] = bs_hsum1<R>(ma | mb);
        SA = bs_add<WC, WH, WC>(SA, HA[DN]);
        SU = bs_add<WC, WH, WC>(SU, HU[DN]);
        if (COUNT) {   // the centre line's own match masks (line i - R) tell whether its label changed
#pragma unroll
            for (int d = 0; d < R; ++d) {
                hA[d] = hA[d + 1];
                hB[d] = hB[d + 1];
            }
            hA[R] = ma;
            hB[R] = mb;
        }
        if (i >= 2 * R) {
            // ---- decide output line a + i - 2R
            const uint32_t decided = bs_gt_const<WC>(SU, TU - 1u) | bs_gt_const<WC>(SA, TA - 1u);
            uint32_t gt = SA.b[WC - 1], eq = ~SA.b[WC - 1];   // 2 cA (WC + 1 bits) against cU
#pragma unroll
            for (int k = WC - 1; k >= 1; --k) {
                gt |= eq & SA.b[k - 1] & ~SU.b[k];
                eq &= ~(SA.b[k - 1] ^ SU.b[k]);
            }
            eq &= ~SU.b[0];
            const uint32_t selA = gt | (eq & a_lower);
            const uint32_t now = decided & mask;
            const uint32_t undec = ~decided & mask;
            // ---- the line buffer is free once the previous line's bulk store has read it
#ifndef MR_KO_STORE
            if (vl == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
            __syncwarp();
            if (n_listed > LIST_CAP - 32) {   // warp-uniform
                n_w2 = n_listed;
                return i - 2 * R;
            }
            // list slot of this word: entries so far + the undecided words of the lanes below (no atomic)
            const uint32_t bal = __ballot_sync(0xffffffffu, undec != 0u);
            const int e = n_listed + __popc(bal & ((1u << vl) - 1u));
            n_listed += __popc(bal);
#ifdef MR_STATS   // experiment: n_nomatch = pixels round 1 leaves undecided, n_changed = listed words, n_fine = exact pixels
            if (COUNT) {
                c_nomatch += __popc(undec);
                c_changed += undec != 0u;
            }
#else
            if (COUNT) {
                c_changed += __popc(now & ~((selA & hA[0]) | (~selA & hB[0])));   // centre label != winner
                c_nomatch += __popc(now & ((a_lab == 0 ? selA : 0u) | (b_lab == 0 ? ~selA : 0u)));
            }
#endif
            {   // labels of the U output pixels, B ^ (selA ? A ^ B : 0): eight pixels per step through the
                // bit -> byte expansion table (one 8-byte entry per byte of selA)
                const uint32_t sel = selA >> R;
                constexpr int GP = (G::U % 4 == 0) ? 4 : 2;
#pragma unroll
                for (int k = 0; k < G::U / GP; ++k) {
                    const uint32_t nib = (sel >> (GP * k)) & ((1u << GP) - 1u);
                    const uint32_t m = (nib * 0x00204081u) & 0x01010101u;   // bit j -> byte j
                    // byte j = m_j ? A : B without leaving the multiply-add pipe: (B4 - m B) has 0 or B in
                    // every byte (no borrow), adding m A then puts A into the zeroed bytes (no carry)
                    const uint32_t v = m * a_lab + (B4 - m * b_lab);
                    if (GP == 4) *reinterpret_cast<uint32_t*>(obuf + 4 * k) = v;
                    else *reinterpret_cast<uint16_t*>(obuf + 2 * k) = (uint16_t)v;
                }
            }
#ifndef MR_KO_FENCE
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
            __syncwarp();
#ifdef MR_KO_STORE
            if (false) {
#else
            if (vl == 0 && obytes) {
#endif
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gline), "r"(obuf0),
                             "r"(obytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            gline += P.job.out_stride2;
            if (undec) {
                reinterpret_cast<uint16_t*>(ws + C::W_W2CODE)[e] = (uint16_t)(((i - 2 * R) << 5) | vl);
                reinterpret_cast<uint32_t*>(ws + C::W_W2MASK)[e] = undec;
            }
            SA = bs_sub<WC, WH>(SA, HA[DO]);
            SU = bs_sub<WC, WH>(SU, HU[DO]);
        }
    }
    n_w2 = n_listed;
    return nl;
}

// exact vote of one still undecided pixel.  code = line-in-block << 5 | vote word, q = bit
template <int R, int NB, bool COUNT>
__device__ __forceinline__ void exact_pixel(const FastParams& P, uint8_t* ws, const Item& it, int blk_a, int code,
                                            int q, unsigned& c_nomatch, unsigned& c_changed) {
    using G = Geo<R>;
    const int y = blk_a + (code >> 5), vl = code & 31;
    uint32_t Pl[2 * R + 1][NB];
    load_planes<R, NB>(ws, it, y, vl, Pl);
    const unsigned best = exact_vote<R, NB>(Pl, q);
    unsigned centre = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) centre |= ((Pl[R][b] >> q) & 1u) << b;
    it.out[(long long)y * P.job.out_stride2 + it.xw0 + G::O0 + vl * G::U + q] = (uint8_t)best;   // over round 1's guess
#ifdef MR_STATS
    if (COUNT) c_changed += 0x10000u;
#else
    if (COUNT) {
        c_nomatch += (best == 0);
        c_changed += (best != centre);
    }
#endif
}

// ---------------------------------------------------------------------------------------------
template <int R, int NB>
__device__ __forceinline__ Item decode_item(const FastParams& P, unsigned item) {
    using G = Geo<R>;
    Item it;
    int c = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (k < P.n_classes && item >= P.cls[k].first_item) c = k;
    const SchedClass cl = P.cls[c];
    const unsigned rem = item - cl.first_item;
    const int chunk = (int)(rem / (unsigned)P.cols), col = (int)(rem - (unsigned)chunk * (unsigned)P.cols);   // neighbouring strips first
    const int sheet = col / P.n_strips, strip = col - sheet * P.n_strips;
    it.px = P.job.px + (long long)sheet * P.job.sheet_stride;
    it.out = P.job.out + (long long)sheet * P.job.out_sheet_stride;
    it.xw0 = (long long)strip * G::SW - 32 * G::HW;
    it.y0 = cl.line0 + chunk * cl.chunk;
    it.y1 = min(it.y0 + cl.chunk, cl.line1);
    it.edge = (strip == 0) || (it.xw0 + 32 * G::AW > P.job.n1);
    return it;
}

// LIN: the input is a label image (bpp = 1) instead of packed RGB: standalone clean
template <int R, int NB, bool COUNT, bool LIN>
__global__ void __launch_bounds__(Cfg<R, NB>::WARPS * 32, 1) k_fast(const FastParams P) {
    using C = Cfg<R, NB>;
    extern __shared__ __align__(128) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: the TMA operands live in uniform registers
    uint8_t* ws = sm + C::SLICES + warp * C::W_BYTES;         // this warp's private slice
    const uint8_t* coarse = sm;
    int* cnt = reinterpret_cast<int*>(ws + C::W_CNT);
    const uint32_t bar = smem_u32(ws + C::W_MBAR);
    uint8_t* stage = ws + C::W_STAGE;
    const uint32_t stage_u32 = smem_u32(stage);
    if (!LIN) {
        const uint4* src = reinterpret_cast<const uint4*>(P.job.coarse);
        uint4* dst = reinterpret_cast<uint4*>(sm);
        for (int i = tid; i < (int)(MR_COARSE_SIZE / 16); i += C::WARPS * 32) dst[i] = __ldg(src + i);
    }
    if (lane == 0) {
        cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // the only block barrier: coarse table + barriers ready
    unsigned c_fine = 0, c_nomatch = 0, c_changed = 0;
    uint32_t parity = 0;

#ifdef MR_TIMELINE
    unsigned long long t_start, t_first = 0, n_lines_done = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
    if (blockIdx.x == 0 && tid == 0) *P.work_reset = 0u;
    unsigned next = 0;
    if (lane == 0) next = atomicAdd(P.work_counter, 1u);
    next = __shfl_sync(0xffffffffu, next, 0);
#pragma unroll 1
    while (next < (unsigned)P.n_items) {
        const Item it = decode_item<R, NB>(P, next);
        const LaneMasks M = lane_masks<R>(P, it.xw0, lane);
        unsigned hint = ((1u << NB) - 1u) * 0x101u;          // candidate memory of this vote word: none yet (VOID, VOID)
        const int cat_last = it.y1 + R;                       // lines [y0-R, y1+R) are categorised
        int ycat = it.y0 - R;                                 // next line to categorise
        const CopyPlan plan = copy_plan<R, (LIN ? 1 : 3)>(P, it, stage_u32);
        // loadable lines of this item (none if the strip window has no bytes inside the sheet)
        const int ylo = max(ycat, -P.job.halo_lo);
        const int yhi = plan.bytes ? min(cat_last, (int)P.job.n2 + P.job.halo_hi) : ylo;
        const int pflo = P.job.halo_lo_px ? max(ylo, 0) : ylo;
        const int pfhi = P.job.halo_hi_px ? min(yhi, (int)P.job.n2) : yhi;
        const uint8_t* cur_src = nullptr;   // address of the line copied last
        bool pend = issue_line_copy(P, plan, ycat, ylo, yhi, pflo, pfhi, bar, lane, cur_src);
        // bytes of one output line of the strip (whole 16-byte chunks, see mr_fast_supported)
        uint32_t obytes = 0;
        if (R > 0) {
            const long long olimit = P.padded ? ((P.job.n1 + 15) / 16) * 16 : P.job.n1;
            const long long ob = olimit - (it.xw0 + 32 * Geo<R>::HW);
            obytes = (uint32_t)(ob < 0 ? 0 : (ob > Geo<R>::SW ? Geo<R>::SW : ob));
        }
#pragma unroll 1
        for (int a = it.y0; a < it.y1; a += C::B) {
            const int vend = min(a + C::B, it.y1);
            if (vend == it.y1) {   // last block: fetch the following item now (earlier would hoard an item
                                   // that an idle warp could already be working on)
                // only lane 0 holds the answer until the item ends: nobody waits for the atomic's trip to L2
                if (lane == 0) next = atomicAdd(P.work_counter, 1u);
            }
            // ---- categorise lines [ycat, vend + R); each arrives by TMA, issued one line ahead
#pragma unroll 1
            for (; ycat < vend + R; ++ycat) {
                if (pend) {
                    mbar_wait(bar, parity);
                    parity ^= 1u;
                }
                bool nxt = false;
                auto release = [&]() {
                    __syncwarp();   // every lane is done with the staging buffer
                    nxt = issue_line_copy(P, plan, ycat + 1, ylo, yhi, pflo, pfhi, bar, lane, cur_src);
                };
                if constexpr (LIN)
                    labels_line<R, NB>(ws, it, ycat, lane, M, pend ? stage : nullptr, release);
                else
                    categorize_line<R, NB, COUNT>(P, coarse, ws, it, ycat, lane, M, pend ? stage : nullptr, release,
                                                  c_fine, c_nomatch);
                pend = nxt;
            }
            if constexpr (R > 0) {
                // ---- votes: round 1 per block (lane = vote word), then round 2 over the word list
                // round 1 produced, densely over the lanes
#pragma unroll 1
                for (int va = a; va < vend;) {
#ifdef MR_KO_VOTE   // timing experiment: no vote at all
                break;
#endif
#ifdef MR_TIMELINE
                if (t_first == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_first));
                n_lines_done += vend - va;
#endif
                int n2 = 0;   // entries round 1 put on the word list
                const int done = block_vote<R, NB, COUNT>(P, ws, it, va, vend - va, lane, M.vvalid, obytes, hint, c_nomatch,
                                                          c_changed, n2);
                __syncwarp();
#ifdef MR_KO_ROUND2
                n2 = 0;
#endif
                if (n2) {   // byte patches follow: the bulk stores of this block must have landed
#ifndef MR_KO_WAIT
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
                    __syncwarp();
                }
#pragma unroll 1
                for (int base = 0; base < n2; base += 32) {
                    const bool have = base + lane < n2;
                    int code = 0;
                    uint32_t mask = 0u;
                    if (have) {
                        code = reinterpret_cast<const uint16_t*>(ws + C::W_W2CODE)[base + lane];
                        mask = reinterpret_cast<const uint32_t*>(ws + C::W_W2MASK)[base + lane];
                    }
                    __syncwarp();   // entries are in registers: the exact-vote list may overwrite them
                    if (have) vote_word<R, NB, COUNT>(P, ws, it, va + (code >> 5), va, code & 31, mask, c_nomatch, c_changed);
                }
                __syncwarp();
                {   // ---- exact votes of what is still undecided, spread evenly over the lanes
#ifdef MR_KO_EXACT
                    const int n = 0;
#else
                    const int n = cnt[2];
#endif
                    for_each_pixel(reinterpret_cast<const uint16_t*>(ws + C::W_W2CODE),
                                   reinterpret_cast<const uint32_t*>(ws + C::W_W2MASK), n, lane, [&](int code, int q) {
                                       exact_pixel<R, NB, COUNT>(P, ws, it, va, code, q, c_nomatch, c_changed);
                                   });
                }
                __syncwarp();
                if (lane == 0) { cnt[1] = 0; cnt[2] = 0; }
                __syncwarp();
                va += done;
                }
            }
        }
        next = __shfl_sync(0xffffffffu, next, 0);   // the item fetched in the last block
    }
    if (R > 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // line buffer still in use
#ifdef MR_TIMELINE
    if (lane == 0 && P.timeline) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        unsigned long long* o = P.timeline + 4 * ((size_t)blockIdx.x * C::WARPS + warp);
        o[0] = t_start; o[1] = t_first; o[2] = t_end; o[3] = n_lines_done;
    }
#endif
    if (COUNT && P.job.counters) {
        for (int o = 16; o > 0; o >>= 1) {
            c_fine += __shfl_xor_sync(0xffffffffu, c_fine, o);
            c_nomatch += __shfl_xor_sync(0xffffffffu, c_nomatch, o);
            c_changed += __shfl_xor_sync(0xffffffffu, c_changed, o);
        }
#ifdef MR_STATS
        c_fine = c_changed >> 16;
        c_changed &= 0xFFFFu;
#endif
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
std::mutex g_cfg_mutex;   // cudaFuncSetAttribute bookkeeping below (contexts may live on several host threads)

// ---- host: the item schedule.  Class 0 hands out most of the lines (f0) in tall chunks whose item
// count fills whole "waves" of warps (an item count just above a multiple of the warp count would
// leave most of the GPU idle at the end of the class); the rest goes out in small chunks that level
// the finish: two-block items first, one-block items last.  Small jobs: one class of small chunks.
// (Measured on B200, scripts/sched_sweep*.sh: tall-class share 0.78 .. 0.96, one to four tail classes
// down to quarter blocks -- everything within 1.5 %; a per-warp timeline shows the warps ending within
// 50 us of each other whatever the job size, 5 % of the warp time on a 0.57 ms job.)
void plan_schedule(FastParams& P, int warps, int B, int R) {
    const long long n2 = P.job.n2;
    const int cols = P.cols;
    double f0 = 0.92;         // measured on B200: 0.86 .. 0.94 within 2 %
    long long tail = 2 * B;
#ifdef MR_SCHED_ENV   // schedule experiments: MAPRAST_F0 (share of the tall class), MAPRAST_TAIL (blocks per tail item)
    if (const char* e = getenv("MAPRAST_F0")) f0 = atof(e);
    if (const char* e = getenv("MAPRAST_TAIL")) tail = atoll(e) * B;
#endif
    const double share = (double)n2 * cols / warps;   // lines per warp
    long long H0 = 0, n0 = 0;
    if (share >= 6.0 * B) {
        double best = 0.0;
        long long hmax = (long long)(f0 * share) / B * B;
        if (hmax > 512) hmax = 512 / B * B;
        for (long long h = hmax; h >= 4 * B; h -= B) {
            const long long n = (long long)(f0 * n2) / h;
            if (n < 1) continue;
            const long long items = n * cols, waves = (items + warps - 1) / warps;
            const double fill = (double)items / (double)(waves * warps);
            const double halo = (double)h / ((double)h + 2 * R * 0.45);
            if (fill * halo > best) { best = fill * halo; H0 = h; n0 = n; }
        }
    }
    int k = 0;
    long long y = 0, items = 0;
    if (n0 > 0) {
        P.cls[k] = SchedClass{0, (int)(n0 * H0), (int)H0, 0u};
        y = n0 * H0;
        items = n0 * cols;
        ++k;
    }
#ifdef MR_SCHED_ENV   // MAPRAST_TAILS = "frac:lines,frac:lines": classes after the tall one (the last takes the rest)
    if (const char* e = getenv("MAPRAST_TAILS")) {
        if (n0 > 0) {
            const char* p = e;
            while (*p && k < 3 && y < n2) {
                const double fr = atof(p);
                while (*p && *p != ':') ++p;
                const long long h = *p ? atoll(p + 1) : tail;
                while (*p && *p != ',') ++p;
                const bool last = !*p;
                if (*p) ++p;
                long long y1 = last ? n2 : y + (long long)(fr * n2) / h * h;
                if (y1 > n2) y1 = n2;
                if (y1 <= y) continue;
                P.cls[k] = SchedClass{(int)y, (int)y1, (int)h, (unsigned)items};
                items += ((y1 - y + h - 1) / h) * cols;
                y = y1;
                ++k;
            }
        }
    }
#endif
    if (k == 1 && y < n2 && n2 - y > 4 * B) {   // behind the tall class: half of the rest in two-block items ...
        const long long h = tail;
        long long y1 = y + (n2 - y) / 2 / h * h;
        if (y1 > y) {
            P.cls[k] = SchedClass{(int)y, (int)y1, (int)h, (unsigned)items};
            items += ((y1 - y) / h) * cols;
            y = y1;
            ++k;
        }
        tail = B;                               // ... the last lines in one-block items (1 % on config 2)
    }
    if (y < n2) {
        long long h = tail;
        if (n0 == 0) {   // small job: about four items per warp, at least one block
            h = (long long)(share / 4.0) / B * B;
            if (h < B) h = B;
        }
        P.cls[k] = SchedClass{(int)y, (int)n2, (int)h, (unsigned)items};
        items += ((n2 - y + h - 1) / h) * cols;
        ++k;
    }
    P.n_classes = k;
    P.n_items = items;
}

template <int R, int NB, bool COUNT, bool LIN>
cudaError_t launch_t(FastParams& P, int sm_count, cudaStream_t s) {
    using C = Cfg<R, NB>;
    auto kern = k_fast<R, NB, COUNT, LIN>;
    constexpr size_t smem = C::TOTAL;
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
    {
        std::lock_guard<std::mutex> lock(g_cfg_mutex);
        if (!configured[dev]) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev] = true;
        }
    }
    // work items: strips x line chunks; the chunk height is chosen so that every warp of the GPU
    // gets several items (dynamic queue), but not so small that the 2R halo lines dominate
    const long long warps = (long long)sm_count * C::WARPS;
    P.n_strips = (int)((P.job.n1 + Geo<R>::SW - 1) / Geo<R>::SW);
    P.cols = P.n_strips * P.job.n_sheets;
    plan_schedule(P, (int)warps, C::B, R);
    if (P.n_items >= (1ll << 31) - (1 << 20)) return cudaErrorNotSupported;
    long long grid = sm_count;
    const long long need = (P.n_items + C::WARPS - 1) / C::WARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, C::WARPS * 32, smem, s>>>(P);
    return cudaGetLastError();
}

template <int R, int NB>
cudaError_t launch_c(FastParams& P, int sm_count, cudaStream_t s) {
    if constexpr (R > 0) {
        if (P.job.bpp == 1)
            return P.job.counters ? launch_t<R, NB, true, true>(P, sm_count, s) : launch_t<R, NB, false, true>(P, sm_count, s);
    }
    return P.job.counters ? launch_t<R, NB, true, false>(P, sm_count, s) : launch_t<R, NB, false, false>(P, sm_count, s);
}

template <int NB>
cudaError_t launch_r(FastParams& P, int sm_count, cudaStream_t s) {
    switch (P.job.R) {
        case 1: return launch_c<1, NB>(P, sm_count, s);
        case 2: return launch_c<2, NB>(P, sm_count, s);
        case 3: return launch_c<3, NB>(P, sm_count, s);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace

#ifdef MR_TIMELINE
static unsigned long long* g_timeline_host = nullptr;
extern "C" int mr_debug_timeline(unsigned long long* out, int n_entries) {   // n_entries x 4 values
    if (!g_timeline_host) return -1;
    cudaDeviceSynchronize();
    return (int)cudaMemcpy(out, g_timeline_host, (size_t)n_entries * 32, cudaMemcpyDeviceToHost);
}
#endif

// bpp == 1: label image in (standalone clean); job.pal.K then holds the largest label value
bool mr_fast_supported(const MrJob& job) {
    if (job.bpp == 1) {
        if (job.R < 1) return false;
    } else if (job.bpp != 3 || job.lut == nullptr || job.coarse == nullptr) {
        return false;
    }
    if (job.pal.K > 126 || job.R < 0 || job.R > 3) return false;
    if (job.n1 <= 0 || job.n2 <= 0 || job.n2 > (1 << 30) || job.n_sheets <= 0) return false;
    const bool a16 = ((uintptr_t)job.px % 16 == 0) && ((uintptr_t)job.out % 16 == 0) && (job.stride2 % 16 == 0) &&
                     (job.out_stride2 % 16 == 0) && (job.sheet_stride % 16 == 0) && (job.out_sheet_stride % 16 == 0);
    if (!a16) return false;
    if (!job.padded && (job.n1 % 16 != 0)) return false;
    if ((job.halo_lo_px || job.halo_hi_px) &&
        (job.n_sheets != 1 || (uintptr_t)job.halo_lo_px % 16 != 0 || (uintptr_t)job.halo_hi_px % 16 != 0)) return false;
    return true;
}

cudaError_t mr_launch_fast(const MrJob& job, int sm_count, cudaStream_t s, int* n_launches, MrWorkRing* ring) {
    if (!mr_fast_supported(job)) return cudaErrorNotSupported;
    if (!ring || !ring->slots) return cudaErrorInvalidValue;
    cudaError_t e = cudaSuccess;
    // item queue of this launch: the next slot of the context's ring.  Every slot is zero when its
    // turn comes: the ring starts zeroed, and each launch zeroes the slot half a ring ahead (no
    // launch can still be running when the slot it used comes up for zeroing MR_WORK_SLOTS / 2
    // launches later).  No memset, no global state: contexts are independent.
    const unsigned seq = ring->next++;
    unsigned int* ctr = ring->slots + (seq % MR_WORK_SLOTS);
    FastParams P;
    P.job = job;
    P.work_counter = ctr;
    P.work_reset = ring->slots + ((seq + MR_WORK_SLOTS / 2) % MR_WORK_SLOTS);
    P.padded = job.padded;
#ifdef MR_TIMELINE
    static unsigned long long* g_timeline = nullptr;
    if (!g_timeline) { cudaMalloc(&g_timeline, 4 * 8 * 148 * 32); }
    cudaMemsetAsync(g_timeline, 0, 4 * 8 * 148 * 32, s);
    P.timeline = g_timeline;
    g_timeline_host = g_timeline;
#endif
    const int K = job.pal.K;
#ifdef MR_QUICK   // kernel experiments: only the BASELINE config 1 / 2 / 4 / 5 instantiations (fast to compile)
    if (job.R == 2 && K > 14 && K <= 30) e = launch_c<2, 5>(P, sm_count, s);
    else if (job.R == 3 && K > 62) e = launch_c<3, 7>(P, sm_count, s);
    else if (job.R == 1 && K <= 14) e = launch_c<1, 4>(P, sm_count, s);
    else e = cudaErrorNotSupported;
#else
    if (job.R == 0) e = launch_c<0, 4>(P, sm_count, s);
    else if (K <= 14) e = launch_r<4>(P, sm_count, s);
    else if (K <= 30) e = launch_r<5>(P, sm_count, s);
    else if (K <= 62) e = launch_r<6>(P, sm_count, s);
    else e = launch_r<7>(P, sm_count, s);
#endif
    if (e == cudaSuccess) *n_launches += 1;
    return e;
}
