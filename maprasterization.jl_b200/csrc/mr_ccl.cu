// mr_ccl.cu -- segment prune on the GPU: the reference's real "clean" button,
// `_clean(A, known; categories, line_threshold, prune_threshold)` (src/clean.jl:59-93).
//
// The reference segments the label image with a sequential raster-scan union-find
// (`fast_scanning!`, src/scan.jl:56-157, threshold 1e-9, match = (:color,)): two pixels end up in
// one segment iff they are 4-connected through equal labels -- provided no segment holds two
// different known categories (then the reference splits it at the pixel where the merge is
// refused, scan.jl:140-146, which depends on the scan order; for a segment whose label is kept
// that case is detected here and reported as MR_ERR_UNSUPPORTED instead of being guessed; a
// segment whose label is not kept -- e.g. the no-match background with clicked points that did
// not match -- becomes 0 whichever way it is split).  So this is connected-component
// labelling + per-component statistics + a per-component rule.
//
// RUN based: a cleaned map has long horizontal runs of equal labels (~30 pixels on the synthetic
// sheets), so everything after the first two passes works on runs, not pixels:
//
//   k_count_stage  (line, 4096-pixel chunk): number of run starts; per thread (16 pixels, one 16-byte load) the
//              run-start mask and its offset in the chunk (`info`, 4 B per 16 pixels); the chunk's run starts
//              and labels compacted into the chunk's staging slot.  THE ONLY PASS THAT READS THE LABELS.
//   k_scan     exclusive scan of the chunk counts -> first run of every chunk / line
//   k_fill_runs  staging slots -> run arrays (start x, label, parent = self): work per run, not per pixel
//   k_merge    one block per line: every run finds the runs of the previous line that overlap it
//              (binary search in the sorted run starts) and unites with those of equal label
//              (lock-free union-find, atomicMin links towards the smaller run index)
//   k_number_roots / k_flatten   roots take a compact component id; every run's parent entry becomes TAG | id
//   k_stats    per run: count, bounding box, label -> atomics on the component's record
//   k_known    known-category plane (sparse list): pixel -> run (binary search) -> component
//   k_decide   per component: the rule of clean.jl:65-88 (Float64 arithmetic as written there)
//   k_runout   per run: the label it will be painted with
//   k_paint    (line, chunk) again: `info` -> run -> label, 16-byte stores (no label read, no block scan)
//
// Run / pixel indices fit 31 bits (n1 * n2 < 2^31; bit 31 tags a numbered root).
// Bound: HBM; algorithmic traffic 2 B/pixel (label in, label out); the labels are read once and written
// once, plus 0.5 B per pixel of `info` (written, read) and ~25 B per run.  The streaming kernels handle
// UB chunks per block (independent loads in flight, one barrier pair for all of them).
#include "mr_internal.h"

namespace {

constexpr uint32_t TAG = 0x80000000u;
constexpr int PT = 16;            // pixels per thread
constexpr int BT = 256;           // threads per block
constexpr int CH = PT * BT;       // pixels per chunk
constexpr int UB = 2;             // chunks per block in the streaming kernels (count: 12 KB of shared memory each)

// per-chunk staging written by k_count_stage: the only pass over the label plane
struct Stage {
    uint32_t* info;             // [chunk][BT]: run-start mask of the thread's 16 pixels | offset of its first run in the chunk << 16
    uint16_t* sx;               // [chunk][CH]: x - chunk start of the chunk's runs, in order (first `count` entries)
    uint8_t* sl;                // [chunk][CH]: their labels
};

struct SegRec {                 // one per component; 32 bytes, read and written as two uint4 (k_init_recs, k_decide, k_stats)
    unsigned int xmin, xmax, ymin, ymax;   // the box: one 16-byte read in k_stats
    unsigned int count;
    unsigned int cmin, cmax;    // known categories present: min / max of the non-zero ones
    unsigned int label;         // input label, later (k_decide) the output label
};

struct Runs {
    uint32_t* start;            // x of the first pixel
    uint32_t* parent;           // union-find over run indices; after k_number_roots / k_flatten: TAG | component id
    uint8_t* label;
    uint8_t* out;               // label the run is painted with
    uint32_t* rootof;           // component id -> its root run (first n_segments entries)
    const uint32_t* cbase;      // first run of every (line, chunk); cbase[n2 * chunks] = number of runs
    uint32_t chunks;            // chunks per line
};

__device__ __forceinline__ uint32_t find_root(uint32_t* parent, uint32_t x) {
    uint32_t p = parent[x];
    while (p != x) {            // path halving; racing writers only ever shorten paths
        const uint32_t g = parent[p];
        if (g != p) parent[x] = g;
        x = p;
        p = g;
    }
    return x;
}

__device__ __forceinline__ void unite(uint32_t* parent, uint32_t a, uint32_t b) {
    while (true) {
        a = find_root(parent, a);
        b = find_root(parent, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }   // link the larger root under the smaller
        const uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// run-start mask of the 16 pixels in w[0..3] (byte-parallel): bit k <=> pixel k differs from the pixel before it;
// prev = label of the pixel before w[0]'s first one (any value > 255: none)
__device__ __forceinline__ uint32_t starts16(const uint32_t (&w)[4], uint32_t prev) {
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t sh = i == 0 ? ((w[0] << 8) | (prev & 0xFFu)) : __funnelshift_l(w[i - 1], w[i], 8);
        const uint32_t t = w[i] ^ sh;
        const uint32_t nz = ((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) >> 7) & 0x01010101u;   // byte != 0 -> bit 0 of the byte
        m |= ((nz * 0x10204080u) >> 28) << (4 * i);                                        // bits 0, 8, 16, 24 -> a nibble
    }
    if (prev > 0xFFu) m |= 1u;
    return m;
}

// The only pass over the label plane.  Block = CT = 128 threads = two 4096-pixel chunk units of one line, 64 threads
// (two warps) per unit; a thread owns GT = 4 groups of 16 consecutive pixels (four 16-byte loads issued together:
// 64 contiguous bytes), so the scan, the left-neighbour exchange and the address arithmetic are paid once per 64
// pixels (the 16-pixels-per-thread version was instruction bound: 250 instructions per 16 pixels, SM busy 82 %,
// DRAM 28 %).  Per group: run-start mask (byte-parallel) and the `info` word (mask | offset of its first run in
// the chunk); the run starts / labels of a unit are compacted in shared memory and leave as contiguous stores.
constexpr int GT = 4;                 // groups of PT pixels per thread
constexpr int CT = UB * (BT / GT);    // threads per block: BT / GT = 64 per unit
__global__ void __launch_bounds__(CT) k_count_stage(const uint8_t* __restrict__ lab, int64_t stride2, uint32_t n1, bool al16,
                                                    uint32_t chunks, uint32_t n_units, uint32_t* __restrict__ cnt, Stage G) {
    __shared__ uint16_t s_x[UB][CH];
    __shared__ uint8_t s_l[UB][CH];
    __shared__ uint32_t s_w[UB][2];                       // run counts of the two warps of a unit
    const uint32_t groups = (chunks + UB - 1) / UB;       // blocks per line
    const uint32_t y = blockIdx.x / groups, c0 = (blockIdx.x - y * groups) * UB;   // line, first chunk of this block (uniform)
    const int u = threadIdx.x / (BT / GT), tu = threadIdx.x % (BT / GT);           // unit of the block, thread within the unit
    const int lane = threadIdx.x & 31, wu = tu >> 5;                               // warp within the unit
    const bool live = c0 + u < chunks;                                             // warp-uniform
    const uint32_t unit = y * chunks + c0 + u;
    const uint32_t x0 = (c0 + u) * CH + tu * (GT * PT);                            // first pixel of this thread
    const uint8_t* line = lab + (int64_t)y * stride2;
    uint32_t w[GT][4], m[GT];
    int nv[GT];
#pragma unroll
    for (int g = 0; g < GT; ++g) {
        const uint32_t x = x0 + g * PT;
        w[g][0] = w[g][1] = w[g][2] = w[g][3] = 0u;
        nv[g] = (!live || x >= n1) ? 0 : (n1 - x < (uint32_t)PT ? (int)(n1 - x) : PT);
        if (nv[g] == PT && al16) {
            const uint4 v = *reinterpret_cast<const uint4*>(line + x);
            w[g][0] = v.x; w[g][1] = v.y; w[g][2] = v.z; w[g][3] = v.w;
        } else if (nv[g] > 0) {       // ragged end of the line / unaligned lines: byte loads (never for groups outside the line)
#pragma unroll
            for (int k = 0; k < PT; ++k)
                if (k < nv[g]) w[g][k >> 2] |= (uint32_t)line[x + k] << (8 * (k & 3));
        }
    }
    // label of the pixel before the thread's first one: the previous thread's last pixel (it has 64 valid pixels
    // whenever this thread has any); first lane of a warp: from memory; first pixel of the line: none
    uint32_t prev = __shfl_up_sync(0xffffffffu, w[GT - 1][3] >> 24, 1);
    if (lane == 0) prev = (x0 > 0 && nv[0] > 0) ? (uint32_t)line[x0 - 1] : 0x100u;
    if (x0 == 0) prev = 0x100u;
    uint32_t pc = 0;
#pragma unroll
    for (int g = 0; g < GT; ++g) {
        m[g] = starts16(w[g], g == 0 ? prev : (w[g - 1][3] >> 24)) & (nv[g] == PT ? 0xFFFFu : ((1u << nv[g]) - 1u));
        pc += __popc(m[g]);
    }
    // exclusive prefix of pc over the 64 threads of the unit
    uint32_t sc = pc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += t;
    }
    if (lane == 31) s_w[u][wu] = sc;
    __syncthreads();
    const uint32_t tot = s_w[u][0] + s_w[u][1];
    uint32_t pos = sc - pc + (wu ? s_w[u][0] : 0u);
    if (live) {
        if (tu == 0) cnt[unit] = tot;
#pragma unroll
        for (int g = 0; g < GT; ++g) {
            G.info[(size_t)unit * BT + tu * GT + g] = m[g] | (pos << 16);
            uint32_t mm = m[g];
            while (mm) {
                const int k = __ffs(mm) - 1;
                mm &= mm - 1;
                s_x[u][pos] = (uint16_t)((tu * GT + g) * PT + k);
                // byte k of the 16: the 8-byte half by two selects, the byte by one permute
                s_l[u][pos] = (uint8_t)__byte_perm(k < 8 ? w[g][0] : w[g][2], k < 8 ? w[g][1] : w[g][3], (uint32_t)(k & 7));
                ++pos;
            }
        }
    }
    __syncthreads();
    if (live) {
        uint16_t* sx = G.sx + (size_t)unit * CH;
        uint8_t* sl = G.sl + (size_t)unit * CH;
        for (uint32_t j = tu; j < tot; j += BT / GT) {
            sx[j] = s_x[u][j];
            sl[j] = s_l[u][j];
        }
    }
}

// staging slots -> run arrays; FU chunk units per block (a unit holds ~140 runs on a filtered sheet: one block
// per unit left half of its threads idle)
constexpr int FU = 8;
__global__ void __launch_bounds__(BT) k_fill_runs(Runs R, Stage G, uint32_t n_units) {
    const uint32_t u0 = blockIdx.x * FU, u1 = min(u0 + FU, n_units);
    const uint32_t base0 = R.cbase[u0], n = R.cbase[u1] - base0;
    for (uint32_t j = threadIdx.x; j < n; j += BT) {
        const uint32_t i = base0 + j;
        uint32_t u = u0;                                  // the unit that owns run i (FU - 1 comparisons)
#pragma unroll
        for (int k = 1; k < FU; ++k)
            if (u0 + k < u1 && R.cbase[u0 + k] <= i) u = u0 + k;
        const uint32_t off = i - R.cbase[u];
        R.start[i] = (u % R.chunks) * CH + G.sx[(size_t)u * CH + off];
        R.label[i] = G.sl[(size_t)u * CH + off];
        R.parent[i] = i;
    }
}

// in-place exclusive scan of n counts, total appended at a[n]; one block.  Every thread owns one contiguous
// segment (a multiple of four entries, 16-byte loads): it sums the segment, the block scans the 1024 sums, the
// thread walks its segment again writing the running prefix.  (The first version walked the array in 13 steps of
// 8192 entries with three barriers and a round trip to L2 each: 57 us for 100 k entries.)
__global__ void __launch_bounds__(1024) k_scan(uint32_t* a, uint64_t n) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t seg = ((n + 1023) / 1024 + 3) & ~(uint64_t)3;
    const uint64_t b = min((uint64_t)threadIdx.x * seg, n), e = min(b + seg, n);
    const uint64_t e4 = b + ((e - b) & ~(uint64_t)3);
    uint32_t mine = 0;
#pragma unroll 8
    for (uint64_t i = b; i < e4; i += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(a + i);
        mine += v.x + v.y + v.z + v.w;
    }
    for (uint64_t i = e4; i < e; ++i) mine += a[i];
    uint32_t s = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
        const uint32_t t = wsum[w];
        if (w < warp) base += t;
        total += t;
    }
    __syncthreads();                       // every segment has been summed before any entry is overwritten
    uint32_t run = base + s - mine;
#pragma unroll 8
    for (uint64_t i = b; i < e4; i += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(a + i);
        uint4 o;
        o.x = run; run += v.x;
        o.y = run; run += v.y;
        o.z = run; run += v.z;
        o.w = run; run += v.w;
        *reinterpret_cast<uint4*>(a + i) = o;
    }
    for (uint64_t i = e4; i < e; ++i) {
        const uint32_t v = a[i];
        a[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) a[n] = total;
}

// one block per line y >= 1 (blockIdx.x = y - 1)
__global__ void __launch_bounds__(128) k_merge(uint32_t n1, Runs R) {
    const uint32_t y = blockIdx.x + 1;
    const uint32_t p0 = R.cbase[(uint64_t)(y - 1) * R.chunks], c0 = R.cbase[(uint64_t)y * R.chunks],
                   c1 = R.cbase[(uint64_t)(y + 1) * R.chunks];
    // The trip count is warp-uniform and the lanes are brought together at the top of every step: the unite loops run
    // for very different times, and without the barrier the lanes of a warp drift apart for good -- on raw labels
    // the binary search then ran with 10 of 32 lanes (ncu line profile: 767 M warp instructions for 36 M runs).
    for (uint32_t ib = c0; ib < c1; ib += blockDim.x) {
        __syncwarp();
        const uint32_t i = ib + threadIdx.x;
        const bool on = i < c1;
        uint32_t s = 0, e = 0, lo = p0, hi = on ? c0 : p0;   // invariant: start[lo] <= s (start[p0] = 0), start[hi] > s or hi == c0
        uint8_t L = 0;
        if (on) {
            s = R.start[i];
            e = (i + 1 < c1) ? R.start[i + 1] : n1;          // [s, e)
            L = R.label[i];
        }
        // last run of the previous line that starts at or before s
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (R.start[mid] <= s) lo = mid; else hi = mid;
        }
        __syncwarp();
        if (on)
            for (uint32_t j = lo; j < c0 && R.start[j] < e; ++j)
                if (R.label[j] == L) unite(R.parent, i, j);
    }
}

// roots take a compact component id: parent[root] = TAG | id, rootof[id] = root.  Four entries per thread
// (one 16-byte load; the run arrays are 16-byte aligned and padded to a multiple of four entries).
__global__ void k_number_roots(uint32_t* __restrict__ parent, uint32_t n, unsigned int* n_roots, uint32_t* __restrict__ rootof) {
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    uint32_t rm = 0u;
    if (i0 < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(parent + i0);
        rm = (v.x == i0 ? 1u : 0u) | (v.y == i0 + 1 && i0 + 1 < n ? 2u : 0u) | (v.z == i0 + 2 && i0 + 2 < n ? 4u : 0u) |
             (v.w == i0 + 3 && i0 + 3 < n ? 8u : 0u);
    }
    const uint32_t c = __popc(rm), lane = threadIdx.x & 31;
    uint32_t sc = c;                                   // inclusive scan of the counts over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += t;
    }
    // one atomic per BLOCK on the single counter (raw label sheets have millions of roots: one atomic per warp
    // serialised on that address, 174 us for 36 M runs)
    __shared__ uint32_t s_wt[8], s_base;
    const uint32_t tot = __shfl_sync(0xffffffffu, sc, 31);
    const int warp = threadIdx.x >> 5;
    if (lane == 0) s_wt[warp] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t all = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { const uint32_t t = s_wt[w]; s_wt[w] = all; all += t; }
        s_base = all ? atomicAdd(n_roots, all) : 0u;
    }
    __syncthreads();
    if (!tot) return;
    uint32_t id = s_base + s_wt[warp] + sc - c;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((rm >> k) & 1u) {
            parent[i0 + k] = TAG | id;
            rootof[id] = i0 + k;
            ++id;
        }
}

// every run takes its component's id: walk towards the root until an entry carries a tag (the root's, or the one
// another thread has already written into a run on the way: same component), then parent[i] = TAG | id.
// Afterwards the component of a run is ONE coalesced read (no second, random read of the root's entry).
// Four consecutive runs per thread (one 16-byte load / store of their entries; the arrays are padded to a
// multiple of four): four independent chains of dependent L2 reads in flight per thread (165 -> 122 us).
// (The same idea in k_merge -- two binary searches per thread advancing together -- gave nothing: 200 -> 204 us.)
__global__ void k_flatten(uint32_t* parent, uint32_t n) {
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    const uint4 v = *reinterpret_cast<const uint4*>(parent + i0);
    uint32_t p[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k >= n) p[k] = TAG;                    // padding entries
    while (!(p[0] & p[1] & p[2] & p[3] & TAG)) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (!(p[k] & TAG)) p[k] = parent[p[k]];     // a stale cached entry is still a valid link of the same component
    }
    if (i0 + 3 < n) *reinterpret_cast<uint4*>(parent + i0) = make_uint4(p[0], p[1], p[2], p[3]);
    else
        for (int k = 0; k < 4 && i0 + k < n; ++k) parent[i0 + k] = p[k];
}

__device__ __forceinline__ uint32_t comp_id(const uint32_t* parent, uint32_t run) { return parent[run] & ~TAG; }

__global__ void k_init_recs(SegRec* rec, unsigned int n, Runs R) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // two 16-byte stores: { xmin, xmax, ymin, ymax } { count, cmin, cmax, label }
    const uint32_t label = R.label[R.rootof[i]];   // the component's label = the label of any of its runs
    uint4* o = reinterpret_cast<uint4*>(rec + i);
    o[0] = make_uint4(0xffffffffu, 0u, 0xffffffffu, 0u);
    o[1] = make_uint4(0u, 0xffffffffu, 0u, label);
}

// one block per line.  (Grouping the runs of a warp by component -- match.any + redux -- before touching the
// records was measured: 280 -> 367 us; the conditional min / max atomics are almost always skipped and the
// count is a fire-and-forget reduction.)
__global__ void __launch_bounds__(128) k_stats(uint32_t n1, Runs R, SegRec* __restrict__ rec) {
    const uint32_t y = blockIdx.x;
    const uint32_t c0 = R.cbase[(uint64_t)y * R.chunks], c1 = R.cbase[(uint64_t)(y + 1) * R.chunks];
    // two runs per thread and step: their component reads and box reads are in flight together
    for (uint32_t ib = c0 + threadIdx.x; ib < c1; ib += 2 * blockDim.x) {
        uint32_t s[2], e[2];
        SegRec* r[2];
        bool on[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t i = ib + k * blockDim.x;
            on[k] = i < c1;
            s[k] = 0; e[k] = 1; r[k] = rec;
            if (on[k]) {
                s[k] = R.start[i];
                e[k] = (i + 1 < c1) ? R.start[i + 1] : n1;
                r[k] = rec + comp_id(R.parent, i);
            }
        }
        uint4 box[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            box[k] = make_uint4(0u, 0xffffffffu, 0u, 0xffffffffu);
            if (on[k]) {
                atomicAdd(&r[k]->count, e[k] - s[k]);
                // the box so far in one 16-byte read (a stale copy only costs a redundant atomic: min / max are monotone)
                box[k] = *reinterpret_cast<const uint4*>(&r[k]->xmin);
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (on[k]) {
                if (box[k].x > s[k]) atomicMin(&r[k]->xmin, s[k]);
                if (box[k].y < e[k] - 1) atomicMax(&r[k]->xmax, e[k] - 1);
                if (box[k].z > y) atomicMin(&r[k]->ymin, y);
                if (box[k].w < y) atomicMax(&r[k]->ymax, y);
            }
    }
}

__global__ void k_known(const int64_t* __restrict__ idx, const uint8_t* __restrict__ cat, int64_t n_known, uint32_t n1,
                        uint32_t n2, Runs R, SegRec* __restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_known) return;
    const int64_t p = idx[i];
    const unsigned c = cat[i];
    if (p < 0 || p >= (int64_t)n1 * n2 || c == 0) return;
    const uint32_t y = (uint32_t)(p / n1), x = (uint32_t)(p % n1);
    uint32_t lo = R.cbase[(uint64_t)y * R.chunks], hi = R.cbase[(uint64_t)(y + 1) * R.chunks];
    while (hi - lo > 1) {   // last run of line y that starts at or before x
        const uint32_t mid = (lo + hi) >> 1;
        if (R.start[mid] <= x) lo = mid; else hi = mid;
    }
    SegRec* r = rec + comp_id(R.parent, lo);
    atomicMin(&r->cmin, c);
    atomicMax(&r->cmax, c);
}

// clean.jl:65-88 for one component; Float64 operations as written in the reference
__global__ void k_decide(SegRec* __restrict__ rec, unsigned int n, const uint8_t* __restrict__ keep,
                         double line_threshold, double prune_threshold, unsigned int* conflict) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SegRec r;
    {   // two 16-byte loads
        const uint4 a = reinterpret_cast<const uint4*>(rec + i)[0], b = reinterpret_cast<const uint4*>(rec + i)[1];
        r.xmin = a.x; r.xmax = a.y; r.ymin = a.z; r.ymax = a.w;
        r.count = b.x; r.cmin = b.y; r.cmax = b.z; r.label = b.w;
    }
    const unsigned L = r.label;
    unsigned out = L;
    if (!keep[L]) {
        out = 0u;   // however the reference would split this segment, every part has a label that is not kept
    } else {
        if (r.cmax != 0 && r.cmin != r.cmax) *conflict = 1u;   // two known categories in one kept segment
        const unsigned h = r.xmax - r.xmin + 1, w = r.ymax - r.ymin + 1;
        const unsigned long long m = h > w ? h : w;
        const double rectangle_area = __ddiv_rn((double)(m * m), 2.0);
        const double ratio = __ddiv_rn((double)r.count, rectangle_area);
        if (((double)r.count < prune_threshold) || (ratio < line_threshold)) {
            const unsigned c = r.cmax;
            if (c == 0) out = 0u;
            else if (!(L == 0 || L == c)) out = c;
        }
    }
    rec[i].label = out;
}

__global__ void k_runout(Runs R, uint32_t n, const SegRec* __restrict__ rec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) R.out[i] = (uint8_t)rec[comp_id(R.parent, i)].label;
}

// PU chunk units per block; per thread: `info` word -> first run and run-start mask of its 16 pixels.  The info
// words, chunk bases and first output labels of all PU units are fetched before any of them is used (three
// dependent loads per unit: independent chains in flight).
constexpr int PU = 4;
__global__ void __launch_bounds__(BT) k_paint(uint32_t n1, uint32_t n_units, Runs R, Stage G, uint8_t* __restrict__ out,
                                              int64_t out_stride2, bool out_al16) {
    const uint32_t groups = (R.chunks + PU - 1) / PU;          // blocks per line
    const uint32_t y = blockIdx.x / groups, c0 = (blockIdx.x - y * groups) * PU;   // line, first chunk of this block (one division)
    const uint32_t unit0 = y * R.chunks + c0;
    uint32_t inf[PU], run[PU];
    uint8_t cur[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        const bool on = c0 + u < R.chunks;
        inf[u] = on ? G.info[(size_t)(unit0 + u) * BT + threadIdx.x] : 0u;
        run[u] = on ? R.cbase[unit0 + u] : 0u;
    }
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        run[u] += inf[u] >> 16;                                  // next run to start
        // the run that continues into this thread (none when its first pixel starts one, or outside the sheet)
        cur[u] = ((inf[u] & 1u) || run[u] == 0u || c0 + u >= R.chunks) ? (uint8_t)0 : R.out[run[u] - 1];
    }
    uint8_t* row = out + (int64_t)y * out_stride2;
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        if (c0 + u >= R.chunks) break;
        const uint32_t x = (c0 + u) * CH + threadIdx.x * PT;
        if (x >= n1) continue;
        const int nv = n1 - x < (uint32_t)PT ? (int)(n1 - x) : PT;
        // the continuing run's label everywhere, then every run start (0.5 per thread on a filtered sheet)
        // overwrites the bytes from its pixel on: work per run, not per pixel; the 16 bytes as two 64-bit halves
        uint32_t r = run[u], mm = inf[u] & 0xFFFFu;
        unsigned long long lo = (unsigned long long)cur[u] * 0x0101010101010101ull, hi = lo;
        while (mm) {
            const int k = __ffs(mm) - 1;
            mm &= mm - 1;
            const unsigned long long L8 = (unsigned long long)R.out[r++] * 0x0101010101010101ull;
            const unsigned long long keep = (1ull << (8 * (k & 7))) - 1ull;     // the bytes of the half before pixel k
            if (k < 8) {
                lo = (lo & keep) | (L8 & ~keep);
                hi = L8;
            } else {
                hi = (hi & keep) | (L8 & ~keep);
            }
        }
        uint8_t* dst = row + x;
        if (nv == PT && out_al16) {
            *reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
        } else {
            for (int k = 0; k < nv; ++k) dst[k] = (uint8_t)((k < 8 ? lo : hi) >> (8 * (k & 7)));
        }
    }
}

Stage make_stage(void* stage_mem, int64_t n1, int64_t n2) {
    const size_t units = (size_t)((n1 + CH - 1) / CH) * (size_t)n2;
    Stage G;
    uint8_t* p = reinterpret_cast<uint8_t*>(stage_mem);
    G.info = reinterpret_cast<uint32_t*>(p);
    G.sx = reinterpret_cast<uint16_t*>(p + units * BT * 4);
    G.sl = p + units * BT * 4 + units * CH * 2;
    return G;
}

Runs make_runs(void* run_mem, unsigned int n_runs, const uint32_t* cbase, int64_t n1) {
    Runs R;
    uint8_t* p = reinterpret_cast<uint8_t*>(run_mem);
    const size_t n4 = ((size_t)n_runs + 3) / 4 * 4;   // keep the arrays 16-byte aligned
    R.start = reinterpret_cast<uint32_t*>(p);
    R.parent = reinterpret_cast<uint32_t*>(p + n4 * 4);
    R.label = p + n4 * 8;
    R.out = p + n4 * 9;
    R.rootof = reinterpret_cast<uint32_t*>(p + n4 * 10);
    R.cbase = cbase;
    R.chunks = (uint32_t)((n1 + CH - 1) / CH);
    return R;
}

bool aligned16(const void* p, int64_t stride) { return ((uintptr_t)p % 16 == 0) && (stride % 16 == 0); }

}  // namespace

size_t mr_prune_cbase_bytes(int64_t n1, int64_t n2) {
    return ((size_t)((n1 + CH - 1) / CH) * (size_t)n2 + 1) * sizeof(uint32_t);
}
// staging of k_count_stage: 4 B per 16 pixels (`info`) + 3 B per pixel slot of the chunk's compacted run starts / labels
size_t mr_prune_stage_bytes(int64_t n1, int64_t n2) {
    const size_t units = (size_t)((n1 + CH - 1) / CH) * (size_t)n2;
    return units * ((size_t)BT * 4 + (size_t)CH * 3) + 16;
}
size_t mr_prune_run_bytes(unsigned int n_runs) { return (((size_t)n_runs + 3) / 4 * 4) * 14 + 16; }
size_t mr_prune_rec_bytes(unsigned int n_segments) { return ((size_t)n_segments + 1) * sizeof(SegRec); }

// phase 1: run starts per (line, chunk) and their exclusive scan; the number of runs ends up in
// cbase[chunks * n2] (device)
cudaError_t mr_launch_prune_count(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, uint32_t* cbase,
                                  void* stage_mem, cudaStream_t s, int* n_launches) {
    const uint32_t chunks = (uint32_t)((n1 + CH - 1) / CH);
    const uint32_t n_units = chunks * (uint32_t)n2;
    k_count_stage<<<(unsigned)n2 * ((chunks + UB - 1) / UB), CT, 0, s>>>(labels, stride2, (uint32_t)n1, aligned16(labels, stride2), chunks,
                                                        n_units, cbase, make_stage(stage_mem, n1, n2));
    k_scan<<<1, 1024, 0, s>>>(cbase, (uint64_t)chunks * (uint64_t)n2);
    *n_launches += 2;
    return cudaGetLastError();
}

// phase 2: runs, vertical merges, component numbering; the number of components ends up in *d_n_roots
cudaError_t mr_launch_prune_components(int64_t n1, int64_t n2, const uint32_t* cbase, const void* stage_mem, void* run_mem,
                                       unsigned int n_runs, unsigned int* d_n_roots, cudaStream_t s, int* n_launches) {
    const Runs R = make_runs(run_mem, n_runs, cbase, n1);
    cudaError_t e = cudaMemsetAsync(d_n_roots, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    const uint32_t n_units = R.chunks * (uint32_t)n2;
    k_fill_runs<<<(n_units + FU - 1) / FU, BT, 0, s>>>(R, make_stage(const_cast<void*>(stage_mem), n1, n2), n_units);
    *n_launches += 1;
    if (n2 > 1) {
        k_merge<<<(unsigned)(n2 - 1), 128, 0, s>>>((uint32_t)n1, R);
        *n_launches += 1;
    }
    k_number_roots<<<(n_runs + 1023) / 1024, 256, 0, s>>>(R.parent, n_runs, d_n_roots, R.rootof);
    k_flatten<<<(n_runs + 1023) / 1024, 256, 0, s>>>(R.parent, n_runs);
    *n_launches += 2;
    return cudaGetLastError();
}

// phase 3: statistics, rule, output.  *d_conflict (device) becomes non-zero when a segment holds
// two different known categories.
cudaError_t mr_launch_prune_apply(int64_t n1, int64_t n2, const uint32_t* cbase, const void* stage_mem, void* run_mem,
                                  unsigned int n_runs, void* rec,
                                  unsigned int n_segments, int64_t n_known, const int64_t* known_idx,
                                  const uint8_t* known_cat, const uint8_t* keep_dev, double line_threshold,
                                  double prune_threshold, uint8_t* out, int64_t out_stride2,
                                  unsigned int* d_conflict, cudaStream_t s, int* n_launches) {
    const Runs R = make_runs(run_mem, n_runs, cbase, n1);
    SegRec* r = reinterpret_cast<SegRec*>(rec);
    cudaError_t e = cudaMemsetAsync(d_conflict, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    k_init_recs<<<(n_segments + 255) / 256, 256, 0, s>>>(r, n_segments, R);
    k_stats<<<(unsigned)n2, 128, 0, s>>>((uint32_t)n1, R, r);
    *n_launches += 2;
    if (n_known > 0) {
        k_known<<<(unsigned)((n_known + 255) / 256), 256, 0, s>>>(known_idx, known_cat, n_known, (uint32_t)n1,
                                                                  (uint32_t)n2, R, r);
        *n_launches += 1;
    }
    k_decide<<<(n_segments + 255) / 256, 256, 0, s>>>(r, n_segments, keep_dev, line_threshold, prune_threshold, d_conflict);
    k_runout<<<(n_runs + 255) / 256, 256, 0, s>>>(R, n_runs, r);
    const uint32_t n_units = R.chunks * (uint32_t)n2;
    k_paint<<<(unsigned)n2 * ((R.chunks + PU - 1) / PU), BT, 0, s>>>((uint32_t)n1, n_units, R, make_stage(const_cast<void*>(stage_mem), n1, n2), out,
                                                  out_stride2, aligned16(out, out_stride2));
    *n_launches += 3;
    return cudaGetLastError();
}
