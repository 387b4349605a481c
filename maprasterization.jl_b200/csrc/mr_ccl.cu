// mr_ccl.cu -- segment prune on the GPU: the reference's real "clean" button,
// `_clean(A, known; categories, line_threshold, prune_threshold)` (src/clean.jl:59-93).
//
// The reference segments the label image with a sequential raster-scan union-find
// (`fast_scanning!`, src/scan.jl:56-157, threshold 1e-9, match = (:color,)): two pixels end up in
// one segment iff they are 4-connected through equal labels -- provided no segment holds two
// different known categories (then the reference splits it at the pixel where the merge is
// refused, scan.jl:140-146, which depends on the scan order; that case is detected here and
// reported as MR_ERR_UNSUPPORTED instead of being guessed).  So this is connected-component
// labelling + per-component statistics + a per-component rule:
//
//   k_runs      lane = pixel of a line: parent = start of its horizontal run of equal labels
//               (warp ballot; a run that began in an earlier 32-pixel group links to that group)
//   k_vmerge    union of vertically adjacent runs (lock-free union-find, atomicMin linking towards
//               the smaller index); only the first pixel of every horizontal overlap does the union
//   k_flatten   parent = root
//   k_number    roots take a compact id (warp-aggregated atomic counter); id tagged into the root
//   k_stats     per (run x 32-pixel group): count, bounding box -> atomics on the component's record
//   k_known     known-category plane (sparse list) -> per-component min / max known category
//   k_relabel   out = rule(component record)  (clean.jl:65-88, Float64 arithmetic as written there)
//
// Pixel index p = y * n1 + x fits 31 bits (n1 * n2 < 2^31; bit 31 tags a numbered root).
// Bound: HBM; algorithmic traffic 2 B/pixel (label in, label out); this first version moves
// ~30 B/pixel (4-byte parent array, five passes) -- see DESIGN.md.
#include "mr_internal.h"

namespace {

constexpr uint32_t TAG = 0x80000000u;

struct SegRec {                 // one per component
    unsigned int count;
    unsigned int xmin, xmax, ymin, ymax;
    unsigned int cmin, cmax;    // known categories present: min / max of the non-zero ones
    unsigned int label;
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

// grid: x = 32-pixel groups of a line, y = lines
__global__ void k_runs(const uint8_t* __restrict__ lab, int64_t stride2, uint32_t n1, uint32_t n2,
                       uint32_t* __restrict__ parent) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t groups = (n1 + 31) / 32;
    const uint64_t gw = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (uint64_t)groups * n2) return;
    const uint32_t y = (uint32_t)(gw / groups), x0 = (uint32_t)(gw % groups) * 32, x = x0 + lane;
    const bool in = x < n1;
    const uint8_t* line = lab + (int64_t)y * stride2;
    const int me = in ? line[x] : -1;
    int left = __shfl_up_sync(0xffffffffu, me, 1);
    if (lane == 0) left = x0 > 0 ? line[x0 - 1] : -1;
    const bool start = in && (left != me);
    const uint32_t starts = __ballot_sync(0xffffffffu, start);
    if (!in) return;
    const uint32_t below = starts & (0xffffffffu >> (31 - lane));   // starts at or before this lane
    const uint32_t base = y * n1;
    parent[base + x] = below ? base + x0 + (31 - __clz(below)) : base + x0 - 1;   // previous group continues
}

__global__ void k_vmerge(const uint8_t* __restrict__ lab, int64_t stride2, uint32_t n1, uint32_t n2,
                         uint32_t* __restrict__ parent) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n1 * (n2 - 1)) return;
    const uint32_t y = (uint32_t)(i / n1) + 1, x = (uint32_t)(i % n1);
    const uint8_t* line = lab + (int64_t)y * stride2;
    const uint8_t* up = line - stride2;
    const uint8_t me = line[x];
    if (up[x] != me) return;
    if (x > 0 && line[x - 1] == me && up[x - 1] == me) return;   // the pixel to the left did this union
    unite(parent, y * n1 + x, (y - 1) * n1 + x);
}

__global__ void k_flatten(uint32_t* __restrict__ parent, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // read-only walk: a path-halving write by another thread could land AFTER the owner stored the
    // root and leave a non-root in the entry.  Every entry is only ever written by its owner here.
    uint32_t r = (uint32_t)i, p = parent[r];
    while (p != r) {
        r = p;
        p = parent[r];
    }
    parent[i] = r;
}

__global__ void k_number(uint32_t* __restrict__ parent, uint64_t n, unsigned int* n_roots) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool root = i < n && parent[i] == (uint32_t)i;
    const uint32_t m = __ballot_sync(0xffffffffu, root);
    if (!m) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(n_roots, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (root) parent[i] = TAG | (base + __popc(m & ((1u << lane) - 1u)));
}

__global__ void k_init_recs(SegRec* rec, unsigned int n) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SegRec r;
    r.count = 0;
    r.xmin = r.ymin = 0xffffffffu;
    r.xmax = r.ymax = 0;
    r.cmin = 0xffffffffu;
    r.cmax = 0;
    r.label = 0;
    rec[i] = r;
}

__device__ __forceinline__ uint32_t comp_id(const uint32_t* parent, uint32_t p) {
    const uint32_t v = parent[p];
    return (v & TAG) ? (v & ~TAG) : (parent[v] & ~TAG);
}

// same geometry as k_runs: one record update per (run x 32-pixel group)
__global__ void k_stats(const uint8_t* __restrict__ lab, int64_t stride2, uint32_t n1, uint32_t n2,
                        const uint32_t* __restrict__ parent, SegRec* __restrict__ rec) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t groups = (n1 + 31) / 32;
    const uint64_t gw = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (uint64_t)groups * n2) return;
    const uint32_t y = (uint32_t)(gw / groups), x0 = (uint32_t)(gw % groups) * 32, x = x0 + lane;
    const bool in = x < n1;
    const uint8_t* line = lab + (int64_t)y * stride2;
    const int me = in ? line[x] : -1;
    const int left = __shfl_up_sync(0xffffffffu, me, 1);
    const bool lead = in && (lane == 0 || left != me);
    const uint32_t leads = __ballot_sync(0xffffffffu, lead);
    const uint32_t valid = __ballot_sync(0xffffffffu, in);
    if (!lead) return;
    const uint32_t above = leads & ~((2u << lane) - 1u);            // next leader after this lane
    const uint32_t end = above ? (uint32_t)(__ffs(above) - 1) : (uint32_t)__popc(valid);   // one past the group's last lane
    const uint32_t len = end - lane;
    const bool true_start = (x == 0) || (line[x - 1] != me);
    const bool true_end = (x + len == n1) || (line[x + len] != me);
    SegRec* r = rec + comp_id(parent, y * n1 + x);
    atomicAdd(&r->count, len);
    if (true_start) atomicMin(&r->xmin, x);
    if (true_end) atomicMax(&r->xmax, x + len - 1);
    if (r->ymin > y) atomicMin(&r->ymin, y);
    if (r->ymax < y) atomicMax(&r->ymax, y);
    if (true_start && r->label != (unsigned)me) r->label = (unsigned)me;   // same value from every writer
}

__global__ void k_known(const int64_t* __restrict__ idx, const uint8_t* __restrict__ cat, int64_t n_known,
                        uint64_t npx, const uint32_t* __restrict__ parent, SegRec* __restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_known) return;
    const int64_t p = idx[i];
    const unsigned c = cat[i];
    if (p < 0 || (uint64_t)p >= npx || c == 0) return;
    SegRec* r = rec + comp_id(parent, (uint32_t)p);
    atomicMin(&r->cmin, c);
    atomicMax(&r->cmax, c);
}

// clean.jl:65-88 for one component; Float64 operations as written in the reference
__device__ __forceinline__ unsigned rule(const SegRec& r, const uint8_t* keep, double line_threshold,
                                         double prune_threshold, unsigned int* conflict) {
    const unsigned L = r.label;
    if (r.cmax != 0 && r.cmin != r.cmax) *conflict = 1u;   // two known categories in one segment
    if (!keep[L]) return 0u;
    const unsigned h = r.xmax - r.xmin + 1, w = r.ymax - r.ymin + 1;
    const unsigned long long m = h > w ? h : w;
    const double rectangle_area = __ddiv_rn((double)(m * m), 2.0);
    const double ratio = __ddiv_rn((double)r.count, rectangle_area);
    if (((double)r.count < prune_threshold) || (ratio < line_threshold)) {
        const unsigned c = r.cmax;
        if (c == 0) return 0u;
        if (!(L == 0 || L == c)) return c;
    }
    return L;
}

__global__ void k_relabel(const uint32_t* __restrict__ parent, const SegRec* __restrict__ rec, uint32_t n1,
                          uint32_t n2, const uint8_t* __restrict__ keep, double line_threshold,
                          double prune_threshold, uint8_t* __restrict__ out, int64_t out_stride2,
                          unsigned int* conflict) {
    __shared__ uint8_t skeep[256];
    if (threadIdx.x < 256) skeep[threadIdx.x] = keep[threadIdx.x];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n1 * n2) return;
    const uint32_t y = (uint32_t)(i / n1), x = (uint32_t)(i % n1);
    const SegRec r = rec[comp_id(parent, (uint32_t)i)];
    out[(int64_t)y * out_stride2 + x] = (uint8_t)rule(r, skeep, line_threshold, prune_threshold, conflict);
}

}  // namespace

size_t mr_prune_parent_bytes(int64_t n1, int64_t n2) { return (size_t)n1 * (size_t)n2 * sizeof(uint32_t); }
size_t mr_prune_rec_bytes(unsigned int n_segments) { return (size_t)n_segments * sizeof(SegRec); }

// phase 1: components; leaves the number of components in *d_n_roots (device)
cudaError_t mr_launch_prune_components(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                                       uint32_t* parent, unsigned int* d_n_roots, cudaStream_t s,
                                       int* n_launches) {
    const uint64_t npx = (uint64_t)n1 * (uint64_t)n2;
    const uint64_t groups = (uint64_t)((n1 + 31) / 32) * (uint64_t)n2;
    const unsigned gb = (unsigned)((groups + 7) / 8);            // 8 warps per block
    const unsigned pb = (unsigned)((npx + 255) / 256);
    cudaError_t e = cudaMemsetAsync(d_n_roots, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    k_runs<<<gb, 256, 0, s>>>(labels, stride2, (uint32_t)n1, (uint32_t)n2, parent);
    if (n2 > 1) {
        const uint64_t nv = (uint64_t)n1 * (uint64_t)(n2 - 1);
        k_vmerge<<<(unsigned)((nv + 255) / 256), 256, 0, s>>>(labels, stride2, (uint32_t)n1, (uint32_t)n2, parent);
        *n_launches += 1;
    }
    k_flatten<<<pb, 256, 0, s>>>(parent, npx);
    k_number<<<pb, 256, 0, s>>>(parent, npx, d_n_roots);
    *n_launches += 3;
    return cudaGetLastError();
}

// phase 2: statistics, rule, output.  *d_conflict (device) becomes non-zero when a segment holds
// two different known categories.
cudaError_t mr_launch_prune_apply(const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                                  const uint32_t* parent, void* rec, unsigned int n_segments,
                                  int64_t n_known, const int64_t* known_idx, const uint8_t* known_cat,
                                  const uint8_t* keep_dev, double line_threshold, double prune_threshold,
                                  uint8_t* out, int64_t out_stride2, unsigned int* d_conflict, cudaStream_t s,
                                  int* n_launches) {
    const uint64_t npx = (uint64_t)n1 * (uint64_t)n2;
    const uint64_t groups = (uint64_t)((n1 + 31) / 32) * (uint64_t)n2;
    SegRec* r = reinterpret_cast<SegRec*>(rec);
    cudaError_t e = cudaMemsetAsync(d_conflict, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    k_init_recs<<<(n_segments + 255) / 256, 256, 0, s>>>(r, n_segments);
    k_stats<<<(unsigned)((groups + 7) / 8), 256, 0, s>>>(labels, stride2, (uint32_t)n1, (uint32_t)n2, parent, r);
    *n_launches += 2;
    if (n_known > 0) {
        k_known<<<(unsigned)((n_known + 255) / 256), 256, 0, s>>>(known_idx, known_cat, n_known, npx, parent, r);
        *n_launches += 1;
    }
    k_relabel<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(parent, r, (uint32_t)n1, (uint32_t)n2, keep_dev,
                                                            line_threshold, prune_threshold, out, out_stride2,
                                                            d_conflict);
    *n_launches += 1;
    return cudaGetLastError();
}
