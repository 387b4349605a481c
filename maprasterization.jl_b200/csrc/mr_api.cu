// mr_api.cu -- the C ABI of libmaprast.so (include/maprast.h): context, palette upload,
// host/device entry points, chunked H2D -> kernel -> D2H pipeline for host callers.
// No CPU compute path exists: every label is produced by a CUDA kernel.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/maprast.h"
#include "mr_internal.h"

struct mr_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t s_main = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    // palette
    bool palette_set = false;
    MrPalette pal{};
    double* d_pal = nullptr;        // mean | lo | hi | lin
    uint8_t* d_lut = nullptr;
    uint8_t* d_coarse = nullptr;
    double* d_qtab = nullptr;       // squared channel differences per category (table build scratch)
    bool lut_valid = false;
    float build_ms = 0.f;
    // known-category plane (sparse)
    int64_t n_known = 0;
    int64_t known_origin = 0;       // sheet line of line 0 of the px of the next calls (bands); 0 for whole images
    int64_t* d_known_idx = nullptr;
    uint8_t* d_known_cat = nullptr;
    // counters
    MrCountersDev* d_counters = nullptr;
    MrWorkRing work;                // item queues of the fast kernel
    bool count_on = false;
    int64_t launches = 0;
    // grow-only device scratch for the host entry points
    uint8_t* d_in = nullptr;  size_t d_in_cap = 0;
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    uint8_t* d_tmp = nullptr; size_t d_tmp_cap = 0;
    // repacking scratch of the device entry points: lines that are not 16-byte aligned are copied
    // into aligned pitches so that they can take the fast kernel
    uint8_t* d_rp_in = nullptr;  size_t d_rp_in_cap = 0;
    uint8_t* d_rp_out = nullptr; size_t d_rp_out_cap = 0;
    uint8_t* d_lab = nullptr;    size_t d_lab_cap = 0;     // categorise-only labels of the RGBA path
    uint8_t* d_f64 = nullptr;    size_t d_f64_cap = 0;     // RGB{Float64} scratch images of blur (ping-pong, host staging)
    uint8_t* d_f64b = nullptr;   size_t d_f64b_cap = 0;
    double* d_tex = nullptr;     bool tex_set = false;     // texture statistics of mr_set_texture_stats (9 K doubles)
    // segment prune scratch: parent array, component records, keep table + two counters
    uint8_t* d_parent = nullptr; size_t d_parent_cap = 0;   // first run of every (line, chunk)
    uint8_t* d_runs = nullptr; size_t d_runs_cap = 0;       // run arrays
    uint8_t* d_rec = nullptr; size_t d_rec_cap = 0;
    uint8_t* d_stage = nullptr; size_t d_stage_cap = 0;     // per-chunk staging of the run-count pass
    uint8_t* d_poly = nullptr; size_t d_poly_cap = 0;       // packed rings of mr_polygons_*
    uint8_t* d_small = nullptr;     // 256-byte keep table | n_roots | conflict
    std::vector<cudaEvent_t> ev_up, ev_done;
    std::string err;
};

static thread_local std::string g_err;

static int fail(mr_ctx* ctx, int code, const std::string& msg) {
    g_err = msg;
    if (ctx) ctx->err = msg;
    return code;
}
static int fail_cuda(mr_ctx* ctx, cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "CUDA error in %s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return fail(ctx, MR_ERR_CUDA, buf);
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);   \
    } while (0)

extern "C" int mr_version(void) { return 100; }

extern "C" const char* mr_last_error(const mr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" mr_ctx* mr_create(int device, int* err) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
        (void)cudaGetLastError();
        fail(nullptr, MR_ERR_CUDA,
             "mr_create: no usable CUDA device (libmaprast has no CPU fallback)");
        if (err) *err = MR_ERR_CUDA;
        return nullptr;
    }
    mr_ctx* ctx = new mr_ctx();
    ctx->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    cudaDeviceProp prop;
    ok = ok && cudaGetDeviceProperties(&prop, device) == cudaSuccess;
    if (ok && prop.major < 10) {
        fail(nullptr, MR_ERR_CUDA, "mr_create: device is not sm_100+ (library is built for sm_100a only)");
        ok = false;
    }
    ok = ok && cudaStreamCreateWithFlags(&ctx->s_main, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_lut, MR_LUT_SIZE) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_coarse, MR_COARSE_SIZE) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_qtab, MR_QTAB_BYTES) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_counters, sizeof(MrCountersDev)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_counters, 0, sizeof(MrCountersDev)) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->work.slots, MR_WORK_SLOTS * sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->work.slots, 0, MR_WORK_SLOTS * sizeof(unsigned int)) == cudaSuccess;
    if (!ok) {
        if (g_err.empty()) fail(nullptr, MR_ERR_CUDA, "mr_create: CUDA initialisation failed");
        if (err) *err = MR_ERR_CUDA;
        mr_destroy(ctx);
        return nullptr;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (err) *err = MR_OK;
    return ctx;
}

extern "C" void mr_destroy(mr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto e : ctx->ev_up) cudaEventDestroy(e);
    for (auto e : ctx->ev_done) cudaEventDestroy(e);
    cudaFree(ctx->d_pal); cudaFree(ctx->d_lut); cudaFree(ctx->d_coarse); cudaFree(ctx->d_qtab);
    cudaFree(ctx->d_known_idx); cudaFree(ctx->d_known_cat); cudaFree(ctx->d_counters); cudaFree(ctx->work.slots);
    cudaFree(ctx->d_in); cudaFree(ctx->d_out); cudaFree(ctx->d_tmp); cudaFree(ctx->d_rp_in); cudaFree(ctx->d_rp_out); cudaFree(ctx->d_lab); cudaFree(ctx->d_f64); cudaFree(ctx->d_f64b); cudaFree(ctx->d_tex);
    cudaFree(ctx->d_parent); cudaFree(ctx->d_runs); cudaFree(ctx->d_rec); cudaFree(ctx->d_stage); cudaFree(ctx->d_poly); cudaFree(ctx->d_small);
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    (void)cudaGetLastError();
    delete ctx;
}

extern "C" int mr_set_palette(mr_ctx* ctx, int K, const double* mean, const double* lo, const double* hi,
                              int colourspace, int channels) {
    if (!ctx) return fail(nullptr, MR_ERR_ARG, "mr_set_palette: null ctx");
    if (K < 1 || K > 254) return fail(ctx, MR_ERR_ARG, "mr_set_palette: K must be in 1..254 (uint8 labels, 255 reserved)");
    if (channels != 3 && channels != 4) return fail(ctx, MR_ERR_ARG, "mr_set_palette: channels must be 3 or 4");
    if (colourspace != MR_COLOURSPACE_RGB && colourspace != MR_COLOURSPACE_LAB)
        return fail(ctx, MR_ERR_ARG, "mr_set_palette: unknown colourspace");
    if (colourspace == MR_COLOURSPACE_LAB && channels != 3)
        return fail(ctx, MR_ERR_ARG, "mr_set_palette: Lab matching takes 3-channel statistics");
    if (!mean || !lo || !hi) return fail(ctx, MR_ERR_ARG, "mr_set_palette: null statistics");
    bool any_finite = false;
    for (int c = 0; c < K; ++c) {
        bool fin = true, inf = true;
        for (int ch = 0; ch < channels; ++ch) {
            double m = mean[c * channels + ch];
            if (isnan(m)) return fail(ctx, MR_ERR_ARG, "mr_set_palette: NaN category mean");
            fin = fin && isfinite(m);
            inf = inf && (isinf(m) && m > 0);
        }
        if (!fin && !inf)
            return fail(ctx, MR_ERR_ARG, "mr_set_palette: a category mean must be all finite or all +Inf (missing)");
        any_finite = any_finite || fin;
    }
    // the reference indexes `missing` and throws when no category has points (categories.jl:80-81)
    if (!any_finite) return fail(ctx, MR_ERR_ARG, "mr_set_palette: all categories are missing");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)K * channels;
    std::vector<double> h(3 * n + 256);
    memcpy(h.data(), mean, n * 8);
    memcpy(h.data() + n, lo, n * 8);
    memcpy(h.data() + 2 * n, hi, n * 8);
    for (int i = 0; i < 256; ++i) {   // sRGB inverse companding (Appendix A.4), host libm
        double v = (double)i / 255.0;
        h[3 * n + i] = (v <= 0.04045) ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4);
    }
    CK(cudaStreamSynchronize(ctx->s_main));
    // one allocation for the largest palette (254 categories x 4 channels x 3 arrays + the sRGB table),
    // made once: a palette change (slider move -> mr_set_palette) costs no cudaMalloc / cudaFree
    if (!ctx->d_pal) CK(cudaMalloc(&ctx->d_pal, ((size_t)254 * 4 * 3 + 256) * 8));
    CK(cudaMemcpy(ctx->d_pal, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    ctx->pal.mean = ctx->d_pal;
    ctx->pal.lo = ctx->d_pal + n;
    ctx->pal.hi = ctx->d_pal + 2 * n;
    ctx->pal.lin = ctx->d_pal + 3 * n;
    ctx->pal.K = K;
    ctx->pal.C = channels;
    ctx->pal.colourspace = colourspace;
    ctx->palette_set = true;
    ctx->tex_set = false;            // texture statistics belong to a palette
    ctx->lut_valid = false;
    ctx->build_ms = 0.f;
    {   // exact label table over the 2^24 colours (4-channel statistics: alpha = 255 in the distance)
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        int nl = 0;
        CK(cudaEventRecord(a, ctx->s_main));
        CK(mr_launch_build_lut(ctx->pal, ctx->d_lut, ctx->d_coarse, ctx->d_qtab, ctx->s_main, &nl));
        CK(cudaEventRecord(b, ctx->s_main));
        CK(cudaStreamSynchronize(ctx->s_main));
        CK(cudaEventElapsedTime(&ctx->build_ms, a, b));
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        ctx->launches += nl;
        ctx->lut_valid = true;
    }
    return MR_OK;
}

extern "C" int mr_set_known_band(mr_ctx* ctx, int64_t n, const int64_t* lin_idx, const uint8_t* cat, int64_t first_line) {
    if (first_line < 0) return fail(ctx, MR_ERR_ARG, "mr_set_known_band: negative first_line");
    const int rc = mr_set_known(ctx, n, lin_idx, cat);
    if (rc == MR_OK) ctx->known_origin = first_line;
    return rc;
}

extern "C" int mr_set_known(mr_ctx* ctx, int64_t n, const int64_t* lin_idx, const uint8_t* cat) {
    if (!ctx) return fail(nullptr, MR_ERR_ARG, "mr_set_known: null ctx");
    if (n < 0 || (n > 0 && (!lin_idx || !cat))) return fail(ctx, MR_ERR_ARG, "mr_set_known: bad arguments");
    for (int64_t k = 0; k < n; ++k)
        if (lin_idx[k] < 0) return fail(ctx, MR_ERR_ARG, "mr_set_known: negative linear index");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->s_main));
    cudaFree(ctx->d_known_idx); ctx->d_known_idx = nullptr;
    cudaFree(ctx->d_known_cat); ctx->d_known_cat = nullptr;
    ctx->n_known = 0;
    ctx->known_origin = 0;
    if (n == 0) return MR_OK;
    CK(cudaMalloc(&ctx->d_known_idx, n * 8));
    CK(cudaMalloc(&ctx->d_known_cat, n));
    CK(cudaMemcpy(ctx->d_known_idx, lin_idx, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_known_cat, cat, n, cudaMemcpyHostToDevice));
    ctx->n_known = n;
    return MR_OK;
}

extern "C" int mr_get_lut(mr_ctx* ctx, uint8_t* lut_host) {
    if (!ctx || !lut_host) return fail(ctx, MR_ERR_ARG, "mr_get_lut: null argument");
    if (!ctx->lut_valid) return fail(ctx, MR_ERR_STATE, "mr_get_lut: no 3-channel palette set");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(lut_host, ctx->d_lut, MR_LUT_SIZE, cudaMemcpyDeviceToHost));
    return MR_OK;
}

extern "C" int mr_palette_build_ms(mr_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, MR_ERR_ARG, "mr_palette_build_ms: null argument");
    *ms = ctx->build_ms;
    return MR_OK;
}

extern "C" int mr_enable_counters(mr_ctx* ctx, int on) {
    if (!ctx) return fail(nullptr, MR_ERR_ARG, "mr_enable_counters: null ctx");
    ctx->count_on = on != 0;
    return MR_OK;
}

extern "C" int mr_read_counters(mr_ctx* ctx, mr_counters* out, int reset) {
    if (!ctx || !out) return fail(ctx, MR_ERR_ARG, "mr_read_counters: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    MrCountersDev h;
    CK(cudaMemcpy(&h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    out->n_pixels = h.n_pixels; out->n_nomatch = h.n_nomatch;
    out->n_fine = h.n_fine; out->n_changed = h.n_changed;
    if (reset) CK(cudaMemset(ctx->d_counters, 0, sizeof h));
    return MR_OK;
}

extern "C" int64_t mr_launch_count(const mr_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" void* mr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void mr_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ job plumbing
static int check_image_args(mr_ctx* ctx, const char* fn, const void* in, const void* out, int64_t n1,
                            int64_t n2, int64_t stride2, int bpp, int64_t out_stride2, int R,
                            int halo_lo, int halo_hi) {
    std::string f(fn);
    if (!ctx) return fail(nullptr, MR_ERR_ARG, f + ": null ctx");
    if (n1 < 0 || n2 < 0) return fail(ctx, MR_ERR_ARG, f + ": negative size");
    if (n1 > 0 && n2 > 0 && (!in || !out)) return fail(ctx, MR_ERR_ARG, f + ": null buffer");
    if (stride2 < n1 * bpp || out_stride2 < n1) return fail(ctx, MR_ERR_ARG, f + ": stride smaller than a line");
    if (R < 0 || R > 7) return fail(ctx, MR_ERR_ARG, f + ": window radius R must be in 0..7");
    if (halo_lo < 0 || halo_hi < 0 || halo_lo > R || halo_hi > R)
        return fail(ctx, MR_ERR_ARG, f + ": halo_lo/halo_hi must be in 0..R");
    return MR_OK;
}

static int check_palette(mr_ctx* ctx, const char* fn, int bpp) {
    std::string f(fn);
    if (!ctx->palette_set) return fail(ctx, MR_ERR_STATE, f + ": mr_set_palette has not been called");
    if (bpp != 3 && bpp != 4) return fail(ctx, MR_ERR_ARG, f + ": bytes_per_px must be 3 (RGB{N0f8}) or 4 (RGBA{N0f8})");
    if (bpp == 3 && ctx->pal.C != 3)
        return fail(ctx, MR_ERR_ARG, f + ": 4-channel statistics need RGBA pixels");
    if (bpp == 4 && ctx->pal.C != 4 && ctx->pal.colourspace == MR_COLOURSPACE_RGB)
        return fail(ctx, MR_ERR_ARG, f + ": RGBA pixels need 4-channel statistics (categories.jl:92)");
    return MR_OK;
}

static MrJob make_job(mr_ctx* ctx, const uint8_t* px, uint8_t* out, int64_t n1, int64_t n2, int64_t stride2,
                      int bpp, int R, int halo_lo, int halo_hi, int64_t out_stride2) {
    MrJob j{};
    j.px = px; j.out = out; j.n1 = n1; j.n2 = n2; j.stride2 = stride2; j.out_stride2 = out_stride2;
    j.sheet_stride = 0; j.out_sheet_stride = 0; j.n_sheets = 1; j.bpp = bpp; j.R = R;
    j.halo_lo = halo_lo; j.halo_hi = halo_hi; j.padded = 0;
    j.lut = (bpp == 3 && ctx->lut_valid) ? ctx->d_lut : nullptr;
    j.coarse = ctx->d_coarse;
    j.pal = ctx->pal;
    j.counters = ctx->count_on ? ctx->d_counters : nullptr;
    return j;
}

static int grow(mr_ctx* ctx, uint8_t** p, size_t* cap, size_t need) {
    if (*cap >= need) return MR_OK;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MR_ERR_NOMEM, "device allocation failed"); }
    *cap = need;
    return MR_OK;
}

// fused (or categorise-only) on device-resident data
// A job that would take the fast kernel if only its lines were 16-byte aligned (arbitrary n1, tight
// or odd strides, unaligned base pointers): copy it into aligned pitches (device to device, ~2x the
// algorithmic traffic) and run the fast kernel there -- still an order of magnitude faster than the
// generic kernel.
static bool repack_pays(const MrJob& job) {
    if (job.n_sheets != 1 || job.halo_lo_px || job.halo_hi_px || job.R > 3) return false;
    if (mr_fast_supported(job)) return false;
    MrJob a = job;   // the same job with aligned geometry
    a.px = a.out = reinterpret_cast<uint8_t*>(uintptr_t(256));
    a.stride2 = (job.n1 * job.bpp + 15) / 16 * 16;
    a.out_stride2 = (job.n1 + 15) / 16 * 16;
    a.padded = 1;
    return mr_fast_supported(a);
}

static int run_repacked(mr_ctx* ctx, const MrJob& job, cudaStream_t s, int* nl) {
    const int64_t T = job.n2 + job.halo_lo + job.halo_hi;
    const int64_t line_in = (job.n1 * job.bpp + 15) / 16 * 16, line_out = (job.n1 + 15) / 16 * 16;
    int rc;
    if ((rc = grow(ctx, &ctx->d_rp_in, &ctx->d_rp_in_cap, (size_t)line_in * T))) return rc;
    if ((rc = grow(ctx, &ctx->d_rp_out, &ctx->d_rp_out_cap, (size_t)line_out * job.n2))) return rc;
    CK(cudaMemcpy2DAsync(ctx->d_rp_in, line_in, job.px - (int64_t)job.halo_lo * job.stride2, job.stride2,
                         job.n1 * job.bpp, T, cudaMemcpyDeviceToDevice, s));
    MrJob a = job;
    a.px = ctx->d_rp_in + (int64_t)job.halo_lo * line_in;
    a.stride2 = line_in;
    a.out = ctx->d_rp_out;
    a.out_stride2 = line_out;
    a.padded = 1;
    cudaError_t e = mr_launch_fast(a, ctx->sm_count, s, nl, &ctx->work);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "fused kernel launch (repacked lines)");
    CK(cudaMemcpy2DAsync(job.out, job.out_stride2, ctx->d_rp_out, line_out, job.n1, job.n2, cudaMemcpyDeviceToDevice, s));
    return MR_OK;
}

static int run_rgba(mr_ctx* ctx, const MrJob& job, cudaStream_t s);

static int run_fused(mr_ctx* ctx, const MrJob& job, cudaStream_t s) {
    if (job.n1 == 0 || job.n2 == 0 || job.n_sheets == 0) return MR_OK;
    if (job.bpp == 4 && ctx->lut_valid && job.n_sheets == 1 && job.R <= 3 && job.pal.K <= 126 && !job.counters &&
        !job.halo_lo_px && !job.halo_hi_px)
        return run_rgba(ctx, job, s);
    int nl = 0;
    cudaError_t e = cudaErrorNotSupported;
    if (mr_fast_supported(job)) {
        e = mr_launch_fast(job, ctx->sm_count, s, &nl, &ctx->work);
    } else if (repack_pays(job)) {
        const int rc = run_repacked(ctx, job, s, &nl);
        ctx->launches += nl;
        return rc;
    }
    if (e == cudaErrorNotSupported) e = mr_launch_generic(job, false, ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "fused kernel launch");
    return MR_OK;
}

// standalone clean (labels in -> labels out).  max_label < 0: not known yet (one reduction pass +
// read-back picks the plane count of the fast kernel); *max_label_out returns it for the next pass.
static int run_clean(mr_ctx* ctx, const MrJob& job, cudaStream_t s, int max_label = -1, int* max_label_out = nullptr) {
    if (job.n1 == 0 || job.n2 == 0) return MR_OK;
    int nl = 0;
    MrJob j = job;
    j.bpp = 1;
    j.lut = nullptr;
    j.pal.K = 0;
    cudaError_t e = cudaErrorNotSupported;
    const bool direct = mr_fast_supported(j), repack = !direct && repack_pays(j);
    if (direct || repack) {   // geometry qualifies (possibly after repacking): the plane count needs the largest label
        if (max_label < 0) {
            if (!ctx->d_small) CK(cudaMalloc(&ctx->d_small, 2048));
            unsigned int* d_max = reinterpret_cast<unsigned int*>(ctx->d_small + 264);
            e = mr_launch_max_label(job.px - (int64_t)job.halo_lo * job.stride2, job.n1, job.n2 + job.halo_lo + job.halo_hi,
                                    job.stride2, d_max, ctx->sm_count, s, &nl);
            if (e != cudaSuccess) return fail_cuda(ctx, e, "label maximum");
            unsigned int m = 0;
            CK(cudaMemcpyAsync(&m, d_max, sizeof m, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            max_label = (int)m;
        }
        if (max_label > 254)
            return fail(ctx, MR_ERR_ARG, "clean: label 255 is reserved (labels are 0 = no match, 1..254 = category)");
        j.pal.K = max_label;
        e = cudaErrorNotSupported;
        if (max_label <= 126) {
            if (direct) {
                e = mr_launch_fast(j, ctx->sm_count, s, &nl, &ctx->work);
            } else {
                const int rc = run_repacked(ctx, j, s, &nl);
                ctx->launches += nl;
                if (max_label_out) *max_label_out = max_label;
                return rc;
            }
        }
    }
    if (max_label_out) *max_label_out = max_label;
    if (e == cudaErrorNotSupported) e = mr_launch_generic(job, true, ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "clean kernel launch");
    return MR_OK;
}
// RGBA{N0f8} sheets on the fast kernels: strip the alpha byte into packed RGB (aligned pitch),
// categorise every line (own + halo) with the label table (alpha = 255), re-evaluate the pixels
// that are not opaque directly, then run the standalone majority filter on the labels.
// (Known-category points: run_with_known calls this with R = 0 and does its fix-up in between.)
static int run_rgba(mr_ctx* ctx, const MrJob& job, cudaStream_t s) {
    const int64_t T = job.n2 + job.halo_lo + job.halo_hi;
    const int64_t line_rgb = (job.n1 * 3 + 15) / 16 * 16, line_lab = (job.n1 + 15) / 16 * 16;
    int rc;
    if ((rc = grow(ctx, &ctx->d_rp_in, &ctx->d_rp_in_cap, (size_t)line_rgb * T))) return rc;
    if ((rc = grow(ctx, &ctx->d_lab, &ctx->d_lab_cap, (size_t)line_lab * T))) return rc;
    if (!ctx->d_small) CK(cudaMalloc(&ctx->d_small, 2048));
    unsigned int* d_special = reinterpret_cast<unsigned int*>(ctx->d_small + 268);
    const uint8_t* px0 = job.px - (int64_t)job.halo_lo * job.stride2;
    int nl = 0;
    cudaError_t e = mr_launch_rgba_strip(px0, job.stride2, job.n1, T, ctx->d_rp_in, line_rgb, d_special, s, &nl);
    if (e == cudaSuccess) {
        MrJob c = job;   // categorise only, all T lines as one sheet
        c.px = ctx->d_rp_in; c.stride2 = line_rgb; c.bpp = 3; c.R = 0; c.n2 = T; c.halo_lo = c.halo_hi = 0;
        c.out = ctx->d_lab; c.out_stride2 = line_lab; c.padded = 1; c.lut = ctx->d_lut; c.counters = nullptr;
        e = mr_launch_fast(c, ctx->sm_count, s, &nl, &ctx->work);
    }
    if (e == cudaSuccess)
        e = mr_launch_alpha_fix(job.pal, px0, job.stride2, job.n1, T, ctx->d_lab, line_lab, d_special, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "RGBA front end");
    uint8_t* own = ctx->d_lab + (int64_t)job.halo_lo * line_lab;
    if (job.R == 0) {
        CK(cudaMemcpy2DAsync(job.out, job.out_stride2, own, line_lab, job.n1, job.n2, cudaMemcpyDeviceToDevice, s));
        return MR_OK;
    }
    MrJob c = job;
    c.px = own; c.stride2 = line_lab; c.bpp = 1; c.lut = nullptr;
    return run_clean(ctx, c, s, ctx->pal.K);
}

// Known-category points (categories.jl:110-112) change single labels BEFORE the majority filter sees
// them: categorise every line the windows read (the band's own lines and its halo lines, wherever they
// live) into an aligned label plane, fix the clicked pixels up, then run the standalone filter on it.
static int run_with_known(mr_ctx* ctx, MrJob job, cudaStream_t s) {
    const int R = job.R;
    const int hl = R > 0 ? job.halo_lo : 0, hh = R > 0 ? job.halo_hi : 0;
    const int64_t T = job.n2 + hl + hh;
    const int64_t line_tmp = (job.n1 + 15) / 16 * 16;   // aligned intermediate labels: both passes can take the fast kernel
    uint8_t* plane = job.out;
    int64_t plane_stride = job.out_stride2;
    if (R > 0) {
        int rc = grow(ctx, &ctx->d_tmp, &ctx->d_tmp_cap, (size_t)line_tmp * T);
        if (rc) return rc;
        plane = ctx->d_tmp;
        plane_stride = line_tmp;
    }
    // categorise-only passes: [halo lines before | own lines | halo lines after] -> plane
    struct Piece { const uint8_t* px; int64_t lines, at; };
    Piece pieces[3];
    int np = 0;
    if (job.halo_lo_px || job.halo_hi_px) {
        if (hl) pieces[np++] = Piece{job.halo_lo_px ? job.halo_lo_px + (int64_t)(job.halo_lo - hl) * job.stride2
                                                    : job.px - (int64_t)hl * job.stride2, hl, 0};
        pieces[np++] = Piece{job.px, job.n2, hl};
        if (hh) pieces[np++] = Piece{job.halo_hi_px ? job.halo_hi_px : job.px + job.n2 * job.stride2, hh, hl + job.n2};
    } else {
        pieces[np++] = Piece{job.px - (int64_t)hl * job.stride2, T, 0};
    }
    for (int k = 0; k < np; ++k) {
        MrJob c = job;
        c.px = pieces[k].px; c.n2 = pieces[k].lines; c.R = 0; c.halo_lo = c.halo_hi = 0;
        c.halo_lo_px = c.halo_hi_px = nullptr;
        c.out = plane + pieces[k].at * plane_stride; c.out_stride2 = plane_stride;
        if (R > 0) c.counters = nullptr;
        int rc = run_fused(ctx, c, s);
        if (rc) return rc;
    }
    int nl = 0;
    cudaError_t e = mr_launch_known_fixup(job, ctx->n_known, ctx->d_known_idx, ctx->d_known_cat, plane, plane_stride,
                                          ctx->known_origin, -(int64_t)hl, T, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "known-category fix-up");
    if (R > 0) {
        MrJob c = job;
        c.px = plane + (int64_t)hl * plane_stride; c.stride2 = plane_stride; c.bpp = 1;
        c.halo_lo = hl; c.halo_hi = hh; c.halo_lo_px = c.halo_hi_px = nullptr;
        return run_clean(ctx, c, s, ctx->pal.K);
    }
    return MR_OK;
}

extern "C" int mr_categorize_clean_dev(mr_ctx* ctx, const uint8_t* px_dev, int64_t n1, int64_t n2,
                                       int64_t stride2, int bpp, int R, int halo_lo, int halo_hi,
                                       uint8_t* labels_dev, int64_t out_stride2, void* stream) {
    int rc = check_image_args(ctx, "mr_categorize_clean_dev", px_dev, labels_dev, n1, n2, stride2, bpp,
                              out_stride2, R, halo_lo, halo_hi);
    if (rc) return rc;
    if ((rc = check_palette(ctx, "mr_categorize_clean_dev", bpp))) return rc;
    CK(cudaSetDevice(ctx->device));
    MrJob job = make_job(ctx, px_dev, labels_dev, n1, n2, stride2, bpp, R, halo_lo, halo_hi, out_stride2);
    if (ctx->n_known > 0) return run_with_known(ctx, job, (cudaStream_t)stream);
    return run_fused(ctx, job, (cudaStream_t)stream);
}

extern "C" int mr_batch_categorize_clean_dev(mr_ctx* ctx, int n_sheets, const uint8_t* px_dev,
                                             int64_t sheet_stride, int64_t n1, int64_t n2, int64_t stride2,
                                             int bpp, int R, uint8_t* labels_dev, int64_t out_sheet_stride,
                                             int64_t out_stride2, void* stream) {
    int rc = check_image_args(ctx, "mr_batch_categorize_clean_dev", px_dev, labels_dev, n1, n2, stride2, bpp,
                              out_stride2, R, 0, 0);
    if (rc) return rc;
    if ((rc = check_palette(ctx, "mr_batch_categorize_clean_dev", bpp))) return rc;
    if (n_sheets < 0 || sheet_stride < stride2 * n2 || out_sheet_stride < out_stride2 * n2)
        return fail(ctx, MR_ERR_ARG, "mr_batch_categorize_clean_dev: bad sheet count/strides");
    if (ctx->n_known > 0)
        return fail(ctx, MR_ERR_ARG, "mr_batch_categorize_clean_dev: known-category points are per-sheet; clear them (mr_set_known n=0)");
    CK(cudaSetDevice(ctx->device));
    MrJob job = make_job(ctx, px_dev, labels_dev, n1, n2, stride2, bpp, R, 0, 0, out_stride2);
    job.n_sheets = n_sheets;
    job.sheet_stride = sheet_stride;
    job.out_sheet_stride = out_sheet_stride;
    return run_fused(ctx, job, (cudaStream_t)stream);
}

extern "C" int mr_clean_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                            int R, int halo_lo, int halo_hi, uint8_t* out_dev, int64_t out_stride2,
                            void* stream) {
    int rc = check_image_args(ctx, "mr_clean_dev", labels_dev, out_dev, n1, n2, stride2, 1, out_stride2, R,
                              halo_lo, halo_hi);
    if (rc) return rc;
    if (labels_dev == out_dev && n1 > 0 && n2 > 0 && R > 0)
        return fail(ctx, MR_ERR_ARG, "mr_clean_dev: in-place operation is not supported");
    CK(cudaSetDevice(ctx->device));
    MrJob job = make_job(ctx, labels_dev, out_dev, n1, n2, stride2, 1, R, halo_lo, halo_hi, out_stride2);
    job.lut = nullptr;
    return run_clean(ctx, job, (cudaStream_t)stream);
}

// Same with an upper bound of the labels supplied by the caller (e.g. the K of the palette that produced
// them): no label-maximum pass and no stream synchronisation -- fully asynchronous.
extern "C" int mr_clean_dev_bounded(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                                    int R, int halo_lo, int halo_hi, int max_label, uint8_t* out_dev,
                                    int64_t out_stride2, void* stream) {
    int rc = check_image_args(ctx, "mr_clean_dev_bounded", labels_dev, out_dev, n1, n2, stride2, 1, out_stride2, R,
                              halo_lo, halo_hi);
    if (rc) return rc;
    if (max_label < 0 || max_label > 254)
        return fail(ctx, MR_ERR_ARG, "mr_clean_dev_bounded: max_label must be in 0..254");
    if (labels_dev == out_dev && n1 > 0 && n2 > 0 && R > 0)
        return fail(ctx, MR_ERR_ARG, "mr_clean_dev_bounded: in-place operation is not supported");
    CK(cudaSetDevice(ctx->device));
    MrJob job = make_job(ctx, labels_dev, out_dev, n1, n2, stride2, 1, R, halo_lo, halo_hi, out_stride2);
    job.lut = nullptr;
    return run_clean(ctx, job, (cudaStream_t)stream, max_label);
}

// ------------------------------------------------------------------ host entry points
static int ensure_events(mr_ctx* ctx, size_t n) {
    while (ctx->ev_up.size() < n) {
        cudaEvent_t a, b;
        CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ctx->ev_up.push_back(a);
        ctx->ev_done.push_back(b);
    }
    return MR_OK;
}

// Chunked pipeline: every input line crosses PCIe once (into one resident device image);
// chunk c's kernel waits for the upload of the last line its windows read, its labels go back
// on a third stream.  H2D, kernel and D2H of different chunks overlap.
// px points at line 0 of the band; lines [-halo_lo, n2 + halo_hi) are read from the host.
static int host_impl(mr_ctx* ctx, const char* fn, const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2,
                     int bpp, int R, int halo_lo, int halo_hi, uint8_t* labels_out, int64_t out_stride2,
                     mr_counters* counters) {
    int rc = check_image_args(ctx, fn, px, labels_out, n1, n2, stride2, bpp, out_stride2, R, halo_lo, halo_hi);
    if (rc) return rc;
    if ((rc = check_palette(ctx, fn, bpp))) return rc;
    if (n1 == 0 || n2 == 0) {
        if (counters) memset(counters, 0, sizeof *counters);
        return MR_OK;
    }
    CK(cudaSetDevice(ctx->device));
    const int64_t T = n2 + halo_lo + halo_hi;            // lines resident on the device
    const int64_t line_in = (n1 * bpp + 15) / 16 * 16;   // 16-byte aligned device lines
    const int64_t line_out = (n1 + 15) / 16 * 16;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)line_in * T))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, (size_t)line_out * n2))) return rc;
    const bool saved_count = ctx->count_on;
    if (counters) {
        ctx->count_on = true;
        CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(MrCountersDev), ctx->s_main));
    }
    int64_t chunk_lines = T;
    if (ctx->n_known == 0) {
        const int64_t target = (int64_t)48 << 20;   // ~48 MiB of input per chunk
        chunk_lines = target / line_in;
        if (chunk_lines < 4 * (R + 1)) chunk_lines = 4 * (R + 1);
        if (chunk_lines > T) chunk_lines = T;
    }
    const int64_t n_up = (T + chunk_lines - 1) / chunk_lines;      // upload pieces over [0, T)
    const int64_t n_k = (n2 + chunk_lines - 1) / chunk_lines;      // kernel chunks over [0, n2)
    if ((rc = ensure_events(ctx, (size_t)(n_up > n_k ? n_up : n_k)))) { ctx->count_on = saved_count; return rc; }
    const uint8_t* host0 = px - (int64_t)halo_lo * stride2;        // device line 0
    uint8_t* d_own = ctx->d_in + (int64_t)halo_lo * line_in;
    rc = MR_OK;
    for (int64_t c = 0; c < n_up && rc == MR_OK; ++c) {
        const int64_t a = c * chunk_lines, b = (a + chunk_lines < T) ? a + chunk_lines : T;
        cudaError_t e = cudaMemcpy2DAsync(ctx->d_in + a * line_in, line_in, host0 + a * stride2, stride2,
                                          n1 * bpp, b - a, cudaMemcpyHostToDevice, ctx->s_h2d);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_up[c], ctx->s_h2d);
        if (e != cudaSuccess) rc = fail_cuda(ctx, e, "H2D copy");
    }
    for (int64_t c = 0; c < n_k && rc == MR_OK; ++c) {
        const int64_t a = c * chunk_lines, b = (a + chunk_lines < n2) ? a + chunk_lines : n2;
        const int hl = (int)((a + halo_lo < R) ? a + halo_lo : R);
        const int hh = (int)((n2 + halo_hi - b < R) ? n2 + halo_hi - b : R);
        const int64_t last_dev_line = halo_lo + b + hh - 1;         // last line this chunk's windows read
        cudaError_t e = cudaStreamWaitEvent(ctx->s_main, ctx->ev_up[last_dev_line / chunk_lines], 0);
        if (e != cudaSuccess) { rc = fail_cuda(ctx, e, "stream wait"); break; }
        MrJob job = make_job(ctx, d_own + a * line_in, ctx->d_out + a * line_out, n1, b - a, line_in, bpp,
                             R, hl, hh, line_out);
        job.padded = 1;   // d_in / d_out lines are padded to 16 bytes by construction
        rc = (ctx->n_known > 0) ? run_with_known(ctx, job, ctx->s_main) : run_fused(ctx, job, ctx->s_main);
        if (rc) break;
        e = cudaEventRecord(ctx->ev_done[c], ctx->s_main);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[c], 0);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(labels_out + a * out_stride2, out_stride2, ctx->d_out + a * line_out, line_out,
                                  n1, b - a, cudaMemcpyDeviceToHost, ctx->s_d2h);
        if (e != cudaSuccess) rc = fail_cuda(ctx, e, "D2H copy");
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->s_h2d);
    cudaError_t e2 = cudaStreamSynchronize(ctx->s_main);
    cudaError_t e3 = cudaStreamSynchronize(ctx->s_d2h);
    ctx->count_on = saved_count;
    if (rc) return rc;
    if (e1 != cudaSuccess) return fail_cuda(ctx, e1, "H2D stream");
    if (e2 != cudaSuccess) return fail_cuda(ctx, e2, "kernel stream");
    if (e3 != cudaSuccess) return fail_cuda(ctx, e3, "D2H stream");
    if (counters) return mr_read_counters(ctx, counters, 1);
    return MR_OK;
}

extern "C" int mr_categorize_clean_host(mr_ctx* ctx, const uint8_t* px, int64_t n1, int64_t n2,
                                        int64_t stride2, int bpp, int R, uint8_t* labels_out,
                                        int64_t out_stride2, mr_counters* counters) {
    return host_impl(ctx, "mr_categorize_clean_host", px, n1, n2, stride2, bpp, R, 0, 0, labels_out,
                     out_stride2, counters);
}

extern "C" int mr_categorize_clean_host_band(mr_ctx* ctx, const uint8_t* px, int64_t n1, int64_t n2,
                                             int64_t stride2, int bpp, int R, int halo_lo, int halo_hi,
                                             uint8_t* labels_out, int64_t out_stride2, mr_counters* counters) {
    return host_impl(ctx, "mr_categorize_clean_host_band", px, n1, n2, stride2, bpp, R, halo_lo, halo_hi,
                     labels_out, out_stride2, counters);
}

extern "C" int mr_clean_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                             int R, int repeat, uint8_t* out, int64_t out_stride2) {
    int rc = check_image_args(ctx, "mr_clean_host", labels, out, n1, n2, stride2, 1, out_stride2, R, 0, 0);
    if (rc) return rc;
    if (repeat < 0) return fail(ctx, MR_ERR_ARG, "mr_clean_host: repeat must be >= 0");
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t line = (n1 + 15) / 16 * 16;   // 16-byte aligned device lines: any n1 takes the fast kernel
    const size_t bytes = (size_t)line * n2;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, bytes))) return rc;
    if ((rc = grow(ctx, &ctx->d_tmp, &ctx->d_tmp_cap, bytes))) return rc;
    uint8_t* a = ctx->d_out;
    uint8_t* b = ctx->d_tmp;
    CK(cudaMemcpy2DAsync(a, line, labels, stride2, n1, n2, cudaMemcpyHostToDevice, ctx->s_main));
    const int passes = (R == 0) ? 0 : repeat;
    int max_label = -1;   // the filter never creates a label: one reduction serves every pass
    for (int p = 0; p < passes; ++p) {
        MrJob job = make_job(ctx, a, b, n1, n2, line, 1, R, 0, 0, line);
        job.lut = nullptr;
        job.padded = 1;
        if ((rc = run_clean(ctx, job, ctx->s_main, max_label, &max_label))) return rc;
        uint8_t* t = a; a = b; b = t;
    }
    CK(cudaMemcpy2DAsync(out, out_stride2, a, line, n1, n2, cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    return MR_OK;
}

// ------------------------------------------------------------------ bench/test utility
// ------------------------------------------------------------------ bands with remote halo lines
extern "C" int mr_categorize_clean_dev_peer(mr_ctx* ctx, const uint8_t* px_dev, int64_t n1, int64_t n2,
                                            int64_t stride2_bytes, int bytes_per_px, int R,
                                            const uint8_t* halo_lo_px, int halo_lo, const uint8_t* halo_hi_px,
                                            int halo_hi, uint8_t* labels_dev, int64_t out_stride2, void* stream) {
    int rc = check_image_args(ctx, "mr_categorize_clean_dev_peer", px_dev, labels_dev, n1, n2, stride2_bytes,
                              bytes_per_px, out_stride2, R, halo_lo, halo_hi);
    if (rc) return rc;
    if ((rc = check_palette(ctx, "mr_categorize_clean_dev_peer", bytes_per_px))) return rc;
    if ((halo_lo > 0 && !halo_lo_px) || (halo_hi > 0 && !halo_hi_px))
        return fail(ctx, MR_ERR_ARG, "mr_categorize_clean_dev_peer: null halo pointer with a non-zero halo");
    CK(cudaSetDevice(ctx->device));
    MrJob job = make_job(ctx, px_dev, labels_dev, n1, n2, stride2_bytes, bytes_per_px, R, halo_lo, halo_hi, out_stride2);
    job.halo_lo_px = halo_lo > 0 ? halo_lo_px : nullptr;
    job.halo_hi_px = halo_hi > 0 ? halo_hi_px : nullptr;
    if (n1 == 0 || n2 == 0) return MR_OK;
    if (!mr_fast_supported(job))
        return fail(ctx, MR_ERR_ARG,
                    "mr_categorize_clean_dev_peer: remote halo lines need the fast path (packed RGB, 16-byte aligned "
                    "lines and pointers, n1 % 16 == 0, R <= 3, K <= 126)");
    if (ctx->n_known > 0) return run_with_known(ctx, job, (cudaStream_t)stream);
    int nl = 0;
    cudaError_t e = mr_launch_fast(job, ctx->sm_count, (cudaStream_t)stream, &nl, &ctx->work);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "fused kernel (remote halos)");
    return MR_OK;
}

extern "C" int mr_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset) {
    if (!dev_ptr || !handle64 || !offset) return fail(nullptr, MR_ERR_ARG, "mr_ipc_export: null argument");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));   // handle of the whole allocation
    if (e != cudaSuccess) return fail_cuda(nullptr, e, "cudaIpcGetMemHandle");
    // base of the allocation: driver entry point fetched through the runtime (no link-time libcuda)
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn) return fail_cuda(nullptr, e, "cudaGetDriverEntryPoint(cuMemGetAddressRange)");
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0)
        return fail(nullptr, MR_ERR_CUDA, "mr_ipc_export: cuMemGetAddressRange failed");
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    *offset = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return MR_OK;
}

extern "C" int mr_ipc_open(mr_ctx* ctx, const void* handle64, void** base_out) {
    if (!ctx || !handle64 || !base_out) return fail(ctx, MR_ERR_ARG, "mr_ipc_open: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *base_out = p;
    return MR_OK;
}

extern "C" int mr_ipc_close(mr_ctx* ctx, void* base) {
    if (!ctx) return fail(nullptr, MR_ERR_ARG, "mr_ipc_close: null ctx");
    if (!base) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(base));
    return MR_OK;
}

// ------------------------------------------------------------------ segment prune (_clean)
static int prune_dev(mr_ctx* ctx, const char* fn, const uint8_t* labels_dev, int64_t n1, int64_t n2,
                     int64_t stride2, const uint8_t* keep256, double line_threshold, double prune_threshold,
                     uint8_t* out_dev, int64_t out_stride2, int64_t* n_segments, cudaStream_t s) {
    std::string f(fn);
    if (!keep256) return fail(ctx, MR_ERR_ARG, f + ": null keep table");
    if (isnan(line_threshold) || isnan(prune_threshold)) return fail(ctx, MR_ERR_ARG, f + ": NaN threshold");
    if (n_segments) *n_segments = 0;
    if (n1 == 0 || n2 == 0) return MR_OK;
    if ((double)n1 * (double)n2 >= 2147483648.0) return fail(ctx, MR_ERR_ARG, f + ": n1 * n2 must be below 2^31");
    if (labels_dev == out_dev) return fail(ctx, MR_ERR_ARG, f + ": in-place operation is not supported");
    int rc;
    const size_t cb_bytes = mr_prune_cbase_bytes(n1, n2);
    if ((rc = grow(ctx, &ctx->d_parent, &ctx->d_parent_cap, cb_bytes))) return rc;   // chunk bases
    uint32_t* cbase = reinterpret_cast<uint32_t*>(ctx->d_parent);
    if ((rc = grow(ctx, &ctx->d_stage, &ctx->d_stage_cap, mr_prune_stage_bytes(n1, n2)))) return rc;
    if (!ctx->d_small) CK(cudaMalloc(&ctx->d_small, 2048));
    unsigned int* d_n_roots = reinterpret_cast<unsigned int*>(ctx->d_small + 256);
    unsigned int* d_conflict = reinterpret_cast<unsigned int*>(ctx->d_small + 260);
    CK(cudaMemcpyAsync(ctx->d_small, keep256, 256, cudaMemcpyHostToDevice, s));
    int nl = 0;
    cudaError_t e = mr_launch_prune_count(labels_dev, n1, n2, stride2, cbase, ctx->d_stage, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "segment prune (run count)");
    unsigned int n_runs = 0;
    CK(cudaMemcpyAsync(&n_runs, reinterpret_cast<const uint8_t*>(cbase) + cb_bytes - sizeof(uint32_t), sizeof n_runs,
                       cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if ((rc = grow(ctx, &ctx->d_runs, &ctx->d_runs_cap, mr_prune_run_bytes(n_runs)))) return rc;
    nl = 0;
    e = mr_launch_prune_components(n1, n2, cbase, ctx->d_stage, ctx->d_runs, n_runs, d_n_roots, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "segment prune (components)");
    unsigned int n_roots = 0;
    CK(cudaMemcpyAsync(&n_roots, d_n_roots, sizeof n_roots, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if ((rc = grow(ctx, &ctx->d_rec, &ctx->d_rec_cap, mr_prune_rec_bytes(n_roots)))) return rc;
    nl = 0;
    e = mr_launch_prune_apply(n1, n2, cbase, ctx->d_stage, ctx->d_runs, n_runs, ctx->d_rec, n_roots, ctx->n_known,
                              ctx->d_known_idx, ctx->d_known_cat, ctx->d_small, line_threshold, prune_threshold, out_dev,
                              out_stride2, d_conflict, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "segment prune (apply)");
    unsigned int conflict = 0;
    CK(cudaMemcpyAsync(&conflict, d_conflict, sizeof conflict, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n_segments) *n_segments = n_roots;
    if (conflict)
        return fail(ctx, MR_ERR_UNSUPPORTED,
                    f + ": a segment holds two different known categories; the reference splits it in scan order "
                        "(src/scan.jl:140-146), which this implementation does not reproduce");
    return MR_OK;
}

extern "C" int mr_clean_segments_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                                     const uint8_t* keep256, double line_threshold, double prune_threshold,
                                     uint8_t* out_dev, int64_t out_stride2, int64_t* n_segments, void* stream) {
    int rc = check_image_args(ctx, "mr_clean_segments_dev", labels_dev, out_dev, n1, n2, stride2, 1, out_stride2, 0, 0, 0);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return prune_dev(ctx, "mr_clean_segments_dev", labels_dev, n1, n2, stride2, keep256, line_threshold,
                     prune_threshold, out_dev, out_stride2, n_segments, (cudaStream_t)stream);
}

extern "C" int mr_clean_segments_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                                      const uint8_t* keep256, double line_threshold, double prune_threshold,
                                      uint8_t* out, int64_t out_stride2, int64_t* n_segments) {
    int rc = check_image_args(ctx, "mr_clean_segments_host", labels, out, n1, n2, stride2, 1, out_stride2, 0, 0, 0);
    if (rc) return rc;
    if (n_segments) *n_segments = 0;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n1 * n2;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, bytes))) return rc;
    if ((rc = grow(ctx, &ctx->d_tmp, &ctx->d_tmp_cap, bytes))) return rc;
    CK(cudaMemcpy2DAsync(ctx->d_out, n1, labels, stride2, n1, n2, cudaMemcpyHostToDevice, ctx->s_main));
    rc = prune_dev(ctx, "mr_clean_segments_host", ctx->d_out, n1, n2, n1, keep256, line_threshold, prune_threshold,
                   ctx->d_tmp, n1, n_segments, ctx->s_main);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(out, out_stride2, ctx->d_tmp, n1, n1, n2, cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    return MR_OK;
}

// ------------------------------------------------------------------ _balance, per-pixel part (src/balance.jl:25-42)
static int balance_check(mr_ctx* ctx, const char* fn, const void* px, const void* out, int64_t n1, int64_t n2,
                         int64_t stride2, int bpp, const double* coef, int64_t out_stride2) {
    std::string f(fn);
    if (!ctx) return fail(nullptr, MR_ERR_ARG, f + ": null ctx");
    if (n1 < 0 || n2 < 0) return fail(ctx, MR_ERR_ARG, f + ": negative size");
    if (bpp != 3) return fail(ctx, MR_ERR_ARG, f + ": pixels must be packed RGB{N0f8} (3 bytes)");
    if (n1 > 0 && n2 > 0 && (!px || !out)) return fail(ctx, MR_ERR_ARG, f + ": null buffer");
    if (!coef) return fail(ctx, MR_ERR_ARG, f + ": null coefficients");
    for (int k = 0; k < 21; ++k)
        if (!isfinite(coef[k])) return fail(ctx, MR_ERR_ARG, f + ": non-finite coefficient");
    if (stride2 < n1 * 3 || out_stride2 < n1 * 24 || out_stride2 % 8 != 0) return fail(ctx, MR_ERR_ARG, f + ": bad stride");
    return MR_OK;
}

extern "C" int mr_balance_dev(mr_ctx* ctx, const uint8_t* px_dev, int64_t n1, int64_t n2, int64_t stride2, int bytes_per_px,
                              const double* coef21, double* out_dev, int64_t out_stride2, void* stream) {
    int rc = balance_check(ctx, "mr_balance_dev", px_dev, out_dev, n1, n2, stride2, bytes_per_px, coef21, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    if ((rc = grow(ctx, &ctx->d_f64, &ctx->d_f64_cap, mr_balance_scratch_bytes(n1, n2)))) return rc;
    int nl = 0;
    cudaError_t e = mr_launch_balance(coef21, px_dev, stride2, n1, n2, reinterpret_cast<double*>(ctx->d_f64), out_dev,
                                      out_stride2, ctx->sm_count, (cudaStream_t)stream, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "balance launch");
    return MR_OK;
}

extern "C" int mr_balance_host(mr_ctx* ctx, const uint8_t* px, int64_t n1, int64_t n2, int64_t stride2, int bytes_per_px,
                               const double* coef21, double* out, int64_t out_stride2) {
    int rc = balance_check(ctx, "mr_balance_host", px, out, n1, n2, stride2, bytes_per_px, coef21, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t in_pitch = (n1 * 3 + 15) / 16 * 16, out_pitch = n1 * 24;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)in_pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_f64b, &ctx->d_f64b_cap, (size_t)out_pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_f64, &ctx->d_f64_cap, mr_balance_scratch_bytes(n1, n2)))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_in, in_pitch, px, stride2, n1 * 3, n2, cudaMemcpyHostToDevice, s));
    int nl = 0;
    cudaError_t e = mr_launch_balance(coef21, ctx->d_in, in_pitch, n1, n2, reinterpret_cast<double*>(ctx->d_f64),
                                      reinterpret_cast<double*>(ctx->d_f64b), out_pitch, ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "balance launch");
    CK(cudaMemcpy2DAsync(out, out_stride2, ctx->d_f64b, out_pitch, n1 * 24, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// ------------------------------------------------------------------ blur + categorise on RGB{Float64}
static int blur_check(mr_ctx* ctx, const char* fn, const void* in, const void* out, int64_t n1, int64_t n2,
                      int64_t in_stride2, int in_bytes_per_px, int R, int repeat, int64_t out_stride2) {
    std::string f(fn);
    if (!ctx) return fail(nullptr, MR_ERR_ARG, f + ": null ctx");
    if (n1 < 0 || n2 < 0) return fail(ctx, MR_ERR_ARG, f + ": negative size");
    if (in_bytes_per_px != 3 && in_bytes_per_px != 24)
        return fail(ctx, MR_ERR_ARG, f + ": pixels must be packed RGB{N0f8} (3 bytes) or RGB{Float64} (24 bytes)");
    if (n1 > 0 && n2 > 0 && (!in || !out)) return fail(ctx, MR_ERR_ARG, f + ": null buffer");
    if (in_stride2 < n1 * in_bytes_per_px || out_stride2 < n1 * 24) return fail(ctx, MR_ERR_ARG, f + ": stride smaller than a line");
    if (in_bytes_per_px == 24 && in_stride2 % 8 != 0) return fail(ctx, MR_ERR_ARG, f + ": Float64 lines must be 8-byte aligned");
    if (out_stride2 % 8 != 0) return fail(ctx, MR_ERR_ARG, f + ": Float64 lines must be 8-byte aligned");
    if (R < 0 || R > 7) return fail(ctx, MR_ERR_ARG, f + ": window radius R must be in 0..7");
    if (repeat < 0) return fail(ctx, MR_ERR_ARG, f + ": negative repeat");
    if (in == out && n1 > 0 && n2 > 0) return fail(ctx, MR_ERR_ARG, f + ": in-place operation is not supported");
    return MR_OK;
}

// `repeat` passes; the first reads the caller's pixels, the last writes `out`, the ones in between ping-pong
// between `out` and a scratch image.  repeat == 0: the conversion A .* 1.0 (a pass with R = 0).
static int blur_dev(mr_ctx* ctx, const void* in, int in_bpp, int64_t in_stride2, int64_t n1, int64_t n2, int R, int repeat,
                    double* out, int64_t out_stride2, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    int passes = repeat, radius = R;
    if (repeat == 0) { passes = 1; radius = 0; }
    const int64_t pitch = n1 * 24;
    int rc;
    if (passes > 1 && (rc = grow(ctx, &ctx->d_f64, &ctx->d_f64_cap, (size_t)pitch * n2))) return rc;
    const void* src = in;
    int64_t src_stride = in_stride2;
    bool src_u8 = in_bpp == 3;
    for (int it = 0; it < passes; ++it) {
        const bool to_out = ((passes - 1 - it) % 2) == 0;
        double* dst = to_out ? out : reinterpret_cast<double*>(ctx->d_f64);
        const int64_t dst_stride = to_out ? out_stride2 : pitch;
        int nl = 0;
        cudaError_t e = mr_launch_blur(src, src_u8, src_stride, n1, n2, radius, dst, dst_stride, ctx->sm_count, s, &nl);
        ctx->launches += nl;
        if (e != cudaSuccess) return fail_cuda(ctx, e, "blur launch");
        src = dst; src_stride = dst_stride; src_u8 = false;
    }
    return MR_OK;
}

extern "C" int mr_blur_dev(mr_ctx* ctx, const void* px_dev, int64_t n1, int64_t n2, int64_t stride2, int bytes_per_px,
                           int R, int repeat, double* out_dev, int64_t out_stride2, void* stream) {
    int rc = blur_check(ctx, "mr_blur_dev", px_dev, out_dev, n1, n2, stride2, bytes_per_px, R, repeat, out_stride2);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return blur_dev(ctx, px_dev, bytes_per_px, stride2, n1, n2, R, repeat, out_dev, out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_blur_host(mr_ctx* ctx, const void* px, int64_t n1, int64_t n2, int64_t stride2, int bytes_per_px,
                            int R, int repeat, double* out, int64_t out_stride2) {
    int rc = blur_check(ctx, "mr_blur_host", px, out, n1, n2, stride2, bytes_per_px, R, repeat, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t in_pitch = (n1 * bytes_per_px + 15) / 16 * 16, out_pitch = n1 * 24;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)in_pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_f64b, &ctx->d_f64b_cap, (size_t)out_pitch * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_in, in_pitch, px, stride2, n1 * bytes_per_px, n2, cudaMemcpyHostToDevice, s));
    rc = blur_dev(ctx, ctx->d_in, bytes_per_px, in_pitch, n1, n2, R, repeat, reinterpret_cast<double*>(ctx->d_f64b), out_pitch, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(out, out_stride2, ctx->d_f64b, out_pitch, n1 * 24, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// categorise (+ known points) on Float64 colours, then the standalone majority filter
static int cat_f64_dev(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2, int R, uint8_t* out,
                       int64_t out_stride2, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    uint8_t* plane = out;
    int64_t plane_stride = out_stride2;
    const int64_t line_tmp = (n1 + 15) / 16 * 16;
    int rc;
    if (R > 0) {
        if ((rc = grow(ctx, &ctx->d_tmp, &ctx->d_tmp_cap, (size_t)line_tmp * n2))) return rc;
        plane = ctx->d_tmp;
        plane_stride = line_tmp;
    }
    int nl = 0;
    // colour-only matching = the PixelProp categoriser with match = color (mr_texture.cu)
    cudaError_t e = mr_launch_categorize_props(ctx->pal, nullptr, 1, px, stride2, nullptr, 0, nullptr, 0, n1, n2, ctx->n_known,
                                               ctx->d_known_idx, ctx->d_known_cat, plane, plane_stride,
                                               (ctx->count_on && R == 0) ? ctx->d_counters : nullptr, ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "Float64 categorise launch");
    if (R > 0) {
        MrJob c = make_job(ctx, plane, out, n1, n2, plane_stride, 1, R, 0, 0, out_stride2);
        c.lut = nullptr;
        c.padded = 1;
        return run_clean(ctx, c, s, ctx->pal.K);
    }
    return MR_OK;
}

static int cat_f64_check(mr_ctx* ctx, const char* fn, const void* px, const void* out, int64_t n1, int64_t n2,
                         int64_t stride2, int R, int64_t out_stride2) {
    std::string f(fn);
    int rc = check_image_args(ctx, fn, px, out, n1, n2, stride2, 24, out_stride2, R, 0, 0);
    if (rc) return rc;
    if (!ctx->palette_set) return fail(ctx, MR_ERR_STATE, f + ": mr_set_palette has not been called");
    if (ctx->pal.C != 3 || ctx->pal.colourspace != MR_COLOURSPACE_RGB)
        return fail(ctx, MR_ERR_ARG, f + ": RGB{Float64} pixels take 3-channel RGB statistics");
    if (stride2 % 8 != 0) return fail(ctx, MR_ERR_ARG, f + ": Float64 lines must be 8-byte aligned");
    return MR_OK;
}

extern "C" int mr_categorize_clean_f64_dev(mr_ctx* ctx, const double* px_dev, int64_t n1, int64_t n2, int64_t stride2,
                                           int R, uint8_t* labels_dev, int64_t out_stride2, void* stream) {
    int rc = cat_f64_check(ctx, "mr_categorize_clean_f64_dev", px_dev, labels_dev, n1, n2, stride2, R, out_stride2);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return cat_f64_dev(ctx, px_dev, n1, n2, stride2, R, labels_dev, out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_categorize_clean_f64_host(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2,
                                            int R, uint8_t* labels_out, int64_t out_stride2) {
    int rc = cat_f64_check(ctx, "mr_categorize_clean_f64_host", px, labels_out, n1, n2, stride2, R, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = n1 * 24, line_out = (n1 + 15) / 16 * 16;
    if ((rc = grow(ctx, &ctx->d_f64b, &ctx->d_f64b_cap, (size_t)pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, (size_t)line_out * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_f64b, pitch, px, stride2, n1 * 24, n2, cudaMemcpyHostToDevice, s));
    rc = cat_f64_dev(ctx, reinterpret_cast<const double*>(ctx->d_f64b), n1, n2, pitch, R, ctx->d_out, line_out, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(labels_out, out_stride2, ctx->d_out, line_out, n1, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// ------------------------------------------------------------------ texture features + PixelProp categoriser
// (_stds src/selectcolors.jl:706-715, _stripes src/stripes.jl:1-35, _categorize with a match tuple
//  src/categories.jl:76-133)
extern "C" int mr_set_texture_stats(mr_ctx* ctx, int K, const double* std_mean, const double* std_lo, const double* std_hi,
                                    const double* stripe_mean, const double* stripe_lo, const double* stripe_hi) {
    if (!ctx) return fail(nullptr, MR_ERR_ARG, "mr_set_texture_stats: null ctx");
    if (!ctx->palette_set) return fail(ctx, MR_ERR_STATE, "mr_set_texture_stats: mr_set_palette has not been called");
    if (K != ctx->pal.K) return fail(ctx, MR_ERR_ARG, "mr_set_texture_stats: K differs from the palette's");
    if (!std_mean || !std_lo || !std_hi || !stripe_mean || !stripe_lo || !stripe_hi)
        return fail(ctx, MR_ERR_ARG, "mr_set_texture_stats: null statistics");
    CK(cudaSetDevice(ctx->device));
    std::vector<double> h((size_t)9 * K);
    memcpy(h.data(), std_mean, (size_t)K * 8);
    memcpy(h.data() + K, std_lo, (size_t)K * 8);
    memcpy(h.data() + 2 * K, std_hi, (size_t)K * 8);
    memcpy(h.data() + 3 * K, stripe_mean, (size_t)2 * K * 8);
    memcpy(h.data() + 5 * K, stripe_lo, (size_t)2 * K * 8);
    memcpy(h.data() + 7 * K, stripe_hi, (size_t)2 * K * 8);
    CK(cudaStreamSynchronize(ctx->s_main));
    if (!ctx->d_tex) CK(cudaMalloc(&ctx->d_tex, (size_t)254 * 9 * 8));
    CK(cudaMemcpy(ctx->d_tex, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    ctx->tex_set = true;
    return MR_OK;
}

// comp = doubles per pixel of the output plane (1: stds, 2: stripes)
static int texture_check(mr_ctx* ctx, const char* fn, const void* px, const void* out, int64_t n1, int64_t n2,
                         int64_t stride2, int comp, int64_t out_stride2) {
    std::string f(fn);
    if (!ctx) return fail(nullptr, MR_ERR_ARG, f + ": null ctx");
    if (n1 < 0 || n2 < 0) return fail(ctx, MR_ERR_ARG, f + ": negative size");
    if (n1 > 0 && n2 > 0 && (!px || !out)) return fail(ctx, MR_ERR_ARG, f + ": null buffer");
    if (stride2 < n1 * 24 || out_stride2 < n1 * 8 * comp) return fail(ctx, MR_ERR_ARG, f + ": stride smaller than a line");
    if (stride2 % 8 != 0 || out_stride2 % 8 != 0) return fail(ctx, MR_ERR_ARG, f + ": Float64 lines must be 8-byte aligned");
    if (px == out && n1 > 0 && n2 > 0) return fail(ctx, MR_ERR_ARG, f + ": in-place operation is not supported");
    return MR_OK;
}

static int stds_dev(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2, double* out,
                    int64_t out_stride2, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    int rc;
    if ((rc = grow(ctx, &ctx->d_f64, &ctx->d_f64_cap, 64))) return rc;
    int nl = 0;
    cudaError_t e = mr_launch_stds(px, stride2, n1, n2, out, out_stride2, reinterpret_cast<double*>(ctx->d_f64),
                                   ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "stds launch");
    return MR_OK;
}

static int stripes_dev(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2, int radius, double* out,
                       int64_t out_stride2, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    int rc;
    if ((rc = grow(ctx, &ctx->d_f64, &ctx->d_f64_cap, mr_stripes_scratch_bytes(n1, n2)))) return rc;
    int nl = 0;
    cudaError_t e = mr_launch_stripes(px, stride2, n1, n2, radius, out, out_stride2, reinterpret_cast<double*>(ctx->d_f64),
                                      ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "stripes launch");
    return MR_OK;
}

extern "C" int mr_stds_dev(mr_ctx* ctx, const double* px_dev, int64_t n1, int64_t n2, int64_t stride2, double* out_dev,
                           int64_t out_stride2, void* stream) {
    int rc = texture_check(ctx, "mr_stds_dev", px_dev, out_dev, n1, n2, stride2, 1, out_stride2);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return stds_dev(ctx, px_dev, n1, n2, stride2, out_dev, out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_stripes_dev(mr_ctx* ctx, const double* px_dev, int64_t n1, int64_t n2, int64_t stride2, int radius,
                              double* out_dev, int64_t out_stride2, void* stream) {
    int rc = texture_check(ctx, "mr_stripes_dev", px_dev, out_dev, n1, n2, stride2, 2, out_stride2);
    if (rc) return rc;
    if (radius < 1 || radius > 20) return fail(ctx, MR_ERR_ARG, "mr_stripes_dev: radius must be in 1..20 (the slider's range)");
    CK(cudaSetDevice(ctx->device));
    return stripes_dev(ctx, px_dev, n1, n2, stride2, radius, out_dev, out_stride2, (cudaStream_t)stream);
}

// host variants: pixels -> d_f64b, plane -> d_in (staging), scratch in d_f64
static int texture_host(mr_ctx* ctx, const char* fn, const double* px, int64_t n1, int64_t n2, int64_t stride2, int comp,
                        int radius, double* out, int64_t out_stride2) {
    int rc = texture_check(ctx, fn, px, out, n1, n2, stride2, comp, out_stride2);
    if (rc) return rc;
    if (comp == 2 && (radius < 1 || radius > 20)) return fail(ctx, MR_ERR_ARG, std::string(fn) + ": radius must be in 1..20 (the slider's range)");
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = n1 * 24, opitch = n1 * 8 * comp;
    if ((rc = grow(ctx, &ctx->d_f64b, &ctx->d_f64b_cap, (size_t)pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)opitch * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_f64b, pitch, px, stride2, n1 * 24, n2, cudaMemcpyHostToDevice, s));
    double* plane = reinterpret_cast<double*>(ctx->d_in);
    const double* dpx = reinterpret_cast<const double*>(ctx->d_f64b);
    rc = comp == 1 ? stds_dev(ctx, dpx, n1, n2, pitch, plane, opitch, s)
                   : stripes_dev(ctx, dpx, n1, n2, pitch, radius, plane, opitch, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(out, out_stride2, plane, opitch, opitch, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

extern "C" int mr_stds_host(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2, double* out,
                            int64_t out_stride2) {
    return texture_host(ctx, "mr_stds_host", px, n1, n2, stride2, 1, 0, out, out_stride2);
}

extern "C" int mr_stripes_host(mr_ctx* ctx, const double* px, int64_t n1, int64_t n2, int64_t stride2, int radius,
                               double* out, int64_t out_stride2) {
    return texture_host(ctx, "mr_stripes_host", px, n1, n2, stride2, 2, radius, out, out_stride2);
}

static int props_check(mr_ctx* ctx, const char* fn, const void* px, int64_t stride2, const void* sd, int64_t sd_stride2,
                       const void* st, int64_t st_stride2, int64_t n1, int64_t n2, int match, int R, const void* out,
                       int64_t out_stride2) {
    std::string f(fn);
    int rc = check_image_args(ctx, fn, px, out, n1, n2, stride2, 24, out_stride2, R, 0, 0);
    if (rc) return rc;
    if (!ctx->palette_set) return fail(ctx, MR_ERR_STATE, f + ": mr_set_palette has not been called");
    if (ctx->pal.C != 3 || ctx->pal.colourspace != MR_COLOURSPACE_RGB)
        return fail(ctx, MR_ERR_ARG, f + ": RGB{Float64} pixels take 3-channel RGB statistics");
    if (match < 1 || match > 7) return fail(ctx, MR_ERR_ARG, f + ": match must be a non-empty subset of color (1) | std (2) | stripe (4)");
    if ((match & 6) && !ctx->tex_set) return fail(ctx, MR_ERR_STATE, f + ": mr_set_texture_stats has not been called for this palette");
    if (n1 > 0 && n2 > 0 && (((match & 2) && !sd) || ((match & 4) && !st))) return fail(ctx, MR_ERR_ARG, f + ": a matched texture plane is null");
    if ((match & 2) && sd_stride2 < n1 * 8) return fail(ctx, MR_ERR_ARG, f + ": std stride smaller than a line");
    if ((match & 4) && st_stride2 < n1 * 16) return fail(ctx, MR_ERR_ARG, f + ": stripe stride smaller than a line");
    if (stride2 % 8 != 0 || ((match & 2) && sd_stride2 % 8 != 0) || ((match & 4) && st_stride2 % 8 != 0))
        return fail(ctx, MR_ERR_ARG, f + ": Float64 lines must be 8-byte aligned");
    return MR_OK;
}

// categorise PixelProp pixels (+ known points), then the standalone majority filter
static int props_dev(mr_ctx* ctx, const double* px, int64_t stride2, const double* sd, int64_t sd_stride2, const double* st,
                     int64_t st_stride2, int64_t n1, int64_t n2, int match, int R, uint8_t* out, int64_t out_stride2,
                     cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    uint8_t* plane = out;
    int64_t plane_stride = out_stride2;
    const int64_t line_tmp = (n1 + 15) / 16 * 16;
    int rc;
    if (R > 0) {
        if ((rc = grow(ctx, &ctx->d_tmp, &ctx->d_tmp_cap, (size_t)line_tmp * n2))) return rc;
        plane = ctx->d_tmp;
        plane_stride = line_tmp;
    }
    int nl = 0;
    cudaError_t e = mr_launch_categorize_props(ctx->pal, ctx->d_tex, match, px, stride2, (match & 2) ? sd : nullptr, sd_stride2,
                                               (match & 4) ? st : nullptr, st_stride2, n1, n2, ctx->n_known, ctx->d_known_idx,
                                               ctx->d_known_cat, plane, plane_stride,
                                               (ctx->count_on && R == 0) ? ctx->d_counters : nullptr, ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "PixelProp categorise launch");
    if (R > 0) {
        MrJob c = make_job(ctx, plane, out, n1, n2, plane_stride, 1, R, 0, 0, out_stride2);
        c.lut = nullptr;
        c.padded = 1;
        return run_clean(ctx, c, s, ctx->pal.K);
    }
    return MR_OK;
}

extern "C" int mr_categorize_clean_props_dev(mr_ctx* ctx, const double* px_dev, int64_t stride2, const double* std_dev,
                                             int64_t std_stride2, const double* stripe_dev, int64_t stripe_stride2,
                                             int64_t n1, int64_t n2, int match, int R, uint8_t* labels_dev,
                                             int64_t out_stride2, void* stream) {
    int rc = props_check(ctx, "mr_categorize_clean_props_dev", px_dev, stride2, std_dev, std_stride2, stripe_dev,
                         stripe_stride2, n1, n2, match, R, labels_dev, out_stride2);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return props_dev(ctx, px_dev, stride2, std_dev, std_stride2, stripe_dev, stripe_stride2, n1, n2, match, R, labels_dev,
                     out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_categorize_clean_props_host(mr_ctx* ctx, const double* px, int64_t stride2, const double* std_plane,
                                              int64_t std_stride2, const double* stripe_plane, int64_t stripe_stride2,
                                              int64_t n1, int64_t n2, int match, int R, uint8_t* labels_out,
                                              int64_t out_stride2) {
    int rc = props_check(ctx, "mr_categorize_clean_props_host", px, stride2, std_plane, std_stride2, stripe_plane,
                         stripe_stride2, n1, n2, match, R, labels_out, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = n1 * 24, line_out = (n1 + 15) / 16 * 16;
    // pixels -> d_f64b; std plane | stripe plane -> d_f64; labels -> d_out
    if ((rc = grow(ctx, &ctx->d_f64b, &ctx->d_f64b_cap, (size_t)pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_f64, &ctx->d_f64_cap, (size_t)pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, (size_t)line_out * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    double* d_sd = reinterpret_cast<double*>(ctx->d_f64);
    double* d_st = d_sd + n1 * n2;
    CK(cudaMemcpy2DAsync(ctx->d_f64b, pitch, px, stride2, n1 * 24, n2, cudaMemcpyHostToDevice, s));
    if (match & 2) CK(cudaMemcpy2DAsync(d_sd, n1 * 8, std_plane, std_stride2, n1 * 8, n2, cudaMemcpyHostToDevice, s));
    if (match & 4) CK(cudaMemcpy2DAsync(d_st, n1 * 16, stripe_plane, stripe_stride2, n1 * 16, n2, cudaMemcpyHostToDevice, s));
    rc = props_dev(ctx, reinterpret_cast<const double*>(ctx->d_f64b), pitch, d_sd, n1 * 8, d_st, n1 * 16, n1, n2, match, R,
                   ctx->d_out, line_out, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(labels_out, out_stride2, ctx->d_out, line_out, n1, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// ------------------------------------------------------------------ _shrink_grow (src/clean.jl:96-140)
// n_erode passes of _erode, then n_dilate passes of _dilate; ping-pong between `out` and a scratch plane so that the
// last pass writes `out`
static int shrink_grow_dev(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int n_erode,
                           int n_dilate, uint8_t* out, int64_t out_stride2, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    const int passes = n_erode + n_dilate;
    if (passes == 0) {
        CK(cudaMemcpy2DAsync(out, out_stride2, labels, stride2, n1, n2, cudaMemcpyDeviceToDevice, s));
        return MR_OK;
    }
    const int64_t pitch = (n1 + 15) / 16 * 16;
    int rc;
    if (passes > 1 && (rc = grow(ctx, &ctx->d_lab, &ctx->d_lab_cap, (size_t)pitch * n2))) return rc;
    const uint8_t* src = labels;
    int64_t src_stride = stride2;
    for (int it = 0; it < passes; ++it) {
        const bool to_out = ((passes - 1 - it) % 2) == 0;
        uint8_t* dst = to_out ? out : ctx->d_lab;
        const int64_t dst_stride = to_out ? out_stride2 : pitch;
        int nl = 0;
        cudaError_t e = mr_launch_morph(it >= n_erode, src, src_stride, n1, n2, dst, dst_stride, ctx->sm_count, s, &nl);
        ctx->launches += nl;
        if (e != cudaSuccess) return fail_cuda(ctx, e, "erode / dilate launch");
        src = dst; src_stride = dst_stride;
    }
    return MR_OK;
}

static int shrink_grow_check(mr_ctx* ctx, const char* fn, const void* labels, const void* out, int64_t n1, int64_t n2,
                             int64_t stride2, int n_erode, int n_dilate, int64_t out_stride2) {
    int rc = check_image_args(ctx, fn, labels, out, n1, n2, stride2, 1, out_stride2, 0, 0, 0);
    if (rc) return rc;
    if (n_erode < 0 || n_dilate < 0) return fail(ctx, MR_ERR_ARG, std::string(fn) + ": negative pass count");
    if (labels == out && n1 > 0 && n2 > 0) return fail(ctx, MR_ERR_ARG, std::string(fn) + ": in-place operation is not supported");
    return MR_OK;
}

extern "C" int mr_shrink_grow_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2, int n_erode,
                                  int n_dilate, uint8_t* out_dev, int64_t out_stride2, void* stream) {
    int rc = shrink_grow_check(ctx, "mr_shrink_grow_dev", labels_dev, out_dev, n1, n2, stride2, n_erode, n_dilate, out_stride2);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return shrink_grow_dev(ctx, labels_dev, n1, n2, stride2, n_erode, n_dilate, out_dev, out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_shrink_grow_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int n_erode,
                                   int n_dilate, uint8_t* out, int64_t out_stride2) {
    int rc = shrink_grow_check(ctx, "mr_shrink_grow_host", labels, out, n1, n2, stride2, n_erode, n_dilate, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = (n1 + 15) / 16 * 16;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, (size_t)pitch * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_in, pitch, labels, stride2, n1, n2, cudaMemcpyHostToDevice, s));
    rc = shrink_grow_dev(ctx, ctx->d_in, n1, n2, pitch, n_erode, n_dilate, ctx->d_out, pitch, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(out, out_stride2, ctx->d_out, pitch, n1, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// ------------------------------------------------------------------ _recolor (src/selectcolors.jl:650-655)
static int recolor_dev(mr_ctx* ctx, const char* fn, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int K,
                       const double* colors, double* out, int64_t out_stride2, cudaStream_t s) {
    std::string f(fn);
    if (K < 1 || K > 254) return fail(ctx, MR_ERR_ARG, f + ": K must be in 1..254");
    if (!colors) return fail(ctx, MR_ERR_ARG, f + ": null colour table");
    if (out_stride2 < n1 * 32 || out_stride2 % 8 != 0) return fail(ctx, MR_ERR_ARG, f + ": output lines hold 32 bytes per pixel, 8-byte aligned");
    if (n1 == 0 || n2 == 0) return MR_OK;
    int rc;
    if ((rc = grow(ctx, &ctx->d_poly, &ctx->d_poly_cap, (size_t)K * 24))) return rc;
    CK(cudaMemcpyAsync(ctx->d_poly, colors, (size_t)K * 24, cudaMemcpyHostToDevice, s));   // pageable: staged before it returns
    int nl = 0;
    cudaError_t e = mr_launch_recolor(labels, stride2, n1, n2, K, reinterpret_cast<const double*>(ctx->d_poly), out, out_stride2,
                                      ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "recolor launch");
    return MR_OK;
}

extern "C" int mr_recolor_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2, int K,
                              const double* colors, double* out_dev, int64_t out_stride2, void* stream) {
    int rc = check_image_args(ctx, "mr_recolor_dev", labels_dev, out_dev, n1, n2, stride2, 1, n1, 0, 0, 0);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return recolor_dev(ctx, "mr_recolor_dev", labels_dev, n1, n2, stride2, K, colors, out_dev, out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_recolor_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, int K,
                               const double* colors, double* out, int64_t out_stride2) {
    int rc = check_image_args(ctx, "mr_recolor_host", labels, out, n1, n2, stride2, 1, n1, 0, 0, 0);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = (n1 + 15) / 16 * 16, opitch = n1 * 32;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)pitch * n2))) return rc;
    if ((rc = grow(ctx, &ctx->d_f64b, &ctx->d_f64b_cap, (size_t)opitch * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_in, pitch, labels, stride2, n1, n2, cudaMemcpyHostToDevice, s));
    rc = recolor_dev(ctx, "mr_recolor_host", ctx->d_in, n1, n2, pitch, K, colors, reinterpret_cast<double*>(ctx->d_f64b), opitch, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(out, out_stride2, ctx->d_f64b, opitch, opitch, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// ------------------------------------------------------------------ polygon masks / polygon painting
// (_polymask src/selectcolors.jl:559-563; (1 .- mask) .* A :480-481; rasterize!(...; fill = i) :494-496)
static int polygons_dev(mr_ctx* ctx, const char* fn, int n_rings, const int32_t* offsets, const double* xy, const uint8_t* fill,
                        int64_t n1, int64_t n2, uint8_t* plane_dev, int64_t stride2, cudaStream_t s) {
    std::string f(fn);
    if (n_rings < 0) return fail(ctx, MR_ERR_ARG, f + ": negative ring count");
    if (n_rings > 0 && (!offsets || !xy)) return fail(ctx, MR_ERR_ARG, f + ": null ring arrays");
    if (n_rings > 0 && offsets[0] != 0) return fail(ctx, MR_ERR_ARG, f + ": offsets[0] must be 0");
    for (int p = 0; p < n_rings; ++p)
        if (offsets[p + 1] < offsets[p]) return fail(ctx, MR_ERR_ARG, f + ": offsets must not decrease");
    const int64_t nv = n_rings > 0 ? offsets[n_rings] : 0;
    for (int64_t k = 0; k < 2 * nv; ++k)
        if (!isfinite(xy[k])) return fail(ctx, MR_ERR_ARG, f + ": vertex coordinates must be finite");
    if (n1 == 0 || n2 == 0) return MR_OK;
    // pack: off | xy | ext | ring_of | fill (8-byte aligned parts)
    const size_t off_b = ((size_t)(n_rings + 1) * 4 + 7) / 8 * 8, xy_b = (size_t)nv * 16, ext_b = (size_t)n_rings * 16;
    const size_t ro_b = ((size_t)nv * 4 + 7) / 8 * 8;
    std::vector<uint8_t> h(off_b + xy_b + ext_b + ro_b + (size_t)n_rings + 8, 0);
    if (n_rings > 0) memcpy(h.data(), offsets, (size_t)(n_rings + 1) * 4);
    if (nv > 0) memcpy(h.data() + off_b, xy, xy_b);
    double* ext = reinterpret_cast<double*>(h.data() + off_b + xy_b);
    int32_t* ring_of = reinterpret_cast<int32_t*>(h.data() + off_b + xy_b + ext_b);
    for (int p = 0; p < n_rings; ++p) {
        double lo = INFINITY, hi = -INFINITY;
        for (int k = offsets[p]; k < offsets[p + 1]; ++k) {
            lo = fmin(lo, xy[2 * k + 1]); hi = fmax(hi, xy[2 * k + 1]);
            ring_of[k] = offsets[p + 1] - offsets[p] >= 3 ? p : -1;
        }
        ext[2 * p] = lo; ext[2 * p + 1] = hi;
    }
    if (fill && n_rings > 0) memcpy(h.data() + off_b + xy_b + ext_b + ro_b, fill, (size_t)n_rings);
    int rc;
    if ((rc = grow(ctx, &ctx->d_poly, &ctx->d_poly_cap, h.size()))) return rc;
    CK(cudaMemcpyAsync(ctx->d_poly, h.data(), h.size(), cudaMemcpyHostToDevice, s));   // pageable: staged before it returns
    int nl = 0;
    cudaError_t e = mr_launch_polygons(n_rings, (int)nv, reinterpret_cast<const int32_t*>(ctx->d_poly),
                                       reinterpret_cast<const double*>(ctx->d_poly + off_b),
                                       reinterpret_cast<const double*>(ctx->d_poly + off_b + xy_b),
                                       reinterpret_cast<const int32_t*>(ctx->d_poly + off_b + xy_b + ext_b),
                                       fill ? ctx->d_poly + off_b + xy_b + ext_b + ro_b : nullptr, n1, n2, plane_dev, stride2,
                                       ctx->sm_count, s, &nl);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "polygon launch");
    return MR_OK;
}

extern "C" int mr_polygons_dev(mr_ctx* ctx, int n_rings, const int32_t* offsets, const double* xy, const uint8_t* fill,
                               int64_t n1, int64_t n2, uint8_t* plane_dev, int64_t stride2, void* stream) {
    int rc = check_image_args(ctx, "mr_polygons_dev", plane_dev, plane_dev, n1, n2, stride2, 1, stride2, 0, 0, 0);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return polygons_dev(ctx, "mr_polygons_dev", n_rings, offsets, xy, fill, n1, n2, plane_dev, stride2, (cudaStream_t)stream);
}

extern "C" int mr_polygons_host(mr_ctx* ctx, int n_rings, const int32_t* offsets, const double* xy, const uint8_t* fill,
                                int64_t n1, int64_t n2, uint8_t* plane, int64_t stride2) {
    int rc = check_image_args(ctx, "mr_polygons_host", plane, plane, n1, n2, stride2, 1, stride2, 0, 0, 0);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = (n1 + 15) / 16 * 16;
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, (size_t)pitch * n2))) return rc;
    cudaStream_t s = ctx->s_main;
    if (fill) CK(cudaMemcpy2DAsync(ctx->d_out, pitch, plane, stride2, n1, n2, cudaMemcpyHostToDevice, s));   // painting keeps the rest
    rc = polygons_dev(ctx, "mr_polygons_host", n_rings, offsets, xy, fill, n1, n2, ctx->d_out, pitch, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(plane, stride2, ctx->d_out, pitch, n1, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

// ------------------------------------------------------------------ fill_gaps (src/clean.jl:2-54)
static int fill_check(mr_ctx* ctx, const char* fn, const void* labels, const void* px, const void* out, int64_t n1,
                      int64_t n2, int64_t stride2, int64_t px_stride2, int bpp, const void* mask, int64_t mask_stride2,
                      const uint8_t* categories, int n_categories, int repeat, int64_t out_stride2) {
    std::string f(fn);
    int rc = check_image_args(ctx, fn, labels, out, n1, n2, stride2, 1, out_stride2, 0, 0, 0);
    if (rc) return rc;
    if ((rc = check_palette(ctx, fn, bpp))) return rc;
    if (n1 > 0 && n2 > 0 && !px) return fail(ctx, MR_ERR_ARG, f + ": null pixel buffer (`original`)");
    if (px_stride2 < n1 * bpp) return fail(ctx, MR_ERR_ARG, f + ": pixel stride smaller than a line");
    if (mask && mask_stride2 < n1) return fail(ctx, MR_ERR_ARG, f + ": mask stride smaller than a line");
    if (n_categories < 0 || n_categories > 254 || (n_categories > 0 && !categories))
        return fail(ctx, MR_ERR_ARG, f + ": bad category list");
    for (int k = 0; k < n_categories; ++k)
        if (categories[k] < 1 || categories[k] > ctx->pal.K)
            return fail(ctx, MR_ERR_ARG, f + ": categories must be in 1..K (stats[i] is the palette row of categories[i])");
    if (repeat < 0) return fail(ctx, MR_ERR_ARG, f + ": negative repeat");
    if (labels == out && n1 > 0 && n2 > 0) return fail(ctx, MR_ERR_ARG, f + ": in-place operation is not supported");
    return MR_OK;
}

// `repeat` passes (selectcolors.jl:692-696); odd passes write `out`, even passes the scratch plane, arranged
// so that the last pass lands in `out`
static int fill_dev(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2, const uint8_t* px,
                    int64_t px_stride2, int bpp, const uint8_t* mask, int64_t mask_stride2, const uint8_t* categories,
                    int n_categories, int repeat, uint8_t* out, int64_t out_stride2, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return MR_OK;
    if (repeat == 0) {
        CK(cudaMemcpy2DAsync(out, out_stride2, labels, stride2, n1, n2, cudaMemcpyDeviceToDevice, s));
        return MR_OK;
    }
    const int64_t pitch = (n1 + 15) / 16 * 16;
    int rc;
    if (repeat > 1 && (rc = grow(ctx, &ctx->d_lab, &ctx->d_lab_cap, (size_t)pitch * n2))) return rc;
    const uint8_t* src = labels;
    int64_t src_stride = stride2;
    for (int it = 0; it < repeat; ++it) {
        const bool to_out = ((repeat - 1 - it) % 2) == 0;
        uint8_t* dst = to_out ? out : ctx->d_lab;
        const int64_t dst_stride = to_out ? out_stride2 : pitch;
        int nl = 0;
        cudaError_t e = mr_launch_fill_gaps(ctx->pal, src, n1, n2, src_stride, px, px_stride2, bpp, mask, mask_stride2,
                                            categories, n_categories, ctx->n_known, ctx->d_known_idx, ctx->d_known_cat,
                                            dst, dst_stride, ctx->sm_count, s, &nl);
        ctx->launches += nl;
        if (e != cudaSuccess) return fail_cuda(ctx, e, "fill_gaps launch");
        src = dst;
        src_stride = dst_stride;
    }
    return MR_OK;
}

extern "C" int mr_fill_gaps_dev(mr_ctx* ctx, const uint8_t* labels_dev, int64_t n1, int64_t n2, int64_t stride2,
                                const uint8_t* px_dev, int64_t px_stride2, int bytes_per_px, const uint8_t* mask_dev,
                                int64_t mask_stride2, const uint8_t* categories, int n_categories, int repeat,
                                uint8_t* out_dev, int64_t out_stride2, void* stream) {
    int rc = fill_check(ctx, "mr_fill_gaps_dev", labels_dev, px_dev, out_dev, n1, n2, stride2, px_stride2, bytes_per_px,
                        mask_dev, mask_stride2, categories, n_categories, repeat, out_stride2);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return fill_dev(ctx, labels_dev, n1, n2, stride2, px_dev, px_stride2, bytes_per_px, mask_dev, mask_stride2, categories,
                    n_categories, repeat, out_dev, out_stride2, (cudaStream_t)stream);
}

extern "C" int mr_fill_gaps_host(mr_ctx* ctx, const uint8_t* labels, int64_t n1, int64_t n2, int64_t stride2,
                                 const uint8_t* px, int64_t px_stride2, int bytes_per_px, const uint8_t* mask,
                                 int64_t mask_stride2, const uint8_t* categories, int n_categories, int repeat,
                                 uint8_t* out, int64_t out_stride2) {
    int rc = fill_check(ctx, "mr_fill_gaps_host", labels, px, out, n1, n2, stride2, px_stride2, bytes_per_px, mask,
                        mask_stride2, categories, n_categories, repeat, out_stride2);
    if (rc) return rc;
    if (n1 == 0 || n2 == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t pitch = (n1 + 15) / 16 * 16, ppitch = (n1 * bytes_per_px + 15) / 16 * 16;
    const size_t plane = (size_t)pitch * n2;
    // d_out: labels in | result ; d_tmp: mask ; d_in: pixels
    if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, 2 * plane))) return rc;
    if ((rc = grow(ctx, &ctx->d_in, &ctx->d_in_cap, (size_t)ppitch * n2))) return rc;
    if (mask && (rc = grow(ctx, &ctx->d_tmp, &ctx->d_tmp_cap, plane))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpy2DAsync(ctx->d_out, pitch, labels, stride2, n1, n2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpy2DAsync(ctx->d_in, ppitch, px, px_stride2, n1 * bytes_per_px, n2, cudaMemcpyHostToDevice, s));
    if (mask) CK(cudaMemcpy2DAsync(ctx->d_tmp, pitch, mask, mask_stride2, n1, n2, cudaMemcpyHostToDevice, s));
    rc = fill_dev(ctx, ctx->d_out, n1, n2, pitch, ctx->d_in, ppitch, bytes_per_px, mask ? ctx->d_tmp : nullptr, pitch,
                  categories, n_categories, repeat, ctx->d_out + plane, pitch, s);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(out, out_stride2, ctx->d_out + plane, pitch, n1, n2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MR_OK;
}

extern "C" int mr_synth_scan_dev(mr_ctx* ctx, uint32_t seed, uint32_t sheet_id, int K,
                                 const uint8_t* palette_rgb_host, int64_t n1, int64_t n2, int64_t line0,
                                 int64_t nlines, int mode, uint8_t* out_dev, int64_t out_stride2, void* stream) {
    if (!ctx) return fail(nullptr, MR_ERR_ARG, "mr_synth_scan_dev: null ctx");
    if (K < 1 || K > 254 || !palette_rgb_host || n1 < 0 || nlines < 0 || out_stride2 < 3 * n1 || line0 < 0 ||
        line0 + nlines > n2)
        return fail(ctx, MR_ERR_ARG, "mr_synth_scan_dev: bad arguments");
    if (n1 == 0 || nlines == 0) return MR_OK;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_small) CK(cudaMalloc(&ctx->d_small, 2048));
    uint8_t* d_pal = ctx->d_small + 1024;   // K * 3 <= 762 bytes; the call synchronises before it returns
    CK(cudaMemcpy(d_pal, palette_rgb_host, (size_t)K * 3, cudaMemcpyHostToDevice));
    int nl = 0;
    cudaError_t e = mr_launch_synth(seed, sheet_id, K, d_pal, n1, line0, nlines, mode, out_dev, out_stride2,
                                    (cudaStream_t)stream, &nl);
    ctx->launches += nl;
    cudaError_t e2 = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "synth launch");
    if (e2 != cudaSuccess) return fail_cuda(ctx, e2, "synth kernel");
    return MR_OK;
}
