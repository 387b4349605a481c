// mr_multi.cu -- several GPUs of one box behind ONE C-ABI object: the library owns a
// context per device and a host thread per device and call, splits the sheet into row bands over n2 (or a
// batch into groups of sheets) and drives the single-GPU entry points of mr_api.cu.  The reference call
// site (src/selectcolors.jl:466-478) stays ONE call.
//   host-resident sheet:   every band reads its R halo lines straight from the caller's sheet -- nothing is
//                          exchanged between GPUs at all;
//   device-resident bands: band i lives on device i; its halo lines are read IN PLACE from the neighbour
//                          bands over NVLink peer access (enabled here, in-process) by the fused kernel's
//                          TMA line copies -- no exchange step, no NCCL.
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/maprast.h"
#include "mr_internal.h"

// One persistent host thread per device slot beyond the first (slot 0 runs on the caller's thread): a call
// publishes a task, bumps the generation and waits until every worker has reported back.  (Creating the
// threads per call cost about 0.15 ms on an 8-GPU box -- a third of a device-resident config 3 pass.)
struct MrWorkers {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> task;
    std::vector<int> rc;
    unsigned long long gen = 0;
    int pending = 0;
    bool quit = false;

    void start(int n) {
        rc.assign(n, 0);
        for (int i = 1; i < n; ++i)
            th.emplace_back([this, i] {
                unsigned long long seen = 0;
                for (;;) {
                    std::function<int(int)> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv_go.wait(lk, [&] { return quit || gen != seen; });
                        if (quit) return;
                        seen = gen;
                        f = task;
                    }
                    const int r = f(i);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        rc[i] = r;
                        if (--pending == 0) cv_done.notify_one();
                    }
                }
            });
    }
    void run(const std::function<int(int)>& f) {
        const int n = (int)rc.size();
        {
            std::lock_guard<std::mutex> lk(mu);
            task = f;
            pending = n - 1;
            ++gen;
        }
        cv_go.notify_all();
        rc[0] = f(0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv_go.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
};

struct mr_multi {
    std::vector<mr_ctx*> ctx;
    std::vector<int> dev;
    MrWorkers workers;
    std::vector<char> peer_prev, peer_next;   // device i can read device i-1 / i+1
    std::vector<int64_t> known_idx;           // whole-sheet known-category points (host copy)
    std::vector<uint8_t> known_cat;
    std::vector<char> ctx_has_known;          // a context holds points from an earlier call
    std::string err;
};

static thread_local std::string g_merr;

static int mfail(mr_multi* m, int code, const std::string& msg) {
    g_merr = msg;
    if (m) m->err = msg;
    return code;
}

// lines [a, b) of band i of n: contiguous, sizes differ by at most 16 lines, multiples of 16 (whole blocks)
static void band_range(int64_t n2, int n, int i, int64_t* a, int64_t* b) {
    const int64_t units = (n2 + 15) / 16;
    int64_t ua = units * i / n, ub = units * (i + 1) / n;
    *a = ua * 16 < n2 ? ua * 16 : n2;
    *b = ub * 16 < n2 ? ub * 16 : n2;
    if (i == n - 1) *b = n2;
}

extern "C" const char* mr_multi_last_error(const mr_multi* m) { return m ? m->err.c_str() : g_merr.c_str(); }

extern "C" void mr_multi_destroy(mr_multi* m) {
    if (!m) return;
    m->workers.stop();
    for (mr_ctx* c : m->ctx) mr_destroy(c);
    delete m;
}

extern "C" mr_multi* mr_multi_create(const int* device_ids, int n_dev, int* err) {
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        (void)cudaGetLastError();
        mfail(nullptr, MR_ERR_CUDA, "mr_multi_create: no usable CUDA device (libmaprast has no CPU fallback)");
        if (err) *err = MR_ERR_CUDA;
        return nullptr;
    }
    if (n_dev <= 0) { n_dev = have; device_ids = nullptr; }
    if (n_dev > 64) {
        mfail(nullptr, MR_ERR_ARG, "mr_multi_create: more than 64 devices");
        if (err) *err = MR_ERR_ARG;
        return nullptr;
    }
    mr_multi* m = new mr_multi();
    for (int i = 0; i < n_dev; ++i) {
        const int d = device_ids ? device_ids[i] : i;
        int e = MR_OK;
        mr_ctx* c = mr_create(d, &e);     // the same device may appear more than once (two contexts on it)
        if (!c) {
            mfail(nullptr, e, std::string("mr_multi_create: ") + mr_last_error(nullptr));
            if (err) *err = e;
            mr_multi_destroy(m);
            return nullptr;
        }
        m->ctx.push_back(c);
        m->dev.push_back(d);
    }
    // peer access between neighbouring bands (device-resident path); failure only disables that path
    m->peer_prev.assign(n_dev, 0);
    m->peer_next.assign(n_dev, 0);
    for (int i = 0; i < n_dev; ++i)
        for (int s = -1; s <= 1; s += 2) {
            const int j = i + s;
            if (j < 0 || j >= n_dev) continue;
            bool ok = m->dev[i] == m->dev[j];
            if (!ok) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]) == cudaSuccess && can &&
                    cudaSetDevice(m->dev[i]) == cudaSuccess) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
                    ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
                }
                (void)cudaGetLastError();
            }
            (s < 0 ? m->peer_prev : m->peer_next)[i] = ok;
        }
    m->ctx_has_known.assign(n_dev, 0);
    m->workers.start(n_dev);
    if (err) *err = MR_OK;
    return m;
}

extern "C" int mr_multi_device_count(const mr_multi* m) { return m ? (int)m->ctx.size() : 0; }

extern "C" mr_ctx* mr_multi_ctx(mr_multi* m, int i) {
    return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr;
}

extern "C" int mr_multi_band_range(const mr_multi* m, int64_t n2, int i, int64_t* line0, int64_t* nlines) {
    if (!m || i < 0 || i >= (int)m->ctx.size() || n2 < 0 || !line0 || !nlines)
        return mfail(const_cast<mr_multi*>(m), MR_ERR_ARG, "mr_multi_band_range: bad arguments");
    int64_t a, b;
    band_range(n2, (int)m->ctx.size(), i, &a, &b);
    *line0 = a;
    *nlines = b - a;
    return MR_OK;
}

// run f(i) on the host thread of every device slot; first failure wins
template <typename F>
static int on_all(mr_multi* m, const char* what, F f) {
    const int n = (int)m->ctx.size();
    m->workers.run(f);
    const std::vector<int>& rc = m->workers.rc;
    for (int i = 0; i < n; ++i)
        if (rc[i] != MR_OK) {
            char head[64];
            snprintf(head, sizeof head, "%s (device slot %d): ", what, i);
            return mfail(m, rc[i], std::string(head) + mr_last_error(m->ctx[i]));
        }
    return MR_OK;
}

extern "C" int mr_multi_set_palette(mr_multi* m, int K, const double* mean, const double* lo, const double* hi,
                                    int colourspace, int channels) {
    if (!m) return mfail(nullptr, MR_ERR_ARG, "mr_multi_set_palette: null handle");
    return on_all(m, "mr_multi_set_palette",
                  [&](int i) { return mr_set_palette(m->ctx[i], K, mean, lo, hi, colourspace, channels); });
}

extern "C" int mr_multi_set_known(mr_multi* m, int64_t n, const int64_t* lin_idx, const uint8_t* cat) {
    if (!m) return mfail(nullptr, MR_ERR_ARG, "mr_multi_set_known: null handle");
    if (n < 0 || (n > 0 && (!lin_idx || !cat))) return mfail(m, MR_ERR_ARG, "mr_multi_set_known: bad arguments");
    for (int64_t k = 0; k < n; ++k)
        if (lin_idx[k] < 0) return mfail(m, MR_ERR_ARG, "mr_multi_set_known: negative linear index");
    m->known_idx.assign(lin_idx, lin_idx + n);
    m->known_cat.assign(cat, cat + n);
    return MR_OK;
}

static void add_counters(mr_counters* sum, const mr_counters& c) {
    sum->n_pixels += c.n_pixels; sum->n_nomatch += c.n_nomatch; sum->n_fine += c.n_fine; sum->n_changed += c.n_changed;
}

// the known-category points of the sheet that band [a - hl, b + hh) can see, as whole-sheet indices
static int band_known(mr_multi* m, int i, int64_t n1, int64_t a, int64_t b, int R) {
    if (m->known_idx.empty() && !m->ctx_has_known[i]) return MR_OK;   // nothing set, nothing to clear (no sync)
    std::vector<int64_t> idx;
    std::vector<uint8_t> cat;
    for (size_t k = 0; k < m->known_idx.size(); ++k) {
        const int64_t j = n1 > 0 ? m->known_idx[k] / n1 : 0;
        if (j >= a - R && j < b + R) { idx.push_back(m->known_idx[k]); cat.push_back(m->known_cat[k]); }
    }
    m->ctx_has_known[i] = !idx.empty();
    return mr_set_known_band(m->ctx[i], (int64_t)idx.size(), idx.data(), cat.data(), a);
}

extern "C" int mr_multi_categorize_clean_host(mr_multi* m, const uint8_t* px, int64_t n1, int64_t n2,
                                              int64_t stride2_bytes, int bytes_per_px, int R, uint8_t* labels_out,
                                              int64_t out_stride2, mr_counters* counters) {
    if (!m) return mfail(nullptr, MR_ERR_ARG, "mr_multi_categorize_clean_host: null handle");
    if (n1 < 0 || n2 < 0 || R < 0 || R > 7) return mfail(m, MR_ERR_ARG, "mr_multi_categorize_clean_host: bad size / radius");
    const int n = (int)m->ctx.size();
    std::vector<mr_counters> cnt(n);
    const int rc = on_all(m, "mr_multi_categorize_clean_host", [&](int i) {
        int64_t a, b;
        band_range(n2, n, i, &a, &b);
        memset(&cnt[i], 0, sizeof(mr_counters));
        if (b <= a && n2 > 0) return MR_OK;             // more devices than blocks of lines
        if (i > 0 && n2 == 0) return MR_OK;
        int r = band_known(m, i, n1, a, b, R);
        if (r) return r;
        const int hl = (int)(a < R ? a : R), hh = (int)(n2 - b < R ? n2 - b : R);
        r = mr_categorize_clean_host_band(m->ctx[i], px ? px + a * stride2_bytes : px, n1, b - a, stride2_bytes, bytes_per_px,
                                          R, hl, hh, labels_out ? labels_out + a * out_stride2 : labels_out, out_stride2,
                                          counters ? &cnt[i] : nullptr);
        return r;
    });
    if (rc) return rc;
    if (counters) {
        memset(counters, 0, sizeof *counters);
        for (int i = 0; i < n; ++i) add_counters(counters, cnt[i]);
    }
    return MR_OK;
}

extern "C" int mr_multi_batch_categorize_clean_host(mr_multi* m, int n_sheets, const uint8_t* px,
                                                    int64_t sheet_stride_bytes, int64_t n1, int64_t n2,
                                                    int64_t stride2_bytes, int bytes_per_px, int R,
                                                    uint8_t* labels_out, int64_t out_sheet_stride, int64_t out_stride2) {
    if (!m) return mfail(nullptr, MR_ERR_ARG, "mr_multi_batch_categorize_clean_host: null handle");
    if (n_sheets < 0) return mfail(m, MR_ERR_ARG, "mr_multi_batch_categorize_clean_host: negative sheet count");
    if (!m->known_idx.empty())
        return mfail(m, MR_ERR_ARG, "mr_multi_batch_categorize_clean_host: known-category points are per sheet; clear them");
    const int n = (int)m->ctx.size();
    return on_all(m, "mr_multi_batch_categorize_clean_host", [&](int i) {
        const int s0 = (int)((int64_t)n_sheets * i / n), s1 = (int)((int64_t)n_sheets * (i + 1) / n);
        int r = m->ctx_has_known[i] ? mr_set_known(m->ctx[i], 0, nullptr, nullptr) : MR_OK;
        m->ctx_has_known[i] = 0;
        for (int s = s0; s < s1 && r == MR_OK; ++s)
            r = mr_categorize_clean_host(m->ctx[i], px + (int64_t)s * sheet_stride_bytes, n1, n2, stride2_bytes, bytes_per_px,
                                         R, labels_out + (int64_t)s * out_sheet_stride, out_stride2, nullptr);
        return r;
    });
}

extern "C" int mr_multi_categorize_clean_dev(mr_multi* m, const uint8_t* const* band_px, int64_t n1, int64_t n2,
                                             int64_t stride2_bytes, int bytes_per_px, int R,
                                             uint8_t* const* band_out, int64_t out_stride2) {
    if (!m) return mfail(nullptr, MR_ERR_ARG, "mr_multi_categorize_clean_dev: null handle");
    if (!band_px || !band_out || n1 < 0 || n2 < 0 || R < 0 || R > 3)
        return mfail(m, MR_ERR_ARG, "mr_multi_categorize_clean_dev: bad arguments (R <= 3 on this path)");
    const int n = (int)m->ctx.size();
    std::vector<int64_t> a(n), b(n);
    for (int i = 0; i < n; ++i) band_range(n2, n, i, &a[i], &b[i]);
    // halo lines come from the nearest non-empty neighbour band; a band shorter than R cannot lend them
    for (int i = 0; i < n; ++i) {
        if (b[i] <= a[i]) continue;
        if (b[i] - a[i] < R && n > 1)
            return mfail(m, MR_ERR_ARG, "mr_multi_categorize_clean_dev: a band has fewer than R lines (use fewer devices)");
        if (i > 0 && a[i] > 0 && !m->peer_prev[i])
            return mfail(m, MR_ERR_UNSUPPORTED, "mr_multi_categorize_clean_dev: no peer access to the previous band's device");
        if (i < n - 1 && b[i] < n2 && !m->peer_next[i])
            return mfail(m, MR_ERR_UNSUPPORTED, "mr_multi_categorize_clean_dev: no peer access to the next band's device");
    }
    const int rc = on_all(m, "mr_multi_categorize_clean_dev", [&](int i) {
        if (b[i] <= a[i]) return MR_OK;
        int r = band_known(m, i, n1, a[i], b[i], R);
        if (r) return r;
        const int hl = (int)(a[i] < R ? a[i] : R), hh = (int)(n2 - b[i] < R ? n2 - b[i] : R);
        int ip = i - 1, in = i + 1;
        while (ip >= 0 && b[ip] <= a[ip]) --ip;
        while (in < n && b[in] <= a[in]) ++in;
        const uint8_t* lo = (hl && ip >= 0) ? band_px[ip] + (b[ip] - a[ip] - hl) * stride2_bytes : nullptr;
        const uint8_t* hi = (hh && in < n) ? band_px[in] : nullptr;
        cudaStream_t s = nullptr;   // the legacy default stream of the band's device: ordered after the caller's work
        r = mr_categorize_clean_dev_peer(m->ctx[i], band_px[i], n1, b[i] - a[i], stride2_bytes, bytes_per_px, R, lo, hl,
                                         hi, hh, band_out[i], out_stride2, s);
        if (r) return r;
        if (cudaSetDevice(m->dev[i]) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
            (void)cudaGetLastError();
            return MR_ERR_CUDA;
        }
        return MR_OK;
    });
    return rc;
}
