// tvd_abi.cu -- the C ABI of libtvd.so (include/tvd.h): model handles, per-device contexts,
// workspaces and the orchestration  K1 -> K4 -> Phase A -> sort -> Phase B -> Phase C.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "tvd_internal.cuh"

// ---------------------------------------------------------------------------------------------
// errors, counters
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_far_mu;
static double g_far_ms = 0.0;
static int64_t g_far_launches = 0;
static double g_near_ms = 0.0, g_post_ms = 0.0;     // near-field walks; everything after the sweep

void tvd_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int set_error(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _buf[512];                                                                     \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                       \
            g_last_error = _buf;                                                                \
            cudaGetLastError();                                                                 \
            return (int)_e;                                                                     \
        }                                                                                       \
    } while (0)

// RAII: entry points that switch the CUDA device restore the caller's device on return
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// grow-only device buffer.  Allocation and release are stream-ordered (cudaMallocAsync pool):
// no device-wide synchronisation, and the memory of a released buffer is reused by the next
// model of the process (a PreparedModel of 72 tesseroids costs microseconds to set up, not
// fifteen cudaMalloc/cudaFree round trips).
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes, cudaStream_t stream)
    {
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFreeAsync(p, stream);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e == cudaSuccess)
            cap = bytes;
        return e;
    }
    void release(cudaStream_t stream)
    {
        if (p)
            cudaFreeAsync(p, stream);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct DeviceCtx {
    int device = -1;
    bool has_table = false;
    PieceTable tab;
    bool glq_packed = false;
    int packed_nf = 0;
    double packed_ratio[TVD_MAX_FUSED] = {0};
    Buf test_rec, glq_rec, geom, tile_rec, ab_rec, wz_rec;
    bool ab_packed = false;
    double ab_radius = 0.0;
    Buf far_out, queue, keys_sorted, pair_result, sort_temp, counters, near_ws;
    Buf pts, result, counts, jac_pieces, jac_out;
    unsigned long long qcap_hint = 0;
    int64_t packed_ppg = -1;
    bool defer_sync = false;            // forward_core leaves the final synchronisation to its caller
    unsigned long long last_nq[TVD_MAX_FUSED] = {0};
    cudaStream_t stream = nullptr;
    void *h_pinned = nullptr;           // small pinned staging area (counters, flags)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_near[2 * TVD_MAX_FUSED] = {nullptr}, ev_end = nullptr;
    cudaEvent_t ev_ready = nullptr;     // piece table and packed records are complete
    std::mutex mu;
};

struct tvd_model {
    int64_t n_tess = 0, n_pieces = 0;
    double delta = 0.0;
    int has_delta = 0;
    int primary = -1;                   // device that ran K1; the other devices copy from it
    std::vector<double> h_top, h_bottom;        // host copy of the split list, fetched on demand
    std::vector<int32_t> h_parent;
    bool h_valid = false;
    std::map<int, DeviceCtx *> ctx;
    std::mutex mu;
};

// pinned staging buffers, recycled between models (pinning memory costs ~100 us)
#define TVD_PINNED_BYTES ((size_t)1 << 20)
#define TVD_PINNED_HEAD 4096            // counters and flags live in the first 4 KB
static std::mutex g_pinned_mu;
static std::vector<void *> g_pinned_free;
static void *pinned_get()
{
    {
        std::lock_guard<std::mutex> lock(g_pinned_mu);
        if (!g_pinned_free.empty()) {
            void *p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, TVD_PINNED_BYTES + TVD_PINNED_HEAD) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
static void pinned_put(void *p)
{
    if (!p)
        return;
    std::lock_guard<std::mutex> lock(g_pinned_mu);
    g_pinned_free.push_back(p);
}

static void free_table(PieceTable &t, cudaStream_t stream)
{
    cudaFreeAsync(t.tess_bounds, stream);
    cudaFreeAsync(t.tess_kind, stream);
    cudaFreeAsync(t.tess_params, stream);
    cudaFreeAsync(t.tess_npieces, stream);
    cudaFreeAsync(t.tess_offset, stream);
    cudaFreeAsync(t.sweep_piece, stream);
    cudaFreeAsync(t.top, stream);
    cudaFreeAsync(t.bottom, stream);
    cudaFreeAsync(t.parent, stream);
    t = PieceTable();
}

static void destroy_ctx(DeviceCtx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    free_table(c->tab, st);
    for (Buf *b : {&c->test_rec, &c->glq_rec, &c->geom, &c->tile_rec, &c->ab_rec, &c->wz_rec, &c->far_out,
                   &c->queue, &c->keys_sorted, &c->pair_result, &c->sort_temp, &c->counters,
                   &c->near_ws, &c->pts, &c->result, &c->counts, &c->jac_pieces, &c->jac_out})
        b->release(st);
    pinned_put(c->h_pinned);
    if (c->ev0)
        cudaEventDestroy(c->ev0);
    if (c->ev1)
        cudaEventDestroy(c->ev1);
    for (cudaEvent_t ev : c->ev_near)
        if (ev)
            cudaEventDestroy(ev);
    if (c->ev_end)
        cudaEventDestroy(c->ev_end);
    if (c->ev_ready)
        cudaEventDestroy(c->ev_ready);
    if (st) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
    }
    delete c;
}

// Contexts (stream, events, pinned staging, small workspaces) are recycled between models of a
// process: the field functions build and drop a model per call, and creating a stream and
// sixteen events costs as much as the whole set-up of a small model.
#define TVD_CTX_POOL 8                      // recycled contexts kept per device
#define TVD_CTX_KEEP_BYTES ((size_t)64 << 20)   // workspaces above this size are not kept
static std::mutex g_ctx_pool_mu;
static std::vector<DeviceCtx *> g_ctx_pool[64];

static void free_ctx(DeviceCtx *c)
{
    if (!c)
        return;
    if (c->device < 0 || c->device >= 64 || !c->stream) {
        destroy_ctx(c);
        return;
    }
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    free_table(c->tab, st);                 // stream-ordered: the next user of this stream follows
    for (Buf *b : {&c->test_rec, &c->glq_rec, &c->geom, &c->tile_rec, &c->ab_rec, &c->wz_rec, &c->far_out,
                   &c->queue, &c->keys_sorted, &c->pair_result, &c->sort_temp, &c->counters,
                   &c->near_ws, &c->pts, &c->result, &c->counts, &c->jac_pieces, &c->jac_out})
        if (b->cap > TVD_CTX_KEEP_BYTES)
            b->release(st);
    c->has_table = false;
    c->glq_packed = false;
    c->packed_nf = 0;
    c->packed_ppg = -1;
    c->ab_packed = false;
    c->qcap_hint = 0;
    c->defer_sync = false;
    {
        std::lock_guard<std::mutex> lock(g_ctx_pool_mu);
        if (g_ctx_pool[c->device].size() < TVD_CTX_POOL) {
            g_ctx_pool[c->device].push_back(c);
            return;
        }
    }
    destroy_ctx(c);
}

// the tesseroid table of the model on c's device, uploaded from the given host arrays
static int upload_tess(tvd_model *m, DeviceCtx *c, const double *bounds, const int32_t *kind,
                       const double *params)
{
    PieceTable &t = c->tab;
    t.n_tess = m->n_tess;
    const size_t nt = (size_t)(m->n_tess > 0 ? m->n_tess : 1);
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMallocAsync(&t.tess_bounds, nt * 6 * sizeof(double), st));
    CUDA_TRY(cudaMallocAsync(&t.tess_kind, nt * sizeof(int32_t), st));
    CUDA_TRY(cudaMallocAsync(&t.tess_params, nt * TVD_NPAR * sizeof(double), st));
    CUDA_TRY(cudaMallocAsync(&t.tess_npieces, nt * sizeof(int32_t), st));
    if (m->n_tess > 0) {
        CUDA_TRY(cudaMemcpyAsync(t.tess_bounds, bounds, m->n_tess * 6 * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(t.tess_kind, kind, m->n_tess * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(t.tess_params, params, m->n_tess * TVD_NPAR * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
    }
    return TVD_OK;
}

static int alloc_pieces(PieceTable &t, int64_t n_pieces, cudaStream_t st)
{
    t.n_pieces = n_pieces;
    const size_t np = (size_t)(n_pieces > 0 ? n_pieces : 1);
    CUDA_TRY(cudaMallocAsync(&t.top, np * sizeof(double), st));
    CUDA_TRY(cudaMallocAsync(&t.bottom, np * sizeof(double), st));
    CUDA_TRY(cudaMallocAsync(&t.parent, np * sizeof(int32_t), st));
    return TVD_OK;
}

static int new_ctx(int device, DeviceCtx **out)
{
    CUDA_TRY(cudaSetDevice(device));
    {
        // keep freed blocks of the stream-ordered pool for the next model instead of returning
        // them to the driver at every synchronisation (up to 2 GiB per device)
        static std::mutex mu;
        static bool done[64] = {false};
        std::lock_guard<std::mutex> lock(mu);
        if (device >= 0 && device < 64 && !done[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                uint64_t keep = 2ull << 30;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
            done[device] = true;
        }
    }
    if (device >= 0 && device < 64) {
        std::lock_guard<std::mutex> lock(g_ctx_pool_mu);
        if (!g_ctx_pool[device].empty()) {
            *out = g_ctx_pool[device].back();
            g_ctx_pool[device].pop_back();
            return TVD_OK;
        }
    }
    DeviceCtx *c = new DeviceCtx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaEventCreateWithFlags(&c->ev0, cudaEventDefault);
    if (e == cudaSuccess)
        e = cudaEventCreate(&c->ev1);
    for (int i = 0; i < 2 * TVD_MAX_FUSED && e == cudaSuccess; i++)
        e = cudaEventCreate(&c->ev_near[i]);
    if (e == cudaSuccess)
        e = cudaEventCreate(&c->ev_end);
    if (e == cudaSuccess)
        e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        c->h_pinned = pinned_get();
        if (!c->h_pinned)
            e = cudaErrorMemoryAllocation;
    }
    if (e != cudaSuccess) {
        destroy_ctx(c);
        CUDA_TRY(e);
    }
    *out = c;
    return TVD_OK;
}

static int group_period(int64_t n_pieces);

// K4, model part: sweep order and GLQ records (once per model and device)
static int ensure_glq(tvd_model *m, DeviceCtx *c, cudaStream_t stream)
{
    if (c->glq_packed || m->n_pieces == 0)
        return TVD_OK;
    const int64_t Q = m->n_pieces;
    CUDA_TRY(tvd_build_sweep_order(c->tab, stream));
    CUDA_TRY(c->glq_rec.reserve((size_t)Q * TVD_GLQ_REC * sizeof(double), stream));
    CUDA_TRY(c->geom.reserve((size_t)Q * 5 * sizeof(double), stream));
    CUDA_TRY(tvd_launch_pack_glq(c->tab, group_period(Q), c->glq_rec.as<double>(),
                                 c->geom.as<double>(), stream));
    c->glq_packed = true;
    return TVD_OK;
}

// Context of `device`.  A device other than the one that built the model receives the piece
// table AND the packed sweep records by device-to-device copies (NVLink when peer access is
// available) instead of a host upload followed by its own sorts and K4.
static int get_ctx(tvd_model *m, int device, DeviceCtx **out)
{
    DeviceCtx *src = nullptr;
    {
        std::lock_guard<std::mutex> lock(m->mu);
        auto it = m->ctx.find(device);
        if (it != m->ctx.end()) {
            *out = it->second;
            return TVD_OK;
        }
        src = m->ctx[m->primary];
    }
    // replicate outside the model lock so that several devices are populated concurrently
    const bool timing = getenv("TVD_DEBUG_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing)
            return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[get_ctx dev %d] %-32s %9.1f us\n", device, what,
                std::chrono::duration<double, std::micro>(now - t_prev).count());
        t_prev = now;
    };
    DeviceCtx *c = nullptr;
    int rc = new_ctx(device, &c);
    if (rc)
        return rc;
    lap("new_ctx (context, stream, events)");
    {
        // the source finishes its own records first; its stream signals when they are complete
        std::lock_guard<std::mutex> lock(src->mu);
        cudaError_t e = cudaSetDevice(src->device);
        if (e == cudaSuccess) {
            rc = ensure_glq(m, src, src->stream);
            if (!rc)
                e = cudaEventRecord(src->ev_ready, src->stream);
        }
        if (e == cudaSuccess && !rc)
            e = cudaSetDevice(device);
        lap("source records ready (lock + K4)");
        if (e == cudaSuccess && !rc) {
            int can = 0;
            if (!getenv("TVD_NO_PEER_ACCESS") &&
                cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can)
                cudaDeviceEnablePeerAccess(src->device, 0);     // "already enabled" is fine
            cudaGetLastError();
            e = cudaStreamWaitEvent(c->stream, src->ev_ready, 0);
        }
        lap("peer access");
        if (e != cudaSuccess && !rc) {
            g_last_error = std::string("piece table replication failed: ") + cudaGetErrorString(e);
            rc = (int)e;
        }
    }
    if (!rc) {
        PieceTable &t = c->tab;
        const PieceTable &o = src->tab;
        cudaStream_t st = c->stream;
        const size_t nt = (size_t)(m->n_tess > 0 ? m->n_tess : 1);
        const size_t np = (size_t)(m->n_pieces > 0 ? m->n_pieces : 1);
        t.n_tess = m->n_tess;
        t.n_pieces = m->n_pieces;
        cudaError_t e = cudaSuccess;
        auto clone = [&](auto *&dst, const auto *from, size_t count) {
            if (e != cudaSuccess)
                return;
            e = cudaMallocAsync((void **)&dst, count * sizeof(*from), st);
            if (e == cudaSuccess && from)
                e = cudaMemcpyPeerAsync(dst, device, from, src->device, count * sizeof(*from), st);
        };
        clone(t.tess_bounds, o.tess_bounds, nt * 6);
        clone(t.tess_kind, o.tess_kind, nt);
        clone(t.tess_params, o.tess_params, nt * TVD_NPAR);
        clone(t.tess_npieces, o.tess_npieces, nt);
        clone(t.tess_offset, o.tess_offset, nt + 1);
        clone(t.top, o.top, np);
        clone(t.bottom, o.bottom, np);
        clone(t.parent, o.parent, np);
        if (m->n_pieces > 0) {
            clone(t.sweep_piece, o.sweep_piece, np);
            if (e == cudaSuccess)
                e = c->glq_rec.reserve(np * TVD_GLQ_REC * sizeof(double), st);
            if (e == cudaSuccess)
                e = c->geom.reserve(np * 5 * sizeof(double), st);
            if (e == cudaSuccess)
                e = cudaMemcpyPeerAsync(c->glq_rec.p, device, src->glq_rec.p, src->device,
                                        np * TVD_GLQ_REC * sizeof(double), st);
            if (e == cudaSuccess)
                e = cudaMemcpyPeerAsync(c->geom.p, device, src->geom.p, src->device,
                                        np * 5 * sizeof(double), st);
            if (e == cudaSuccess)
                c->glq_packed = true;
        }
        lap("copies enqueued");
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);      // once per (model, device)
        lap("copies done");
        if (e != cudaSuccess) {
            g_last_error = std::string("piece table replication failed: ") + cudaGetErrorString(e);
            cudaGetLastError();
            rc = (int)e;
        }
    }
    if (rc) {
        free_ctx(c);
        return rc;
    }
    c->has_table = true;
    {
        std::lock_guard<std::mutex> lock(m->mu);
        auto it = m->ctx.find(device);
        if (it != m->ctx.end()) {       // another thread populated this device meanwhile
            *out = it->second;
            free_ctx(c);
            return TVD_OK;
        }
        m->ctx[device] = c;
    }
    *out = c;
    return TVD_OK;
}

// ---------------------------------------------------------------------------------------------
// ABI: misc
// ---------------------------------------------------------------------------------------------
extern "C" int tvd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" const char *tvd_last_error(void) { return g_last_error.c_str(); }

extern "C" int tvd_warm_devices(int n_devices, const int *device_ids)
{
    // CUDA contexts are created lazily, ~0.3 s each; creating those of all devices at once (and
    // while the caller is still flattening its model) takes them off the first call's critical path
    const int avail = tvd_device_count();
    if (n_devices < 1 || avail == 0)
        return TVD_OK;
    std::vector<std::thread> threads;
    for (int d = 0; d < n_devices; d++) {
        const int id = device_ids ? device_ids[d] : d;
        if (id < 0 || id >= avail)
            continue;
        threads.emplace_back([id]() {
            if (cudaSetDevice(id) == cudaSuccess)
                cudaFree(nullptr);
            cudaGetLastError();
        });
    }
    for (auto &t : threads)
        t.join();
    return TVD_OK;
}

extern "C" int64_t tvd_launch_count(int reset)
{
    if (reset)
        return g_launches.exchange(0);
    return g_launches.load();
}

extern "C" double tvd_far_kernel_ms(int reset, int64_t *n_launches)
{
    std::lock_guard<std::mutex> lock(g_far_mu);
    const double ms = g_far_ms;
    if (n_launches)
        *n_launches = g_far_launches;
    if (reset) {
        g_far_ms = g_near_ms = g_post_ms = 0.0;
        g_far_launches = 0;
    }
    return ms;
}

extern "C" void tvd_phase_ms(int reset, double *out)
{
    std::lock_guard<std::mutex> lock(g_far_mu);
    if (out) {
        out[0] = g_far_ms;
        out[1] = g_near_ms;
        out[2] = g_post_ms;
        out[3] = (double)g_far_launches;
    }
    if (reset) {
        g_far_ms = g_near_ms = g_post_ms = 0.0;
        g_far_launches = 0;
    }
}

extern "C" double tvd_measure_fp64_peak(int device, int iters)
{
    DeviceGuard guard;
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return -1.0;
    }
    return tvd_fp64_probe(iters > 0 ? iters : 4096, nullptr);
}

extern "C" int tvd_debug_trig(int device, int64_t n, const double *x, double *cos_out,
                              double *sin_out)
{
    g_last_error.clear();
    if (n < 0 || (n > 0 && (!x || !cos_out || !sin_out)))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_debug_trig: bad argument");
    if (tvd_device_count() <= device || device < 0)
        return set_error(TVD_ERR_NO_DEVICE, "tvd_debug_trig: no such CUDA device");
    if (n == 0)
        return TVD_OK;
    DeviceGuard guard;
    CUDA_TRY(cudaSetDevice(device));
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)n * 3 * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = tvd_debug_trig_launch(n, d, d + n, d + 2 * n, nullptr);
    if (e == cudaSuccess)
        e = cudaMemcpy(cos_out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        e = cudaMemcpy(sin_out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    CUDA_TRY(e);
    return TVD_OK;
}

extern "C" int tvd_debug_math(int device, int func, int64_t n, const double *x, double *out)
{
    g_last_error.clear();
    if (n < 0 || func < 0 || func > 7 || (n > 0 && (!x || !out)))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_debug_math: bad argument");
    if (tvd_device_count() <= device || device < 0)
        return set_error(TVD_ERR_NO_DEVICE, "tvd_debug_math: no such CUDA device");
    if (n == 0)
        return TVD_OK;
    DeviceGuard guard;
    CUDA_TRY(cudaSetDevice(device));
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)n * 2 * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = tvd_debug_math_launch(func, n, d, d + n, nullptr);
    if (e == cudaSuccess)
        e = cudaMemcpy(out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    CUDA_TRY(e);
    return TVD_OK;
}

// ---------------------------------------------------------------------------------------------
// ABI: model
// ---------------------------------------------------------------------------------------------
extern "C" int tvd_model_create(int64_t n_tess, const double *bounds, const int32_t *kind,
                                const double *params, double delta, int device, tvd_model **out)
{
    g_last_error.clear();
    if (!out || n_tess < 0 || (n_tess > 0 && (!bounds || !kind || !params)))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_model_create: bad argument");
    if (n_tess > 0x7fffffffLL)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_model_create: more than 2^31-1 tesseroids");
    *out = nullptr;
    if (tvd_device_count() <= device || device < 0)
        return set_error(TVD_ERR_NO_DEVICE, "tvd_model_create: no such CUDA device (no CPU fallback)");
    for (int64_t i = 0; i < n_tess; i++)
        if (kind[i] < TVD_DENS_CONST || kind[i] > TVD_DENS_SINE)
            return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_model_create: unknown density kind");

    // TVD_DEBUG_TIMING=1 prints the host time of each stage (development aid)
    const bool timing = getenv("TVD_DEBUG_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing)
            return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[tvd_model_create] %-28s %8.1f us\n", what,
                std::chrono::duration<double, std::micro>(now - t_prev).count());
        t_prev = now;
    };
    DeviceGuard guard;
    tvd_model *m = new tvd_model();
    m->n_tess = n_tess;
    m->has_delta = std::isnan(delta) ? 0 : 1;
    m->delta = m->has_delta ? delta : 0.0;
    m->primary = device;

    DeviceCtx *c = nullptr;
    int rc = new_ctx(device, &c);
    if (rc) {
        delete m;
        return rc;
    }
    lap("context (stream, events)");
    cudaStream_t st = c->stream;
    int *d_status = nullptr;
    int64_t *d_offsets = nullptr;
    double *d_keep = nullptr;
    auto fail = [&](int code) {
        cudaFreeAsync(d_status, st);
        cudaFreeAsync(d_keep, st);
        if (c->tab.tess_offset != d_offsets)
            cudaFreeAsync(d_offsets, st);
        free_ctx(c);
        delete m;
        return code;
    };
    rc = upload_tess(m, c, bounds, kind, params);
    if (rc)
        return fail(rc);
    lap("tesseroid table upload");
#define MODEL_TRY(expr)                                                                  \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            g_last_error = std::string(#expr " failed: ") + cudaGetErrorString(_e);      \
            cudaGetLastError();                                                          \
            return fail((int)_e);                                                        \
        }                                                                                \
    } while (0)
    // pinned landing area for the two scalars the host needs: [0] total pieces, [1] status
    int64_t *h_total = reinterpret_cast<int64_t *>(c->h_pinned);
    int *h_status = reinterpret_cast<int *>(h_total + 1);
    MODEL_TRY(cudaMallocAsync(&d_status, sizeof(int), st));
    MODEL_TRY(cudaMallocAsync(&d_offsets, (size_t)(n_tess + 1) * sizeof(int64_t), st));
    MODEL_TRY(cudaMallocAsync(&d_keep, (size_t)(n_tess > 0 ? n_tess : 1) * TVD_K1_KEEP * 2 *
                                           sizeof(double), st));
    MODEL_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    MODEL_TRY(tvd_launch_k1_count(c->tab, m->delta, m->has_delta, d_status, d_keep, st));
    MODEL_TRY(tvd_launch_scan(c->tab.tess_npieces, d_offsets, n_tess, st));
    MODEL_TRY(cudaMemcpyAsync(h_total, d_offsets + n_tess, sizeof(int64_t), cudaMemcpyDeviceToHost,
                              st));
    MODEL_TRY(cudaMemcpyAsync(h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    lap("K1 count + scan enqueued");
    MODEL_TRY(cudaStreamSynchronize(st));      // the host pages of bounds/kind/params are free again
    lap("synchronise");
    const int64_t total = *h_total;
    const int status = *h_status;
    if (status != 0) {
        set_error(status, "tvd_model_create: a tesseroid needs more than 256 radial pieces "
                          "(delta too small for this density)");
        return fail(status);
    }
    if (total > 0x7fffffffLL) {
        set_error(TVD_ERR_TOO_MANY_PIECES, "tvd_model_create: more than 2^31-1 radial pieces");
        return fail(TVD_ERR_TOO_MANY_PIECES);
    }
    m->n_pieces = total;
    rc = alloc_pieces(c->tab, total, st);
    if (rc)
        return fail(rc);
    MODEL_TRY(tvd_launch_k1_emit(c->tab, d_offsets, m->delta, m->has_delta, d_keep, st));
    c->tab.tess_offset = d_offsets;     // owned by the piece table from here on
    MODEL_TRY(cudaFreeAsync(d_status, st));
    d_status = nullptr;
    MODEL_TRY(cudaFreeAsync(d_keep, st));
    d_keep = nullptr;
    // no synchronisation here: everything that uses the table is ordered after it, either on
    // this stream or behind a synchronisation of it (get_ctx, forward_core)
    MODEL_TRY(cudaEventRecord(c->ev_ready, st));
    lap("K1 emit enqueued");
#undef MODEL_TRY
    c->has_table = true;
    m->ctx[device] = c;
    *out = m;
    return TVD_OK;
}

extern "C" void tvd_model_free(tvd_model *m)
{
    if (!m)
        return;
    DeviceGuard guard;
    for (auto &kv : m->ctx)
        free_ctx(kv.second);
    delete m;
}

extern "C" int64_t tvd_model_num_pieces(const tvd_model *m) { return m ? m->n_pieces : 0; }

extern "C" int tvd_model_get_pieces(const tvd_model *cm, double *top, double *bottom,
                                    int32_t *parent)
{
    tvd_model *m = const_cast<tvd_model *>(cm);
    if (!m)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_model_get_pieces: null model");
    {
        std::lock_guard<std::mutex> lock(m->mu);
        if (!m->h_valid) {
            DeviceGuard guard;
            DeviceCtx *c = m->ctx[m->primary];
            std::lock_guard<std::mutex> lc(c->mu);
            CUDA_TRY(cudaSetDevice(c->device));
            const int64_t n = m->n_pieces;
            m->h_top.resize(n);
            m->h_bottom.resize(n);
            m->h_parent.resize(n);
            if (n > 0) {
                CUDA_TRY(cudaMemcpyAsync(m->h_top.data(), c->tab.top, n * sizeof(double),
                                         cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaMemcpyAsync(m->h_bottom.data(), c->tab.bottom, n * sizeof(double),
                                         cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaMemcpyAsync(m->h_parent.data(), c->tab.parent, n * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, c->stream));
            }
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            m->h_valid = true;
        }
    }
    if (top)
        memcpy(top, m->h_top.data(), m->n_pieces * sizeof(double));
    if (bottom)
        memcpy(bottom, m->h_bottom.data(), m->n_pieces * sizeof(double));
    if (parent)
        memcpy(parent, m->h_parent.data(), m->n_pieces * sizeof(int32_t));
    return TVD_OK;
}

// ---------------------------------------------------------------------------------------------
// summation plan
// ---------------------------------------------------------------------------------------------
// Groups of the piece table start at multiples of this many pieces (it divides TVD_TILE, so
// every shared-memory tile of a group starts on it too; the "same tesseroid / same longitude
// extent" flags of the GLQ records are reset there).  Small models get a finer period so that a
// 72-tesseroid shell is swept by a dozen thread blocks instead of one.
static int group_period(int64_t n_pieces) { return n_pieces <= 16384 ? 16 : TVD_TILE; }

static int plan_groups(int64_t n_pieces, int64_t n_points_total)
{
    // Pure function of (Q, P_total): every rank / device of a job derives the same plan, and with
    // it the same summation order.  A work unit of the sweep is (512 points) x (one group); the
    // plan aims at ~128 units per SM of ONE device for the whole job, so that a job split over 8
    // devices still gives each of them >= 16 units per SM (the last wave of units is then a few
    // per cent of the run, whatever the number of points per device; at 490 units on 148 SMs it
    // was 17 %).
    const int64_t ppb = (int64_t)TVD_FAR_THREADS;
    const int64_t blocks_x = (n_points_total + ppb - 1) / ppb;
    if (blocks_x <= 0 || n_pieces <= 0)
        return 1;
    const int64_t period = group_period(n_pieces);
    int64_t gmax = (n_pieces + period - 1) / period;
    if (gmax > 4096)
        gmax = 4096;                    // Phase C adds the group sums of a point one by one
    const int64_t target = 148 * 128;
    int64_t g = (target + blocks_x - 1) / blocks_x;
    if (g > gmax)
        g = gmax;
    if (g < 1)
        g = 1;
    return (int)g;
}

extern "C" int tvd_plan_groups(const tvd_model *m, int64_t n_points_total)
{
    return m ? plan_groups(m->n_pieces, n_points_total) : 1;
}

// ---------------------------------------------------------------------------------------------
// forward on one device (device pointers), the core: nf fields in one far-field sweep
// ---------------------------------------------------------------------------------------------
struct FieldSet {
    int nf = 0;
    int field[TVD_MAX_FUSED];       // ascending field ids (= the sweep's slot order)
    double ratio[TVD_MAX_FUSED];
    int mask = 0;
};

// after the stream of a forward call has drained: timings, error flags, statistics
static int forward_finish(DeviceCtx *c, const FieldSet &fs, int64_t Q, int64_t P,
                          const unsigned long long *nq, int64_t *stats)
{
    const int nf = fs.nf;
    const unsigned long long *h_cnt = reinterpret_cast<unsigned long long *>(c->h_pinned) + 8;
    const unsigned long long all_pairs = (unsigned long long)P * (unsigned long long)Q;
    {
        float ms = 0.f;
        double near_ms = 0.0, post_ms = 0.0;
        for (int f = 0; f < nf; f++)
            if (nq[f] > 0 &&
                cudaEventElapsedTime(&ms, c->ev_near[2 * f], c->ev_near[2 * f + 1]) == cudaSuccess)
                near_ms += ms;
        if (cudaEventElapsedTime(&ms, c->ev1, c->ev_end) == cudaSuccess)
            post_ms = ms;
        std::lock_guard<std::mutex> lock(g_far_mu);
        g_near_ms += near_ms;
        g_post_ms += post_ms;
    }
    for (int f = 0; f < nf; f++)
        if (h_cnt[8 + 8 * f + 4] != 0)
            return set_error(TVD_ERR_STACK_OVERFLOW, "Tesseroid stack overflow");
    if (stats) {
        for (int f = 0; f < nf; f++) {
            const unsigned long long *hc = h_cnt + 8 + 8 * f;
            int64_t *st = stats + (size_t)f * TVD_NSTATS;
            const int64_t far_pairs = (int64_t)(all_pairs - nq[f]);
            const int64_t near_nodes = (int64_t)hc[0], near_leaves = (int64_t)hc[1];
            st[TVD_ST_NODES] = far_pairs + near_nodes;
            st[TVD_ST_LEAVES] = far_pairs + near_leaves;
            st[TVD_ST_SPLITS] = (int64_t)hc[2];
            st[TVD_ST_TOO_SMALL] = (int64_t)hc[3];
            st[TVD_ST_PIECES] = Q;
            st[TVD_ST_NONROOT_NODES] = near_nodes - (int64_t)nq[f];
            // hc[5] = walked pairs whose root was integrated after all (one root leaf each)
            st[TVD_ST_NONROOT_LEAVES] = near_leaves - (int64_t)hc[5];
            st[TVD_ST_NEAR_PAIRS] = (int64_t)nq[f];
        }
    }
    return TVD_OK;
}

// what the caller knows about the radii of the points: all bitwise equal to `r0`, not all equal,
// or unknown (device-resident points: a kernel checks)
struct RadiusInfo {
    int uniform = -1;
    double r0 = 0.0;
};

static int forward_core(tvd_model *m, DeviceCtx *c, const FieldSet &fs, int64_t P,
                        const double *d_lon, const double *d_sinlat, const double *d_coslat,
                        const double *d_radius, int stack_size, int n_groups,
                        double *const *d_results /*[nf]*/, int64_t *stats /*[nf][TVD_NSTATS]*/,
                        int32_t *d_leaf_counts, cudaStream_t stream, RadiusInfo rinfo,
                        double *d_jac_out = nullptr /*sensitivity mode: [n_tess][P], nf == 1*/)
{
    const int64_t Q = m->n_pieces;
    const int nf = fs.nf;
    if (stats)
        memset(stats, 0, (size_t)nf * TVD_NSTATS * sizeof(int64_t));
    if (P == 0)
        return TVD_OK;
    if (P > 0x7fffffffLL)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward: more than 2^31-1 points per device");
    // the piece table was produced on the context's own stream (tvd_model_create, get_ctx)
    if (stream != c->stream)
        CUDA_TRY(cudaStreamWaitEvent(stream, c->ev_ready, 0));
    if (Q == 0) {
        for (int f = 0; f < nf; f++)
            CUDA_TRY(cudaMemsetAsync(d_results[f], 0, P * sizeof(double), stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        return TVD_OK;
    }
    if (n_groups <= 0)
        n_groups = plan_groups(Q, P);
    const int period = group_period(Q);
    int64_t ppg = (Q + n_groups - 1) / n_groups;
    ppg = ((ppg + period - 1) / period) * period;
    // groups that would start beyond Q are dropped (they would add exact zeros)
    n_groups = (int)((Q + ppg - 1) / ppg);
    const int64_t tiles_per_group = (ppg + TVD_TILE - 1) / TVD_TILE;

    // K4: GLQ records once per model, test records per (field set, ratios), tile records per
    // (field set, ratios, grouping)
    {
        const int rc = ensure_glq(m, c, stream);
        if (rc)
            return rc;
    }
    bool same_test = c->packed_nf == nf;
    for (int f = 0; f < nf && same_test; f++)
        same_test = c->packed_ratio[f] == fs.ratio[f];
    if (!same_test || c->packed_ppg != ppg) {
        RatioList rl;
        rl.nf = nf;
        for (int f = 0; f < TVD_MAX_FUSED; f++)
            rl.ratio[f] = f < nf ? fs.ratio[f] : 0.0;
        if (!same_test) {
            CUDA_TRY(c->test_rec.reserve((size_t)Q * tvd_test_rec_len(nf) * sizeof(double), stream));
            CUDA_TRY(tvd_launch_pack_test(c->tab, c->geom.as<double>(), rl,
                                          c->test_rec.as<double>(), stream));
        }
        CUDA_TRY(c->tile_rec.reserve((size_t)n_groups * tiles_per_group * TVD_TILE_REC *
                                     sizeof(double), stream));
        CUDA_TRY(tvd_launch_pack_tiles(c->tab, c->geom.as<double>(), rl, ppg, n_groups,
                                       c->tile_rec.as<double>(), stream));
        c->packed_nf = nf;
        c->packed_ppg = ppg;
        for (int f = 0; f < nf; f++)
            c->packed_ratio[f] = fs.ratio[f];
    }
    const int n_cnt = 8 + 8 * TVD_MAX_FUSED;
    CUDA_TRY(c->counters.reserve(n_cnt * sizeof(unsigned long long), stream));
    unsigned long long *d_cnt = c->counters.as<unsigned long long>();
    // pinned landing area of the device counters (TVD_PINNED_HEAD bytes)
    unsigned long long *h_cnt = reinterpret_cast<unsigned long long *>(c->h_pinned) + 8;
    // points at one height: 2r*rc and r^2 + rc^2 per piece instead of per pair
    const double *d_ab = nullptr, *d_wz = nullptr;
    if (!d_jac_out) {
        if (rinfo.uniform < 0) {
            int *d_flag = reinterpret_cast<int *>(d_cnt + 7);
            int *h_flag = reinterpret_cast<int *>(c->h_pinned);
            double *h_r0 = reinterpret_cast<double *>(c->h_pinned) + 1;
            CUDA_TRY(tvd_launch_check_uniform(P, d_radius, d_flag, stream));
            CUDA_TRY(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(h_r0, d_radius, sizeof(double), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            rinfo.uniform = *h_flag == 1 ? 1 : 0;
            rinfo.r0 = *h_r0;
        }
        if (rinfo.uniform == 1 && rinfo.r0 == rinfo.r0) {
            if (!c->ab_packed || memcmp(&c->ab_radius, &rinfo.r0, sizeof(double)) != 0) {
                CUDA_TRY(c->ab_rec.reserve((size_t)Q * TVD_AB_REC * sizeof(double), stream));
                CUDA_TRY(c->wz_rec.reserve((size_t)Q * TVD_WZ_REC * sizeof(double), stream));
                CUDA_TRY(tvd_launch_pack_ab(c->tab, c->glq_rec.as<double>(), rinfo.r0,
                                            c->ab_rec.as<double>(), c->wz_rec.as<double>(),
                                            stream));
                c->ab_packed = true;
                c->ab_radius = rinfo.r0;
            }
            d_ab = c->ab_rec.as<double>();
            d_wz = c->wz_rec.as<double>();
        }
    }
    CUDA_TRY(c->far_out.reserve((size_t)nf * n_groups * P * sizeof(double), stream));
    if (d_jac_out)
        CUDA_TRY(c->jac_pieces.reserve((size_t)Q * P * sizeof(double), stream));
    const unsigned long long all_pairs = (unsigned long long)P * (unsigned long long)Q;
    unsigned long long qcap = c->qcap_hint;
    if (qcap == 0) {
        // first guess: 64 deferred pairs per point and field, at least 2^20, at most 2 GiB in all
        // (an overflow is detected below and repaired by a test-only sweep with exact sizes)
        qcap = 64ull * (unsigned long long)P;
        if (qcap < (1ull << 20))
            qcap = 1ull << 20;
        const unsigned long long cap2g = (2ull << 30) / (8ull * (unsigned long long)nf);
        if (qcap > cap2g)
            qcap = cap2g;
    }
    if (qcap > all_pairs)
        qcap = all_pairs;
    if (const char *dbg = getenv("TVD_DEBUG_QCAP"))     // tests force the overflow path with this
        qcap = strtoull(dbg, nullptr, 10) > 0 ? strtoull(dbg, nullptr, 10) : qcap;

    // [0..nf) = queue counts, [8 + 8 f ..) = near-field counters of field slot f
    unsigned long long nq[TVD_MAX_FUSED] = {0};
    for (int attempt = 0; attempt < 2; attempt++) {
        CUDA_TRY(c->queue.reserve((size_t)nf * qcap * sizeof(unsigned long long), stream));
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, n_cnt * sizeof(unsigned long long), stream));
        FarArgs fa;
        fa.mask = fs.mask;
        fa.n_points = P;
        fa.lon = d_lon;
        fa.sinlat = d_sinlat;
        fa.coslat = d_coslat;
        fa.radius = d_radius;
        fa.n_pieces = Q;
        fa.test_rec = c->test_rec.as<double>();
        fa.glq_rec = c->glq_rec.as<double>();
        fa.tile_rec = c->tile_rec.as<double>();
        fa.ab_rec = d_ab;
        fa.wz_rec = d_wz;
        fa.n_groups = n_groups;
        fa.pieces_per_group = ppg;
        fa.tiles_per_group = (int)tiles_per_group;
        fa.far_out = c->far_out.as<double>();
        fa.queue = c->queue.as<unsigned long long>();
        fa.qcount = d_cnt;
        fa.qcap = qcap;
        fa.jac = d_jac_out ? c->jac_pieces.as<double>() : nullptr;
        if (attempt == 0) {
            CUDA_TRY(cudaEventRecord(c->ev0, stream));
            CUDA_TRY(tvd_launch_far(fa, stream));
            CUDA_TRY(cudaEventRecord(c->ev1, stream));
        } else {
            // the far sums of the first sweep stand; only the queues are rebuilt
            CUDA_TRY(tvd_launch_queue_only(fa, stream));
        }
        CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, nf * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (int f = 0; f < nf; f++)
            nq[f] = h_cnt[f];
        float ms = 0.f;
        if (attempt == 0 && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) {
            std::lock_guard<std::mutex> lock(g_far_mu);
            g_far_ms += ms;
            g_far_launches += 1;
        }
        unsigned long long worst = 0;
        for (int f = 0; f < nf; f++)
            worst = nq[f] > worst ? nq[f] : worst;
        if (worst <= qcap)
            break;
        // queue overflow: the exact sizes are now known, refill the queues with a test-only sweep
        qcap = worst;
        c->qcap_hint = worst + worst / 4;
        if (attempt == 1)
            return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward: near-field queue overflow twice");
    }

    int pbits = 1;
    while (pbits < 32 && (1ll << pbits) < P)
        pbits++;
    {
        // sort / walk workspaces once, for the largest queue of the set
        unsigned long long nmax = 0;
        for (int f = 0; f < nf; f++)
            nmax = nq[f] > nmax ? nq[f] : nmax;
        if (nmax > 0) {
            size_t temp_bytes = 0;
            CUDA_TRY(tvd_sort_keys(nullptr, temp_bytes, c->queue.as<unsigned long long>(),
                                   c->keys_sorted.as<unsigned long long>(), (int64_t)nmax,
                                   32 + pbits, stream));
            CUDA_TRY(c->keys_sorted.reserve((size_t)nmax * sizeof(unsigned long long), stream));
            CUDA_TRY(c->pair_result.reserve((size_t)nmax * sizeof(double), stream));
            CUDA_TRY(c->sort_temp.reserve(temp_bytes ? temp_bytes : 16, stream));
            CUDA_TRY(c->near_ws.reserve(tvd_near_workspace_bytes((int64_t)nmax), stream));
        }
    }
    for (int f = 0; f < nf; f++) {
        const unsigned long long *d_keys = nullptr;
        const unsigned long long n = nq[f];
        if (n > 0) {
            const unsigned long long *q_in = c->queue.as<unsigned long long>() + (size_t)f * qcap;
            size_t temp_bytes = c->sort_temp.cap;
            CUDA_TRY(tvd_sort_keys(c->sort_temp.p, temp_bytes, q_in,
                                   c->keys_sorted.as<unsigned long long>(), (int64_t)n, 32 + pbits,
                                   stream));
            d_keys = c->keys_sorted.as<unsigned long long>();
            NearArgs na;
            na.field = fs.field[f];
            na.n_pairs = (int64_t)n;
            na.keys = d_keys;
            na.lon = d_lon;
            na.sinlat = d_sinlat;
            na.coslat = d_coslat;
            na.radius = d_radius;
            na.table = c->tab;
            na.ratio = fs.ratio[f];
            na.stack_size = stack_size;
            na.pair_result = c->pair_result.as<double>();
            na.counters = d_cnt + 8 + 8 * f;
            na.leaf_counts = d_leaf_counts;
            na.n_points = P;
            na.batch = 1;
            na.workspace = c->near_ws.p;
            CUDA_TRY(cudaEventRecord(c->ev_near[2 * f], stream));
            CUDA_TRY(tvd_launch_near(na, stream));
            CUDA_TRY(cudaEventRecord(c->ev_near[2 * f + 1], stream));
        }
        if (d_jac_out)
            CUDA_TRY(tvd_launch_jac_reduce(c->tab, P, c->jac_pieces.as<double>(), (int64_t)n,
                                           d_keys, c->pair_result.as<double>(), d_jac_out, stream));
        else
            CUDA_TRY(tvd_launch_combine(P, n_groups,
                                        c->far_out.as<double>() + (size_t)f * n_groups * P,
                                        (int64_t)n, d_keys, c->pair_result.as<double>(),
                                        d_results[f], stream));
        // keys_sorted / pair_result are reused by the next field: the stream orders the kernels
    }
    CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, n_cnt * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(c->ev_end, stream));
    for (int f = 0; f < TVD_MAX_FUSED; f++)
        c->last_nq[f] = nq[f];
    if (c->defer_sync)
        return TVD_OK;                  // the caller enqueues its result copies, then finishes
    CUDA_TRY(cudaStreamSynchronize(stream));
    return forward_finish(c, fs, Q, P, nq, stats);
}

// validate and order a list of fields; a set the sweep is not compiled for is rejected
static int make_field_set(int n_fields, const int *fields, const double *ratios, FieldSet &fs,
                          int *order /*[n_fields]: slot of the caller's i-th field*/)
{
    if (n_fields < 1 || n_fields > TVD_MAX_FUSED || !fields || !ratios)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward: bad field list");
    fs.nf = n_fields;
    fs.mask = 0;
    for (int i = 0; i < n_fields; i++) {
        if (fields[i] < 0 || fields[i] >= TVD_NFIELDS || !(ratios[i] > 0) ||
            (fs.mask & (1 << fields[i])))
            return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward: bad field or ratio");
        fs.mask |= 1 << fields[i];
    }
    if (!tvd_far_mask_supported(fs.mask))
        return set_error(TVD_ERR_BAD_ARGUMENT,
                         "tvd_forward_fused: this combination of fields is not compiled in");
    int slot = 0;
    for (int f = 0; f < TVD_NFIELDS; f++) {
        if (!(fs.mask & (1 << f)))
            continue;
        for (int i = 0; i < n_fields; i++)
            if (fields[i] == f) {
                fs.field[slot] = f;
                fs.ratio[slot] = ratios[i];
                order[i] = slot;
            }
        slot++;
    }
    return TVD_OK;
}

extern "C" int tvd_fused_supported(int n_fields, const int *fields)
{
    if (n_fields < 1 || n_fields > TVD_MAX_FUSED || !fields)
        return 0;
    int mask = 0;
    for (int i = 0; i < n_fields; i++) {
        if (fields[i] < 0 || fields[i] >= TVD_NFIELDS || (mask & (1 << fields[i])))
            return 0;
        mask |= 1 << fields[i];
    }
    return tvd_far_mask_supported(mask) ? 1 : 0;
}

extern "C" int tvd_forward_device(int field, tvd_model *m, int device, int64_t n_points,
                                  const double *d_lon, const double *d_sinlat,
                                  const double *d_coslat, const double *d_radius, double ratio,
                                  int stack_size, int n_groups, double *d_result, int64_t *stats,
                                  void *stream)
{
    g_last_error.clear();
    if (!m || field < 0 || field >= TVD_NFIELDS || n_points < 0 || !(ratio > 0))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward_device: bad argument");
    FieldSet fs;
    int order[1];
    int rc = make_field_set(1, &field, &ratio, fs, order);
    if (rc)
        return rc;
    DeviceGuard guard;
    DeviceCtx *c = nullptr;
    rc = get_ctx(m, device, &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(cudaSetDevice(device));
    double *res[1] = {d_result};
    return forward_core(m, c, fs, n_points, d_lon, d_sinlat, d_coslat, d_radius, stack_size,
                        n_groups, res, stats, nullptr, (cudaStream_t)stream, RadiusInfo());
}

extern "C" int tvd_forward_fused_device(int n_fields, const int *fields, const double *ratios,
                                        tvd_model *m, int device, int64_t n_points,
                                        const double *d_lon, const double *d_sinlat,
                                        const double *d_coslat, const double *d_radius,
                                        int stack_size, int n_groups, double *const *d_results,
                                        int64_t *stats, void *stream)
{
    g_last_error.clear();
    if (!m || n_points < 0 || !d_results)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward_fused_device: bad argument");
    FieldSet fs;
    int order[TVD_MAX_FUSED];
    int rc = make_field_set(n_fields, fields, ratios, fs, order);
    if (rc)
        return rc;
    DeviceGuard guard;
    DeviceCtx *c = nullptr;
    rc = get_ctx(m, device, &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(cudaSetDevice(device));
    double *res[TVD_MAX_FUSED];
    for (int i = 0; i < n_fields; i++)
        res[order[i]] = d_results[i];
    std::vector<int64_t> st((size_t)n_fields * TVD_NSTATS, 0);
    rc = forward_core(m, c, fs, n_points, d_lon, d_sinlat, d_coslat, d_radius, stack_size, n_groups,
                      res, st.data(), nullptr, (cudaStream_t)stream, RadiusInfo());
    if (!rc && stats)
        for (int i = 0; i < n_fields; i++)
            memcpy(stats + (size_t)i * TVD_NSTATS, st.data() + (size_t)order[i] * TVD_NSTATS,
                   TVD_NSTATS * sizeof(int64_t));
    return rc;
}

// are all radii of a host slice bitwise equal?  (the sweep then takes per-piece constants)
static RadiusInfo scan_radius(const double *radius, int64_t n)
{
    RadiusInfo r;
    r.uniform = 1;
    r.r0 = n > 0 ? radius[0] : 0.0;
    for (int64_t i = 1; i < n; i++)
        if (memcmp(radius + i, radius, sizeof(double)) != 0) {
            r.uniform = 0;
            break;
        }
    return r;
}

// one pass of one device: points [p0, p0 + P) of the caller's arrays, host slices in, host slices
// out (the context is locked and its device current)
static int forward_host_pass(tvd_model *m, DeviceCtx *c, const FieldSet &fs, const int *order,
                             int n_fields, int64_t P, int64_t p0, int64_t P_total,
                             const double *lon, const double *sinlat, const double *coslat,
                             const double *radius, int stack_size, int n_groups,
                             double *results /*[n_fields][P_total], caller order*/,
                             int64_t *stats /*[n_fields][TVD_NSTATS], slot order*/,
                             int32_t *leaf_counts)
{
    int rc = TVD_OK;
    if (stats)
        memset(stats, 0, (size_t)n_fields * TVD_NSTATS * sizeof(int64_t));
    if (P == 0)
        return TVD_OK;
    cudaStream_t st = c->stream;
    CUDA_TRY(c->pts.reserve((size_t)P * 4 * sizeof(double), st));
    CUDA_TRY(c->result.reserve((size_t)n_fields * P * sizeof(double), st));
    double *d_pts = c->pts.as<double>();
    const double *src[4] = {lon + p0, sinlat + p0, coslat + p0, radius + p0};
    // small slices travel through the pinned staging area: one truly asynchronous copy each way
    // instead of four (plus one per field) staged pageable copies
    double *stage = reinterpret_cast<double *>(reinterpret_cast<char *>(c->h_pinned) + TVD_PINNED_HEAD);
    const bool staged_in = (size_t)P * 4 * sizeof(double) <= TVD_PINNED_BYTES;
    const bool staged_out = (size_t)P * n_fields * sizeof(double) <= TVD_PINNED_BYTES && !leaf_counts;
    if (staged_in) {
        for (int k = 0; k < 4; k++)
            memcpy(stage + (size_t)k * P, src[k], P * sizeof(double));
        CUDA_TRY(cudaMemcpyAsync(d_pts, stage, (size_t)P * 4 * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
    } else {
        for (int k = 0; k < 4; k++)
            CUDA_TRY(cudaMemcpyAsync(d_pts + (size_t)k * P, src[k], P * sizeof(double),
                                     cudaMemcpyHostToDevice, st));
    }
    const RadiusInfo rinfo = scan_radius(radius + p0, P);
    int32_t *d_counts = nullptr;
    if (leaf_counts) {
        CUDA_TRY(c->counts.reserve((size_t)m->n_tess * P * sizeof(int32_t) + 16, st));
        d_counts = c->counts.as<int32_t>();
        CUDA_TRY(tvd_launch_init_counts(c->tab, P, d_counts, st));
    }
    double *d_res[TVD_MAX_FUSED];
    for (int f = 0; f < n_fields; f++)
        d_res[f] = c->result.as<double>() + (size_t)f * P;
    c->defer_sync = true;
    rc = forward_core(m, c, fs, P, d_pts, d_pts + P, d_pts + 2 * P, d_pts + 3 * P, stack_size,
                      n_groups, d_res, stats, d_counts, st, rinfo);
    c->defer_sync = false;
    if (rc)
        return rc;
    if (staged_out) {
        // results are contiguous [slot][P] on the device: one copy, then scatter on the host
        CUDA_TRY(cudaMemcpyAsync(stage, c->result.p, (size_t)n_fields * P * sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
    } else {
        for (int i = 0; i < n_fields; i++)
            CUDA_TRY(cudaMemcpyAsync(results + (size_t)i * P_total + p0, d_res[order[i]],
                                     P * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (leaf_counts && m->n_tess > 0)
        CUDA_TRY(cudaMemcpy2DAsync(leaf_counts + p0, (size_t)P_total * sizeof(int32_t), d_counts,
                                   (size_t)P * sizeof(int32_t), (size_t)P * sizeof(int32_t),
                                   (size_t)m->n_tess, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (staged_out)
        for (int i = 0; i < n_fields; i++)
            memcpy(results + (size_t)i * P_total + p0, stage + (size_t)order[i] * P,
                   P * sizeof(double));
    if (m->n_pieces == 0)
        return TVD_OK;
    return forward_finish(c, fs, m->n_pieces, P, c->last_nq, stats);
}

// Points of one pass.  The device buffers that grow with the number of points (points, results,
// the group sums far[nf][G][P], the near-field queues) are bounded by processing a device's share
// in slabs, the way the reference streams points one by one (_tesseroid.pyx:65): the group sums
// stay under 2 GiB and a slab under 2^24 points.  A point's value depends on (point, model,
// n_groups) only, and n_groups is the plan of the WHOLE job, so slabbing does not change a bit.
static int64_t point_slab(int nf, int n_groups)
{
    int64_t slab = (int64_t)((2ull << 30) / (8ull * (unsigned long long)nf *
                                             (unsigned long long)(n_groups > 0 ? n_groups : 1)));
    if (slab > (1ll << 24))
        slab = 1ll << 24;
    slab = slab / TVD_FAR_THREADS * TVD_FAR_THREADS;
    if (slab < TVD_FAR_THREADS)
        slab = TVD_FAR_THREADS;
    if (const char *dbg = getenv("TVD_DEBUG_POINT_SLAB"))   // tests force several slabs with this
        slab = atoll(dbg) > 0 ? atoll(dbg) : slab;
    return slab;
}

// one device's share of tvd_forward
static int forward_host_slice(tvd_model *m, int device, const FieldSet &fs, const int *order,
                              int n_fields, int64_t P, int64_t p0, int64_t P_total,
                              const double *lon, const double *sinlat, const double *coslat,
                              const double *radius, int stack_size, int n_groups,
                              double *results /*[n_fields][P_total], caller order*/,
                              int64_t *stats /*[n_fields][TVD_NSTATS], slot order*/,
                              int32_t *leaf_counts)
{
    DeviceCtx *c = nullptr;
    int rc = get_ctx(m, device, &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(cudaSetDevice(device));
    const int64_t slab = point_slab(n_fields, n_groups);
    if (P <= slab)
        return forward_host_pass(m, c, fs, order, n_fields, P, p0, P_total, lon, sinlat, coslat,
                                 radius, stack_size, n_groups, results, stats, leaf_counts);
    if (stats)
        memset(stats, 0, (size_t)n_fields * TVD_NSTATS * sizeof(int64_t));
    std::vector<int64_t> st_pass((size_t)n_fields * TVD_NSTATS, 0);
    for (int64_t s0 = 0; s0 < P; s0 += slab) {
        const int64_t n = P - s0 < slab ? P - s0 : slab;
        rc = forward_host_pass(m, c, fs, order, n_fields, n, p0 + s0, P_total, lon, sinlat, coslat,
                               radius, stack_size, n_groups, results, stats ? st_pass.data() : nullptr,
                               leaf_counts);
        if (rc)
            return rc;
        if (stats)
            for (size_t k = 0; k < st_pass.size(); k++)
                stats[k] += st_pass[k];
    }
    if (stats)
        for (int f = 0; f < n_fields; f++)
            stats[(size_t)f * TVD_NSTATS + TVD_ST_PIECES] = m->n_pieces;
    return TVD_OK;
}

static int forward_host(int n_fields, const int *fields, const double *ratios, tvd_model *m,
                        int64_t n_points, const double *lon, const double *sinlat,
                        const double *coslat, const double *radius, int stack_size,
                        int n_devices, const int *device_ids, double *results, int64_t *stats,
                        int32_t *leaf_counts)
{
    g_last_error.clear();
    if (!m || n_points < 0 || n_devices < 1 ||
        (n_points > 0 && (!lon || !sinlat || !coslat || !radius || !results)))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward: bad argument");
    FieldSet fs;
    int order[TVD_MAX_FUSED];
    int rc = make_field_set(n_fields, fields, ratios, fs, order);
    if (rc)
        return rc;
    if (leaf_counts && n_fields != 1)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_forward: leaf counts need a single field");
    const int ndev_avail = tvd_device_count();
    if (ndev_avail == 0)
        return set_error(TVD_ERR_NO_DEVICE, "tvd_forward: no CUDA device (no CPU fallback)");
    if (stats)
        memset(stats, 0, (size_t)n_fields * TVD_NSTATS * sizeof(int64_t));
    if (n_points == 0)
        return TVD_OK;
    int nparts = n_devices;
    if ((int64_t)nparts > n_points)
        nparts = (int)n_points;
    for (int d = 0; d < nparts; d++) {
        const int id = device_ids ? device_ids[d] : d;
        if (id < 0 || id >= ndev_avail)
            return set_error(TVD_ERR_NO_DEVICE, "tvd_forward: no such CUDA device");
    }
    DeviceGuard guard;
    const int n_groups = plan_groups(m->n_pieces, n_points);
    // _split_arrays (tesseroid.py:318-323): n = size//nparts, the last chunk takes the remainder
    const int64_t chunk = n_points / nparts;
    std::vector<int> rcs(nparts, 0);
    std::vector<std::string> errs(nparts);
    std::vector<std::vector<int64_t>> st(nparts,
                                         std::vector<int64_t>((size_t)n_fields * TVD_NSTATS, 0));
    auto work = [&](int d) {
        const int id = device_ids ? device_ids[d] : d;
        const int64_t p0 = (int64_t)d * chunk;
        const int64_t p1 = (d == nparts - 1) ? n_points : p0 + chunk;
        rcs[d] = forward_host_slice(m, id, fs, order, n_fields, p1 - p0, p0, n_points, lon, sinlat,
                                    coslat, radius, stack_size, n_groups, results, st[d].data(),
                                    leaf_counts);
        if (rcs[d])
            errs[d] = g_last_error;
    };
    if (nparts == 1) {
        work(0);
    } else {
        // Devices that do not hold the model yet are populated first, all at once (CUDA contexts
        // are created concurrently, the records arrive by device-to-device copies): a worker that
        // already computes would otherwise hold the source context's lock for its whole sweep.
        {
            const auto t0 = std::chrono::steady_clock::now();
            struct Lap {
                std::chrono::steady_clock::time_point t0;
                int n;
                ~Lap()
                {
                    if (getenv("TVD_DEBUG_TIMING"))
                        fprintf(stderr, "[tvd_forward] prepare phase on %d devices %9.1f us\n", n,
                                std::chrono::duration<double, std::micro>(
                                    std::chrono::steady_clock::now() - t0).count());
                }
            } lap_prepare{t0, nparts};
            std::vector<std::thread> threads;
            for (int d = 0; d < nparts; d++)
                threads.emplace_back([&, d]() {
                    DeviceCtx *c = nullptr;
                    rcs[d] = get_ctx(m, device_ids ? device_ids[d] : d, &c);
                    if (rcs[d])
                        errs[d] = g_last_error;
                });
            for (auto &t : threads)
                t.join();
            for (int d = 0; d < nparts; d++)
                if (rcs[d]) {
                    g_last_error = errs[d];
                    return rcs[d];
                }
        }
        const bool timing = getenv("TVD_DEBUG_TIMING") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> threads;
        for (int d = 0; d < nparts; d++)
            threads.emplace_back(work, d);
        for (auto &t : threads)
            t.join();
        if (timing)
            fprintf(stderr, "[tvd_forward] compute phase on %d devices %9.1f us\n", nparts,
                    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0)
                        .count());
    }
    for (int d = 0; d < nparts; d++)
        if (rcs[d]) {
            g_last_error = errs[d];
            return rcs[d];
        }
    if (stats) {
        for (int i = 0; i < n_fields; i++) {
            int64_t *out = stats + (size_t)i * TVD_NSTATS;
            for (int d = 0; d < nparts; d++)
                for (int k = 0; k < TVD_NSTATS; k++)
                    out[k] += st[d][(size_t)order[i] * TVD_NSTATS + k];
            out[TVD_ST_PIECES] = m->n_pieces;
        }
    }
    return TVD_OK;
}

extern "C" int tvd_forward(int field, tvd_model *m, int64_t n_points, const double *lon,
                           const double *sinlat, const double *coslat, const double *radius,
                           double ratio, int stack_size, int n_devices, const int *device_ids,
                           double *result, int64_t *stats, int32_t *leaf_counts)
{
    return forward_host(1, &field, &ratio, m, n_points, lon, sinlat, coslat, radius, stack_size,
                        n_devices, device_ids, result, stats, leaf_counts);
}

extern "C" int tvd_forward_fused(int n_fields, const int *fields, const double *ratios,
                                 tvd_model *m, int64_t n_points, const double *lon,
                                 const double *sinlat, const double *coslat, const double *radius,
                                 int stack_size, int n_devices, const int *device_ids,
                                 double *results, int64_t *stats)
{
    return forward_host(n_fields, fields, ratios, m, n_points, lon, sinlat, coslat, radius,
                        stack_size, n_devices, device_ids, results, stats, nullptr);
}

extern "C" int tvd_sensitivity(int field, tvd_model *m, int64_t n_points, const double *lon,
                               const double *sinlat, const double *coslat, const double *radius,
                               double ratio, int stack_size, int device, double *jac,
                               int64_t *stats)
{
    g_last_error.clear();
    if (!m || n_points < 0 || (n_points > 0 && (!lon || !sinlat || !coslat || !radius || !jac)))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_sensitivity: bad argument");
    FieldSet fs;
    int order[1];
    int rc = make_field_set(1, &field, &ratio, fs, order);
    if (rc)
        return rc;
    if (device < 0 || device >= tvd_device_count())
        return set_error(TVD_ERR_NO_DEVICE, "tvd_sensitivity: no such CUDA device (no CPU fallback)");
    if (stats)
        memset(stats, 0, TVD_NSTATS * sizeof(int64_t));
    const int64_t P = n_points, T = m->n_tess;
    if (P == 0 || T == 0)
        return TVD_OK;
    DeviceGuard guard;
    DeviceCtx *c = nullptr;
    rc = get_ctx(m, device, &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(cudaSetDevice(device));
    // The per-piece and per-tesseroid rows of a slab of points live on the device
    // ((Q + T) * slab * 8 bytes): the points are processed in slabs that keep that under ~2 GiB.
    // A row's value at a point does not depend on the slab (one group, pieces summed in order).
    int64_t slab = (int64_t)((2ull << 30) / ((unsigned long long)(m->n_pieces + T) * sizeof(double)));
    slab = slab / 512 * 512;
    if (slab < 512)
        slab = 512;
    if (const char *dbg = getenv("TVD_DEBUG_JAC_SLAB"))     // tests force several slabs with this
        slab = atoll(dbg) > 0 ? atoll(dbg) : slab;
    if (slab > P)
        slab = P;
    cudaStream_t st = c->stream;
    CUDA_TRY(c->pts.reserve((size_t)slab * 4 * sizeof(double), st));
    CUDA_TRY(c->jac_out.reserve((size_t)T * slab * sizeof(double), st));
    double *d_pts = c->pts.as<double>();
    std::vector<int64_t> st_slab(TVD_NSTATS, 0);
    for (int64_t p0 = 0; p0 < P; p0 += slab) {
        const int64_t n = P - p0 < slab ? P - p0 : slab;
        const double *src[4] = {lon + p0, sinlat + p0, coslat + p0, radius + p0};
        for (int k = 0; k < 4; k++)
            CUDA_TRY(cudaMemcpyAsync(d_pts + (size_t)k * n, src[k], n * sizeof(double),
                                     cudaMemcpyHostToDevice, st));
        double *res[1] = {nullptr};
        rc = forward_core(m, c, fs, n, d_pts, d_pts + n, d_pts + 2 * n, d_pts + 3 * n, stack_size,
                          1, res, st_slab.data(), nullptr, st, RadiusInfo(),
                          c->jac_out.as<double>());
        if (rc)
            return rc;
        // rows of the slab [T][n] -> columns p0 .. p0 + n of the caller's [T][P]
        CUDA_TRY(cudaMemcpy2DAsync(jac + p0, (size_t)P * sizeof(double), c->jac_out.p,
                                   (size_t)n * sizeof(double), (size_t)n * sizeof(double),
                                   (size_t)T, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (stats)
            for (int k = 0; k < TVD_NSTATS; k++)
                stats[k] += st_slab[k];
    }
    if (stats)
        stats[TVD_ST_PIECES] = m->n_pieces;
    return TVD_OK;
}
