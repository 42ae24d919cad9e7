// tvd_abi.cu -- the C ABI of libtvd.so (include/tvd.h): model handles, per-device contexts,
// workspaces and the orchestration  K1 -> K4 -> Phase A -> sort -> Phase B -> Phase C.
#include <atomic>
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

// grow-only device buffer
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess)
            cap = bytes;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
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
    Buf test_rec, glq_rec, geom, tile_rec, ab_rec;
    bool ab_packed = false;
    double ab_radius = 0.0;
    Buf far_out, queue, keys_sorted, pair_result, sort_temp, counters;
    Buf pts, result, counts, jac_pieces, jac_out;
    unsigned long long qcap_hint = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_near[2 * TVD_MAX_FUSED] = {nullptr}, ev_end = nullptr;
    std::mutex mu;
};

struct tvd_model {
    int64_t n_tess = 0, n_pieces = 0;
    double delta = 0.0;
    int has_delta = 0;
    std::vector<double> h_bounds, h_params, h_top, h_bottom;
    std::vector<int32_t> h_kind, h_npieces, h_parent;
    std::map<int, DeviceCtx *> ctx;
    std::mutex mu;
};

static void free_table(PieceTable &t)
{
    cudaFree(t.tess_bounds);
    cudaFree(t.tess_kind);
    cudaFree(t.tess_params);
    cudaFree(t.tess_npieces);
    cudaFree(t.tess_offset);
    cudaFree(t.sweep_piece);
    cudaFree(t.top);
    cudaFree(t.bottom);
    cudaFree(t.parent);
    t = PieceTable();
}

static void free_ctx(DeviceCtx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    free_table(c->tab);
    c->test_rec.release();
    c->glq_rec.release();
    c->geom.release();
    c->tile_rec.release();
    c->ab_rec.release();
    c->far_out.release();
    c->queue.release();
    c->keys_sorted.release();
    c->pair_result.release();
    c->sort_temp.release();
    c->counters.release();
    c->pts.release();
    c->result.release();
    c->counts.release();
    c->jac_pieces.release();
    c->jac_out.release();
    if (c->ev0)
        cudaEventDestroy(c->ev0);
    if (c->ev1)
        cudaEventDestroy(c->ev1);
    for (cudaEvent_t ev : c->ev_near)
        if (ev)
            cudaEventDestroy(ev);
    if (c->ev_end)
        cudaEventDestroy(c->ev_end);
    if (c->stream)
        cudaStreamDestroy(c->stream);
    delete c;
}

static int upload_tess(tvd_model *m, DeviceCtx *c)
{
    PieceTable &t = c->tab;
    t.n_tess = m->n_tess;
    const size_t nt = (size_t)(m->n_tess > 0 ? m->n_tess : 1);
    CUDA_TRY(cudaMalloc(&t.tess_bounds, nt * 6 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.tess_kind, nt * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&t.tess_params, nt * TVD_NPAR * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.tess_npieces, nt * sizeof(int32_t)));
    if (m->n_tess > 0) {
        CUDA_TRY(cudaMemcpy(t.tess_bounds, m->h_bounds.data(), m->n_tess * 6 * sizeof(double),
                            cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(t.tess_kind, m->h_kind.data(), m->n_tess * sizeof(int32_t),
                            cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(t.tess_params, m->h_params.data(),
                            m->n_tess * TVD_NPAR * sizeof(double), cudaMemcpyHostToDevice));
    }
    return TVD_OK;
}

static int alloc_pieces(PieceTable &t, int64_t n_pieces)
{
    t.n_pieces = n_pieces;
    const size_t np = (size_t)(n_pieces > 0 ? n_pieces : 1);
    CUDA_TRY(cudaMalloc(&t.top, np * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.bottom, np * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.parent, np * sizeof(int32_t)));
    return TVD_OK;
}

static int new_ctx(int device, DeviceCtx **out)
{
    CUDA_TRY(cudaSetDevice(device));
    DeviceCtx *c = new DeviceCtx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess)
        e = cudaEventCreate(&c->ev1);
    for (int i = 0; i < 2 * TVD_MAX_FUSED && e == cudaSuccess; i++)
        e = cudaEventCreate(&c->ev_near[i]);
    if (e == cudaSuccess)
        e = cudaEventCreate(&c->ev_end);
    if (e != cudaSuccess) {
        free_ctx(c);
        CUDA_TRY(e);
    }
    *out = c;
    return TVD_OK;
}

// context of `device`, replicating the piece table from the host copy if needed
static int get_ctx(tvd_model *m, int device, DeviceCtx **out)
{
    {
        std::lock_guard<std::mutex> lock(m->mu);
        auto it = m->ctx.find(device);
        if (it != m->ctx.end()) {
            *out = it->second;
            return TVD_OK;
        }
    }
    // replicate outside the lock so that several devices are populated concurrently
    DeviceCtx *c = nullptr;
    int rc = new_ctx(device, &c);
    if (rc)
        return rc;
    rc = upload_tess(m, c);
    if (!rc)
        rc = alloc_pieces(c->tab, m->n_pieces);
    if (!rc && m->n_pieces > 0) {
        cudaError_t e = cudaMemcpy(c->tab.top, m->h_top.data(), m->n_pieces * sizeof(double),
                                   cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy(c->tab.bottom, m->h_bottom.data(), m->n_pieces * sizeof(double),
                           cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy(c->tab.parent, m->h_parent.data(), m->n_pieces * sizeof(int32_t),
                           cudaMemcpyHostToDevice);
        if (e == cudaSuccess && m->n_tess > 0)
            e = cudaMemcpy(c->tab.tess_npieces, m->h_npieces.data(),
                           m->n_tess * sizeof(int32_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            std::vector<int64_t> off((size_t)m->n_tess + 1, 0);
            for (int64_t i = 0; i < m->n_tess; i++)
                off[i + 1] = off[i] + m->h_npieces[i];
            e = cudaMalloc(&c->tab.tess_offset, off.size() * sizeof(int64_t));
            if (e == cudaSuccess)
                e = cudaMemcpy(c->tab.tess_offset, off.data(), off.size() * sizeof(int64_t),
                               cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) {
            g_last_error = std::string("piece table replication failed: ") + cudaGetErrorString(e);
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
    if (n < 0 || func < 0 || func > 6 || (n > 0 && (!x || !out)))
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_debug_math: bad argument");
    if (tvd_device_count() <= device || device < 0)
        return set_error(TVD_ERR_NO_DEVICE, "tvd_debug_math: no such CUDA device");
    if (n == 0)
        return TVD_OK;
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

    tvd_model *m = new tvd_model();
    m->n_tess = n_tess;
    m->has_delta = std::isnan(delta) ? 0 : 1;
    m->delta = m->has_delta ? delta : 0.0;
    m->h_bounds.assign(bounds, bounds + 6 * n_tess);
    m->h_kind.assign(kind, kind + n_tess);
    m->h_params.assign(params, params + TVD_NPAR * n_tess);

    DeviceCtx *c = nullptr;
    int rc = new_ctx(device, &c);
    if (rc) {
        delete m;
        return rc;
    }
    auto fail = [&](int code) {
        free_ctx(c);
        delete m;
        return code;
    };
    rc = upload_tess(m, c);
    if (rc)
        return fail(rc);
    int *d_status = nullptr;
    int64_t *d_offsets = nullptr;
    auto cleanup_tmp = [&]() {
        cudaFree(d_status);
        if (c->tab.tess_offset != d_offsets)
            cudaFree(d_offsets);
    };
#define MODEL_TRY(expr)                                                                  \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            g_last_error = std::string(#expr " failed: ") + cudaGetErrorString(_e);      \
            cudaGetLastError();                                                          \
            cleanup_tmp();                                                               \
            return fail((int)_e);                                                        \
        }                                                                                \
    } while (0)
    MODEL_TRY(cudaMalloc(&d_status, sizeof(int)));
    MODEL_TRY(cudaMalloc(&d_offsets, (size_t)(n_tess + 1) * sizeof(int64_t)));
    MODEL_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), c->stream));
    MODEL_TRY(tvd_launch_k1_count(c->tab, m->delta, m->has_delta, d_status, c->stream));
    MODEL_TRY(tvd_launch_scan(c->tab.tess_npieces, d_offsets, n_tess, c->stream));
    int64_t total = 0;
    int status = 0;
    MODEL_TRY(cudaMemcpyAsync(&total, d_offsets + n_tess, sizeof(int64_t), cudaMemcpyDeviceToHost,
                              c->stream));
    MODEL_TRY(cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MODEL_TRY(cudaStreamSynchronize(c->stream));
    if (status != 0) {
        cleanup_tmp();
        set_error(status, "tvd_model_create: a tesseroid needs more than 256 radial pieces "
                          "(delta too small for this density)");
        return fail(status);
    }
    if (total > 0x7fffffffLL) {
        cleanup_tmp();
        set_error(TVD_ERR_TOO_MANY_PIECES, "tvd_model_create: more than 2^31-1 radial pieces");
        return fail(TVD_ERR_TOO_MANY_PIECES);
    }
    m->n_pieces = total;
    rc = alloc_pieces(c->tab, total);
    if (rc) {
        cleanup_tmp();
        return fail(rc);
    }
    MODEL_TRY(tvd_launch_k1_emit(c->tab, d_offsets, m->delta, m->has_delta, c->stream));
    c->tab.tess_offset = d_offsets;     // owned by the piece table from here on
    m->h_top.resize(total);
    m->h_bottom.resize(total);
    m->h_parent.resize(total);
    m->h_npieces.resize(n_tess);
    if (total > 0) {
        MODEL_TRY(cudaMemcpyAsync(m->h_top.data(), c->tab.top, total * sizeof(double),
                                  cudaMemcpyDeviceToHost, c->stream));
        MODEL_TRY(cudaMemcpyAsync(m->h_bottom.data(), c->tab.bottom, total * sizeof(double),
                                  cudaMemcpyDeviceToHost, c->stream));
        MODEL_TRY(cudaMemcpyAsync(m->h_parent.data(), c->tab.parent, total * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, c->stream));
        MODEL_TRY(cudaMemcpyAsync(m->h_npieces.data(), c->tab.tess_npieces,
                                  n_tess * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    MODEL_TRY(cudaStreamSynchronize(c->stream));
#undef MODEL_TRY
    cleanup_tmp();
    c->has_table = true;
    m->ctx[device] = c;
    *out = m;
    return TVD_OK;
}

extern "C" void tvd_model_free(tvd_model *m)
{
    if (!m)
        return;
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &kv : m->ctx)
        free_ctx(kv.second);
    cudaSetDevice(cur);
    delete m;
}

extern "C" int64_t tvd_model_num_pieces(const tvd_model *m) { return m ? m->n_pieces : 0; }

extern "C" int tvd_model_get_pieces(const tvd_model *m, double *top, double *bottom,
                                    int32_t *parent)
{
    if (!m)
        return set_error(TVD_ERR_BAD_ARGUMENT, "tvd_model_get_pieces: null model");
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
static int plan_groups(int64_t n_pieces, int64_t n_points_total)
{
    // Pure function of (Q, P_total): every rank / device of a job derives the same plan.
    //  - few points: split the piece table into enough groups for ~2 waves of blocks on 148 SMs,
    //    but never groups smaller than 4 tiles;
    //  - many points: one group, unless the last wave of blocks would leave most SMs idle; then
    //    the smallest number of groups (<= 8) that fills the last wave to >= 97 %.
    const int64_t ppb = (int64_t)TVD_FAR_THREADS;
    const int64_t blocks_x = (n_points_total + ppb - 1) / ppb;
    if (blocks_x <= 0 || n_pieces <= 0)
        return 1;
    const int64_t slots = 148 * TVD_FAR_MINBLOCKS;      // blocks resident at once
    int64_t gmax = n_pieces / (4 * TVD_TILE);
    if (gmax < 1)
        gmax = 1;
    if (gmax > 65535)
        gmax = 65535;
    int64_t g;
    if (blocks_x < 2 * slots) {
        g = (2 * slots + blocks_x - 1) / blocks_x;
    } else {
        g = 1;
        double best = 0.0;
        for (int64_t cand = 1; cand <= 8; cand++) {
            const int64_t blocks = blocks_x * cand;
            const int64_t waves = (blocks + slots - 1) / slots;
            const double fill = (double)blocks / (double)(waves * slots);
            if (fill > best + 1e-12) {
                best = fill;
                g = cand;
            }
            if (fill >= 0.97)
                break;
        }
    }
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

static int forward_core(tvd_model *m, DeviceCtx *c, const FieldSet &fs, int64_t P,
                        const double *d_lon, const double *d_sinlat, const double *d_coslat,
                        const double *d_radius, int stack_size, int n_groups,
                        double *const *d_results /*[nf]*/, int64_t *stats /*[nf][TVD_NSTATS]*/,
                        int32_t *d_leaf_counts, cudaStream_t stream,
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
    if (Q == 0) {
        for (int f = 0; f < nf; f++)
            CUDA_TRY(cudaMemsetAsync(d_results[f], 0, P * sizeof(double), stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        return TVD_OK;
    }
    if (n_groups <= 0)
        n_groups = plan_groups(Q, P);
    int64_t ppg = (Q + n_groups - 1) / n_groups;
    ppg = ((ppg + TVD_TILE - 1) / TVD_TILE) * TVD_TILE;
    // groups that would start beyond Q are dropped (they would add exact zeros)
    n_groups = (int)((Q + ppg - 1) / ppg);

    // K4: GLQ records once per model, test records per (field set, ratios)
    if (!c->glq_packed) {
        CUDA_TRY(tvd_build_sweep_order(c->tab, stream));
        CUDA_TRY(c->glq_rec.reserve((size_t)Q * TVD_GLQ_REC * sizeof(double)));
        CUDA_TRY(c->geom.reserve((size_t)Q * 5 * sizeof(double)));
        CUDA_TRY(tvd_launch_pack_glq(c->tab, c->glq_rec.as<double>(), c->geom.as<double>(), stream));
        c->glq_packed = true;
    }
    bool same_test = c->packed_nf == nf;
    for (int f = 0; f < nf && same_test; f++)
        same_test = c->packed_ratio[f] == fs.ratio[f];
    if (!same_test) {
        RatioList rl;
        rl.nf = nf;
        for (int f = 0; f < TVD_MAX_FUSED; f++)
            rl.ratio[f] = f < nf ? fs.ratio[f] : 0.0;
        CUDA_TRY(c->test_rec.reserve((size_t)Q * tvd_test_rec_len(nf) * sizeof(double)));
        CUDA_TRY(c->tile_rec.reserve((size_t)((Q + TVD_TILE - 1) / TVD_TILE) * TVD_TILE_REC *
                                     sizeof(double)));
        CUDA_TRY(tvd_launch_pack_test(c->tab, c->geom.as<double>(), rl, c->test_rec.as<double>(),
                                      c->tile_rec.as<double>(), stream));
        c->packed_nf = nf;
        for (int f = 0; f < nf; f++)
            c->packed_ratio[f] = fs.ratio[f];
    }
    const int n_cnt = 8 + 8 * TVD_MAX_FUSED;
    CUDA_TRY(c->counters.reserve(n_cnt * sizeof(unsigned long long)));
    // points at one height: 2r*rc and r^2 + rc^2 per piece instead of per pair
    const double *d_ab = nullptr;
    if (!d_jac_out) {
        int *d_flag = reinterpret_cast<int *>(c->counters.as<unsigned long long>() + 7);
        int h_flag = 0;
        double h_r0 = 0.0;
        CUDA_TRY(tvd_launch_check_uniform(P, d_radius, d_flag, stream));
        CUDA_TRY(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&h_r0, d_radius, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (h_flag == 1 && h_r0 == h_r0) {
            if (!c->ab_packed || memcmp(&c->ab_radius, &h_r0, sizeof(double)) != 0) {
                CUDA_TRY(c->ab_rec.reserve((size_t)Q * TVD_AB_REC * sizeof(double)));
                CUDA_TRY(tvd_launch_pack_ab(c->tab, c->glq_rec.as<double>(), h_r0,
                                            c->ab_rec.as<double>(), stream));
                c->ab_packed = true;
                c->ab_radius = h_r0;
            }
            d_ab = c->ab_rec.as<double>();
        }
    }
    CUDA_TRY(c->far_out.reserve((size_t)nf * n_groups * P * sizeof(double)));
    if (d_jac_out)
        CUDA_TRY(c->jac_pieces.reserve((size_t)Q * P * sizeof(double)));
    unsigned long long qcap = c->qcap_hint;
    if (qcap == 0) {
        qcap = 64ull * (unsigned long long)P;
        if (qcap < (1ull << 20))
            qcap = 1ull << 20;
    }
    const unsigned long long all_pairs = (unsigned long long)P * (unsigned long long)Q;
    if (qcap > all_pairs)
        qcap = all_pairs;
    if (const char *dbg = getenv("TVD_DEBUG_QCAP"))     // tests force the overflow path with this
        qcap = strtoull(dbg, nullptr, 10) > 0 ? strtoull(dbg, nullptr, 10) : qcap;

    // [0..nf) = queue counts, [8 + 8 f ..) = near-field counters of field slot f
    unsigned long long *d_cnt = c->counters.as<unsigned long long>();
    unsigned long long nq[TVD_MAX_FUSED] = {0};
    for (int attempt = 0; attempt < 2; attempt++) {
        CUDA_TRY(c->queue.reserve((size_t)nf * qcap * sizeof(unsigned long long)));
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
        fa.n_groups = n_groups;
        fa.pieces_per_group = ppg;
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
        CUDA_TRY(cudaMemcpyAsync(nq, d_cnt, nf * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
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
    for (int f = 0; f < nf; f++) {
        const unsigned long long *d_keys = nullptr;
        const unsigned long long n = nq[f];
        if (n > 0) {
            const unsigned long long *q_in = c->queue.as<unsigned long long>() + (size_t)f * qcap;
            CUDA_TRY(c->keys_sorted.reserve((size_t)n * sizeof(unsigned long long)));
            CUDA_TRY(c->pair_result.reserve((size_t)n * sizeof(double)));
            size_t temp_bytes = 0;
            CUDA_TRY(tvd_sort_keys(nullptr, temp_bytes, q_in,
                                   c->keys_sorted.as<unsigned long long>(), (int64_t)n, 32 + pbits,
                                   stream));
            CUDA_TRY(c->sort_temp.reserve(temp_bytes ? temp_bytes : 16));
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
    unsigned long long h_cnt[8 + 8 * TVD_MAX_FUSED];
    CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(c->ev_end, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
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
    DeviceCtx *c = nullptr;
    rc = get_ctx(m, device, &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(cudaSetDevice(device));
    double *res[1] = {d_result};
    return forward_core(m, c, fs, n_points, d_lon, d_sinlat, d_coslat, d_radius, stack_size,
                        n_groups, res, stats, nullptr, (cudaStream_t)stream);
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
                      res, st.data(), nullptr, (cudaStream_t)stream);
    if (!rc && stats)
        for (int i = 0; i < n_fields; i++)
            memcpy(stats + (size_t)i * TVD_NSTATS, st.data() + (size_t)order[i] * TVD_NSTATS,
                   TVD_NSTATS * sizeof(int64_t));
    return rc;
}

// one device's share of tvd_forward: host slices in, host slices out
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
    if (stats)
        memset(stats, 0, (size_t)n_fields * TVD_NSTATS * sizeof(int64_t));
    if (P == 0)
        return TVD_OK;
    CUDA_TRY(c->pts.reserve((size_t)P * 4 * sizeof(double)));
    CUDA_TRY(c->result.reserve((size_t)n_fields * P * sizeof(double)));
    double *d_pts = c->pts.as<double>();
    const double *src[4] = {lon + p0, sinlat + p0, coslat + p0, radius + p0};
    for (int k = 0; k < 4; k++)
        CUDA_TRY(cudaMemcpyAsync(d_pts + (size_t)k * P, src[k], P * sizeof(double),
                                 cudaMemcpyHostToDevice, c->stream));
    int32_t *d_counts = nullptr;
    if (leaf_counts) {
        CUDA_TRY(c->counts.reserve((size_t)m->n_tess * P * sizeof(int32_t) + 16));
        d_counts = c->counts.as<int32_t>();
        CUDA_TRY(tvd_launch_init_counts(c->tab, P, d_counts, c->stream));
    }
    double *d_res[TVD_MAX_FUSED];
    for (int f = 0; f < n_fields; f++)
        d_res[f] = c->result.as<double>() + (size_t)f * P;
    rc = forward_core(m, c, fs, P, d_pts, d_pts + P, d_pts + 2 * P, d_pts + 3 * P, stack_size,
                      n_groups, d_res, stats, d_counts, c->stream);
    if (rc)
        return rc;
    for (int i = 0; i < n_fields; i++)
        CUDA_TRY(cudaMemcpyAsync(results + (size_t)i * P_total + p0, d_res[order[i]],
                                 P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (leaf_counts && m->n_tess > 0)
        CUDA_TRY(cudaMemcpy2DAsync(leaf_counts + p0, (size_t)P_total * sizeof(int32_t), d_counts,
                                   (size_t)P * sizeof(int32_t), (size_t)P * sizeof(int32_t),
                                   (size_t)m->n_tess, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
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
        std::vector<std::thread> threads;
        for (int d = 0; d < nparts; d++)
            threads.emplace_back(work, d);
        for (auto &t : threads)
            t.join();
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
    DeviceCtx *c = nullptr;
    rc = get_ctx(m, device, &c);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(c->pts.reserve((size_t)P * 4 * sizeof(double)));
    CUDA_TRY(c->jac_out.reserve((size_t)T * P * sizeof(double)));
    double *d_pts = c->pts.as<double>();
    const double *src[4] = {lon, sinlat, coslat, radius};
    for (int k = 0; k < 4; k++)
        CUDA_TRY(cudaMemcpyAsync(d_pts + (size_t)k * P, src[k], P * sizeof(double),
                                 cudaMemcpyHostToDevice, c->stream));
    double *res[1] = {nullptr};
    rc = forward_core(m, c, fs, P, d_pts, d_pts + P, d_pts + 2 * P, d_pts + 3 * P, stack_size, 1,
                      res, stats, nullptr, c->stream, c->jac_out.as<double>());
    if (rc)
        return rc;
    CUDA_TRY(cudaMemcpyAsync(jac, c->jac_out.p, (size_t)T * P * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return TVD_OK;
}
