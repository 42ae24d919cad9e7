// tvd_preprocess.cu -- point-independent preprocessing kernels (sm_100a).
//
// K1  density-based radial discretisation, one warp per tesseroid
//     (reference: _density_based_discretization / _divider_calculation,
//      code/tesseroid_density/tesseroid.py:237-300; NumPy semantics restated literally)
// --  sweep order: tesseroids stably sorted by longitude extent (tvd_build_sweep_order)
// K4  per-piece root records for the far-field sweep, in sweep order: the point-independent
//     halves of distance_n_size (_tesseroid.pyx:104-123) and scale_nodes (:152-176), the density
//     at the radial GLQ nodes (:274), the node weights, the far-test records per ratio set, the
//     bounding spheres of the tiles and the per-piece constants of the one-height specialisation.
#include <cstdlib>
#include "tvd_internal.cuh"

// ---------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------
// numpy.linspace(start, stop, 101)[i]  (tesseroid.py:258,293): i*step + start, last = stop;
// a zero step takes NumPy's denormal branch (i/100)*delta + start.
__device__ __forceinline__ double linspace101(double start, double stop, double delta,
                                              double step, int i)
{
    if (i == 100)
        return stop;
    if (step == 0.0)
        return ((double)i / 100.0) * delta + start;
    return (double)i * step + start;
}

// numpy.isclose(a, b), rtol=1e-5, atol=1e-8 (tesseroid.py:262)
__device__ __forceinline__ bool isclose_default(double a, double b)
{
    if (isfinite(a) && isfinite(b))
        return fabs(a - b) <= (1e-8 + 1e-5 * fabs(b));
    return a == b;
}

// One WARP per tesseroid.  The 101 density samples of an interval are spread over the lanes
// (sample i on lane i mod 32), minima / maxima / the first maximum of |rho_n - line| are warp
// reductions with NumPy's tie and NaN rules, and the FIFO of pending intervals
// (tesseroid.py:267-282) is a per-warp array in shared memory that is appended to and never
// wraps (a tesseroid of k pieces creates 2k - 1 intervals).
#define K1_WARPS 4
#define K1_MAX_IV (2 * TVD_MAX_K1_PIECES)
#define K1_KEEP TVD_K1_KEEP     // pieces per tesseroid the count pass keeps for the emit pass
#define K1_FULL 0xffffffffu

struct K1Best {                 // running first-maximum (numpy.argmax) with NaN propagation
    double best;                // largest diff so far (-1 before the first sample)
    int best_i;                 // its sample index (the smallest index among equals)
    int nan_i;                  // index of the first NaN sample, 1000 if none
};
__device__ __forceinline__ void k1_best_merge(K1Best &a, double best, int best_i, int nan_i)
{
    if (best > a.best || (best == a.best && best_i < a.best_i)) {
        a.best = best;
        a.best_i = best_i;
    }
    if (nan_i < a.nan_i)
        a.nan_i = nan_i;
}

// tesseroid.py:286-300, all lanes return the same divider and max_diff
__device__ __forceinline__ void divider_calculation(int lane, double top, double bottom, int kind,
                                                    const double *p, double rho_min,
                                                    double rho_max, double &divider,
                                                    double &max_diff)
{
    const double delta = top - bottom;
    const double step = delta / 100.0;
    const double range = rho_max - rho_min;
    double nd[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = lane + 32 * j;
        h[j] = linspace101(bottom, top, delta, step, i <= 100 ? i : 100);
        nd[j] = (tvd_density_cr(kind, p, h[j]) - rho_min) / range;
    }
    const double nd0 = __shfl_sync(K1_FULL, nd[0], 0);       // heights[0] == bottom exactly
    const double nd100 = __shfl_sync(K1_FULL, nd[3], 4);     // sample 100 = lane 4, slot 3
    const double slope = (nd100 - nd0) / (top - bottom);
    K1Best b = {-1.0, 1000, 1000};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = lane + 32 * j;
        if (i > 100)
            continue;
        const double line = slope * (h[j] - bottom) + nd0;
        const double diff = fabs(nd[j] - line);
        if (diff != diff) {                 // numpy max/argmax propagate the first NaN
            if (i < b.nan_i)
                b.nan_i = i;
        } else if (diff > b.best) {
            b.best = diff;
            b.best_i = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(K1_FULL, b.best, o);
        const int oi = __shfl_xor_sync(K1_FULL, b.best_i, o);
        const int on = __shfl_xor_sync(K1_FULL, b.nan_i, o);
        k1_best_merge(b, ob, oi, on);
    }
    if (b.nan_i <= 100) {
        max_diff = nan("");
        divider = linspace101(bottom, top, delta, step, b.nan_i);
    } else {
        max_diff = b.best;
        divider = linspace101(bottom, top, delta, step, b.best_i);
    }
}

// EMIT = false: counts[t] = number of pieces, and the pieces themselves into `keep` when there
// are at most K1_KEEP of them; EMIT = true: pieces written at offsets[t] (copied from `keep`
// when the count pass stored them, recomputed otherwise).
template <bool EMIT>
__global__ void __launch_bounds__(32 * K1_WARPS)
k1_discretize(int64_t n_tess, const double *__restrict__ bounds, const int32_t *__restrict__ kinds,
              const double *__restrict__ params, double delta, int has_delta,
              int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
              double *__restrict__ out_top, double *__restrict__ out_bottom,
              int32_t *__restrict__ out_parent, int *__restrict__ status,
              double *__restrict__ keep)
{
    __shared__ double s_top[K1_WARPS][K1_MAX_IV], s_bot[K1_WARPS][K1_MAX_IV];
    __shared__ unsigned char s_emit[K1_WARPS][K1_MAX_IV];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * K1_WARPS + wib;
    if (t >= n_tess)
        return;
    const double top = bounds[6 * t + 4], bottom = bounds[6 * t + 5];
    int64_t base = 0;
    if (EMIT) {
        base = offsets[t];
        const int cnt = counts[t];
        if (cnt <= K1_KEEP) {               // the count pass kept them
            if (lane < cnt) {
                out_top[base + lane] = keep[(t * K1_KEEP + lane) * 2 + 0];
                out_bottom[base + lane] = keep[(t * K1_KEEP + lane) * 2 + 1];
                out_parent[base + lane] = (int32_t)t;
            }
            return;
        }
    }
    const int kind = kinds[t];
    double p[TVD_NPAR];
#pragma unroll
    for (int i = 0; i < TVD_NPAR; i++)
        p[i] = params[TVD_NPAR * t + i];
    bool single = (kind == TVD_DENS_CONST) || !has_delta;   // tesseroid.py:222
    double rho_min = 0, rho_max = 0;
    if (!single) {
        // tesseroid.py:258-264
        const double d = top - bottom, step = d / 100.0;
        bool has_nan = false;
        double lo = __longlong_as_double(0x7ff0000000000000ll), hi = -lo;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = lane + 32 * j;
            if (i > 100)
                continue;
            const double rho = tvd_density_cr(kind, p, linspace101(bottom, top, d, step, i));
            if (rho != rho)
                has_nan = true;
            if (rho < lo)
                lo = rho;
            if (rho > hi)
                hi = rho;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double olo = __shfl_xor_sync(K1_FULL, lo, o), ohi = __shfl_xor_sync(K1_FULL, hi, o);
            if (olo < lo)
                lo = olo;
            if (ohi > hi)
                hi = ohi;
        }
        has_nan = __any_sync(K1_FULL, has_nan);
        rho_min = lo;
        rho_max = hi;
        // every sample NaN leaves +-inf here; the reference then has min = max = NaN as well
        if (has_nan)
            rho_min = rho_max = nan("");
        single = isclose_default(rho_min, rho_max);
    }
    if (single) {
        if (lane == 0) {
            if (EMIT) {
                out_top[base] = top;
                out_bottom[base] = bottom;
                out_parent[base] = (int32_t)t;
            } else {
                counts[t] = 1;
                keep[(t * K1_KEEP) * 2 + 0] = top;
                keep[(t * K1_KEEP) * 2 + 1] = bottom;
            }
        }
        return;
    }
    // FIFO of pending intervals (tesseroid.py:267-282)
    double *iv_top = s_top[wib], *iv_bot = s_bot[wib];
    unsigned char *iv_emit = s_emit[wib];
    const double tesseroid_size = top - bottom;
    int n_iv = 1, nout = 0;
    if (lane == 0) {
        iv_top[0] = top;
        iv_bot[0] = bottom;
    }
    __syncwarp();
    for (int head = 0; head < n_iv; head++) {
        const double it = iv_top[head], ib = iv_bot[head];
        double divider, max_diff;
        divider_calculation(lane, it, ib, kind, p, rho_min, rho_max, divider, max_diff);
        const double size_ratio = (it - ib) / tesseroid_size;
        if (max_diff * size_ratio > delta) {
            const int npending = n_iv - head - 1;
            if (npending + 2 + nout > TVD_MAX_K1_PIECES) {
                if (lane == 0) {
                    atomicExch(status, TVD_ERR_TOO_MANY_PIECES);
                    if (!EMIT)
                        counts[t] = 0;
                }
                return;
            }
            if (lane == 0) {
                iv_top[n_iv] = it;
                iv_bot[n_iv] = divider;
                iv_top[n_iv + 1] = divider;
                iv_bot[n_iv + 1] = ib;
                iv_emit[head] = 0;
            }
            n_iv += 2;
        } else {
            if (lane == 0)
                iv_emit[head] = 1;
            nout++;
        }
        __syncwarp();
    }
    if (!EMIT && lane == 0)
        counts[t] = nout;
    if (EMIT || nout <= K1_KEEP) {
        // emitted intervals in processing order = the reference's emission order
        int rank0 = 0;
        for (int i0 = 0; i0 < n_iv; i0 += 32) {
            const int i = i0 + lane;
            const bool em = i < n_iv && iv_emit[i] != 0;
            const unsigned m = __ballot_sync(K1_FULL, em);
            if (em) {
                const int r = rank0 + __popc(m & ((1u << lane) - 1u));
                if (EMIT) {
                    out_top[base + r] = iv_top[i];
                    out_bottom[base + r] = iv_bot[i];
                    out_parent[base + r] = (int32_t)t;
                } else {
                    keep[(t * K1_KEEP + r) * 2 + 0] = iv_top[i];
                    keep[(t * K1_KEEP + r) * 2 + 1] = iv_bot[i];
                }
            }
            rank0 += __popc(m);
        }
    }
}

cudaError_t tvd_launch_k1_count(const PieceTable &t, double delta, int has_delta, int *d_status,
                                double *d_keep, cudaStream_t stream)
{
    if (t.n_tess == 0)
        return cudaSuccess;
    const unsigned blocks = (unsigned)((t.n_tess + K1_WARPS - 1) / K1_WARPS);
    k1_discretize<false><<<blocks, 32 * K1_WARPS, 0, stream>>>(
        t.n_tess, t.tess_bounds, t.tess_kind, t.tess_params, delta, has_delta, t.tess_npieces,
        nullptr, nullptr, nullptr, nullptr, d_status, d_keep);
    tvd_count_launch();
    return cudaGetLastError();
}

cudaError_t tvd_launch_k1_emit(const PieceTable &t, const int64_t *d_offsets, double delta,
                               int has_delta, double *d_keep, cudaStream_t stream)
{
    if (t.n_tess == 0)
        return cudaSuccess;
    const unsigned blocks = (unsigned)((t.n_tess + K1_WARPS - 1) / K1_WARPS);
    k1_discretize<true><<<blocks, 32 * K1_WARPS, 0, stream>>>(
        t.n_tess, t.tess_bounds, t.tess_kind, t.tess_params, delta, has_delta, t.tess_npieces,
        d_offsets, t.top, t.bottom, t.parent, nullptr, d_keep);
    tvd_count_launch();
    return cudaGetLastError();
}

// Exclusive scan of int32 counts into int64 offsets (offsets[n] = total).  One block walks the
// array in 1024-element strips carrying the running total: O(n) and launch-latency bound,
// called once per model.
__global__ void __launch_bounds__(1024)
scan_counts(const int32_t *__restrict__ counts, int64_t *__restrict__ offsets, int64_t n)
{
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (int64_t start = 0; start < n; start += 1024) {
        const int64_t i = start + threadIdx.x;
        const int64_t v = i < n ? (int64_t)counts[i] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o)
                x += y;
        }
        if (lane == 31)
            warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o)
                    w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t incl = x + (warp > 0 ? warp_sums[warp - 1] : 0) + carry;
        if (i < n)
            offsets[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023)
            carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        offsets[n] = carry_s;
}

cudaError_t tvd_launch_scan(const int32_t *d_counts, int64_t *d_offsets, int64_t n,
                            cudaStream_t stream)
{
    scan_counts<<<1, 1024, 0, stream>>>(d_counts, d_offsets, n);
    tvd_count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// sweep order: tesseroids stably sorted by the bit patterns of (w, e); any total order works,
// the point is that tesseroids with the same extent in longitude (a column of a regular mesh)
// become neighbours, so the sweep computes cos(lon - lonc) once per column
// ---------------------------------------------------------------------------------------------
__global__ void sweep_keys(int64_t n, const double *__restrict__ bounds, int which,
                           const int32_t *__restrict__ idx_in, unsigned long long *__restrict__ keys,
                           int32_t *__restrict__ idx_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int64_t t = idx_in ? idx_in[i] : i;
    keys[i] = (unsigned long long)__double_as_longlong(bounds[6 * t + which]);
    if (idx_out)
        idx_out[i] = (int32_t)t;
}
__global__ void sweep_counts(int64_t n, const int32_t *__restrict__ order,
                             const int32_t *__restrict__ npieces, int32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = npieces[order[i]];
}
__global__ void sweep_expand(int64_t n, const int32_t *__restrict__ order,
                             const int64_t *__restrict__ tess_offset,
                             const int64_t *__restrict__ sweep_offset,
                             int32_t *__restrict__ sweep_piece)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int64_t t = order[i];
    const int64_t q0 = tess_offset[t], np = tess_offset[t + 1] - q0, o = sweep_offset[i];
    for (int64_t k = 0; k < np; k++)
        sweep_piece[o + k] = (int32_t)(q0 + k);
}

cudaError_t tvd_build_sweep_order(PieceTable &t, cudaStream_t stream)
{
    if (t.sweep_piece != nullptr || t.n_pieces == 0)
        return cudaSuccess;
    const int64_t n = t.n_tess;
    // stream-ordered temporaries (cudaMallocAsync pool): no device-wide synchronisation
    unsigned long long *key_a = nullptr, *key_b = nullptr;
    int32_t *idx_a = nullptr, *idx_b = nullptr, *cnt = nullptr;
    int64_t *off = nullptr;
    void *temp = nullptr;
    size_t temp_bytes = 0;
    cudaError_t e = cudaMallocAsync(&t.sweep_piece, (size_t)t.n_pieces * sizeof(int32_t), stream);
    auto done = [&](cudaError_t err) {
        cudaFreeAsync(key_a, stream);
        cudaFreeAsync(key_b, stream);
        cudaFreeAsync(idx_a, stream);
        cudaFreeAsync(idx_b, stream);
        cudaFreeAsync(cnt, stream);
        cudaFreeAsync(off, stream);
        cudaFreeAsync(temp, stream);
        if (err != cudaSuccess) {
            cudaFreeAsync(t.sweep_piece, stream);
            t.sweep_piece = nullptr;
        }
        return err;
    };
    if (e != cudaSuccess)
        return done(e);
#define SW_TRY(x)              \
    do {                       \
        e = (x);               \
        if (e != cudaSuccess)  \
            return done(e);    \
    } while (0)
    SW_TRY(cudaMallocAsync(&key_a, n * sizeof(unsigned long long), stream));
    SW_TRY(cudaMallocAsync(&key_b, n * sizeof(unsigned long long), stream));
    SW_TRY(cudaMallocAsync(&idx_a, n * sizeof(int32_t), stream));
    SW_TRY(cudaMallocAsync(&idx_b, n * sizeof(int32_t), stream));
    SW_TRY(cudaMallocAsync(&cnt, n * sizeof(int32_t), stream));
    SW_TRY(cudaMallocAsync(&off, (n + 1) * sizeof(int64_t), stream));
    SW_TRY(tvd_sort_pairs(nullptr, temp_bytes, key_a, key_b, idx_a, idx_b, n, stream));
    SW_TRY(cudaMallocAsync(&temp, temp_bytes ? temp_bytes : 16, stream));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    // LSD: first by e, then (stably) by w
    sweep_keys<<<blocks, 256, 0, stream>>>(n, t.tess_bounds, 1, nullptr, key_a, idx_a);
    SW_TRY(tvd_sort_pairs(temp, temp_bytes, key_a, key_b, idx_a, idx_b, n, stream));
    sweep_keys<<<blocks, 256, 0, stream>>>(n, t.tess_bounds, 0, idx_b, key_a, nullptr);
    SW_TRY(tvd_sort_pairs(temp, temp_bytes, key_a, key_b, idx_b, idx_a, n, stream));
    sweep_counts<<<blocks, 256, 0, stream>>>(n, idx_a, t.tess_npieces, cnt);
    SW_TRY(tvd_launch_scan(cnt, off, n, stream));
    sweep_expand<<<blocks, 256, 0, stream>>>(n, idx_a, t.tess_offset, off, t.sweep_piece);
    tvd_count_launch(4);
    SW_TRY(cudaGetLastError());
#undef SW_TRY
    return done(cudaSuccess);
}

// ---------------------------------------------------------------------------------------------
// K4: root records
// ---------------------------------------------------------------------------------------------
// Records are stored in SWEEP order (PieceTable::sweep_piece).
// glq record  [16]: lonc0 lonc1 sinlatc0 sinlatc1 coslatc0 coslatc1 rc0 rc1 rc0^2 rc1^2
//                   w00 w01 w10 w11 (w[j][k] = rho_k * kappa_jk * scale)
//                   flags (int64)  model-order piece index (int64)
// geometry     [5]: cx cy cz (Cartesian centre of the piece)  rt^2  max(Llon, Llat)
// test record: cx cy cz K_0 .. K_{nf-1} with K_f = rt^2 - thr2(ratio_f); thr2 is the conservative
//   far-field threshold on the squared centre distance, see tvd_far.cu; pairs that do not clear
//   it are re-examined with the literal test.
#define TVD_GEOM_REC 5
__global__ void __launch_bounds__(128)
k4_pack_glq(int64_t n_pieces, const double *__restrict__ tess_bounds,
            const int32_t *__restrict__ tess_kind, const double *__restrict__ tess_params,
            const double *__restrict__ ptop, const double *__restrict__ pbottom,
            const int32_t *__restrict__ parent, const int32_t *__restrict__ sweep_piece,
            int period, double *__restrict__ glq_rec, double *__restrict__ geom)
{
    // one thread per sweep position qs; q = the piece (model order) visited there
    const int64_t qs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qs >= n_pieces)
        return;
    const int64_t q = sweep_piece[qs];
    const int64_t t = parent[q];
    const double w = tess_bounds[6 * t + 0], e = tess_bounds[6 * t + 1];
    const double s = tess_bounds[6 * t + 2], n = tess_bounds[6 * t + 3];
    const double top = ptop[q], bottom = pbottom[q];
    const int kind = tess_kind[t];
    double p[TVD_NPAR];
#pragma unroll
    for (int i = 0; i < TVD_NPAR; i++)
        p[i] = tess_params[TVD_NPAR * t + i];

    // distance_n_size, point-independent part (_tesseroid.pyx:111-122)
    const double rt = 0.5 * (top + bottom) + TVD_R;
    const double lont = TVD_D2R * 0.5 * (w + e);
    const double latt = TVD_D2R * 0.5 * (s + n);
    double sinlatt, coslatt, sinlont, coslont, sin_n, cos_n, sin_s, cos_s;
    cr_sincos(latt, sinlatt, coslatt);
    cr_sincos(lont, sinlont, coslont);
    cr_sincos(TVD_D2R * n, sin_n, cos_n);
    cr_sincos(TVD_D2R * s, sin_s, cos_s);
    const double rtop = top + TVD_R;
    const double arg_lon = sinlatt * sinlatt + (coslatt * coslatt) * cr_cos(TVD_D2R * (e - w));
    const double arg_lat = sin_n * sin_s + cos_n * cos_s;
    const double Llon = rtop * acos(arg_lon);
    const double Llat = rtop * acos(arg_lat);
    // The far test must hold for the sizes the REFERENCE computes.  L = rtop*acos(arg) with
    // L ~ sqrt(2 (1 - arg)): a difference of a few ulp in arg (libm vs libdevice sin/cos/acos)
    // moves L by ~4e-16 / (2 (1 - arg)) relative, which for very small cells exceeds the fixed
    // 2e-6 guard of the threshold (1 - arg = 1.5e-10 for a 0.001 degree cell).  The sizes that
    // enter the conservative test are therefore inflated by that conditioning (the near-field
    // walk, tvd_near.cu, repeats the literal test for everything this defers).
    const double ulon = 1.0e-15 / fmax(1.0 - arg_lon, 1.0e-16);
    const double ulat = 1.0e-15 / fmax(1.0 - arg_lat, 1.0e-16);
    double *gm = geom + TVD_GEOM_REC * qs;
    gm[0] = rt * (coslatt * coslont);          // centre of the piece, geocentric Cartesian
    gm[1] = rt * (coslatt * sinlont);
    gm[2] = rt * sinlatt;
    gm[3] = rt * rt;
    // fmax ignores a single NaN (a NaN size never splits)
    gm[4] = fmax(Llon * (1.0 + ulon), Llat * (1.0 + ulat));

    // scale_nodes (_tesseroid.pyx:162-175)
    const double dlon = e - w, dlat = n - s, dr = top - bottom;
    const double mlon = 0.5 * (e + w), mlat = 0.5 * (n + s);
    const double mr = 0.5 * (top + bottom + 2. * TVD_R);
    double lonc[2], sinlatc[2], coslatc[2], rc[2], rcsq[2], dens[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double node = i == 0 ? -TVD_NODE : TVD_NODE;
        lonc[i] = TVD_D2R * (0.5 * dlon * node + mlon);
        const double latc = TVD_D2R * (0.5 * dlat * node + mlat);
        cr_sincos(latc, sinlatc[i], coslatc[i]);
        rc[i] = (0.5 * dr * node + mr);
        rcsq[i] = rc[i] * rc[i];
        dens[i] = tvd_density_cr(kind, p, rc[i] - TVD_R);   // :274 (or the constant, :250)
    }
    const double scale = TVD_D2R * dlon * TVD_D2R * dlat * dr * 0.125;
    double *g = glq_rec + TVD_GLQ_REC * qs;
    g[0] = lonc[0];
    g[1] = lonc[1];
    g[2] = sinlatc[0];
    g[3] = sinlatc[1];
    g[4] = coslatc[0];
    g[5] = coslatc[1];
    g[6] = rc[0];
    g[7] = rc[1];
    g[8] = rcsq[0];
    g[9] = rcsq[1];
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double kappa = rcsq[k] * coslatc[j];          // :248
            g[10 + 2 * j + k] = dens[k] * kappa * scale;
        }
    }
    // flags relative to the previous record (reset at every multiple of `period`, where groups
    // and tiles of the sweep may start), two 32-bit words: [0] = same
    // tesseroid (a further radial piece) -> the sweep keeps its whole angular half; [1] = same
    // extent in longitude (bitwise equal w and e) -> it keeps the longitude part
    int same_tess = 0, same_lon = 0;
    if ((qs % period) != 0) {
        const int64_t tp = parent[sweep_piece[qs - 1]];
        if (tp == t) {
            same_tess = 1;
            same_lon = 1;
        } else if (__double_as_longlong(tess_bounds[6 * tp + 0]) == __double_as_longlong(w) &&
                   __double_as_longlong(tess_bounds[6 * tp + 1]) == __double_as_longlong(e)) {
            same_lon = 1;
        }
    }
    reinterpret_cast<int *>(g + 14)[0] = same_tess;
    reinterpret_cast<int *>(g + 14)[1] = same_lon;
    reinterpret_cast<long long *>(g)[15] = q;      // the piece's index in the model-order table
}

// Conservative far-field threshold.  The literal test (divisions, :135,140) splits iff
// d <= ratio*Llon or d <= ratio*Llat (NaN sizes never split).  The sweep evaluates
// d^2 = r^2 + rt^2 - 2 p.c from Cartesian position vectors with an absolute error far below
// 1 m^2, and the sizes L of root cells are reproduced to far better than 1e-6, so
// d2 > thr^2*(1+2e-6) + 2 guarantees that the literal test says "no split"; everything else is
// deferred to the literal walk.
__global__ void __launch_bounds__(128)
k4_pack_test(int64_t n_pieces, const double *__restrict__ geom, RatioList ratios,
             double *__restrict__ test_rec, bool no_far)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pieces)
        return;
    const int trec = tvd_test_rec_len(ratios.nf);
    const double *gm = geom + TVD_GEOM_REC * q;
    double *tr = test_rec + (int64_t)trec * q;
    tr[0] = gm[0];
    tr[1] = gm[1];
    tr[2] = gm[2];
    for (int f = 0; f < ratios.nf; f++) {
        const double thr = ratios.ratio[f] * gm[4];             // = max(ratio*Llon, ratio*Llat)
        double thr2 = thr * thr * (1.0 + 2e-6) + 2.0;           // NaN if both sizes are NaN
        if (no_far)
            thr2 = __longlong_as_double(0x7ff8000000000000ll);
        tr[3 + f] = gm[3] - thr2;       // a NaN threshold defers every pair of the piece
    }
    for (int f = 3 + ratios.nf; f < trec; f++)
        tr[f] = 0.0;
}

// Tile records.  For the pieces of one tile: C = mean of their centres, R = largest distance of
// a centre from C, thr_max_f = largest far-field threshold.  A point with
//   |p - C| > R + thr_max_f (1 + 2e-6) + 3 m
// satisfies d_pq >= |p - C| - R > thr_q (1 + 2e-6) + 3 for every piece q of the tile, hence
// d_pq^2 > thr_q^2 (1 + 2e-6) + 2 + (rounding slack): every per-pair far test of the tile passes.
__global__ void __launch_bounds__(128)
k4_pack_tiles(int64_t n_pieces, const double *__restrict__ geom, RatioList ratios,
              int64_t ppg, int tiles_per_group, int64_t n_tiles, double *__restrict__ tile_rec,
              bool no_far)
{
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles)
        return;
    const int64_t g = tile / tiles_per_group, k = tile - g * tiles_per_group;
    const int64_t q0 = g * ppg + k * TVD_TILE;
    int64_t q1 = q0 + TVD_TILE;
    if (q1 > (g + 1) * ppg)
        q1 = (g + 1) * ppg;
    if (q1 > n_pieces)
        q1 = n_pieces;
    double *tr = tile_rec + TVD_TILE_REC * tile;
    if (q0 >= q1) {                     // a tile slot beyond the end of its group: never read
        for (int f = 0; f < TVD_TILE_REC; f++)
            tr[f] = 0.0;
        return;
    }
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int64_t q = q0; q < q1; q++) {
        cx += geom[TVD_GEOM_REC * q + 0];
        cy += geom[TVD_GEOM_REC * q + 1];
        cz += geom[TVD_GEOM_REC * q + 2];
    }
    const double inv = 1.0 / (double)(q1 - q0);
    cx *= inv;
    cy *= inv;
    cz *= inv;
    double r2max = 0.0, lmax = 0.0;
    bool bad = no_far;
    for (int64_t q = q0; q < q1; q++) {
        const double dx = geom[TVD_GEOM_REC * q + 0] - cx, dy = geom[TVD_GEOM_REC * q + 1] - cy;
        const double dz = geom[TVD_GEOM_REC * q + 2] - cz;
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double L = geom[TVD_GEOM_REC * q + 4];
        if (!(r2 == r2) || !(L == L))
            bad = true;                     // NaN geometry: this tile always takes the per-pair path
        r2max = fmax(r2max, r2);
        lmax = fmax(lmax, L);
    }
    tr[0] = cx;
    tr[1] = cy;
    tr[2] = cz;
    tr[3] = cx * cx + cy * cy + cz * cz;
    const double R = sqrt(r2max) * (1.0 + 1e-9) + 1e-3;
    for (int f = 0; f < TVD_MAX_FUSED; f++) {
        double lim = __longlong_as_double(0x7ff0000000000000ll);        // +inf: never far
        if (f < ratios.nf && !bad) {
            const double x = R + ratios.ratio[f] * lmax * (1.0 + 2e-6) + 3.0;
            lim = x * x * (1.0 + 1e-12) + 1.0;
            if (!(lim == lim))
                lim = __longlong_as_double(0x7ff0000000000000ll);
        }
        tr[4 + f] = lim;
    }
}

cudaError_t tvd_launch_pack_glq(const PieceTable &t, int period, double *d_glq, double *d_geom,
                                cudaStream_t stream)
{
    if (t.n_pieces == 0)
        return cudaSuccess;
    const int threads = 128;
    const unsigned blocks = (unsigned)((t.n_pieces + threads - 1) / threads);
    k4_pack_glq<<<blocks, threads, 0, stream>>>(t.n_pieces, t.tess_bounds, t.tess_kind,
                                                t.tess_params, t.top, t.bottom, t.parent,
                                                t.sweep_piece, period, d_glq, d_geom);
    tvd_count_launch();
    return cudaGetLastError();
}

static bool tvd_debug_no_far()
{
    // test knob: every pair goes through the literal near-field walk (no far-field shortcut)
    const char *nf = getenv("TVD_DEBUG_NO_FAR");
    return nf && nf[0] == '1';
}

cudaError_t tvd_launch_pack_test(const PieceTable &t, const double *d_geom, RatioList ratios,
                                 double *d_test, cudaStream_t stream)
{
    if (t.n_pieces == 0)
        return cudaSuccess;
    const int threads = 128;
    const unsigned blocks = (unsigned)((t.n_pieces + threads - 1) / threads);
    k4_pack_test<<<blocks, threads, 0, stream>>>(t.n_pieces, d_geom, ratios, d_test,
                                                 tvd_debug_no_far());
    tvd_count_launch();
    return cudaGetLastError();
}

cudaError_t tvd_launch_pack_tiles(const PieceTable &t, const double *d_geom, RatioList ratios,
                                  int64_t pieces_per_group, int n_groups, double *d_tile,
                                  cudaStream_t stream)
{
    if (t.n_pieces == 0)
        return cudaSuccess;
    const int threads = 128;
    const int tiles_per_group = (int)((pieces_per_group + TVD_TILE - 1) / TVD_TILE);
    const int64_t n_tiles = (int64_t)n_groups * tiles_per_group;
    k4_pack_tiles<<<(unsigned)((n_tiles + threads - 1) / threads), threads, 0, stream>>>(
        t.n_pieces, d_geom, ratios, pieces_per_group, tiles_per_group, n_tiles, d_tile,
        tvd_debug_no_far());
    tvd_count_launch();
    return cudaGetLastError();
}

__global__ void init_counts(int64_t n_tess, int64_t n_points, const int32_t *__restrict__ npieces,
                            int32_t *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tess * n_points)
        counts[i] = npieces[i / n_points];
}

cudaError_t tvd_launch_init_counts(const PieceTable &t, int64_t n_points, int32_t *counts,
                                   cudaStream_t stream)
{
    const int64_t n = t.n_tess * n_points;
    if (n == 0)
        return cudaSuccess;
    init_counts<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(t.n_tess, n_points,
                                                                 t.tess_npieces, counts);
    tvd_count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// debug: the double-double sin / cos / exp, exposed so that tests can check them against an
// arbitrary-precision reference
// ---------------------------------------------------------------------------------------------
__global__ void debug_math_kernel(int func, int64_t n, const double *__restrict__ x,
                                  double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    double v;
    switch (func) {
    case 0: v = dd_sin(x[i]); break;
    case 1: v = dd_cos(x[i]); break;
    case 2: v = dd_exp(x[i]); break;
    case 3: v = cr_sin(x[i]); break;          // table-based correctly rounded (tvd_crtrig.cuh)
    case 4: v = cr_cos(x[i]); break;
    case 5: { double c; cr_sincos(x[i], v, c); break; }
    case 6: { double s; cr_sincos(x[i], s, v); break; }
    default: v = cr_exp(x[i]); break;
    }
    out[i] = v;
}

cudaError_t tvd_debug_math_launch(int func, int64_t n, const double *d_x, double *d_out,
                                  cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    debug_math_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(func, n, d_x, d_out);
    tvd_count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// points-at-one-height specialisation of the far-field sweep
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k4_pack_ab(int64_t n_pieces, const double *__restrict__ glq_rec, double radius,
           double *__restrict__ ab, double *__restrict__ wz)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pieces)
        return;
    const double *g = glq_rec + TVD_GLQ_REC * q;
    const double r_sqr = radius * radius;      // radius**2 (:239), as the sweep computes it
    const double r2 = 2.0 * radius;            // 2*radius  (:247)
    double *o = ab + TVD_AB_REC * q;
    o[0] = r2 * g[6];
    o[1] = r2 * g[7];
    o[2] = r_sqr + g[8];
    o[3] = r_sqr + g[9];
    // w[j][k]*rc[k] and w[j][k]*radius, the products the sweep forms for points at individual heights
    double *z = wz + TVD_WZ_REC * q;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            z[2 * j + k] = g[10 + 2 * j + k] * g[6 + k];
            z[4 + 2 * j + k] = g[10 + 2 * j + k] * radius;
        }
}

cudaError_t tvd_launch_pack_ab(const PieceTable &t, const double *d_glq, double radius,
                               double *d_ab, double *d_wz, cudaStream_t stream)
{
    if (t.n_pieces == 0)
        return cudaSuccess;
    k4_pack_ab<<<(unsigned)((t.n_pieces + 255) / 256), 256, 0, stream>>>(t.n_pieces, d_glq, radius,
                                                                        d_ab, d_wz);
    tvd_count_launch();
    return cudaGetLastError();
}

__global__ void check_uniform(int64_t n, const double *__restrict__ radius, int *__restrict__ flag)
{
    const long long first = __double_as_longlong(radius[0]);
    bool same = true;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        same = same && (__double_as_longlong(radius[i]) == first);
    if (!same)
        *flag = 0;
}

cudaError_t tvd_launch_check_uniform(int64_t n, const double *d_radius, int *d_flag,
                                     cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(int), stream);
    if (e != cudaSuccess || n == 0)
        return e;
    const int one = 1;
    e = cudaMemcpyAsync(d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess)
        return e;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1184)
        blocks = 1184;
    check_uniform<<<blocks, 256, 0, stream>>>(n, d_radius, d_flag);
    tvd_count_launch();
    return cudaGetLastError();
}
