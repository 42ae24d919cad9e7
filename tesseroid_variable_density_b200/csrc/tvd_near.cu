// tvd_near.cu -- Phase B (near-field pairs, literal adaptive subdivision) and Phase C (combine).
//
// Phase B restates rediscretizer (_tesseroid.pyx:34-98) for the (piece, point) pairs that the
// far-field sweep could not prove to be far: one thread walks one pair depth-first in the
// reference's own LIFO order with the literal distance_n_size / divisions / split /
// scale_nodes / kernel arithmetic, so leaf sets (and hence subdivision counts) and the
// per-pair summation order are the reference's.  The explicit stack holds one record per
// subdivision LEVEL (parent w, s, dlon, dlat and the next child to visit) instead of one row per
// pending node: children are regenerated with the reference's own expressions
// w + i*dlon, w + (i+1)*dlon (:192-195) when they are popped, and the reference's row count
// (`stktop`) is tracked separately to reproduce its overflow error (:88-89).
//
// Pairs arrive sorted by (point, piece): neighbouring threads walk similar trees, and Phase C
// adds a point's pair results in piece order, which makes every point's value independent of
// how the points were sharded over blocks, groups or GPUs.
#include "tvd_internal.cuh"

// Near-field pairs are the worst-conditioned ones (cos psi and l^2 cancel hardest for the closest
// cells), and there are few of them: the trigonometry of the integration nodes (scale_nodes and
// cos(lon - lonc)) is evaluated in double-double and rounded once (tvd_ddmath.cuh), i.e. it
// equals glibc's result wherever glibc rounds correctly.

struct Level {
    double w, s, dlon, dlat;
    int nlon, nlat, next;
};

struct NodeTrig {            // scale_nodes output (_tesseroid.pyx:152-176)
    double lonc[2], sinlatc[2], coslatc[2], rc[2], scale;
};

__device__ __forceinline__ void scale_nodes(double w, double e, double s, double n, double top,
                                            double bottom, NodeTrig &o)
{
    const double dlon = e - w, dlat = n - s, dr = top - bottom;
    const double mlon = 0.5 * (e + w), mlat = 0.5 * (n + s);
    const double mr = 0.5 * (top + bottom + 2. * TVD_R);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double node = i == 0 ? -TVD_NODE : TVD_NODE;
        o.lonc[i] = TVD_D2R * (0.5 * dlon * node + mlon);
        const double latc = TVD_D2R * (0.5 * dlat * node + mlat);
        dd_sincos(latc, o.sinlatc[i], o.coslatc[i]);
        o.rc[i] = (0.5 * dr * node + mr);
    }
    o.scale = TVD_D2R * dlon * TVD_D2R * dlat * dr * 0.125;
}

// kernel{V,x,y,z}[_variable] (_tesseroid.pyx:231-520) and the Fatiando 0.5 tensor integrands.
template <int FIELD>
__device__ double glq_literal(bool variable, int kind, const double *p, double lon, double sinlat,
                              double coslat, double radius, const NodeTrig &nd)
{
    const double r_sqr = radius * radius;
    double dens[2] = {1.0, 1.0};
    if (variable) {
        dens[0] = tvd_density(kind, p, nd.rc[0] - TVD_R);
        dens[1] = tvd_density(kind, p, nd.rc[1] - TVD_R);
    }
    double result = 0.0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        // sin(lonc - lon) = -sin(lon - lonc) exactly (odd function, same |argument|)   (:401)
        double coslon, sinlon;
        dd_sincos(lon - nd.lonc[i], sinlon, coslon);
        sinlon = -sinlon;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double kphi = coslat * nd.sinlatc[j] - sinlat * nd.coslatc[j] * coslon;
            const double cospsi = sinlat * nd.sinlatc[j] + coslat * nd.coslatc[j] * coslon;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const double rc = nd.rc[k];
                const double rc_sqr = rc * rc;
                const double l_sqr = r_sqr + rc_sqr - 2 * radius * rc * cospsi;
                const double kappa = rc_sqr * nd.coslatc[j];
                const double dk = variable ? dens[k] * kappa : kappa;
                const double l = sqrt(l_sqr);
                double term;
                if (FIELD == TVD_POTENTIAL) {
                    term = dk / l;
                } else if (FIELD == TVD_GX) {
                    term = dk * rc * kphi / (l_sqr * l);
                } else if (FIELD == TVD_GY) {
                    term = dk * (rc * nd.coslatc[j] * sinlon / (l_sqr * l));
                } else if (FIELD == TVD_GZ) {
                    term = dk * (rc * cospsi - radius) / (l_sqr * l);
                } else {
                    const double dx = rc * kphi;
                    const double dy = rc * nd.coslatc[j] * sinlon;
                    const double dz = rc * cospsi - radius;
                    const double l5 = l_sqr * l_sqr * l;
                    double num;
                    if (FIELD == TVD_GXX) num = 3 * (dx * dx) - l_sqr;
                    else if (FIELD == TVD_GXY) num = 3 * dx * dy;
                    else if (FIELD == TVD_GXZ) num = 3 * dx * dz;
                    else if (FIELD == TVD_GYY) num = 3 * (dy * dy) - l_sqr;
                    else if (FIELD == TVD_GYZ) num = 3 * dy * dz;
                    else num = 3 * (dz * dz) - l_sqr;
                    term = dk * num / l5;
                }
                result += term;
            }
        }
    }
    if (FIELD == TVD_GZ)
        result *= -1;
    if (variable)
        return result * nd.scale;
    return result * nd.scale * p[0];
}

template <int FIELD>
__global__ void __launch_bounds__(128)
near_walk(const NearArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_pairs)
        return;
    const unsigned long long key = a.keys[i];
    const int64_t pi = (int64_t)(key >> 32);
    const int64_t q = (int64_t)(key & 0xffffffffull);
    const double lon = a.lon[pi], sinlat = a.sinlat[pi], coslat = a.coslat[pi], radius = a.radius[pi];
    const int64_t t = a.table.parent[q];
    const double top = a.table.top[q], bottom = a.table.bottom[q];
    const int kind = a.table.tess_kind[t];
    const bool variable = kind != TVD_DENS_CONST;
    double p[TVD_NPAR];
#pragma unroll
    for (int k = 0; k < TVD_NPAR; k++)
        p[k] = a.table.tess_params[TVD_NPAR * t + k];

    Level stack[TVD_MAX_DEPTH];
    int depth = 0;
    int stktop = 0;                 // the reference's row index of the stack top (:73)
    double w = a.table.tess_bounds[6 * t + 0], e = a.table.tess_bounds[6 * t + 1];
    double s = a.table.tess_bounds[6 * t + 2], n = a.table.tess_bounds[6 * t + 3];
    double res = 0.0;
    unsigned long long nodes = 0, leaves = 0, splits = 0, too_small = 0;
    bool overflow = false;
    const double rt = 0.5 * (top + bottom) + TVD_R;       // :111, invariant within a piece
    const double rtop = top + TVD_R;                      // :119
    stktop -= 1;                                          // pop of the root (:81)
    while (true) {
        nodes++;
        // distance_n_size (:104-123).  The split decision compares d with ratio*L where, for the
        // small cells of a deep subdivision, L = rtop*acos(1 - tiny) is dominated by rounding: the
        // reference's decision is then reproduced only by reproducing its roundings, i.e. with
        // correctly rounded sin/cos (double-double, slow).  The geometry is therefore evaluated
        // with the fast sin/cos first and accepted when the decision is safe against the
        // difference between the two evaluations; otherwise it is redone in double-double.
        double distance, Llon, Llat;
        {
            const double lont = TVD_D2R * 0.5 * (w + e);
            const double latt = TVD_D2R * 0.5 * (s + n);
            const double sinlatt = tvd_sin(latt), coslatt = tvd_cos(latt);
            const double cospsi = sinlat * sinlatt + coslat * coslatt * tvd_cos(lon - lont);
            const double d2 = radius * radius + rt * rt - 2 * radius * rt * cospsi;
            distance = sqrt(d2);
            const double arg_lon = sinlatt * sinlatt + (coslatt * coslatt) * tvd_cos(TVD_D2R * (e - w));
            const double arg_lat = tvd_sin(TVD_D2R * n) * tvd_sin(TVD_D2R * s) +
                                   tvd_cos(TVD_D2R * n) * tvd_cos(TVD_D2R * s);
            Llon = rtop * acos(arg_lon);
            Llat = rtop * acos(arg_lat);
            // relative uncertainties: 1-ulp differences in the trig values move the acos arguments
            // by ~2e-16 absolute (L ~ sqrt(2 (1 - arg))) and d^2 by ~2 r rt 2e-16 ~ 0.02 m^2
            const double u_d = 0.02 / d2 + 1e-9;
            const double u_lon = 4e-16 / (1.0 - arg_lon), u_lat = 4e-16 / (1.0 - arg_lat);
            const double tl = a.ratio * Llon, tt = a.ratio * Llat;
            const bool safe = fabs(distance - tl) > (u_d + u_lon) * tl &&
                              fabs(distance - tt) > (u_d + u_lat) * tt;
            if (!safe) {      // also taken for NaN / zero sizes (comparisons false)
                double sl, cl, sn, cn, ss, cs;
                dd_sincos(latt, sl, cl);
                dd_sincos(TVD_D2R * n, sn, cn);
                dd_sincos(TVD_D2R * s, ss, cs);
                const double cp = sinlat * sl + coslat * cl * dd_cos(lon - lont);
                distance = sqrt(radius * radius + rt * rt - 2 * radius * rt * cp);
                Llon = rtop * acos(sl * sl + (cl * cl) * dd_cos(TVD_D2R * (e - w)));
                Llat = rtop * acos(sn * ss + cn * cs);
            }
        }
        // divisions (:129-146)
        int nlon = 1, nlat = 1;
        if (distance <= a.ratio * Llon) {
            if (Llon <= 0.1)
                too_small++;
            else
                nlon = 2;
        }
        if (distance <= a.ratio * Llat) {
            if (Llat <= 0.1) {
                // the reference overwrites error = -1 here, so a cell too small in both
                // dimensions still counts once
                if (!(distance <= a.ratio * Llon && Llon <= 0.1))
                    too_small++;
            } else {
                nlat = 2;
            }
        }
        const int new_cells = nlon * nlat;
        if (new_cells > 1) {
            if (stktop + new_cells > a.stack_size || depth >= TVD_MAX_DEPTH) {   // :88-89
                overflow = true;
                break;
            }
            splits++;
            Level &L = stack[depth++];
            L.w = w;
            L.s = s;
            L.dlon = (e - w) / nlon;                      // split (:187-188)
            L.dlat = (n - s) / nlat;
            L.nlon = nlon;
            L.nlat = nlat;
            L.next = new_cells - 1;
            stktop += new_cells;
        } else {
            NodeTrig nd;
            scale_nodes(w, e, s, n, top, bottom, nd);
            res += glq_literal<FIELD>(variable, kind, p, lon, sinlat, coslat, radius, nd);
            leaves++;
        }
        // pop the next pending node in LIFO order
        while (depth > 0 && stack[depth - 1].next < 0)
            depth--;
        if (depth == 0)
            break;
        Level &L = stack[depth - 1];
        const int c = L.next--;
        stktop -= 1;
        const int ci = c / L.nlat, cj = c - ci * L.nlat;
        w = L.w + ci * L.dlon;                            // :192-195
        e = L.w + (ci + 1) * L.dlon;
        s = L.s + cj * L.dlat;
        n = L.s + (cj + 1) * L.dlat;
    }
    a.pair_result[i] = res;
    atomicAdd(&a.counters[0], nodes);
    atomicAdd(&a.counters[1], leaves);
    atomicAdd(&a.counters[2], splits);
    atomicAdd(&a.counters[3], too_small);
    if (overflow)
        atomicExch(&a.counters[4], 1ull);
    if (splits == 0)
        atomicAdd(&a.counters[5], 1ull);   // pairs whose root was integrated after all
    if (a.leaf_counts)
        atomicAdd(&a.leaf_counts[t * a.n_points + pi], (int)leaves - 1);
}

template <int FIELD>
static cudaError_t launch_near_field(const NearArgs &a, cudaStream_t stream)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.n_pairs + threads - 1) / threads);
    near_walk<FIELD><<<blocks, threads, 0, stream>>>(a);
    tvd_count_launch();
    return cudaGetLastError();
}

cudaError_t tvd_launch_near(const NearArgs &a, cudaStream_t stream)
{
    if (a.n_pairs == 0)
        return cudaSuccess;
    switch (a.field) {
    case TVD_POTENTIAL: return launch_near_field<TVD_POTENTIAL>(a, stream);
    case TVD_GX: return launch_near_field<TVD_GX>(a, stream);
    case TVD_GY: return launch_near_field<TVD_GY>(a, stream);
    case TVD_GZ: return launch_near_field<TVD_GZ>(a, stream);
    case TVD_GXX: return launch_near_field<TVD_GXX>(a, stream);
    case TVD_GXY: return launch_near_field<TVD_GXY>(a, stream);
    case TVD_GXZ: return launch_near_field<TVD_GXZ>(a, stream);
    case TVD_GYY: return launch_near_field<TVD_GYY>(a, stream);
    case TVD_GYZ: return launch_near_field<TVD_GYZ>(a, stream);
    case TVD_GZZ: return launch_near_field<TVD_GZZ>(a, stream);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// Phase C
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
combine(int64_t n_points, int n_groups, const double *__restrict__ far_out, int64_t n_pairs,
        const unsigned long long *__restrict__ keys, const double *__restrict__ pair_result,
        double *__restrict__ result)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points)
        return;
    double r = far_out[p];
    for (int g = 1; g < n_groups; g++)
        r += far_out[(int64_t)g * n_points + p];
    if (n_pairs > 0) {
        // lower bound of (p << 32) in the sorted keys
        const unsigned long long target = (unsigned long long)p << 32;
        int64_t lo = 0, hi = n_pairs;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        double near = 0.0;
        bool any = false;
        while (lo < n_pairs && (int64_t)(keys[lo] >> 32) == p) {
            near += pair_result[lo];
            any = true;
            lo++;
        }
        if (any)
            r += near;
    }
    result[p] = r;
}

cudaError_t tvd_launch_combine(int64_t n_points, int n_groups, const double *far_out,
                               int64_t n_pairs, const unsigned long long *keys,
                               const double *pair_result, double *result, cudaStream_t stream)
{
    if (n_points == 0)
        return cudaSuccess;
    combine<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(n_points, n_groups, far_out,
                                                                    n_pairs, keys, pair_result,
                                                                    result);
    tvd_count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// sensitivity mode: J[t][p] = sum of the far-field values of t's pieces (piece order) + the
// near-field results of t's deferred pieces (piece order) -- the same summation structure as a
// forward run of the one-tesseroid model [t].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jac_reduce(int64_t n_tess, int64_t n_points, const int64_t *__restrict__ tess_offset,
           const double *__restrict__ jq, int64_t n_pairs,
           const unsigned long long *__restrict__ keys, const double *__restrict__ pair_result,
           double *__restrict__ out)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points)
        return;
    for (int64_t t = blockIdx.y; t < n_tess; t += gridDim.y) {
    const int64_t q0 = tess_offset[t], q1 = tess_offset[t + 1];
    double far = 0.0;
    bool any_near = false;
    for (int64_t q = q0; q < q1; q++) {
        const double v = jq[q * n_points + p];
        if (v != v)
            any_near = true;
        else
            far += v;
    }
    double r = far;
    if (any_near && n_pairs > 0) {
        double near = 0.0;
        for (int64_t q = q0; q < q1; q++) {
            const double v = jq[q * n_points + p];
            if (v == v)
                continue;
            const unsigned long long key = ((unsigned long long)p << 32) | (unsigned long long)q;
            int64_t lo = 0, hi = n_pairs;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (keys[mid] < key)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            // a pair that is not in the queue produced a genuine NaN in the sweep
            near += (lo < n_pairs && keys[lo] == key) ? pair_result[lo] : v;
        }
        r += near;
    }
    out[t * n_points + p] = r;
    }
}

cudaError_t tvd_launch_jac_reduce(const PieceTable &t, int64_t n_points, const double *jac_pieces,
                                  int64_t n_pairs, const unsigned long long *keys,
                                  const double *pair_result, double *jac_out, cudaStream_t stream)
{
    if (n_points == 0 || t.n_tess == 0)
        return cudaSuccess;
    dim3 grid((unsigned)((n_points + 255) / 256), (unsigned)(t.n_tess < 65535 ? t.n_tess : 65535));
    jac_reduce<<<grid, 256, 0, stream>>>(t.n_tess, n_points, t.tess_offset, jac_pieces, n_pairs,
                                         keys, pair_result, jac_out);
    tvd_count_launch();
    return cudaGetLastError();
}
