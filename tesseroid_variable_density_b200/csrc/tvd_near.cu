// tvd_near.cu -- Phase B (near-field pairs, literal adaptive subdivision) and Phase C (combine).
//
// Phase B restates rediscretizer (_tesseroid.pyx:34-98) for the (piece, point) pairs that the
// far-field sweep could not prove to be far.  Every node of the subdivision tree is tested with
// the literal distance_n_size / divisions arithmetic (:104-146), its children are generated with
// the reference's own expressions w + i*dlon, w + (i+1)*dlon (:187-195) and every leaf is
// integrated with scale_nodes + kernel* (:152-176, :231-520), so leaf sets -- and hence
// subdivision counts -- are the reference's.
//
// A WARP walks one pair at a time (north_star subsystem 2):
//   * the pending nodes live in a per-warp double-ended queue in shared memory (4 doubles + one
//     word per entry: top/bottom are invariant within a piece).  The queue only holds nodes that
//     are already known to split; each step pops up to 8 of them, the 32 lanes generate and test
//     their (up to 4) children, and warp votes compact the children that split again back into
//     the queue and the leaves into a 64-entry leaf buffer;
//   * nodes are taken from the FRONT (breadth-first: the frontier of these trees is a ring of
//     ~4 pi D^2 cells per level, so two levels fit) and from the BACK (depth-first) when the
//     queue runs short of room; a pair that still would not fit is redone depth-first one node
//     at a time, which needs at most 3 entries per level;
//   * whenever 32 leaves are buffered (or the tree is exhausted) the 32 lanes integrate one leaf
//     each: leaf n of a pair is always summed by lane n mod 32 in increasing n, and the lanes
//     are added by a fixed butterfly.  The traversal, and with it the summation order, is a
//     function of the pair alone -- not of how pairs are batched, scheduled or sharded;
//   * the reference's stack row index `stktop` after a node is popped is -1 + the sum of the
//     push indices along its path, so each entry carries it and ValueError('Tesseroid stack
//     overflow') (:88-89) is reproduced without replaying the reference's LIFO order;
//   * the trigonometry is correctly rounded (tvd_crtrig.cuh: the values glibc returned to the
//     reference), both for the split decisions and for the integration nodes.
// Warps take batches of consecutive pairs from an atomic counter (the trees differ in size by
// orders of magnitude); the roots of a batch are tested one per lane.
//
// Pairs arrive sorted by (point, piece); Phase C adds a point's pair results in piece order,
// which makes every point's value independent of how the points were sharded over blocks, groups
// or GPUs.
#include <cstdlib>
#include "tvd_internal.cuh"

// queue entries per warp (powers of two): the breadth-first frontier holds ~2 x pi D^2 splitting
// cells, so 256 entries serve the default ratios of the potential and gx, gy, gz (D <= 2.5) with
// twice the resident warps, and 512 the tensor's D = 8; larger ratios fall back on depth-first
// steps (slower, never wrong)
#define NW_CAP_SMALL 256
#define NW_CAP_LARGE 512
#define NW_RATIO_SMALL 3.2          // largest ratio served by the small queue
#define NW_LEAFCAP 64               // leaf buffer entries per warp
#define NW_CTX 8                    // doubles of per-pair constants kept in shared memory
#ifndef NW_WARPS
#define NW_WARPS 4                  // warps per block
#endif
#define NW_FULL 0xffffffffu
// The walk calls sin/cos from nine places; out of line the kernel is a third of the size, which
// matters because few warps are resident (instruction fetch stalls were the top stall reason
// with the inlined form, profiles/near_walk_r2.md).
#ifndef NW_INLINE_TRIG
#define NW_SINCOS(ARG_, S_, C_)                    \
    do {                                           \
        const double2 _r = cr_sincos_call(ARG_);   \
        S_ = _r.x;                                 \
        C_ = _r.y;                                 \
    } while (0)
#define NW_COS(ARG_) cr_cos_call(ARG_)
#else
#define NW_SINCOS(ARG_, S_, C_) cr_sincos(ARG_, S_, C_)
#define NW_COS(ARG_) cr_cos(ARG_)
#endif

template <int NW_CAP>
struct NwWarp {                     // per-warp shared-memory state
    double w[NW_CAP], e[NW_CAP], s[NW_CAP], n[NW_CAP];
    // bit 0: splits in longitude, bit 1: splits in latitude, bits 2-7: depth,
    // bits 8-: 1 + the reference's stktop after this node was popped
    int meta[NW_CAP];
    double lw[NW_LEAFCAP], le[NW_LEAFCAP], ls[NW_LEAFCAP], ln[NW_LEAFCAP];
    double ctx[NW_CTX];             // lon sinlat coslat radius rt rtop ratio of the current pair
};

struct PairCtx {                    // what a (piece, point) pair contributes to every node
    double lon, sinlat, coslat, radius, r_sqr;
    double rt, rtop, dr;            // :111, :119, top - bottom
    double rc[2], a2[2], b[2];      // radial nodes; 2*radius*rc[k]; r_sqr + rc[k]^2
    double dkap[2];                 // density(rc[k] - R) * rc[k]^2   (or rc[k]^2: constant density)
    double amp;                     // constant density: the density (:250), else 1
    double ratio;
    int stack_size;
};

// distance_n_size + divisions (_tesseroid.pyx:104-146) of one cell: returns nlon-1 | (nlat-1) << 1,
// and bit 2 when the cell should split but is below the 0.1 m floor (the reference's error
// count).  Out of line: it is called for the roots and inside the walk.
// distance_n_size + divisions (_tesseroid.pyx:104-146) of one cell: returns nlon-1 | (nlat-1) << 1,
// and bit 2 when the cell should split but is below the 0.1 m floor (the reference's error
// count).  *key (optional) = ratio * max(Llon, Llat) / distance: >= 1 for a cell that splits, and
// the larger the deeper its subtree (an estimate, used for scheduling only).
template <bool VEC>
__device__ __forceinline__ int node_test_body(double w, double e, double s, double n, double lon,
                                              double sinlat, double coslat, double radius,
                                              double rt, double rtop, double ratio, float *key)
{
    const double lont = TVD_D2R * 0.5 * (w + e);
    const double latt = TVD_D2R * 0.5 * (s + n);
    double sinlatt, coslatt, sin_n, cos_n, sin_s, cos_s, cos_dlon, cos_ew;
    if (VEC) {
        // the five sin/cos arguments of a node are known up front: one batched evaluation (five
        // interleaved dependency chains) instead of five calls in a row -- the walk of a deep
        // pair is a chain of dependent steps, its latency is what counts
        const double arg[5] = {latt, TVD_D2R * n, TVD_D2R * s, lon - lont, TVD_D2R * (e - w)};
        double sv[5], cv[5];
        cr_sincos_vec<5>(arg, sv, cv);
        sinlatt = sv[0], coslatt = cv[0], sin_n = sv[1], cos_n = cv[1];
        sin_s = sv[2], cos_s = cv[2], cos_dlon = cv[3], cos_ew = cv[4];
    } else {
        // one thread per pair (near_roots): registers matter more than latency
        NW_SINCOS(latt, sinlatt, coslatt);
        NW_SINCOS(TVD_D2R * n, sin_n, cos_n);
        NW_SINCOS(TVD_D2R * s, sin_s, cos_s);
        cos_dlon = NW_COS(lon - lont);
        cos_ew = NW_COS(TVD_D2R * (e - w));
    }
    const double cospsi = sinlat * sinlatt + coslat * coslatt * cos_dlon;
    const double distance = sqrt(radius * radius + rt * rt - 2 * radius * rt * cospsi);
    const double Llon = rtop * acos(sinlatt * sinlatt + (coslatt * coslatt) * cos_ew);
    const double Llat = rtop * acos(sin_n * sin_s + cos_n * cos_s);
    int code = 0;
    const bool near_lon = distance <= ratio * Llon, near_lat = distance <= ratio * Llat;
    // the reference overwrites error = -1, so a cell too small in both dimensions counts once
    const bool small_lon = near_lon && Llon <= 0.1, small_lat = near_lat && Llat <= 0.1;
    if (near_lon && !small_lon)
        code |= 1;
    if (near_lat && !small_lat)
        code |= 2;
    if (small_lon || small_lat)
        code |= 4;
    if (key)
        *key = (float)(ratio * fmax(Llon, Llat)) / (float)distance;
    return code;
}
// Out-of-line form for the walk.  The pair's constants travel through the warp's shared memory
// (NwWarp::ctx): eleven doubles by value would spill to the local-memory stack at every call.
template <bool VEC>
static __device__ __noinline__ int node_test_call(double w, double e, double s, double n,
                                                  const double *ctx)
{
    return node_test_body<VEC>(w, e, s, n, ctx[0], ctx[1], ctx[2], ctx[3], ctx[4], ctx[5], ctx[6],
                               nullptr);
}

// scale_nodes (_tesseroid.pyx:152-176) + kernel{V,x,y,z}[_variable] (:231-520) and the Fatiando 0.5
// tensor integrands for one leaf cell.  cos psi and l^2 keep the reference's operation order;
// l^-1 comes from the MUFU seed + one third-order correction instead of sqrt and a division.
template <int FIELD, bool VEC>
__device__ __forceinline__ double leaf_glq(const PairCtx &c, double w, double e, double s, double n)
{
    constexpr bool NEED_SIN = FIELD == TVD_GY || FIELD == TVD_GXY || FIELD == TVD_GYY ||
                              FIELD == TVD_GYZ;
    const double dlon = e - w, dlat = n - s;
    const double mlon = 0.5 * (e + w), mlat = 0.5 * (n + s);
    double coslon[2], sinlon[2], sinlatc[2], coslatc[2];
    if (VEC) {
        // sin(lonc - lon) = -sin(lon - lonc) exactly (odd function, same |argument|)   (:401)
        double arg[4], sv[4], cv[4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double node = i == 0 ? -TVD_NODE : TVD_NODE;
            arg[i] = TVD_D2R * (0.5 * dlat * node + mlat);
            arg[2 + i] = c.lon - TVD_D2R * (0.5 * dlon * node + mlon);
        }
        cr_sincos_vec<4>(arg, sv, cv);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            sinlatc[i] = sv[i];
            coslatc[i] = cv[i];
            coslon[i] = cv[2 + i];
            sinlon[i] = NEED_SIN ? -sv[2 + i] : 0.0;
        }
    } else {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double node = i == 0 ? -TVD_NODE : TVD_NODE;
        const double lonc = TVD_D2R * (0.5 * dlon * node + mlon);
        const double latc = TVD_D2R * (0.5 * dlat * node + mlat);
        NW_SINCOS(latc, sinlatc[i], coslatc[i]);
        if (NEED_SIN) {
            // sin(lonc - lon) = -sin(lon - lonc) exactly (odd function, same |argument|)   (:401)
            double sl;
            NW_SINCOS(c.lon - lonc, sl, coslon[i]);
            sinlon[i] = -sl;
        } else {
            coslon[i] = NW_COS(c.lon - lonc);
            sinlon[i] = 0.0;
        }
    }
    }
    const double scale = TVD_D2R * dlon * TVD_D2R * dlat * c.dr * 0.125;
    double result = 0.0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double kphi = c.coslat * sinlatc[j] - c.sinlat * coslatc[j] * coslon[i];
            const double cospsi = c.sinlat * sinlatc[j] + c.coslat * coslatc[j] * coslon[i];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const double rc = c.rc[k];
                const double l_sqr = c.b[k] - c.a2[k] * cospsi;     // r^2 + rc^2 - 2 r rc cos psi
                const double dk = c.dkap[k] * coslatc[j];           // [density *] kappa
                const double y = rsqrt_fast(l_sqr);
                double term;
                if (FIELD == TVD_POTENTIAL) {
                    term = dk * y;
                } else if (FIELD == TVD_GX) {
                    term = dk * rc * kphi * (y * y * y);
                } else if (FIELD == TVD_GY) {
                    term = dk * (rc * coslatc[j] * sinlon[i] * (y * y * y));
                } else if (FIELD == TVD_GZ) {
                    term = dk * (rc * cospsi - c.radius) * (y * y * y);
                } else {
                    const double dx = rc * kphi;
                    const double dy = rc * coslatc[j] * sinlon[i];
                    const double dz = rc * cospsi - c.radius;
                    const double y2 = y * y;
                    const double y5 = y2 * y2 * y;
                    double num;
                    if (FIELD == TVD_GXX) num = 3 * (dx * dx) - l_sqr;
                    else if (FIELD == TVD_GXY) num = 3 * dx * dy;
                    else if (FIELD == TVD_GXZ) num = 3 * dx * dz;
                    else if (FIELD == TVD_GYY) num = 3 * (dy * dy) - l_sqr;
                    else if (FIELD == TVD_GYZ) num = 3 * dy * dz;
                    else num = 3 * (dz * dz) - l_sqr;
                    term = dk * num * y5;
                }
                result += term;
            }
        }
    }
    if (FIELD == TVD_GZ)
        result *= -1;
    return result * scale * c.amp;
}

// everything of a pair that does not depend on the cell
__device__ __forceinline__ void make_pair_ctx(const NearArgs &a, int64_t pi, int64_t q, int64_t t,
                                              PairCtx &c, double dens0, double dens1)
{
    c.lon = a.lon[pi];
    c.sinlat = a.sinlat[pi];
    c.coslat = a.coslat[pi];
    c.radius = a.radius[pi];
    c.r_sqr = c.radius * c.radius;
    const double top = a.table.top[q], bottom = a.table.bottom[q];
    c.rt = 0.5 * (top + bottom) + TVD_R;
    c.rtop = top + TVD_R;
    c.dr = top - bottom;
    const double mr = 0.5 * (top + bottom + 2. * TVD_R);
    const bool variable = a.table.tess_kind[t] != TVD_DENS_CONST;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double node = k == 0 ? -TVD_NODE : TVD_NODE;
        c.rc[k] = (0.5 * c.dr * node + mr);
        const double rc_sqr = c.rc[k] * c.rc[k];
        c.a2[k] = 2 * c.radius * c.rc[k];
        c.b[k] = c.r_sqr + rc_sqr;
        // variable density: dens * kappa with kappa = rc_sqr * coslatc (:274-275); constant
        // density: kappa alone, the density multiplies the sum (:250)
        c.dkap[k] = variable ? (k == 0 ? dens0 : dens1) * rc_sqr : rc_sqr;
    }
    c.amp = variable ? 1.0 : a.table.tess_params[TVD_NPAR * t];
    c.ratio = a.ratio;
    c.stack_size = a.stack_size;
}

// densities at the two radial GLQ nodes of a piece (:274), 1 for a constant density
__device__ __forceinline__ void pair_densities(const NearArgs &a, int64_t q, int64_t t,
                                               double &dens0, double &dens1)
{
    dens0 = dens1 = 1.0;
    const int kind = a.table.tess_kind[t];
    if (kind == TVD_DENS_CONST)
        return;
    double p[TVD_NPAR];
#pragma unroll
    for (int k = 0; k < TVD_NPAR; k++)
        p[k] = a.table.tess_params[TVD_NPAR * t + k];
    const double top = a.table.top[q], bottom = a.table.bottom[q];
    const double mr = 0.5 * (top + bottom + 2. * TVD_R), hdr = 0.5 * (top - bottom);
    dens0 = tvd_density(kind, p, (hdr * -TVD_NODE + mr) - TVD_R);
    dens1 = tvd_density(kind, p, (hdr * TVD_NODE + mr) - TVD_R);
}

// Roots, one THREAD per pair: the literal test of the root cell.  A root that does not split
// after all (the sweep's far test is conservative) is integrated here.  A root that splits but
// is barely near-field (class 1: ratio*L/d < 2) usually has only leaves below it -- the
// millions of shallow pairs of a regional grid a few km above the model: its children are
// tested here too and, if none of them splits, integrated here, in the walk's own summation
// order (leaf n on lane n, butterfly: (v0 + v2) + (v1 + v3)), so a pair's bits do not depend
// on which kernel finished it.  Everything else is filed by the estimated depth of its tree
// (class = exponent of ratio*L/d), so that the walk can start with the deepest trees -- they are
// orders of magnitude longer than the median and would otherwise finish alone at the end.
//   ws layout: code[n] (uint8: split code, 0 = done), klass[n] (uint8), order[n] (int32),
//   bins[64] (uint32: [0..31] pairs per class, [32..63] scatter cursors)
template <int FIELD>
__global__ void __launch_bounds__(128)
near_roots(const NearArgs a)
{
    __shared__ double s_ctx[128][NW_CTX + 1];       // + 1: rows on different banks
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < a.n_pairs;
    int code = 0, small = 0, klass = 255;
    unsigned nodes = 0, leaves = 0, splits = 0, rootleaf = 0;
    bool overflow = false;
    if (have) {
        const unsigned long long key = a.keys[i];
        const int64_t pi = (int64_t)(key >> 32), q = (int64_t)(key & 0xffffffffull);
        const int64_t t = a.table.parent[q];
        double dens0, dens1;
        pair_densities(a, q, t, dens0, dens1);
        PairCtx c;
        make_pair_ctx(a, pi, q, t, c, dens0, dens1);
        const double w = a.table.tess_bounds[6 * t + 0], e = a.table.tess_bounds[6 * t + 1];
        const double s = a.table.tess_bounds[6 * t + 2], n = a.table.tess_bounds[6 * t + 3];
        float depth_key = 0.f;
        const int r = node_test_body<false>(w, e, s, n, c.lon, c.sinlat, c.coslat, c.radius, c.rt,
                                            c.rtop, c.ratio, &depth_key);
        code = r & 3;
        small = (r >> 2) & 1;
        nodes = 1;
        if (code == 0) {
            a.pair_result[i] = leaf_glq<FIELD, false>(c, w, e, s, n);
            leaves = rootleaf = 1;
        } else {
            const int cells = code == 3 ? 4 : 2;
            if (-1 + cells > c.stack_size)
                overflow = true;        // stktop + new_cells > STACK_SIZE at the root (:88)
            int ex = 0;
            frexpf(depth_key, &ex);     // key in [2^(ex-1), 2^ex)
            klass = ex < 0 ? 0 : (ex > 31 ? 31 : ex);
            if (klass <= 1 && !overflow && a.shallow) {
                double *ctx = s_ctx[threadIdx.x];
                ctx[0] = c.lon;
                ctx[1] = c.sinlat;
                ctx[2] = c.coslat;
                ctx[3] = c.radius;
                ctx[4] = c.rt;
                ctx[5] = c.rtop;
                ctx[6] = c.ratio;
                // split (:187-195): children in push order (i, j) = (0,0),(0,1),(1,0),(1,1)
                const int nlon = 1 + (code & 1), nlat = 1 + ((code >> 1) & 1);
                const double dlon = (e - w) / nlon, dlat = (n - s) / nlat;
                bool shallow = true;
                int csmall = 0;
                for (int ch = 0; ch < cells && shallow; ch++) {
                    const int ci = nlat == 2 ? (ch >> 1) : ch, cj = ch - ci * nlat;
                    const int rr = node_test_call<false>(w + ci * dlon, w + (ci + 1) * dlon,
                                                         s + cj * dlat, s + (cj + 1) * dlat, ctx);
                    shallow = (rr & 3) == 0;
                    csmall += (rr >> 2) & 1;
                }
                if (shallow) {
                    double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
                    for (int ch = 0; ch < cells; ch++) {
                        const int ci = nlat == 2 ? (ch >> 1) : ch, cj = ch - ci * nlat;
                        const double val = leaf_glq<FIELD, false>(c, w + ci * dlon, w + (ci + 1) * dlon,
                                                           s + cj * dlat, s + (cj + 1) * dlat);
                        if (ch == 0) v[0] = val;
                        else if (ch == 1) v[1] = val;
                        else if (ch == 2) v[2] = val;
                        else v[3] = val;
                    }
                    // the walk's order: lane n holds leaf n, lanes are added by the butterfly
                    a.pair_result[i] = cells == 4 ? (v[0] + v[2]) + (v[1] + v[3]) : v[0] + v[1];
                    nodes += cells;
                    leaves += cells;
                    splits = 1;
                    small += csmall;
                    if (a.leaf_counts)
                        atomicAdd(&a.leaf_counts[t * a.n_points + pi], cells - 1);
                    klass = 255;
                    code = 0;
                }
            }
        }
        a.root_code[i] = (unsigned char)code;
        a.root_class[i] = (unsigned char)klass;
    }
    {
        // one atomic per (warp, class): a slab's millions of pairs fall into two or three classes
        const unsigned peers = __match_any_sync(NW_FULL, klass);
        if (klass <= 31 && (threadIdx.x & 31) == __ffs(peers) - 1)
            atomicAdd(a.bins + klass, (unsigned)__popc(peers));
    }
    const unsigned n_nodes = __reduce_add_sync(NW_FULL, nodes);
    const unsigned n_leaves = __reduce_add_sync(NW_FULL, leaves);
    const unsigned n_splits = __reduce_add_sync(NW_FULL, splits);
    const unsigned n_rootleaf = __reduce_add_sync(NW_FULL, rootleaf);
    const unsigned n_small = __reduce_add_sync(NW_FULL, (unsigned)small);
    const bool any_over = __any_sync(NW_FULL, overflow);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.counters[0], (unsigned long long)n_nodes);
        if (n_leaves)
            atomicAdd(&a.counters[1], (unsigned long long)n_leaves);
        if (n_splits)
            atomicAdd(&a.counters[2], (unsigned long long)n_splits);
        if (n_rootleaf)
            atomicAdd(&a.counters[5], (unsigned long long)n_rootleaf);
        if (n_small)
            atomicAdd(&a.counters[3], (unsigned long long)n_small);
        if (any_over)
            atomicExch(&a.counters[4], 1ull);
    }
}

// order[] = the pairs that split, deepest class first (any order within a class: a pair's value
// does not depend on when it is walked)
__global__ void __launch_bounds__(256)
near_order(const NearArgs a)
{
    __shared__ unsigned start[32];
    if (threadIdx.x < 32) {
        unsigned s = 0;
        for (int k = 31; k > (int)threadIdx.x; k--)
            s += a.bins[k];
        start[threadIdx.x] = s;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = i < a.n_pairs ? a.root_class[i] : 255;
    const unsigned peers = __match_any_sync(NW_FULL, k);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (k <= 31 && lane == leader)
        base = atomicAdd(a.bins + 32 + k, (unsigned)__popc(peers));
    base = __shfl_sync(NW_FULL, base, leader);
    if (k <= 31)
        a.order[start[k] + base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)i;
}

#ifndef NW_MINBLOCKS
#define NW_MINBLOCKS 2              // resident blocks per SM the register allocation is sized for
#endif
template <int FIELD, int NW_CAP>
__global__ void __launch_bounds__(32 * NW_WARPS, NW_MINBLOCKS)
near_walk(const NearArgs a)
{
    extern __shared__ __align__(16) unsigned char nw_smem[];
    NwWarp<NW_CAP> &S = reinterpret_cast<NwWarp<NW_CAP> *>(nw_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    // warp-uniform totals (every lane holds the same numbers)
    unsigned long long t_nodes = 0, t_leaves = 0, t_splits = 0, t_small = 0;
    bool overflow = false;
    const int64_t n_walk = (int64_t)__reduce_add_sync(NW_FULL, a.bins[lane]);

    for (;;) {
        unsigned long long b = 0;
        if (lane == 0)
            b = atomicAdd(a.counters + 6, 1ull);
        b = __shfl_sync(NW_FULL, b, 0);
        const int64_t i0 = (int64_t)b * a.batch;
        if (i0 >= n_walk)
            break;
        const int nb = (int)((n_walk - i0) < a.batch ? (n_walk - i0) : a.batch);

        // ---- the trees of the batch, one pair at a time -----------------------------------------
        for (int jp = 0; jp < nb; jp++) {
            const int64_t ip = a.order[i0 + jp];
            const int rc0 = a.root_code[ip];
            const unsigned long long key = a.keys[ip];
            const int64_t pi = (int64_t)(key >> 32), q = (int64_t)(key & 0xffffffffull);
            const int64_t t = a.table.parent[q];
            double d0, d1;
            pair_densities(a, q, t, d0, d1);
            PairCtx c;
            make_pair_ctx(a, pi, q, t, c, d0, d1);
            unsigned p_nodes = 0, p_leaves = 0, p_splits = 0, p_small = 0;
            double acc = 0.0;
            // mode 0: breadth-first with depth-first relief; mode 1: depth-first, one node a step
            for (int mode = 0; mode < 2; mode++) {
                p_nodes = p_leaves = p_splits = p_small = 0;
                acc = 0.0;
                int head = 0, count = 1, nleaf = 0;
                bool queue_full = false;
                __syncwarp();
                if (lane == 0) {
                    S.ctx[0] = c.lon;
                    S.ctx[1] = c.sinlat;
                    S.ctx[2] = c.coslat;
                    S.ctx[3] = c.radius;
                    S.ctx[4] = c.rt;
                    S.ctx[5] = c.rtop;
                    S.ctx[6] = c.ratio;
                    S.w[0] = a.table.tess_bounds[6 * t + 0];
                    S.e[0] = a.table.tess_bounds[6 * t + 1];
                    S.s[0] = a.table.tess_bounds[6 * t + 2];
                    S.n[0] = a.table.tess_bounds[6 * t + 3];
                    S.meta[0] = rc0;                // depth 0, stktop after pop = -1
                }
                __syncwarp();
                while (count > 0 || nleaf > 0) {
                    if (count > 0) {
                        const int room = NW_CAP - count;
                        int k;
                        bool from_back;
                        if (mode == 1) {
                            k = 1;
                            from_back = true;
                        } else if (room >= 64) {
                            k = count < 8 ? count : 8;
                            from_back = false;
                        } else {
                            k = count < 8 ? count : 8;
                            if (k > room / 3)
                                k = room / 3;
                            from_back = true;
                        }
                        if (k == 0 || room < 3) {
                            queue_full = true;
                            break;
                        }
                        const int idx = lane >> 2, ch = lane & 3;
                        bool valid = idx < k;
                        double cw = 0, ce = 0, cs = 0, cn = 0;
                        int cmeta = 0;
                        if (valid) {
                            const int pos = from_back ? ((head + count - 1 - idx) & (NW_CAP - 1))
                                                      : ((head + idx) & (NW_CAP - 1));
                            const double pw = S.w[pos], pe = S.e[pos], ps = S.s[pos], pn = S.n[pos];
                            const int pm = S.meta[pos];
                            const int nlon = 1 + (pm & 1), nlat = 1 + ((pm >> 1) & 1);
                            valid = ch < nlon * nlat;
                            // split (:187-195): children in push order (i, j) = (0,0),(0,1),(1,0),(1,1)
                            const int ci = nlat == 2 ? (ch >> 1) : ch, cj = ch - ci * nlat;
                            const double dlon = (pe - pw) / nlon, dlat = (pn - ps) / nlat;
                            cw = pw + ci * dlon;
                            ce = pw + (ci + 1) * dlon;
                            cs = ps + cj * dlat;
                            cn = ps + (cj + 1) * dlat;
                            // a child pushed with index ch is popped with ch siblings still below it
                            cmeta = ((((pm >> 2) & 63) + 1) << 2) | (((pm >> 8) + ch) << 8);
                        }
                        __syncwarp();       // every parent is in registers before slots are reused
                        if (!from_back)
                            head = (head + k) & (NW_CAP - 1);
                        count -= k;
                        int code = 0, small = 0;
                        if (valid) {
                            const int r = node_test_call<true>(cw, ce, cs, cn, S.ctx);
                            code = r & 3;
                            small = (r >> 2) & 1;
                        }
                        const bool split = valid && code != 0;
                        const bool leaf = valid && code == 0;
                        if (split) {
                            const int stk = (cmeta >> 8) - 1, cells = code == 3 ? 4 : 2;
                            if (stk + cells > c.stack_size || ((cmeta >> 2) & 63) >= TVD_MAX_DEPTH - 1)
                                overflow = true;                        // :88-89
                        }
                        const unsigned m_valid = __ballot_sync(NW_FULL, valid);
                        const unsigned m_split = __ballot_sync(NW_FULL, split);
                        const unsigned m_leaf = __ballot_sync(NW_FULL, leaf);
                        p_nodes += __popc(m_valid);
                        p_splits += __popc(m_split);
                        p_leaves += __popc(m_leaf);
                        p_small += __reduce_add_sync(NW_FULL, (unsigned)small);
                        if (split) {
                            const int pos = (head + count + __popc(m_split & lt_mask)) & (NW_CAP - 1);
                            S.w[pos] = cw;
                            S.e[pos] = ce;
                            S.s[pos] = cs;
                            S.n[pos] = cn;
                            S.meta[pos] = cmeta | code;
                        }
                        if (leaf) {
                            const int pos = nleaf + __popc(m_leaf & lt_mask);
                            S.lw[pos] = cw;
                            S.le[pos] = ce;
                            S.ls[pos] = cs;
                            S.ln[pos] = cn;
                        }
                        count += __popc(m_split);
                        nleaf += __popc(m_leaf);
                        __syncwarp();
                        if (__any_sync(NW_FULL, overflow)) {
                            overflow = true;
                            break;
                        }
                    }
                    if (nleaf >= 32 || (count == 0 && nleaf > 0)) {
                        const int m = nleaf < 32 ? nleaf : 32;
                        if (lane < m)
                            acc += leaf_glq<FIELD, true>(c, S.lw[lane], S.le[lane], S.ls[lane], S.ln[lane]);
                        __syncwarp();
                        const int rest = nleaf - m;         // < 32: no overlap with the slots read
                        double mw = 0, me = 0, ms = 0, mn = 0;
                        if (lane < rest) {
                            mw = S.lw[32 + lane];
                            me = S.le[32 + lane];
                            ms = S.ls[32 + lane];
                            mn = S.ln[32 + lane];
                        }
                        __syncwarp();
                        if (lane < rest) {
                            S.lw[lane] = mw;
                            S.le[lane] = me;
                            S.ls[lane] = ms;
                            S.ln[lane] = mn;
                        }
                        nleaf = rest;
                        __syncwarp();
                    }
                }
                if (!queue_full || overflow)
                    break;
                if (mode == 1)
                    overflow = true;    // unreachable: depth-first needs <= 3 entries per level
            }
            // fixed butterfly: every lane ends with the same bits
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc += __shfl_xor_sync(NW_FULL, acc, o);
            if (lane == 0) {
                a.pair_result[ip] = acc;
                if (a.leaf_counts)
                    atomicAdd(&a.leaf_counts[t * a.n_points + pi], (int)p_leaves - 1);
            }
            t_nodes += p_nodes;
            t_leaves += p_leaves;
            t_splits += p_splits + 1;       // + the root
            t_small += p_small;
            if (overflow)
                break;
        }
        if (__any_sync(NW_FULL, overflow)) {
            overflow = true;
            break;
        }
    }
    if (lane == 0) {
        atomicAdd(&a.counters[0], t_nodes);
        atomicAdd(&a.counters[1], t_leaves);
        atomicAdd(&a.counters[2], t_splits);
        atomicAdd(&a.counters[3], t_small);
        if (overflow)
            atomicExch(&a.counters[4], 1ull);
    }
}

template <int FIELD, int NW_CAP>
static cudaError_t launch_near_walk(NearArgs a, cudaStream_t stream)
{
    static int blocks_per_sm[64] = {0};
    const size_t smem = sizeof(NwWarp<NW_CAP>) * NW_WARPS;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    if (blocks_per_sm[dev] == 0) {
        e = cudaFuncSetAttribute(near_walk<FIELD, NW_CAP>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, near_walk<FIELD, NW_CAP>,
                                                          32 * NW_WARPS, smem);
        if (e != cudaSuccess)
            return e;
        blocks_per_sm[dev] = nb > 0 ? nb : 1;
    }
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess)
        return e;
    const int64_t warps = (int64_t)sms * blocks_per_sm[dev] * NW_WARPS;
    // pairs per grab: the deepest trees come first and are taken one at a time (the number of
    // pairs that split is only known on the device; all pairs bound it)
    int64_t batch = a.n_pairs / (64 * warps);
    batch = batch < 1 ? 1 : (batch > 32 ? 32 : batch);
    a.batch = (int)batch;
    const int64_t n_batches = (a.n_pairs + batch - 1) / batch;
    int64_t blocks = (n_batches + NW_WARPS - 1) / NW_WARPS;
    if (blocks > (int64_t)sms * blocks_per_sm[dev])
        blocks = (int64_t)sms * blocks_per_sm[dev];
    near_walk<FIELD, NW_CAP><<<(unsigned)blocks, 32 * NW_WARPS, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int FIELD>
static cudaError_t launch_near_field(NearArgs a, cudaStream_t stream)
{
    // workspace: code[n] klass[n] (bytes), then order[n] (int32), then bins[64] (uint32)
    unsigned char *ws = reinterpret_cast<unsigned char *>(a.workspace);
    const size_t n = (size_t)a.n_pairs;
    const size_t off_order = ((2 * n + 15) / 16) * 16;
    a.root_code = ws;
    a.root_class = ws + n;
    a.order = reinterpret_cast<int32_t *>(ws + off_order);
    a.bins = reinterpret_cast<unsigned *>(ws + off_order + 4 * n);
    cudaError_t e = cudaMemsetAsync(a.bins, 0, 64 * sizeof(unsigned), stream);
    if (e != cudaSuccess)
        return e;
    // The root kernel's one-level expansion pays when there are millions of barely-near pairs; on
    // a small problem it only lengthens the kernel by its few threads that take it.  Either way
    // the pair gets the same bits (the walk's own summation order), so this is a free choice.
    a.shallow = a.n_pairs >= 65536 ? 1 : 0;
    if (const char *dbg = getenv("TVD_DEBUG_SHALLOW"))     // tests force either way with this
        a.shallow = dbg[0] == '1' ? 1 : 0;
    near_roots<FIELD><<<(unsigned)((a.n_pairs + 127) / 128), 128, 0, stream>>>(a);
    near_order<<<(unsigned)((a.n_pairs + 255) / 256), 256, 0, stream>>>(a);
    tvd_count_launch(3);
    if (a.ratio <= NW_RATIO_SMALL)
        return launch_near_walk<FIELD, NW_CAP_SMALL>(a, stream);
    return launch_near_walk<FIELD, NW_CAP_LARGE>(a, stream);
}

// bytes of NearArgs::workspace for n pairs
size_t tvd_near_workspace_bytes(int64_t n_pairs)
{
    const size_t n = (size_t)n_pairs;
    return ((2 * n + 15) / 16) * 16 + 4 * n + 64 * sizeof(unsigned) + 16;
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
