// tvd_far.cu -- Phase A: the far-field sweep (the dominant kernel), sm_100a FP64.
//
// One thread owns one computation point and walks the piece records of its piece group in
// sweep order (tesseroids sorted by longitude extent, each followed by its radial pieces).
// Records are staged through shared memory in TVD_TILE-piece tiles by bulk asynchronous copies
// (cp.async.bulk, i.e. the TMA 1-D path, completing on an mbarrier; double buffered) and are read
// as broadcast LDS.128 (all lanes read the same record).
//
//   per tile      one test of the warp against the bounding sphere of the tile's pieces: if every
//                 lane is beyond the tile's largest threshold, the tile is swept without
//                 per-pair tests (the common case);
//   per pair      otherwise a conservative far-field test on the squared centre distance
//                 (3 DFMA + 1 DSETP, no sqrt, no libm): if it clears the threshold, the literal
//                 test of the reference (divisions, _tesseroid.pyx:129-146) is guaranteed to say
//                 "no split"; every other pair is appended to the near-field queue
//                 (warp-aggregated atomic) and walked literally by Phase B (tvd_near.cu);
//   far pairs     the 2x2x2 GLQ of the root cell (kernel*, _tesseroid.pyx:231-520) with the
//                 point-independent node data precomputed in the record: its longitude part is
//                 recomputed once per mesh column, its latitude part once per tesseroid, the
//                 radial part per piece; cos psi and l^2 keep the reference's operation order
//                 and are NOT contracted.
//
// The kernel is compiled for a mask of fields (single fields and a few fused sets), for the
// sensitivity mode (JAC) and for points at one height (UR).  The per-point sum over the pieces of
// a group is sequential in sweep order; groups are summed in group order by Phase C, so a point's
// value does not depend on how points are sharded.
#include "tvd_internal.cuh"

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA 1-D)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], "
                 "%2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// cos / sincos of the longitude difference.
//
// libdevice's cos() costs ~20 FP64 + ~25 integer/branch/table-load instructions (special-case
// tests, a Payne-Hanek slow path, coefficients fetched from a global table).  The sweep only
// needs |x| up to a few pi, so it carries its own: Cody-Waite reduction by pi/2 (33-bit head, so
// k*head is exact) keeping the reduction tail, then the fdlibm kernels (error < 1 ulp; for the
// small arguments of regional models the result is correctly rounded in almost every case,
// i.e. it agrees with glibc's cos, which the reference calls).  When every lane of the warp
// has |x| <= pi/4 (regional models) no reduction is done at all.
// ---------------------------------------------------------------------------------------------
// cos (and, if WANT_SIN, sin) of two arguments.  Every lane's result is a function of its own
// argument alone: the warp votes below only choose between code paths that return the same bits
// (a point's value must not depend on which points share its warp, i.e. on how the points were
// chunked over devices or ranks).
//   * |x| < pi/4 for the whole warp: the reduction is skipped; it would return k = 0, r = x and a
//     zero tail, so kcos(x, 0) / ksin(x, 0) ARE the general path's values;
//   * huge, infinite or NaN arguments: libdevice, decided per lane after the common path.
template <bool WANT_SIN>
__device__ __forceinline__ void trig2(double x0, double x1, double &c0, double &c1, double &s0,
                                      double &s1)
{
    // |x| < pi/4 decided on the high words with integer compares (the FP64 pipe is the scarce
    // resource): hi(|x|) < hi(pi/4) = 0x3fe921fb; NaN and inf fail the test
    const bool small = ((__double2hiint(x0) & 0x7fffffff) < 0x3fe921fb) &&
                       ((__double2hiint(x1) & 0x7fffffff) < 0x3fe921fb);
    if (__all_sync(0xffffffffu, small)) {
        c0 = kcos(x0, 0.0);
        c1 = kcos(x1, 0.0);
        if (WANT_SIN) {
            s0 = ksin(x0, 0.0);
            s1 = ksin(x1, 0.0);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        double r, y;
        const int k = reduce_pio2(i == 0 ? x0 : x1, r, y);
        const double cr = kcos(r, y), sr = ksin(r, y);
        // cos(x): k&3 = 0: cr, 1: -sr, 2: -cr, 3: sr ;  sin(x): 0: sr, 1: cr, 2: -sr, 3: -cr
        double c = (k & 1) ? sr : cr;
        if ((k + 1) & 2)
            c = -c;
        double sn = (k & 1) ? cr : sr;
        if (k & 2)
            sn = -sn;
        if (i == 0) {
            c0 = c;
            s0 = sn;
        } else {
            c1 = c;
            s1 = sn;
        }
    }
    // the Cody-Waite reduction is only valid for moderate arguments
    if (!(fabs(x0) < 1.0e5)) {      // also catches NaN/inf
        if (WANT_SIN)
            sincos(x0, &s0, &c0);
        else
            c0 = cos(x0);
    }
    if (!(fabs(x1) < 1.0e5)) {
        if (WANT_SIN)
            sincos(x1, &s1, &c1);
        else
            c1 = cos(x1);
    }
}

// ---------------------------------------------------------------------------------------------
// The sweep is compiled for a MASK of fields (bit f = tvd_field f).  A single field is the
// one-bit mask; fused sets (e.g. potential + gz + gzz) share the far test, the angular half, l^2
// and its inverse square root between their fields (a field's values then agree with its
// single-field sweep to a few ulp: l^-3 is refined directly when only gx/gy/gz are asked for).
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int mask_count(int m)
{
    int c = 0;
    for (int f = 0; f < TVD_NFIELDS; f++)
        c += (m >> f) & 1;
    return c;
}
__host__ __device__ constexpr int mask_slot(int m, int f)      // index of field f among the set bits
{
    int c = 0;
    for (int g = 0; g < f; g++)
        c += (m >> g) & 1;
    return c;
}
__host__ __device__ constexpr bool mask_has(int m, int f) { return (m >> f) & 1; }

template <int MASK>
struct MaskTraits {
    static constexpr int nf = mask_count(MASK);
    static constexpr int trec = tvd_test_rec_len(nf);
    static constexpr bool needs_sinlon = (MASK & ((1 << TVD_GY) | (1 << TVD_GXY) | (1 << TVD_GYY) |
                                                  (1 << TVD_GYZ))) != 0;
    static constexpr bool needs_kphi = (MASK & ((1 << TVD_GX) | (1 << TVD_GXX) | (1 << TVD_GXY) |
                                                (1 << TVD_GXZ))) != 0;
    static constexpr bool needs_y3 = (MASK & ~(1 << TVD_POTENTIAL)) != 0;
    static constexpr bool needs_y5 = (MASK >> TVD_GXX) != 0;
    static constexpr bool needs_dz = (MASK & ((1 << TVD_GZ) | (1 << TVD_GXZ) | (1 << TVD_GYZ) |
                                              (1 << TVD_GZZ))) != 0;
    // gz is the only user of dz: w*dz = (w*rc)*cospsi - w*r is formed directly from per-piece
    // weights (one FMA per node instead of an FMA and a product)
    static constexpr bool fold_dz = (MASK & (1 << TVD_GZ)) != 0 &&
                                    (MASK & ((1 << TVD_GXZ) | (1 << TVD_GYZ) | (1 << TVD_GZZ))) == 0;
    static constexpr bool needs_dx = (MASK & ((1 << TVD_GXX) | (1 << TVD_GXY) | (1 << TVD_GXZ))) != 0;
    static constexpr bool needs_dy = (MASK & ((1 << TVD_GXY) | (1 << TVD_GYY) | (1 << TVD_GYZ))) != 0;
};

// per-thread constants of one computation point
struct PointRegs {
    double lon, sinlat, coslat, r, r_sqr, r2;
    double p2x, p2y, p2z, neg_r_sqr;    // -2 r (cos lat cos lon, cos lat sin lon, sin lat), -r^2
};

// The angular half of the root GLQ: everything that depends on the piece only through its
// lon/lat extent.  The sweep visits the pieces sorted by longitude extent, so
//   - the longitude part (cos/sin of lon - lonc_i) is recomputed only when the extent in
//     longitude changes (once per column of a regular mesh),
//   - the latitude part (cos psi_ij, k_phi_ij) once per tesseroid,
//   - the radial pieces of one tesseroid share both.
struct Angular {
    double coslon[2];
    double sinlon[2];
    double cospsi[2][2];     // [i][j]
    double kphi[2][2];
    double clc[2];           // coslatc[j]
};

template <int MASK>
__device__ __forceinline__ void glq_lon(const PointRegs &pt, const double *__restrict__ g,
                                        Angular &A)
{
    typedef MaskTraits<MASK> T;
    const double2 lonc = *reinterpret_cast<const double2 *>(g + 0);
    // sin(lonc - lon) = -sin(lon - lonc) exactly (odd function, same |argument|)   (:401)
    double s0 = 0.0, s1 = 0.0;
    trig2<T::needs_sinlon>(pt.lon - lonc.x, pt.lon - lonc.y, A.coslon[0], A.coslon[1], s0, s1);
    A.sinlon[0] = -s0;
    A.sinlon[1] = -s1;
}

template <int MASK>
__device__ __forceinline__ void glq_lat(const PointRegs &pt, const double *__restrict__ g,
                                        Angular &A)
{
    typedef MaskTraits<MASK> T;
    const double2 sinlatc = *reinterpret_cast<const double2 *>(g + 2);
    const double2 coslatc = *reinterpret_cast<const double2 *>(g + 4);
    // reference order: sinlat*sinlatc[j] + coslat*coslatc[j]*coslon   (:244), no contraction
    const double ss[2] = {pt.sinlat * sinlatc.x, pt.sinlat * sinlatc.y};
    const double cc[2] = {pt.coslat * coslatc.x, pt.coslat * coslatc.y};
    const double slc[2] = {sinlatc.x, sinlatc.y};
    A.clc[0] = coslatc.x;
    A.clc[1] = coslatc.y;
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            A.cospsi[i][j] = ss[j] + cc[j] * A.coslon[i];
            if (T::needs_kphi)
                A.kphi[i][j] = pt.coslat * slc[j] - pt.sinlat * A.clc[j] * A.coslon[i];   // :322
        }
    }
}

// The radial half: the eight nodes of one piece, all fields of the mask.  out[slot] receives the
// piece's contribution to each field (kernel units, reference sign conventions).
template <int MASK, bool UR>
__device__ __forceinline__ void glq_radial(const PointRegs &pt, const double *__restrict__ g,
                                           const double *__restrict__ ab,
                                           const double *__restrict__ wz, const Angular &A,
                                           double *__restrict__ out)
{
    typedef MaskTraits<MASK> T;
    const double2 rc = *reinterpret_cast<const double2 *>(g + 6);
    const double2 rcsq = *reinterpret_cast<const double2 *>(g + 8);
    const double2 w0 = *reinterpret_cast<const double2 *>(g + 10);   // w[0][0], w[0][1]
    const double2 w1 = *reinterpret_cast<const double2 *>(g + 12);   // w[1][0], w[1][1]
    // r_sqr + rc_sqr - 2*radius*rc[k]*cospsi   (:247), no contraction.  When every point of the
    // call has the same radius (a grid at one height), 2*radius*rc[k] and r_sqr + rc_sqr do not
    // depend on the point and come precomputed (same roundings) with the tile.
    double a[2], b[2];
    if (UR) {
        const double2 a01 = *reinterpret_cast<const double2 *>(ab + 0);
        const double2 b01 = *reinterpret_cast<const double2 *>(ab + 2);
        a[0] = a01.x;
        a[1] = a01.y;
        b[0] = b01.x;
        b[1] = b01.y;
    } else {
        b[0] = pt.r_sqr + rcsq.x;
        b[1] = pt.r_sqr + rcsq.y;
        a[0] = pt.r2 * rc.x;
        a[1] = pt.r2 * rc.y;
    }
    const double rck[2] = {rc.x, rc.y};
    const double wjk[2][2] = {{w0.x, w0.y}, {w1.x, w1.y}};     // rho_k * kappa_jk * scale
    // field-specific node weights, same products as the reference's integrands
    double wx[2][2], wy[2][2];
    if (mask_has(MASK, TVD_GX) || mask_has(MASK, TVD_GY)) {
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                wx[j][k] = wjk[j][k] * rck[k];                 // kappa*rc[k]*kphi            (:328)
                wy[j][k] = wx[j][k] * A.clc[j];                // kappa*rc[k]*coslatc*sinlon  (:408)
            }
    }
    // gz alone needs dz only inside w*dz = w*(rc*cospsi - radius) (:487): with wr = w*rc and
    // wq = w*radius per (j, k) it is ONE FMA per node, fma(wr, cospsi, -wq).  Points at one height
    // get wr and wq precomputed with the tile (same products, same roundings).
    double wr[2][2], wq[2][2];
    if (T::fold_dz) {
        if (UR) {
            const double2 r0 = *reinterpret_cast<const double2 *>(wz + 0);
            const double2 r1 = *reinterpret_cast<const double2 *>(wz + 2);
            const double2 q0 = *reinterpret_cast<const double2 *>(wz + 4);
            const double2 q1 = *reinterpret_cast<const double2 *>(wz + 6);
            wr[0][0] = r0.x, wr[0][1] = r0.y, wr[1][0] = r1.x, wr[1][1] = r1.y;
            wq[0][0] = q0.x, wq[0][1] = q0.y, wq[1][0] = q1.x, wq[1][1] = q1.y;
        } else {
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    wr[j][k] = wjk[j][k] * rck[k];
                    wq[j][k] = wjk[j][k] * pt.r;
                }
        }
    }

    double sum[T::nf];
#pragma unroll
    for (int s = 0; s < T::nf; s++)
        sum[s] = 0.0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double cospsi = A.cospsi[i][j];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const double l_sqr = b[k] - a[k] * cospsi;
                const double w = wjk[j][k];
                double y = 0.0, y3 = 0.0, y5 = 0.0, dx = 0.0, dy = 0.0, dz = 0.0;
                if (T::needs_y5 || mask_has(MASK, TVD_POTENTIAL)) {
                    y = rsqrt_fast(l_sqr);
                    if (T::needs_y3) {
                        const double y2 = y * y;
                        y3 = y2 * y;
                        if (T::needs_y5)
                            y5 = y3 * y2;
                    }
                } else {
                    y3 = rsqrt3_fast(l_sqr);       // gx, gy, gz only need l^-3
                }
                // rc*cospsi - radius (:487); the product is not ill-conditioned enough to need
                // the reference's separate rounding (<= 5e-14 relative)
                if (T::needs_dz && !T::fold_dz)
                    dz = fma(rck[k], cospsi, -pt.r);
                if (T::needs_dx)
                    dx = rck[k] * A.kphi[i][j];
                if (T::needs_dy)
                    dy = rck[k] * A.clc[j] * A.sinlon[i];
                if (mask_has(MASK, TVD_POTENTIAL)) {
                    double &acc = sum[mask_slot(MASK, TVD_POTENTIAL)];
                    acc = fma(w, y, acc);                                   // kappa/sqrt(l_sqr)
                }
                if (mask_has(MASK, TVD_GX)) {
                    double &acc = sum[mask_slot(MASK, TVD_GX)];
                    acc = fma(wx[j][k] * A.kphi[i][j], y3, acc);
                }
                if (mask_has(MASK, TVD_GY)) {
                    double &acc = sum[mask_slot(MASK, TVD_GY)];
                    acc = fma(wy[j][k] * A.sinlon[i], y3, acc);
                }
                if (mask_has(MASK, TVD_GZ)) {
                    double &acc = sum[mask_slot(MASK, TVD_GZ)];
                    if (T::fold_dz)
                        acc = fma(-fma(wr[j][k], cospsi, -wq[j][k]), y3, acc);
                    else
                        acc = fma(-(w * dz), y3, acc);                      // z down (:490)
                }
                // gradient tensor (Fatiando 0.5 integrands, z up)
                if (mask_has(MASK, TVD_GXX)) {
                    double &acc = sum[mask_slot(MASK, TVD_GXX)];
                    acc = fma(w * fma(3.0 * dx, dx, -l_sqr), y5, acc);
                }
                if (mask_has(MASK, TVD_GXY)) {
                    double &acc = sum[mask_slot(MASK, TVD_GXY)];
                    acc = fma(w * (3.0 * dx * dy), y5, acc);
                }
                if (mask_has(MASK, TVD_GXZ)) {
                    double &acc = sum[mask_slot(MASK, TVD_GXZ)];
                    acc = fma(w * (3.0 * dx * dz), y5, acc);
                }
                if (mask_has(MASK, TVD_GYY)) {
                    double &acc = sum[mask_slot(MASK, TVD_GYY)];
                    acc = fma(w * fma(3.0 * dy, dy, -l_sqr), y5, acc);
                }
                if (mask_has(MASK, TVD_GYZ)) {
                    double &acc = sum[mask_slot(MASK, TVD_GYZ)];
                    acc = fma(w * (3.0 * dy * dz), y5, acc);
                }
                if (mask_has(MASK, TVD_GZZ)) {
                    double &acc = sum[mask_slot(MASK, TVD_GZZ)];
                    acc = fma(w * fma(3.0 * dz, dz, -l_sqr), y5, acc);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < T::nf; s++)
        out[s] = sum[s];
}

// acc += v if lhs > rhs, as one compare and one predicated add (the compiler's own code for
// `if (far) acc += v` re-evaluates the compare, adds unconditionally and selects both halves)
__device__ __forceinline__ void add_if_gt(double &acc, double v, double lhs, double rhs)
{
    asm("{\n"
        ".reg .pred q;\n"
        "setp.gt.f64 q, %2, %3;\n"
        "@q add.f64 %0, %0, %1;\n"
        "}\n"
        : "+d"(acc)
        : "d"(v), "d"(lhs), "d"(rhs));
}

// shared-memory layout of one sweep instantiation (doubles per stage, bytes in all)
template <int MASK, bool UR>
struct FarSmem {
    static constexpr int n_test = TVD_TILE * MaskTraits<MASK>::trec;
    static constexpr int n_glq = TVD_TILE * TVD_GLQ_REC;
    static constexpr int n_ab = UR ? TVD_TILE * TVD_AB_REC : 0;
    static constexpr int n_wz = (UR && MaskTraits<MASK>::fold_dz) ? TVD_TILE * TVD_WZ_REC : 0;
    static constexpr int n_doubles = 2 * (n_test + n_glq + n_ab + n_wz);
    static constexpr size_t bytes = (size_t)n_doubles * sizeof(double) + 2 * sizeof(uint64_t);
};

template <int MASK, int THREADS, int MINBLOCKS, bool JAC, bool UR>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
far_sweep(const FarArgs a)
{
    typedef MaskTraits<MASK> T;
    constexpr int NF = T::nf;
    constexpr int TREC = T::trec;
    // dynamic shared memory (some field sets exceed the 48 KB of static shared memory):
    // two stages of test, GLQ, one-height (a, b) and one-height (w*rc, w*r) records, two mbarriers
    typedef FarSmem<MASK, UR> L;
    extern __shared__ __align__(16) unsigned char far_smem[];
    double *const s_base = reinterpret_cast<double *>(far_smem);
    // stage st of each record kind (pointer arithmetic, not pointer arrays: those would live in
    // local memory)
    auto s_test = [&](int st) { return s_base + st * L::n_test; };
    auto s_glq = [&](int st) { return s_base + 2 * L::n_test + st * L::n_glq; };
    auto s_ab = [&](int st) { return s_base + 2 * (L::n_test + L::n_glq) + st * L::n_ab; };
    auto s_wz = [&](int st) {
        return s_base + 2 * (L::n_test + L::n_glq + L::n_ab) + st * L::n_wz;
    };
    uint64_t *const s_bar = reinterpret_cast<uint64_t *>(s_base + L::n_doubles);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int g = blockIdx.y;
    const int64_t q_begin = (int64_t)g * a.pieces_per_group;
    int64_t q_end = q_begin + a.pieces_per_group;
    if (q_end > a.n_pieces)
        q_end = a.n_pieces;
    const int n_tiles = (int)((q_end - q_begin + TVD_TILE - 1) / TVD_TILE);

    PointRegs pt;
    Angular ang;
    const int64_t pidx = (int64_t)blockIdx.x * THREADS + tid;
    const bool valid = pidx < a.n_points;
    double acc[NF];
    {
        const int64_t pi = valid ? pidx : 0;     // padding threads shadow point 0
        pt.lon = a.lon[pi];
        pt.sinlat = a.sinlat[pi];
        pt.coslat = a.coslat[pi];
        pt.r = a.radius[pi];
        pt.r_sqr = pt.r * pt.r;        // radius**2 (:239)
        pt.r2 = 2.0 * pt.r;            // 2*radius  (:247)
        double sl, cl;
        sincos(pt.lon, &sl, &cl);
        pt.p2x = -pt.r2 * (pt.coslat * cl);
        pt.p2y = -pt.r2 * (pt.coslat * sl);
        pt.p2z = -pt.r2 * pt.sinlat;
        pt.neg_r_sqr = -pt.r_sqr;
#pragma unroll
        for (int s = 0; s < NF; s++)
            acc[s] = 0.0;
    }

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile) {
        const int st = tile & 1;
        const int64_t q0 = q_begin + (int64_t)tile * TVD_TILE;
        int64_t nq = q_end - q0;
        if (nq > TVD_TILE)
            nq = TVD_TILE;
        const uint32_t bt = (uint32_t)(nq * TREC * sizeof(double));
        const uint32_t bg = (uint32_t)(nq * TVD_GLQ_REC * sizeof(double));
        const uint32_t ba = UR ? (uint32_t)(nq * TVD_AB_REC * sizeof(double)) : 0u;
        const uint32_t bw = (UR && T::fold_dz) ? (uint32_t)(nq * TVD_WZ_REC * sizeof(double)) : 0u;
        mbar_expect_tx(&s_bar[st], bt + bg + ba + bw);
        bulk_g2s(s_test(st), a.test_rec + q0 * TREC, bt, &s_bar[st]);
        bulk_g2s(s_glq(st), a.glq_rec + q0 * TVD_GLQ_REC, bg, &s_bar[st]);
        if (UR)
            bulk_g2s(s_ab(st), a.ab_rec + q0 * TVD_AB_REC, ba, &s_bar[st]);
        if (UR && T::fold_dz)
            bulk_g2s(s_wz(st), a.wz_rec + q0 * TVD_WZ_REC, bw, &s_bar[st]);
    };

    if (tid == 0) {
        if (n_tiles > 0)
            issue(0);
        if (n_tiles > 1)
            issue(1);
    }

    for (int tile = 0; tile < n_tiles; tile++) {
        const int st = tile & 1;
        mbar_wait(&s_bar[st], (uint32_t)((tile >> 1) & 1));
        const int64_t q0 = q_begin + (int64_t)tile * TVD_TILE;
        int nq = (int)((q_end - q0) < TVD_TILE ? (q_end - q0) : TVD_TILE);
        const double *tt = s_test(st);
        const double *gg = s_glq(st);
        const double *aa = s_ab(st);
        const double *ww = s_wz(st);
        // d^2-part of the far test of piece qi, computed one piece ahead so that its dependent
        // FMA chain overlaps the radial work of the previous piece
        // conservative far-field test: d^2 - thr_f^2 = r^2 + (rt^2 - thr_f^2) - 2 p.c > 0 with
        // p, c the Cartesian position vectors of the point and the piece centre
        // (FMA allowed: only the guarded comparison depends on it, see K4)
        auto far_value = [&](int qi) -> double {
            const double *trec = tt + qi * TREC;
            if (NF == 1) {
                const double2 cxy = *reinterpret_cast<const double2 *>(trec + 0);
                const double2 czk = *reinterpret_cast<const double2 *>(trec + 2);
                double t = fma(pt.p2x, cxy.x, czk.y);
                t = fma(pt.p2y, cxy.y, t);
                return fma(pt.p2z, czk.x, t);
            } else {
                double t = pt.p2x * trec[0];
                t = fma(pt.p2y, trec[1], t);
                return fma(pt.p2z, trec[2], t);
            }
        };
        // Tile-level test: if every lane of the warp is farther from the bounding sphere of the
        // tile's piece centres than the tile's largest threshold (tvd_preprocess.cu:k4_pack_tiles),
        // every per-pair test of the tile is known to pass: the tile is swept without tests,
        // votes or predicated accumulation.  The far sums are the same either way.
        bool tile_far = true;
        {
            const double *tr = a.tile_rec + ((int64_t)g * a.tiles_per_group + tile) * TVD_TILE_REC;
            double tc = fma(pt.p2x, __ldg(tr + 0), __ldg(tr + 3));
            tc = fma(pt.p2y, __ldg(tr + 1), tc);
            tc = fma(pt.p2z, __ldg(tr + 2), tc);
            const double d2c = tc - pt.neg_r_sqr;            // |p - C|^2
#pragma unroll
            for (int s = 0; s < NF; s++)
                tile_far = tile_far && (d2c > __ldg(tr + 4 + s));
            tile_far = __all_sync(0xffffffffu, tile_far || !valid);
        }
        if (tile_far) {
            for (int qi = 0; qi < nq; qi++) {
                const double *grec = gg + qi * TVD_GLQ_REC;
                const int2 flags = *reinterpret_cast<const int2 *>(grec + 14);
                if (flags.y == 0)
                    glq_lon<MASK>(pt, grec, ang);
                if (flags.x == 0)
                    glq_lat<MASK>(pt, grec, ang);
                double v[NF];
                glq_radial<MASK, UR>(pt, grec, aa + (UR ? qi * TVD_AB_REC : 0),
                                     ww + ((UR && T::fold_dz) ? qi * TVD_WZ_REC : 0), ang, v);
                if (JAC) {
                    if (valid)
                        a.jac[reinterpret_cast<const long long *>(grec)[15] * a.n_points + pidx] = v[0];
                } else {
#pragma unroll
                    for (int s = 0; s < NF; s++)
                        acc[s] += v[s];
                }
            }
        } else {
            double t_next = nq > 0 ? far_value(0) : 0.0;
            for (int qi = 0; qi < nq; qi++) {
                const double *trec = tt + qi * TREC;
                const double *grec = gg + qi * TVD_GLQ_REC;
                // record flags (previous piece of the same tile), two 32-bit words: x = same tesseroid
                // (a further radial piece) -> keep the whole angular half; y = same extent in longitude
                // -> keep its longitude part.  Slot 15 = the piece's index in the model-order table.
                const int2 flags = *reinterpret_cast<const int2 *>(grec + 14);   // {same tesseroid, same lon}
                const long long q_model = reinterpret_cast<const long long *>(grec)[15];
                const double t = t_next;
                if (qi + 1 < nq)
                    t_next = far_value(qi + 1);
                bool far[NF];
                if (NF == 1) {
                    far[0] = t > pt.neg_r_sqr;
                } else {
#pragma unroll
                    for (int s = 0; s < NF; s++)
                        far[s] = (t + trec[3 + s]) > pt.neg_r_sqr;
                }
#pragma unroll
                for (int s = 0; s < NF; s++) {
                    // hot path: one vote and one uniform branch.  Padding threads shadow a real point
                    // and are filtered out only inside the (rare) deferral branch.
                    if (__any_sync(0xffffffffu, !far[s])) {
                        const bool defer = !far[s] && valid;
                        const unsigned m = __ballot_sync(0xffffffffu, defer);
                        unsigned long long base = 0;
                        const int leader = __ffs(m) - 1;
                        if (m != 0 && lane == leader)
                            base = atomicAdd(a.qcount + s, (unsigned long long)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, leader & 31);
                        if (defer) {
                            const unsigned long long idx = base + __popc(m & ((1u << lane) - 1u));
                            if (idx < a.qcap)
                                a.queue[(unsigned long long)s * a.qcap + idx] =
                                    ((unsigned long long)pidx << 32) | (unsigned long long)q_model;
                        }
                    }
                }
                // every lane evaluates the root GLQ (it contains warp-wide votes); lanes whose pair
                // was deferred simply drop the value
                if (flags.y == 0)
                    glq_lon<MASK>(pt, grec, ang);
                if (flags.x == 0)
                    glq_lat<MASK>(pt, grec, ang);
                double v[NF];
                glq_radial<MASK, UR>(pt, grec, aa + (UR ? qi * TVD_AB_REC : 0),
                                     ww + ((UR && T::fold_dz) ? qi * TVD_WZ_REC : 0), ang, v);
                if (JAC) {
                    // sensitivity mode: one value per (piece, point); NaN marks a deferred pair
                    if (valid)
                        a.jac[q_model * a.n_points + pidx] =
                            far[0] ? v[0] : __longlong_as_double(0x7ff8000000000000ll);
                } else {
#pragma unroll
                    for (int s = 0; s < NF; s++)
                        add_if_gt(acc[s], v[s], NF == 1 ? t : t + trec[3 + s], pt.neg_r_sqr);
                }
            }
        }
        __syncthreads();    // everyone is done with stage st
        if (tid == 0 && tile + 2 < n_tiles)
            issue(tile + 2);
    }

    if (valid && !JAC) {
#pragma unroll
        for (int s = 0; s < NF; s++)
            a.far_out[((int64_t)s * a.n_groups + g) * a.n_points + pidx] = acc[s];
    }
}

// ---------------------------------------------------------------------------------------------
// Queue-only sweep: re-runs just the far tests and refills the near-field queues.  Used when
// the queues of the main sweep overflowed: the far sums of that sweep are already complete (a
// deferred pair is left out of them whether or not it found room in the queue), and the exact
// queue sizes are known from its counters.
// ---------------------------------------------------------------------------------------------
template <int NF>
__global__ void __launch_bounds__(TVD_FAR_THREADS)
queue_sweep(const FarArgs a)
{
    constexpr int TREC = tvd_test_rec_len(NF);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    const int64_t q_begin = (int64_t)g * a.pieces_per_group;
    int64_t q_end = q_begin + a.pieces_per_group;
    if (q_end > a.n_pieces)
        q_end = a.n_pieces;
    const int64_t pidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pidx < a.n_points;
    const int64_t pi = valid ? pidx : 0;
    const double lon = a.lon[pi], sinlat = a.sinlat[pi], coslat = a.coslat[pi], r = a.radius[pi];
    double sl, cl;
    sincos(lon, &sl, &cl);
    const double r2 = 2.0 * r;
    const double p2x = -r2 * (coslat * cl), p2y = -r2 * (coslat * sl), p2z = -r2 * sinlat;
    const double neg_r_sqr = -(r * r);
    for (int64_t q = q_begin; q < q_end; q++) {
        const double *trec = a.test_rec + q * TREC;      // same address for every lane
        // identical arithmetic to far_sweep, so the same pairs are deferred
        double t;
        if (NF == 1) {
            t = fma(p2x, trec[0], trec[3]);
            t = fma(p2y, trec[1], t);
            t = fma(p2z, trec[2], t);
        } else {
            t = p2x * trec[0];
            t = fma(p2y, trec[1], t);
            t = fma(p2z, trec[2], t);
        }
#pragma unroll
        for (int s = 0; s < NF; s++) {
            const bool far = (NF == 1 ? t : t + trec[3 + s]) > neg_r_sqr;
            if (__any_sync(0xffffffffu, !far)) {
                const bool defer = !far && valid;
                const unsigned m = __ballot_sync(0xffffffffu, defer);
                unsigned long long base = 0;
                const int leader = __ffs(m) - 1;
                if (m != 0 && lane == leader)
                    base = atomicAdd(a.qcount + s, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader & 31);
                if (defer) {
                    const unsigned long long idx = base + __popc(m & ((1u << lane) - 1u));
                    if (idx < a.qcap)
                        a.queue[(unsigned long long)s * a.qcap + idx] =
                            ((unsigned long long)pidx << 32) |
                            (unsigned long long)reinterpret_cast<const long long *>(
                                a.glq_rec + q * TVD_GLQ_REC)[15];
                }
            }
        }
    }
}

cudaError_t tvd_launch_queue_only(const FarArgs &a, cudaStream_t stream)
{
    if (a.n_points == 0 || a.n_pieces == 0)
        return cudaSuccess;
    int nf = 0;
    for (int f = 0; f < TVD_NFIELDS; f++)
        nf += (a.mask >> f) & 1;
    dim3 grid((unsigned)((a.n_points + TVD_FAR_THREADS - 1) / TVD_FAR_THREADS),
              (unsigned)a.n_groups);
    switch (nf) {
    case 1: queue_sweep<1><<<grid, TVD_FAR_THREADS, 0, stream>>>(a); break;
    case 2: queue_sweep<2><<<grid, TVD_FAR_THREADS, 0, stream>>>(a); break;
    case 3: queue_sweep<3><<<grid, TVD_FAR_THREADS, 0, stream>>>(a); break;
    case 4: queue_sweep<4><<<grid, TVD_FAR_THREADS, 0, stream>>>(a); break;
    case 5: queue_sweep<5><<<grid, TVD_FAR_THREADS, 0, stream>>>(a); break;
    case 6: queue_sweep<6><<<grid, TVD_FAR_THREADS, 0, stream>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    tvd_count_launch();
    return cudaGetLastError();
}

template <int MASK, bool JAC, bool UR>
static cudaError_t launch_sweep(const FarArgs &a, cudaStream_t stream)
{
    constexpr int THREADS = TVD_FAR_THREADS;
    constexpr size_t smem = FarSmem<MASK, UR>::bytes;
    static bool configured[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= 64)
        return cudaErrorInvalidDevice;
    if (!configured[dev]) {
        e = cudaFuncSetAttribute(far_sweep<MASK, THREADS, TVD_FAR_MINBLOCKS, JAC, UR>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        configured[dev] = true;
    }
    dim3 grid((unsigned)((a.n_points + THREADS - 1) / THREADS), (unsigned)a.n_groups);
    far_sweep<MASK, THREADS, TVD_FAR_MINBLOCKS, JAC, UR><<<grid, THREADS, smem, stream>>>(a);
    tvd_count_launch();
    return cudaGetLastError();
}

template <int MASK>
static cudaError_t launch_mask(const FarArgs &a, cudaStream_t stream)
{
    if (a.jac != nullptr) {
        if constexpr (mask_count(MASK) == 1)
            return launch_sweep<MASK, true, false>(a, stream);
        else
            return cudaErrorInvalidValue;
    }
    if (a.ab_rec != nullptr)
        return launch_sweep<MASK, false, true>(a, stream);
    return launch_sweep<MASK, false, false>(a, stream);
}

#define TVD_M(f) (1 << (f))
bool tvd_far_mask_supported(int mask)
{
    switch (mask) {
    case TVD_M(TVD_POTENTIAL): case TVD_M(TVD_GX): case TVD_M(TVD_GY): case TVD_M(TVD_GZ):
    case TVD_M(TVD_GXX): case TVD_M(TVD_GXY): case TVD_M(TVD_GXZ): case TVD_M(TVD_GYY):
    case TVD_M(TVD_GYZ): case TVD_M(TVD_GZZ):
    case TVD_MASK_V_GZ: case TVD_MASK_V_GZ_GZZ: case TVD_MASK_V_G: case TVD_MASK_G:
    case TVD_MASK_TENSOR:
        return true;
    }
    return false;
}

cudaError_t tvd_launch_far(const FarArgs &a, cudaStream_t stream)
{
    if (a.n_points == 0 || a.n_pieces == 0)
        return cudaSuccess;
    switch (a.mask) {
    case TVD_M(TVD_POTENTIAL): return launch_mask<TVD_M(TVD_POTENTIAL)>(a, stream);
    case TVD_M(TVD_GX): return launch_mask<TVD_M(TVD_GX)>(a, stream);
    case TVD_M(TVD_GY): return launch_mask<TVD_M(TVD_GY)>(a, stream);
    case TVD_M(TVD_GZ): return launch_mask<TVD_M(TVD_GZ)>(a, stream);
    case TVD_M(TVD_GXX): return launch_mask<TVD_M(TVD_GXX)>(a, stream);
    case TVD_M(TVD_GXY): return launch_mask<TVD_M(TVD_GXY)>(a, stream);
    case TVD_M(TVD_GXZ): return launch_mask<TVD_M(TVD_GXZ)>(a, stream);
    case TVD_M(TVD_GYY): return launch_mask<TVD_M(TVD_GYY)>(a, stream);
    case TVD_M(TVD_GYZ): return launch_mask<TVD_M(TVD_GYZ)>(a, stream);
    case TVD_M(TVD_GZZ): return launch_mask<TVD_M(TVD_GZZ)>(a, stream);
    case TVD_MASK_V_GZ: return launch_mask<TVD_MASK_V_GZ>(a, stream);
    case TVD_MASK_V_GZ_GZZ: return launch_mask<TVD_MASK_V_GZ_GZZ>(a, stream);
    case TVD_MASK_V_G: return launch_mask<TVD_MASK_V_G>(a, stream);
    case TVD_MASK_G: return launch_mask<TVD_MASK_G>(a, stream);
    case TVD_MASK_TENSOR: return launch_mask<TVD_MASK_TENSOR>(a, stream);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// FP64 peak probe (roofline denominator): 16 independent DFMA chains per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe(double *out, int iters, double x, double y)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        acc[i] = (double)(threadIdx.x + i);
    // 16 independent chains, 16 x 16 = 256 DFMA per loop iteration: an FP64 warp instruction
    // occupies the issue port for two cycles and nothing else issues in its shadow, so the loop
    // overhead (3 instructions) must be negligible for the probe to see the pipe's peak
    for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < 16; i++)
                acc[i] = fma(acc[i], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        s += acc[i];
    if (s == 123.456)
        out[0] = s;
}

double tvd_fp64_probe(int iters, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d_out = nullptr;
    if (cudaMalloc(&d_out, sizeof(double)) != cudaSuccess)
        return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256;
    dfma_probe<<<blocks, threads, 0, stream>>>(d_out, 64, 0.999999, 1e-9);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, stream);
        dfma_probe<<<blocks, threads, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16.0 * (double)((iters + 15) / 16 * 16) * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best)
            best = tf;
    }
    tvd_count_launch(6);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

// ---------------------------------------------------------------------------------------------
// debug: the sweep's own cos/sin, exposed so that tests can measure them against glibc
// ---------------------------------------------------------------------------------------------
__global__ void debug_trig_kernel(int64_t n, const double *__restrict__ x, double *__restrict__ c,
                                  double *__restrict__ s)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double xi = x[i < n ? i : n - 1];
    double c0, c1, s0, s1;
    trig2<true>(xi, xi, c0, c1, s0, s1);
    if (i < n) {
        c[i] = c0;
        s[i] = s0;
    }
}

cudaError_t tvd_debug_trig_launch(int64_t n, const double *d_x, double *d_c, double *d_s,
                                  cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    debug_trig_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, d_x, d_c, d_s);
    tvd_count_launch();
    return cudaGetLastError();
}
