// tvd_internal.cuh -- shared declarations of the libtvd translation units (sm_100a only).
//
// All translation units are compiled with -fmad=false: the reference binary (x86-64, no FMA
// contraction) rounds every product and sum separately, and the ill-conditioned expressions
// (cos psi, l^2) must round the same way.  Where a fused multiply-add is
// wanted it is written explicitly as fma().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tvd.h"

#define TVD_R 6378137.0                     // fatiando.constants.MEAN_EARTH_RADIUS (_tesseroid.pyx:23)
#define TVD_D2R 0.017453292519943295        // numpy.pi/180. (_tesseroid.pyx:21)
#define TVD_NODE 0.577350269189625731058868041146   // _tesseroid.pyx:27-28

// Far-field record sizes (doubles per radial piece); see DESIGN.md "data layout".
// test record: cx cy cz K_0 .. K_{nf-1}, padded to an even number of doubles (16-byte records)
__host__ __device__ constexpr int tvd_test_rec_len(int nf) { return ((3 + nf) + 1) / 2 * 2; }
#define TVD_GLQ_REC 16
#define TVD_MAX_FUSED 6                     // fields per fused sweep
// fused field sets the sweep is compiled for (besides the ten single fields)
#define TVD_MASK_V_GZ ((1 << TVD_POTENTIAL) | (1 << TVD_GZ))
#define TVD_MASK_V_GZ_GZZ ((1 << TVD_POTENTIAL) | (1 << TVD_GZ) | (1 << TVD_GZZ))
#define TVD_MASK_V_G ((1 << TVD_POTENTIAL) | (1 << TVD_GX) | (1 << TVD_GY) | (1 << TVD_GZ))
#define TVD_MASK_G ((1 << TVD_GX) | (1 << TVD_GY) | (1 << TVD_GZ))
#define TVD_MASK_TENSOR ((1 << TVD_GXX) | (1 << TVD_GXY) | (1 << TVD_GXZ) | (1 << TVD_GYY) | \
                         (1 << TVD_GYZ) | (1 << TVD_GZZ))
#ifndef TVD_TILE
#define TVD_TILE 96                         // pieces per shared-memory tile (static smem <= 48 KB for 6 fused fields)
#endif
// far-field sweep launch shape: threads (= points) per block, resident blocks per SM
#ifndef TVD_FAR_THREADS
#define TVD_FAR_THREADS 512
#endif
#ifndef TVD_FAR_MINBLOCKS
#define TVD_FAR_MINBLOCKS 1
#endif
#define TVD_MAX_K1_PIECES 256               // cap on radial pieces of ONE tesseroid
#define TVD_MAX_DEPTH 64                    // levels of the near-field subdivision stack

// ---------------------------------------------------------------------------------------------
// sin / cos.  fdlibm kernels + Cody-Waite reduction by pi/2 with a 33-bit head (k*head exact)
// keeping the reduction tail.  Error < 1 ulp, and equal to glibc's sin/cos -- what the
// reference calls -- in 97-99 % of cases; libdevice's versions are used only for huge arguments.
// ---------------------------------------------------------------------------------------------
#define TVD_PIO4 0.78539816339744830962
__device__ __forceinline__ double kcos(double x, double y)
{
    const double z = x * x;
    double r = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    r = fma(z, r, -2.75573143513906633035e-07);
    r = fma(z, r, 2.48015872894767294178e-05);
    r = fma(z, r, -1.38888888888741095749e-03);
    r = fma(z, r, 4.16666666666666019037e-02);
    r = r * z;
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + fma(z, r, -(x * y)));
}
__device__ __forceinline__ double kcos0(double x)      // tail y == 0
{
    const double z = x * x;
    double r = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    r = fma(z, r, -2.75573143513906633035e-07);
    r = fma(z, r, 2.48015872894767294178e-05);
    r = fma(z, r, -1.38888888888741095749e-03);
    r = fma(z, r, 4.16666666666666019037e-02);
    r = r * z;
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    return w + fma(z, r, (1.0 - w) - hz);
}
__device__ __forceinline__ double ksin(double x, double y)
{
    const double z = x * x;
    const double v = z * x;
    double r = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    r = fma(z, r, 2.75573137070700676789e-06);
    r = fma(z, r, -1.98412698298579493134e-04);
    r = fma(z, r, 8.33333333332248946124e-03);
    // x - ((z*(0.5*y - v*r) - y) - v*S1)
    return x - ((z * fma(-v, r, 0.5 * y) - y) - v * -1.66666666666666324348e-01);
}
__device__ __forceinline__ double ksin0(double x)
{
    const double z = x * x;
    const double v = z * x;
    double r = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    r = fma(z, r, 2.75573137070700676789e-06);
    r = fma(z, r, -1.98412698298579493134e-04);
    r = fma(z, r, 8.33333333332248946124e-03);
    r = fma(z, r, -1.66666666666666324348e-01);
    return fma(v, r, x);
}
// general argument: returns quadrant, reduced argument r and its tail y (x = k*pi/2 + r + y)
__device__ __forceinline__ int reduce_pio2(double x, double &r, double &y)
{
    const double t = fma(x, 6.36619772367581382433e-01, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    const double r0 = fma(-kd, 1.57079632673412561417e+00, x);    // exact
    const double w = kd * 6.07710050650619224932e-11;
    r = r0 - w;
    y = (r0 - r) - w;
    return k;
}
// scalar sin and cos of one argument (per-thread, no warp votes): K4, Phase B, densities
__device__ __forceinline__ void tvd_sincos(double x, double &s, double &c)
{
    if (!(fabs(x) < 1.0e5)) {       // huge, inf or NaN
        sincos(x, &s, &c);
        return;
    }
    if (fabs(x) <= TVD_PIO4) {
        s = ksin0(x);
        c = kcos0(x);
        return;
    }
    double r, y;
    const int k = reduce_pio2(x, r, y);
    const double cr = kcos(r, y), sr = ksin(r, y);
    c = (k & 1) ? sr : cr;
    if ((k + 1) & 2)
        c = -c;
    s = (k & 1) ? cr : sr;
    if (k & 2)
        s = -s;
}
__device__ __forceinline__ double tvd_sin(double x)
{
    double s, c;
    tvd_sincos(x, s, c);
    return s;
}
__device__ __forceinline__ double tvd_cos(double x)
{
    double s, c;
    tvd_sincos(x, s, c);
    return c;
}

// ---------------------------------------------------------------------------------------------
// FP64 building blocks
// ---------------------------------------------------------------------------------------------
// 1/sqrt(x) to ~1 ulp: MUFU.RSQ64H seed (>= 20 bits) + one third-order correction
//   e = 1 - x*y0^2 ;  y = y0*(1 + e/2 + 3e^2/8)        (error ~ (5/16) e^3 < 2^-60)
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double t = x * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = e * p;
    return fma(y0, q, y0);
}

// x^(-3/2) directly from the same seed (one operation fewer than cubing rsqrt_fast):
//   e = 1 - x*y0^2 ;  x^(-3/2) = y0^3 (1 - e)^(-3/2) = y0^3 (1 + 3e/2 + 15e^2/8 + O(e^3))
__device__ __forceinline__ double rsqrt3_fast(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double s = y0 * y0;
    const double e = fma(-x, s, 1.0);
    const double c = s * y0;
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    return fma(c, q, c);
}

// ---------------------------------------------------------------------------------------------
// density-in-depth models (tvd_density_kind), evaluated in the reference scripts' own
// operation order (no contraction: the TU is built with -fmad=false)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double tvd_density(int kind, const double *__restrict__ p, double h)
{
    switch (kind) {
    case TVD_DENS_CONST:
        return p[0];
    case TVD_DENS_LINEAR: {
        double r = h + p[1];
        return p[0] * r + p[2];
    }
    case TVD_DENS_EXP:
        return p[0] * exp(p[1] * (h - p[2]) / p[3]) + p[4];
    default:
        return p[0] * tvd_sin(p[1] * h) + p[2];
    }
}

// The same models with sin/exp evaluated in double-double and rounded once (tvd_ddmath.cuh):
// used where the last bit decides discrete outcomes (K1) and where it is free (K4).
#include "tvd_ddmath.cuh"
#include "tvd_crtrig.cuh"
__device__ __forceinline__ double tvd_density_cr(int kind, const double *__restrict__ p, double h)
{
    switch (kind) {
    case TVD_DENS_CONST:
        return p[0];
    case TVD_DENS_LINEAR: {
        double r = h + p[1];
        return p[0] * r + p[2];
    }
    case TVD_DENS_EXP:
        return p[0] * cr_exp(p[1] * (h - p[2]) / p[3]) + p[4];     // same bits as dd_exp
    default:
        return p[0] * cr_sin(p[1] * h) + p[2];      // same bits as dd_sin (tvd_crtrig.cuh)
    }
}

// canonical piece table (device, SoA), output of the K1 preprocessing
struct PieceTable {
    int64_t n_tess = 0, n_pieces = 0;
    double *tess_bounds = nullptr;   // [n_tess][6]
    int32_t *tess_kind = nullptr;    // [n_tess]
    double *tess_params = nullptr;   // [n_tess][TVD_NPAR]
    int32_t *tess_npieces = nullptr; // [n_tess]
    int64_t *tess_offset = nullptr;  // [n_tess + 1] first piece of each tesseroid
    int32_t *sweep_piece = nullptr;  // [n_pieces] sweep position -> piece (pieces sorted by the
                                     // longitude extent of their tesseroid, see tvd_build_sweep_order)
    double *top = nullptr;           // [n_pieces]
    double *bottom = nullptr;        // [n_pieces]
    int32_t *parent = nullptr;       // [n_pieces]
};

// ---- launchers implemented in the .cu files (all asynchronous on `stream`) -------------------
// K1: density-based radial discretisation.  counts pass then emit pass.
// d_keep: scratch [n_tess][TVD_K1_KEEP][2], pieces of tesseroids with few pieces handed from the
// count pass to the emit pass
#define TVD_K1_KEEP 8
cudaError_t tvd_launch_k1_count(const PieceTable &t, double delta, int has_delta, int *d_status,
                                double *d_keep, cudaStream_t stream);
cudaError_t tvd_launch_scan(const int32_t *d_counts, int64_t *d_offsets, int64_t n,
                            cudaStream_t stream);   // exclusive scan, d_offsets[n] = total
cudaError_t tvd_launch_k1_emit(const PieceTable &t, const int64_t *d_offsets, double delta,
                               int has_delta, double *d_keep, cudaStream_t stream);
// K4: per-piece root records for the far-field sweep.  The GLQ records depend on the model
// only; the test records on the distance-size ratio of each field of the sweep.
// `period`: the "same tesseroid / same longitude extent" flags of the GLQ records are reset at
// every multiple of it (groups and tiles of the sweep start there)
cudaError_t tvd_launch_pack_glq(const PieceTable &t, int period, double *d_glq, double *d_geom,
                                cudaStream_t stream);
struct RatioList {
    int nf;
    double ratio[TVD_MAX_FUSED];
};
cudaError_t tvd_launch_pack_test(const PieceTable &t, const double *d_geom, RatioList ratios,
                                 double *d_test, cudaStream_t stream);
// per-tile record of the sweep: Cx Cy Cz |C|^2 lim_0 .. lim_5 (bounding sphere of the tile's piece
// centres; a point farther than sqrt(lim_f) from C is far from every piece of the tile for field f).
// Tiles are those of the sweep's grouping: tile k of group g covers the pieces
// [g*ppg + k*TVD_TILE, ...) and its record is number g*ceil(ppg/TVD_TILE) + k.
#define TVD_TILE_REC 10
cudaError_t tvd_launch_pack_tiles(const PieceTable &t, const double *d_geom, RatioList ratios,
                                  int64_t pieces_per_group, int n_groups, double *d_tile,
                                  cudaStream_t stream);
// points-at-one-height specialisation: per piece 2r*rc0 2r*rc1 r^2+rc0^2 r^2+rc1^2 (sweep order)
#define TVD_AB_REC 4
// ... and, for gz, the node weights times rc[k] and times the radius:
// w00*rc0 w01*rc1 w10*rc0 w11*rc1  w00*r w01*r w10*r w11*r (sweep order)
#define TVD_WZ_REC 8
cudaError_t tvd_launch_pack_ab(const PieceTable &t, const double *d_glq, double radius,
                               double *d_ab, double *d_wz, cudaStream_t stream);
// *d_flag = 1 if every radius[i] is bitwise equal to radius[0], else 0
cudaError_t tvd_launch_check_uniform(int64_t n, const double *d_radius, int *d_flag,
                                     cudaStream_t stream);
// Phase A: far-field sweep
bool tvd_far_mask_supported(int mask);
struct FarArgs {
    int mask;                        // bit f set: tvd_field f is computed
    int64_t n_points;
    const double *lon, *sinlat, *coslat, *radius;
    int64_t n_pieces;
    const double *test_rec, *glq_rec, *tile_rec;
    const double *ab_rec;            // non-NULL: all points share one radius (see TVD_AB_REC)
    const double *wz_rec;            // with ab_rec: TVD_WZ_REC
    int n_groups;
    int64_t pieces_per_group;
    int tiles_per_group;             // ceil(pieces_per_group / TVD_TILE)
    double *far_out;                 // [nf][n_groups][n_points]
    unsigned long long *queue;       // [nf][qcap] near-field pairs: (point << 32) | piece
    unsigned long long *qcount;      // [nf]
    unsigned long long qcap;
    double *jac;                     // sensitivity mode (single field): [n_pieces][n_points],
                                     // NaN marks a deferred pair; far_out is not written
};
cudaError_t tvd_launch_far(const FarArgs &a, cudaStream_t stream);
cudaError_t tvd_launch_queue_only(const FarArgs &a, cudaStream_t stream);   // refill after overflow
// sensitivity mode: rows of pieces -> rows of tesseroids, near-field terms added in piece order
cudaError_t tvd_launch_jac_reduce(const PieceTable &t, int64_t n_points, const double *jac_pieces,
                                  int64_t n_pairs, const unsigned long long *keys,
                                  const double *pair_result, double *jac_out, cudaStream_t stream);
// Order in which the far-field sweep visits the pieces: tesseroids stably sorted by their
// longitude extent (w, e), each followed by its radial pieces.  Fills t.sweep_piece.
cudaError_t tvd_build_sweep_order(PieceTable &t, cudaStream_t stream);
cudaError_t tvd_sort_pairs(void *d_temp, size_t &temp_bytes, const unsigned long long *keys_in,
                           unsigned long long *keys_out, const int32_t *vals_in,
                           int32_t *vals_out, int64_t n, cudaStream_t stream);
// sort of the near-field pair keys (CUB radix sort; bookkeeping, tvd_sort.cu)
cudaError_t tvd_sort_keys(void *d_temp, size_t &temp_bytes, const unsigned long long *in,
                          unsigned long long *out, int64_t n, int end_bit, cudaStream_t stream);
// Phase B: near-field pairs, literal subdivision walk
struct NearArgs {
    int field;
    int64_t n_pairs;
    const unsigned long long *keys;  // sorted
    const double *lon, *sinlat, *coslat, *radius;
    PieceTable table;
    double ratio;
    int stack_size;
    double *pair_result;             // [n_pairs]
    unsigned long long *counters;    // nodes, leaves, splits, too_small, overflow flag,
                                     // root-leaf pairs, [6] = work counter (zeroed by the caller)
    int32_t *leaf_counts;            // NULL or [n_tess][n_points]
    int64_t n_points;
    int batch;                       // pairs per work grab (set by the launcher)
    int shallow;                     // root kernel also finishes barely-near pairs (set by the launcher)
    void *workspace;                 // tvd_near_workspace_bytes(n_pairs) bytes
    // carved out of the workspace by the launcher
    unsigned char *root_code, *root_class;
    int32_t *order;
    unsigned *bins;
};
size_t tvd_near_workspace_bytes(int64_t n_pairs);
cudaError_t tvd_launch_near(const NearArgs &a, cudaStream_t stream);
// Phase C: result[p] = sum_g far[g][p] + sum of p's near-field pair results (key order)
cudaError_t tvd_launch_combine(int64_t n_points, int n_groups, const double *far_out,
                               int64_t n_pairs, const unsigned long long *keys,
                               const double *pair_result, double *result, cudaStream_t stream);
cudaError_t tvd_launch_init_counts(const PieceTable &t, int64_t n_points, int32_t *counts,
                                   cudaStream_t stream);
double tvd_fp64_probe(int iters, cudaStream_t stream);
cudaError_t tvd_debug_trig_launch(int64_t n, const double *d_x, double *d_c, double *d_s,
                                  cudaStream_t stream);
cudaError_t tvd_debug_math_launch(int func, int64_t n, const double *d_x, double *d_out,
                                  cudaStream_t stream);

void tvd_count_launch(int n = 1);
