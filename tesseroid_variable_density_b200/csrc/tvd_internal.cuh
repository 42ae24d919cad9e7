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
#define TVD_TEST_REC 4
#define TVD_GLQ_REC 16
#define TVD_TILE 64                         // pieces per shared-memory tile
// far-field sweep launch shape: threads per block, points per thread, resident blocks per SM
#ifndef TVD_FAR_THREADS
#define TVD_FAR_THREADS 512
#endif
#ifndef TVD_FAR_PPT
#define TVD_FAR_PPT 1
#endif
#ifndef TVD_FAR_MINBLOCKS
#define TVD_FAR_MINBLOCKS 1
#endif
#define TVD_MAX_K1_PIECES 256               // cap on radial pieces of ONE tesseroid
#define TVD_MAX_DEPTH 64                    // levels of the near-field subdivision stack

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
        return p[0] * sin(p[1] * h) + p[2];
    }
}

// canonical piece table (device, SoA), output of the K1 preprocessing
struct PieceTable {
    int64_t n_tess = 0, n_pieces = 0;
    double *tess_bounds = nullptr;   // [n_tess][6]
    int32_t *tess_kind = nullptr;    // [n_tess]
    double *tess_params = nullptr;   // [n_tess][TVD_NPAR]
    int32_t *tess_npieces = nullptr; // [n_tess]
    double *top = nullptr;           // [n_pieces]
    double *bottom = nullptr;        // [n_pieces]
    int32_t *parent = nullptr;       // [n_pieces]
};

// ---- launchers implemented in the .cu files (all asynchronous on `stream`) -------------------
// K1: density-based radial discretisation.  counts pass then emit pass.
cudaError_t tvd_launch_k1_count(const PieceTable &t, double delta, int has_delta, int *d_status,
                                cudaStream_t stream);
cudaError_t tvd_launch_scan(const int32_t *d_counts, int64_t *d_offsets, int64_t n,
                            cudaStream_t stream);   // exclusive scan, d_offsets[n] = total
cudaError_t tvd_launch_k1_emit(const PieceTable &t, const int64_t *d_offsets, double delta,
                               int has_delta, cudaStream_t stream);
// K4: per-piece root records for the far-field sweep of one (field, ratio)
cudaError_t tvd_launch_pack(int field, const PieceTable &t, double ratio, double *d_test,
                            double *d_glq, cudaStream_t stream);
// Phase A: far-field sweep
struct FarArgs {
    int field;
    int64_t n_points;
    const double *lon, *sinlat, *coslat, *radius;
    int64_t n_pieces;
    const double *test_rec, *glq_rec;
    int n_groups;
    int64_t pieces_per_group;
    double *far_out;                 // [n_groups][n_points]
    unsigned long long *queue;       // near-field pairs: (point << 32) | piece
    unsigned long long *qcount;
    unsigned long long qcap;
};
cudaError_t tvd_launch_far(const FarArgs &a, cudaStream_t stream);
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
    unsigned long long *counters;    // nodes, leaves, splits, too_small, overflow flag
    int32_t *leaf_counts;            // NULL or [n_tess][n_points]
    int64_t n_points;
};
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

void tvd_count_launch(int n = 1);
