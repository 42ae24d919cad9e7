/*
 * tvd.h -- C ABI of libtvd.so, the B200 (sm_100a) implementation of the
 * tesseroid_density forward-modelling hot path.
 *
 * Plain C, plain pointers and sizes, no C++/torch types, no exceptions across the boundary.
 * Every entry point cites the reference interface it replaces (paths relative to
 * pinga-lab/tesseroid-variable-density).
 *
 * The reference's native boundary is ONE CALL PER RADIAL PIECE,
 *     _tesseroid.<field>(bounds, density, ratio, STACK_SIZE, lons, sinlats, coslats, radii, result)
 *     (code/tesseroid_density/_tesseroid.pyx:204-213, 282-291, 362-371, 442-451)
 * driven by the Python loop of _forward_model (code/tesseroid_density/tesseroid.py:198-234).
 * That granularity is wrong for a GPU, so the ABI is batched: one call preprocesses the whole
 * model (tesseroid.py:217-223, :237-300) and one call evaluates a field of the whole model on
 * all points (tesseroid.py:225-233 + _tesseroid.pyx:34-98).
 *
 * Conventions
 *   - all floating point is IEEE float64; angles of tesseroid bounds in DEGREES, point
 *     longitudes in RADIANS, exactly what the reference passes at its own boundary
 *     (tesseroid.py:209-211, _convert_coords :125-139);
 *   - results are in "kernel units" (field / G, SI), exactly what the reference's Cython
 *     functions accumulate into `result`; the unit scaling by G, SI2MGAL, SI2EOTVOS stays in
 *     the Python layer (tesseroid.py:381,439,497,560);
 *   - return value: 0 ok; TVD_ERR_STACK_OVERFLOW (-1) mirrors
 *     ValueError('Tesseroid stack overflow') (_tesseroid.pyx:88-89); other negative values are
 *     argument errors; positive values are CUDA runtime errors (cudaError_t).  The text of the
 *     last error of the calling thread is returned by tvd_last_error();
 *   - host buffers are caller-owned, C-contiguous, never retained after return;
 *   - the library owns all device memory; the model handle is freed by tvd_model_free();
 *   - calls are synchronous and thread-safe for distinct handles / outputs.
 */
#ifndef TVD_H
#define TVD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Fields.  0..3 exist in the reference (_tesseroid.pyx:204,282,362,442); 4..9 are the gradient
 * tensor listed in the reference docstring (tesseroid.py:56-65) but only implemented in its
 * upstream, Fatiando 0.5. */
enum tvd_field {
    TVD_POTENTIAL = 0, TVD_GX = 1, TVD_GY = 2, TVD_GZ = 3,
    TVD_GXX = 4, TVD_GXY = 5, TVD_GXZ = 6, TVD_GYY = 7, TVD_GYZ = 8, TVD_GZZ = 9,
    TVD_NFIELDS = 10
};

/* Density-in-depth models; they replace the reference's Python callables `density(height)`
 * (_tesseroid.pyx:274,354,434,515), which cannot run on a GPU.  h = height in metres.
 *   CONST   rho = p0                              plain number      (_tesseroid.pyx:218-224)
 *   LINEAR  rho = p0*(h + p1) + p2                code/scripts/linear_density.py:73-75
 *   EXP     rho = p0*exp(p1*(h - p2)/p3) + p4     code/scripts/exponential_density.py:97-98
 *   SINE    rho = p0*sin(p1*h) + p2               code/scripts/sine_density.py:93-94      */
enum tvd_density_kind { TVD_DENS_CONST = 0, TVD_DENS_LINEAR = 1, TVD_DENS_EXP = 2,
                        TVD_DENS_SINE = 3 };
#define TVD_NPAR 5

/* stats[] slots filled by tvd_forward*.  "node" = one pop of the reference's subdivision
 * stack (_tesseroid.pyx:74-81), "leaf" = a node that is integrated (:91-96). */
enum tvd_stat {
    TVD_ST_NODES = 0,          /* stack pops                                              */
    TVD_ST_LEAVES = 1,         /* GLQ evaluations                                         */
    TVD_ST_SPLITS = 2,         /* nodes that were split (:87-90)                          */
    TVD_ST_TOO_SMALL = 3,      /* "-error_code": cells that should split but are <=0.1 m  */
                               /*   (:136-137,141-142) -> RuntimeWarning in Python        */
    TVD_ST_PIECES = 4,         /* radial pieces of the model (Q)                          */
    TVD_ST_NONROOT_NODES = 5,
    TVD_ST_NONROOT_LEAVES = 6,
    TVD_ST_NEAR_PAIRS = 7,     /* (piece, point) pairs that left the fast far-field path  */
    TVD_NSTATS = 8
};

#define TVD_OK 0
#define TVD_ERR_STACK_OVERFLOW (-1)
#define TVD_ERR_BAD_ARGUMENT (-2)
#define TVD_ERR_TOO_MANY_PIECES (-3)
#define TVD_ERR_NO_DEVICE (-4)

typedef struct tvd_model tvd_model;   /* opaque: flattened radial-piece table, per device */

/* Number of visible CUDA devices (0 if none). */
int tvd_device_count(void);

/* Create the CUDA contexts of the given devices (NULL = 0..n_devices-1) concurrently.  Optional:
 * contexts are otherwise created by the first call that uses a device; a caller that knows its
 * devices can overlap this with its own preparation (the Python layer does so while it flattens
 * the model, the counterpart of the reference forking its pool, tesseroid.py:180). */
int tvd_warm_devices(int n_devices, const int *device_ids);

/* Text of the last error raised on the calling thread ("" if none). */
const char *tvd_last_error(void);

/*
 * Build the model: density-based radial discretisation of every tesseroid on the GPU and
 * the flattened piece table.
 * Replaces: the per-tesseroid part of _forward_model (tesseroid.py:217-224) and
 *           _density_based_discretization / _divider_calculation (tesseroid.py:237-300).
 *
 *   n_tess     number of tesseroids (already filtered as _check_tesseroid does, :142-161)
 *   bounds     [n_tess][6]  w, e, s, n, top, bottom  (np.array(tesseroid.get_bounds()), :221)
 *   kind       [n_tess]     tvd_density_kind
 *   params     [n_tess][TVD_NPAR]
 *   delta      density-variation threshold; NaN means delta=None (:222): no radial splitting
 *   device     CUDA device that runs the preprocessing (the table is replicated lazily to the
 *              other devices used by tvd_forward)
 *   out        receives the handle
 */
int tvd_model_create(int64_t n_tess, const double *bounds, const int32_t *kind,
                     const double *params, double delta, int device, tvd_model **out);

void tvd_model_free(tvd_model *model);

/* Number of radial pieces Q (>= n_tess). */
int64_t tvd_model_num_pieces(const tvd_model *model);

/*
 * Copy out the flattened split list: for piece q, its top/bottom and the index of the input
 * tesseroid it came from.  Pieces are in model order; the pieces of one tesseroid are in the
 * reference's FIFO emission order (tesseroid.py:267-282).  Any pointer may be NULL.
 * This is what _density_based_discretization(bounds, density, delta) returns (:237).
 */
int tvd_model_get_pieces(const tvd_model *model, double *top, double *bottom, int32_t *parent);

/*
 * Evaluate one field of the whole model on all points (HOST buffers; host<->device copies
 * inside the call).  Replaces: the piece loop of _forward_model (tesseroid.py:225-233) and
 * rediscretizer + kernels (_tesseroid.pyx:34-98, 231-520), and the njobs point chunking
 * (tesseroid.py:179-195, 303-324) when n_devices > 1.
 *
 *   field        tvd_field
 *   n_points     P
 *   lon_rad, sinlat, coslat, radius   [P] as produced by _convert_coords (tesseroid.py:125-139)
 *   ratio        distance-size ratio D (> 0)
 *   stack_size   the reference's STACK_SIZE (tesseroid.py:102); only used to reproduce the
 *                overflow error
 *   n_devices, device_ids   points are split into n_devices contiguous chunks exactly like
 *                _split_arrays (tesseroid.py:318-323); device_ids==NULL means 0..n_devices-1.
 *                A device works through its chunk in slabs of at most 2^24 points (the reference
 *                streams points one at a time, _tesseroid.pyx:65), so device memory does not grow
 *                with P; chunking and slabbing never change a bit of the result
 *   result       [P] OVERWRITTEN with the field in kernel units (the reference accumulates into
 *                a zero-filled array, tesseroid.py:121)
 *   stats        NULL or int64[TVD_NSTATS]
 *   leaf_counts  NULL or int32[n_tess][P]: GLQ evaluations per (input tesseroid, point), summed
 *                over the tesseroid's radial pieces (debug / parity; small problems only)
 */
int tvd_forward(int field, tvd_model *model, int64_t n_points, const double *lon_rad,
                const double *sinlat, const double *coslat, const double *radius,
                double ratio, int stack_size, int n_devices, const int *device_ids,
                double *result, int64_t *stats, int32_t *leaf_counts);

/*
 * Several fields of the same model on the same points in ONE far-field sweep:
 * the fields share the far test, the angular half of the quadrature, l^2 and its inverse square
 * root; each field keeps its own distance-size ratio, near-field queue and literal walk; its
 * subdivision counts are those of a separate tvd_forward call and its values agree with it to a
 * few ulp (bit-identical for {gx, gy, gz} and for the six tensor components).  The reference needs one full pass
 * per field (tesseroid.py:377,435,493,556).
 *
 *   n_fields, fields, ratios   distinct tvd_field ids and the ratio D of each
 *   results      [n_fields][P] in the caller's field order, kernel units, OVERWRITTEN
 *   stats        NULL or int64[n_fields][TVD_NSTATS]
 * Compiled-in sets: any single field; {potential, gz}; {potential, gz, gzz};
 * {potential, gx, gy, gz};
 * {gx, gy, gz}; the six tensor components.  Other sets return TVD_ERR_BAD_ARGUMENT
 * (tvd_fused_supported tells in advance).
 */
int tvd_forward_fused(int n_fields, const int *fields, const double *ratios, tvd_model *model,
                      int64_t n_points, const double *lon_rad, const double *sinlat,
                      const double *coslat, const double *radius, int stack_size, int n_devices,
                      const int *device_ids, double *results, int64_t *stats);

/*
 * Sensitivity (Jacobian) mode: the contribution of every input tesseroid to every point,
 * jac[t][p], instead of their sum.  The reference builds such matrices one
 * column at a time with `field(lon, lat, height, [tesseroid_t], dens=1)`
 * ("Use this, e.g., for sensitivity matrix building", tesseroid.py:347-349); row t here is
 * bit-identical to that call.  Single device; jac is a HOST buffer [n_tess][n_points], kernel
 * units.  The points are processed in slabs so that the device holds at most ~2 GiB of rows
 * ((n_pieces + n_tess) * slab * 8 bytes).
 */
int tvd_sensitivity(int field, tvd_model *model, int64_t n_points, const double *lon_rad,
                    const double *sinlat, const double *coslat, const double *radius,
                    double ratio, int stack_size, int device, double *jac, int64_t *stats);

/* 1 if tvd_forward_fused accepts this set of fields, else 0. */
int tvd_fused_supported(int n_fields, const int *fields);

/*
 * Same computation with DEVICE-resident point arrays and result on `device`, enqueued on
 * `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).  Used by the
 * one-process-per-GPU launcher (each rank owns a contiguous point shard) and by bench.py for
 * the kernel-only number.  `n_groups` fixes how the piece table is partitioned into
 * independently summed groups (0 = choose from n_points); ranks that must produce results
 * bit-identical to a single-device run pass the same n_groups.
 * The call returns after the stream has drained.
 */
int tvd_forward_device(int field, tvd_model *model, int device, int64_t n_points,
                       const double *d_lon_rad, const double *d_sinlat, const double *d_coslat,
                       const double *d_radius, double ratio, int stack_size, int n_groups,
                       double *d_result, int64_t *stats, void *stream);

/* tvd_forward_fused with DEVICE-resident points and results on `device` (the rank-per-GPU form of
 * the fused sweep): d_results[i] is the device array of the caller's i-th field. */
int tvd_forward_fused_device(int n_fields, const int *fields, const double *ratios,
                             tvd_model *model, int device, int64_t n_points,
                             const double *d_lon_rad, const double *d_sinlat,
                             const double *d_coslat, const double *d_radius, int stack_size,
                             int n_groups, double *const *d_results, int64_t *stats, void *stream);

/* The n_groups that tvd_forward would choose for a job of n_points_total points over this model
 * (pure function of its arguments; lets independent ranks agree on the summation plan). */
int tvd_plan_groups(const tvd_model *model, int64_t n_points_total);

/* Kernel launches issued by the library on the calling thread since the last call to this
 * function with reset != 0 (bench.py's "gpu_launches"). */
int64_t tvd_launch_count(int reset);

/* Device-time (ms, CUDA events on the launching stream) of the dominant kernel (far-field
 * sweep) accumulated since the last reset, and the number of launches it covers. */
double tvd_far_kernel_ms(int reset, int64_t *n_launches);

/* Device time (ms, CUDA events on the launching stream) by phase, accumulated since the last
 * reset: out[0] far-field sweeps, out[1] near-field walks (rediscretizer proper,
 * _tesseroid.pyx:34-98), out[2] everything enqueued after the sweep (key sort, walks, combine),
 * out[3] number of sweeps.  tvd_far_kernel_ms resets the same counters. */
void tvd_phase_ms(int reset, double *out);

/* FP64 peak probe: runs a dependent-chain-free DFMA loop on `device` and returns the measured
 * FP64 rate in TFLOP/s (FMA = 2 flops).  Roofline denominator, see DESIGN.md. */
double tvd_measure_fp64_peak(int device, int iters);

/* Debug/test hook: the far-field sweep's own cos/sin of the longitude difference evaluated on
 * `device` for n host values (tests measure them against glibc, which the reference calls at
 * _tesseroid.pyx:242,401). */
int tvd_debug_trig(int device, int64_t n, const double *x, double *cos_out, double *sin_out);

/* Debug/test hook: the double-double sin (func 0), cos (1) and exp (2) that K1 uses for the
 * density samples (tesseroid.py:259,294), and the table-based correctly rounded sin (3), cos (4)
 * and sincos (5: its sine, 6: its cosine) of the near-field walk (the values glibc returned to
 * the reference at _tesseroid.pyx:113-121,168-169,242) and the table-based correctly rounded
 * exp (7) of K1/K4, evaluated on `device` for n host values. */
int tvd_debug_math(int device, int func, int64_t n, const double *x, double *out);

#ifdef __cplusplus
}
#endif
#endif /* TVD_H */
