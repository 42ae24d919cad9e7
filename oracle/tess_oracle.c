/*
 * tess_oracle.c -- CPU restatement of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package links, loads or calls this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, and there only as the checker / reported baseline.
 *
 * It restates, literally and in the reference's own operation order (plain C, glibc libm,
 * no FMA contraction: compile WITHOUT -march=native / -mfma / -ffast-math), the algorithm of
 * pinga-lab/tesseroid-variable-density:
 *
 *   rediscretizer (points x LIFO stack)          code/tesseroid_density/_tesseroid.pyx:34-98
 *   distance_n_size                              _tesseroid.pyx:104-123
 *   divisions                                    _tesseroid.pyx:129-146
 *   scale_nodes                                  _tesseroid.pyx:152-176
 *   split                                        _tesseroid.pyx:182-198
 *   kernelV/x/y/z (+ _variable)                  _tesseroid.pyx:231-520
 *   _forward_model (model loop, piece order)     code/tesseroid_density/tesseroid.py:198-234
 *   _density_based_discretization                tesseroid.py:237-283
 *   _divider_calculation                         tesseroid.py:286-300
 *   _split_arrays (point chunking, njobs)        tesseroid.py:303-324
 *
 * The gradient-tensor kernels (fields 4..9) do not exist in the reference; they follow the
 * integrands of Fatiando 0.5 (un-vendored dependency, environment.yml:18) as listed in
 * SURVEY.md section 8 a15 and are pinned only by the analytical shell solutions
 * (code/scripts/linear_density.py:25-30): PARITY UNPINNED for fields 4..9.
 *
 * Pinned (tests/test_oracle_*.py) against the reference's stored golden results and against the
 * unmodified reference compiled into oracle/_ref by oracle/build_ref.py.
 *
 * The Python density callables of the reference scripts become parametric density kinds:
 *   0 CONST   rho = p0                                     (plain number, _tesseroid.pyx:218-224)
 *   1 LINEAR  rho(h) = p0*(h + p1) + p2                    (linear_density.py:73-75, neuquen_basin.py:119-121)
 *   2 EXP     rho(h) = p0*exp(p1*(h - p2)/p3) + p4         (exponential_density.py:97-98, neuquen_basin.py:142-147)
 *   3 SINE    rho(h) = p0*sin(p1*h) + p2                   (sine_density.py:93-94)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_NPAR 5
#define MEAN_EARTH_RADIUS 6378137.0

static const double d2r = 3.14159265358979323846 / 180.;
static const double nodes[2] = {-0.577350269189625731058868041146,
                                0.577350269189625731058868041146};

enum { F_POT = 0, F_GX, F_GY, F_GZ, F_GXX, F_GXY, F_GXZ, F_GYY, F_GYZ, F_GZZ, F_COUNT };
enum { D_CONST = 0, D_LINEAR, D_EXP, D_SINE };

/* stats layout shared with include/tvd.h */
enum { ST_NODES = 0, ST_LEAVES, ST_SPLITS, ST_TOO_SMALL, ST_PIECES, ST_NONROOT_NODES,
       ST_NONROOT_LEAVES, ST_RESERVED, ST_COUNT };

double orc_density_eval(int kind, const double *p, double h)
{
    switch (kind) {
    case D_CONST:
        return p[0];
    case D_LINEAR: {
        double r = h + p[1];
        return p[0] * r + p[2];
    }
    case D_EXP:
        return p[0] * exp(p[1] * (h - p[2]) / p[3]) + p[4];
    case D_SINE:
        return p[0] * sin(p[1] * h) + p[2];
    }
    return NAN;
}

/* numpy.linspace(start, stop, 101): y[i] = i*step + start, last sample forced to stop;
 * step == 0 -> (i/100)*delta + start.  tesseroid.py:258,293 */
static void linspace101(double start, double stop, double *y)
{
    double delta = stop - start;
    double step = delta / 100.;
    int i;
    if (step == 0) {
        for (i = 0; i < 101; i++)
            y[i] = (i / 100.) * delta + start;
    } else {
        for (i = 0; i < 101; i++)
            y[i] = i * step + start;
    }
    y[100] = stop;
}

/* numpy.isclose(a, b) with the default rtol=1e-5, atol=1e-8 (tesseroid.py:262) */
static int isclose_default(double a, double b)
{
    if (isfinite(a) && isfinite(b))
        return fabs(a - b) <= (1e-8 + 1e-5 * fabs(b));
    return a == b;
}

/* tesseroid.py:286-300 */
static void divider_calculation(double top, double bottom, int kind, const double *p,
                                double rho_min, double rho_max, double *divider,
                                double *max_diff)
{
    double heights[101], nd[101];
    double slope, line, diff, best = -1.0;
    int i, arg = 0, has_nan = 0;
    linspace101(bottom, top, heights);
    for (i = 0; i < 101; i++)
        nd[i] = (orc_density_eval(kind, p, heights[i]) - rho_min) / (rho_max - rho_min);
    slope = (nd[100] - nd[0]) / (top - bottom);
    for (i = 0; i < 101; i++) {
        line = slope * (heights[i] - bottom) + nd[0];
        diff = fabs(nd[i] - line);
        if (isnan(diff)) {
            if (!has_nan) {     /* numpy max/argmax propagate the first NaN */
                has_nan = 1;
                arg = i;
            }
        } else if (!has_nan && diff > best) {
            best = diff;
            arg = i;
        }
    }
    *max_diff = has_nan ? NAN : best;
    *divider = heights[arg];
}

/* tesseroid.py:237-283.  Emits (top, bottom) of the radial pieces in the reference's FIFO
 * emission order.  Returns the number of pieces, or -1 if it exceeds max_out. */
int orc_density_discretize(double top, double bottom, int kind, const double *p, double delta,
                           int max_out, double *out_top, double *out_bottom)
{
    double heights[101];
    double rho_min, rho_max, rho;
    double tesseroid_size = top - bottom;
    double *qt, *qb;
    int head = 0, tail = 0, cap = 64, nout = 0, i, has_nan = 0;

    linspace101(bottom, top, heights);
    rho_min = rho_max = orc_density_eval(kind, p, heights[0]);
    for (i = 0; i < 101; i++) {
        rho = orc_density_eval(kind, p, heights[i]);
        if (isnan(rho))
            has_nan = 1;
        if (rho < rho_min)
            rho_min = rho;
        if (rho > rho_max)
            rho_max = rho;
    }
    if (has_nan)
        rho_min = rho_max = NAN;
    if (isclose_default(rho_min, rho_max)) {
        if (max_out < 1)
            return -1;
        out_top[0] = top;
        out_bottom[0] = bottom;
        return 1;
    }
    qt = (double *)malloc(sizeof(double) * cap);
    qb = (double *)malloc(sizeof(double) * cap);
    qt[tail] = top;
    qb[tail] = bottom;
    tail++;
    while (head < tail) {
        double t = qt[head], b = qb[head], divider, max_diff, size_ratio;
        head++;
        divider_calculation(t, b, kind, p, rho_min, rho_max, &divider, &max_diff);
        size_ratio = (t - b) / tesseroid_size;
        if (max_diff * size_ratio > delta) {
            if (tail + 2 > cap) {
                cap *= 2;
                qt = (double *)realloc(qt, sizeof(double) * cap);
                qb = (double *)realloc(qb, sizeof(double) * cap);
            }
            qt[tail] = t;
            qb[tail] = divider;
            tail++;
            qt[tail] = divider;
            qb[tail] = b;
            tail++;
        } else {
            if (nout >= max_out) {
                free(qt);
                free(qb);
                return -1;
            }
            out_top[nout] = t;
            out_bottom[nout] = b;
            nout++;
        }
    }
    free(qt);
    free(qb);
    return nout;
}

/* _tesseroid.pyx:104-123 */
static inline double distance_n_size(double w, double e, double s, double n, double top,
                                     double bottom, double lon, double sinlat, double coslat,
                                     double radius, double *Llon, double *Llat, double *Lr)
{
    double rt, rtop, lont, latt, sinlatt, coslatt, cospsi, distance;
    rt = 0.5 * (top + bottom) + MEAN_EARTH_RADIUS;
    lont = d2r * 0.5 * (w + e);
    latt = d2r * 0.5 * (s + n);
    sinlatt = sin(latt);
    coslatt = cos(latt);
    cospsi = sinlat * sinlatt + coslat * coslatt * cos(lon - lont);
    distance = sqrt(radius * radius + rt * rt - 2 * radius * rt * cospsi);
    rtop = top + MEAN_EARTH_RADIUS;
    *Llon = rtop * acos(sinlatt * sinlatt + (coslatt * coslatt) * cos(d2r * (e - w)));
    *Llat = rtop * acos(sin(d2r * n) * sin(d2r * s) + cos(d2r * n) * cos(d2r * s));
    *Lr = top - bottom;
    return distance;
}

/* _tesseroid.pyx:129-146 */
static inline int divisions(double distance, double Llon, double Llat, double Lr, double ratio,
                            int *nlon, int *nlat, int *new_cells)
{
    int error = 0;
    (void)Lr;
    *nlon = 1;
    *nlat = 1;
    if (distance <= ratio * Llon) {
        if (Llon <= 0.1)
            error = -1;
        else
            *nlon = 2;
    }
    if (distance <= ratio * Llat) {
        if (Llat <= 0.1)
            error = -1;
        else
            *nlat = 2;
    }
    *new_cells = (*nlon) * (*nlat);
    return error;
}

/* _tesseroid.pyx:152-176 */
static inline double scale_nodes(double w, double e, double s, double n, double top,
                                 double bottom, double *lonc, double *sinlatc, double *coslatc,
                                 double *rc)
{
    double dlon, dlat, dr, mlon, mlat, mr, latc, scale;
    int i;
    dlon = e - w;
    dlat = n - s;
    dr = top - bottom;
    mlon = 0.5 * (e + w);
    mlat = 0.5 * (n + s);
    mr = 0.5 * (top + bottom + 2. * MEAN_EARTH_RADIUS);
    for (i = 0; i < 2; i++) {
        lonc[i] = d2r * (0.5 * dlon * nodes[i] + mlon);
        latc = d2r * (0.5 * dlat * nodes[i] + mlat);
        sinlatc[i] = sin(latc);
        coslatc[i] = cos(latc);
        rc[i] = (0.5 * dr * nodes[i] + mr);
    }
    scale = d2r * dlon * d2r * dlat * dr * 0.125;
    return scale;
}

/* _tesseroid.pyx:182-198 */
static inline int split(double w, double e, double s, double n, double top, double bottom,
                        int nlon, int nlat, double (*stack)[6], int stktop)
{
    int i, j;
    double dlon = (e - w) / nlon;
    double dlat = (n - s) / nlat;
    for (i = 0; i < nlon; i++) {
        for (j = 0; j < nlat; j++) {
            stktop += 1;
            stack[stktop][0] = w + i * dlon;
            stack[stktop][1] = w + (i + 1) * dlon;
            stack[stktop][2] = s + j * dlat;
            stack[stktop][3] = s + (j + 1) * dlat;
            stack[stktop][4] = top;
            stack[stktop][5] = bottom;
        }
    }
    return stktop;
}

/*
 * The GLQ kernels.  `variable` selects the *_variable flavour (density evaluated inside the
 * sum at rc[k] - R, _tesseroid.pyx:274,354,434,515) or the constant flavour (multiplied at
 * the end, :250,329,409,491).
 */
static inline double glq_kernel(int field, int variable, int kind, const double *p,
                                double lon, double sinlat, double coslat, double radius,
                                double scale, const double *lonc, const double *sinlatc,
                                const double *coslatc, const double *rc)
{
    int i, j, k;
    double kappa, r_sqr, coslon, sinlon = 0, kphi, l_sqr, cospsi, rc_sqr, dens = 1.0;
    double deltax, deltay, deltaz, term, l_5;
    double result = 0;
    r_sqr = radius * radius;
    for (i = 0; i < 2; i++) {
        coslon = cos(lon - lonc[i]);
        if (field == F_GY || field == F_GXY || field == F_GYY || field == F_GYZ)
            sinlon = sin(lonc[i] - lon);
        for (j = 0; j < 2; j++) {
            kphi = coslat * sinlatc[j] - sinlat * coslatc[j] * coslon;
            cospsi = sinlat * sinlatc[j] + coslat * coslatc[j] * coslon;
            for (k = 0; k < 2; k++) {
                rc_sqr = rc[k] * rc[k];
                l_sqr = r_sqr + rc_sqr - 2 * radius * rc[k] * cospsi;
                kappa = rc_sqr * coslatc[j];
                if (variable)
                    dens = orc_density_eval(kind, p, rc[k] - MEAN_EARTH_RADIUS);
                switch (field) {
                case F_POT:
                    term = variable ? dens * kappa / sqrt(l_sqr) : kappa / sqrt(l_sqr);
                    break;
                case F_GX:
                    term = variable ? dens * kappa * rc[k] * kphi / pow(l_sqr, 1.5)
                                    : kappa * rc[k] * kphi / pow(l_sqr, 1.5);
                    break;
                case F_GY:
                    term = variable
                               ? dens * kappa * (rc[k] * coslatc[j] * sinlon / pow(l_sqr, 1.5))
                               : kappa * (rc[k] * coslatc[j] * sinlon / pow(l_sqr, 1.5));
                    break;
                case F_GZ:
                    term = variable
                               ? dens * kappa * (rc[k] * cospsi - radius) / pow(l_sqr, 1.5)
                               : kappa * (rc[k] * cospsi - radius) / pow(l_sqr, 1.5);
                    break;
                default:
                    /* Fatiando 0.5 tensor integrands (z up); PARITY UNPINNED */
                    deltax = rc[k] * kphi;
                    deltay = rc[k] * coslatc[j] * sinlon;
                    deltaz = rc[k] * cospsi - radius;
                    l_5 = pow(l_sqr, 2.5);
                    switch (field) {
                    case F_GXX: term = kappa * (3 * (deltax * deltax) - l_sqr) / l_5; break;
                    case F_GXY: term = kappa * (3 * deltax * deltay) / l_5; break;
                    case F_GXZ: term = kappa * (3 * deltax * deltaz) / l_5; break;
                    case F_GYY: term = kappa * (3 * (deltay * deltay) - l_sqr) / l_5; break;
                    case F_GYZ: term = kappa * (3 * deltay * deltaz) / l_5; break;
                    default:    term = kappa * (3 * (deltaz * deltaz) - l_sqr) / l_5; break;
                    }
                    if (variable)
                        term = dens * term;
                    break;
                }
                result += term;
            }
        }
    }
    if (field == F_GZ)
        result *= -1;
    if (variable)
        return result * scale;
    return result * scale * p[0];
}

typedef struct {
    int64_t nodes, leaves, splits, too_small, nonroot_nodes, nonroot_leaves;
    int overflow;
} walk_stats;

/* _tesseroid.pyx:34-98 for one piece and the point range [p0, p1). */
static void rediscretizer(int field, int variable, int kind, const double *dp,
                          const double *bounds, double ratio, int stack_size, int64_t p0,
                          int64_t p1, const double *lons, const double *sinlats,
                          const double *coslats, const double *radii, double *result,
                          int32_t *leaf_counts, double (*stack)[6], walk_stats *st)
{
    int64_t l;
    int i, nlon, nlat, new_cells, stktop, error, root;
    double lonc[2], sinlatc[2], coslatc[2], rc[2];
    double w, e, s, n, top, bottom, Llon, Llat, Lr, distance, scale, res;
    for (l = p0; l < p1; l++) {
        double lon = lons[l], sinlat = sinlats[l], coslat = coslats[l], radius = radii[l];
        int32_t nleaf = 0;
        res = 0;
        for (i = 0; i < 6; i++)
            stack[0][i] = bounds[i];
        stktop = 0;
        root = 1;
        while (stktop >= 0) {
            w = stack[stktop][0];
            e = stack[stktop][1];
            s = stack[stktop][2];
            n = stack[stktop][3];
            top = stack[stktop][4];
            bottom = stack[stktop][5];
            stktop -= 1;
            st->nodes++;
            if (!root)
                st->nonroot_nodes++;
            distance = distance_n_size(w, e, s, n, top, bottom, lon, sinlat, coslat, radius,
                                       &Llon, &Llat, &Lr);
            error = divisions(distance, Llon, Llat, Lr, ratio, &nlon, &nlat, &new_cells);
            st->too_small -= error;
            if (new_cells > 1) {
                if (stktop + nlon * nlat > stack_size) {
                    st->overflow = 1;   /* ValueError('Tesseroid stack overflow'), :88-89 */
                    return;
                }
                st->splits++;
                stktop = split(w, e, s, n, top, bottom, nlon, nlat, stack, stktop);
            } else {
                scale = scale_nodes(w, e, s, n, top, bottom, lonc, sinlatc, coslatc, rc);
                res += glq_kernel(field, variable, kind, dp, lon, sinlat, coslat, radius, scale,
                                  lonc, sinlatc, coslatc, rc);
                nleaf++;
                st->leaves++;
                if (!root)
                    st->nonroot_leaves++;
            }
            root = 0;
        }
        result[l] += res;
        if (leaf_counts)
            leaf_counts[l] += nleaf;
    }
}

/*
 * tesseroid.py:198-234 (_forward_model) + :179-195 (the njobs point chunking).
 *
 * bounds      [n_tess][6]  w, e, s, n, top, bottom (degrees, metres)
 * kind        [n_tess]     density kind; kind < 0 -> tesseroid skipped (None / no density)
 * params      [n_tess][ORC_NPAR]
 * delta       NaN -> delta=None (no density-based discretisation)
 * result      [n_points]   accumulated in place, kernel units (field / G)
 * leaf_counts NULL or [n_tess][n_points] int32, leaves per (tesseroid, point), summed over
 *             the radial pieces; must be zero-filled by the caller
 * stats       NULL or int64[8], see ST_*
 * n_threads   point chunks evaluated concurrently (the reference's njobs, tesseroid.py:318-323)
 *
 * returns 0, or -1 on stack overflow (the reference raises ValueError), or -2 if a tesseroid
 * needs more than 4096 radial pieces.
 */
typedef struct {
    int field, stack_size, t, nparts, status;
    int64_t n_tess, n_points;
    const double *bounds, *params, *lon, *sinlat, *coslat, *radius;
    const int32_t *kind;
    double delta, ratio;
    double *result;
    int32_t *leaf_counts;
    walk_stats st;
    int64_t npieces;
} chunk_job;

static void *run_chunk(void *arg)
{
    chunk_job *j = (chunk_job *)arg;
    /* _split_arrays: n = size//nparts, the last chunk takes the remainder */
    int64_t chunk = j->n_points / j->nparts;
    int64_t p0 = (int64_t)j->t * chunk;
    int64_t p1 = (j->t == j->nparts - 1) ? j->n_points : p0 + chunk;
    double (*stack)[6] = malloc(sizeof(double[6]) * (size_t)(j->stack_size + 8));
    double *ptop = malloc(sizeof(double) * 4096), *pbot = malloc(sizeof(double) * 4096);
    int64_t it;
    memset(&j->st, 0, sizeof(j->st));
    j->npieces = 0;
    j->status = 0;
    for (it = 0; it < j->n_tess && !j->status; it++) {
        const double *b = j->bounds + 6 * it;
        const double *dp = j->params + ORC_NPAR * it;
        int kd = j->kind[it], np = 1, ip;
        int variable = kd != D_CONST;
        double pb[6];
        if (kd < 0)
            continue;
        ptop[0] = b[4];
        pbot[0] = b[5];
        if (variable && !isnan(j->delta)) {
            np = orc_density_discretize(b[4], b[5], kd, dp, j->delta, 4096, ptop, pbot);
            if (np < 0) {
                j->status = -2;
                break;
            }
        }
        memcpy(pb, b, sizeof(pb));
        for (ip = 0; ip < np; ip++) {
            pb[4] = ptop[ip];
            pb[5] = pbot[ip];
            j->npieces++;
            rediscretizer(j->field, variable, kd, dp, pb, j->ratio, j->stack_size, p0, p1,
                          j->lon, j->sinlat, j->coslat, j->radius, j->result,
                          j->leaf_counts ? j->leaf_counts + it * j->n_points : NULL, stack,
                          &j->st);
            if (j->st.overflow) {
                j->status = -1;
                break;
            }
        }
    }
    free(stack);
    free(ptop);
    free(pbot);
    return NULL;
}

int orc_forward(int field, int64_t n_tess, const double *bounds, const int32_t *kind,
                const double *params, double delta, double ratio, int stack_size,
                int64_t n_points, const double *lon, const double *sinlat,
                const double *coslat, const double *radius, double *result,
                int32_t *leaf_counts, int64_t *stats, int n_threads)
{
    int status = 0, t, nparts;
    int64_t tot[ST_COUNT];
    chunk_job *jobs;
    pthread_t *th;
    memset(tot, 0, sizeof(tot));
    if (field < 0 || field >= F_COUNT)
        return -3;
    nparts = n_threads < 1 ? 1 : n_threads;
    if ((int64_t)nparts > n_points)
        nparts = n_points > 0 ? (int)n_points : 1;
    jobs = (chunk_job *)calloc((size_t)nparts, sizeof(chunk_job));
    th = (pthread_t *)calloc((size_t)nparts, sizeof(pthread_t));
    for (t = 0; t < nparts; t++) {
        chunk_job *j = &jobs[t];
        j->field = field; j->stack_size = stack_size; j->t = t; j->nparts = nparts;
        j->n_tess = n_tess; j->n_points = n_points; j->bounds = bounds; j->params = params;
        j->lon = lon; j->sinlat = sinlat; j->coslat = coslat; j->radius = radius;
        j->kind = kind; j->delta = delta; j->ratio = ratio; j->result = result;
        j->leaf_counts = leaf_counts;
        if (t > 0)
            pthread_create(&th[t], NULL, run_chunk, j);
    }
    run_chunk(&jobs[0]);
    for (t = 1; t < nparts; t++)
        pthread_join(th[t], NULL);
    for (t = 0; t < nparts; t++) {
        walk_stats *st = &jobs[t].st;
        tot[ST_NODES] += st->nodes;
        tot[ST_LEAVES] += st->leaves;
        tot[ST_SPLITS] += st->splits;
        tot[ST_TOO_SMALL] += st->too_small;
        tot[ST_NONROOT_NODES] += st->nonroot_nodes;
        tot[ST_NONROOT_LEAVES] += st->nonroot_leaves;
        if (jobs[t].status && !status)
            status = jobs[t].status;
    }
    tot[ST_PIECES] = jobs[0].npieces;
    free(jobs);
    free(th);
    if (stats)
        memcpy(stats, tot, sizeof(tot));
    return status;
}
