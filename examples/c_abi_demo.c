/*
 * Plain-C use of libtvd.so (include/tvd.h): a 3 x 3 mesh of 1-degree tesseroids with an
 * exponential density, gz and potential on a 5 x 5 grid at 10 km.
 *
 *   gcc -O2 -I include examples/c_abi_demo.c -o c_abi_demo \
 *       -L tesseroid_variable_density_b200/csrc -ltvd -lm \
 *       -Wl,-rpath,$PWD/tesseroid_variable_density_b200/csrc
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "tvd.h"

#define R 6378137.0
#define G 0.00000000006673

int main(void)
{
    enum { NT = 9, NP = 25 };
    double bounds[NT][6], params[NT][TVD_NPAR] = {{0}};
    int32_t kind[NT];
    const double top = 0.0, bottom = -35000.0, b = 3.0;
    const double A = (3300.0 - 2670.0) / (1.0 - exp(-b)), C = 3300.0 - A;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double *t = bounds[3 * i + j];
            t[0] = -1.5 + j; t[1] = t[0] + 1.0;       /* w, e (degrees) */
            t[2] = -1.5 + i; t[3] = t[2] + 1.0;       /* s, n           */
            t[4] = top; t[5] = bottom;                /* metres         */
            kind[3 * i + j] = TVD_DENS_EXP;           /* A*exp(-b*(h - bottom)/T) + C */
            double *p = params[3 * i + j];
            p[0] = A; p[1] = -b; p[2] = bottom; p[3] = top - bottom; p[4] = C;
        }
    /* points as the reference's _convert_coords produces them */
    double lon[NP], sinlat[NP], coslat[NP], radius[NP], gz[NP], pot[NP];
    const double d2r = 3.14159265358979323846 / 180.0;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) {
            const double la = (-2.0 + i) * d2r, lo = (-2.0 + j) * d2r;
            lon[5 * i + j] = lo;
            sinlat[5 * i + j] = sin(la);
            coslat[5 * i + j] = cos(la);
            radius[5 * i + j] = R + 10e3;
        }
    if (tvd_device_count() < 1) {
        fprintf(stderr, "no CUDA device: libtvd has no CPU fallback\n");
        return 2;
    }
    tvd_model *model = NULL;
    int rc = tvd_model_create(NT, &bounds[0][0], kind, &params[0][0], 0.1, 0, &model);
    if (rc) {
        fprintf(stderr, "tvd_model_create: %d %s\n", rc, tvd_last_error());
        return 1;
    }
    int64_t stats[TVD_NSTATS];
    rc = tvd_forward(TVD_GZ, model, NP, lon, sinlat, coslat, radius, 2.5, 200, 1, NULL, gz, stats, NULL);
    if (!rc)
        rc = tvd_forward(TVD_POTENTIAL, model, NP, lon, sinlat, coslat, radius, 1.0, 200, 1, NULL,
                         pot, NULL, NULL);
    if (rc) {
        fprintf(stderr, "tvd_forward: %d %s\n", rc, tvd_last_error());
        return 1;
    }
    printf("pieces %lld, gz leaves %lld, near pairs %lld\n", (long long)tvd_model_num_pieces(model),
           (long long)stats[TVD_ST_LEAVES], (long long)stats[TVD_ST_NEAR_PAIRS]);
    printf("gz at the centre point   %.6f mGal\n", gz[12] * G * 1e5);
    printf("potential at the centre  %.6f m2/s2\n", pot[12] * G);
    tvd_model_free(model);
    return 0;
}
