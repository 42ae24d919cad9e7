// host check of the table-based correctly rounded sin/cos core against quad and long double
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <quadmath.h>
#include "../../tesseroid_variable_density_b200/csrc/tvd_crtrig.cuh"

static uint64_t st = 88172645463325252ull;
static double rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (st >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 10000000;
    double ranges[][2] = {{-0.7853, 0.7853}, {-1.5708, 1.5708}, {-7, 7}, {-1e-3, 1e-3}, {-400, 400}, {-1e-8,1e-8}};
    for (auto &rg : ranges) {
        long bad_s = 0, bad_c = 0, notok = 0, bad_ok_s = 0, bad_ok_c = 0;
        for (long i = 0; i < n; i++) {
            double x = rg[0] + (rg[1] - rg[0]) * rnd();
            double rh = x, rl = 0; int q = 0;
            if (fabs(x) > 0.78539816339744828) cr_reduce_pio2(x, rh, rl, q);
            double sv, cv; bool ok;
            cr_sincos_core<true>(rh, rl, sv, cv, ok);
            double s, c;
            switch (q & 3) { case 0: s = sv; c = cv; break; case 1: s = cv; c = -sv; break;
                             case 2: s = -sv; c = -cv; break; default: s = -cv; c = sv; }
            double gs = (double)sinq((__float128)x), gc = (double)cosq((__float128)x);
            if (!ok) notok++;
            if (s != gs) { bad_s++; if (ok) { bad_ok_s++; if (bad_ok_s < 4) printf("  sin x=%a got %a quad %a sinl %La\n", x, s, gs, sinl((long double)x)); } }
            if (c != gc) { bad_c++; if (ok) { bad_ok_c++; if (bad_ok_c < 4) printf("  cos x=%a got %a quad %a cosl %La\n", x, c, gc, cosl((long double)x)); } }
        }
        printf("range [%g, %g]: n=%ld  not-ok %ld (%.2e)  sin != quad %ld (with ok: %ld)  cos != quad %ld (with ok: %ld)\n",
               rg[0], rg[1], n, notok, (double)notok / n, bad_s, bad_ok_s, bad_c, bad_ok_c);
    }
    double eranges[][2] = {{-1.0, 1.0}, {-40.0, 5.0}, {-690.0, 690.0}, {-1e-6, 1e-6}};
    for (auto &rg : eranges) {
        long bad = 0, notok = 0, bad_ok = 0;
        for (long i = 0; i < n; i++) {
            double x = rg[0] + (rg[1] - rg[0]) * rnd();
            int k; bool ok;
            double r = cr_exp_core(x, k, ok);
            double got = scalbn(r, k);
            double want = (double)expq((__float128)x);
            if (!ok) notok++;
            if (got != want) { bad++; if (ok) { bad_ok++; if (bad_ok < 4) printf("  exp x=%a got %a quad %a\n", x, got, want); } }
        }
        printf("exp range [%g, %g]: n=%ld  not-ok %ld (%.2e)  exp != quad %ld (with ok: %ld)\n",
               rg[0], rg[1], n, notok, (double)notok / n, bad, bad_ok);
    }
    return 0;
}
// build: g++ -O2 -ffp-contract=off -mfma crtrig_host_test.cpp -lquadmath -lm
