// tvd_ddmath.cuh -- sin, cos and exp evaluated in double-double arithmetic and rounded once.
//
// Why: the density-based discretisation (K1) takes an argmax over 101 samples of the normalised
// density (tesseroid.py:297-299).  For a sinusoidal density with a whole number of periods over
// the tesseroid the candidate maxima are mathematically equal, so the winner -- and with it the
// whole radial split list -- is decided by the last bit of sin().  The reference evaluates the
// density with NumPy/glibc, whose sin/cos/exp are correctly rounded in all but a tiny fraction
// of cases (the glibc of the paper's time was correctly rounded outright).  Evaluating in
// double-double (~2^-100) and rounding once reproduces the correctly rounded value, so K1 agrees
// with the reference even in those tie situations.  This is O(tesseroids), not O(pairs), so the
// cost (a few hundred flops per call) is irrelevant.
#pragma once

struct dd_t {
    double hi, lo;
};

__device__ __forceinline__ dd_t dd_two_sum(double a, double b)
{
    const double s = a + b;
    const double bb = s - a;
    return {s, (a - (s - bb)) + (b - bb)};
}
__device__ __forceinline__ dd_t dd_quick_two_sum(double a, double b)
{
    const double s = a + b;
    return {s, b - (s - a)};
}
__device__ __forceinline__ dd_t dd_two_prod(double a, double b)
{
    const double p = a * b;
    return {p, fma(a, b, -p)};
}
__device__ __forceinline__ dd_t dd_add(dd_t a, dd_t b)
{
    dd_t s = dd_two_sum(a.hi, b.hi);
    const dd_t t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return dd_quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd_t dd_mul(dd_t a, dd_t b)
{
    dd_t p = dd_two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return dd_quick_two_sum(p.hi, p.lo);
}

// 1/n! for n = 2..29 as double-double (exact rational arithmetic)
__device__ const double DD_INV_FACT[28][2] = {
    {0.5, 0.0},
    {0.16666666666666666, 9.25185853854297e-18},
    {0.041666666666666664, 2.3129646346357427e-18},
    {0.008333333333333333, 1.1564823173178714e-19},
    {0.001388888888888889, -5.300543954373577e-20},
    {0.0001984126984126984, 1.7209558293420705e-22},
    {2.48015873015873e-05, 2.1511947866775882e-23},
    {2.7557319223985893e-06, -1.858393274046472e-22},
    {2.755731922398589e-07, 2.3767714622250297e-23},
    {2.505210838544172e-08, -1.448814070935912e-24},
    {2.08767569878681e-09, -1.20734505911326e-25},
    {1.6059043836821613e-10, 1.2585294588752098e-26},
    {1.1470745597729725e-11, 2.0655512752830745e-28},
    {7.647163731819816e-13, 7.03872877733453e-30},
    {4.779477332387385e-14, 4.399205485834081e-31},
    {2.8114572543455206e-15, 1.6508842730861433e-31},
    {1.5619206968586225e-16, 1.1910679660273754e-32},
    {8.22063524662433e-18, 2.2141894119604265e-34},
    {4.110317623312165e-19, 1.4412973378659527e-36},
    {1.9572941063391263e-20, -1.3643503830087908e-36},
    {8.896791392450574e-22, -7.911402614872376e-38},
    {3.868170170630684e-23, -8.843177655482344e-40},
    {1.6117375710961184e-24, -3.6846573564509766e-41},
    {6.446950284384474e-26, -1.9330404233703465e-42},
    {2.4795962632247976e-27, -1.2953730964765229e-43},
    {9.183689863795546e-29, 1.4303150396787322e-45},
    {3.279889237069838e-30, 1.5117542744029879e-46},
    {1.1309962886447716e-31, 1.0498015412959506e-47},
};
__device__ __forceinline__ dd_t dd_inv_fact(int n)      // 2 <= n <= 29
{
    return {DD_INV_FACT[n - 2][0], DD_INV_FACT[n - 2][1]};
}

// sin(r), cos(r) for a double-double |r| <= ~0.8 by their Taylor series in double-double
__device__ inline void dd_sincos_kernel(dd_t r, dd_t &s, dd_t &c)
{
    const dd_t z = dd_mul(r, r);
    // sin r = r (1 - z/3! + z^2/5! - ... ), terms up to r^27
    dd_t ps = dd_inv_fact(27);
    for (int n = 25; n >= 3; n -= 2) {
        ps = dd_mul(ps, z);
        dd_t cn = dd_inv_fact(n);
        ps = dd_add({cn.hi, cn.lo}, {-ps.hi, -ps.lo});
    }
    // ps = 1/3! - z/5! + ...   ->  sin = r - r z ps
    dd_t t = dd_mul(dd_mul(r, z), ps);
    s = dd_add(r, {-t.hi, -t.lo});
    // cos r = 1 - z/2! + z^2/4! - ..., terms up to r^28
    dd_t pc = dd_inv_fact(28);
    for (int n = 26; n >= 2; n -= 2) {
        pc = dd_mul(pc, z);
        dd_t cn = dd_inv_fact(n);
        pc = dd_add({cn.hi, cn.lo}, {-pc.hi, -pc.lo});
    }
    // pc = 1/2! - z/4! + ...   ->  cos = 1 - z pc
    t = dd_mul(z, pc);
    c = dd_add({1.0, 0.0}, {-t.hi, -t.lo});
}

// correctly rounded (with overwhelming probability) sin and cos of a double
__device__ inline void dd_sincos(double x, double &sn, double &cs)
{
    if (!(fabs(x) < 1.0e6)) {       // huge, inf, NaN: libdevice
        sincos(x, &sn, &cs);
        return;
    }
    // x = k*pi/2 + r;  pi/2 = P1 + P2 + P3 + P4, P1 and P2 with 33 significant bits so that
    // k*P1 and k*P2 are exact for |k| < 2^20
    const double k = rint(x * 0.6366197723675814);
    const double a = fma(-k, 1.5707963267341256, x);                    // exact
    const dd_t s1 = dd_two_sum(a, -(k * 6.077100506303966e-11));        // exact product
    const dd_t p3 = dd_two_prod(k, 2.0222662487959506e-21);
    const dd_t s2 = dd_two_sum(s1.hi, -p3.hi);
    const double lo = ((s1.lo + s2.lo) - p3.lo) - k * 1.0085854035872483e-37;
    const dd_t r = dd_quick_two_sum(s2.hi, lo);
    dd_t s, c;
    dd_sincos_kernel(r, s, c);
    const long long q = (long long)k;
    const double sv = s.hi, cv = c.hi;      // hi of a normalised double-double = RN(hi + lo)
    switch (q & 3) {
    case 0: sn = sv; cs = cv; break;
    case 1: sn = cv; cs = -sv; break;
    case 2: sn = -sv; cs = -cv; break;
    default: sn = -cv; cs = sv; break;
    }
}
__device__ inline double dd_sin(double x)
{
    double s, c;
    dd_sincos(x, s, c);
    return s;
}
__device__ inline double dd_cos(double x)
{
    double s, c;
    dd_sincos(x, s, c);
    return c;
}

// correctly rounded (with overwhelming probability) exp of a double
__device__ inline double dd_exp(double x)
{
    if (!(fabs(x) < 700.0))         // overflow/underflow range, inf, NaN: libdevice
        return exp(x);
    // x = k ln2 + r;  ln2 = L1 (32 bits) + L2 + L3
    const double k = rint(x * 1.4426950408889634);
    const double a = fma(-k, 0.6931471803691238, x);                    // exact
    const dd_t p2 = dd_two_prod(k, 1.9082149292705877e-10);
    const dd_t s1 = dd_two_sum(a, -p2.hi);
    const double lo = (s1.lo - p2.lo) - k * 1.1612227229362532e-26;
    const dd_t r = dd_quick_two_sum(s1.hi, lo);
    // exp r = 1 + r (1 + r/2! + r^2/3! + ...), |r| <= 0.347, terms up to r^26
    dd_t p = dd_inv_fact(26);
    for (int n = 25; n >= 2; n--) {
        p = dd_mul(p, r);
        p = dd_add(dd_inv_fact(n), p);
    }
    // p = 1/2! + r/3! + ...  ->  exp = 1 + r + r^2 p
    dd_t e = dd_mul(dd_mul(r, r), p);
    e = dd_add(e, r);
    e = dd_add({1.0, 0.0}, e);
    return scalbn(e.hi, (int)k);
}
