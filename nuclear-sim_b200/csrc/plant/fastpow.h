// x ** y for positive finite x on the device.
//
// Python's float ** float is libm pow.  libdevice's pow (<= 2 ulp) is 25 % of the step kernel's executed instructions:
// ~55 instructions of argument screening (signs, integers, infinities, NaNs, zeros) in front of a ~170-instruction
// double-double log / exp core, 78 calls per plant-step, every one with a positive finite base.  nps_pow_pos is that
// core alone, written for this operand class:
//     x = 2^e * m,  m in [sqrt(1/2), sqrt(2));   u = (m - 1)/(m + 1) as a double-double (one division, the residual by
//     FMA);  log m = 2 atanh u = 2u + u^3 * P(u^2)  (12 series terms, truncation < 2^-65);  log x = e ln2 + log m
//     accumulated as a double-double;  (Ph, Pl) = y * log x with the product's rounding error recovered by FMA;
//     exp: k = round(Ph / ln2), r = Ph - k ln2 (two FMAs) + Pl, degree-14 Taylor polynomial, 2^k by exponent arithmetic.
// Measured against exact arithmetic (mpmath, 200 bits) on the operand ranges of this model: max error 1.06 ulp; never
// more than 1 ulp away from glibc's pow on 2e7 random (x in 1e-6..1e6, y in -3..4) and near-1 operands - i.e. inside
// libdevice pow's own error bound, which is what the parity tolerances already absorb.  Outside the guarded range
// (x not a positive normal number, |y| >= 1024, |y log x| >= 64) the caller falls back to libdevice pow.
// Coefficients sit in constant memory so the FMA chain reads them as operands instead of materialising each 64-bit
// literal with two moves.  The step uses it on the device only: the host build (test infrastructure) keeps libm pow, the
// reference's own function, and compiles this header just to test the algorithm without a GPU (nps_oracle_fastpow).
#pragma once
#include <string.h>
#include "hd.h"

namespace nps {

// 2/(2k+1), k = 1..12 (atanh series)  then  1/k!, k = 2..14 (exp series)
#define NPS_POW_TAB { \
    2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13, 2.0 / 15, 2.0 / 17, 2.0 / 19, 2.0 / 21, 2.0 / 23, 2.0 / 25, \
    1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, \
    1.0 / 479001600, 1.0 / 6227020800.0, 1.0 / 87178291200.0}
#if defined(__CUDACC__)
__constant__ double nps_pow_tab_dev[25] = NPS_POW_TAB;
#endif
static const double nps_pow_tab_host[25] = NPS_POW_TAB;
#if defined(__CUDA_ARCH__)
#define nps_pow_tab nps_pow_tab_dev
#else
#define nps_pow_tab nps_pow_tab_host
#endif

NPS_HD long long nps_bits(double v) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    long long b; memcpy(&b, &v, sizeof(b)); return b;
#endif
}
NPS_HD double nps_from_bits(long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double v; memcpy(&v, &b, sizeof(v)); return v;
#endif
}
NPS_HD double nps_rcp_f32(double g) {     // 1/g to single precision, correctly rounded on both sides
#if defined(__CUDA_ARCH__)
    return (double)__frcp_rn((float)g);
#else
    return (double)(1.0f / (float)g);
#endif
}

// log x as a double-double (Lh, Ll) for a positive normal x; false outside the guarded range.
NPS_HD bool nps_pow_log(double x, double& Lh_out, double& Ll_out) {
    const long long ix = nps_bits(x);
    const int be = (int)((ix >> 52) & 0x7ff);
    if (!(ix > 0) || (unsigned)(be - 23) >= 2000u) return false;
    int e = be - 1023;
    double m = nps_from_bits((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double f = m - 1.0;                       // exact
    const double g = m + 1.0;
    const double g_lo = m - (g - 1.0);              // m + 1 == g + g_lo exactly
    const double u = f / g;
    double r = fma(-u, g, f);
    r = fma(-u, g_lo, r);                           // f - u * (g + g_lo)
    const double u_lo = r * nps_rcp_f32(g);
    const double u2 = u * u;
    const double u4 = u2 * u2;
    double pe = nps_pow_tab[10], po = nps_pow_tab[11];     // even / odd halves of the series: two independent FMA chains
#pragma unroll
    for (int k = 8; k >= 0; k -= 2) { pe = fma(pe, u4, nps_pow_tab[k]); po = fma(po, u4, nps_pow_tab[k + 1]); }
    const double p = fma(po, u2, pe);
    const double lh = 2.0 * u;
    const double ll = fma(u2 * u, p, 2.0 * u_lo);
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double ed = (double)e;
    const double t_hi = ed * LN2_HI;                // exact: LN2_HI has 21 trailing zero bits
    const double Lh = t_hi + lh;
    double Ll = (t_hi - Lh) + lh;                   // fast two-sum (|t_hi| >= |lh| or t_hi == 0)
    Ll += fma(ed, LN2_LO, ll);
    Lh_out = Lh; Ll_out = Ll;
    return true;
}

// exp(y * (Lh + Ll)); false outside the guarded range (|y| >= 1024 or |y log x| >= 64).
NPS_HD bool nps_pow_exp(double y, double Lh, double Ll, double& out) {
    if (!(fabs(y) < 1024.0)) return false;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10, LOG2E = 1.44269504088896338700e+00;
    const double Ph = y * Lh;
    const double Pl = fma(y, Lh, -Ph) + y * Ll;
    if (!(fabs(Ph) < 64.0)) return false;           // results within 1e-28 .. 1e28: the error budget was verified there
    const double magic = 6755399441055744.0;        // 1.5 * 2^52: adding it rounds to the nearest integer
    const double kd = (Ph * LOG2E + magic) - magic;
    const double r1 = fma(-kd, LN2_HI, Ph);
    const double rr = fma(-kd, LN2_LO, r1) + Pl;
    const double r2 = rr * rr;
    double qe = nps_pow_tab[24], qo = nps_pow_tab[23];     // q(r) = sum_j tab[12 + j] r^j, split by parity of j
#pragma unroll
    for (int k = 22; k >= 12; k -= 2) { qe = fma(qe, r2, nps_pow_tab[k]); if (k - 1 >= 13) qo = fma(qo, r2, nps_pow_tab[k - 1]); }
    const double q = fma(qo, rr, qe);
    const double ex = 1.0 + fma(r2, q, rr);
    out = nps_from_bits(nps_bits(ex) + ((long long)kd * (1LL << 52)));
    return true;
}

NPS_HD bool nps_pow_pos(double x, double y, double& out) {
    double Lh, Ll;
    if (!(fabs(y) < 1024.0) || !nps_pow_log(x, Lh, Ll)) return false;
    return nps_pow_exp(y, Lh, Ll, out);
}


// ------------------------------------------------------------------------------------------------
// Table-driven variant (tuning builds with -DNPS_POW_TABLES; NOT the shipped path: it measured 2 % slower on B200 because
// its table loads miss L1 under the frame's streaming traffic, profiles/r02_ab_pow.txt): the same double-double log / exp structure with two 128-entry tables
// (csrc/tools/make_pow_tables.py -> fastpow_tables.inc) in place of the division + 12-term atanh series and the
// degree-14 exponential series:
//   x = 2^e z, z in [1, 2);  i = 7 leading mantissa bits;  r = z * invc_i - 1 (one FMA, |r| < 2^-8);
//   log x = e ln2 + logc_i + log1p(r), log1p by its series to r^7, accumulated as hi + lo (absolute error ~2^-68);
//   (Ph, Pl) = y * log x;  k = round(Ph * 128 / ln2), r' = Ph - k ln2/128 + Pl (|r'| < 0.0028);
//   x^y = 2^(k >> 7) * t_(k & 127) * (1 + r' + ... + r'^5 / 120).
// ------------------------------------------------------------------------------------------------
#include "fastpow_tables.inc"
#if defined(__CUDACC__)
__device__ const double nps_pow_log_table_dev[128 * 3] = NPS_POW_LOG_TABLE;
__device__ const double nps_pow_exp_table_dev[128 * 2] = NPS_POW_EXP_TABLE;
#endif
static const double nps_pow_log_table_host[128 * 3] = NPS_POW_LOG_TABLE;
static const double nps_pow_exp_table_host[128 * 2] = NPS_POW_EXP_TABLE;
#if defined(__CUDA_ARCH__)
#define NPS_LOGT(i) __ldg(nps_pow_log_table_dev + (i))
#define NPS_EXPT(i) __ldg(nps_pow_exp_table_dev + (i))
#else
#define NPS_LOGT(i) nps_pow_log_table_host[i]
#define NPS_EXPT(i) nps_pow_exp_table_host[i]
#endif

NPS_HD bool nps_pow_log_tab(double x, double& Lh_out, double& Ll_out) {
    const long long ix = nps_bits(x);
    const int be = (int)((ix >> 52) & 0x7ff);
    if (!(ix > 0) || (unsigned)(be - 23) >= 2000u) return false;
    const int e = be - 1023;
    const int i = (int)((ix >> 45) & 127);
    const double z = nps_from_bits((ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);     // [1, 2)
    const double invc = NPS_LOGT(3 * i), logc_hi = NPS_LOGT(3 * i + 1), logc_lo = NPS_LOGT(3 * i + 2);
    const double r = fma(z, invc, -1.0);
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double ed = (double)e;
    const double a = ed * LN2_HI;                   // exact: LN2_HI has 21 trailing zero bits
    const double s1 = a + logc_hi;
    const double e1 = (a - s1) + logc_hi;           // fast two-sum: a == 0 or |a| >= ln2 > logc_hi
    const double s2 = s1 + r;
    const double bb = s2 - s1;
    const double e2 = (s1 - (s2 - bb)) + (r - bb);  // two-sum (r may exceed s1 when e == 0 and i is small)
    const double r2 = r * r;
    // log1p(r) - r = -r^2/2 + r^3/3 - r^4/4 + r^5/5 - r^6/6 + r^7/7
    const double q = fma(r, fma(r, fma(r, fma(r, 1.0 / 7, -1.0 / 6), 1.0 / 5), -1.0 / 4), 1.0 / 3);
    const double tail = fma(r2 * r, q, -0.5 * r2);
    const double lo = ((e1 + e2) + fma(ed, LN2_LO, logc_lo)) + tail;
    const double Lh = s2 + lo;
    Lh_out = Lh;
    Ll_out = (s2 - Lh) + lo;
    return true;
}

NPS_HD bool nps_pow_exp_tab(double y, double Lh, double Ll, double& out) {
    if (!(fabs(y) < 1024.0)) return false;
    const double Ph = y * Lh;
    const double Pl = fma(y, Lh, -Ph) + y * Ll;
    if (!(fabs(Ph) < 64.0)) return false;
    const double magic = 6755399441055744.0;        // 1.5 * 2^52
    const double kd = (Ph * NPS_POW_128_LN2 + magic) - magic;
    const long long k = (long long)kd;
    const double r1 = fma(-kd, NPS_POW_LN2_128_HI, Ph);            // exact: |kd| < 2^14, 24 trailing zero bits in the constant
    const double rr = fma(-kd, NPS_POW_LN2_128_LO, r1) + Pl;
    const int j = (int)(k & 127);
    const double t_hi = NPS_EXPT(2 * j), t_lo = NPS_EXPT(2 * j + 1);
    const double r2 = rr * rr;
    // exp(rr) - 1 = rr + rr^2 (1/2 + rr/6 + rr^2 (1/24 + rr/120))
    const double p = fma(r2, fma(r2, fma(rr, 1.0 / 120, 1.0 / 24), fma(rr, 1.0 / 6, 0.5)), rr);
    const double res = t_hi + fma(t_hi, p, t_lo);
    out = nps_from_bits(nps_bits(res) + ((k >> 7) << 52));
    return true;
}

NPS_HD bool nps_pow_pos_tab(double x, double y, double& out) {
    double Lh, Ll;
    if (!(fabs(y) < 1024.0) || !nps_pow_log_tab(x, Lh, Ll)) return false;
    return nps_pow_exp_tab(y, Lh, Ll, out);
}

#undef NPS_LOGT
#undef NPS_EXPT
#undef nps_pow_tab
#undef NPS_POW_TAB
}  // namespace nps
