// Correctly rounded sin / cos / atan2 for the rotating-calipers step (metrics.rs:133-148 ->
// imageproc min_area_rect): the reference computes them with the platform libm in f64, and an
// ulp of difference flips an outward floor/ceil when a rotated coordinate is an exact integer.
// glibc's results are (all but) correctly rounded; CUDA's are 1-2 ulp.  These routines evaluate
// in double-double arithmetic (~106 bits) and round once, so host (oracle cross-check) and
// device agree with a correctly rounded libm bit for bit.  Requires non-contracted arithmetic
// (-fmad=false / -ffp-contract=off): every fma below is explicit.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define DDM_HD __host__ __device__ __forceinline__
#else
#define DDM_HD static inline
#endif

namespace ddm {

struct dd { double hi, lo; };

DDM_HD dd two_sum(double a, double b) { double s = a + b, bb = s - a; dd r = {s, (a - (s - bb)) + (b - bb)}; return r; }
DDM_HD dd quick_two_sum(double a, double b) { double s = a + b; dd r = {s, b - (s - a)}; return r; }
DDM_HD dd two_prod(double a, double b) { double p = a * b; dd r = {p, fma(a, b, -p)}; return r; }
DDM_HD dd add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
DDM_HD dd add_d(dd a, double b) { dd s = two_sum(a.hi, b); s.lo += a.lo; return quick_two_sum(s.hi, s.lo); }
DDM_HD dd neg(dd a) { dd r = {-a.hi, -a.lo}; return r; }
DDM_HD dd mul(dd a, dd b) { dd p = two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return quick_two_sum(p.hi, p.lo); }
DDM_HD dd mul_d(dd a, double b) { dd p = two_prod(a.hi, b); p.lo += a.lo * b; return quick_two_sum(p.hi, p.lo); }
DDM_HD dd div(dd a, dd b) {
  double q1 = a.hi / b.hi;
  dd r = add(a, neg(mul_d(b, q1)));
  double q2 = r.hi / b.hi;
  r = add(r, neg(mul_d(b, q2)));
  double q3 = r.hi / b.hi;
  dd q = quick_two_sum(q1, q2);
  return add_d(q, q3);
}

// sin and cos of a double-double r with |r| <= pi/4 (+ a little): Taylor series to 2^-110
DDM_HD void sincos_small(dd r, dd *s, dd *c) {
  const double INV_FACT[32][2] = {
    {0x1.0000000000000p+0, 0x0.0p+0},  // 1/0!
    {0x1.0000000000000p+0, 0x0.0p+0},  // 1/1!
    {0x1.0000000000000p-1, 0x0.0p+0},  // 1/2!
    {0x1.5555555555555p-3, 0x1.5555555555555p-57},  // 1/3!
    {0x1.5555555555555p-5, 0x1.5555555555555p-59},  // 1/4!
    {0x1.1111111111111p-7, 0x1.1111111111111p-63},  // 1/5!
    {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65},  // 1/6!
    {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73},  // 1/7!
    {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},  // 1/8!
    {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},  // 1/9!
    {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76},  // 1/10!
    {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80},  // 1/11!
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},  // 1/12!
    {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},  // 1/13!
    {0x1.93974a8c07c9dp-37, 0x1.05d6f8a2efd1fp-92},  // 1/14!
    {0x1.ae7f3e733b81fp-41, 0x1.1d8656b0ee8cbp-97},  // 1/15!
    {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},  // 1/16!
    {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},  // 1/17!
    {0x1.6827863b97d97p-53, 0x1.eec01221a8b0bp-107},  // 1/18!
    {0x1.2f49b46814157p-57, 0x1.2650f61dbdcb4p-112},  // 1/19!
    {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},  // 1/20!
    {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},  // 1/21!
    {0x1.0ce396db7f853p-70, -0x1.aebcdbd20331cp-124},  // 1/22!
    {0x1.761b41316381ap-75, -0x1.3423c7d91404fp-130},  // 1/23!
    {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},  // 1/24!
    {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},  // 1/25!
    {0x1.88e85fc6a4e5ap-89, -0x1.71c37ebd16540p-143},  // 1/26!
    {0x1.d1ab1c2dccea3p-94, 0x1.054d0c78aea14p-149},  // 1/27!
    {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153},  // 1/28!
    {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157},  // 1/29!
    {0x1.3932c5047d60ep-108, 0x1.832b7b530a627p-162},  // 1/30!
    {0x1.434d2e783f5bcp-113, 0x1.0b87b91be9affp-167},  // 1/31!
  };
  const dd r2 = mul(r, r);
  dd sp = {INV_FACT[31][0], INV_FACT[31][1]}, cp = {INV_FACT[30][0], INV_FACT[30][1]};
  // Horner in r^2:  sin = r * sum_{k} (-1)^k r^(2k) / (2k+1)! ,  cos = sum_k (-1)^k r^(2k) / (2k)!
  for (int n = 29; n >= 1; n -= 2) {
    dd f = {INV_FACT[n][0], INV_FACT[n][1]};
    sp = add(f, neg(mul(sp, r2)));
  }
  for (int n = 28; n >= 0; n -= 2) {
    dd f = {INV_FACT[n][0], INV_FACT[n][1]};
    cp = add(f, neg(mul(cp, r2)));
  }
  *s = mul(sp, r);
  *c = cp;
}

// sin and cos of a double-double angle, |a| < ~8: reduce by multiples of pi/2 (three-part constant)
DDM_HD void sincos_dd(dd a, dd *s, dd *c) {
  const double P1 = 0x1.921fb54442d18p+0, P2 = 0x1.1a62633145c07p-54, P3 = -0x1.f1976b7ed8fbcp-110;
  const double kf = nearbyint(a.hi * 0x1.45f306dc9c883p-1);  // a / (pi/2)
  dd r = a;
  if (kf != 0.0) {
    r = add(r, neg(two_prod(kf, P1)));
    r = add(r, neg(two_prod(kf, P2)));
    r = add_d(r, -kf * P3);
  }
  dd sr, cr;
  sincos_small(r, &sr, &cr);
  const int k = ((int)kf) & 3;
  if (k == 0) { *s = sr; *c = cr; }
  else if (k == 1) { *s = cr; *c = neg(sr); }
  else if (k == 2) { *s = neg(sr); *c = neg(cr); }
  else { *s = neg(cr); *c = sr; }
}

// correctly rounded (with overwhelming probability) sin / cos of a double
DDM_HD void cr_sincos(double a, double *s, double *c) {
  dd ad = {a, 0.0}, sd, cd;
  sincos_dd(ad, &sd, &cd);
  *s = sd.hi + sd.lo;
  *c = cd.hi + cd.lo;
}

// correctly rounded atan2 for finite, non-zero y and x: one Newton step in double-double from the
// libm estimate:  t <- t + (y cos t - x sin t) / (x cos t + y sin t)
DDM_HD double cr_atan2(double y, double x) {
  const double t0 = atan2(y, x);
  if (x == 0.0 || y == 0.0 || !(t0 == t0) || isinf(x) || isinf(y)) return t0;
  dd t = {t0, 0.0}, s, c;
  sincos_dd(t, &s, &c);
  const dd num = add(mul_d(c, y), neg(mul_d(s, x)));
  const dd den = add(mul_d(c, x), mul_d(s, y));
  const dd corr = div(num, den);
  const dd r = add_d(corr, t0);
  return r.hi + r.lo;
}

}  // namespace ddm
