// interval.cuh — the interval arithmetic the reference's bound computations run on, for device code.
//
// The reference bounds factors with Boost.Interval under a policy WITHOUT directed rounding and WITHOUT checking
// (src/common.h:43-60: rounded_transc_exact + save_state_nothing + checking_base), i.e. plain double arithmetic driven
// by Boost's sign-case analysis.  Boost is not vendored in the reference tree, so the case analysis below restates the
// library's published algorithms (numeric/interval/arith.hpp, arith2.hpp, transc.hpp, detail/division.hpp,
// constants.hpp; Boost >= 1.55 per the reference README) — "parity unpinned" in the sense of DESIGN.md section 4: it is
// checked against the oracle's independent restatement and against the enclosure property, not against Boost itself.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace rdisgpu {

struct Ival {
  double lo, hi;
};

#define RDIS_IV __device__ __forceinline__

RDIS_IV Ival iv(double l, double u) { return Ival{l, u}; }
RDIS_IV Ival iv_point(double x) { return Ival{x, x}; }
RDIS_IV double iv_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
RDIS_IV double iv_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
RDIS_IV Ival iv_whole() { return Ival{-iv_inf(), iv_inf()}; }
RDIS_IV Ival iv_empty() { return Ival{iv_nan(), iv_nan()}; }
RDIS_IV double iv_width(Ival x) { return x.hi - x.lo; }
RDIS_IV double iv_median(Ival x) { return (x.lo + x.hi) / 2.0; }
RDIS_IV bool iv_zero_in(Ival x) { return !(x.lo > 0.0) && !(x.hi < 0.0); }

RDIS_IV Ival iv_neg(Ival x) { return iv(-x.hi, -x.lo); }
RDIS_IV Ival iv_add(Ival x, Ival y) { return iv(x.lo + y.lo, x.hi + y.hi); }
RDIS_IV Ival iv_add(Ival x, double y) { return iv(x.lo + y, x.hi + y); }
RDIS_IV Ival iv_sub(Ival x, Ival y) { return iv(x.lo - y.hi, x.hi - y.lo); }
RDIS_IV Ival iv_sub(Ival x, double y) { return iv(x.lo - y, x.hi - y); }

// operator*(interval, interval): arith.hpp's nine-case analysis on the signs of the bounds
RDIS_IV Ival iv_mul(Ival x, Ival y) {
  const double xl = x.lo, xu = x.hi, yl = y.lo, yu = y.hi;
  if (xl < 0.0) {
    if (xu > 0.0) {
      if (yl < 0.0) {
        if (yu > 0.0) return iv(fmin(xl * yu, xu * yl), fmax(xl * yl, xu * yu));  // M * M
        return iv(xu * yl, xl * yl);                                             // M * N
      }
      if (yu > 0.0) return iv(xl * yu, xu * yu);  // M * P
      return iv(0.0, 0.0);                        // M * Z
    }
    if (yl < 0.0) {
      if (yu > 0.0) return iv(xl * yu, xl * yl);  // N * M
      return iv(xu * yu, xl * yl);                // N * N
    }
    if (yu > 0.0) return iv(xl * yu, xu * yl);  // N * P
    return iv(0.0, 0.0);                        // N * Z
  }
  if (xu > 0.0) {
    if (yl < 0.0) {
      if (yu > 0.0) return iv(xu * yl, xu * yu);  // P * M
      return iv(xu * yl, xl * yu);                // P * N
    }
    if (yu > 0.0) return iv(xl * yl, xu * yu);  // P * P
    return iv(0.0, 0.0);                        // P * Z
  }
  return iv(0.0, 0.0);  // Z * ?
}
RDIS_IV Ival iv_mul(Ival x, double y) {
  if (y < 0.0) return iv(x.hi * y, x.lo * y);
  if (y == 0.0) return iv(0.0, 0.0);
  return iv(x.lo * y, x.hi * y);
}

// operator/(interval, interval): detail/division.hpp
RDIS_IV Ival iv_div_non_zero(Ival x, Ival y) {
  const double xl = x.lo, xu = x.hi, yl = y.lo, yu = y.hi;
  if (xu < 0.0) return (yu < 0.0) ? iv(xu / yl, xl / yu) : iv(xl / yl, xu / yu);
  if (xl < 0.0) return (yu < 0.0) ? iv(xu / yu, xl / yu) : iv(xl / yl, xu / yl);
  return (yu < 0.0) ? iv(xu / yu, xl / yl) : iv(xl / yu, xu / yl);
}
RDIS_IV Ival iv_div(Ival x, Ival y) {
  if (iv_zero_in(y)) {
    const bool x_zero = (x.lo == 0.0 && x.hi == 0.0);
    if (y.lo != 0.0) {
      if (y.hi != 0.0) return x_zero ? iv(0.0, 0.0) : iv_whole();  // div_zero
      // y = [yl, 0]: div_negative
      if (x_zero) return iv(0.0, 0.0);
      if (x.hi < 0.0) return iv(x.hi / y.lo, iv_inf());
      if (x.lo < 0.0) return iv_whole();
      return iv(-iv_inf(), x.lo / y.lo);
    }
    if (y.hi != 0.0) {  // y = [0, yu]: div_positive
      if (x_zero) return iv(0.0, 0.0);
      if (x.hi < 0.0) return iv(-iv_inf(), x.hi / y.hi);
      if (x.lo < 0.0) return iv_whole();
      return iv(x.lo / y.hi, iv_inf());
    }
    return iv_empty();
  }
  return iv_div_non_zero(x, y);
}
RDIS_IV Ival iv_div(Ival x, double y) {
  if (y == 0.0) return iv_empty();
  return (y < 0.0) ? iv(x.hi / y, x.lo / y) : iv(x.lo / y, x.hi / y);
}
RDIS_IV Ival iv_inverse(Ival x) {  // interval_lib::multiplicative_inverse
  if (iv_zero_in(x)) {
    if (x.lo != 0.0) {
      if (x.hi != 0.0) return iv_whole();
      return iv(-iv_inf(), 1.0 / x.lo);
    }
    if (x.hi != 0.0) return iv(1.0 / x.hi, iv_inf());
    return iv_empty();
  }
  return iv(1.0 / x.hi, 1.0 / x.lo);
}

RDIS_IV Ival iv_square(Ival x) {  // arith2.hpp
  if (x.hi < 0.0) return iv(x.hi * x.hi, x.lo * x.lo);
  if (x.lo > 0.0) return iv(x.lo * x.lo, x.hi * x.hi);
  return iv(0.0, (-x.lo > x.hi) ? x.lo * x.lo : x.hi * x.hi);
}
RDIS_IV Ival iv_sqrt(Ival x) {
  if (x.hi < 0.0) return iv_empty();
  return iv((x.lo <= 0.0) ? 0.0 : sqrt(x.lo), sqrt(x.hi));
}

// pow(interval, int): arith2.hpp, binary exponentiation of the bounds (pow_aux)
RDIS_IV double iv_pow_pos(double x_, int pwr) {
  double x = x_, y = (pwr & 1) ? x_ : 1.0;
  pwr >>= 1;
  while (pwr > 0) {
    x = x * x;
    if (pwr & 1) y = x * y;
    pwr >>= 1;
  }
  return y;
}
RDIS_IV Ival iv_pow_nonneg(Ival x, int pwr) {
  if (pwr == 0) return iv(1.0, 1.0);  // x^0 (Boost: interval<T>(1) when the interval is not empty)
  if (x.hi < 0.0) {
    const double yl = iv_pow_pos(-x.hi, pwr), yu = iv_pow_pos(-x.lo, pwr);
    return (pwr & 1) ? iv(-yu, -yl) : iv(yl, yu);
  }
  if (x.lo < 0.0) {
    if (pwr & 1) return iv(-iv_pow_pos(-x.lo, pwr), iv_pow_pos(x.hi, pwr));
    return iv(0.0, iv_pow_pos(fmax(-x.lo, x.hi), pwr));
  }
  return iv(iv_pow_pos(x.lo, pwr), iv_pow_pos(x.hi, pwr));
}
RDIS_IV Ival iv_pow(Ival x, int pwr) { return (pwr < 0) ? iv_inverse(iv_pow_nonneg(x, -pwr)) : iv_pow_nonneg(x, pwr); }

// constants.hpp (double): pi enclosed by two adjacent doubles
RDIS_IV double iv_pi_lo() { return (3373259426.0 + 273688.0 / 2097152.0) / 1073741824.0; }
RDIS_IV double iv_pi_hi() { return (3373259426.0 + 273689.0 / 2097152.0) / 1073741824.0; }

// transc.hpp: cos through fmod(x, 2 pi) and the monotone pieces of [0, 2 pi]; sin(x) = cos(x - pi/2)
RDIS_IV Ival iv_cos(Ival x) {
  const Ival pi2 = iv(iv_pi_lo() * 2.0, iv_pi_hi() * 2.0);
  // fmod(x, pi2) (arith2.hpp): n = floor(x.lo / (x.lo < 0 ? pi2.lo : pi2.hi)); x - n * pi2
  const double yb = (x.lo < 0.0) ? pi2.lo : pi2.hi;
  const double n = floor(x.lo / yb);
  Ival tmp = iv_sub(x, iv_mul(pi2, n));
  if (iv_width(tmp) >= pi2.lo) return iv(-1.0, 1.0);
  bool negate = false;
  if (tmp.lo >= iv_pi_hi()) {  // -cos(tmp - pi): the recursive call's own fmod is the identity (0 <= lower < 2 pi)
    tmp = iv_sub(tmp, iv(iv_pi_lo(), iv_pi_hi()));
    negate = true;
    if (iv_width(tmp) >= pi2.lo) return iv(-1.0, 1.0);
  }
  const double l = tmp.lo, u = tmp.hi;
  Ival r;
  if (u <= iv_pi_lo())
    r = iv(cos(u), cos(l));
  else if (u <= pi2.lo)
    r = iv(-1.0, cos(fmin(pi2.lo - u, l)));
  else
    r = iv(-1.0, 1.0);
  return negate ? iv_neg(r) : r;
}
RDIS_IV Ival iv_sin(Ival x) { return iv_cos(iv_sub(x, iv(iv_pi_lo() / 2.0, iv_pi_hi() / 2.0))); }

// power(NumericInterval, Numeric) of the reference, src/util/numeric.cpp:26-43 — including its quirk for negative
// exponents (the recursive call's result is discarded, so the INVERSE of the un-powered interval is returned) and the
// implicit double -> int conversion of the exponent in boost::numeric::pow(interval, int)
RDIS_IV Ival iv_rdis_power(Ival x, double e) {
  if (e == 0.0) return iv(1.0, 1.0);
  if (e == 1.0) return x;
  if (e == 2.0) return iv_square(x);
  if (e < 0.0) return iv_inverse(x);
  return iv_pow(x, (int)e);
}

}  // namespace rdisgpu
