// interval_oracle.hpp — ORACLE (test infrastructure, never on the product path): the reference's interval bounds.
//   Factor::computeBounds                              /root/reference/src/Factor.cpp:122-139
//   OptimizableFunction::computeBounds                 src/OptimizableFunction.cpp:186-216
//   NonlinearProductFactor::computeFactorBounds        src/NonlinearProductFactor.cpp:120-145
//   power(NumericInterval, Numeric)                    src/util/numeric.cpp:26-43
//   BundleAdjustmentFactor::computeFactorBounds / getVarVals / evalFactor(IntervalVec) / angleAxisRotatePoint(IntervalVec)
//                                                      src/bundleadjust/BundleAdjustmentFactor.cpp:47-52,67-93,104-157,186-232
//   simpleNormalize                                    src/bundleadjust/BundleAdjustmentCommon.cpp:43-53
// PARITY UNPINNED: NumericInterval is boost::numeric::interval<double> under the no-rounding / no-checking policy of
// src/common.h:43-60; Boost (>= 1.55, version not pinned by the reference, README.md:38-40) is not vendored under
// /root/reference and is not installed here.  The class below restates the library's published algorithms
// (numeric/interval/arith.hpp, arith2.hpp, transc.hpp, detail/division.hpp, constants.hpp) in terms of plain double
// operations, which is what that policy reduces them to.  The reference holds no golden bound values; the tests pin
// this file by the ENCLOSURE property (every sampled point value lies inside the bound), hand-derived cases, and an
// independent interval library (mpmath.iv) evaluating the same expressions (tests/test_bounds.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

#include "rdis_oracle.hpp"

namespace oracle {

class Interval {
 public:
  Interval() : l_(0), u_(0) {}
  Interval(double v) : l_(v), u_(v) {}  // NOLINT: implicit like boost's interval(T const&)
  Interval(double l, double u) : l_(l), u_(u) {}
  double lower() const { return l_; }
  double upper() const { return u_; }
  static Interval whole() { return Interval(-std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()); }
  static Interval empty() { return Interval(std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::quiet_NaN()); }

 private:
  double l_, u_;
};

namespace ivl {
inline bool neg(double x) { return x < 0.0; }
inline bool pos(double x) { return x > 0.0; }
inline bool zero(double x) { return x == 0.0; }
inline bool zero_in(const Interval& x) { return !pos(x.lower()) && !neg(x.upper()); }
inline double width(const Interval& x) { return x.upper() - x.lower(); }
inline double median(const Interval& x) { return (x.lower() + x.upper()) / 2.0; }
}  // namespace ivl

inline Interval operator-(const Interval& x) { return Interval(-x.upper(), -x.lower()); }
inline Interval operator+(const Interval& x, const Interval& y) { return Interval(x.lower() + y.lower(), x.upper() + y.upper()); }
inline Interval operator-(const Interval& x, const Interval& y) { return Interval(x.lower() - y.upper(), x.upper() - y.lower()); }
inline Interval sub_scalar(const Interval& x, double y) { return Interval(x.lower() - y, x.upper() - y); }

// arith.hpp operator*(interval, interval)
inline Interval operator*(const Interval& x, const Interval& y) {
  using namespace ivl;
  const double xl = x.lower(), xu = x.upper(), yl = y.lower(), yu = y.upper();
  if (neg(xl)) {
    if (pos(xu)) {
      if (neg(yl)) {
        if (pos(yu)) return Interval(std::min(xl * yu, xu * yl), std::max(xl * yl, xu * yu));
        return Interval(xu * yl, xl * yl);
      }
      if (pos(yu)) return Interval(xl * yu, xu * yu);
      return Interval(0.0, 0.0);
    }
    if (neg(yl)) {
      if (pos(yu)) return Interval(xl * yu, xl * yl);
      return Interval(xu * yu, xl * yl);
    }
    if (pos(yu)) return Interval(xl * yu, xu * yl);
    return Interval(0.0, 0.0);
  }
  if (pos(xu)) {
    if (neg(yl)) {
      if (pos(yu)) return Interval(xu * yl, xu * yu);
      return Interval(xu * yl, xl * yu);
    }
    if (pos(yu)) return Interval(xl * yl, xu * yu);
    return Interval(0.0, 0.0);
  }
  return Interval(0.0, 0.0);
}
// arith.hpp operator*(interval, T)
inline Interval mul_scalar(const Interval& x, double y) {
  if (ivl::neg(y)) return Interval(x.upper() * y, x.lower() * y);
  if (ivl::zero(y)) return Interval(0.0, 0.0);
  return Interval(x.lower() * y, x.upper() * y);
}
// detail/division.hpp
inline Interval operator/(const Interval& x, const Interval& y) {
  using namespace ivl;
  const double xl = x.lower(), xu = x.upper(), yl = y.lower(), yu = y.upper();
  const double inf = std::numeric_limits<double>::infinity();
  if (zero_in(y)) {
    const bool xz = zero(xl) && zero(xu);
    if (!zero(yl)) {
      if (!zero(yu)) return xz ? Interval(0.0, 0.0) : Interval::whole();  // div_zero
      if (xz) return Interval(0.0, 0.0);                                  // div_negative
      if (neg(xu)) return Interval(xu / yl, inf);
      if (neg(xl)) return Interval::whole();
      return Interval(-inf, xl / yl);
    }
    if (!zero(yu)) {  // div_positive
      if (xz) return Interval(0.0, 0.0);
      if (neg(xu)) return Interval(-inf, xu / yu);
      if (neg(xl)) return Interval::whole();
      return Interval(xl / yu, inf);
    }
    return Interval::empty();
  }
  if (neg(xu)) return neg(yu) ? Interval(xu / yl, xl / yu) : Interval(xl / yl, xu / yu);
  if (neg(xl)) return neg(yu) ? Interval(xu / yu, xl / yu) : Interval(xl / yl, xu / yl);
  return neg(yu) ? Interval(xu / yu, xl / yl) : Interval(xl / yu, xu / yl);
}
inline Interval div_scalar(const Interval& x, double y) {
  if (ivl::zero(y)) return Interval::empty();
  return ivl::neg(y) ? Interval(x.upper() / y, x.lower() / y) : Interval(x.lower() / y, x.upper() / y);
}
inline Interval multiplicative_inverse(const Interval& x) {
  using namespace ivl;
  const double inf = std::numeric_limits<double>::infinity();
  if (zero_in(x)) {
    if (!zero(x.lower())) {
      if (!zero(x.upper())) return Interval::whole();
      return Interval(-inf, 1.0 / x.lower());
    }
    if (!zero(x.upper())) return Interval(1.0 / x.upper(), inf);
    return Interval::empty();
  }
  return Interval(1.0 / x.upper(), 1.0 / x.lower());
}
// arith2.hpp
inline Interval square(const Interval& x) {
  const double xl = x.lower(), xu = x.upper();
  if (ivl::neg(xu)) return Interval(xu * xu, xl * xl);
  if (ivl::pos(xl)) return Interval(xl * xl, xu * xu);
  return Interval(0.0, (-xl > xu) ? xl * xl : xu * xu);
}
inline Interval sqrt(const Interval& x) {
  if (ivl::neg(x.upper())) return Interval::empty();
  return Interval(!ivl::pos(x.lower()) ? 0.0 : std::sqrt(x.lower()), std::sqrt(x.upper()));
}
inline double pow_positive(double base, int pwr) {  // detail::pow_dn / pow_up without rounding
  double x = base, y = (pwr & 1) ? base : 1.0;
  for (pwr >>= 1; pwr > 0; pwr >>= 1) {
    x = x * x;
    if (pwr & 1) y = x * y;
  }
  return y;
}
inline Interval pow(const Interval& x, int pwr) {
  if (pwr < 0) return multiplicative_inverse(pow(x, -pwr));
  if (pwr == 0) return Interval(1.0, 1.0);
  const double xl = x.lower(), xu = x.upper();
  if (ivl::neg(xu)) {
    const double yl = pow_positive(-xu, pwr), yu = pow_positive(-xl, pwr);
    return (pwr & 1) ? Interval(-yu, -yl) : Interval(yl, yu);
  }
  if (ivl::neg(xl)) {
    if (pwr & 1) return Interval(-pow_positive(-xl, pwr), pow_positive(xu, pwr));
    return Interval(0.0, pow_positive(std::max(-xl, xu), pwr));
  }
  return Interval(pow_positive(xl, pwr), pow_positive(xu, pwr));
}
// constants.hpp: the two doubles around pi
inline double pi_lower() { return (3373259426.0 + 273688.0 / (1 << 21)) / (1 << 30); }
inline double pi_upper() { return (3373259426.0 + 273689.0 / (1 << 21)) / (1 << 30); }
// transc.hpp
inline Interval cos(const Interval& x) {
  const Interval pi2(pi_lower() * 2.0, pi_upper() * 2.0);
  // arith2.hpp fmod(x, pi2)
  const double yb = ivl::neg(x.lower()) ? pi2.lower() : pi2.upper();
  const double n = std::floor(x.lower() / yb);
  const Interval tmp = x - mul_scalar(pi2, n);
  if (ivl::width(tmp) >= pi2.lower()) return Interval(-1.0, 1.0);
  if (tmp.lower() >= pi_upper()) return -cos(tmp - Interval(pi_lower(), pi_upper()));
  const double l = tmp.lower(), u = tmp.upper();
  if (u <= pi_lower()) return Interval(std::cos(u), std::cos(l));
  if (u <= pi2.lower()) return Interval(-1.0, std::cos(std::min(pi2.lower() - u, l)));
  return Interval(-1.0, 1.0);
}
inline Interval sin(const Interval& x) { return cos(x - Interval(pi_lower() / 2.0, pi_upper() / 2.0)); }

// src/util/numeric.cpp:26-43 (the negative-exponent branch discards the recursive result: kept as is)
inline Interval power(Interval ival, Numeric exp) {
  if (exp == 0.) {
    ival = Interval(1., 1.);
  } else if (exp == 1.) {
  } else if (exp == 2.) {
    ival = square(ival);
  } else if (exp < 0) {
    power(ival, -exp);
    ival = multiplicative_inverse(ival);
  } else {
    ival = pow(ival, (int)exp);  // boost::numeric::pow(interval, int): the double exponent converts implicitly
  }
  return ival;
}

// Variable as the bound computation sees it: point if assigned (and not ignored), domain hull otherwise.
template <class IsPoint>
inline Interval var_interval(const Variable& v, IsPoint is_point) {
  return is_point(v) ? Interval(v.eval()) : Interval(v.getDomain().lo, v.getDomain().hi);
}

// src/NonlinearProductFactor.cpp:120-145
template <class IsPoint>
inline Interval nlpf_factor_bounds(const NonlinearProductFactor& f, IsPoint is_point) {
  Interval feval(1);
  const auto& vars = f.getVariables();
  for (size_t i = 0; i < vars.size(); ++i) {
    Interval val = var_interval(*vars[i], is_point);
    const auto& t = f.terms()[i];
    if (t.hasConstant()) val = sub_scalar(val, t.constant);
    if (t.hasExp()) val = power(val, t.exponent);
    if (t.useSine) val = sin(val);
    feval = feval * val;
  }
  return mul_scalar(feval, f.coeff());
}

// src/bundleadjust/BundleAdjustmentFactor.cpp:104-157 + 186-232
template <class IsPoint>
inline Interval ba_factor_bounds(const BundleAdjustmentFactor& f, IsPoint is_point) {
  Interval vals[BA_NSLOTS];
  for (int i = 0; i < BA_NSLOTS; ++i) vals[i] = var_interval(*f.getVariables()[i], is_point);
  Interval p[3] = {vals[PT_X], vals[PT_X + 1], vals[PT_X + 2]};
  const Interval vun[3] = {vals[0], vals[1], vals[2]};
  const Interval theta = sqrt(square(vun[0]) + square(vun[1]) + square(vun[2]));
  Interval v[3];
  {
    const Interval norm = sqrt(square(vun[0]) + square(vun[1]) + square(vun[2]));
    for (int i = 0; i < 3; ++i) v[i] = vun[i] / norm;
  }
  Interval costheta, sintheta;
  if (ivl::width(theta) < 1e-6) {
    costheta = Interval(std::cos(ivl::median(theta)));
    sintheta = Interval(std::sin(ivl::median(theta)));
  } else {
    costheta = cos(theta);
    sintheta = sin(theta);
  }
  const Interval oneminus = Interval(1) - costheta;
  Interval vxp[3];
  vxp[0] = v[1] * p[2] - v[2] * p[1];
  vxp[1] = v[2] * p[0] - v[0] * p[2];
  vxp[2] = v[0] * p[1] - v[1] * p[0];
  const Interval vdp = v[0] * p[0] + v[1] * p[1] + v[2] * p[2];
  Interval q[3];
  for (int i = 0; i < 3; ++i) q[i] = p[i] * costheta + vxp[i] * sintheta + v[i] * oneminus * vdp;
  q[0] = q[0] + vals[3];
  q[1] = q[1] + vals[4];
  q[2] = q[2] + vals[5];
  Interval px = -q[0] / q[2];
  Interval py = -q[1] / q[2];
  const Interval r2 = square(px) + square(py);
  const Interval dstn = Interval(1) + vals[7] * r2 + vals[8] * square(r2);
  px = vals[6] * dstn * px;
  py = vals[6] * dstn * py;
  const Interval ex = square(sub_scalar(px, f.obsX()));
  const Interval ey = square(sub_scalar(py, f.obsY()));
  return div_scalar(ex + ey, 2.0);
}

}  // namespace oracle
