// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the line-search / conjugate-gradient driver the reference
// uses for its default subspace optimizer.  The reference keeps this code in
//   /root/reference/external/include/minimize_nrc.h
// (Bracketmethod::bracket :81-151, Dbrent::minimize :300-404, Df1dim :410-448,
//  Dlinemethod::linmin :492-514, Frprmn::minimize :619-691).
//
// Parity pin: `make -C oracle ref` compiles the reference's own header (where it
// lies, unmodified) into oracle/_ref/liboracle_refnrc.so behind the same C API;
// tests/test_oracle_nr_pin.py demands bit-identical trajectories between this
// restatement and that build.  Every floating-point expression below therefore
// keeps the operand order of the cited lines.
#pragma once
#include <cmath>
#include <limits>
#include <utility>
#include <vector>

namespace oracle {
namespace nr {

typedef std::vector<double> Vec;

// std::max semantics (returns the first argument on ties / NaN in the second
// comparison) — minimize_nrc.h:60-63 forwards to std::max.
static inline double pick_max(double a, double b) { return (a < b) ? b : a; }

// What bracket() hands to the 1-D minimiser: three abscissas and their values.
struct Triple {
  double ax, bx, cx;
  double fa, fb, fc;
};

// minimize_nrc.h:81-151.  `line` is any callable double(double).
template <class Line>
Triple bracket_minimum(double a, double b, Line& line) {
  const double kGold = 1.618034;
  const double kMaxGrow = 100.0;
  const double kTiny = 1.0e-20;

  Triple t;
  t.ax = a;
  t.bx = b;
  t.fa = line(t.ax);
  t.fb = line(t.bx);
  if (t.fb > t.fa) {  // walk downhill from a to b (:92-95)
    std::swap(t.ax, t.bx);
    std::swap(t.fb, t.fa);
  }
  t.cx = t.bx + kGold * (t.bx - t.ax);  // :98
  t.fc = line(t.cx);

  while (t.fb > t.fc) {  // :101
    const double r = (t.bx - t.ax) * (t.fb - t.fc);
    const double q = (t.bx - t.cx) * (t.fb - t.fa);
    const double qmr = q - r;
    double u = t.bx - ((t.bx - t.cx) * q - (t.bx - t.ax) * r) /
                          (2.0 * std::copysign(pick_max(std::fabs(qmr), kTiny), qmr));  // :107-108
    const double ulim = t.bx + kMaxGrow * (t.cx - t.bx);                             // :109
    double fu;

    if ((t.bx - u) * (u - t.cx) > 0.0) {  // parabolic point between b and c (:112)
      fu = line(u);
      if (fu < t.fc) {  // minimum between b and c (:115-120)
        t.ax = t.bx;
        t.bx = u;
        t.fa = t.fb;
        t.fb = fu;
        return t;
      } else if (fu > t.fb) {  // minimum between a and u (:121-125)
        t.cx = u;
        t.fc = fu;
        return t;
      }
      u = t.cx + kGold * (t.cx - t.bx);  // :127
      fu = line(u);
    } else if ((t.cx - u) * (u - ulim) > 0.0) {  // between c and the limit (:129)
      fu = line(u);
      if (fu < t.fc) {
        // :133-134 — the new u uses the *old* cx (argument evaluated before the shift),
        // and the function is then evaluated at the new u.
        const double stepped = u + kGold * (u - t.cx);
        t.bx = t.cx;
        t.cx = u;
        u = stepped;
        const double fstepped = line(u);
        t.fb = t.fc;
        t.fc = fu;
        fu = fstepped;
      }
    } else if ((u - ulim) * (ulim - t.cx) >= 0.0) {  // clamp to the limit (:136-139)
      u = ulim;
      fu = line(u);
    } else {  // reject the parabola (:140-144)
      u = t.cx + kGold * (t.cx - t.bx);
      fu = line(u);
    }
    // discard the oldest point (:146-147)
    t.ax = t.bx;
    t.bx = t.cx;
    t.cx = u;
    t.fa = t.fb;
    t.fb = t.fc;
    t.fc = fu;
  }
  return t;
}

struct LineMin {
  double xmin;
  double fmin;
};

// Derivative-Brent, minimize_nrc.h:300-404.  `line(x)` returns f along the line and
// `line.slope(x)` the derivative at the point of the most recent line(x) call.
// Throws const char* exactly like the reference when 100 iterations do not converge.
template <class Line>
LineMin dbrent_minimize(const Triple& t, Line& line, double tol = 3.0e-8) {
  const int kMaxIter = 100;
  const double kZeps = std::numeric_limits<double>::epsilon() * 1.0e-3;

  double a = (t.ax < t.cx ? t.ax : t.cx);  // :312-313
  double b = (t.ax > t.cx ? t.ax : t.cx);
  double x, w, v;
  x = w = v = t.bx;
  double fx, fw, fv;
  fw = fv = fx = line(x);
  double dx, dw, dv;
  dw = dv = dx = line.slope(x);
  double d = 0.0, e = 0.0;
  double u, fu, du;

  for (int it = 0; it < kMaxIter; ++it) {
    const double xm = 0.5 * (a + b);
    const double tol1 = tol * std::fabs(x) + kZeps;
    const double tol2 = 2.0 * tol1;
    if (std::fabs(x - xm) <= (tol2 - 0.5 * (b - a))) {  // :324
      return LineMin{x, fx};
    }

    if (std::fabs(e) > tol1) {  // :329
      double d1 = 2.0 * (b - a);
      double d2 = d1;
      if (dw != dx) d1 = (w - x) * dx / (dx - dw);  // secant through w (:332)
      if (dv != dx) d2 = (v - x) * dx / (dx - dv);  // secant through v (:333)
      const double u1 = x + d1;
      const double u2 = x + d2;
      const bool ok1 = (a - u1) * (u1 - b) > 0.0 && dx * d1 <= 0.0;  // :339-340
      const bool ok2 = (a - u2) * (u2 - b) > 0.0 && dx * d2 <= 0.0;
      const double olde = e;
      e = d;
      if (ok1 || ok2) {
        if (ok1 && ok2)
          d = (std::fabs(d1) < std::fabs(d2) ? d1 : d2);
        else if (ok1)
          d = d1;
        else
          d = d2;
        if (std::fabs(d) <= std::fabs(0.5 * olde)) {  // :350
          u = x + d;
          if (u - a < tol2 || b - u < tol2) d = std::copysign(tol1, xm - x);
        } else {
          e = (dx >= 0.0 ? a - x : b - x);  // bisect (:355)
          d = 0.5 * e;
        }
      } else {
        e = (dx >= 0.0 ? a - x : b - x);  // :359
        d = 0.5 * e;
      }
    } else {
      e = (dx >= 0.0 ? a - x : b - x);  // :362
      d = 0.5 * e;
    }

    if (std::fabs(d) >= tol1) {  // :365
      u = x + d;
      fu = line(u);
    } else {
      u = x + std::copysign(tol1, d);
      fu = line(u);
      if (fu > fx) {  // smallest downhill step goes uphill: done (:371-376)
        return LineMin{x, fx};
      }
    }
    du = line.slope(u);  // :379

    if (fu <= fx) {  // :382-389
      if (u >= x)
        a = x;
      else
        b = x;
      v = w; fv = fw; dv = dw;
      w = x; fw = fx; dw = dx;
      x = u; fx = fu; dx = du;
    } else {  // :390-401
      if (u < x)
        a = u;
      else
        b = u;
      if (fu <= fw || w == x) {
        v = w; fv = fw; dv = dw;
        w = u; fw = fu; dw = du;
      } else if (fu < fv || v == x || v == w) {
        v = u; fv = fu; dv = du;
      }
    }
  }
  throw("Too many iterations in routine dbrent");  // :403
}

// The 1-D restriction of an n-D function, minimize_nrc.h:410-448.
// FD must offer  double operator()(const Vec&)  and  void df(const Vec&, Vec&).
template <class FD>
struct LineRestriction {
  const Vec& origin;
  const Vec& dir;
  FD& fd;
  Vec trial;
  Vec grad;
  LineRestriction(const Vec& p, const Vec& xi, FD& f)
      : origin(p), dir(xi), fd(f), trial(p.size()), grad(p.size()) {}
  double operator()(double s) {
    const size_t n = origin.size();
    for (size_t j = 0; j < n; ++j) trial[j] = origin[j] + s * dir[j];  // :434
    return fd(trial);
  }
  double slope(double /*s: always the point of the latest operator() call*/) {
    double acc = 0.0;
    fd.df(trial, grad);  // :441
    const size_t n = origin.size();
    for (size_t j = 0; j < n; ++j) acc += grad[j] * dir[j];  // :442-443
    return acc;
  }
};

// Polak–Ribière conjugate gradient, minimize_nrc.h:586-692.  State is public because
// the caller (CGDSubspaceOptimizer.cpp:40-64) reads p / fret / iter after a throw.
template <class FD>
struct PolakRibiere {
  FD& fd;
  Vec p, xi;
  int iter;
  double fret;
  const double ftol;
  const int maxiters;

  PolakRibiere(FD& f, int maxit = 300, double tol = 3.0e-8)
      : fd(f), iter(0), fret(std::numeric_limits<double>::max()), ftol(tol), maxiters(maxit) {}

  // Dlinemethod::linmin, :492-514
  double line_minimise() {
    LineRestriction<FD> line(p, xi, fd);
    Triple t = bracket_minimum(0.0, 1.0, line);
    LineMin m = dbrent_minimize(t, line);
    const size_t n = p.size();
    for (size_t j = 0; j < n; ++j) {
      xi[j] *= m.xmin;
      p[j] += xi[j];
    }
    return m.fmin;
  }

  Vec minimize(const Vec& start) {
    const double kEps = 1.0e-18;
    const double kGtol = 1.0e-8;
    const size_t n = start.size();
    p = start;
    Vec g(n), h(n);
    xi.resize(n);
    double fp = fd(p);  // :634
    fd.df(p, xi);       // :635
    for (size_t j = 0; j < n; ++j) {
      g[j] = -xi[j];
      xi[j] = h[j] = g[j];
    }
    for (int its = 0; its < maxiters; ++its) {
      iter = its;
      fret = line_minimise();  // :646
      if (2.0 * std::fabs(fret - fp) <= ftol * (std::fabs(fret) + std::fabs(fp) + kEps)) {  // :648-649
        return p;
      }
      fp = fret;
      fd.df(p, xi);  // :654
      double test = 0.0;
      const double den = pick_max(std::fabs(fp), 1.0);
      for (size_t j = 0; j < n; ++j) {
        const double temp = std::fabs(xi[j]) * pick_max(std::fabs(p[j]), 1.0) / den;  // :659
        if (temp > test) test = temp;
      }
      if (test < kGtol) return p;  // :663
      double gg = 0.0, dgg = 0.0;
      for (size_t j = 0; j < n; ++j) {
        gg += g[j] * g[j];
        dgg += (xi[j] + g[j]) * xi[j];  // :672, Polak–Ribière
      }
      if (gg == 0.0) return p;  // :676
      const double gam = dgg / gg;
      for (size_t j = 0; j < n; ++j) {
        g[j] = -xi[j];
        xi[j] = h[j] = g[j] + gam * h[j];
      }
    }
    throw("Too many iterations in frprmn");  // :690 — the normal exit at maxiters
  }
};

}  // namespace nr
}  // namespace oracle
