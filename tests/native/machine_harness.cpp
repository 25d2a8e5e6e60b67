// Test harness (CPU): drives rdis_b200/csrc/cgd_machine.cuh — the resumable state machine the GPU
// thread groups run — with a single "lane" and a plain callback objective, and compares the whole
// trajectory bit-for-bit with the nested-call driver of the oracle (oracle/nr_minimize.hpp, itself
// pinned bit-for-bit against the reference's own external/include/minimize_nrc.h).
// With -DUSE_REFERENCE_NRC (-I/root/reference/external/include) the comparison is made directly
// against the reference header.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../rdis_b200/csrc/cgd_machine.cuh"
#ifdef USE_REFERENCE_NRC
namespace rdis { typedef double Numeric; }
#include "minimize_nrc.h"
#else
#include "../../oracle/nr_minimize.hpp"
#endif

typedef std::vector<double> Vec;

struct TestFunction {
  int kind;
  int n;
  Vec lo, hi, w;
  long nf = 0, ng = 0;
  double clampv(int j, double v) const { return (lo[j] <= v && v <= hi[j]) ? v : (v < lo[j] ? lo[j] : hi[j]); }
  double operator()(const Vec& xr) {
    ++nf;
    Vec x(n);
    for (int j = 0; j < n; ++j) x[j] = clampv(j, xr[j]);
    double s = 0;
    switch (kind) {
      case 0:  // Rosenbrock chain
        for (int j = 0; j + 1 < n; ++j) s += 100 * (x[j + 1] - x[j] * x[j]) * (x[j + 1] - x[j] * x[j]) + (1 - x[j]) * (1 - x[j]);
        if (n == 1) s = (1 - x[0]) * (1 - x[0]);
        break;
      case 1:  // weighted quadratic + quartic
        for (int j = 0; j < n; ++j) s += w[j] * x[j] * x[j] + 0.01 * x[j] * x[j] * x[j] * x[j] + 0.3 * x[j];
        break;
      case 2:  // sinusoid chain (non-convex)
        for (int j = 0; j < n; ++j) s += 0.6 * x[j] + 0.1 * x[j] * x[j];
        for (int j = 0; j + 1 < n; ++j) s += 12 * std::sin(x[j]) * std::sin(x[j + 1]);
        break;
      default:  // linear: unbounded below except for the clamp
        for (int j = 0; j < n; ++j) s += w[j] * x[j];
    }
    return s;
  }
  void df(const Vec& xr, Vec& g) {
    ++ng;
    Vec x(n);
    for (int j = 0; j < n; ++j) x[j] = clampv(j, xr[j]);
    g.assign(n, 0.0);
    switch (kind) {
      case 0:
        for (int j = 0; j + 1 < n; ++j) {
          const double t = x[j + 1] - x[j] * x[j];
          g[j] += -400 * t * x[j] - 2 * (1 - x[j]);
          g[j + 1] += 200 * t;
        }
        if (n == 1) g[0] = -2 * (1 - x[0]);
        break;
      case 1:
        for (int j = 0; j < n; ++j) g[j] = 2 * w[j] * x[j] + 0.04 * x[j] * x[j] * x[j] + 0.3;
        break;
      case 2:
        for (int j = 0; j < n; ++j) g[j] = 0.6 + 0.2 * x[j];
        for (int j = 0; j + 1 < n; ++j) {
          g[j] += 12 * std::cos(x[j]) * std::sin(x[j + 1]);
          g[j + 1] += 12 * std::sin(x[j]) * std::cos(x[j + 1]);
        }
        break;
      default:
        for (int j = 0; j < n; ++j) g[j] = w[j];
    }
  }
};

// The reference's value cache around the objective, as one "factor" over all variables: Variable::assign only
// notifies (marks the value dirty) when a coordinate moves by >= 1e-12 (src/Variable.cpp:66-88); operator() returns
// the cached value unless dirty (src/Factor.cpp:110-119); df() always uses the assigned point.  With it the
// objective is NOT a function of the point alone, which is what CgdMachine's `faithful` mode exists for.
struct CachedFunction {
  TestFunction fn;
  int n;
  Vec last;
  bool assigned = false, dirty = true;
  double cached = 0;
  long nf = 0, ng = 0;
  explicit CachedFunction(const TestFunction& f) : fn(f), n(f.n), last(f.n, 0.0) {}
  void assign(const Vec& xr) {
    for (int j = 0; j < n; ++j) {
      const double v = fn.clampv(j, xr[j]);
      if (!assigned || !(std::fabs(v - last[j]) < 1e-12)) dirty = true;
      last[j] = v;
    }
    assigned = true;
  }
  double operator()(const Vec& xr) {
    ++nf;
    assign(xr);
    if (dirty) {
      cached = fn(last);
      dirty = false;
    }
    return cached;
  }
  void df(const Vec& xr, Vec& g) {
    ++ng;
    assign(xr);
    fn.df(last, g);
  }
};

struct Outcome {
  Vec p;
  double fret;
  int iter;
  long nf, ng;
};

template <class Fn>
static Outcome run_nested(Fn fn, const Vec& x0, int maxiters, double ftol) {
#ifdef USE_REFERENCE_NRC
  rdis::nrc::Frprmn<Fn> cg(fn, maxiters, ftol);
#else
  oracle::nr::PolakRibiere<Fn> cg(fn, maxiters, ftol);
#endif
  try {
    cg.minimize(x0);
  } catch (const char*) {
  }
  return Outcome{cg.p, cg.fret, cg.iter, fn.nf, fn.ng};
}

// the kernel's loop (solve_kernels.cuh: solve_problem) with one lane
template <class Fn>
static Outcome run_machine(Fn fn, const Vec& x0, int maxiters, double ftol, bool faithful) {
  using namespace rdisgpu;
  const int n = fn.n;
  Vec p = x0, xi(n), g(n), h(n), trial(n), grad(n);
  CgdMachine m;
  m.start(maxiters, ftol, faithful);
  while (!m.done()) {
    switch (m.req) {
      case REQ_INIT_GRAD: {
        const double f = fn(p);
        fn.df(p, grad);
        for (int j = 0; j < n; ++j) { g[j] = -grad[j]; xi[j] = h[j] = g[j]; }
        m.on_init(f);
        break;
      }
      case REQ_VALUE:
      case REQ_VALUE_SLOPE: {
        for (int j = 0; j < n; ++j) trial[j] = p[j] + m.alpha * xi[j];
        const double f = fn(trial);
        double sl = 0.0;
        if (m.req == REQ_VALUE_SLOPE) {
          fn.df(trial, grad);
          for (int j = 0; j < n; ++j) sl += grad[j] * xi[j];
        }
        m.on_eval(f, sl);
        break;
      }
      case REQ_MOVE:
        for (int j = 0; j < n; ++j) { xi[j] *= m.alpha; p[j] += xi[j]; }
        m.on_moved();
        break;
      case REQ_GRADIENT: {
        fn.df(p, xi);
        double tnum = 0, gg = 0, dgg = 0;
        for (int j = 0; j < n; ++j) {
          const double pj = std::fabs(p[j]);
          const double t = std::fabs(xi[j]) * (pj < 1.0 ? 1.0 : pj);
          if (t > tnum) tnum = t;
          gg += g[j] * g[j];
          dgg += (xi[j] + g[j]) * xi[j];
        }
        m.on_gradient(tnum, gg, dgg);
        break;
      }
      case REQ_DIRECTION:
        for (int j = 0; j < n; ++j) { g[j] = -xi[j]; xi[j] = h[j] = g[j] + m.gam * h[j]; }
        m.on_directed();
        break;
      default:
        std::abort();
    }
  }
  return Outcome{p, m.fret, m.iter, fn.nf, fn.ng};
}

int main(int argc, char** argv) {
  const int trials = argc > 1 ? std::atoi(argv[1]) : 400;
  const bool cached = argc > 2 && std::atoi(argv[2]) != 0;  // objective behind the reference's value cache
  int plain_machine_differs = 0;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(-1, 1);
  int bad = 0;
  long total_nested = 0, total_machine = 0;
  for (int t = 0; t < trials; ++t) {
    TestFunction fn;
    fn.kind = t % 4;
    fn.n = 1 + (int)(rng() % 12);
    const double box = (fn.kind == 3) ? 50.0 : (t % 7 == 0 ? 1.5 : 30.0);
    fn.lo.assign(fn.n, -box);
    fn.hi.assign(fn.n, box);
    fn.w.resize(fn.n);
    for (double& w : fn.w) w = 0.2 + 2.0 * std::fabs(U(rng));
    Vec x0(fn.n);
    for (double& v : x0) v = (fn.kind == 2 ? 6.0 : 2.0) * U(rng);
    const int maxiters = (t % 5 == 0) ? 3 : 25;
    Outcome a, b;
    if (cached) {
      // the start point is assigned and evaluated by the caller first (CGDSubspaceOptimizer.cpp:33-37)
      CachedFunction cf(fn);
      cf(x0);
      a = run_nested(cf, x0, maxiters, 3e-8);
      b = run_machine(cf, x0, maxiters, 3e-8, true);
      const Outcome c = run_machine(cf, x0, maxiters, 3e-8, false);
      bool same_c = (a.fret == c.fret) && (a.iter == c.iter);
      for (int j = 0; j < fn.n; ++j) same_c = same_c && (a.p[j] == c.p[j]);
      if (!same_c) ++plain_machine_differs;
    } else {
      a = run_nested(fn, x0, maxiters, 3e-8);
      b = run_machine(fn, x0, maxiters, 3e-8, false);
    }
    bool same = (a.fret == b.fret) && (a.iter == b.iter);
    for (int j = 0; j < fn.n; ++j) same = same && (a.p[j] == b.p[j]);
    if (!same) {
      ++bad;
      if (bad < 6) std::printf("MISMATCH trial %d kind %d n %d: fret %.17g vs %.17g iter %d vs %d\n", t, fn.kind, fn.n, a.fret, b.fret, a.iter, b.iter);
    }
    total_nested += a.nf + a.ng;
    total_machine += b.nf + b.ng;
  }
  std::printf("trials %d mismatches %d  evaluations nested %ld machine %ld\n", trials, bad, total_nested, total_machine);
  if (cached) std::printf("cached objective: the non-faithful machine departs from the reference on %d trials\n", plain_machine_differs);
  return bad == 0 ? 0 : 1;
}
