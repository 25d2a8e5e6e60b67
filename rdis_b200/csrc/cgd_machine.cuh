// cgd_machine.cuh — the reference's Polak–Ribière / derivative-Brent subspace solve, restated as a
// resumable state machine so that it can live inside a GPU thread group.
//
// The reference runs the solve as nested calls with a callback functor
//   Frprmn::minimize -> Dlinemethod::linmin -> Bracketmethod::bracket / Dbrent::minimize -> Df1dim
//   (external/include/minimize_nrc.h:619-691, 492-514, 81-151, 300-404, 410-448)
// On the GPU the "callback" is a cooperative sweep over the component's factors by a whole thread
// group (sub-warp tile, CTA, or cooperative grid), so control is inverted: the machine publishes a
// request (evaluate the objective at p + alpha*xi, with or without the slope; move p; rebuild the
// direction), the group performs it, and feeds the reduced scalars back.  Every lane of the group
// runs an identical copy of the machine on identical inputs, so no broadcast is needed.
//
// Evaluations the reference repeats at an already-evaluated point (fp=func(p) at :634 after the
// caller's own evaluation, fa=func(0) at :88, fx=funcd(bx) at :315) are not re-requested where the
// objective is a deterministic function of the point.  It is NOT one under the reference's value cache:
// Variable::assign suppresses the dirty notification of a move below 1e-12 (src/Variable.cpp:69-73), so
// Factor::eval may return the value of an earlier, nearby point (src/Factor.cpp:110-119), and the
// re-evaluation at :88 may then see a different number than the line search left in fret.  A machine started
// `faithful` therefore requests fa = func(0) as well (PH_BR_FA), and every kernel takes fx at :315 from the
// evaluation that comes back, so a group that emulates the cache (every solve kernel does) follows the
// reference's evaluation sequence event for event.  All scalar arithmetic keeps the operand order of the
// cited lines.
//
// The header is __host__ __device__ clean: tests/native/machine_harness.cpp compiles it with g++ and
// demands bit-identical request sequences against the reference's own header (oracle/_ref).
#pragma once

#if defined(__CUDACC__)
#define RDIS_HD __host__ __device__ __forceinline__
#else
#define RDIS_HD inline
#endif

#include <math.h>

namespace rdisgpu {

enum MachineRequest : int {
  REQ_INIT_GRAD = 0,  // value + full gradient at p; then xi = h = g = -grad          (:634-641)
  REQ_VALUE = 1,      // value at p + alpha*xi                                          (Df1dim::operator(), :432-436)
  REQ_VALUE_SLOPE = 2,// value and directional derivative at p + alpha*xi               (:432-447)
  REQ_MOVE = 3,       // p += alpha*xi (alpha = xmin of the line search)                (:508-511)
  REQ_GRADIENT = 4,   // full gradient at p into xi; report test / gg / dgg             (:654-673)
  REQ_DIRECTION = 5,  // g = -xi; xi = h = g + gam*h                                    (:681-685)
  REQ_DONE = 6
};

enum MachineStatus : int {
  ST_FTOL = 0, ST_GTOL = 1, ST_GG_ZERO = 2, ST_MAXITERS = 3, ST_DBRENT_ITMAX = 4,
  ST_EMPTY = 5, ST_NONFINITE = 6, ST_BRACKET_CAP = 7
};

struct CgdMachine {
  // ---- request published to the group ----
  int req;
  double alpha;  // abscissa for REQ_VALUE / REQ_VALUE_SLOPE, step for REQ_MOVE
  double gam;    // for REQ_DIRECTION

  // ---- results ----
  double fret;   // Frprmn::fret (DBL_MAX until the first line search finishes, :609)
  int iter;      // Frprmn::iter
  int status;
  int n_value;   // evaluations requested (value only)
  int n_slope;   // evaluations requested with derivatives

  // ---- Frprmn ----
  double fp, ftol;
  int its, maxiters;

  // ---- bracket / dbrent share storage where lifetimes do not overlap ----
  double ax, bx, cx, fa, fb, fc;       // bracket triple
  double u, fu;                        // trial abscissa / value
  double a, b, d, e;                   // dbrent interval and step memory
  double v, w, x, fv, fw, fx, dv, dw, dx;
  int phase;
  int db_iter;
  int br_iter;
  bool small_step;
  bool faithful;  // request fa = func(ax = 0) at the top of every line search (minimize_nrc.h:88)

  enum Phase : int {
    PH_INIT = 0,
    PH_BR_FA, PH_BR_FB, PH_BR_FC, PH_BR_PARAB_INSIDE, PH_BR_PARAB_BEYOND, PH_BR_PARAB_BEYOND2, PH_BR_SHIFT,
    PH_DB_FIRST, PH_DB_EVAL,
    PH_MOVED, PH_GRAD, PH_DIRECTED
  };

  static constexpr int kBracketCap = 2000;  // device safety net; the reference's loop is unbounded (:101)

  RDIS_HD static double pick_max(double p, double q) { return (p < q) ? q : p; }  // std::max, :60-63

  RDIS_HD void start(int maxiters_, double ftol_, bool faithful_ = false) {
    faithful = faithful_;
    maxiters = maxiters_;
    ftol = ftol_;
    fret = 1.7976931348623157e308;  // std::numeric_limits<double>::max(), :609
    iter = 0;
    its = 0;
    status = ST_MAXITERS;
    n_value = 0;
    n_slope = 0;
    phase = PH_INIT;
    req = REQ_INIT_GRAD;
    alpha = 0.0;
    gam = 0.0;
    ++n_slope;
  }

  // ---- feed-back entry points -----------------------------------------------------------
  // after REQ_INIT_GRAD: f = objective at the start point
  RDIS_HD void on_init(double f) {
    fp = f;
    begin_line_search();
  }
  // after REQ_VALUE / REQ_VALUE_SLOPE
  RDIS_HD void on_eval(double f, double slope) { advance(f, slope); }
  // after REQ_MOVE
  RDIS_HD void on_moved() {
    // :648-652
    if (2.0 * fabs(fret - fp) <= ftol * (fabs(fret) + fabs(fp) + 1.0e-18)) {
      finish(ST_FTOL);
      return;
    }
    fp = fret;
    req = REQ_GRADIENT;
    ++n_slope;
    phase = PH_GRAD;
  }
  // after REQ_GRADIENT: test_num = max_j |xi_j| * max(|p_j|, 1); gg, dgg as in :670-673
  RDIS_HD void on_gradient(double test_num, double gg, double dgg) {
    const double den = pick_max(fabs(fp), 1.0);
    const double test = test_num / den;  // monotone in the numerator, so max-then-divide == divide-then-max (:657-661)
    if (test < 1.0e-8) {
      finish(ST_GTOL);
      return;
    }
    if (gg == 0.0) {
      finish(ST_GG_ZERO);
      return;
    }
    gam = dgg / gg;
    req = REQ_DIRECTION;
    phase = PH_DIRECTED;
  }
  // after REQ_DIRECTION
  RDIS_HD void on_directed() {
    ++its;
    if (its >= maxiters) {
      finish(ST_MAXITERS);  // throw("Too many iterations in frprmn"), :690
      return;
    }
    begin_line_search();
  }

  RDIS_HD bool done() const { return req == REQ_DONE; }

 private:
  RDIS_HD void finish(int st) {
    status = st;
    req = REQ_DONE;
  }
  RDIS_HD void ask_value(double at, int next_phase) {
    if (!(fabs(at) <= 1.7976931348623157e308)) {  // NaN or inf abscissa: the reference asserts (!isnan), CGD.cpp:172
      finish(ST_NONFINITE);
      return;
    }
    alpha = at;
    req = REQ_VALUE;
    phase = next_phase;
    ++n_value;
  }
  RDIS_HD void ask_value_slope(double at, int next_phase) {
    if (!(fabs(at) <= 1.7976931348623157e308)) {
      finish(ST_NONFINITE);
      return;
    }
    alpha = at;
    req = REQ_VALUE_SLOPE;
    phase = next_phase;
    ++n_slope;
  }

  // linmin(): bracket(0, 1) then dbrent (:496-507).  fa = f(p + 0*xi): fp unless the group emulates the
  // reference's value cache (see the header comment), in which case it is asked for.
  RDIS_HD void begin_line_search() {
    iter = its;  // :645
    ax = 0.0;
    bx = 1.0;
    fa = fp;
    br_iter = 0;
    if (faithful)
      ask_value(ax, PH_BR_FA);
    else
      ask_value(bx, PH_BR_FB);
  }

  RDIS_HD void line_search_done(double xmin, double fmin) {
    fret = fmin;  // :646
    alpha = xmin;
    req = REQ_MOVE;
    phase = PH_MOVED;
  }

  // Top of the dbrent for-loop (:320-378): decide the next trial point.
  RDIS_HD void dbrent_loop() {
    if (db_iter >= 100) {
      finish(ST_DBRENT_ITMAX);  // throw, :403 — p and fret keep their pre-linmin values
      return;
    }
    const double tol = 3.0e-8;
    const double zeps = 2.220446049250313e-16 * 1.0e-3;
    const double xm = 0.5 * (a + b);
    const double tol1 = tol * fabs(x) + zeps;
    const double tol2 = 2.0 * tol1;
    if (fabs(x - xm) <= (tol2 - 0.5 * (b - a))) {
      line_search_done(x, fx);
      return;
    }
    if (fabs(e) > tol1) {
      double d1 = 2.0 * (b - a);
      double d2 = d1;
      if (dw != dx) d1 = (w - x) * dx / (dx - dw);
      if (dv != dx) d2 = (v - x) * dx / (dx - dv);
      const double u1 = x + d1;
      const double u2 = x + d2;
      const bool ok1 = (a - u1) * (u1 - b) > 0.0 && dx * d1 <= 0.0;
      const bool ok2 = (a - u2) * (u2 - b) > 0.0 && dx * d2 <= 0.0;
      const double olde = e;
      e = d;
      if (ok1 || ok2) {
        if (ok1 && ok2)
          d = (fabs(d1) < fabs(d2) ? d1 : d2);
        else if (ok1)
          d = d1;
        else
          d = d2;
        if (fabs(d) <= fabs(0.5 * olde)) {
          const double ut = x + d;
          if (ut - a < tol2 || b - ut < tol2) d = copysign(tol1, xm - x);
        } else {
          e = (dx >= 0.0 ? a - x : b - x);
          d = 0.5 * e;
        }
      } else {
        e = (dx >= 0.0 ? a - x : b - x);
        d = 0.5 * e;
      }
    } else {
      e = (dx >= 0.0 ? a - x : b - x);
      d = 0.5 * e;
    }
    if (fabs(d) >= tol1) {
      u = x + d;
      small_step = false;
    } else {
      u = x + copysign(tol1, d);
      small_step = true;
    }
    ask_value_slope(u, PH_DB_EVAL);
  }

  // One objective evaluation came back.  The phase switch only records what the evaluation was for;
  // the loop bodies of bracket / dbrent exist ONCE below it (stage dispatch) instead of being inlined
  // at each of their call sites: a third of the code, and thread groups that share a warp but sit
  // in different phases reconverge between the stages.
  enum Stage : int { SG_NONE = 0, SG_SHIFT, SG_BRACKET_LOOP, SG_BEGIN_DBRENT, SG_DBRENT_LOOP };

  RDIS_HD void advance(double f, double slope) {
    int stage = SG_NONE;
    switch (phase) {
      case PH_BR_FA:  // :88
        fa = f;
        ask_value(bx, PH_BR_FB);
        break;
      case PH_BR_FB: {  // :89-98
        fb = f;
        if (fb > fa) {
          double t = ax; ax = bx; bx = t;
          t = fb; fb = fa; fa = t;
        }
        cx = bx + 1.618034 * (bx - ax);
        ask_value(cx, PH_BR_FC);
        break;
      }
      case PH_BR_FC:
        fc = f;
        stage = SG_BRACKET_LOOP;
        break;
      case PH_BR_PARAB_INSIDE:  // :113-128
        fu = f;
        if (fu < fc) {
          ax = bx; bx = u; fa = fb; fb = fu;
          stage = SG_BEGIN_DBRENT;
        } else if (fu > fb) {
          cx = u; fc = fu;
          stage = SG_BEGIN_DBRENT;
        } else {
          u = cx + 1.618034 * (cx - bx);
          ask_value(u, PH_BR_SHIFT);
        }
        break;
      case PH_BR_PARAB_BEYOND:  // :130-135
        fu = f;
        if (fu < fc) {
          const double stepped = u + 1.618034 * (u - cx);
          bx = cx; cx = u; u = stepped;
          fb = fc; fc = fu;
          ask_value(u, PH_BR_PARAB_BEYOND2);
        } else {
          stage = SG_SHIFT;
        }
        break;
      case PH_BR_PARAB_BEYOND2:
      case PH_BR_SHIFT:
        fu = f;
        stage = SG_SHIFT;
        break;
      case PH_DB_FIRST:  // :315-316
        fw = fv = fx = f;
        dw = dv = dx = slope;
        stage = SG_DBRENT_LOOP;
        break;
      case PH_DB_EVAL: {  // :365-401
        fu = f;
        if (small_step && fu > fx) {
          line_search_done(x, fx);
          break;
        }
        const double du = slope;
        if (fu <= fx) {
          if (u >= x) a = x; else b = x;
          v = w; fv = fw; dv = dw;
          w = x; fw = fx; dw = dx;
          x = u; fx = fu; dx = du;
        } else {
          if (u < x) a = u; else b = u;
          if (fu <= fw || w == x) {
            v = w; fv = fw; dv = dw;
            w = u; fw = fu; dw = du;
          } else if (fu < fv || v == x || v == w) {
            v = u; fv = fu; dv = du;
          }
        }
        ++db_iter;
        stage = SG_DBRENT_LOOP;
        break;
      }
      default:
        finish(ST_NONFINITE);
        break;
    }
    if (stage == SG_SHIFT) {  // :146-147
      ax = bx; bx = cx; cx = u;
      fa = fb; fb = fc; fc = fu;
      stage = SG_BRACKET_LOOP;
    }
    if (stage == SG_BRACKET_LOOP) {  // top of the bracket while-loop, :101-148
      if (!(fb > fc)) {
        stage = SG_BEGIN_DBRENT;
      } else if (++br_iter > kBracketCap) {
        finish(ST_BRACKET_CAP);
      } else {
        const double r = (bx - ax) * (fb - fc);
        const double q = (bx - cx) * (fb - fa);
        const double qmr = q - r;
        u = bx - ((bx - cx) * q - (bx - ax) * r) / (2.0 * copysign(pick_max(fabs(qmr), 1.0e-20), qmr));
        const double ulim = bx + 100.0 * (cx - bx);
        int next_phase;
        if ((bx - u) * (u - cx) > 0.0) {
          next_phase = PH_BR_PARAB_INSIDE;
        } else if ((cx - u) * (u - ulim) > 0.0) {
          next_phase = PH_BR_PARAB_BEYOND;
        } else if ((u - ulim) * (ulim - cx) >= 0.0) {
          u = ulim;
          next_phase = PH_BR_SHIFT;
        } else {
          u = cx + 1.618034 * (cx - bx);
          next_phase = PH_BR_SHIFT;
        }
        ask_value(u, next_phase);
      }
    }
    if (stage == SG_BEGIN_DBRENT) {  // Dbrent::minimize prologue (:312-316); f(bx) is known (fb), only the slope is new
      a = (ax < cx ? ax : cx);
      b = (ax > cx ? ax : cx);
      x = w = v = bx;
      d = 0.0;
      e = 0.0;
      db_iter = 0;
      ask_value_slope(x, PH_DB_FIRST);
    }
    if (stage == SG_DBRENT_LOOP) dbrent_loop();
  }
};

}  // namespace rdisgpu
