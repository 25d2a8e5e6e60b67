// factors.cuh — device-side factor arithmetic and the HBM layout it reads.
//
// HBM layout (all resident for the life of a context; int32 indices, fp64 values):
//   xval  double[V]    committed value of every variable (mirror of xbd.x outside running solves)
//   xbd   double2[V]   .x = value of the variable (p of the running solve for its variables)
//                      .y = search direction xi for variables of a running solve, NaN for
//                           everything else ("frozen": read as-is, never clamped)
//   dom   double2[V]   {lb, ub} of the single-interval domain
//   NLPF  rowptr i32[F+1] | evid i32[E] | expo f64[E] | konst f64[E] | sine u8[E] | coeff f64[F]
//   BA    cam i32[F] | pt i32[F] | obs double2[F]        (variable ids implied: 9c+p, 9*ncams+3i+d)
//   incidence (variable-major CSR, built at finalize):
//     NLPF  vrow i32[V+1] | vedge i32[E] (edge ids, ascending factor id) | efac i32[E] (edge -> factor)
//     BA    crow i32[ncams+1] | cfac i32[F]  and  prow i32[npts+1] | pfac i32[F]  (factor ids ascending)
//   scratch: gedge f64[E] per-edge partials, gvec/hvec/xsave f64[V] CG vectors, fstamp i32[F]
//
// Arithmetic follows the reference expression by expression (operand order kept):
//   NonlinearProductFactor::evalFactor / getDerivative   src/NonlinearProductFactor.cpp:186-209,149-178
//   power()                                              src/util/numeric.cpp:12-23
//   BundleAdjustmentFactor forward model                 src/bundleadjust/BundleAdjustmentFactor.cpp:266-335,
//                                                        BundleAdjustmentFactor.h:80-107, BundleAdjustmentCommon.h:64-93
//   BundleAdjustmentFactor::computeGradient(vals, grad)  src/bundleadjust/BundleAdjustmentFactor.cpp:351-554
//   VariableDomain::closestVal                           src/VariableDomain.cpp:157-163
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "trig.cuh"

namespace rdisgpu {

enum FactorKind : int { KIND_NONE = -1, KIND_NLPF = 0, KIND_BA = 1 };

struct GraphView {
  int kind;
  int64_t V, F, E;
  double2* xbd;
  double* xval;  // dense mirror of xbd[].x for frozen variables (8 B/variable: what the streaming sweep gathers)
  const double2* dom;
  // NLPF
  const int32_t* rowptr;
  const int32_t* evid;
  const double* expo;
  const double* konst;
  const uint8_t* sine;
  const double* coeff;
  const int32_t* vrow;
  const int32_t* vedge;
  const int32_t* efac;
  // NLPF term table (nlpf_resident.cuh): the distinct (k, e, sine) expressions of each variable, variable-major
  const int32_t* tvrow;   // i32[V+1]
  const int32_t* eterm;   // i32[E]: edge -> its term (global term id)
  const double* t_expo;   // f64[U]
  const double* t_konst;  // f64[U]
  const uint8_t* t_sine;  // u8[U]
  int32_t* vloc;          // i32[V] scratch: variable -> slot inside the component that owns it
  int32_t* floc;          // i32[F] scratch: factor -> slot inside the component that owns it
  // BA
  const int32_t* cam;
  const int32_t* pt;
  const double2* obs;
  int32_t ncams, npts;
  const int32_t* crow;
  const int32_t* cfac;
  const int32_t* prow;
  const int32_t* pfac;
  // constant overlay (nullable)
  const uint8_t* fconst_on;
  const double* fconst_val;
  // the reference's per-factor value cache (src/Factor.h:228-234) and Variable::assign's change flags, strict mode
  // only (nullable): fcache f64[F] last value Factor::eval computed, fdirty u8[F], vchg u8[V]
  double* fcache;
  uint8_t* fdirty;
  uint8_t* vchg;
  // scratch
  double* gedge;
  double* gvec;
  double* hvec;
  double* xsave;
  int32_t* fstamp;
};

__device__ __forceinline__ double clamp_to_domain(double v, double2 d) {
  if (d.x <= v && v <= d.y) return v;
  if (v < d.x) return d.x;
  return d.y;
}

// Value of variable `vid` as a factor sees it.  kMode 0 / 1: frozen variables (direction slot NaN) are read as
// stored; variables of the running solve are p (mode 0) or p + alpha*xi (mode 1) clamped into their domain
// (SubfunctionFD::quickAssignVals, src/optimizers/CGDSubspaceOptimizer.cpp:160-184).  p + alpha*xi is a product
// rounded, then a sum rounded — Df1dim's xt[j] = p[j] + x*xi[j] (minimize_nrc.h:434) on a target without FMA.
// kMode 2: the ASSIGNED value (the dense mirror xval = Variable::eval() of the reference): the strict kernels
// assign first and evaluate afterwards, like the reference.
// `dirv` receives xi (0 for frozen variables) so callers can form directional derivatives.
constexpr int kAtP = 0, kOnLine = 1, kAssigned = 2;
template <int kMode>
__device__ __forceinline__ double load_var(const GraphView& G, int32_t vid, double alpha, double& dirv) {
  if (kMode == kAssigned) {
    dirv = 0.0;
    return G.xval[vid];
  }
  const double2 xb = G.xbd[vid];
  if (xb.y != xb.y) {  // frozen
    dirv = 0.0;
    return xb.x;
  }
  dirv = xb.y;
  const double raw = (kMode == kOnLine) ? __dadd_rn(xb.x, __dmul_rn(alpha, xb.y)) : xb.x;  // never contracted
  return clamp_to_domain(raw, __ldg(&G.dom[vid]));
}

// sin / cos: trig.cuh (bit-identical to the CUDA math library, and host-compilable for the oracle's devtrig twin)
__device__ __forceinline__ double rdis_power(double val, double e) {
  if (e == 0.) return 1.;
  if (e == 1.) return val;
  if (e == 2.) return val * val;
  return pow(val, e);
}

// ------------------------------------------------------------------------------------------
// NonlinearProductFactor
// ------------------------------------------------------------------------------------------
struct NlpfOps {
  static constexpr int kMaxArityFast = 4;  // arities above this take the O(arity^2) recompute path (the generators stop at 4; 8 cost 242 registers)

  // f_j at abscissa alpha; if kSlope also d f_j / d alpha = sum_i (d f_j/d x_i) * xi_i.
  template <int kAlongLine>
  __device__ static __forceinline__ double value(const GraphView& G, int64_t fid, double alpha, bool kSlope,
                                                 double& slope) {
    const int32_t e0 = __ldg(&G.rowptr[fid]);
    const int32_t e1 = __ldg(&G.rowptr[fid + 1]);
    const double c = __ldg(&G.coeff[fid]);
    const int ar = e1 - e0;
    double prod = 1.0;
    if (!kSlope) {
      for (int32_t e = e0; e < e1; ++e) {
        double dirv;
        double val = load_var<kAlongLine>(G, __ldg(&G.evid[e]), alpha, dirv);
        const double k = __ldg(&G.konst[e]);
        const double ex = __ldg(&G.expo[e]);
        if (k != 0) val -= k;
        if (ex != 1) val = rdis_power(val, ex);
        if (__ldg(&G.sine[e])) val = rdis_sin(val);
        prod *= val;
      }
      slope = 0.0;
      return prod * c;
    }
    if (ar <= kMaxArityFast) {
      double t[kMaxArityFast], dt[kMaxArityFast], dir[kMaxArityFast];
      bool plain[kMaxArityFast];
#pragma unroll
      for (int i = 0; i < kMaxArityFast; ++i) {
        if (i < ar) {
          const int32_t e = e0 + i;
          term<kAlongLine>(G, e, alpha, t[i], dt[i], plain[i], dir[i]);
          prod *= t[i];
        }
      }
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < kMaxArityFast; ++i) {
        if (i < ar && dir[i] != 0.0) {
          double pe = 1.0;  // getDerivative(vid_i): product in slot order, own slot replaced by its derivative
#pragma unroll
          for (int j = 0; j < kMaxArityFast; ++j) {
            if (j < ar) {
              if (j == i) {
                if (!plain[j]) pe *= dt[j];
              } else {
                pe *= t[j];
              }
            }
          }
          s = s + (pe * c) * dir[i];  // product rounded, then the sum (no contraction anywhere: the resident kernel folds the same way)
        }
      }
      slope = s;
      return prod * c;
    }
    // generic arity
    double s = 0.0;
    for (int32_t ei = e0; ei < e1; ++ei) {
      double ti, dti, diri;
      bool pl;
      term<kAlongLine>(G, ei, alpha, ti, dti, pl, diri);
      prod *= ti;
      if (diri != 0.0) s = s + (partial<kAlongLine>(G, e0, e1, ei, alpha) * c) * diri;
    }
    slope = s;
    return prod * c;
  }

  // All partials of factor fid at the current point (alpha ignored, direction ignored):
  // writes gedge[e] for e in the factor's row; returns the factor value.
  template <int kMode = kAtP>
  __device__ static __forceinline__ double gradient(const GraphView& G, int64_t fid, double* gout /*row base*/) {
    const int32_t e0 = __ldg(&G.rowptr[fid]);
    const int32_t e1 = __ldg(&G.rowptr[fid + 1]);
    const double c = __ldg(&G.coeff[fid]);
    const int ar = e1 - e0;
    double prod = 1.0;
    if (ar <= kMaxArityFast) {
      double t[kMaxArityFast], dt[kMaxArityFast], dir[kMaxArityFast];
      bool plain[kMaxArityFast];
#pragma unroll
      for (int i = 0; i < kMaxArityFast; ++i) {
        if (i < ar) {
          term<kMode>(G, e0 + i, 0.0, t[i], dt[i], plain[i], dir[i]);
          prod *= t[i];
        }
      }
#pragma unroll
      for (int i = 0; i < kMaxArityFast; ++i) {
        if (i < ar) {
          double pe = 1.0;
#pragma unroll
          for (int j = 0; j < kMaxArityFast; ++j) {
            if (j < ar) {
              if (j == i) {
                if (!plain[j]) pe *= dt[j];
              } else {
                pe *= t[j];
              }
            }
          }
          gout[i] = pe * c;
        }
      }
      return prod * c;
    }
    for (int32_t ei = e0; ei < e1; ++ei) {
      double ti, dti, diri;
      bool pl;
      term<kMode>(G, ei, 0.0, ti, dti, pl, diri);
      prod *= ti;
      gout[ei - e0] = partial<kMode>(G, e0, e1, ei, 0.0) * c;
    }
    return prod * c;
  }
  __device__ static __forceinline__ int64_t edge_base(const GraphView& G, int64_t fid) { return __ldg(&G.rowptr[fid]); }
  __device__ static __forceinline__ int arity(const GraphView& G, int64_t fid) {
    return __ldg(&G.rowptr[fid + 1]) - __ldg(&G.rowptr[fid]);
  }

  // d/dx_v of the sum of the stamped factors, incident edges visited in ascending factor id
  // (the order productGradient accumulates them in, src/State.h:157-194).
  __device__ static __forceinline__ double gather_var(const GraphView& G, int32_t vid, int32_t stamp, bool filter) {
    double acc = 0.0;
    bool first = true;
    const int32_t r0 = __ldg(&G.vrow[vid]), r1 = __ldg(&G.vrow[vid + 1]);
    if (!filter) {
      // all factors contribute: fetch the partials eight at a time (independent loads in flight),
      // fold them in the same ascending order
      for (int32_t r = r0; r < r1; r += 8) {
        int32_t e[8];
        double ge[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = (r + i < r1) ? __ldg(&G.vedge[r + i]) : -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) ge[i] = (e[i] >= 0) ? G.gedge[e[i]] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (e[i] >= 0) {
            acc = first ? ge[i] : acc + ge[i];
            first = false;
          }
        }
      }
      return acc;
    }
    for (int32_t r = r0; r < r1; ++r) {
      const int32_t e = __ldg(&G.vedge[r]);
      if (G.fstamp[__ldg(&G.efac[e])] != stamp) continue;
      const double ge = G.gedge[e];
      acc = first ? ge : acc + ge;
      first = false;
    }
    return acc;
  }

  // ---- strict kernels (strict_kernels.cuh): evaluate from the ASSIGNED state, as the reference does ----
  __device__ static __forceinline__ double strict_value(const GraphView& G, int64_t fid) {
    double sl;
    return value<kAssigned>(G, fid, 0.0, false, sl);
  }
  __device__ static __forceinline__ double strict_gradient(const GraphView& G, int64_t fid, double* gout) {
    return gradient<kAssigned>(G, fid, gout);
  }
  // did Variable::assign notify this factor (a variable of the running solve moved by >= 1e-12)?
  __device__ static __forceinline__ bool any_own_changed(const GraphView& G, int64_t fid) {
    const int32_t e0 = __ldg(&G.rowptr[fid]), e1 = __ldg(&G.rowptr[fid + 1]);
    bool chg = false;
    for (int32_t e = e0; e < e1; ++e) {
      const int32_t v = __ldg(&G.evid[e]);
      const double d = G.xbd[v].y;
      if (d == d && G.vchg[v]) chg = true;
    }
    return chg;
  }

 private:
  // value term t = [sin]((x-k)^e) and the own-slot derivative factor e*(x-k)^(e-1)*[cos((x-k)^e)];
  // `plain` marks the e==1, no-sine case the reference skips (derivative 1, constant not subtracted).
  template <int kAlongLine>
  __device__ static __forceinline__ void term(const GraphView& G, int32_t e, double alpha, double& t, double& dt,
                                              bool& plain, double& dirv) {
    const double xv = load_var<kAlongLine>(G, __ldg(&G.evid[e]), alpha, dirv);
    const double k = __ldg(&G.konst[e]);
    const double ex = __ldg(&G.expo[e]);
    const bool sn = __ldg(&G.sine[e]) != 0;
    double val = xv;
    if (k != 0) val -= k;
    plain = (ex == 1) && !sn;
    if (plain) {
      t = val;
      dt = 1.0;
      return;
    }
    if (ex == 1) {  // sine of x - k: the general path below reduces to dt = 1.0 * 1.0 * cos(val) bit for bit (nlpf_term_grad)
      double sv, cv;
      rdis_sincos(val, sv, cv);
      t = sv;
      dt = cv;
      return;
    }
    val = rdis_power(val, ex);
    const double inner = xv - k;
    const double innerexp = rdis_power(inner, ex);
    double dv = rdis_power(inner, ex - 1.0);
    dv *= ex;
    if (sn) {
      // val and innerexp are the same number (k == 0: x and x - 0; else x - k both times), so one
      // argument reduction serves both; each result is bit-identical to the separate call
      double sv, cv;
      rdis_sincos(val, sv, cv);
      dv *= (innerexp == val) ? cv : rdis_cos(innerexp);
      t = sv;  // same value as the value-only path, so a point has one value
    } else {
      t = val;
    }
    dt = dv;
  }
  // getDerivative for the variable of edge `etarget` (recompute path, any arity)
  template <int kAlongLine>
  __device__ static __forceinline__ double partial(const GraphView& G, int32_t e0, int32_t e1, int32_t etarget,
                                                   double alpha) {
    double pe = 1.0;
    for (int32_t e = e0; e < e1; ++e) {
      double t, dt, dirv;
      bool pl;
      term<kAlongLine>(G, e, alpha, t, dt, pl, dirv);
      if (e == etarget) {
        if (!pl) pe *= dt;
      } else {
        pe *= t;
      }
    }
    return pe;
  }
};

// Quotients by a denominator that is used many times.  The reference divides (IEEE, correctly rounded).
//   kExact     a / b, the division instruction sequence every time (strict kernels)
//   kCorrected the reciprocal y = RN(1/b) once, then q = RN(a*y), r = a - b*q (exact, one fma), RN(q + r*y):
//              the correction step the division sequence itself ends with.  The result is the correctly rounded
//              quotient unless a/b lies within 3*2^-54 ulp of a rounding boundary after a first guess more than one
//              ulp off (probability ~2^-52 per quotient; tests/test_gpu_parity.py counts zero mismatches against
//              `/` on 2^28 operand pairs), or at singular operands (b = 0, inf; signed zeros).
enum DivMode : int { kExact = 0, kCorrected = 1 };
template <int kDiv>
struct QuotBy {
  double b, y;
  __device__ __forceinline__ explicit QuotBy(double b_) : b(b_), y(kDiv == kExact ? 0.0 : 1.0 / b_) {}
  __device__ __forceinline__ double operator()(double a) const {
    if (kDiv == kExact) return a / b;
    const double q = __dmul_rn(a, y);
    const double r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
  }
};

// ------------------------------------------------------------------------------------------
// BundleAdjustmentFactor.  Slot order: rot xyz, trans xyz, focal, k1, k2, point xyz
// (src/bundleadjust/BundleAdjustmentCommon.h:36-59).
// ------------------------------------------------------------------------------------------
struct BaOps {
  struct Fwd {
    double a0, a1, a2, theta;   // normalised axis, angle
    double c0, c1, c2, adp;     // axis x point, axis . point
    double P0, P1, P2;          // camera-frame point
    double pp0, pp1, r2, dist;
    double res0, res1;
    double s, c;                // sin/cos(theta)
  };

  // Camera-only part of the forward model: theta = |r|, axis = r/theta, sin/cos(theta)
  // (normalize + the theta > 0 branch of angleAxisRotateTranslatePoint, BundleAdjustmentFactor.cpp:266-335).
  // It depends on the three rotation variables only, so the block kernels hoist it out of the
  // per-factor work; the arithmetic is the same expression for expression.
  __device__ static __forceinline__ void rotation(double r0, double r1, double r2v, Fwd& m) {
    const double nrm = sqrt(r0 * r0 + r1 * r1 + r2v * r2v);
    if (nrm != 0.0) {
      m.a0 = r0 / nrm; m.a1 = r1 / nrm; m.a2 = r2v / nrm;
    } else {
      m.a0 = r0; m.a1 = r1; m.a2 = r2v;
    }
    m.theta = nrm;
    if (nrm > 0.0) {
      rdis_sincos(nrm, m.s, m.c);  // bit-identical to sincos() (tests/native/trig_check.cu), inlined and host-reproducible
    } else {
      m.s = 0.0; m.c = 1.0;  // sin(0), cos(0): what the gradient code recomputes from theta
    }
  }

  // Rotate + translate + project + distort + residual, given rotation(): x[3..11] are read.
  // The two quotients by the depth share one reciprocal (QuotBy above; kExact in strict mode).
  template <int kDiv = kCorrected>
  __device__ static __forceinline__ double project(const double* x, double2 ob, Fwd& m) {
    const double q0 = x[9], q1 = x[10], q2 = x[11];
    m.c0 = m.a1 * q2 - m.a2 * q1;
    m.c1 = m.a2 * q0 - m.a0 * q2;
    m.c2 = m.a0 * q1 - m.a1 * q0;
    double P0, P1, P2;
    if (m.theta > 0.0) {
      const double omc = 1 - m.c;
      m.adp = m.a0 * q0 + m.a1 * q1 + m.a2 * q2;
      P0 = q0 * m.c + m.c0 * m.s + m.a0 * omc * m.adp;
      P1 = q1 * m.c + m.c1 * m.s + m.a1 * omc * m.adp;
      P2 = q2 * m.c + m.c2 * m.s + m.a2 * omc * m.adp;
    } else {
      m.adp = 0;
      P0 = q0 + m.c0; P1 = q1 + m.c1; P2 = q2 + m.c2;
    }
    P0 += x[3]; P1 += x[4]; P2 += x[5];
    m.P0 = P0; m.P1 = P1; m.P2 = P2;
    const QuotBy<kDiv> by_depth(P2);
    m.pp0 = by_depth(-P0);
    m.pp1 = by_depth(-P1);
    m.r2 = m.pp0 * m.pp0 + m.pp1 * m.pp1;
    m.dist = 1 + m.r2 * (x[7] + x[8] * m.r2);
    const double pix0 = x[6] * m.dist * m.pp0;
    const double pix1 = x[6] * m.dist * m.pp1;
    m.res0 = (pix0 - ob.x);
    m.res1 = (pix1 - ob.y);
    return (m.res0 * m.res0 + m.res1 * m.res1) / 2.0;
  }

  template <int kDiv = kCorrected>
  __device__ static __forceinline__ double forward(const double* x, double2 ob, Fwd& m) {
    rotation(x[0], x[1], x[2], m);
    return project<kDiv>(x, ob, m);
  }

  // 12 partials in slot order.  The reference divides by P_z^2, P_z and |r| 24 times per observation
  // (through_projection is expanded six times, BundleAdjustmentFactor.cpp:418-422 and repeats); QuotBy keeps those
  // quotients (correctly rounded) while paying for three reciprocals.
  template <int kDiv = kCorrected>
  __device__ static __forceinline__ void partials(const double* x, const Fwd& m, double* g) {
    const double f = x[6], k1 = x[7], k2 = x[8];
    const double q[3] = {x[9], x[10], x[11]};
    const double a[3] = {m.a0, m.a1, m.a2};
    const double axp[3] = {m.c0, m.c1, m.c2};
    const double P[3] = {m.P0, m.P1, m.P2};
    const double s = m.s, c = m.c, adp = m.adp;
    const double t1 = 2.0 * (k1 + 2.0 * k2 * m.r2);
    const double vnorm = m.theta;
    const double P22 = P[2] * P[2];
    const double pp00 = m.pp0 * m.pp0, pp01 = m.pp0 * m.pp1, pp11 = m.pp1 * m.pp1;
    const double J00 = m.dist + t1 * pp00, J01 = t1 * pp01, J10 = t1 * pp01, J11 = m.dist + t1 * pp11;
    const double omc = (1 - c);

    const QuotBy<kDiv> byP22(P22), byP2(P[2]), byVn(vnorm);
    auto chain = [&](double dP0, double dP1, double dP2) -> double {
      const double dppx = byP22(P[0] * dP2 - P[2] * dP0);
      const double dppy = byP22(P[1] * dP2 - P[2] * dP1);
      const double drx = m.res0 * (J00 * dppx + J01 * dppy);
      const double dry = m.res1 * (J10 * dppx + J11 * dppy);
      return f * (drx + dry);
    };

    // d P / d axis (3x3) and d P / d theta
    const double A00 = (adp + a[0] * q[0]) * omc;
    const double A01 = q[2] * s + a[0] * q[1] * omc;
    const double A02 = -q[1] * s + a[0] * q[2] * omc;
    const double T0 = -q[0] * s + axp[0] * c + a[0] * adp * s;
    const double A10 = -q[2] * s + a[1] * q[0] * omc;
    const double A11 = (adp + a[1] * q[1]) * omc;
    const double A12 = q[0] * s + a[1] * q[2] * omc;
    const double T1 = -q[1] * s + axp[1] * c + a[1] * adp * s;
    const double A20 = q[1] * s + a[2] * q[0] * omc;
    const double A21 = -q[0] * s + a[2] * q[1] * omc;
    const double A22 = (adp + a[2] * q[2]) * omc;
    const double T2 = -q[2] * s + axp[2] * c + a[2] * adp * s;

    // rotation vector, component x
    {
      const double d0 = byVn(a[1] * a[1] + a[2] * a[2]);
      const double d1 = byVn(-a[0] * a[1]);
      const double d2 = byVn(-a[0] * a[2]);
      g[0] = chain(A00 * d0 + A01 * d1 + A02 * d2 + T0 * a[0], A10 * d0 + A11 * d1 + A12 * d2 + T1 * a[0],
                     A20 * d0 + A21 * d1 + A22 * d2 + T2 * a[0]);
    }
    {
      const double d0 = byVn(-a[0] * a[1]);
      const double d1 = byVn(a[0] * a[0] + a[2] * a[2]);
      const double d2 = byVn(-a[1] * a[2]);
      g[1] = chain(A00 * d0 + A01 * d1 + A02 * d2 + T0 * a[1], A10 * d0 + A11 * d1 + A12 * d2 + T1 * a[1],
                     A20 * d0 + A21 * d1 + A22 * d2 + T2 * a[1]);
    }
    {
      const double d0 = byVn(-a[0] * a[2]);
      const double d1 = byVn(-a[1] * a[2]);
      const double d2 = byVn(a[0] * a[0] + a[1] * a[1]);
      g[2] = chain(A00 * d0 + A01 * d1 + A02 * d2 + T0 * a[2], A10 * d0 + A11 * d1 + A12 * d2 + T1 * a[2],
                     A20 * d0 + A21 * d1 + A22 * d2 + T2 * a[2]);
    }
    // translation
    g[3] = byP2((m.res0 * J00 + m.res1 * J10) * -f);
    g[4] = byP2((m.res0 * J01 + m.res1 * J11) * -f);
    {
      const double dpx = J00 * P[0] + J01 * P[1];
      const double dpy = J10 * P[0] + J11 * P[1];
      g[5] = byP22((m.res0 * dpx + m.res1 * dpy) * f);
    }
    // intrinsics
    g[6] = m.res0 * (m.dist * m.pp0) + m.res1 * (m.dist * m.pp1);
    g[7] = m.res0 * (f * m.r2 * m.pp0) + m.res1 * (f * m.r2 * m.pp1);
    g[8] = m.res0 * (f * m.r2 * m.r2 * m.pp0) + m.res1 * (f * m.r2 * m.r2 * m.pp1);
    // point: columns of the rotation matrix
    g[9] = chain(c * (1.0 - a[0] * a[0]) + a[0] * a[0], a[2] * s + a[0] * a[1] * (1.0 - c),
                   -a[1] * s + a[0] * a[2] * (1.0 - c));
    g[10] = chain(-a[2] * s + a[0] * a[1] * (1.0 - c), c * (1.0 - a[1] * a[1]) + a[1] * a[1],
                    a[0] * s + a[1] * a[2] * (1.0 - c));
    g[11] = chain(a[1] * s + a[0] * a[2] * (1.0 - c), -a[0] * s + a[1] * a[2] * (1.0 - c),
                    c * (1.0 - a[2] * a[2]) + a[2] * a[2]);
  }

  __device__ static __forceinline__ int32_t slot_vid(const GraphView& G, int32_t cam, int32_t pt, int s) {
    return (s < 9) ? (9 * cam + s) : (9 * G.ncams + 3 * pt + (s - 9));
  }

  template <int kAlongLine, int kDiv = kCorrected>
  __device__ static __forceinline__ double value(const GraphView& G, int64_t fid, double alpha, bool kSlope,
                                                 double& slope) {
    const int32_t cam = __ldg(&G.cam[fid]);
    const int32_t pt = __ldg(&G.pt[fid]);
    const double2 ob = __ldg(&G.obs[fid]);
    double x[12], dir[12];
#pragma unroll
    for (int s = 0; s < 12; ++s) x[s] = load_var<kAlongLine>(G, slot_vid(G, cam, pt, s), alpha, dir[s]);
    Fwd m;
    const double fv = forward<kDiv>(x, ob, m);
    if (kSlope) {
      double g[12];
      partials<kDiv>(x, m, g);
      double sl = 0.0;
#pragma unroll
      for (int s = 0; s < 12; ++s)
        if (dir[s] != 0.0) sl += g[s] * dir[s];
      slope = sl;
    } else {
      slope = 0.0;
    }
    return fv;
  }

  template <int kMode = kAtP, int kDiv = kCorrected>
  __device__ static __forceinline__ double gradient(const GraphView& G, int64_t fid, double* gout) {
    const int32_t cam = __ldg(&G.cam[fid]);
    const int32_t pt = __ldg(&G.pt[fid]);
    const double2 ob = __ldg(&G.obs[fid]);
    double x[12], dirv;
#pragma unroll
    for (int s = 0; s < 12; ++s) x[s] = load_var<kMode>(G, slot_vid(G, cam, pt, s), 0.0, dirv);
    Fwd m;
    const double fv = forward<kDiv>(x, ob, m);
    double g[12];
    partials<kDiv>(x, m, g);
#pragma unroll
    for (int s = 0; s < 12; ++s) gout[s] = g[s];
    return fv;
  }
  __device__ static __forceinline__ double strict_value(const GraphView& G, int64_t fid) {
    double sl;
    return value<kAssigned, kExact>(G, fid, 0.0, false, sl);
  }
  __device__ static __forceinline__ double strict_gradient(const GraphView& G, int64_t fid, double* gout) {
    return gradient<kAssigned, kExact>(G, fid, gout);
  }
  __device__ static __forceinline__ bool any_own_changed(const GraphView& G, int64_t fid) {
    const int32_t cam = __ldg(&G.cam[fid]);
    const int32_t pt = __ldg(&G.pt[fid]);
    bool chg = false;
#pragma unroll
    for (int s = 0; s < 12; ++s) {
      const int32_t v = slot_vid(G, cam, pt, s);
      const double d = G.xbd[v].y;
      if (d == d && G.vchg[v]) chg = true;
    }
    return chg;
  }
  __device__ static __forceinline__ int64_t edge_base(const GraphView&, int64_t fid) { return fid * 12; }
  __device__ static __forceinline__ int arity(const GraphView&, int64_t) { return 12; }

  __device__ static __forceinline__ double gather_var(const GraphView& G, int32_t vid, int32_t stamp, bool filter) {
    const int32_t ncv = 9 * G.ncams;
    const int32_t* row;
    const int32_t* lst;
    int32_t blk, slot;
    if (vid < ncv) {
      blk = vid / 9; slot = vid - 9 * blk; row = G.crow; lst = G.cfac;
    } else {
      const int32_t o = vid - ncv;
      blk = o / 3; slot = 9 + (o - 3 * blk); row = G.prow; lst = G.pfac;
    }
    double acc = 0.0;
    bool first = true;
    const int32_t r0 = __ldg(&row[blk]), r1 = __ldg(&row[blk + 1]);
    for (int32_t r = r0; r < r1; ++r) {
      const int32_t fid = __ldg(&lst[r]);
      if (filter && G.fstamp[fid] != stamp) continue;
      const double ge = G.gedge[(int64_t)fid * 12 + slot];
      acc = first ? ge : acc + ge;
      first = false;
    }
    return acc;
  }
};

}  // namespace rdisgpu
