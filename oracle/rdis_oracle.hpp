// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under rdis_b200/ may include, link or
// execute this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs do, and only as the checker / CPU baseline.
//
// Single-threaded CPU restatement (std-only C++17) of the reference's subspace-solve
// hot path, keeping its object graph (heap Variables / Factors, per-factor value cache,
// dirty-flag fan-out on assign, sorted-map gradient merge) so that CPU timings reflect
// the reference's design.  Citations are relative to /root/reference/.
//
//   Variable::assign change filter            src/Variable.cpp:66-88
//   VariableDomain::closestVal                src/VariableDomain.cpp:130-163 (one interval)
//   Factor::eval / cache / onVar*             src/Factor.cpp:110-119,154-181, src/Factor.h:228-234
//   Factor::computeGradient                   src/Factor.cpp:142-151
//   power()                                   src/util/numeric.cpp:12-23
//   NonlinearProductFactor                    src/NonlinearProductFactor.cpp:27-54,57-117,149-209
//   SimpleSumFactor                           src/SimpleSumFactor.cpp:29-54,95-186
//   BundleAdjustmentFactor                    src/bundleadjust/BundleAdjustmentFactor.cpp:55-64,160-185,266-335,338-554
//                                             + BundleAdjustmentFactor.h:80-107, BundleAdjustmentCommon.h:36-93
//   OptimizableFunction::evalFactors          src/OptimizableFunction.cpp:95-135
//   computeGradientOfSum / productGradient    src/OptimizableFunction.cpp:248-262, src/State.h:157-194
//   SubspaceOptimizer / CGDSubspaceOptimizer  src/SubspaceOptimizer.cpp:12-53, src/optimizers/CGDSubspaceOptimizer.cpp:19-184
//
// PARITY STATUS.  The CG / line-search driver is pinned bit-for-bit against the
// reference's own minimize_nrc.h (see nr_minimize.hpp).  The reference as a whole
// cannot be built here (Boost is required by every translation unit and is absent), so
// factor arithmetic is pinned by (a) the documented optimum in data/testpoly.txt:18-22,
// (b) the ten known minima of src/main.cpp:107-135 (SimpleSumFactor functions),
// (c) the reference's own finite-difference gradient criterion (src/Factor.cpp:191-228).
// Bundle-adjustment and sinusoid objective values have no golden numbers anywhere in
// the reference: for those the arithmetic parity is UNPINNED beyond (c).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifdef ORACLE_USE_REFERENCE_NRC
// Build variant that drives the solve with the reference's own, unmodified header.
namespace rdis { typedef double Numeric; }
#include "minimize_nrc.h"  // found via -I/root/reference/external/include
#else
#include "nr_minimize.hpp"
#endif

#ifdef ORACLE_DEVICE_TRIG
// "devtrig" twin: sin / cos are the DEVICE's routines compiled for the host (rdis_b200/csrc/trig.cuh: explicit
// fma, bit-identical to the CUDA math library) instead of glibc's.  Every other operation of the path (+ - * /
// sqrt, compares) is correctly rounded IEEE on both sides, so the strict device mode must reproduce this twin
// BIT FOR BIT; the twin against the glibc build measures what a different libm alone does to a solve.
#include "trig.cuh"  // found via -I rdis_b200/csrc
#endif

namespace oracle {

typedef double Numeric;
#ifdef ORACLE_DEVICE_TRIG
inline Numeric orc_sin(Numeric x) { return rdisgpu::rdis_sin(x); }
inline Numeric orc_cos(Numeric x) { return rdisgpu::rdis_cos(x); }
#else
inline Numeric orc_sin(Numeric x) { return std::sin(x); }
inline Numeric orc_cos(Numeric x) { return std::cos(x); }
#endif
typedef long long VariableID;
typedef long long FactorID;

struct Counters {
  long long factor_eval_calls = 0;    // reference counter g_numFactEvals (src/Factor.cpp:114): counts cache hits too
  long long factor_recomputes = 0;    // evalFactor() bodies actually run
  long long factor_grad_calls = 0;    // Factor::computeGradient calls
  long long f_evals = 0;              // SubfunctionFD::operator()
  long long df_evals = 0;             // SubfunctionFD::df
};

// src/util/numeric.cpp:12-23
inline Numeric power(Numeric val, Numeric e) {
  if (e == 0.) return 1.;
  if (e == 1.) return val;
  if (e == 2.) return val * val;
  return std::pow(val, e);
}

// One closed interval; the reference's CGD path asserts a single subinterval
// (src/optimizers/CGDSubspaceOptimizer.cpp:118-119).
struct Domain {
  Numeric lo = -std::numeric_limits<Numeric>::max();
  Numeric hi = std::numeric_limits<Numeric>::max();
  // src/VariableDomain.cpp:157-163 with boost::numeric::in(val, si) == (lo <= val && val <= hi)
  Numeric closestVal(Numeric val) const {
    if (lo <= val && val <= hi) return val;
    if (val < lo) return lo;
    return hi;
  }
};

class Factor;

class Variable {
 public:
  Variable(VariableID id, Domain d) : id_(id), dom_(d) {}
  VariableID getID() const { return id_; }
  const Domain& getDomain() const { return dom_; }
  void setDomain(Domain d) { dom_ = d; }
  bool isAssigned() const { return assigned_; }
  Numeric eval() const { return value_; }
  void addFactor(Factor* f) { factors_.push_back(f); }
  const std::vector<Factor*>& getFactors() const { return factors_; }
  inline void assign(Numeric newval);
  inline void unassign();
  // When false the 1e-12 notification filter of src/Variable.cpp:69-73 is bypassed
  // (every assign marks the factors dirty).  Lets tests bound the effect of the filter.
  static bool& changeFilter() { static bool on = true; return on; }
  Numeric samp_lo = 0, samp_hi = 0;  // sampling interval (Variable.h), host bookkeeping only

 private:
  VariableID id_;
  Domain dom_;
  bool assigned_ = false;
  Numeric value_ = 0;
  std::vector<Factor*> factors_;
};

// Sorted (vid -> d f/d x_vid) map with boost::container::flat_map behaviour.
class PartialGradient {
 public:
  typedef std::pair<VariableID, Numeric> Entry;
  void clear() { e_.clear(); }
  bool empty() const { return e_.empty(); }
  size_t size() const { return e_.size(); }
  void reserve(size_t n) { e_.reserve(n); }
  std::vector<Entry>& entries() { return e_; }
  const std::vector<Entry>& entries() const { return e_; }
  Numeric& operator[](VariableID vid) {
    auto it = std::lower_bound(e_.begin(), e_.end(), vid,
                               [](const Entry& a, VariableID v) { return a.first < v; });
    if (it == e_.end() || it->first != vid) it = e_.insert(it, Entry(vid, 0.0));
    return it->second;
  }
  const Numeric* find(VariableID vid) const {
    auto it = std::lower_bound(e_.begin(), e_.end(), vid,
                               [](const Entry& a, VariableID v) { return a.first < v; });
    if (it == e_.end() || it->first != vid) return nullptr;
    return &it->second;
  }

 private:
  std::vector<Entry> e_;
};

// src/State.h:157-194 with the MinSum semiring's Product (= +): merge g2 into g1.
inline void productGradient(PartialGradient& g1, const PartialGradient& g2) {
  if (g1.empty()) {
    g1 = g2;
    return;
  }
  auto& a = g1.entries();
  const auto& b = g2.entries();
  a.reserve(a.size() + b.size());
  size_t i = 0, j = 0;
  while (j < b.size()) {
    if (i == a.size() || a[i].first > b[j].first) {
      a.insert(a.begin() + i, b[j]);  // O(n) middle insert, as flat_map does
      ++i;
      ++j;
    } else if (a[i].first == b[j].first) {
      a[i].second = a[i].second + b[j].second;
      ++i;
      ++j;
    } else {
      ++i;
    }
  }
}

class Factor {
 public:
  explicit Factor(FactorID id) : id_(id) {}
  virtual ~Factor() {}
  FactorID getID() const { return id_; }
  const std::vector<Variable*>& getVariables() const { return vars_; }
  bool isAssigned() const { return assignedConstant_; }
  bool areAllVarsAssigned() const { return numAssigned_ == (long long)vars_.size(); }

  // src/Factor.cpp:110-119
  Numeric eval(Counters& c) const {
    assert(assignedConstant_ || areAllVarsAssigned());
    if (!assignedConstant_) {
      if (dirty_) {  // src/Factor.h:228-234
        cached_ = evalFactor();
        dirty_ = false;
        ++c.factor_recomputes;
      }
      ++c.factor_eval_calls;
    }
    return cached_;
  }
  Numeric evalNoCache() const { return evalFactor(); }

  // The NLPF and BA overrides (src/NonlinearProductFactor.cpp:57-117,
  // src/bundleadjust/BundleAdjustmentFactor.cpp:338-348) ignore the assigned-constant
  // flag; only the base version (src/Factor.cpp:142-151, used by SimpleSumFactor)
  // returns early.
  virtual void computeGradient(PartialGradient& g) const {
    if (assignedConstant_) return;
    g.reserve(vars_.size());
    for (const Variable* v : vars_) g[v->getID()] = getDerivative(v->getID());
  }

  // "assigned constant" = factor simplified away by the tree search (host bookkeeping,
  // src/Factor.cpp:244-341); the hot path only needs the flag and the frozen value.
  void setAssignedConstant(bool on, Numeric value) {
    assignedConstant_ = on;
    if (on) cached_ = value;
    else dirty_ = true;
  }

  virtual void onVarAssigned(VariableID vid, Numeric /*newVal*/) {
    int s = slotOf(vid);
    if (!slotAssigned_[s]) {
      ++numAssigned_;
      slotAssigned_[s] = 1;
      dirty_ = true;
    }
  }
  virtual void onVarChanged(VariableID, Numeric /*oldVal*/, Numeric /*newVal*/) { dirty_ = true; }
  virtual void onVarUnassigned(VariableID vid, Numeric /*oldVal*/) {
    int s = slotOf(vid);
    if (slotAssigned_[s]) {
      slotAssigned_[s] = 0;
      dirty_ = true;
      --numAssigned_;
    }
  }

  virtual Numeric evalFactor() const = 0;
  virtual Numeric getDerivative(VariableID /*vid*/) const {
    assert(false);
    return 0;
  }

 protected:
  void attach(Variable* v) {
    vars_.push_back(v);
    slotAssigned_.push_back(0);
    v->addFactor(this);
  }
  int slotOf(VariableID vid) const {
    for (size_t i = 0; i < vars_.size(); ++i)
      if (vars_[i]->getID() == vid) return (int)i;
    assert(false);
    return -1;
  }
  FactorID id_;
  std::vector<Variable*> vars_;
  std::vector<char> slotAssigned_;
  long long numAssigned_ = 0;
  bool assignedConstant_ = false;
  mutable bool dirty_ = true;
  mutable Numeric cached_ = 0;
};

// src/Variable.cpp:66-88
inline void Variable::assign(Numeric newval) {
  if (assigned_) {
    const Numeric tol = 1e-12;
    const bool same = changeFilter() ? (std::fabs(newval - value_) < tol) : false;  // approxeq, src/common.h:66-68
    if (!same) {
      for (size_t k = 0; k < factors_.size(); ++k) factors_[k]->onVarChanged(id_, value_, newval);
    }
  } else {
    assigned_ = true;
    for (size_t k = 0; k < factors_.size(); ++k) factors_[k]->onVarAssigned(id_, newval);
  }
  value_ = newval;
}
// src/Variable.cpp:90-101
inline void Variable::unassign() {
  assert(assigned_);
  assigned_ = false;
  for (size_t k = 0; k < factors_.size(); ++k) factors_[k]->onVarUnassigned(id_, value_);
  value_ = 0;
}

// f = c * prod_i t_i,  t_i = [sin]((x_i - k_i)^{e_i})   (src/NonlinearProductFactor.h:16-21)
class NonlinearProductFactor : public Factor {
 public:
  struct Term {
    Numeric exponent, constant;
    bool useSine;
    bool hasExp() const { return exponent != 1; }
    bool hasConstant() const { return constant != 0; }
  };
  NonlinearProductFactor(FactorID id, Numeric coeff = 1, bool useExponential = false)
      : Factor(id), coeff_(coeff), useExp_(useExponential) {}
  // src/NonlinearProductFactor.cpp:27-54 (a zero exponent drops the variable; duplicates ignored)
  void addVariable(Variable* v, Numeric exponent = 1, Numeric constant = 0, bool useSine = false) {
    if (exponent == 0.0) return;
    for (const Variable* have : vars_)
      if (have == v) return;
    attach(v);
    terms_.push_back(Term{exponent, constant, useSine});
  }
  void setCoeff(Numeric c) { coeff_ = c; }
  Numeric coeff() const { return coeff_; }
  const std::vector<Term>& terms() const { return terms_; }

  // src/NonlinearProductFactor.cpp:186-209
  Numeric evalFactor() const override {
    Numeric prod(1);
    for (size_t i = 0; i < vars_.size(); ++i) {
      Numeric val = vars_[i]->eval();
      const Term& t = terms_[i];
      if (t.hasConstant()) val -= t.constant;
      if (t.hasExp()) val = power(val, t.exponent);
      if (t.useSine) val = orc_sin(val);
      prod *= val;
    }
    Numeric fe = prod;
    if (useExp_) fe = std::exp(-fe);
    fe *= coeff_;
    return fe;
  }
  // src/NonlinearProductFactor.cpp:149-178
  Numeric getDerivative(VariableID vid) const override {
    Numeric prod(1);
    for (size_t i = 0; i < vars_.size(); ++i) {
      const Term& t = terms_[i];
      Numeric val = vars_[i]->eval();
      if (vars_[i]->getID() == vid) {
        if (!t.hasExp() && !t.useSine) continue;  // d/dx of plain x is 1 (note: constant not subtracted, :160)
        val -= t.constant;
        const Numeric inner = val;
        const Numeric innerexp = power(inner, t.exponent);
        val = power(val, t.exponent - 1.0);
        val *= t.exponent;
        if (t.useSine) val *= orc_cos(innerexp);
        prod *= val;
      } else {
        if (t.hasConstant()) val -= t.constant;
        if (t.hasExp()) val = power(val, t.exponent);
        if (t.useSine) val = orc_sin(val);
        prod *= val;
      }
    }
    return prod * coeff_;
  }
  // src/NonlinearProductFactor.cpp:57-117: no assigned-constant early-out
  void computeGradient(PartialGradient& g) const override {
    assert(!useExp_);
    g.clear();
    g.reserve(vars_.size());
    for (const Variable* v : vars_) g[v->getID()] = getDerivative(v->getID());
  }

 private:
  Numeric coeff_;
  bool useExp_;
  std::vector<Term> terms_;
};

// f = c * (k + sum_i a_i x_i^{e_i})^e with an incrementally maintained inner sum
// (src/SimpleSumFactor.cpp:95-186).  CPU-only: used by the ten known-answer functions.
class SimpleSumFactor : public Factor {
 public:
  struct Term {
    Numeric exponent, coeff;
    bool hasExp() const { return exponent != 1; }
    bool hasCoeff() const { return coeff != 1; }
  };
  SimpleSumFactor(FactorID id, Numeric constant = 0, Numeric exponent = 1, Numeric coefficient = 1)
      : Factor(id), constant_(constant), exponent_(exponent), coeff_(coefficient) {}
  void addVariable(Variable* v, Numeric exponent = 1, Numeric coefficient = 1) {
    for (const Variable* have : vars_)
      if (have == v) return;
    attach(v);
    terms_.push_back(Term{exponent, coefficient});
    dirty_ = true;
  }
  Numeric evalVariable(int slot, Numeric val) const {  // :120-126
    const Term& t = terms_[slot];
    if (t.hasExp()) val = power(val, t.exponent);
    if (t.hasCoeff()) val *= t.coeff;
    return val;
  }
  Numeric evalFactor() const override {  // :140-146
    Numeric s = partial_;
    if (constant_ != 0) s += constant_;
    if (exponent_ != 1) s = power(s, exponent_);
    if (coeff_ != 1) s *= coeff_;
    return s;
  }
  Numeric getDerivative(VariableID vid) const override {  // :95-117
    Numeric d = 1.0;
    if (coeff_ != 1) d *= coeff_;
    if (!(std::fabs(exponent_ - 0.0) < 1e-10)) {
      Numeric fe = partial_ + constant_;
      fe = power(fe, exponent_ - 1.0);
      d *= exponent_ * fe;
    }
    const int s = slotOf(vid);
    const Term& t = terms_[s];
    if (t.hasCoeff()) d *= t.coeff;
    if (!(std::fabs(t.exponent - 0.0) < 1e-10)) {
      Numeric ve = vars_[s]->eval();
      ve = power(ve, t.exponent - 1.0);
      d *= t.exponent * ve;
    }
    return d;
  }
  void onVarAssigned(VariableID vid, Numeric nv) override {  // :157-165
    Factor::onVarAssigned(vid, nv);
    partial_ += evalVariable(slotOf(vid), nv);
  }
  void onVarChanged(VariableID vid, Numeric ov, Numeric nv) override {  // :168-175
    Factor::onVarChanged(vid, ov, nv);
    const int s = slotOf(vid);
    partial_ -= evalVariable(s, ov);
    partial_ += evalVariable(s, nv);
  }
  void onVarUnassigned(VariableID vid, Numeric ov) override {  // :178-186
    Factor::onVarUnassigned(vid, ov);
    partial_ -= evalVariable(slotOf(vid), ov);
  }

 private:
  Numeric constant_, exponent_, coeff_;
  Numeric partial_ = 0;
  std::vector<Term> terms_;
};

// Slot order of a bundle-adjustment factor's 12 variables
// (src/bundleadjust/BundleAdjustmentCommon.h:36-59, SEPARATE_THETA_VAR undefined).
enum BASlot { ROT_X = 0, ROT_Y, ROT_Z, TRANS_X, TRANS_Y, TRANS_Z, FOCAL, RDL_K1, RDL_K2, PT_X, PT_Y, PT_Z, BA_NSLOTS };

struct BAForward {
  Numeric axis[3];   // normalised rotation axis (or the raw vector when theta == 0)
  Numeric theta;
  Numeric axp[3];    // axis x point
  Numeric adp;       // axis . point
  Numeric P[3];      // point in the camera frame
  Numeric pp[2];     // after perspective division
  Numeric r2, dist;  // radial distortion
  Numeric pix[2];
  Numeric res[2];
};

// Forward model; returns the factor value 1/2 |pix - obs|^2.
// src/bundleadjust/BundleAdjustmentFactor.cpp:266-335 (rotate+translate),
// BundleAdjustmentFactor.h:80-107 (divide, distort, error), BundleAdjustmentCommon.h:80-93 (normalize).
inline Numeric ba_forward(const Numeric* x, Numeric obsx, Numeric obsy, BAForward& m) {
  const Numeric raw[3] = {x[ROT_X], x[ROT_Y], x[ROT_Z]};
  Numeric* P = m.P;
  P[0] = x[PT_X];
  P[1] = x[PT_Y];
  P[2] = x[PT_Z];
  const Numeric norm = std::sqrt(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2]);
  for (int i = 0; i < 3; ++i) m.axis[i] = (norm != 0.0) ? raw[i] / norm : raw[i];
  m.theta = norm;
  const Numeric* a = m.axis;
  // cross(a, P): BundleAdjustmentCommon.h:64-68
  m.axp[0] = a[1] * P[2] - a[2] * P[1];
  m.axp[1] = a[2] * P[0] - a[0] * P[2];
  m.axp[2] = a[0] * P[1] - a[1] * P[0];
  if (m.theta > 0.0) {
    const Numeric c = orc_cos(m.theta);
    const Numeric s = orc_sin(m.theta);
    const Numeric omc = 1 - c;
    m.adp = a[0] * P[0] + a[1] * P[1] + a[2] * P[2];
    for (int i = 0; i < 3; ++i) P[i] = P[i] * c + m.axp[i] * s + a[i] * omc * m.adp;  // :305-307
  } else {
    m.adp = 0;
    for (int i = 0; i < 3; ++i) P[i] = P[i] + m.axp[i];  // first-order rotation, :326-329
  }
  P[0] += x[TRANS_X];
  P[1] += x[TRANS_Y];
  P[2] += x[TRANS_Z];
  m.pp[0] = -P[0] / P[2];
  m.pp[1] = -P[1] / P[2];
  m.r2 = m.pp[0] * m.pp[0] + m.pp[1] * m.pp[1];
  m.dist = 1 + m.r2 * (x[RDL_K1] + x[RDL_K2] * m.r2);
  m.pix[0] = x[FOCAL] * m.dist * m.pp[0];
  m.pix[1] = x[FOCAL] * m.dist * m.pp[1];
  m.res[0] = (m.pix[0] - obsx);
  m.res[1] = (m.pix[1] - obsy);
  return (m.res[0] * m.res[0] + m.res[1] * m.res[1]) / 2.0;
}

// Closed-form gradient in slot order; src/bundleadjust/BundleAdjustmentFactor.cpp:351-554.
// Expressions keep the reference's operand order; only the bookkeeping (arrays and
// loops instead of ~90 named scalars) differs.
inline Numeric ba_gradient(const Numeric* x, Numeric obsx, Numeric obsy, Numeric* grad) {
  BAForward m;
  const Numeric fval = ba_forward(x, obsx, obsy, m);
  const Numeric f = x[FOCAL], k1 = x[RDL_K1], k2 = x[RDL_K2];
  const Numeric t1 = 2.0 * (k1 + 2.0 * k2 * m.r2);
  const Numeric q[3] = {x[PT_X], x[PT_Y], x[PT_Z]};
  const Numeric s = orc_sin(m.theta), c = orc_cos(m.theta);
  const Numeric* a = m.axis;
  const Numeric* P = m.P;
  const Numeric vnorm = m.theta;
  const Numeric P22 = P[2] * P[2];
  const Numeric pp00 = m.pp[0] * m.pp[0], pp01 = m.pp[0] * m.pp[1], pp11 = m.pp[1] * m.pp[1];
  // d pix / d pp  (:386-389)
  const Numeric J[2][2] = {{m.dist + t1 * pp00, t1 * pp01}, {t1 * pp01, m.dist + t1 * pp11}};

  // Chain a 3-vector dP (derivative of the camera-frame point) through
  // divide -> distort -> residual, :418-422 and its repeats.
#ifdef ORACLE_TWIN_RECIP
  // perturbation twin (not reference behaviour): every quotient by P_z^2, P_z and |r| becomes a product with
  // the denominator's reciprocal (<= 1 ulp away per quotient) — what round 1's device gradient did
  const Numeric iP22 = 1.0 / P22, iP2 = 1.0 / P[2], ivn = 1.0 / vnorm;
#define ORC_DIV_P22(x) ((x) * iP22)
#define ORC_DIV_P2(x) ((x) * iP2)
#define ORC_DIV_VN(x) ((x) * ivn)
#else
#define ORC_DIV_P22(x) ((x) / P22)
#define ORC_DIV_P2(x) ((x) / P[2])
#define ORC_DIV_VN(x) ((x) / vnorm)
#endif
  auto through_projection = [&](const Numeric dP[3]) -> Numeric {
    const Numeric dppx = ORC_DIV_P22(P[0] * dP[2] - P[2] * dP[0]);
    const Numeric dppy = ORC_DIV_P22(P[1] * dP[2] - P[2] * dP[1]);
    const Numeric drx = m.res[0] * (J[0][0] * dppx + J[0][1] * dppy);
    const Numeric dry = m.res[1] * (J[1][0] * dppx + J[1][1] * dppy);
    return f * (drx + dry);
  };

  // sign of the Levi-Civita symbol used for the s-terms: row i, column j, third index k
  static const int third[3][3] = {{-1, 2, 1}, {2, -1, 0}, {1, 0, -1}};
  static const Numeric sgnA[3][3] = {{0, 1, -1}, {-1, 0, 1}, {1, -1, 0}};  // d P_i / d axis_j  (:391-404)

  // d P / d (normalised axis) and d P / d theta
  Numeric dPda[3][3], dPdth[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      if (i == j)
        dPda[i][j] = (m.adp + a[i] * q[i]) * (1 - c);
      else
        dPda[i][j] = (sgnA[i][j] * q[third[i][j]]) * s + a[i] * q[j] * (1 - c);
    }
    dPdth[i] = -q[i] * s + m.axp[i] * c + a[i] * m.adp * s;
  }
  // rotation-vector components: d axis / d raw_m (:408-410, :426-428, :444-446), d theta / d raw_m = axis_m
  for (int mth = 0; mth < 3; ++mth) {
    Numeric dadv[3];
    for (int j = 0; j < 3; ++j) {
      if (j == mth) {
        const int o1 = (mth == 0) ? 1 : 0;
        const int o2 = (mth == 2) ? 1 : 2;
        dadv[j] = ORC_DIV_VN(a[o1] * a[o1] + a[o2] * a[o2]);
      } else {
        const int lo = std::min(j, mth), hi = std::max(j, mth);
        dadv[j] = ORC_DIV_VN(-a[lo] * a[hi]);
      }
    }
    Numeric dP[3];
    for (int i = 0; i < 3; ++i)
      dP[i] = dPda[i][0] * dadv[0] + dPda[i][1] * dadv[1] + dPda[i][2] * dadv[2] + dPdth[i] * a[mth];
    grad[ROT_X + mth] = through_projection(dP);
  }
  // point coordinates: columns of the rotation matrix (:480-512)
  static const Numeric sgnQ[3][3] = {{0, -1, 1}, {1, 0, -1}, {-1, 1, 0}};  // d P_i / d q_j
  for (int j = 0; j < 3; ++j) {
    Numeric dP[3];
    for (int i = 0; i < 3; ++i) {
      if (i == j) {
        dP[i] = c * (1.0 - a[i] * a[i]) + a[i] * a[i];
      } else {
        const int lo = std::min(i, j), hi = std::max(i, j);
        dP[i] = (sgnQ[i][j] * a[third[i][j]]) * s + a[lo] * a[hi] * (1.0 - c);
      }
    }
    grad[PT_X + j] = through_projection(dP);
  }
  // translation (:514-526)
  grad[TRANS_X] = ORC_DIV_P2((m.res[0] * J[0][0] + m.res[1] * J[1][0]) * -f);
  grad[TRANS_Y] = ORC_DIV_P2((m.res[0] * J[0][1] + m.res[1] * J[1][1]) * -f);
  {
    const Numeric dpx = J[0][0] * P[0] + J[0][1] * P[1];
    const Numeric dpy = J[1][0] * P[0] + J[1][1] * P[1];
    grad[TRANS_Z] = ORC_DIV_P22((m.res[0] * dpx + m.res[1] * dpy) * f);
  }
  // intrinsics (:528-538)
  grad[FOCAL] = m.res[0] * (m.dist * m.pp[0]) + m.res[1] * (m.dist * m.pp[1]);
  grad[RDL_K1] = m.res[0] * (f * m.r2 * m.pp[0]) + m.res[1] * (f * m.r2 * m.pp[1]);
  grad[RDL_K2] = m.res[0] * (f * m.r2 * m.r2 * m.pp[0]) + m.res[1] * (f * m.r2 * m.r2 * m.pp[1]);
  return fval;
}

class BundleAdjustmentFactor : public Factor {
 public:
  BundleAdjustmentFactor(FactorID id, long long cam, long long pt, Numeric ox, Numeric oy)
      : Factor(id), cam_(cam), pt_(pt), ox_(ox), oy_(oy) {}
  void addVariable(Variable* v) { attach(v); }  // must be called in BASlot order
  long long camera() const { return cam_; }
  long long point() const { return pt_; }
  Numeric obsX() const { return ox_; }
  Numeric obsY() const { return oy_; }
  Numeric evalFactor() const override {  // :55-64,160-165
    Numeric vals[BA_NSLOTS];
    for (int i = 0; i < BA_NSLOTS; ++i) vals[i] = vars_[i]->eval();
    BAForward m;
    return ba_forward(vals, ox_, oy_, m);
  }
  void computeGradient(PartialGradient& g) const override {  // :338-348
    Numeric vals[BA_NSLOTS], gr[BA_NSLOTS];
    for (int i = 0; i < BA_NSLOTS; ++i) vals[i] = vars_[i]->eval();
    ba_gradient(vals, ox_, oy_, gr);
    for (int i = 0; i < BA_NSLOTS; ++i) g[vars_[i]->getID()] = gr[i];
  }

 private:
  long long cam_, pt_;
  Numeric ox_, oy_;
};

// The plugin base the optimizers talk to (src/OptimizableFunction.h).
class OptimizableFunction {
 public:
  OptimizableFunction() {}
  OptimizableFunction(const OptimizableFunction&) = delete;
  ~OptimizableFunction() {
    for (Factor* f : factors) delete f;
    for (Variable* v : variables) delete v;
  }
  Variable* addVariable(Domain d) {
    Variable* v = new Variable((VariableID)variables.size(), d);
    variables.push_back(v);
    return v;
  }
  void onVarAssigned(VariableID, Numeric) {}  // no-op hook, src/OptimizableFunction.h:65-69

  // src/OptimizableFunction.cpp:95-135, MinSum semiring: Product = +, identity 0.
  Numeric evalFactors(const std::vector<Factor*>& fs, bool useCached = true) {
#ifdef ORACLE_TWIN_TREEFOLD
    // perturbation twin (not reference behaviour): the factor values are summed pairwise (a balanced tree, the
    // shape of a GPU butterfly / block reduction) instead of left to right
    std::vector<Numeric> vals;
    vals.reserve(fs.size());
    for (Factor* fp : fs) {
      const Factor& f = *fp;
      if (!f.isAssigned() && !f.areAllVarsAssigned()) continue;
      vals.push_back(useCached ? f.eval(counters) : f.evalNoCache());
    }
    if (vals.empty()) return 0.0;
    for (size_t stride = 1; stride < vals.size(); stride *= 2)
      for (size_t i = 0; i + stride < vals.size(); i += 2 * stride) vals[i] = vals[i] + vals[i + stride];
    return vals[0];
#else
    Numeric feval = 0.0;
    for (Factor* fp : fs) {
      const Factor& f = *fp;
      if (!f.isAssigned() && !f.areAllVarsAssigned()) continue;
      const Numeric nf = useCached ? f.eval(counters) : f.evalNoCache();
      feval = feval + nf;
    }
    return feval;
#endif
  }
  // src/OptimizableFunction.cpp:248-262
  void computeGradientOfSum(const std::vector<Factor*>& fs, PartialGradient& gradient) {
    gradient.clear();
    gradient.reserve(variables.size());
    PartialGradient pg;
    for (const Factor* f : fs) {
      pg.clear();
      f->computeGradient(pg);
      ++counters.factor_grad_calls;
      productGradient(gradient, pg);
    }
  }
  // src/OptimizableFunction.cpp:138-178 restricted to "assign everything, sweep uncached"
  Numeric eval() { return evalFactors(factors, true); }

  std::vector<Variable*> variables;
  std::vector<Factor*> factors;
  std::vector<Numeric> xinit;  // initial state carried by a loaded problem (BundleAdjustmentFunction.h)
  Counters counters;
  // Optional evaluation trace (test instrumentation, not reference behaviour): every
  // SubfunctionFD call appends {is_df, assigned point, value or gradient}.
  struct TraceRec {
    int is_df;
    std::vector<Numeric> x, out;
  };
  bool trace_on = false;
  std::vector<TraceRec> trace;
  enum Kind { KIND_NLPF = 0, KIND_BA = 1, KIND_SIMPLESUM = 2 } kind = KIND_NLPF;
  long long ncams = 0, npts = 0;  // BA only
};

// src/SubspaceOptimizer.{h,cpp}
class SubspaceOptimizer {
 public:
  explicit SubspaceOptimizer(OptimizableFunction& f_) : f(f_), maxiters(50), ftol(3.0e-8) {}
  virtual ~SubspaceOptimizer() {}
  void setParameters(long long ssmaxit, Numeric ssftol) {  // SSmaxit / SSftol, :23-35
    if (ssmaxit > 0) maxiters = (size_t)ssmaxit;
    if (ssftol > 0) ftol = ssftol;
  }
  virtual Numeric optimize(const std::vector<Variable*>& vars, const std::vector<Factor*>& factors,
                           std::vector<Numeric>& xval, Numeric& deltaFval, bool printdbg) = 0;
  size_t lastIters = 0;
  Numeric lastInitialFval = 0;  // initialFval of the last optimize() (test instrumentation)

 protected:
  void quickAssignVals(const std::vector<Variable*>& vars, const std::vector<Numeric>& xval, bool sanitize) {  // :38-53
    for (size_t i = 0; i < vars.size(); ++i) {
      const Numeric val = sanitize ? vars[i]->getDomain().closestVal(xval[i]) : xval[i];
      vars[i]->assign(val);
      f.onVarAssigned(vars[i]->getID(), val);
    }
  }
  OptimizableFunction& f;
  size_t maxiters;
  Numeric ftol;
};

// src/optimizers/CGDSubspaceOptimizer.{h,cpp}
class CGDSubspaceOptimizer : public SubspaceOptimizer {
 public:
  explicit CGDSubspaceOptimizer(OptimizableFunction& f_) : SubspaceOptimizer(f_) {}

  // :102-184
  struct SubfunctionFD {
    OptimizableFunction& func;
    const std::vector<Variable*>& vars;
    const std::vector<Factor*>& facs;
    PartialGradient& pg;
    Numeric operator()(const std::vector<Numeric>& x) {  // :124-132
      quickAssignVals(x);
      ++func.counters.f_evals;
      const Numeric r = func.evalFactors(facs, true);
      if (func.trace_on) {
        OptimizableFunction::TraceRec t;
        t.is_df = 0;
        for (const Variable* v : vars) t.x.push_back(v->eval());
        t.out.push_back(r);
        func.trace.push_back(t);
      }
      return r;
    }
    void df(const std::vector<Numeric>& x, std::vector<Numeric>& deriv) {  // :135-157
      deriv.assign(vars.size(), 0);
      quickAssignVals(x);
      ++func.counters.df_evals;
      pg.clear();
      func.computeGradientOfSum(facs, pg);
      for (size_t i = 0; i < vars.size(); ++i) {
        const Numeric* d = pg.find(vars[i]->getID());
        deriv[i] = (d == nullptr ? 0 : *d);
      }
      if (func.trace_on) {
        OptimizableFunction::TraceRec t;
        t.is_df = 1;
        for (const Variable* v : vars) t.x.push_back(v->eval());
        t.out = deriv;
        func.trace.push_back(t);
      }
    }
    bool quickAssignVals(const std::vector<Numeric>& xval) {  // :160-184
      bool res = true;
      for (size_t i = 0; i < vars.size(); ++i) {
        const Numeric val = vars[i]->getDomain().closestVal(xval[i]);
        assert(!std::isnan(xval[i]));
        if (val != xval[i]) res = false;
        vars[i]->assign(val);
        func.onVarAssigned(vars[i]->getID(), val);
      }
      return res;
    }
  };

  // :19-98
  Numeric optimize(const std::vector<Variable*>& vars, const std::vector<Factor*>& gdfs,
                   std::vector<Numeric>& xval, Numeric& deltaFval, bool /*printdbg*/) override {
    assert(xval.size() == vars.size());
    if (gdfs.empty()) {
      deltaFval = 0;
      lastInitialFval = 0;
      return 0;
    }
    quickAssignVals(vars, xval, true);
    SubfunctionFD sfd{f, vars, gdfs, pgtmp};
    const Numeric initialFval = sfd(xval);
    lastInitialFval = initialFval;
    const std::vector<Numeric> initxval(xval.begin(), xval.end());
#ifdef ORACLE_USE_REFERENCE_NRC
    rdis::nrc::Frprmn<SubfunctionFD> gdmin(sfd, (int)maxiters, ftol);
#else
    nr::PolakRibiere<SubfunctionFD> gdmin(sfd, (int)maxiters, ftol);
#endif
    try {
      gdmin.minimize(xval);
    } catch (const char*) {
      // "Too many iterations ..." is the normal exit at maxiters (:42-58)
    }
    sfd.quickAssignVals(gdmin.p);
    Numeric fret = gdmin.fret;
    if (fret > initialFval) {  // :66-80
      sfd.quickAssignVals(initxval);
      fret = sfd(initxval);
    }
    for (size_t i = 0; i < vars.size(); ++i) xval[i] = vars[i]->eval();
    deltaFval = (fret - initialFval);
    lastIters = (size_t)gdmin.iter;
    return fret;
  }

 private:
  PartialGradient pgtmp;
};

}  // namespace oracle
