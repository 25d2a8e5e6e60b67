// rdis_host.cpp — see rdis_host.h.  Everything that evaluates a factor calls the C-ABI of
// include/rdis_gpu.h; nothing here computes a factor value on the host.
#include "rdis_host.h"

#include <algorithm>
#include <limits>
#include <cassert>
#include <chrono>
#include <cstring>
#include <iostream>
#include <numeric>

#include "../../include/rdis_gpu.h"

namespace rdis {

// ------------------------------------------------------------------------------------------
// Variable / Factor
// ------------------------------------------------------------------------------------------
Span<Factor*> Variable::getFactors() const {
  if (!m_owner) return Span<Factor*>{nullptr, 0};
  m_owner->ensureIncidence();
  const size_t a = m_owner->incOff[(size_t)m_id], b = m_owner->incOff[(size_t)m_id + 1];
  return Span<Factor*>{m_owner->incList.data() + a, b - a};
}

void Variable::assign(Numeric newval, bool notifyFactors) {  // src/Variable.cpp:66-88
  if (m_isAssigned) {
    // The reference skips the cache-invalidation fan-out when |new - old| <= 1e-12 but still stores the
    // new value (:69-78, :87).  The host keeps no per-factor cache (the device's strict mode does), and the default
    // onVarChanged is a no-op, so the fan-out is skipped altogether.
  } else {
    m_isAssigned = true;
    if (notifyFactors)
      for (Factor* f : getFactors()) f->onVarAssigned(m_id, newval);
  }
  m_value = newval;
  if (m_owner) m_owner->noteAssigned(m_id);  // queued for the next host->device flush
}

void Variable::unassign() {  // src/Variable.cpp:90-102
  if (!m_isAssigned) throw std::logic_error("Variable::unassign on an unassigned variable");
  m_isAssigned = false;
  for (Factor* f : getFactors()) f->onVarUnassigned(m_id, m_value);
  m_value = 0;
}

void Factor::addVariable(Variable* vp) {
  if (!m_owner) throw std::logic_error("Factor::addVariable: factors are created by an OptimizableFunction");
  std::vector<Variable*>& pool = m_owner->varPool;
  if (nvars == 0) voff = pool.size();
  if (voff + nvars != pool.size()) throw std::logic_error("Factor::addVariable: only the most recently created factor can grow");
  pool.push_back(vp);
  ++nvars;
  m_owner->incidenceValid = false;
  if (vp->isAssigned()) ++numVarsAssigned;
}

Span<Variable*> Factor::getVariables() const {
  return Span<Variable*>{m_owner ? m_owner->varPool.data() + voff : nullptr, nvars};
}

void NonlinearProductFactor::addVariable(Variable* vp, Numeric exponent, Numeric constant, bool useSine) {
  Factor::addVariable(vp);
  std::vector<Term>& tp = m_owner->termPool;
  if (tp.size() + 1 != m_owner->varPool.size()) tp.resize(m_owner->varPool.size() - 1);  // bundle-adjustment factors carry no terms
  tp.push_back(Term{exponent, constant, useSine});
}

Span<NonlinearProductFactor::Term> NonlinearProductFactor::getTerms() const {
  return Span<Term>{m_owner ? m_owner->termPool.data() + voff : nullptr, nvars};
}

void Factor::assign(Numeric fval, VariableID assignmentKey) {
  isAssignedConstant = true;
  vidAssigned = assignmentKey;
  assignedVal = fval;
  if (m_owner) m_owner->noteFactorConst(this);
}

void Factor::unassign(VariableID assignmentKey) {
  if (!isAssignedConstant || vidAssigned != assignmentKey) throw std::logic_error("Factor::unassign: wrong assignment key");
  isAssignedConstant = false;
  vidAssigned = -1;
  if (m_owner) m_owner->noteFactorConst(this);
}

// ------------------------------------------------------------------------------------------
// OptimizableFunction
// ------------------------------------------------------------------------------------------
OptimizableFunction::OptimizableFunction() : kind(-1), ncams(0), npts(0), ctx(nullptr) {}

OptimizableFunction::~OptimizableFunction() {
  for (CachedBatch& c : batchCache)
    if (c.batch) rdisgpu_batch_destroy(c.batch);
  if (ctx) rdisgpu_destroy(ctx);  // variables and factors live in the arenas (the reference deletes them one by one, :43-54)
}

OptimizableFunction::CachedBatch* OptimizableFunction::findWave(const std::vector<ComponentProblem>& problems) {
  const size_t n = problems.size();
  for (CachedBatch& c : batchCache) {
    if (c.var_off.size() != n + 1 || c.pvars.size() != (size_t)c.var_off[n] || c.pfacs.size() != (size_t)c.fac_off[n]) continue;
    bool same = true;
    for (size_t k = 0; k < n && same; ++k) {
      const ComponentProblem& p = problems[k];
      const size_t v0 = (size_t)c.var_off[k], nv = (size_t)c.var_off[k + 1] - v0;
      const size_t f0 = (size_t)c.fac_off[k], nf = (size_t)c.fac_off[k + 1] - f0;
      same = p.vars.size() == nv && p.factors.size() == nf &&
             (nv == 0 || std::memcmp(p.vars.data(), c.pvars.data() + v0, nv * sizeof(Variable*)) == 0) &&
             (nf == 0 || std::memcmp(p.factors.data(), c.pfacs.data() + f0, nf * sizeof(Factor*)) == 0);
    }
    if (same) {
      c.stamp = ++batchClock;
      return &c;
    }
  }
  return nullptr;
}

void OptimizableFunction::rememberWave(CachedBatch& c, const std::vector<ComponentProblem>& problems) {
  c.pvars.clear();
  c.pfacs.clear();
  c.pvars.reserve(c.vids.size());
  c.pfacs.reserve(c.fids.size());
  for (const ComponentProblem& p : problems) {
    c.pvars.insert(c.pvars.end(), p.vars.begin(), p.vars.end());
    c.pfacs.insert(c.pfacs.end(), p.factors.begin(), p.factors.end());
  }
}

OptimizableFunction::CachedBatch* OptimizableFunction::cachedBatch(const std::vector<int64_t>& var_off, const std::vector<int32_t>& vids,
                                                                   const std::vector<int64_t>& fac_off, const std::vector<int64_t>& fids) {
  auto mix = [](unsigned long long h, const void* p, size_t n) {  // 8 bytes at a time; equality is checked on a hit anyway
    const unsigned char* b = static_cast<const unsigned char*>(p);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
      unsigned long long w;
      std::memcpy(&w, b + i, 8);
      h = (h ^ w) * 0x9E3779B97F4A7C15ULL;
      h ^= h >> 29;
    }
    for (; i < n; ++i) h = (h ^ b[i]) * 1099511628211ULL;
    return h;
  };
  unsigned long long key = 1469598103934665603ULL;
  key = mix(key, var_off.data(), var_off.size() * 8);
  key = mix(key, vids.data(), vids.size() * 4);
  key = mix(key, fac_off.data(), fac_off.size() * 8);
  key = mix(key, fids.data(), fids.size() * 8);
  for (CachedBatch& c : batchCache)
    if (c.key == key && c.vids == vids && c.fids == fids && c.var_off == var_off && c.fac_off == fac_off) {
      c.stamp = ++batchClock;
      return &c;
    }
  constexpr size_t kMaxCached = 8;
  CachedBatch* slot = nullptr;
  if (batchCache.size() < kMaxCached) {
    batchCache.emplace_back();
    slot = &batchCache.back();
  } else {  // evict the least recently used
    slot = &batchCache[0];
    for (CachedBatch& c : batchCache)
      if (c.stamp < slot->stamp) slot = &c;
    rdisgpu_batch_destroy(slot->batch);
    slot->batch = nullptr;
  }
  slot->key = key;
  slot->stamp = ++batchClock;
  slot->vids = vids;
  slot->fids = fids;
  slot->var_off = var_off;
  slot->fac_off = fac_off;
  check(rdisgpu_batch_create_csr(ctx, (int64_t)var_off.size() - 1, var_off.data(), vids.data(), fac_off.data(), fids.data(), &slot->batch),
        "rdisgpu_batch_create_csr");
  slot->pvars.clear();
  slot->pfacs.clear();
  return slot;
}

void OptimizableFunction::reserve(size_t nFactors, size_t nEdges) {
  factors.reserve(nFactors);
  varPool.reserve(nEdges);
  if (kind != Factor::BUNDLE_ADJUSTMENT) termPool.reserve(nEdges);
}

void OptimizableFunction::ensureIncidence() const {
  if (incidenceValid) return;
  const size_t V = variables.size();
  incOff.assign(V + 1, 0);
  for (const Variable* v : varPool) ++incOff[(size_t)v->getID() + 1];
  for (size_t i = 0; i < V; ++i) incOff[i + 1] += incOff[i];
  incList.resize(varPool.size());
  std::vector<size_t> cur(incOff.begin(), incOff.end() - 1);
  for (Factor* f : factors)  // ascending factor id: the order the reference's per-variable lists are filled in
    for (const Variable* v : f->getVariables()) incList[cur[(size_t)v->getID()]++] = f;
  incidenceValid = true;
}

void OptimizableFunction::check(int rc, const char* what) const {
  if (rc == RDISGPU_OK) return;
  std::string msg = std::string(what) + ": " + rdisgpu_last_error(ctx);
  throw std::runtime_error(msg);
}

Variable* OptimizableFunction::addVariable(Numeric lb, Numeric ub) {
  if (ctx) throw std::logic_error("addVariable after init");
  variableArena.emplace_back((VariableID)variables.size(), VariableDomain(lb, ub));
  Variable* v = &variableArena.back();
  v->m_owner = this;
  variables.push_back(v);
  incidenceValid = false;
  return v;
}

NonlinearProductFactor* OptimizableFunction::addProductFactor(Numeric coefficient) {
  if (ctx) throw std::logic_error("addProductFactor after init");
  if (kind == Factor::BUNDLE_ADJUSTMENT) throw std::logic_error("one factor family per function");
  kind = Factor::NONLINEAR_PRODUCT;
  productArena.emplace_back((FactorID)factors.size(), coefficient);
  NonlinearProductFactor* f = &productArena.back();
  f->m_owner = this;
  factors.push_back(f);
  return f;
}

void OptimizableFunction::declareBundleAdjustment(int32_t ncams_, int32_t npts_) {
  if (kind == Factor::NONLINEAR_PRODUCT) throw std::logic_error("one factor family per function");
  if ((VariableCount)variables.size() != 9LL * ncams_ + 3LL * npts_)
    throw std::logic_error("declareBundleAdjustment: expected 9*ncams + 3*npts variables");
  kind = Factor::BUNDLE_ADJUSTMENT;
  ncams = ncams_;
  npts = npts_;
}

BundleAdjustmentFactor* OptimizableFunction::addObservation(int32_t cam, int32_t pt, Numeric obsx, Numeric obsy) {
  if (ctx) throw std::logic_error("addObservation after init");
  if (kind != Factor::BUNDLE_ADJUSTMENT) throw std::logic_error("declareBundleAdjustment first");
  if (cam < 0 || cam >= ncams || pt < 0 || pt >= npts) throw std::out_of_range("addObservation: camera / point id");
  observationArena.emplace_back((FactorID)factors.size(), cam, pt, obsx, obsy);
  BundleAdjustmentFactor* f = &observationArena.back();
  f->m_owner = this;
  for (int p = 0; p < 9; ++p) f->addVariable(variables[9LL * cam + p]);                // getCamVID
  for (int d = 0; d < 3; ++d) f->addVariable(variables[9LL * ncams + 3LL * pt + d]);   // getPointVID
  factors.push_back(f);
  return f;
}

void OptimizableFunction::exportProductFactors(std::vector<int64_t>& rowptr, std::vector<int32_t>& vid, std::vector<double>& expo,
                                               std::vector<double>& konst, std::vector<uint8_t>& sine, std::vector<double>& coeff) const {
  // the flat arrays of rdisgpu_add_nlpf: the pools ARE the CSR (factors are appended in id order), one linear pass
  const size_t F = factors.size(), E = varPool.size();
  rowptr.resize(F + 1);
  coeff.resize(F);
  vid.resize(E);
  expo.resize(E);
  konst.resize(E);
  sine.resize(E);
  rowptr[F] = (int64_t)E;
  for (size_t j = F; j-- > 0;) {
    const NonlinearProductFactor* nf = static_cast<const NonlinearProductFactor*>(factors[j]);
    rowptr[j] = (nf->nvars == 0) ? rowptr[j + 1] : (int64_t)nf->voff;  // a factor without variables owns an empty run
    coeff[j] = nf->getCoefficient();
  }
  for (size_t e = 0; e < E; ++e) {
    vid[e] = (int32_t)varPool[e]->getID();
    expo[e] = termPool[e].exponent;
    konst[e] = termPool[e].constant;
    sine[e] = termPool[e].useSine ? 1 : 0;
  }
}

void OptimizableFunction::init(int device) {
  if (ctx) throw std::logic_error("OptimizableFunction::init called twice");
  if (variables.empty() || factors.empty() || kind < 0) throw std::logic_error("init: empty function");
  int rc = rdisgpu_create(&ctx, device);
  if (rc != RDISGPU_OK) {
    ctx = nullptr;
    throw std::runtime_error(std::string("rdisgpu_create: ") + rdisgpu_last_error(nullptr) + " (no CPU fallback)");
  }
  const int64_t V = (int64_t)variables.size(), F = (int64_t)factors.size();
  std::vector<double> lb(V), ub(V);
  for (int64_t i = 0; i < V; ++i) {
    lb[i] = variables[i]->getDomain().min();
    ub[i] = variables[i]->getDomain().max();
  }
  check(rdisgpu_set_vars(ctx, V, lb.data(), ub.data()), "rdisgpu_set_vars");
  if (kind == Factor::NONLINEAR_PRODUCT) {
    std::vector<int64_t> rowptr;
    std::vector<int32_t> vid;
    std::vector<double> expo, konst, coeff;
    std::vector<uint8_t> sine;
    exportProductFactors(rowptr, vid, expo, konst, sine, coeff);
    check(rdisgpu_add_nlpf(ctx, F, rowptr.data(), vid.data(), expo.data(), konst.data(), sine.data(), coeff.data()),
          "rdisgpu_add_nlpf");
  } else {
    std::vector<int32_t> cam(F), pt(F);
    std::vector<double> obs(2 * F);
    for (int64_t j = 0; j < F; ++j) {
      const BundleAdjustmentFactor* bf = static_cast<const BundleAdjustmentFactor*>(factors[j]);
      cam[j] = bf->getCamera();
      pt[j] = bf->getPoint();
      obs[2 * j] = bf->obsX();
      obs[2 * j + 1] = bf->obsY();
    }
    check(rdisgpu_add_ba(ctx, F, cam.data(), pt.data(), obs.data(), ncams, npts), "rdisgpu_add_ba");
  }
  check(rdisgpu_finalize(ctx), "rdisgpu_finalize");
  dirtyFlag.assign((size_t)V, 0);
  dirtyVids.clear();
  for (Variable* v : variables)
    if (v->isAssigned()) noteAssigned(v->getID());
  for (Factor* f : factors)
    if (f->isAssigned()) dirtyFactors.push_back(f);
}

void OptimizableFunction::noteAssigned(VariableID vid) {
  if (dirtyFlag.empty()) return;  // before init: init() collects every assigned variable
  if (!dirtyFlag[(size_t)vid]) {
    dirtyFlag[(size_t)vid] = 1;
    dirtyVids.push_back((int32_t)vid);
  }
}

void OptimizableFunction::noteFactorConst(Factor* f) {
  if (!ctx) return;
  dirtyFactors.push_back(f);
}

void OptimizableFunction::flushAssignments() {
  if (!ctx) throw std::logic_error("flushAssignments before init");
  if (!dirtyVids.empty()) {
    std::vector<double> x(dirtyVids.size());
    size_t n = 0;
    for (int32_t vid : dirtyVids) {
      dirtyFlag[(size_t)vid] = 0;
      if (!variables[(size_t)vid]->isAssigned()) continue;  // unassigned again since: its value is never read
      dirtyVids[n] = vid;
      x[n] = variables[(size_t)vid]->m_value;
      ++n;
    }
    if (n) check(rdisgpu_set_x(ctx, (int64_t)n, dirtyVids.data(), x.data()), "rdisgpu_set_x");
    dirtyVids.clear();
  }
  if (!dirtyFactors.empty()) {
    std::vector<int64_t> fid;
    std::vector<double> val;
    std::vector<uint8_t> on;
    for (Factor* f : dirtyFactors) {
      fid.push_back(f->getID());
      val.push_back(f->assignedVal);
      on.push_back(f->isAssigned() ? 1 : 0);
    }
    check(rdisgpu_set_factor_const(ctx, (int64_t)fid.size(), fid.data(), val.data(), on.data()), "rdisgpu_set_factor_const");
    dirtyFactors.clear();
  }
}

Numeric OptimizableFunction::eval() {
  Numeric ferr = 0;
  return evalFactors(factors, ferr);
}

Numeric OptimizableFunction::evalFactors(const FactorPtrVec& fctrs, Numeric& ferr) {
  ferr = 0;  // no partial simplification on this path (approxfactors error is tracked by the tree search)
  if (fctrs.empty()) return 0;
  flushAssignments();
  double sum = 0;
  if (&fctrs == &factors) {
    // the all-factor streaming sweep, when no factor would be skipped: every variable assigned
    bool all = true;
    for (const Variable* v : variables)
      if (!v->isAssigned()) {
        all = false;
        break;
      }
    if (all) {
      check(rdisgpu_eval(ctx, 0, nullptr, &sum, nullptr), "rdisgpu_eval");
      return sum;
    }
  }
  // the reference skips factors that are not fully assigned (src/OptimizableFunction.cpp:108-112)
  std::vector<int64_t> fid;
  fid.reserve(fctrs.size());
  for (const Factor* fp : fctrs)
    if (fp->areAllVarsAssigned() || fp->isAssigned()) fid.push_back(fp->getID());
  if (fid.empty()) return 0;
  check(rdisgpu_eval(ctx, (int64_t)fid.size(), fid.data(), &sum, nullptr), "rdisgpu_eval");
  return sum;
}

void OptimizableFunction::computeGradient(const FactorPtrVec& facs, const VariablePtrVec& vars, NumericVec& gradient) {
  gradient.assign(vars.size(), 0.0);
  if (vars.empty()) return;
  flushAssignments();
  std::vector<int64_t> fid(facs.size());
  std::vector<int32_t> vid(vars.size());
  for (size_t i = 0; i < facs.size(); ++i) fid[i] = facs[i]->getID();
  for (size_t i = 0; i < vars.size(); ++i) vid[i] = (int32_t)vars[i]->getID();
  // a non-null pointer with nf == 0 means "no factors" (null would mean "all factors")
  static const int64_t none = 0;
  check(rdisgpu_grad(ctx, (int64_t)fid.size(), fid.empty() ? &none : fid.data(), (int64_t)vid.size(), vid.data(), gradient.data()),
        "rdisgpu_grad");
}

NumericInterval OptimizableFunction::computeBounds(const FactorPtrVec& fctrs, VariableID assignedVID) {
  NumericInterval out{0.0, 0.0};  // semiring Product identity (MinSum), src/OptimizableFunction.cpp:192
  std::vector<int64_t> fid;
  for (const Factor* f : fctrs)  // :196-213
    if (assignedVID < 0 || (f->isAssigned() && f->getAssignedKey() == assignedVID) || !f->isAssigned()) fid.push_back(f->getID());
  if (fid.empty()) return out;
  flushAssignments();
  std::vector<uint8_t> assigned(variables.size());
  for (size_t v = 0; v < variables.size(); ++v) assigned[v] = variables[v]->isAssigned() ? 1 : 0;
  double sum[2] = {0.0, 0.0};
  check(rdisgpu_bounds(ctx, assigned.data(), (int64_t)fid.size(), fid.data(), nullptr, nullptr, sum), "rdisgpu_bounds");
  out.lo = sum[0];
  out.hi = sum[1];
  return out;
}

void OptimizableFunction::computeBoundsBatch(const std::vector<FactorPtrVec>& lists, std::vector<NumericInterval>& out) {
  out.assign(lists.size(), NumericInterval{0.0, 0.0});
  if (lists.empty()) return;
  std::vector<int64_t> off(1, 0), fid;
  for (const FactorPtrVec& l : lists) {
    for (const Factor* f : l) fid.push_back(f->getID());
    off.push_back((int64_t)fid.size());
  }
  flushAssignments();
  std::vector<uint8_t> assigned(variables.size());
  for (size_t v = 0; v < variables.size(); ++v) assigned[v] = variables[v]->isAssigned() ? 1 : 0;
  std::vector<double> sums(2 * lists.size(), 0.0);
  check(rdisgpu_bounds_lists(ctx, assigned.data(), (int64_t)lists.size(), off.data(), fid.data(), sums.data()), "rdisgpu_bounds_lists");
  for (size_t l = 0; l < lists.size(); ++l) out[l] = NumericInterval{sums[2 * l], sums[2 * l + 1]};
}

// ------------------------------------------------------------------------------------------
// SubspaceOptimizer
// ------------------------------------------------------------------------------------------
SubspaceOptimizer::SubspaceOptimizer(OptimizableFunction& f_)
    : f(f_), doAscent(!f_.isMinSum()), maxiters(50), ftol(3.0e-8) {  // src/SubspaceOptimizer.cpp:12-18
  if (doAscent) throw "only the MinSum semiring is supported";           // src/RDISOptimizer.cpp:57-61
}

void SubspaceOptimizer::setParameters(const ParameterMap& options) {  // src/SubspaceOptimizer.cpp:22-35
  if (options.count("SSmaxit")) maxiters = (size_t)options.at("SSmaxit");
  if (options.count("SSftol")) ftol = options.at("SSftol");
  if (maxiters == 0) throw std::invalid_argument("SSmaxit must be positive");
}

void SubspaceOptimizer::quickAssignVals(const VariablePtrVec& vars, const NumericVec& xval, bool sanitizeVals) {
  assert(vars.size() == xval.size());  // src/SubspaceOptimizer.cpp:38-53
  for (size_t i = 0; i < vars.size(); ++i) {
    Numeric val = sanitizeVals ? vars[i]->getDomain().closestVal(xval[i]) : xval[i];
    if (val != xval[i]) std::cout << "var " << vars[i]->getID() << " xval " << xval[i] << " changed to " << val << std::endl;
    vars[i]->assign(val);
    f.onVarAssigned(vars[i]->getID(), val);
  }
}

// ------------------------------------------------------------------------------------------
// CudaSubspaceOptimizer
// ------------------------------------------------------------------------------------------
Numeric CudaSubspaceOptimizer::optimize(const VariablePtrVec& vars, const FactorPtrVec& gdfs, NumericVec& xval,
                                        Numeric& deltaFval, const bool printdbg) {
  assert(xval.size() == vars.size());
  if (gdfs.empty()) {  // src/optimizers/CGDSubspaceOptimizer.cpp:26-29
    deltaFval = 0;
    return 0;
  }
  std::vector<ComponentProblem> one(1);
  one[0].vars = vars;
  one[0].factors = gdfs;
  one[0].xval = xval;
  optimizeBatch(one, printdbg);
  xval = one[0].xval;
  deltaFval = one[0].deltaFval;
  return one[0].fval;
}

Numeric CudaSubspaceOptimizer::optimizeBatch(std::vector<ComponentProblem>& problems, const bool printdbg) {
  const int64_t n = (int64_t)problems.size();
  if (n == 0) return 0;
  typedef std::chrono::steady_clock Clock;
  const Clock::time_point t0 = Clock::now();
  size_t listed = 0;
  for (const ComponentProblem& p : problems) {
    if (p.xval.size() != p.vars.size()) throw std::invalid_argument("optimizeBatch: xval / vars size mismatch");
    listed += p.vars.size() + p.factors.size();
  }
  // worth keeping resident: many problems, or few (one) with long index lists (ladybug: 49 camera components list 31 843
  // factors; the 4755-variable top-level block 11 950)
  const bool keep = !useLM && (n >= 64 || listed >= 4096);
  // A wave the tree search comes back to (same objects, same order) is recognised from its pointer runs: no id lists
  // are rebuilt, only the start values are packed.
  OptimizableFunction::CachedBatch* wave = keep ? f.findWave(problems) : nullptr;
  x0.clear();
  if (wave) {
    for (const ComponentProblem& p : problems) x0.insert(x0.end(), p.xval.begin(), p.xval.end());
  } else {
    var_off.assign(1, 0);
    fac_off.assign(1, 0);
    vids.clear();
    fids.clear();
    for (ComponentProblem& p : problems) {
      for (size_t i = 0; i < p.vars.size(); ++i) {
        vids.push_back((int32_t)p.vars[i]->getID());
        x0.push_back(p.xval[i]);  // the device clamps start values into the domain (quickAssignVals(vars, xval, true), CGD.cpp:33)
      }
      for (const Factor* fp : p.factors) fids.push_back(fp->getID());
      var_off.push_back((int64_t)vids.size());
      fac_off.push_back((int64_t)fids.size());
    }
    if (keep) {
      // a wave: its index lists stay resident (a revisit of the same sibling set uploads start values only)
      wave = f.cachedBatch(var_off, vids, fac_off, fids);
      f.rememberWave(*wave, problems);
    }
  }
  const std::vector<int64_t>& voff = wave ? wave->var_off : var_off;
  // every other variable the factors read must be current on the device
  const Clock::time_point tf0 = Clock::now();
  f.flushAssignments();
  const Clock::time_point tf1 = Clock::now();
  xout.resize(x0.size());
  finit.resize((size_t)n);
  fend.resize((size_t)n);
  iters.resize((size_t)n);
  status.resize((size_t)n);
  nfe.resize((size_t)n);
  nge.resize((size_t)n);
  const Clock::time_point t1 = Clock::now();
  if (useLM) {
    const double opts[4] = {1e-3, 1e-15, 1e-15, ftol};  // src/optimizers/LMSubspaceOptimizer.cpp:83-86
    f.check(rdisgpu_solve_lm_csr(f.device(), n, var_off.data(), vids.data(), fac_off.data(), fids.data(), x0.data(), (int)maxiters,
                                 opts, xout.data(), finit.data(), fend.data(), iters.data(), status.data(), nfe.data(), nge.data()),
            "rdisgpu_solve_lm_csr");
  } else if (wave) {
    f.check(rdisgpu_batch_solve_cgd(wave->batch, x0.data(), (int)maxiters, ftol), "rdisgpu_batch_solve_cgd");
    tm.fetch_ms -= std::chrono::duration<double, std::milli>(Clock::now() - t1).count();  // fetch_ms = device_ms - time to enqueue
    f.check(rdisgpu_batch_fetch_csr(wave->batch, xout.data(), finit.data(), fend.data(), iters.data(), status.data(), nfe.data(),
                                    nge.data()),
            "rdisgpu_batch_fetch_csr");
  } else {
    f.check(rdisgpu_solve_cgd_csr(f.device(), n, var_off.data(), vids.data(), fac_off.data(), fids.data(), x0.data(), (int)maxiters,
                                  ftol, xout.data(), finit.data(), fend.data(), iters.data(), status.data(), nfe.data(), nge.data()),
            "rdisgpu_solve_cgd_csr");
  }
  const Clock::time_point t2 = Clock::now();
  Numeric total = 0;
  for (int64_t k = 0; k < n; ++k) {
    ComponentProblem& p = problems[(size_t)k];
    if (p.factors.empty()) {  // nothing to optimise: xval untouched, returns 0 (CGD.cpp:26-29)
      p.fval = 0;
      p.deltaFval = 0;
      p.iters = 0;
      p.status = status[(size_t)k];
      continue;
    }
    for (size_t i = 0; i < p.vars.size(); ++i) {
      const Numeric v = xout[(size_t)voff[(size_t)k] + i];
      p.xval[i] = v;
      Variable* var = p.vars[i];
      if (!var->isAssigned()) {
        var->assign(v);  // factor bookkeeping (assigned counts); the queued upload is dropped below
      }
      var->setFromDevice(v);  // the solve committed the value on the device (CGD.cpp:61,84-86)
      f.onVarAssigned(var->getID(), v);
    }
    p.fval = fend[(size_t)k];
    p.deltaFval = fend[(size_t)k] - finit[(size_t)k];
    p.iters = iters[(size_t)k];
    p.status = status[(size_t)k];
    total += p.fval;
    if (printdbg)
      std::cout << "CUDA CGD subspace result (steps " << p.iters << "): " << p.fval << ", diff: " << p.deltaFval
                << ", init: " << finit[(size_t)k] << std::endl;
  }
  // values written by the solve are already resident: drop their queued uploads
  for (int32_t vid : f.dirtyVids) f.dirtyFlag[(size_t)vid] = 0;
  f.dirtyVids.clear();
  const Clock::time_point t3 = Clock::now();
  tm.pack_ms += std::chrono::duration<double, std::milli>(t1 - t0).count() - std::chrono::duration<double, std::milli>(tf1 - tf0).count();
  tm.flush_ms += std::chrono::duration<double, std::milli>(tf1 - tf0).count();
  tm.device_ms += std::chrono::duration<double, std::milli>(t2 - t1).count();
  if (wave) tm.fetch_ms += std::chrono::duration<double, std::milli>(t2 - t1).count();
  tm.writeback_ms += std::chrono::duration<double, std::milli>(t3 - t2).count();
  ++tm.calls;
  return total;
}

// ------------------------------------------------------------------------------------------
// ComponentBatcher
// ------------------------------------------------------------------------------------------
namespace {
struct DisjointSets {
  std::vector<int64_t> parent;
  explicit DisjointSets(size_t n) : parent(n) { std::iota(parent.begin(), parent.end(), (int64_t)0); }
  int64_t find(int64_t a) {
    while (parent[(size_t)a] != a) {
      parent[(size_t)a] = parent[(size_t)parent[(size_t)a]];
      a = parent[(size_t)a];
    }
    return a;
  }
  void unite(int64_t a, int64_t b) {
    a = find(a);
    b = find(b);
    if (a != b) parent[(size_t)std::max(a, b)] = std::min(a, b);  // root = smallest vertex: deterministic labels
  }
};
}  // namespace

void ComponentBatcher::createChildren(const OptimizableFunction& func, const VariableIDVec& componentVars,
                                      std::vector<ChildComponent>& children) {
  children.clear();
  OptimizableFunction& fn = const_cast<OptimizableFunction&>(func);
  const VariablePtrVec& vars = fn.getVariables();
  const FactorPtrVec& facs = func.getFactors();
  const int64_t V = (int64_t)vars.size(), F = (int64_t)facs.size();
  // vertices: variables 0..V-1, factors V..V+F-1 (ConnectivityGraph's vertex numbering)
  DisjointSets ds((size_t)(V + F));
  std::vector<uint8_t> inComp((size_t)V, 0);
  for (VariableID vid : componentVars)
    if (!vars[(size_t)vid]->isAssigned()) inComp[(size_t)vid] = 1;
  std::vector<uint8_t> factorSeen((size_t)F, 0);
  for (VariableID vid : componentVars) {
    if (!inComp[(size_t)vid]) continue;
    for (const Factor* fp : vars[(size_t)vid]->getFactors()) {
      if (fp->isAssigned()) continue;  // an assigned factor has no edges (onFactorAssigned, ConnectivityGraph.cpp:258-271)
      ds.unite(vid, V + fp->getID());
      factorSeen[(size_t)fp->getID()] = 1;
    }
  }
  // edges to unassigned variables outside `componentVars` cannot exist for a proper component; a
  // factor reaching one would make the caller's set not closed — include such variables defensively
  // is NOT done: the reference asserts closure implicitly through its connectivity graph.
  std::map<int64_t, size_t> rootToChild;
  for (VariableID vid : componentVars) {
    if (!inComp[(size_t)vid]) continue;
    const int64_t r = ds.find(vid);
    auto it = rootToChild.find(r);
    if (it == rootToChild.end()) {
      it = rootToChild.emplace(r, children.size()).first;
      children.emplace_back();
    }
    children[it->second].vars.push_back(vid);
  }
  for (FactorID fid = 0; fid < F; ++fid) {
    if (!factorSeen[(size_t)fid]) continue;
    const int64_t r = ds.find(V + fid);
    children[rootToChild.at(r)].factors.push_back(fid);
  }
  for (ChildComponent& c : children) {
    std::sort(c.vars.begin(), c.vars.end());
    std::sort(c.factors.begin(), c.factors.end());
  }
  std::stable_sort(children.begin(), children.end(), [](const ChildComponent& a, const ChildComponent& b) {
    if (a.vars.size() != b.vars.size()) return a.vars.size() < b.vars.size();
    return a.vars.front() < b.vars.front();
  });
}

void ComponentBatcher::createChildrenOnDevice(OptimizableFunction& func, const VariableIDVec& componentVars,
                                              std::vector<ChildComponent>& children) {
  children.clear();
  if (!func.device()) throw std::logic_error("createChildrenOnDevice before init");
  func.flushAssignments();  // assigned-constant factors carry no edges: the device overlay must be current
  const VariablePtrVec& vars = func.getVariables();
  const int64_t V = (int64_t)vars.size(), F = (int64_t)func.getFactors().size();
  std::vector<uint8_t> assigned((size_t)V, 1);  // everything outside the component counts as removed
  for (VariableID vid : componentVars)
    if (!vars[(size_t)vid]->isAssigned()) assigned[(size_t)vid] = 0;
  std::vector<int32_t> vlabel((size_t)V), flabel((size_t)std::max<int64_t>(F, 1));
  int32_t ncomp = 0;
  if (rdisgpu_components(func.device(), assigned.data(), vlabel.data(), flabel.data(), &ncomp, nullptr) != RDISGPU_OK)
    throw std::runtime_error(std::string("rdisgpu_components: ") + rdisgpu_last_error(func.device()));
  // label = smallest variable id of the component: children in ascending label order, then the reference's size order
  std::vector<int32_t> slot((size_t)V, -1);
  children.reserve((size_t)ncomp);
  for (int64_t v = 0; v < V; ++v) {
    const int32_t l = vlabel[(size_t)v];
    if (l < 0) continue;
    if (slot[(size_t)l] < 0) {
      slot[(size_t)l] = (int32_t)children.size();
      children.emplace_back();
    }
    children[(size_t)slot[(size_t)l]].vars.push_back(v);  // ascending v: already sorted
  }
  for (int64_t f = 0; f < F; ++f) {
    const int32_t l = flabel[(size_t)f];
    if (l >= 0) children[(size_t)slot[(size_t)l]].factors.push_back(f);
  }
  std::stable_sort(children.begin(), children.end(), [](const ChildComponent& a, const ChildComponent& b) {
    if (a.vars.size() != b.vars.size()) return a.vars.size() < b.vars.size();
    return a.vars.front() < b.vars.front();
  });
}

void ComponentBatcher::optimizeSiblings(CudaSubspaceOptimizer& ssopt, OptimizableFunction& func, std::vector<ComponentProblem>& children,
                                        const std::vector<NumericInterval>& uab, Numeric parentFmin, Numeric childFmin0, bool useBounds,
                                        SiblingWave& out) {
  const size_t n = children.size();
  if (uab.size() != n) throw std::invalid_argument("optimizeSiblings: one unassigned bound per child");
  out.outcome.assign(n, SIB_PRUNED);
  out.value.assign(n, std::numeric_limits<Numeric>::quiet_NaN());
  out.evaluated = 0;
  // the state to roll back to for children the reference would not have optimised
  std::vector<NumericVec> before(n);
  std::vector<std::vector<char>> wasAssigned(n);
  for (size_t k = 0; k < n; ++k) {
    before[k] = children[k].xval;
    for (const Variable* v : children[k].vars) wasAssigned[k].push_back(v->isAssigned() ? 1 : 0);
  }
  ssopt.optimizeBatch(children, false);  // speculative: every sibling, one device call

  // replay of the sequential loop (src/RDISOptimizer.cpp:291-314)
  Numeric childFmin = childFmin0;
  Numeric assignedLB = 0;
  for (size_t k = 0; k < n; ++k) assignedLB += uab[k].lower();  // no child evaluated yet (Component.cpp:322-336)
  std::vector<char> rollback(n, 1);
  for (size_t k = 0; k < n; ++k) {
    const Numeric fmin_k = childFmin + uab[k].lower();           // Component::computeFMin
    Numeric fx;
    if (useBounds && fmin_k < uab[k].lower()) {                  // checkUnassignedBound: not optimised, set to its bound
      out.outcome[k] = SIB_BOUND_SKIPPED;
      fx = uab[k].lower();
    } else {
      out.outcome[k] = SIB_OPTIMISED;
      fx = children[k].fval;
      rollback[k] = 0;
    }
    out.value[k] = fx;
    ++out.evaluated;
    childFmin += uab[k].lower();                                 // Component::onChildEvaluated
    childFmin -= fx;
    assignedLB += fx - uab[k].lower();
    if (k + 1 < n && useBounds && parentFmin <= assignedLB) break;  // checkAssignedBound: the rest is never visited
  }
  out.childFmin = childFmin;
  out.assignedLower = assignedLB;
  // roll back what the reference would not have touched: host objects and, through Variable::assign, the device mirror
  for (size_t k = 0; k < n; ++k) {
    if (!rollback[k]) continue;
    ComponentProblem& p = children[k];
    for (size_t i = 0; i < p.vars.size(); ++i) {
      if (wasAssigned[k][i]) {
        p.vars[i]->assign(before[k][i]);
      } else if (p.vars[i]->isAssigned()) {
        p.vars[i]->unassign();
      }
      p.xval[i] = before[k][i];
    }
    p.fval = out.value[k];
    p.deltaFval = 0;
  }
  func.flushAssignments();
}

void ComponentBatcher::leafProblem(OptimizableFunction& func, const ChildComponent& child, const NumericVec& fallback,
                                   ComponentProblem& out) {
  VariablePtrVec& vars = func.getVariables();
  const FactorPtrVec& facs = func.getFactors();
  out.vars.clear();
  out.factors.clear();
  out.xval.clear();
  std::vector<uint8_t> mine;  // membership of the child's variables, by position in a sorted list
  for (VariableID vid : child.vars) {
    Variable* v = vars[(size_t)vid];
    out.vars.push_back(v);
    out.xval.push_back(v->isAssigned() ? v->eval() : fallback.at((size_t)vid));
  }
  for (FactorID fid : child.factors) {
    const Factor* fp = facs[(size_t)fid];
    bool ready = true;  // every variable is either assigned or about to be (a variable of this problem)
    for (const Variable* v : fp->getVariables()) {
      if (v->isAssigned()) continue;
      if (!std::binary_search(child.vars.begin(), child.vars.end(), v->getID())) {
        ready = false;
        break;
      }
    }
    if (ready) out.factors.push_back(const_cast<Factor*>(fp));
  }
}

}  // namespace rdis
