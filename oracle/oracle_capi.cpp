// ORACLE — TEST INFRASTRUCTURE ONLY (see rdis_oracle.hpp).  Flat C API over the CPU
// restatement so that tests / bench.py's cpu_baseline leg can drive it through ctypes,
// plus restatements of the reference's problem builders:
//   BundleAdjustmentFunction::load / setDomain   src/bundleadjust/BundleAdjustmentFunction.cpp:50-250,402-477
//   PolynomialFunction::load / readFactor        src/PolynomialFunction.cpp:60-215
//   makeHighDimSinusoid                          src/OptimizableFunctionGenerator.cpp:660-760
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

#include "rdis_oracle.hpp"
#include "lm_oracle.hpp"
#include "interval_oracle.hpp"

using namespace oracle;

namespace {

struct Handle {
  std::unique_ptr<OptimizableFunction> fn;
  std::unique_ptr<CGDSubspaceOptimizer> cgd;
  std::unique_ptr<LMSubspaceOptimizer> lmopt;
  Handle() : fn(new OptimizableFunction()), cgd(new CGDSubspaceOptimizer(*fn)), lmopt(new LMSubspaceOptimizer(*fn)) {}
};

Handle* H(void* h) { return static_cast<Handle*>(h); }

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}
std::string lower(std::string s) {
  for (char& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
bool parse_number(const std::string& s, double& out) {  // lexical_cast<Numeric>: whole token must parse
  if (s.empty()) return false;
  char* end = nullptr;
  out = std::strtod(s.c_str(), &end);
  return end != nullptr && *end == '\0';
}
Domain parse_domain(const std::string& s) {  // src/VariableDomain.cpp:63-75, single "lo:hi"
  std::vector<std::string> tok;
  std::string cur;
  for (char c : s) {
    if (c == ' ' || c == '~' || c == ':' || c == ',') {
      if (!cur.empty()) tok.push_back(cur), cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  if (!cur.empty()) tok.push_back(cur);
  Domain d;
  if (tok.size() >= 2) {
    parse_number(tok.front(), d.lo);
    parse_number(tok.back(), d.hi);
  }
  return d;
}

// src/bundleadjust/BundleAdjustmentFunction.cpp:402-477 (intervals use the no-rounding policy,
// src/common.h:46-60, so plain arithmetic reproduces them).
void ba_domain(int slot, double init, Domain& dom, double& slo, double& shi) {
  const double dsf = 1000.0;
  const double pi = 3.141592653589793238462643383279502884;
  double dlo, dhi;
  auto scale = [](double lo, double hi, double k, double& olo, double& ohi) {
    double a = lo * k, b = hi * k;
    olo = std::min(a, b);
    ohi = std::max(a, b);
  };
  switch (slot) {
    case ROT_X: case ROT_Y: case ROT_Z:
      slo = -1 * pi; shi = 1 * pi;
      scale(slo, shi, dsf, dlo, dhi);
      break;
    case TRANS_X: case TRANS_Y: case TRANS_Z: case PT_X: case PT_Y: case PT_Z:
      slo = init + -1 * 1e2; shi = init + 1 * 1e2;
      scale(slo, shi, dsf, dlo, dhi);
      break;
    case FOCAL: {
      slo = init + -1 * 1e2; shi = init + 1 * 1e2;
      double lowerb = std::min(slo, slo * dsf);
      lowerb = std::max(lowerb, 0.0);
      dlo = lowerb; dhi = shi * dsf;
      break;
    }
    case RDL_K1:
      slo = init + -1 * 1e-4; shi = init + 1 * 1e-4;
      dlo = -1e-1; dhi = 1e-1;
      break;
    default:  // RDL_K2
      slo = init + -1 * 1e-6; shi = init + 1 * 1e-6;
      dlo = -1e-3; dhi = 1e-3;
      break;
  }
  dom.lo = std::min(dlo, slo);  // hull(dom, sit), :468
  dom.hi = std::max(dhi, shi);
}

void build_ba(OptimizableFunction& fn, long long ncams, long long npts, const double* lb, const double* ub,
              long long F, const int32_t* cam, const int32_t* pt, const double* obs) {
  fn.kind = OptimizableFunction::KIND_BA;
  fn.ncams = ncams;
  fn.npts = npts;
  const long long V = 9 * ncams + 3 * npts;
  for (long long v = 0; v < V; ++v) fn.addVariable(Domain{lb[v], ub[v]});
  for (long long j = 0; j < F; ++j) {
    auto* f = new BundleAdjustmentFactor(j, cam[j], pt[j], obs[2 * j], obs[2 * j + 1]);
    fn.factors.push_back(f);
    for (int p = 0; p < 9; ++p) f->addVariable(fn.variables[9 * cam[j] + p]);               // getCamVID, BundleAdjustmentFunction.h:88-90
    for (int d = 0; d < 3; ++d) f->addVariable(fn.variables[9 * ncams + 3 * pt[j] + d]);    // getPointVID, :93-96
  }
}

void build_nlpf(OptimizableFunction& fn, long long V, const double* lb, const double* ub, long long F,
                const int64_t* rowptr, const int32_t* vid, const double* expo, const double* konst,
                const uint8_t* sine, const double* coeff) {
  fn.kind = OptimizableFunction::KIND_NLPF;
  for (long long v = 0; v < V; ++v) fn.addVariable(Domain{lb[v], ub[v]});
  for (long long j = 0; j < F; ++j) {
    auto* f = new NonlinearProductFactor(j, coeff[j], false);
    fn.factors.push_back(f);
    for (int64_t e = rowptr[j]; e < rowptr[j + 1]; ++e)
      f->addVariable(fn.variables[vid[e]], expo[e], konst[e], sine[e] != 0);
  }
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

void* orc_create_nlpf(int64_t V, const double* lb, const double* ub, int64_t F, const int64_t* rowptr,
                      const int32_t* vid, const double* expo, const double* konst, const uint8_t* sine,
                      const double* coeff) {
  Handle* h = new Handle();
  build_nlpf(*h->fn, V, lb, ub, F, rowptr, vid, expo, konst, sine, coeff);
  return h;
}

void* orc_create_ba(int32_t ncams, int32_t npts, const double* lb, const double* ub, int64_t F,
                    const int32_t* cam, const int32_t* pt, const double* obs_xy) {
  Handle* h = new Handle();
  build_ba(*h->fn, ncams, npts, lb, ub, F, cam, pt, obs_xy);
  return h;
}

// BAL text format, first `ncams` cameras / `npts` points (<=0: all).
void* orc_load_bal(const char* path, int64_t ncams_in, int64_t npts_in) {
  std::ifstream ifs(path);
  if (!ifs.is_open()) return nullptr;
  long long nc = 0, np = 0, nobs = 0;
  ifs >> nc >> np >> nobs;
  const long long ncams = ncams_in <= 0 ? nc : ncams_in;
  const long long npts = npts_in <= 0 ? np : npts_in;
  if (ncams > nc || npts > np) return nullptr;
  std::vector<int32_t> cam, pt;
  std::vector<double> obs;
  for (long long i = 0; i < nobs; ++i) {
    long long c, p;
    double ox, oy;
    ifs >> c >> p >> ox >> oy;
    if (c >= ncams || p >= npts) continue;  // :167-169
    cam.push_back((int32_t)c);
    pt.push_back((int32_t)p);
    obs.push_back(ox);
    obs.push_back(oy);
  }
  const long long V = 9 * ncams + 3 * npts;
  std::vector<double> x0(V), lb(V), ub(V), slo(V), shi(V);
  for (long long c = 0; c < nc; ++c) {
    for (int p = 0; p < 9; ++p) {
      double val;
      ifs >> val;
      if (c >= ncams) continue;
      const long long v = 9 * c + p;
      x0[v] = val;
      Domain d;
      ba_domain(p, val, d, slo[v], shi[v]);
      lb[v] = d.lo;
      ub[v] = d.hi;
    }
  }
  for (long long i = 0; i < np; ++i) {
    for (int d3 = 0; d3 < 3; ++d3) {
      double val;
      ifs >> val;
      if (i >= npts) continue;
      const long long v = 9 * ncams + 3 * i + d3;
      x0[v] = val;
      Domain d;
      ba_domain(PT_X + d3, val, d, slo[v], shi[v]);
      lb[v] = d.lo;
      ub[v] = d.hi;
    }
  }
  if (!ifs) return nullptr;
  Handle* h = new Handle();
  build_ba(*h->fn, ncams, npts, lb.data(), ub.data(), (long long)cam.size(), cam.data(), pt.data(), obs.data());
  h->fn->xinit = x0;
  for (long long v = 0; v < V; ++v) {
    h->fn->variables[v]->samp_lo = slo[v];
    h->fn->variables[v]->samp_hi = shi[v];
  }
  return h;
}

// Polynomial text format (data/testpoly.txt header describes it).
void* orc_load_poly(const char* path) {
  std::ifstream ifs(path);
  if (!ifs.is_open()) return nullptr;
  Handle* h = new Handle();
  OptimizableFunction& fn = *h->fn;
  fn.kind = OptimizableFunction::KIND_NLPF;
  Domain defdom;
  std::vector<std::string> names;
  std::vector<char> explicit_dom;
  auto get_var = [&](const std::string& name, const Domain* d) -> Variable* {  // OptimizableFunction::addVariable, :57-76
    for (size_t i = 0; i < names.size(); ++i)
      if (names[i] == name) return fn.variables[i];
    names.push_back(name);
    explicit_dom.push_back(d != nullptr);
    return fn.addVariable(d ? *d : defdom);
  };
  std::string line;
  while (std::getline(ifs, line)) {
    if (line.empty() || line[0] == '#') continue;
    size_t eq = line.find('=');
    if (eq != std::string::npos) {  // readVariable, :97-116
      std::string name = trim(line.substr(0, eq));
      Domain d = parse_domain(trim(line.substr(eq + 1)));
      if (lower(name) == "default") {
        defdom = d;  // applies to variables created from here on (:104-105)
      } else {
        get_var(name, &d);
      }
      continue;
    }
    if (trim(line).empty()) continue;
    auto* f = new NonlinearProductFactor((FactorID)fn.factors.size());  // readFactor, :150-215
    fn.factors.push_back(f);
    std::stringstream ss(line);
    std::string piece;
    while (std::getline(ss, piece, ',')) {
      size_t caret = piece.find('^');
      std::string name = lower(trim(piece.substr(0, caret)));
      double exponent = 1.0;
      Variable* v = nullptr;
      if (caret == std::string::npos) {
        double konst;
        if (parse_number(name, konst)) f->setCoeff(konst);
        else v = get_var(name, nullptr);
      } else {
        v = get_var(name, nullptr);
        parse_number(trim(piece.substr(caret + 1)), exponent);
      }
      if (v != nullptr) f->addVariable(v, exponent, 0, false);
    }
  }
  return h;
}

// makeHighDimSinusoid(treeHeight, branches, maxArity, allowOddArityFactors)
void* orc_make_sinusoid(int64_t treeHeight, int64_t branches, int64_t maxArity, int allowOdd) {
  const double twopi = 2.000001 * 3.141592653;
  char buf[64];
  std::snprintf(buf, sizeof buf, "%g", 10 * twopi);  // boost::format default stream precision (6 significant digits), :669-671
  double bound = 0;
  parse_number(buf, bound);
  Handle* h = new Handle();
  OptimizableFunction& fn = *h->fn;
  fn.kind = OptimizableFunction::KIND_NLPF;
  maxArity = std::min<int64_t>(maxArity, treeHeight + 1);
  const long long hh = treeHeight, k = branches;
  const long long nvars = (k == 1) ? hh + 1
                                   : (long long)((std::round(std::pow((double)k, (double)hh + 1)) - 1) / (k - 1));  // :682-683
  for (long long v = 0; v < nvars; ++v) {
    Variable* var = fn.addVariable(Domain{-bound, bound});
    var->samp_lo = -twopi;
    var->samp_hi = twopi;
  }
  FactorID fid = 0;
  std::vector<Variable*> chain;
  for (long long ar = 1; ar <= maxArity; ++ar) {
    if (ar > 1 && (ar & 1) && !allowOdd) continue;  // :698
    long long lasth = hh;
    for (long long vid = nvars - 1; vid >= 0; --vid) {
      const long long lastVidAtNextH =
          (k == 1) ? lasth - 1
                   : (long long)((std::round(std::pow((double)k, (double)lasth)) - 1.0) / (k - 1.0) - 1.0);  // :704-706
      const long long varheight = (vid > lastVidAtNextH ? lasth : --lasth);
      if (varheight + 1 < ar) continue;
      chain.clear();
      long long cur = vid;
      for (long long c = 0; c < ar; ++c) {  // walk towards the root, :719-726
        chain.push_back(fn.variables[cur]);
        cur = (long long)std::floor(((double)cur - 1.0) / (double)k);
      }
      auto* f = new NonlinearProductFactor(fid++, ar > 1 ? 12 : 0.6, false);
      fn.factors.push_back(f);
      while (!chain.empty()) {  // root-most ancestor first, :734-737
        f->addVariable(chain.back(), 1, 0, ar > 1);
        chain.pop_back();
      }
    }
  }
  for (long long vid = 0; vid < nvars; ++vid) {  // 0.1 x^2 per variable, :743-747
    auto* f = new NonlinearProductFactor(fid++, 0.1, false);
    fn.factors.push_back(f);
    f->addVariable(fn.variables[vid], 2, 0, false);
  }
  return h;
}

// The ten known-answer functions of the reference's debug test (src/main.cpp:103-135), built from
// SimpleSumFactors as src/OptimizableFunctionGenerator.cpp:215-543 builds them.  Each is described
// here as a table: default domain, per-variable domain overrides, and per factor
// {constant, exponent, coefficient, {variable, term exponent, term coefficient}...}.
// Returns nullptr for idx outside 0..9; *expected receives the minimum main.cpp asserts (tol 1e-5).
void* orc_make_debug_function(int idx, double* expected) {
  struct T { int v; double e, a; };
  struct FDesc { double k, e, c; std::vector<T> terms; };
  struct Desc { double lo, hi; int nvars; std::vector<FDesc> fs; double want; std::vector<std::pair<int, Domain>> over; };
  auto lin = [](std::initializer_list<int> vs) {  // plain sum of variables: k=0, e=1, c=1
    FDesc f{0, 1, 1, {}};
    for (int v : vs) f.terms.push_back(T{v, 1, 1});
    return f;
  };
  auto with_exp = [](FDesc f, double e) { f.e = e; return f; };
  Desc d;
  switch (idx) {
    case 0:  // makeSimplePoly :215-236
      d = Desc{-3, 4, 3, {lin({0, 1}), lin({0, 2})}, -12, {}};
      break;
    case 1:  // makeNonDecompPoly :266-288
      d = Desc{-2, 4, 3, {lin({0, 1}), lin({0, 2}), lin({1, 2})}, -12, {}};
      break;
    case 2:  // makeSetToConstFactor :291-296 (third factor raised to the power 0)
      d = Desc{-2, 4, 3, {lin({0, 1}), lin({0, 2}), with_exp(lin({1, 2}), 0)}, -7, {}};
      break;
    case 3:  // makeSetToConstFactor2 :299-322
      d = Desc{1, 4, 3, {lin({0, 1}), lin({0, 2}), FDesc{0, 1, 1, {T{1, -6, 1}, T{2, -6, 1}}}}, 5.01399, {}};
      break;
    case 4:  // make2ComponentPoly :387-412
      d = Desc{1, 4, 6, {lin({0, 2}), lin({0, 3}), lin({1, 4}), lin({1, 5})}, 8, {}};
      break;
    case 5:  // makeTreePoly :325-349
      d = Desc{1, 4, 5, {lin({0, 1}), lin({0, 2}), lin({1, 3}), lin({2, 4})}, 8, {}};
      break;
    case 6:  // makeMinStateWrongPoly :239-262
      d = Desc{1, 4, 2, {FDesc{0, 2, -1, {T{0, 1, 1}, T{1, 1, 1}}}}, -3.0625, {{0, Domain{-2, 1}}, {1, Domain{0.25, 0.5}}}};
      break;
    case 7:  // makeCrossPoly :478-512
      d = Desc{1, 1.3, 8, {lin({0, 2, 4, 6}), lin({1, 3, 5, 7}), lin({0, 2, 5, 7}), lin({1, 3, 4, 6})}, 16, {}};
      break;
    case 8:  // makePowellsFunction :515-549
      d = Desc{-5, 5, 4,
               {FDesc{0, 2, 0.5, {T{0, 1, 1}, T{1, 1, 10}}}, FDesc{0, 2, 5.0 * 0.5, {T{2, 1, 1}, T{3, 1, -1}}},
                FDesc{0, 4, 0.5, {T{1, 1, 1}, T{2, 1, -2}}}, FDesc{0, 4, 10.0 * 0.5, {T{0, 1, 1}, T{3, 1, -1}}}},
               0, {}};
      break;
    case 9: {  // makeTreePoly3 :416-474: fourteen squared pair sums
      const int pairs[14][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 2}, {1, 3}, {1, 4}, {2, 5}, {2, 6}, {3, 7}, {3, 8}, {4, 9}, {4, 10}, {0, 5}};
      d = Desc{-2, 2, 11, {}, 0, {}};
      for (auto& p : pairs) d.fs.push_back(with_exp(lin({p[0], p[1]}), 2));
      break;
    }
    default:
      return nullptr;
  }
  Handle* h = new Handle();
  OptimizableFunction& fn = *h->fn;
  fn.kind = OptimizableFunction::KIND_SIMPLESUM;
  for (int v = 0; v < d.nvars; ++v) {
    Domain dom{d.lo, d.hi};
    for (auto& o : d.over)
      if (o.first == v) dom = o.second;
    Variable* var = fn.addVariable(dom);
    var->samp_lo = dom.lo;
    var->samp_hi = dom.hi;
  }
  for (size_t j = 0; j < d.fs.size(); ++j) {
    auto* f = new SimpleSumFactor((FactorID)j, d.fs[j].k, d.fs[j].e, d.fs[j].c);
    fn.factors.push_back(f);
    for (const T& t : d.fs[j].terms) f->addVariable(fn.variables[t.v], t.e, t.a);
  }
  if (expected) *expected = d.want;
  return h;
}

void orc_destroy(void* h) { delete H(h); }

int64_t orc_num_vars(void* h) { return (int64_t)H(h)->fn->variables.size(); }
int64_t orc_num_factors(void* h) { return (int64_t)H(h)->fn->factors.size(); }
int orc_kind(void* h) { return (int)H(h)->fn->kind; }
int64_t orc_ncams(void* h) { return H(h)->fn->ncams; }
int64_t orc_npts(void* h) { return H(h)->fn->npts; }
int64_t orc_num_edges(void* h) {
  int64_t e = 0;
  for (Factor* f : H(h)->fn->factors) e += (int64_t)f->getVariables().size();
  return e;
}
void orc_get_bounds(void* h, double* lb, double* ub, double* samp_lo, double* samp_hi) {
  for (Variable* v : H(h)->fn->variables) {
    lb[v->getID()] = v->getDomain().lo;
    ub[v->getID()] = v->getDomain().hi;
    if (samp_lo) samp_lo[v->getID()] = v->samp_lo;
    if (samp_hi) samp_hi[v->getID()] = v->samp_hi;
  }
}
int orc_get_xinit(void* h, double* x) {
  auto& xi = H(h)->fn->xinit;
  if (xi.empty()) return 0;
  std::memcpy(x, xi.data(), xi.size() * sizeof(double));
  return 1;
}
void orc_export_nlpf(void* h, int64_t* rowptr, int32_t* vid, double* expo, double* konst, uint8_t* sine, double* coeff) {
  int64_t e = 0, j = 0;
  for (Factor* fp : H(h)->fn->factors) {
    auto* f = static_cast<NonlinearProductFactor*>(fp);
    rowptr[j] = e;
    coeff[j] = f->coeff();
    for (size_t i = 0; i < f->getVariables().size(); ++i, ++e) {
      vid[e] = (int32_t)f->getVariables()[i]->getID();
      expo[e] = f->terms()[i].exponent;
      konst[e] = f->terms()[i].constant;
      sine[e] = f->terms()[i].useSine ? 1 : 0;
    }
    ++j;
  }
  rowptr[j] = e;
}
void orc_export_ba(void* h, int32_t* cam, int32_t* pt, double* obs_xy) {
  int64_t j = 0;
  for (Factor* fp : H(h)->fn->factors) {
    auto* f = static_cast<BundleAdjustmentFactor*>(fp);
    cam[j] = (int32_t)f->camera();
    pt[j] = (int32_t)f->point();
    obs_xy[2 * j] = f->obsX();
    obs_xy[2 * j + 1] = f->obsY();
    ++j;
  }
}

void orc_set_change_filter(int on) { Variable::changeFilter() = (on != 0); }

void orc_set_x(void* h, int64_t n, const int32_t* vid, const double* x) {
  auto& vars = H(h)->fn->variables;
  for (int64_t i = 0; i < n; ++i) vars[vid ? vid[i] : i]->assign(x[i]);
}
void orc_get_x(void* h, int64_t n, const int32_t* vid, double* x) {
  auto& vars = H(h)->fn->variables;
  for (int64_t i = 0; i < n; ++i) x[i] = vars[vid ? vid[i] : i]->eval();
}
void orc_unassign(void* h, int64_t n, const int32_t* vid) {
  auto& vars = H(h)->fn->variables;
  for (int64_t i = 0; i < n; ++i) {
    Variable* v = vars[vid ? vid[i] : i];
    if (v->isAssigned()) v->unassign();
  }
}
void orc_set_factor_const(void* h, int64_t n, const int64_t* fid, const double* val, const uint8_t* on) {
  auto& fs = H(h)->fn->factors;
  for (int64_t i = 0; i < n; ++i) fs[fid[i]]->setAssignedConstant(on[i] != 0, val[i]);
}

// evalFactors over a factor-id list (nullptr = all factors, in id order).
double orc_eval(void* h, int64_t nf, const int64_t* fid, double* per_factor, int use_cache) {
  OptimizableFunction& fn = *H(h)->fn;
  std::vector<Factor*> fs;
  if (fid == nullptr) fs = fn.factors;
  else
    for (int64_t i = 0; i < nf; ++i) fs.push_back(fn.factors[fid[i]]);
  if (per_factor) {
    for (size_t i = 0; i < fs.size(); ++i) {
      const Factor& f = *fs[i];
      per_factor[i] = (!f.isAssigned() && !f.areAllVarsAssigned()) ? 0.0
                      : (use_cache ? f.eval(fn.counters) : f.evalNoCache());
    }
  }
  return fn.evalFactors(fs, use_cache != 0);
}

// Factor::computeBounds per listed factor (nullptr = all) and OptimizableFunction::computeBounds' interval sum over the
// list.  point[v] != 0: the variable enters as its current value (it must be assigned), otherwise as its domain hull —
// the caller expresses "assigned and not in vidsToIgnore" through this mask.
void orc_bounds(void* h, const uint8_t* point, int64_t nf, const int64_t* fid, double* lower, double* upper, double* sum2) {
  OptimizableFunction& fn = *H(h)->fn;
  auto is_point = [&](const Variable& v) { return point[v.getID()] != 0; };
  Interval total(0.0);  // semiring Product identity, src/OptimizableFunction.cpp:192
  const int64_t n = fid ? nf : (int64_t)fn.factors.size();
  for (int64_t i = 0; i < n; ++i) {
    Factor* f = fn.factors[fid ? fid[i] : i];
    bool all_points = true;
    for (const Variable* v : f->getVariables()) all_points = all_points && is_point(*v);
    Interval b;
    if (f->isAssigned() || all_points) {
      b = Interval(f->isAssigned() ? f->eval(fn.counters) : f->evalNoCache());  // src/Factor.cpp:128
    } else if (fn.kind == OptimizableFunction::KIND_NLPF) {
      b = nlpf_factor_bounds(*static_cast<NonlinearProductFactor*>(f), is_point);
    } else {
      b = ba_factor_bounds(*static_cast<BundleAdjustmentFactor*>(f), is_point);
    }
    if (lower) lower[i] = b.lower();
    if (upper) upper[i] = b.upper();
    total = total + b;
  }
  if (sum2) {
    sum2[0] = total.lower();
    sum2[1] = total.upper();
  }
}

// computeGradientOfSum restricted to `vid` (SubfunctionFD::df without the assign).
void orc_grad(void* h, int64_t nf, const int64_t* fid, int64_t nv, const int32_t* vid, double* g) {
  OptimizableFunction& fn = *H(h)->fn;
  std::vector<Factor*> fs;
  if (fid == nullptr) fs = fn.factors;
  else
    for (int64_t i = 0; i < nf; ++i) fs.push_back(fn.factors[fid[i]]);
  PartialGradient pg;
  fn.computeGradientOfSum(fs, pg);
  for (int64_t i = 0; i < nv; ++i) {
    const double* d = pg.find(vid ? vid[i] : i);
    g[i] = d ? *d : 0.0;
  }
}

// Per-factor gradient rows in the factor's own slot order (Factor::computeGradient).
void orc_factor_grad(void* h, int64_t fid, double* out_by_slot) {
  Factor* f = H(h)->fn->factors[fid];
  PartialGradient pg;
  f->computeGradient(pg);
  for (size_t i = 0; i < f->getVariables().size(); ++i) {
    const double* d = pg.find(f->getVariables()[i]->getID());
    out_by_slot[i] = d ? *d : 0.0;
  }
}

// One CGDSubspaceOptimizer::optimize call.  x_inout has nv entries.
double orc_solve_cgd(void* h, int64_t nv, const int32_t* vid, int64_t nf, const int64_t* fid, double* x_inout,
                     int maxiters, double ftol, double* delta, int* iters) {
  Handle* hd = H(h);
  OptimizableFunction& fn = *hd->fn;
  std::vector<Variable*> vars;
  std::vector<Factor*> fs;
  for (int64_t i = 0; i < nv; ++i) vars.push_back(fn.variables[vid[i]]);
  for (int64_t i = 0; i < nf; ++i) fs.push_back(fn.factors[fid[i]]);
  std::vector<double> xval(x_inout, x_inout + nv);
  hd->cgd->setParameters(maxiters, ftol);
  // precondition of the caller (src/RDISOptimizer.cpp:1180): vars already hold xval
  for (int64_t i = 0; i < nv; ++i) vars[i]->assign(xval[i]);
  double d = 0;
  const double fret = hd->cgd->optimize(vars, fs, xval, d, false);
  std::memcpy(x_inout, xval.data(), nv * sizeof(double));
  if (delta) *delta = d;
  if (iters) *iters = (int)hd->cgd->lastIters;
  return fret;
}

// A batch of solves described CSR-style.  nthreads<=1: sequential on this handle.
// nthreads>1: each worker thread runs its share on its own replica of the function
// ("N independent single-threaded processes", the reference being non-reentrant).
// x_inout is the concatenation of the problems' variable values.  Returns wall seconds.
double orc_solve_cgd_batch(void* h, int64_t nprobs, const int64_t* var_off, const int32_t* vids,
                           const int64_t* fac_off, const int64_t* fids, double* x_inout, int maxiters,
                           double ftol, double* f_end, double* f_init, int32_t* iters, int nthreads,
                           void** replicas /* nthreads handles incl. h, or null */) {
  auto run_range = [&](void* hh, int64_t lo, int64_t hi) {
    for (int64_t p = lo; p < hi; ++p) {
      double d = 0;
      int it = 0;
      const double fe = orc_solve_cgd(hh, var_off[p + 1] - var_off[p], vids + var_off[p], fac_off[p + 1] - fac_off[p],
                                      fids + fac_off[p], x_inout + var_off[p], maxiters, ftol, &d, &it);
      if (f_end) f_end[p] = fe;
      if (f_init) f_init[p] = H(hh)->cgd->lastInitialFval;  // the value itself (fe - d would round)
      if (iters) iters[p] = it;
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads <= 1 || replicas == nullptr) {
    run_range(h, 0, nprobs);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
      const int64_t lo = nprobs * t / nthreads, hi = nprobs * (t + 1) / nthreads;
      th.emplace_back(run_range, replicas[t], lo, hi);
    }
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// One LMSubspaceOptimizer::optimize call (PARITY UNPINNED, see lm_oracle.hpp).  info8 (nullable) receives
// {||e||^2 at start, ||e||^2 at end, iterations, stop code, #func, #jac, #linear solves, ||J^T e||_inf}.
double orc_solve_lm(void* h, int64_t nv, const int32_t* vid, int64_t nf, const int64_t* fid, double* x_inout, int maxiters,
                    double ftol, double* delta, double* info8) {
  Handle* hd = H(h);
  OptimizableFunction& fn = *hd->fn;
  std::vector<Variable*> vars;
  std::vector<Factor*> fs;
  for (int64_t i = 0; i < nv; ++i) vars.push_back(fn.variables[vid[i]]);
  for (int64_t i = 0; i < nf; ++i) fs.push_back(fn.factors[fid[i]]);
  std::vector<double> xval(x_inout, x_inout + nv);
  hd->lmopt->setParameters(maxiters, ftol);
  for (int64_t i = 0; i < nv; ++i) vars[i]->assign(xval[i]);
  double d = 0;
  const double fret = hd->lmopt->optimize(vars, fs, xval, d, false);
  std::memcpy(x_inout, xval.data(), nv * sizeof(double));
  if (delta) *delta = d;
  if (info8) {
    const lm::Info& I = hd->lmopt->lastInfo;
    const double v[8] = {I.e0, I.e, (double)I.iters, (double)I.stop, (double)I.nfev, (double)I.njev, (double)I.nlss, I.jte_inf};
    std::memcpy(info8, v, sizeof v);
  }
  return fret;
}

double orc_solve_lm_batch(void* h, int64_t nprobs, const int64_t* var_off, const int32_t* vids, const int64_t* fac_off,
                          const int64_t* fids, double* x_inout, int maxiters, double ftol, double* f_end, double* f_init,
                          int32_t* iters, int32_t* stop) {
  auto t0 = std::chrono::steady_clock::now();
  for (int64_t p = 0; p < nprobs; ++p) {
    double d = 0, info[8];
    const double fe = orc_solve_lm(h, var_off[p + 1] - var_off[p], vids + var_off[p], fac_off[p + 1] - fac_off[p],
                                   fids + fac_off[p], x_inout + var_off[p], maxiters, ftol, &d, info);
    if (f_end) f_end[p] = fe;
    if (f_init) f_init[p] = fe - d;
    if (iters) iters[p] = (int32_t)info[2];
    if (stop) stop[p] = (int32_t)info[3];
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void orc_get_counters(void* h, int64_t* out5) {
  const Counters& c = H(h)->fn->counters;
  out5[0] = c.factor_eval_calls;
  out5[1] = c.factor_recomputes;
  out5[2] = c.factor_grad_calls;
  out5[3] = c.f_evals;
  out5[4] = c.df_evals;
}
void orc_reset_counters(void* h) { H(h)->fn->counters = Counters(); }

void orc_trace_enable(void* h, int on) {
  H(h)->fn->trace_on = (on != 0);
  H(h)->fn->trace.clear();
}
int64_t orc_trace_count(void* h) { return (int64_t)H(h)->fn->trace.size(); }
// record i: returns is_df; x (nv values) and out (1 value, or nv gradient entries) are copied out
int orc_trace_get(void* h, int64_t i, double* x, double* out) {
  const auto& t = H(h)->fn->trace[(size_t)i];
  std::memcpy(x, t.x.data(), t.x.size() * sizeof(double));
  std::memcpy(out, t.out.data(), t.out.size() * sizeof(double));
  return t.is_df;
}

const char* orc_variant() {
#ifdef ORACLE_USE_REFERENCE_NRC
  return "reference-minimize_nrc.h";
#else
  return "restated-nr";
#endif
}

}  // extern "C"
#pragma GCC visibility pop
