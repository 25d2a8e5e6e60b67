// rdis_host.h — host side of the drop-in boundary, C++ (std-only).
//
// The reference is C++11 on Boost; Boost is not available in this image, so the reference's own
// headers cannot be compiled against.  This header mirrors the part of the reference's plugin surface
// that the subspace-solve path touches — same class and member names, argument meaning and error
// behaviour — so that CudaSubspaceOptimizer drops into the `SubspaceOptimizer &` slot the tree
// search holds (src/RDISOptimizer.h:86-89,217; src/optimizers/BCDOptimizer.h:23-24):
//
//   rdis::Numeric / VariableID / FactorID / NumericVec      src/common.h:25-37
//   rdis::VariableDomain (single interval) ::closestVal     src/VariableDomain.cpp:130-163
//   rdis::Variable  assign / unassign / eval / isAssigned   src/Variable.h:23-112, src/Variable.cpp:66-102
//   rdis::Factor    getID / getVariables / containsVar / areAllVarsAssigned / isAssigned /
//                   getAssignedKey / assign / unassign      src/Factor.h:116-213, src/Factor.cpp:110-188,244-341
//   rdis::NonlinearProductFactor (coefficient, per-edge exponent / constant / useSine)
//                                                           src/NonlinearProductFactor.h:27-55,100-113
//   rdis::BundleAdjustmentFactor (camera, point, pixel)     src/bundleadjust/BundleAdjustmentFactor.h:129-136
//   rdis::OptimizableFunction  init / eval / evalFactors / computeGradient / onVarAssigned
//                                                           src/OptimizableFunction.h:33-110, .cpp:79-135,234-262
//   rdis::SubspaceOptimizer  setParameters / optimize / quickAssignVals
//                                                           src/SubspaceOptimizer.h:24-58, .cpp:12-53
//   rdis::CudaSubspaceOptimizer : SubspaceOptimizer         replaces src/optimizers/CGDSubspaceOptimizer.{h,cpp}
//   rdis::ComponentBatcher                                  Component::createChildren src/Component.cpp:508-549,
//                                                           Component::init :50-80, ComponentComparator :603-608,
//                                                           gdfs builder src/RDISOptimizer.cpp:1049-1059,
//                                                           sibling loops :184-211, :291-314
//
// Factor ARITHMETIC is not on the host: Factor objects are descriptors that are flattened once
// (OptimizableFunction::init) into the HBM-resident CSR behind include/rdis_gpu.h, and every evaluation
// goes through that C-ABI.  There is no CPU fallback — without a usable sm_100 device init() throws.
#ifndef RDIS_HOST_H_
#define RDIS_HOST_H_

#include <cstdint>
#include <deque>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

struct rdisgpu_ctx;
struct rdisgpu_batch;

namespace rdis {

typedef double Numeric;
typedef long long int VariableCount;
typedef long long int VariableID;
typedef long long int FactorID;
typedef std::vector<Numeric> NumericVec;
struct NumericInterval {  // boost::numeric::interval<Numeric> as the callers use it: lower() / upper()
  Numeric lo, hi;
  Numeric lower() const { return lo; }
  Numeric upper() const { return hi; }
};
typedef std::vector<VariableID> VariableIDVec;
typedef std::vector<FactorID> FactorIDVec;

class Factor;
class Variable;
class OptimizableFunction;
struct ComponentProblem;

// A read-only view of a run of a flat pool, with the part of std::vector's interface the callers use.  The host objects
// keep NO per-object containers: a factor's variables / terms and a variable's incident factors are runs of pools owned
// by the function (so building the 4.2 M factors of BASELINE config 4 allocates a handful of arrays, not 10 M vectors).
template <class T>
struct Span {
  const T* b;
  size_t n;
  const T* begin() const { return b; }
  const T* end() const { return b + n; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  const T& operator[](size_t i) const { return b[i]; }
};

// One closed interval (every configuration of the hot path has exactly one subinterval per variable;
// CGDSubspaceOptimizer asserts it, src/optimizers/CGDSubspaceOptimizer.cpp:118-119).
class VariableDomain {
 public:
  VariableDomain() : lo_(0), hi_(0), slo_(0), shi_(0) {}
  VariableDomain(Numeric lo, Numeric hi) : lo_(lo), hi_(hi), slo_(lo), shi_(hi) {}
  Numeric min() const { return lo_; }
  Numeric max() const { return hi_; }
  void setSamplingInterval(Numeric lo, Numeric hi) { slo_ = lo; shi_ = hi; }
  Numeric samplingMin() const { return slo_; }
  Numeric samplingMax() const { return shi_; }
  Numeric closestVal(const Numeric val) const {  // src/VariableDomain.cpp:157-163
    if (lo_ <= val && val <= hi_) return val;
    if (val < lo_) return lo_;
    return hi_;
  }

 private:
  Numeric lo_, hi_, slo_, shi_;
};

class Variable {
 public:
  Variable(VariableID id, const VariableDomain& dom) : m_id(id), m_domain(dom), m_isAssigned(false), m_value(0), m_owner(nullptr) {}
  void assign(Numeric newval, bool notifyFactors = true);  // src/Variable.cpp:66-88
  void unassign();                                          // src/Variable.cpp:90-102
  Numeric eval() const {
    if (!m_isAssigned) throw std::logic_error("Variable::eval on an unassigned variable");
    return m_value;
  }
  const VariableID& getID() const { return m_id; }
  const VariableDomain& getDomain() const { return m_domain; }
  void setDomain(const VariableDomain& d) { m_domain = d; }
  Span<Factor*> getFactors() const;  // the factors the variable appears in, ascending factor id (incidence built on first use)
  bool isAssigned() const { return m_isAssigned; }

 private:
  friend class OptimizableFunction;
  friend class CudaSubspaceOptimizer;
  void setFromDevice(Numeric v) {  // the device already holds v: no upload is queued
    m_isAssigned = true;
    m_value = v;
  }
  VariableID m_id;
  VariableDomain m_domain;
  bool m_isAssigned;
  Numeric m_value;
  OptimizableFunction* m_owner;
};
typedef std::vector<Variable*> VariablePtrVec;

class Factor {
 public:
  enum Kind { NONLINEAR_PRODUCT = 0, BUNDLE_ADJUSTMENT = 1 };
  explicit Factor(FactorID id_)
      : id(id_), voff(0), nvars(0), numVarsAssigned(0), isAssignedConstant(false), vidAssigned(-1), assignedVal(0), m_owner(nullptr) {}
  virtual ~Factor() {}
  virtual Kind kind() const = 0;
  // appends to the function's variable pool: only the most recently created factor can grow (every builder works so)
  virtual void addVariable(Variable* vp);
  FactorID getID() const { return id; }
  Span<Variable*> getVariables() const;
  size_t numVars() const { return nvars; }
  bool isAssigned() const { return isAssignedConstant; }
  VariableID getAssignedKey() const { return vidAssigned; }
  bool areAllVarsAssigned() const { return numVarsAssigned == (VariableCount)nvars; }
  bool isVarInFactor(VariableID vid) const {
    for (const Variable* v : getVariables())
      if (v->getID() == vid) return true;
    return false;
  }
  virtual bool containsVar(VariableID vid) const { return !isAssignedConstant && isVarInFactor(vid); }
  // Simplification to a constant by the tree search (src/Factor.cpp:287-341): eval() then returns fval
  // while the analytic gradient is unchanged; mirrored to the device overlay (rdisgpu_set_factor_const).
  virtual void assign(Numeric fval, VariableID assignmentKey);
  virtual void unassign(VariableID assignmentKey);
  // cache invalidation hooks of the reference (src/Factor.cpp:154-181): the device keeps no per-factor
  // cache, so only the assigned-variable count is maintained
  virtual void onVarAssigned(VariableID, Numeric) { ++numVarsAssigned; }
  virtual void onVarChanged(VariableID, Numeric, Numeric) {}
  virtual void onVarUnassigned(VariableID, Numeric) { --numVarsAssigned; }

 protected:
  friend class OptimizableFunction;
  FactorID id;
  size_t voff;      // first slot of this factor's run in the function's variable (and term) pool
  uint32_t nvars;
  VariableCount numVarsAssigned;
  bool isAssignedConstant;
  VariableID vidAssigned;
  Numeric assignedVal;
  OptimizableFunction* m_owner;
};
typedef std::vector<Factor*> FactorPtrVec;

// c * prod_i [sin]((x_i - k_i)^{e_i})   (src/NonlinearProductFactor.cpp:186-209)
class NonlinearProductFactor : public Factor {
 public:
  struct Term {
    Numeric exponent, constant;
    bool useSine;
  };
  NonlinearProductFactor(FactorID id_, Numeric coeff) : Factor(id_), coefficient(coeff) {}
  Kind kind() const override { return NONLINEAR_PRODUCT; }
  void addVariable(Variable* vp) override { addVariable(vp, 1.0, 0.0, false); }
  void addVariable(Variable* vp, Numeric exponent, Numeric constant, bool useSine);
  Numeric getCoefficient() const { return coefficient; }
  Span<Term> getTerms() const;

 private:
  Numeric coefficient;
};

// one reprojection observation: 9 camera + 3 point variables, enum order of
// src/bundleadjust/BundleAdjustmentCommon.h:36-59
class BundleAdjustmentFactor : public Factor {
 public:
  BundleAdjustmentFactor(FactorID id_, int32_t cam, int32_t pt, Numeric obsx, Numeric obsy)
      : Factor(id_), camera(cam), point(pt), ox(obsx), oy(obsy) {}
  Kind kind() const override { return BUNDLE_ADJUSTMENT; }
  int32_t getCamera() const { return camera; }
  int32_t getPoint() const { return point; }
  Numeric obsX() const { return ox; }
  Numeric obsY() const { return oy; }

 private:
  int32_t camera, point;
  Numeric ox, oy;
};

// Stand-in for boost::program_options::variables_map (src/SubspaceOptimizer.cpp:22-35 reads
// "SSmaxit" and "SSftol" from it).
struct ParameterMap : public std::map<std::string, Numeric> {
  size_t count(const std::string& k) const { return find(k) == end() ? 0 : 1; }
};

class OptimizableFunction {
 public:
  OptimizableFunction();
  virtual ~OptimizableFunction();  // deletes variables and factors (src/OptimizableFunction.cpp:43-54)
  OptimizableFunction(const OptimizableFunction&) = delete;
  OptimizableFunction& operator=(const OptimizableFunction&) = delete;

  // ---- construction (what the loaders / generators of the reference do) ----
  Variable* addVariable(Numeric lb, Numeric ub);
  NonlinearProductFactor* addProductFactor(Numeric coefficient);
  // bundle adjustment: variables must have been created as ncams*9 camera then npts*3 point
  // variables (BundleAdjustmentFunction.h:88-96)
  void declareBundleAdjustment(int32_t ncams, int32_t npts);
  BundleAdjustmentFactor* addObservation(int32_t cam, int32_t pt, Numeric obsx, Numeric obsy);

  // Flattens the object graph into the device-resident layout (OptimizableFunction::init,
  // src/OptimizableFunction.cpp:57-76).  Throws std::runtime_error when no sm_100 device is usable.
  virtual void init(int device = 0);

  // capacity hints for the builders (optional): avoids pool re-allocation while millions of factors are added
  void reserve(size_t nFactors, size_t nEdges);
  // the flat arrays rdisgpu_add_nlpf takes (NonlinearProductFactor functions): what init() uploads
  void exportProductFactors(std::vector<int64_t>& rowptr, std::vector<int32_t>& vid, std::vector<double>& expo, std::vector<double>& konst,
                            std::vector<uint8_t>& sine, std::vector<double>& coeff) const;
  VariableCount getNumVars() const { return (VariableCount)variables.size(); }
  VariablePtrVec& getVariables() { return variables; }
  FactorPtrVec const& getFactors() const { return factors; }

  // ---- evaluation: all on the device ----
  virtual Numeric eval();                                                     // src/OptimizableFunction.cpp:89-92
  virtual Numeric evalFactors(const FactorPtrVec& fctrs, Numeric& ferr);      // :95-135 (ferr: simplification error, 0 here)
  // gradient of sum_{f in facs} f restricted to `vars` (computeGradient + the scatter of
  // SubfunctionFD::df, src/OptimizableFunction.cpp:234-262, CGDSubspaceOptimizer.cpp:135-157)
  virtual void computeGradient(const FactorPtrVec& facs, const VariablePtrVec& vars, NumericVec& gradient);
  // Interval bounds of the sum of `fctrs` for branch & bound (src/OptimizableFunction.cpp:181-216): factors that are
  // unassigned, or assigned under the key `assignedVID` (any factor when assignedVID < 0), each bounded by
  // Factor::computeBounds (assigned variables as points, the rest as their domain hull) — through rdisgpu_bounds.
  virtual NumericInterval computeBounds(const FactorPtrVec& fctrs, VariableID assignedVID = -1);
  NumericInterval computeBounds() { return computeBounds(factors, -1); }
  // The same for MANY factor lists in one device call (every child of a decomposition: Component::computeBounds,
  // src/Component.cpp:240-241,592-599), each list's interval sum folded on the device in list order (rdisgpu_bounds_lists).
  virtual void computeBoundsBatch(const std::vector<FactorPtrVec>& lists, std::vector<NumericInterval>& out);

  // hook the reference calls after every Variable::assign made by a subspace optimizer (no-op there,
  // src/OptimizableFunction.h:65-69); here the host->device mirroring is driven by Variable::assign itself
  virtual void onVarAssigned(const VariableID, const Numeric) {}
  virtual void onVarUnassigned(const VariableID) {}

  // ---- device mirror ----
  rdisgpu_ctx* device() const { return ctx; }
  void flushAssignments();  // uploads the values of variables assigned / changed since the last flush
  bool isMinSum() const { return true; }  // the only semiring RDISOptimizer accepts (src/RDISOptimizer.cpp:57-61)

 private:
  friend class Variable;
  friend class Factor;
  friend class CudaSubspaceOptimizer;
  void noteAssigned(VariableID vid);
  void noteFactorConst(Factor* f);
  void check(int rc, const char* what) const;

  friend class NonlinearProductFactor;
  void ensureIncidence() const;  // variable -> factors CSR, built once after construction (counting sort, ascending factor id)
  VariablePtrVec variables;
  FactorPtrVec factors;
  // arenas and pools: objects are constructed in chunked arenas (stable addresses, no per-object allocation); a factor's
  // variables / terms are the run [voff, voff + nvars) of varPool / termPool
  std::deque<Variable> variableArena;
  std::deque<NonlinearProductFactor> productArena;
  std::deque<BundleAdjustmentFactor> observationArena;
  std::vector<Variable*> varPool;
  std::vector<NonlinearProductFactor::Term> termPool;
  mutable std::vector<size_t> incOff;       // V + 1
  mutable std::vector<Factor*> incList;     // one entry per (variable, factor) pair
  mutable bool incidenceValid = false;
  int kind;  // -1 none, Factor::Kind otherwise
  int32_t ncams, npts;
  rdisgpu_ctx* ctx;
  std::vector<int32_t> dirtyVids;
  std::vector<uint8_t> dirtyFlag;
  std::vector<Factor*> dirtyFactors;
  // Sibling batches the tree search comes back to (alternating minimisation re-poses the same components with new start
  // values, src/RDISOptimizer.cpp:1148-1181): their index lists stay resident on the device (rdisgpu_batch), keyed by the
  // lists themselves; a revisit uploads start values only.  Owned here so that they die before the context.
  struct CachedBatch {
    unsigned long long key = 0, stamp = 0;
    std::vector<int32_t> vids;
    std::vector<int64_t> fids, var_off, fac_off;
    // the objects the lists were packed from, problem after problem: a call whose problems hold the SAME pointers in the
    // same order is recognised by comparing pointer runs (sequential memory), without chasing 4e5 pointers for their ids
    std::vector<const Variable*> pvars;
    std::vector<const Factor*> pfacs;
    rdisgpu_batch* batch = nullptr;
  };
  std::vector<CachedBatch> batchCache;
  unsigned long long batchClock = 0;
  CachedBatch* cachedBatch(const std::vector<int64_t>& var_off, const std::vector<int32_t>& vids, const std::vector<int64_t>& fac_off,
                           const std::vector<int64_t>& fids);
  CachedBatch* findWave(const std::vector<ComponentProblem>& problems);
  void rememberWave(CachedBatch& c, const std::vector<ComponentProblem>& problems);
};

class SubspaceOptimizer {
 public:
  explicit SubspaceOptimizer(OptimizableFunction& f_);
  virtual ~SubspaceOptimizer() {}
  virtual void setParameters(const ParameterMap& options);
  // vars / factors define the subfunction; xval holds the start values in the order of `vars` and the
  // final (domain-clamped) values on return; deltaFval = f(x_end) - f(x_init); returns f(x_end).
  virtual Numeric optimize(const VariablePtrVec& vars, const FactorPtrVec& factors, NumericVec& xval, Numeric& deltaFval,
                           const bool printdbg) = 0;

 protected:
  void quickAssignVals(const VariablePtrVec& vars, const NumericVec& xval, bool sanitizeVals);
  OptimizableFunction& f;
  const bool doAscent;
  size_t maxiters;
  Numeric ftol;
};

// One subspace problem of a sibling batch.
struct ComponentProblem {
  VariablePtrVec vars;
  FactorPtrVec factors;
  NumericVec xval;        // in: start values; out: final values
  Numeric fval = 0;       // out: f(x_end) (return value of optimize)
  Numeric deltaFval = 0;  // out
  int iters = 0, status = 0;
};

class CudaSubspaceOptimizer : public SubspaceOptimizer {
 public:
  explicit CudaSubspaceOptimizer(OptimizableFunction& f_) : SubspaceOptimizer(f_), useLM(false) {}
  Numeric optimize(const VariablePtrVec& vars, const FactorPtrVec& factors, NumericVec& xval, Numeric& deltaFval,
                   const bool printdbg) override;
  // The sibling-component batch (src/RDISOptimizer.cpp:184-211, 291-314 loop over children one at a
  // time; here the whole wave is ONE device call).  Problems must not share variables or factors.
  // Returns the sum of the problems' final objective values.
  Numeric optimizeBatch(std::vector<ComponentProblem>& problems, const bool printdbg);
  // wall-clock milliseconds spent in optimizeBatch since the last resetTiming(): recognising / packing the wave, the
  // device calls (upload + solve + download, synchronous), the write-back into the host objects
  struct Timing { double pack_ms = 0, flush_ms = 0, device_ms = 0, fetch_ms = 0, writeback_ms = 0; long calls = 0; };
  const Timing& timing() const { return tm; }
  void resetTiming() { tm = Timing(); }

 protected:
  CudaSubspaceOptimizer(OptimizableFunction& f_, bool lm) : SubspaceOptimizer(f_), useLM(lm) {}
  const bool useLM;  // false: conjugate gradient (rdisgpu_solve_cgd_csr); true: Levenberg-Marquardt (rdisgpu_solve_lm_csr)

 private:
  std::vector<int64_t> var_off, fac_off, fids, nfe, nge;
  std::vector<int32_t> vids, iters, status;
  std::vector<double> x0, xout, finit, fend;
  Timing tm;
};

// Replaces src/optimizers/LMSubspaceOptimizer.{h,cpp} (selected by --useCGD 0, src/bundleadjust/optBA.cpp:155-158).
// Same interface; the solve is levmar's dlevmar_der on the device: no clamping, no revert, stop codes 1..7 in
// ComponentProblem::status.  PARITY UNPINNED (levmar is not vendored with the reference).
class CudaLMSubspaceOptimizer : public CudaSubspaceOptimizer {
 public:
  explicit CudaLMSubspaceOptimizer(OptimizableFunction& f_) : CudaSubspaceOptimizer(f_, true) {}
};

// Sibling components of a set of variables: connected components of the bipartite variable / factor
// graph in which an edge (v, f) exists iff v is unassigned and f is not assigned to a constant
// (src/ConnectivityGraph.cpp:211-285).  Membership is what the reference's Euler-tour connectivity
// yields: variables and factors of a child sorted by id (Component::init, src/Component.cpp:50-80),
// children ordered by number of variables, smallest first (ComponentComparator, :603-608); ties are
// broken by smallest variable id (the reference's tie order is an artefact of its splay trees).
struct ChildComponent {
  VariableIDVec vars;
  FactorIDVec factors;
};
class ComponentBatcher {
 public:
  static void createChildren(const OptimizableFunction& func, const VariableIDVec& componentVars,
                             std::vector<ChildComponent>& children);
  // Same children, labelled on the GPU (rdisgpu_components: min-label propagation over the device CSR) — for
  // million-variable graphs, where the host union-find over heap objects is the slow part.  Needs init().
  static void createChildrenOnDevice(OptimizableFunction& func, const VariableIDVec& componentVars,
                                     std::vector<ChildComponent>& children);
  // The reference's sibling loop WITH branch & bound (src/RDISOptimizer.cpp:291-314) over ONE batched device call.
  // Sequentially the reference (a) gives child k the budget fmin_k = childFmin + uab_k.lower (Component::computeFMin,
  // src/Component.cpp:389-395) and skips its optimisation when fmin_k < uab_k.lower (checkUnassignedBound,
  // RDISOptimizer.cpp:894-913: the child is set to its lower bound), (b) after every child replaces the child's lower
  // bound by its value in childFmin and in the parent's assigned bound (Component::onChildEvaluated,
  // src/Component.cpp:252-343), and (c) leaves the loop as soon as the parent's fmin <= its assigned lower bound
  // (checkAssignedBound, RDISOptimizer.cpp:916-934): later siblings are never optimised.  All three decisions depend on
  // the VALUES of earlier siblings, so a wave cannot know in advance which children the reference would have solved.
  // Siblings share no variable and no factor, so each child's solve is independent of the others: the wave is solved
  // speculatively in one batch, the reference's decisions are then REPLAYED in sibling order on the results, and the
  // children the reference would not have optimised are rolled back (host and device state restored).  The outcome is
  // identical to the sequential loop; the price is the wasted speculative solves.
  enum SiblingOutcome { SIB_OPTIMISED = 0, SIB_BOUND_SKIPPED = 1, SIB_PRUNED = 2 };
  struct SiblingWave {
    std::vector<int> outcome;        // per child, SiblingOutcome
    std::vector<Numeric> value;      // per child: f(x_end), the lower bound (skipped), or NaN (pruned: never evaluated)
    Numeric childFmin = 0;           // the parent's estimator after the loop
    Numeric assignedLower = 0;       // the parent's assigned lower bound after the loop
    size_t evaluated = 0;            // children the loop visited (optimised or skipped)
  };
  // children: in the reference's sibling order; uab: their unassigned bounds; parentFmin / childFmin0: Component::fmin and
  // the childFmin decompose() computed (fmin - partialEval - sum of lower bounds); useBounds = RDISOptimizer::useBounds.
  static void optimizeSiblings(CudaSubspaceOptimizer& ssopt, OptimizableFunction& func, std::vector<ComponentProblem>& children,
                               const std::vector<NumericInterval>& uab, Numeric parentFmin, Numeric childFmin0, bool useBounds,
                               SiblingWave& out);
  // The subspace problem of optimising ALL variables of a child (a leaf visit):
  // gdfs = the child's factors whose other variables are all assigned (src/RDISOptimizer.cpp:1049-1059),
  // start values = current values of the variables that are assigned, else `fallback[vid]`.
  static void leafProblem(OptimizableFunction& func, const ChildComponent& child, const NumericVec& fallback,
                          ComponentProblem& out);
};

}  // namespace rdis
#endif  // RDIS_HOST_H_
