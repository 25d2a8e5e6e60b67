// rdis_builders.h — the reference's problem builders for the host layer: they construct the
// reference-shaped object graph (rdis_host.h) that OptimizableFunction::init() flattens to the device.
//
//   rdis::BundleAdjustmentFunction::load / setDomain / getCamVID / getPointVID / blocks
//        src/bundleadjust/BundleAdjustmentFunction.cpp:50-250, 402-477, BundleAdjustmentFunction.h:51-96
//   rdis::makeHighDimSinusoid
//        src/OptimizableFunctionGenerator.cpp:660-760
// Host-only code (text parsing, integer bookkeeping); no factor is evaluated here.
#ifndef RDIS_BUILDERS_H_
#define RDIS_BUILDERS_H_

#include <string>

#include "rdis_host.h"

namespace rdis {

class BundleAdjustmentFunction : public OptimizableFunction {
 public:
  BundleAdjustmentFunction() : m_numCameras(0), m_numPoints(0), m_numObservations(0) {}
  // BAL text format ("Bundle Adjustment in the Large"): header `ncams npts nobs`, nobs lines
  // `cam pt x y`, then 9 parameters per camera, then 3 coordinates per point.  numcams / numpoints <= 0
  // mean "all"; observations of dropped cameras / points are skipped (:167-169).  Unlike the reference,
  // load() does not call init(): the caller chooses the device with init(device).
  bool load(const std::string& file, VariableCount numcams = -1, VariableCount numpoints = -1);
  VariableID getCamVID(VariableCount cid, int param) const { return 9 * cid + param; }                      // .h:88-91
  VariableID getPointVID(VariableCount pid, int coord) const { return 9 * m_numCameras + 3 * pid + coord; } // .h:93-96
  VariableCount getNumCameras() const { return m_numCameras; }
  VariableCount getNumPoints() const { return m_numPoints; }
  VariableCount getNumObservationsInFile() const { return m_numObservations; }
  const NumericVec& getInitialState() const { return xinit; }
  // variable blocks: one per camera (9 variables), then one per point (3 variables) (.h:51-78)
  VariableCount getNumBlocks() const { return m_numCameras + m_numPoints; }
  void getBlockRangeByBlkId(VariableCount blockid, VariableID& lo, VariableID& hi) const;
  VariableCount getBlockID(VariableID vid) const;

 private:
  void setDomain(VariableID vid, Numeric initialVal);  // :402-477
  VariableCount m_numCameras, m_numPoints, m_numObservations;
  NumericVec xinit;
};

// Sinusoid "tree" test function: complete `branches`-ary tree of height `treeHeight`; for every arity
// a in 1..maxArity (odd a > 1 only if allowed) and every variable with at least a-1 ancestors one factor
// over the variable and its a-1 ancestors (root-most first), coefficient 12 with sin terms (a > 1) or 0.6 with
// plain terms (a = 1); plus 0.1 * x^2 per variable; domains +-10 * 2.000001 * pi printed with 6 significant
// digits, sampling interval +-2.000001 * pi.  Caller owns the result.
OptimizableFunction* makeHighDimSinusoid(VariableCount treeHeight, VariableCount branches, VariableCount maxArity,
                                         bool allowOddArityFactors);

}  // namespace rdis
#endif  // RDIS_BUILDERS_H_
