// rdis_builders.cpp — see rdis_builders.h.
#include "rdis_builders.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

namespace rdis {

namespace {
enum BaSlot { ROT_X = 0, ROT_Y, ROT_Z, TRANS_X, TRANS_Y, TRANS_Z, FOCAL, RDL_K1, RDL_K2, POINT_X, POINT_Y, POINT_Z };
}

bool BundleAdjustmentFunction::load(const std::string& file, VariableCount numcams, VariableCount numpoints) {
  std::ifstream ifs(file.c_str());
  if (!ifs.is_open()) {
    std::cerr << "BundleAdjustmentFunction::Load: Failed to open file for reading " << file << std::endl;
    return false;
  }
  long long numCameras = 0, numPoints = 0;
  ifs >> numCameras >> numPoints >> m_numObservations;
  if (!ifs) return false;
  m_numCameras = (numcams <= 0 ? numCameras : numcams);
  m_numPoints = (numpoints <= 0 ? numPoints : numpoints);
  if (numCameras < m_numCameras || numPoints < m_numPoints) return false;

  // variables first (domains are set once the initial values have been read, like the reference's "0:0")
  for (VariableCount i = 0; i < 9 * m_numCameras + 3 * m_numPoints; ++i) addVariable(0.0, 0.0);
  declareBundleAdjustment((int32_t)m_numCameras, (int32_t)m_numPoints);

  for (long long i = 0; i < m_numObservations; ++i) {
    long long camID = 0, pointID = 0;
    Numeric obsX = 0, obsY = 0;
    ifs >> camID >> pointID >> obsX >> obsY;
    if (!ifs) return false;
    if (camID >= m_numCameras || pointID >= m_numPoints) continue;
    addObservation((int32_t)camID, (int32_t)pointID, obsX, obsY);
  }

  xinit.assign((size_t)getNumVars(), 0.0);
  for (long long cid = 0; cid < numCameras; ++cid) {
    Numeric cameraParams[9];
    for (int p = 0; p < 9; ++p) ifs >> cameraParams[p];
    if (cid >= m_numCameras) continue;
    for (int p = 0; p < 9; ++p) {
      const VariableID vid = getCamVID(cid, p);
      xinit[(size_t)vid] = cameraParams[p];
      setDomain(vid, cameraParams[p]);
    }
  }
  for (long long pid = 0; pid < numPoints; ++pid) {
    Numeric point[3];
    for (int d = 0; d < 3; ++d) ifs >> point[d];
    if (pid >= m_numPoints) continue;
    for (int d = 0; d < 3; ++d) {
      const VariableID vid = getPointVID(pid, d);
      xinit[(size_t)vid] = point[d];
      setDomain(vid, point[d]);
    }
  }
  if (!ifs) return false;
  return getNumVars() > 0 && !getFactors().empty();
}

void BundleAdjustmentFunction::setDomain(VariableID vid, Numeric initialVal) {
  const int type = (vid < 9 * m_numCameras) ? (int)(vid % 9) : (int)(POINT_X + (vid - 9 * m_numCameras) % 3);
  const Numeric dsf = 1000.0;
  const Numeric pi = 3.141592653589793238462643383279502884;
  Numeric slo, shi, dlo, dhi;
  // interval * scalar without rounding: hull of the two products
  auto times = [](Numeric lo, Numeric hi, Numeric k, Numeric& olo, Numeric& ohi) {
    const Numeric a = lo * k, b = hi * k;
    olo = std::min(a, b);
    ohi = std::max(a, b);
  };
  switch (type) {
    case ROT_X: case ROT_Y: case ROT_Z:
      slo = -1 * pi; shi = 1 * pi;
      times(slo, shi, dsf, dlo, dhi);
      break;
    case TRANS_X: case TRANS_Y: case TRANS_Z: case POINT_X: case POINT_Y: case POINT_Z:
      slo = initialVal + -1 * 1e2; shi = initialVal + 1 * 1e2;
      times(slo, shi, dsf, dlo, dhi);
      break;
    case FOCAL: {
      slo = initialVal + -1 * 1e2; shi = initialVal + 1 * 1e2;
      Numeric lower = std::min(slo, slo * dsf);
      lower = std::max(lower, 0.0);
      dlo = lower; dhi = shi * dsf;
      break;
    }
    case RDL_K1:
      slo = initialVal + -1 * 1e-4; shi = initialVal + 1 * 1e-4;
      dlo = -1e-1; dhi = 1e-1;
      break;
    default:  // RDL_K2
      slo = initialVal + -1 * 1e-6; shi = initialVal + 1 * 1e-6;
      dlo = -1e-3; dhi = 1e-3;
      break;
  }
  VariableDomain dom(std::min(dlo, slo), std::max(dhi, shi));  // hull(dom, sit), :468
  dom.setSamplingInterval(slo, shi);
  getVariables()[(size_t)vid]->setDomain(dom);
}

void BundleAdjustmentFunction::getBlockRangeByBlkId(VariableCount blockid, VariableID& lo, VariableID& hi) const {
  if (blockid < m_numCameras) {
    lo = 9 * blockid;
    hi = lo + 8;
  } else {
    lo = 9 * m_numCameras + 3 * (blockid - m_numCameras);
    hi = lo + 2;
  }
}

VariableCount BundleAdjustmentFunction::getBlockID(VariableID vid) const {
  return (vid < 9 * m_numCameras) ? vid / 9 : m_numCameras + (vid - 9 * m_numCameras) / 3;
}

OptimizableFunction* makeHighDimSinusoid(VariableCount treeHeight, VariableCount branches, VariableCount maxArity,
                                         bool allowOddArityFactors) {
  const Numeric twopi = 2.000001 * 3.141592653;
  // the reference builds the domain from the string "-%1%:%1%" % (10 * twopi): 6 significant digits
  char buf[64];
  std::snprintf(buf, sizeof buf, "%g", 10 * twopi);
  const Numeric bound = std::strtod(buf, nullptr);
  OptimizableFunction* poly = new OptimizableFunction();
  maxArity = std::min(maxArity, treeHeight + 1);
  const VariableCount h = treeHeight, k = branches;
  VariableCount nvars = h + 1;
  if (k != 1) {
    VariableCount pw = 1;
    for (VariableCount i = 0; i < h + 1; ++i) pw *= k;  // k^(h+1), exact in integers
    nvars = (pw - 1) / (k - 1);
  }
  for (VariableCount v = 0; v < nvars; ++v) {
    Variable* var = poly->addVariable(-bound, bound);
    VariableDomain d(-bound, bound);
    d.setSamplingInterval(-twopi, twopi);
    var->setDomain(d);
  }
  // depth of a vertex of the complete k-ary tree in BFS numbering
  auto firstAtDepth = [&](VariableCount d) -> VariableID {
    if (k == 1) return d;
    VariableCount pw = 1;
    for (VariableCount i = 0; i < d; ++i) pw *= k;
    return (pw - 1) / (k - 1);
  };
  {  // capacity of the pools: factors and edges per arity class (a vertex at depth d has factors of arity <= d + 1)
    size_t nF = (size_t)nvars, nE = (size_t)nvars;
    for (VariableCount ar = 1; ar <= maxArity; ++ar) {
      if (ar > 1 && (ar & 1) && !allowOddArityFactors) continue;
      const size_t cnt = (size_t)(nvars - firstAtDepth(ar - 1));
      nF += cnt;
      nE += cnt * (size_t)ar;
    }
    poly->reserve(nF, nE);
  }
  VariablePtrVec& variables = poly->getVariables();
  std::vector<Variable*> chain;
  for (VariableCount ar = 1; ar <= maxArity; ++ar) {
    if (ar > 1 && (ar & 1) && !allowOddArityFactors) continue;
    VariableCount depth = h;
    for (VariableID vid = nvars - 1; vid >= 0; --vid) {
      while (vid < firstAtDepth(depth)) --depth;  // depth of vid
      if (depth + 1 < ar) continue;               // needs ar-1 ancestors
      chain.clear();
      VariableID cur = vid;
      for (VariableCount c = 0; c < ar; ++c) {     // the variable and its ancestors, leaf first
        chain.push_back(variables[(size_t)cur]);
        cur = (cur >= 1) ? (cur - 1) / k : -1;  // parent = floor((cur - 1) / k)
      }
      NonlinearProductFactor* f = poly->addProductFactor(ar > 1 ? 12 : 0.6);
      for (size_t i = chain.size(); i-- > 0;) f->addVariable(chain[i], 1, 0, ar > 1);  // root-most first
    }
  }
  for (VariableID vid = 0; vid < nvars; ++vid) poly->addProductFactor(0.1)->addVariable(variables[(size_t)vid], 2, 0, false);
  return poly;
}

}  // namespace rdis
