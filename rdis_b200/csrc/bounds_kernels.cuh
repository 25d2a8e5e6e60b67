// bounds_kernels.cuh — interval bounds of factors for branch & bound:
//   Factor::computeBounds                               src/Factor.cpp:122-139
//   NonlinearProductFactor::computeFactorBounds         src/NonlinearProductFactor.cpp:120-145, power(): src/util/numeric.cpp:26-43
//   BundleAdjustmentFactor::computeFactorBounds         src/bundleadjust/BundleAdjustmentFactor.cpp:47-52, getVarVals :67-93,
//     evalFactor(IntervalVec) :104-157, angleAxisRotatePoint(IntervalVec) :186-232, simpleNormalize BundleAdjustmentCommon.cpp:43-53
// One thread per factor.  A variable enters as the point [x, x] when the caller marks it assigned and as its domain
// hull [lb, ub] otherwise (VariableDomain::interval); a factor whose variables are all assigned, or that is an
// assigned constant, returns its point value (Factor.cpp:128).  Operation order follows the cited lines.
#pragma once
#include "factors.cuh"
#include "interval.cuh"

namespace rdisgpu {

__device__ __forceinline__ Ival bounds_var(const GraphView& G, const uint8_t* assigned, int32_t vid, bool& all_assigned) {
  if (assigned[vid]) {
    const double x = G.xbd[vid].x;
    return iv_point(x);
  }
  all_assigned = false;
  const double2 d = __ldg(&G.dom[vid]);
  return iv(d.x, d.y);
}

__device__ __forceinline__ Ival nlpf_factor_bounds(const GraphView& G, const uint8_t* assigned, int64_t f, bool& all_assigned) {
  Ival feval = iv_point(1.0);
  const int32_t e0 = __ldg(&G.rowptr[f]), e1 = __ldg(&G.rowptr[f + 1]);
  for (int32_t e = e0; e < e1; ++e) {
    Ival val = bounds_var(G, assigned, __ldg(&G.evid[e]), all_assigned);
    const double k = __ldg(&G.konst[e]), ex = __ldg(&G.expo[e]);
    if (k != 0) val = iv_sub(val, k);
    if (ex != 1) val = iv_rdis_power(val, ex);
    if (__ldg(&G.sine[e])) val = iv_sin(val);
    feval = iv_mul(feval, val);
  }
  return iv_mul(feval, __ldg(&G.coeff[f]));
}

__device__ __forceinline__ Ival ba_factor_bounds(const GraphView& G, const uint8_t* assigned, int64_t f, bool& all_assigned) {
  const int32_t cam = __ldg(&G.cam[f]), pt = __ldg(&G.pt[f]);
  Ival vals[12];
#pragma unroll
  for (int s = 0; s < 12; ++s) vals[s] = bounds_var(G, assigned, BaOps::slot_vid(G, cam, pt, s), all_assigned);
  // angleAxisRotatePoint
  Ival p[3] = {vals[9], vals[10], vals[11]};
  const Ival theta = iv_sqrt(iv_add(iv_add(iv_square(vals[0]), iv_square(vals[1])), iv_square(vals[2])));
  Ival v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = iv_div(vals[i], theta);  // simpleNormalize: the norm is the same expression
  Ival ct, st;
  if (iv_width(theta) < 1e-6) {
    const double m = iv_median(theta);
    ct = iv_point(cos(m));
    st = iv_point(sin(m));
  } else {
    ct = iv_cos(theta);
    st = iv_sin(theta);
  }
  const Ival om = iv_sub(iv_point(1.0), ct);
  Ival vcp[3];
  vcp[0] = iv_sub(iv_mul(v[1], p[2]), iv_mul(v[2], p[1]));
  vcp[1] = iv_sub(iv_mul(v[2], p[0]), iv_mul(v[0], p[2]));
  vcp[2] = iv_sub(iv_mul(v[0], p[1]), iv_mul(v[1], p[0]));
  const Ival vdp = iv_add(iv_add(iv_mul(v[0], p[0]), iv_mul(v[1], p[1])), iv_mul(v[2], p[2]));
  Ival q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    q[i] = iv_add(iv_add(iv_mul(p[i], ct), iv_mul(vcp[i], st)), iv_mul(iv_mul(v[i], om), vdp));
  // evalFactor
  q[0] = iv_add(q[0], vals[3]); q[1] = iv_add(q[1], vals[4]); q[2] = iv_add(q[2], vals[5]);
  Ival px = iv_div(iv_neg(q[0]), q[2]);
  Ival py = iv_div(iv_neg(q[1]), q[2]);
  const Ival r2 = iv_add(iv_square(px), iv_square(py));
  const Ival dstn = iv_add(iv_add(iv_point(1.0), iv_mul(vals[7], r2)), iv_mul(vals[8], iv_square(r2)));
  px = iv_mul(iv_mul(vals[6], dstn), px);
  py = iv_mul(iv_mul(vals[6], dstn), py);
  const double2 ob = __ldg(&G.obs[f]);
  const Ival ex = iv_square(iv_sub(px, ob.x));
  const Ival ey = iv_square(iv_sub(py, ob.y));
  return iv_div(iv_add(ex, ey), 2.0);
}

template <class Ops>
__global__ void __launch_bounds__(128) factor_bounds_kernel(GraphView G, const uint8_t* assigned, const int32_t* fids, int64_t nf,
                                                            double* lower, double* upper) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nf; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = fids ? fids[k] : k;
    bool all_assigned = true;
    Ival b = (G.kind == KIND_NLPF) ? nlpf_factor_bounds(G, assigned, f, all_assigned) : ba_factor_bounds(G, assigned, f, all_assigned);
    if (G.fconst_on != nullptr && G.fconst_on[f]) {
      b = iv_point(G.fconst_val[f]);  // isAssignedConstant: eval()
    } else if (all_assigned) {
      double sl;
      b = iv_point(Ops::template value<false>(G, f, 0.0, false, sl));  // areAllVarsAssigned: eval()
    }
    lower[k] = b.lo;
    upper[k] = b.hi;
  }
}

// Bounds of MANY factor lists in one launch (the unassigned bounds of every child of a decomposition,
// Component::computeBounds, src/Component.cpp:592-599 -> OptimizableFunction::computeBounds, OptimizableFunction.cpp:181-216):
// one warp per list, 32 factors at a time bounded lane-parallel, then folded by lane 0's replica IN LIST ORDER
// (bounds = Product(bounds, fb) factor by factor, :194-211) — the sum the reference computes, not a tree.
template <class Ops>
__global__ void __launch_bounds__(128) list_bounds_kernel(GraphView G, const uint8_t* assigned, const int64_t* list_off, const int32_t* fids,
                                                          int64_t nlists, double* sums) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t l = warp0; l < nlists; l += nwarps) {
    const int64_t k0 = list_off[l], k1 = list_off[l + 1];
    double lo = 0.0, hi = 0.0;  // semiring Product identity
    for (int64_t kb = k0; kb < k1; kb += 32) {
      const int64_t k = kb + lane;
      Ival b = iv_point(0.0);
      if (k < k1) {
        const int64_t f = fids[k];
        bool all_assigned = true;
        b = (G.kind == KIND_NLPF) ? nlpf_factor_bounds(G, assigned, f, all_assigned) : ba_factor_bounds(G, assigned, f, all_assigned);
        if (G.fconst_on != nullptr && G.fconst_on[f]) {
          b = iv_point(G.fconst_val[f]);
        } else if (all_assigned) {
          double sl;
          b = iv_point(Ops::template value<false>(G, f, 0.0, false, sl));
        }
      }
      const int cnt = (int)((k1 - kb < 32) ? (k1 - kb) : 32);
      for (int i = 0; i < cnt; ++i) {
        lo += __shfl_sync(0xffffffffu, b.lo, i);
        hi += __shfl_sync(0xffffffffu, b.hi, i);
      }
    }
    if (lane == 0) {
      sums[2 * l] = lo;
      sums[2 * l + 1] = hi;
    }
  }
}

}  // namespace rdisgpu
