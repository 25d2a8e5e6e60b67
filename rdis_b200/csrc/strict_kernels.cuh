// strict_kernels.cuh — CGDSubspaceOptimizer::optimize reproduced EVENT FOR EVENT and OPERATION FOR OPERATION.
//
// rdisgpu_set_option(ctx, "strict", 1) routes every subspace solve through this kernel.  It is the parity
// instrument of the library: with the factor arithmetic compiled without FMA contraction (the whole library is),
// true divisions (QuotBy<kExact>), and every sum folded in the reference's own order, a solve is BIT-IDENTICAL to
// the CPU oracle built with the device's sin/cos (oracle "devtrig" twin) — x, f_init, f_end, iterations, number
// of evaluations — on every component shape (tests/test_gpu_parity.py::test_strict_*).  What it reproduces beyond
// the fast kernels:
//
//   * order of every accumulation: evalFactors sums the factor values left to right in list order
//     (src/OptimizableFunction.cpp:108-132); computeGradientOfSum merges per variable in ascending factor id, the first
//     contribution copied, not added to 0 (src/State.h:157-194); Df1dim::df, Frprmn's gg / dgg are left-to-right over
//     the variables (minimize_nrc.h:440-446, 668-673);
//   * the value cache and its 1e-12 change filter: Variable::assign does not notify the factors of a move below 1e-12
//     (src/Variable.cpp:66-88, approxeq), so Factor::eval returns the value cached at an earlier nearby point
//     (src/Factor.cpp:110-119) while Factor::computeGradient always uses the current values.  The cache (value + dirty
//     flag per factor) lives in HBM and PERSISTS across calls in strict mode, like the reference's Factor objects
//     (rdisgpu_set_x applies the same filter);
//   * the reference's evaluation sequence: fa = func(0) at the top of every line search (CgdMachine `faithful`), the
//     re-assignments of CGDSubspaceOptimizer.cpp:61-80 (quickAssignVals(gdmin.p), restore + re-evaluation if worse).
//
// One CTA per problem, any component shape; thread 0 runs the state machine and the sequential folds.  It is slow by
// construction (a 906-term left-to-right sum is a 906-deep dependency chain) — 3-6x the fast kernels on the bundle
// adjustment wave — and still two orders of magnitude faster than the CPU path.
#pragma once
#include "solve_kernels.cuh"

namespace rdisgpu {

constexpr int kStrictThreads = 128;
constexpr double kAssignTol = 1e-12;  // Variable::assign, src/Variable.cpp:69

struct StrictShared {
  CgdMachine m;
  double f, slope, tnum, gg, dgg;
};

// Variable::assign over the problem's variables + the dirty fan-out (Factor::onVarChanged).
// mode: 0 = closestVal(p), 1 = closestVal(p + alpha*xi), 2 = closestVal(saved start), 3 = raw start (the caller's
// precondition assign, src/RDISOptimizer.cpp:1180: not sanitised)
template <class Ops>
__device__ __forceinline__ void strict_assign(const GraphView& G, const int32_t* vids, int nv, const int32_t* fids, int nf,
                                              int mode, double alpha) {
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const int32_t vid = vids[j];
    const double2 xb = G.xbd[vid];
    double raw;
    if (mode == 1) raw = __dadd_rn(xb.x, __dmul_rn(alpha, xb.y));  // Df1dim: xt[j] = p[j] + x*xi[j]
    else if (mode == 0) raw = xb.x;
    else raw = G.xsave[vid];
    const double val = (mode == 3) ? raw : clamp_to_domain(raw, G.dom[vid]);
    const double old = G.xval[vid];
    G.vchg[vid] = (fabs(val - old) < kAssignTol) ? 0 : 1;
    G.xval[vid] = val;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nf; k += blockDim.x) {
    const int32_t fid = fids[k];
    if (Ops::any_own_changed(G, fid)) G.fdirty[fid] = 1;
  }
  __syncthreads();
}

// evalFactors(list, useCached): recompute the dirty factors, then the left-to-right sum (thread 0 -> sh.f).
template <class Ops>
__device__ __forceinline__ void strict_eval(const GraphView& G, StrictShared& sh, const int32_t* fids, int nf) {
  for (int k = threadIdx.x; k < nf; k += blockDim.x) {
    const int32_t fid = fids[k];
    if (G.fconst_on != nullptr && G.fconst_on[fid]) continue;  // assigned constant: eval() returns the stored constant
    if (G.fdirty[fid]) {
      G.fcache[fid] = Ops::strict_value(G, fid);
      G.fdirty[fid] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < nf; ++k) {
      const int32_t fid = fids[k];
      const double v = (G.fconst_on != nullptr && G.fconst_on[fid]) ? G.fconst_val[fid] : G.fcache[fid];
      s = s + v;
    }
    sh.f = s;
  }
  __syncthreads();
}

// computeGradientOfSum: every factor's partials from the assigned state (no cache involved), then per variable
// the merge in ascending factor id.  dst: 0 = xi (Frprmn: func.df(p, xi)), 1 = scratch gvec slot is NOT used;
// the derivative of variable j lands in G.xbd[vid].y when to_xi, else in G.hvec-independent scratch `dout[j]`.
template <class Ops>
__device__ __forceinline__ void strict_partials(const GraphView& G, const int32_t* fids, int nf) {
  for (int k = threadIdx.x; k < nf; k += blockDim.x) {
    const int32_t fid = fids[k];
    Ops::strict_gradient(G, fid, G.gedge + Ops::edge_base(G, fid));
  }
  __syncthreads();
}

template <class Ops>
__global__ void __launch_bounds__(kStrictThreads) solve_strict_kernel(GraphView G, BatchView B, double* dscr, int maxiters, double ftol) {
  __shared__ StrictShared sh;
  const int pidx = blockIdx.x;
  const ProblemDesc P = B.probs[pidx];
  const int32_t* vids = B.vids + P.var_off;
  const int32_t* fids = B.fids + P.fac_off;
  const int nv = P.nv, nf = P.nf;
  const int32_t stamp = pidx;
  double* deriv = dscr + P.var_off;  // per-variable derivative scratch of this problem (Df1dim's dft)

  if (nf == 0) {  // CGD.cpp:26-29
    for (int j = threadIdx.x; j < nv; j += blockDim.x)
      B.xout[P.var_off + j] = (B.x0 != nullptr) ? B.x0[P.var_off + j] : G.xbd[vids[j]].x;
    if (threadIdx.x == 0) B.res[pidx] = ResultRec{0.0, 0.0, 0, ST_EMPTY, 0, 0};
    return;
  }

  // ---- claim; the caller's assign of xval (raw), then quickAssignVals(vars, xval, sanitize) + sfd(xval) ----
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const int32_t vid = vids[j];
    const double xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : G.xbd[vid].x;
    G.xsave[vid] = xv;
    G.xbd[vid] = make_double2(xv, 0.0);
  }
  for (int k = threadIdx.x; k < nf; k += blockDim.x) G.fstamp[fids[k]] = stamp;
  if (threadIdx.x == 0) sh.m.start(maxiters, ftol, /*faithful=*/true);
  __syncthreads();
  strict_assign<Ops>(G, vids, nv, fids, nf, 3, 0.0);

  double f_init = 0.0;
  while (true) {
    const int req = sh.m.req;  // uniform: written by thread 0 before the last barrier
    const double alpha = sh.m.alpha;
    if (req == REQ_DONE) break;
    if (req == REQ_INIT_GRAD) {
      // initialFval = sfd(xval) (CGD.cpp:37); fp = func(p) (minimize_nrc.h:634: same point, cache hit); func.df(p, xi)
      strict_assign<Ops>(G, vids, nv, fids, nf, 0, 0.0);
      strict_eval<Ops>(G, sh, fids, nf);
      strict_partials<Ops>(G, fids, nf);
      for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        const int32_t vid = vids[j];
        const double gneg = -Ops::gather_var(G, vid, stamp, true);
        G.gvec[vid] = gneg;
        G.hvec[vid] = gneg;
        G.xbd[vid].y = gneg;
      }
      f_init = sh.f;
      __syncthreads();
      if (threadIdx.x == 0) sh.m.on_init(sh.f);
    } else if (req == REQ_VALUE || req == REQ_VALUE_SLOPE) {
      strict_assign<Ops>(G, vids, nv, fids, nf, 1, alpha);
      strict_eval<Ops>(G, sh, fids, nf);
      if (req == REQ_VALUE_SLOPE) {  // Df1dim::df (minimize_nrc.h:438-447): same point, gradient, df1 += dft[j]*xi[j]
        strict_partials<Ops>(G, fids, nf);
        for (int j = threadIdx.x; j < nv; j += blockDim.x) deriv[j] = Ops::gather_var(G, vids[j], stamp, true);
        __syncthreads();
        if (threadIdx.x == 0) {
          double df1 = 0.0;
          for (int j = 0; j < nv; ++j) df1 = __dadd_rn(df1, __dmul_rn(deriv[j], G.xbd[vids[j]].y));
          sh.slope = df1;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) sh.m.on_eval(sh.f, (req == REQ_VALUE_SLOPE) ? sh.slope : 0.0);
    } else if (req == REQ_MOVE) {  // minimize_nrc.h:508-511
      for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        const int32_t vid = vids[j];
        double2 xb = G.xbd[vid];
        xb.y = __dmul_rn(xb.y, alpha);
        xb.x = __dadd_rn(xb.x, xb.y);
        G.xbd[vid] = xb;
      }
      __syncthreads();
      if (threadIdx.x == 0) sh.m.on_moved();
    } else if (req == REQ_GRADIENT) {  // func.df(p, xi), then the loop of :656-673
      strict_assign<Ops>(G, vids, nv, fids, nf, 0, 0.0);
      strict_partials<Ops>(G, fids, nf);
      for (int j = threadIdx.x; j < nv; j += blockDim.x) G.xbd[vids[j]].y = Ops::gather_var(G, vids[j], stamp, true);
      __syncthreads();
      if (threadIdx.x == 0) {
        double tnum = 0.0, gg = 0.0, dgg = 0.0;
        for (int j = 0; j < nv; ++j) {
          const int32_t vid = vids[j];
          const double2 xb = G.xbd[vid];
          const double pj = fabs(xb.x);
          const double t = __dmul_rn(fabs(xb.y), (pj < 1.0) ? 1.0 : pj);
          tnum = (t > tnum) ? t : tnum;
          const double gj = G.gvec[vid];
          gg = __dadd_rn(gg, __dmul_rn(gj, gj));
          dgg = __dadd_rn(dgg, __dmul_rn(__dadd_rn(xb.y, gj), xb.y));
        }
        sh.m.on_gradient(tnum, gg, dgg);
      }
    } else {  // REQ_DIRECTION, :681-685
      const double gam = sh.m.gam;
      for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        const int32_t vid = vids[j];
        const double gj = -G.xbd[vid].y;
        const double hj = __dadd_rn(gj, __dmul_rn(gam, G.hvec[vid]));
        G.gvec[vid] = gj;
        G.hvec[vid] = hj;
        G.xbd[vid].y = hj;
      }
      __syncthreads();
      if (threadIdx.x == 0) sh.m.on_directed();
    }
    __syncthreads();
  }

  // ---- CGD.cpp:61-89: quickAssignVals(gdmin.p); if worse than the start: re-assign the start and re-evaluate ----
  strict_assign<Ops>(G, vids, nv, fids, nf, 0, 0.0);
  double fret = sh.m.fret;
  if (fret > f_init) {
    strict_assign<Ops>(G, vids, nv, fids, nf, 2, 0.0);
    strict_eval<Ops>(G, sh, fids, nf);
    fret = sh.f;
  }
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const int32_t vid = vids[j];
    const double val = G.xval[vid];  // xval[i] = vars[i]->eval()
    G.xbd[vid] = make_double2(val, __longlong_as_double(0x7ff8000000000000LL));
    B.xout[P.var_off + j] = val;
  }
  for (int k = threadIdx.x; k < nf; k += blockDim.x) G.fstamp[fids[k]] = -1;
  if (threadIdx.x == 0) {
    ResultRec r;
    r.f_init = f_init;
    r.f_end = fret;
    r.iters = sh.m.iter;
    r.status = sh.m.status;
    r.n_value = sh.m.n_value;
    r.n_slope = sh.m.n_slope;
    B.res[pidx] = r;
  }
}

// rdisgpu_set_x in strict mode: Variable::assign with the change filter + dirty fan-out over the incidence lists.
__global__ void scatter_x_strict_kernel(GraphView G, int64_t n, const int32_t* vid, const double* x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = vid ? vid[i] : (int32_t)i;
    const double val = x[i];
    const double old = G.xval[v];
    G.xval[v] = val;
    G.xbd[v] = make_double2(val, __longlong_as_double(0x7ff8000000000000LL));
    if (fabs(val - old) < kAssignTol) continue;
    if (G.kind == KIND_NLPF) {
      for (int32_t r = G.vrow[v]; r < G.vrow[v + 1]; ++r) G.fdirty[G.efac[G.vedge[r]]] = 1;
    } else {
      const int32_t ncv = 9 * G.ncams;
      const int32_t* row = (v < ncv) ? G.crow : G.prow;
      const int32_t* lst = (v < ncv) ? G.cfac : G.pfac;
      const int32_t blk = (v < ncv) ? v / 9 : (v - ncv) / 3;
      for (int32_t r = row[blk]; r < row[blk + 1]; ++r) G.fdirty[lst[r]] = 1;
    }
  }
}

__global__ void unset_const_dirty_kernel(uint8_t* fdirty, int64_t n, const int32_t* fid, const uint8_t* on) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!on[i]) fdirty[fid[i]] = 1;
}

}  // namespace rdisgpu
