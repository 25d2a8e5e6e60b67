// lm_kernels.cuh — batched LMSubspaceOptimizer::optimize on the device (PARITY UNPINNED).
//
// Reference: src/optimizers/LMSubspaceOptimizer.cpp:29-171 — residual of factor j is hx_j = sqrt(2 f_j)
// (:199), the Jacobian is dense row-major n x m with J_ji = (d f_j / d x_i) / hx_j (:257-276),
// n = max(|factors|, m) with zero padding rows (:48-49, :200-202), no domain clamping anywhere
// (:281-297), options {tau 1e-3, eps1 1e-15, eps2 1e-15, eps3 = ftol} (:83-86), and the solver is
// levmar's dlevmar_der with x = NULL.  levmar is not vendored with the reference and no reference test
// runs this optimizer, so the arithmetic below follows the published algorithm of levmar 2.6
// (LEVMAR_DER, lm_core.c) as restated in oracle/lm_oracle.hpp — against which it is tested.
//
// Mapping: one CTA per component (m <= kLmMaxVars variables, any number of factors).
//   func / jacf   one thread per factor (strided): f_j, its partials (Ops::gradient), the dense row
//                 of J in an HBM scratch (L2-resident at these sizes: 906 x 9 doubles for the largest
//                 ladybug camera block)
//   J^T J, J^T e  one thread per (i, j >= ) pair and row chunk; chunk partials folded in fixed order.
//                 The per-component blocks are at most 32 x 32 (9 x 9 / 3 x 3 on bundle adjustment):
//                 far below a tensor-core tile — a DMMA path only pays for a genuinely dense large
//                 component, which this round does not cover.
//   (J^T J + mu I) dp = J^T e   LU with partial pivoting and implicit scaling by one thread
//                 (operation order identical to the oracle's), m <= 32
//   control flow  thread 0, decisions broadcast through shared memory
#pragma once
#include "factors.cuh"
#include "ba_block_kernels.cuh"
#include "solve_kernels.cuh"

namespace rdisgpu {

constexpr int kLmMaxVars = 32;
constexpr int kLmThreads = 128;

struct LmView {
  int32_t* vloc;            // i32[V], -1 = not a variable of a running LM problem
  double* scratch;          // per problem: jac[n*m] | e[n] | hx[n]
  const int64_t* scr_off;   // [nprobs + 1] offsets into scratch
};

struct LmShared {
  double p[kLmMaxVars], pDp[kLmMaxVars], Dp[kLmMaxVars], jacTe[kLmMaxVars], diag[kLmMaxVars], work[kLmMaxVars];
  double A[kLmMaxVars * kLmMaxVars];   // J^T J (lower triangle mirrored), augmented in place
  double LU[kLmMaxVars * kLmMaxVars];
  double part[kLmThreads];             // chunk partials of the normal equations
  double red[2][kLmThreads / 32][2];
  int32_t vids[kLmMaxVars];
  int idx[kLmMaxVars];
  double s_sumf, s_e2;
  int decision;
};

__device__ __forceinline__ int32_t slot_var(const GraphView& G, NlpfOps*, int64_t fid, int s) {
  return __ldg(&G.evid[__ldg(&G.rowptr[fid]) + s]);
}
__device__ __forceinline__ int32_t slot_var(const GraphView& G, BaOps*, int64_t fid, int s) {
  return BaOps::slot_vid(G, __ldg(&G.cam[fid]), __ldg(&G.pt[fid]), s);
}

// block-wide sum of two values, every thread gets the totals (fixed order)
__device__ __forceinline__ void lm_block_sum2(LmShared& sh, int& flip, double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh.red[flip][warp][0] = a;
    sh.red[flip][warp][1] = b;
  }
  __syncthreads();
  a = sh.red[flip][0][0];
  b = sh.red[flip][0][1];
  for (int w = 1; w < kLmThreads / 32; ++w) {
    a += sh.red[flip][w][0];
    b += sh.red[flip][w][1];
  }
  flip ^= 1;
}

// A x = B, LU with partial pivoting and implicit scaling (oracle/lm_oracle.hpp: ax_eq_b_lu), one thread.
__device__ inline bool lm_ax_eq_b_lu(LmShared& sh, int m) {
  double* a = sh.LU;
  double* x = sh.Dp;
  for (int i = 0; i < m * m; ++i) a[i] = sh.A[(i / m) * kLmMaxVars + (i % m)];
  for (int i = 0; i < m; ++i) x[i] = sh.jacTe[i];
  for (int i = 0; i < m; ++i) {
    double mx = 0.0;
    for (int j = 0; j < m; ++j) {
      const double t = fabs(a[i * m + j]);
      if (t > mx) mx = t;
    }
    if (mx == 0.0) return false;
    sh.work[i] = 1.0 / mx;
  }
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i < j; ++i) {
      double sum = a[i * m + j];
      for (int k = 0; k < i; ++k) sum -= a[i * m + k] * a[k * m + j];
      a[i * m + j] = sum;
    }
    double mx = 0.0;
    int maxi = -1;
    for (int i = j; i < m; ++i) {
      double sum = a[i * m + j];
      for (int k = 0; k < j; ++k) sum -= a[i * m + k] * a[k * m + j];
      a[i * m + j] = sum;
      const double t = sh.work[i] * fabs(sum);
      if (t >= mx) {
        mx = t;
        maxi = i;
      }
    }
    if (maxi < 0) return false;  // NaN column
    if (j != maxi) {
      for (int k = 0; k < m; ++k) {
        const double t = a[maxi * m + k];
        a[maxi * m + k] = a[j * m + k];
        a[j * m + k] = t;
      }
      sh.work[maxi] = sh.work[j];
    }
    sh.idx[j] = maxi;
    if (a[j * m + j] == 0.0) a[j * m + j] = 2.220446049250313e-16;
    if (j != m - 1) {
      const double t = 1.0 / a[j * m + j];
      for (int i = j + 1; i < m; ++i) a[i * m + j] *= t;
    }
  }
  int k = 0;
  for (int i = 0; i < m; ++i) {
    const int j = sh.idx[i];
    double sum = x[j];
    x[j] = x[i];
    if (k != 0) {
      for (int jj = k - 1; jj < i; ++jj) sum -= a[i * m + jj] * x[jj];
    } else if (sum != 0.0) {
      k = i + 1;
    }
    x[i] = sum;
  }
  for (int i = m - 1; i >= 0; --i) {
    double sum = x[i];
    for (int j = i + 1; j < m; ++j) sum -= a[i * m + j] * x[j];
    x[i] = sum / a[i * m + i];
  }
  return true;
}

enum LmDecision : int { LM_CONTINUE = 0, LM_STOP = 1, LM_TRY = 2, LM_ACCEPT = 3, LM_REJECT = 4 };

template <class Ops>
__global__ void __launch_bounds__(kLmThreads) solve_lm_block_kernel(GraphView G, BatchView B, LmView L, const int32_t* order,
                                                                    int itmax, double tau, double eps1, double eps2, double eps3) {
  __shared__ LmShared sh;
  const int pidx = order ? order[blockIdx.x] : (int)blockIdx.x;
  const ProblemDesc P = B.probs[pidx];
  const int m = P.nv, nf = P.nf;
  const int n = (nf > m) ? nf : m;
  const int32_t* fids = B.fids + P.fac_off;
  const int tid = threadIdx.x;
  int flip = 0;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  if (nf == 0) {  // nothing to optimise (same contract as the CGD path)
    for (int i = tid; i < m; i += kLmThreads) {
      const int32_t vid = B.vids[P.var_off + i];
      B.xout[P.var_off + i] = (B.x0 != nullptr) ? B.x0[P.var_off + i] : G.xbd[vid].x;
    }
    if (tid == 0) B.res[pidx] = ResultRec{0.0, 0.0, 0, 0 /* stop code 0: nothing to do */, 0, 0};
    return;
  }
  double* jac = L.scratch + L.scr_off[pidx];
  double* e = jac + (size_t)n * m;
  double* hx = e + n;

  for (int i = tid; i < m; i += kLmThreads) {
    const int32_t vid = B.vids[P.var_off + i];
    sh.vids[i] = vid;
    sh.p[i] = (B.x0 != nullptr) ? B.x0[P.var_off + i] : G.xbd[vid].x;
    L.vloc[vid] = i;
  }
  __syncthreads();

  // func: assign q (no clamping), hx_j = sqrt(2 f_j), returns sum f_j and sum hx_j^2 to every thread
  auto func = [&](const double* q, double& sumf, double& e2) {
    for (int i = tid; i < m; i += kLmThreads) {
      G.xbd[sh.vids[i]] = make_double2(q[i], qnan);
      G.xval[sh.vids[i]] = q[i];
    }
    __syncthreads();
    double sf = 0.0, s2 = 0.0;
    for (int k = tid; k < n; k += kLmThreads) {
      double h = 0.0;
      if (k < nf) {
        double sl;
        const int32_t fid = fids[k];
        double fv = Ops::template value<false>(G, fid, 0.0, false, sl);
        if (G.fconst_on != nullptr && G.fconst_on[fid]) fv = G.fconst_val[fid];
        sf += fv;
        h = sqrt(fv * 2.0);
      }
      hx[k] = h;
      s2 += h * h;
    }
    lm_block_sum2(sh, flip, sf, s2);
    sumf = sf;
    e2 = s2;
  };

  double ival, p_eL2;
  func(sh.p, ival, p_eL2);
  for (int k = tid; k < n; k += kLmThreads) e[k] = -hx[k];
  __syncthreads();

  double mu = 0.0, Dp_L2 = 1.7976931348623157e308, jacTe_inf = 0.0;
  int nu = 2, stop = 0, nfev = 1, njev = 0, k_it = 0;
  if (!(fabs(p_eL2) <= 1.7976931348623157e308)) stop = 7;
  const int npairs = m * (m + 1) / 2 + m;  // lower triangle of J^T J, then J^T e
  int nchunk = kLmThreads / npairs;
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 8) nchunk = 8;

  for (k_it = 0; k_it < itmax && !stop; ++k_it) {
    if (p_eL2 <= eps3) {
      stop = 6;
      break;
    }
    // ---- jacf at p (the device state already holds p) ----
    for (int k = tid; k < n; k += kLmThreads) {
      double* row = jac + (size_t)k * m;
      for (int i = 0; i < m; ++i) row[i] = 0.0;
      if (k < nf) {
        const int32_t fid = fids[k];
        double* ge = G.gedge + Ops::edge_base(G, fid);
        double fv = Ops::gradient(G, fid, ge);
        if (G.fconst_on != nullptr && G.fconst_on[fid]) fv = G.fconst_val[fid];
        const double feval = sqrt(fv * 2.0);
        const int ar = Ops::arity(G, fid);
        for (int s = 0; s < ar; ++s) {
          const int li = L.vloc[slot_var(G, (Ops*)nullptr, fid, s)];
          if (li >= 0) row[li] = ge[s] / feval;
        }
      }
    }
    ++njev;
    __syncthreads();
    // ---- J^T J (lower triangle) and J^T e ----
    for (int base = 0; base < npairs; base += kLmThreads / nchunk) {
      const int q = base + tid / nchunk;   // pair index
      const int c = tid % nchunk;          // row chunk
      double acc = 0.0;
      if (tid < (kLmThreads / nchunk) * nchunk && q < npairs) {
        const int l0 = (int)((long long)n * c / nchunk), l1 = (int)((long long)n * (c + 1) / nchunk);
        if (q < m * (m + 1) / 2) {
          int i = 0;
          while ((i + 1) * (i + 2) / 2 <= q) ++i;
          const int j = q - i * (i + 1) / 2;
          for (int l = l0; l < l1; ++l) acc += jac[(size_t)l * m + i] * jac[(size_t)l * m + j];
        } else {
          const int i = q - m * (m + 1) / 2;
          for (int l = l0; l < l1; ++l) acc += jac[(size_t)l * m + i] * e[l];
        }
      }
      sh.part[tid] = acc;
      __syncthreads();
      if (c == 0 && tid < (kLmThreads / nchunk) * nchunk && q < npairs) {
        double t = sh.part[tid];
        for (int cc = 1; cc < nchunk; ++cc) t += sh.part[tid + cc];
        if (q < m * (m + 1) / 2) {
          int i = 0;
          while ((i + 1) * (i + 2) / 2 <= q) ++i;
          const int j = q - i * (i + 1) / 2;
          sh.A[i * kLmMaxVars + j] = t;
          sh.A[j * kLmMaxVars + i] = t;
        } else {
          sh.jacTe[q - m * (m + 1) / 2] = t;
        }
      }
      __syncthreads();
    }
    // ---- scalar part: thread-uniform values recomputed by every thread from shared memory ----
    double p_L2 = 0.0;
    jacTe_inf = 0.0;
    for (int i = 0; i < m; ++i) {
      const double t = fabs(sh.jacTe[i]);
      if (jacTe_inf < t) jacTe_inf = t;
      p_L2 += sh.p[i] * sh.p[i];
    }
    if (tid == 0)
      for (int i = 0; i < m; ++i) sh.diag[i] = sh.A[i * kLmMaxVars + i];
    __syncthreads();
    if (jacTe_inf <= eps1) {
      Dp_L2 = 0.0;
      stop = 1;
      break;
    }
    if (k_it == 0) {
      double t = -1.7976931348623157e308;
      for (int i = 0; i < m; ++i)
        if (sh.diag[i] > t) t = sh.diag[i];
      mu = tau * t;
    }
    // ---- inner loop: adaptive damping ----
    while (true) {
      if (tid == 0) {
        for (int i = 0; i < m; ++i) sh.A[i * kLmMaxVars + i] += mu;
        const bool issolved = lm_ax_eq_b_lu(sh, m);
        int dec = LM_REJECT;
        if (issolved) {
          double d2 = 0.0;
          for (int i = 0; i < m; ++i) {
            sh.pDp[i] = sh.p[i] + sh.Dp[i];
            d2 += sh.Dp[i] * sh.Dp[i];
          }
          sh.s_e2 = d2;
          dec = LM_TRY;
        }
        sh.decision = dec;
      }
      __syncthreads();
      int dec = sh.decision;
      if (dec == LM_TRY) {
        Dp_L2 = sh.s_e2;
        if (Dp_L2 <= eps2 * eps2 * p_L2) {
          stop = 2;
          break;
        }
        if (Dp_L2 >= (p_L2 + eps2) / (1e-12 * 1e-12)) {
          stop = 4;
          break;
        }
        double sumf, pDp_eL2;
        func(sh.pDp, sumf, pDp_eL2);
        ++nfev;
        if (!(fabs(pDp_eL2) <= 1.7976931348623157e308)) {
          stop = 7;
          break;
        }
        double dL = 0.0;
        for (int i = 0; i < m; ++i) dL += sh.Dp[i] * (mu * sh.Dp[i] + sh.jacTe[i]);
        const double dF = p_eL2 - pDp_eL2;
        if (dL > 0.0 && dF > 0.0) {
          double t = (2.0 * dF / dL - 1.0);
          t = 1.0 - t * t * t;
          mu = mu * ((t >= 0.3333333334) ? t : 0.3333333334);
          nu = 2;
          __syncthreads();
          for (int i = tid; i < m; i += kLmThreads) sh.p[i] = sh.pDp[i];
          for (int k = tid; k < n; k += kLmThreads) e[k] = -hx[k];
          p_eL2 = pDp_eL2;
          __syncthreads();
          break;
        }
      }
      mu *= nu;
      const int nu2 = nu << 1;
      if (nu2 <= nu) {
        stop = 5;
        break;
      }
      nu = nu2;
      __syncthreads();
      if (tid == 0)
        for (int i = 0; i < m; ++i) sh.A[i * kLmMaxVars + i] = sh.diag[i];
      __syncthreads();
    }
  }
  if (k_it >= itmax) stop = 3;  // levmar: "if(k>=itmax) stop=3" overrides a stop raised in the last iteration

  // ---- commit: LMSSOpt::quickAssignVals(xval) without clamping, then fval = evalFactors (:102-110) ----
  __syncthreads();
  double fval, e2;
  func(sh.p, fval, e2);
  for (int i = tid; i < m; i += kLmThreads) {
    B.xout[P.var_off + i] = sh.p[i];
    L.vloc[sh.vids[i]] = -1;
  }
  if (tid == 0) {
    ResultRec r;
    r.f_init = ival;
    r.f_end = fval;
    r.iters = k_it;
    r.status = stop;
    r.n_value = nfev;
    r.n_slope = njev;
    B.res[pidx] = r;
  }
}

// ------------------------------------------------------------------------------------------
// Bundle-adjustment point blocks (3 variables, <= 32 observations): the same Levenberg-Marquardt
// iteration in registers, one lane per observation, 32/G problems per warp — the LM counterpart of
// solve_ba_points_kernel.  Every pass of the loop is ONE evaluation (value + the three point partials)
// of every unfinished problem of the warp, executed converged; the 3 x 3 normal equations are reduced
// with tile shuffles and solved redundantly by every lane (LDL^T: J^T J + mu I is positive definite).
// Same control flow, stop codes and counters as solve_lm_block_kernel; sums are folded in shuffle-tree
// order instead of row order, so results agree with it (and with the oracle) to rounding.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lm_solve_spd3(const double (&A)[6], double mu, const double (&b)[3], double (&x)[3]) {
  // A = [a00 a10 a11 a20 a21 a22] (lower triangle), solve (A + mu I) x = b
  const double a00 = A[0] + mu, a10 = A[1], a11 = A[2] + mu, a20 = A[3], a21 = A[4], a22 = A[5] + mu;
  const double d0 = a00;
  if (!(d0 > 0.0)) return false;
  const double l10 = a10 / d0, l20 = a20 / d0;
  const double d1 = a11 - l10 * a10;
  if (!(d1 > 0.0)) return false;
  const double l21 = (a21 - l20 * a10) / d1;
  const double d2 = a22 - l20 * a20 - l21 * (a21 - l20 * a10);
  if (!(d2 > 0.0)) return false;
  const double y0 = b[0];
  const double y1 = b[1] - l10 * y0;
  const double y2 = b[2] - l20 * y0 - l21 * y1;
  const double z2 = y2 / d2;
  const double z1 = y1 / d1 - l21 * z2;
  const double z0 = y0 / d0 - l10 * z1 - l20 * z2;
  x[0] = z0; x[1] = z1; x[2] = z2;
  return true;
}

__global__ void __launch_bounds__(32) solve_lm_ba_points_kernel(GraphView Gv, BatchView B, const int32_t* order,
                                                                const PointWarpTask* tasks, int itmax, double tau, double eps1,
                                                                double eps2, double eps3) {
  const PointWarpTask t = tasks[blockIdx.x];
  const int lg = t.lg;
  const int G = 1 << lg;
  const int lane = threadIdx.x & 31;
  const int r = lane & (G - 1);
  const int slot = lane >> lg;
  const int pidx = (slot < t.count) ? order[t.first + slot] : -1;
  const bool run = (pidx >= 0);
  ProblemDesc P;
  P.var_off = 0; P.fac_off = 0; P.nv = 0; P.nf = 0;
  if (run) P = B.probs[pidx];
  const int nf = P.nf;
  const int32_t v0 = run ? B.vids[P.var_off] : 0;
  const unsigned kFull = 0xffffffffu;

  double p[3], q[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    p[j] = 0.0;
    if (run) p[j] = (B.x0 != nullptr) ? B.x0[P.var_off + j] : Gv.xbd[v0 + j].x;
    q[j] = p[j];
  }
  const bool have = run && (r < nf);
  double x[12];
  double2 ob = make_double2(0.0, 0.0);
  BaOps::Fwd m;
  bool fc_on = false;
  double fc_val = 0.0;
#pragma unroll
  for (int s = 0; s < 12; ++s) x[s] = 0.0;
  m.a0 = m.a1 = m.a2 = m.theta = m.s = 0.0; m.c = 1.0;
  if (have) {
    const int32_t fid = B.fids[P.fac_off + r];
    const int32_t cam = __ldg(&Gv.cam[fid]);
    ob = __ldg(&Gv.obs[fid]);
#pragma unroll
    for (int s = 0; s < 9; ++s) x[s] = Gv.xbd[9 * cam + s].x;
    BaOps::rotation(x[0], x[1], x[2], m);
    if (Gv.fconst_on != nullptr && Gv.fconst_on[fid]) {
      fc_on = true;
      fc_val = Gv.fconst_val[fid];
    }
  }

  // levmar state (replicated in every lane of the tile)
  double A[6] = {0, 0, 0, 0, 0, 0}, Jte[3] = {0, 0, 0}, Dp[3] = {0, 0, 0};
  double mu = 0.0, p_eL2 = 0.0, ival = 0.0, fval = 0.0, p_L2 = 0.0;
  int nu = 2, stop = 0, nfev = 0, njev = 0, k_it = 0;
  bool fin = !run, first = true, at_commit = false;

  while (true) {
    if (__all_sync(kFull, fin)) break;
    // ---- one evaluation at q: value, residual, the three point partials ----
    double fv = 0.0, h = 0.0, j0 = 0.0, j1 = 0.0, j2 = 0.0;
    if (have && !fin) {
      x[9] = q[0]; x[10] = q[1]; x[11] = q[2];  // no clamping (LMSubspaceOptimizer.cpp:281-297)
      fv = BaOps::project(x, ob, m);
      double gq[12];
      BaOps::partials(x, m, gq);
      if (fc_on) fv = fc_val;
      h = sqrt(fv * 2.0);
      j0 = gq[9] / h; j1 = gq[10] / h; j2 = gq[11] / h;
    }
    // tile reductions: sum f, ||hx||^2, J^T J (lower), J^T e with e = -hx
    double red[11] = {fv, h * h, j0 * j0, j1 * j0, j1 * j1, j2 * j0, j2 * j1, j2 * j2, j0 * (-h), j1 * (-h), j2 * (-h)};
    for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 11; ++i) red[i] += __shfl_xor_sync(kFull, red[i], o);
    }
    if (fin) continue;

    // ---- the scalar part of levmar's loop (diverges between the tiles of a warp) ----
    const double e2 = red[1];
    bool new_point = false;  // q became the current point p: redo the top of the outer loop
    if (at_commit) {         // final evalFactors at the returned point (LMSubspaceOptimizer.cpp:110)
      fval = red[0];
      fin = true;
    } else if (first) {
      first = false;
      ival = red[0];
      p_eL2 = e2;
      nfev = 1;
      if (!(fabs(p_eL2) <= 1.7976931348623157e308)) stop = 7;
      new_point = true;
    } else {
      ++nfev;
      const double pDp_eL2 = e2;
      bool accepted = false;
      if (!(fabs(pDp_eL2) <= 1.7976931348623157e308)) {
        stop = 7;
      } else {
        double dL = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
        const double dF = p_eL2 - pDp_eL2;
        if (dL > 0.0 && dF > 0.0) {
          double tt = (2.0 * dF / dL - 1.0);
          tt = 1.0 - tt * tt * tt;
          mu = mu * ((tt >= 0.3333333334) ? tt : 0.3333333334);
          nu = 2;
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = q[i];
          p_eL2 = pDp_eL2;
          accepted = true;
        }
      }
      if (stop) {
        ++k_it;  // the for-loop's increment after the inner break
      } else if (accepted) {
        ++k_it;
        new_point = true;
      } else {
        mu *= nu;
        const int nu2 = nu << 1;
        if (nu2 <= nu) {
          stop = 5;
          ++k_it;
        }
        nu = nu2;
      }
    }
    // top of the outer loop at a (new) current point: stop tests, normal equations of THIS evaluation
    if (!fin && !stop && new_point) {
      if (k_it >= itmax) {
        stop = 3;
      } else if (p_eL2 <= eps3) {
        stop = 6;
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) A[i] = red[2 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) Jte[i] = red[8 + i];
        ++njev;
        double jinf = 0.0;
        p_L2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double a = fabs(Jte[i]);
          if (jinf < a) jinf = a;
          p_L2 += p[i] * p[i];
        }
        if (jinf <= eps1) {
          stop = 1;
        } else if (k_it == 0) {
          double mx = A[0];
          if (A[2] > mx) mx = A[2];
          if (A[5] > mx) mx = A[5];
          mu = tau * mx;
        }
      }
    }
    // next trial point (also after a rejection): solve, step tests
    if (!fin && !stop) {
      while (true) {
        if (lm_solve_spd3(A, mu, Jte, Dp)) {
          double d2 = 0.0;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            q[i] = p[i] + Dp[i];
            d2 += Dp[i] * Dp[i];
          }
          if (d2 <= eps2 * eps2 * p_L2) {
            stop = 2;
            ++k_it;
          } else if (d2 >= (p_L2 + eps2) / (1e-12 * 1e-12)) {
            stop = 4;
            ++k_it;
          }
          break;
        }
        mu *= nu;  // the linear system could not be solved: reject without an evaluation
        const int nu2 = nu << 1;
        if (nu2 <= nu) {
          stop = 5;
          ++k_it;
          break;
        }
        nu = nu2;
      }
    }
    if (!fin && stop) {  // leave the loop: one more evaluation at the returned point p
      if (k_it >= itmax) stop = 3;
      at_commit = true;
#pragma unroll
      for (int i = 0; i < 3; ++i) q[i] = p[i];
    }
  }

  if (!run) return;
  if (r == 0) {
    const double qn = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Gv.xbd[v0 + j] = make_double2(p[j], qn);
      Gv.xval[v0 + j] = p[j];
      B.xout[P.var_off + j] = p[j];
    }
    ResultRec res;
    res.f_init = ival;
    res.f_end = fval;
    res.iters = k_it;
    res.status = stop;
    res.n_value = nfev;
    res.n_slope = njev;
    B.res[pidx] = res;
  }
}

}  // namespace rdisgpu
