// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's Levenberg–Marquardt subspace optimizer:
//   LMSubspaceOptimizer::optimize            src/optimizers/LMSubspaceOptimizer.cpp:29-171
//   LMSSOpt::evalFunc / evalJacf             src/optimizers/LMSubspaceOptimizer.cpp:176-204, 207-278
//   LMSSOpt::quickAssignVals (no clamping)   src/optimizers/LMSubspaceOptimizer.cpp:281-297
// and of the solver it calls, dlevmar_der of levmar (M. Lourakis).
//
// PARITY UNPINNED.  levmar is NOT vendored in the reference (README.md:38-40, CMakeLists.txt:189-206:
// an optional, un-versioned external), the reference's default build has USE_LEVMAR off
// (CMakeLists.txt:14) and no test, CLI default or golden value exercises this optimizer.  `levmar_der`
// below restates the published algorithm of levmar 2.6's LEVMAR_DER (lm_core.c) from its documented
// behaviour: e = x - hx with x = NULL (zeros), stop codes 1..7, mu_0 = tau * max diag(J^T J),
// gain-ratio update mu *= max(1/3, 1 - (2 dF/dL - 1)^3), nu doubling on rejection, EPSILON = 1e-12,
// J^T J accumulated rows-last-to-first for n*m < 32*32 and row by row otherwise.  The linear solver is
// an LU decomposition with partial pivoting and implicit row scaling (levmar's built-in AX_EQ_B_LU; a
// LAPACK build of levmar would use Bunch–Kaufman LDL^T — same system, different rounding).
#pragma once
#include <cmath>
#include <limits>
#include <vector>

#include "rdis_oracle.hpp"

namespace oracle {
namespace lm {

// A x = B for a dense m x m system (row-major), LU with partial pivoting and implicit scaling.
// Returns false for a singular matrix (levmar: "Singular matrix A in AX_EQ_B_LU").
inline bool ax_eq_b_lu(const std::vector<double>& A, const std::vector<double>& B, std::vector<double>& x, int m) {
  std::vector<double> a(A), work(m);
  std::vector<int> idx(m);
  x = B;
  for (int i = 0; i < m; ++i) {  // implicit scaling of each row
    double mx = 0.0;
    for (int j = 0; j < m; ++j) {
      const double t = std::fabs(a[i * m + j]);
      if (t > mx) mx = t;
    }
    if (mx == 0.0) return false;
    work[i] = 1.0 / mx;
  }
  for (int j = 0; j < m; ++j) {  // Crout
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
      const double t = work[i] * std::fabs(sum);
      if (t >= mx) {
        mx = t;
        maxi = i;
      }
    }
    if (j != maxi) {
      for (int k = 0; k < m; ++k) std::swap(a[maxi * m + k], a[j * m + k]);
      work[maxi] = work[j];
    }
    idx[j] = maxi;
    if (a[j * m + j] == 0.0) a[j * m + j] = std::numeric_limits<double>::epsilon();
    if (j != m - 1) {
      const double t = 1.0 / a[j * m + j];
      for (int i = j + 1; i < m; ++i) a[i * m + j] *= t;
    }
  }
  int k = 0;
  for (int i = 0; i < m; ++i) {  // forward substitution, unscrambling the permutation
    const int j = idx[i];
    double sum = x[j];
    x[j] = x[i];
    if (k != 0) {
      for (int jj = k - 1; jj < i; ++jj) sum -= a[i * m + jj] * x[jj];
    } else if (sum != 0.0) {
      k = i + 1;
    }
    x[i] = sum;
  }
  for (int i = m - 1; i >= 0; --i) {  // back substitution
    double sum = x[i];
    for (int j = i + 1; j < m; ++j) sum -= a[i * m + j] * x[j];
    x[i] = sum / a[i * m + i];
  }
  return true;
}

struct Info {
  double e0 = 0, e = 0, jte_inf = 0, dp2 = 0, mu_over_max = 0;
  int iters = 0, stop = 0, nfev = 0, njev = 0, nlss = 0;
};

// dlevmar_der with x = NULL.  func(p, hx), jacf(p, jac[n*m] row-major).
template <class Func, class Jacf>
int levmar_der(Func&& func, Jacf&& jacf, std::vector<double>& p, int m, int n, int itmax, const double opts[4], Info& info) {
  const double tau = opts[0], eps1 = opts[1], eps2 = opts[2], eps2_sq = opts[2] * opts[2], eps3 = opts[3];
  const double EPSILON = 1e-12, ONE_THIRD = 0.3333333334;
  std::vector<double> e(n), hx(n), jacTe(m), jac((size_t)n * m), jacTjac((size_t)m * m), Dp(m), diag(m), pDp(m);
  double mu = 0.0, p_eL2, jacTe_inf = 0.0, pDp_eL2, p_L2, Dp_L2 = std::numeric_limits<double>::max(), dF, dL;
  int nu = 2, stop = 0, nfev = 0, njev = 0, nlss = 0, k;
  const long long nm = (long long)n * m;

  func(p, hx);
  nfev = 1;
  p_eL2 = 0.0;
  for (int i = 0; i < n; ++i) {  // e = x - hx = -hx, ||e||^2
    e[i] = -hx[i];
    p_eL2 += e[i] * e[i];
  }
  info.e0 = p_eL2;
  if (!std::isfinite(p_eL2)) stop = 7;

  for (k = 0; k < itmax && !stop; ++k) {
    if (p_eL2 <= eps3) {
      stop = 6;
      break;
    }
    jacf(p, jac);
    ++njev;
    if (nm < 32 * 32) {  // small problem: rows last to first, lower triangle
      for (double& v : jacTjac) v = 0.0;
      for (double& v : jacTe) v = 0.0;
      for (int l = n; l-- > 0;) {
        const double* jaclm = &jac[(size_t)l * m];
        for (int i = m; i-- > 0;) {
          const double alpha = jaclm[i];
          for (int j = i + 1; j-- > 0;) jacTjac[(size_t)i * m + j] += jaclm[j] * alpha;
          jacTe[i] += alpha * e[l];
        }
      }
      for (int i = m; i-- > 0;)
        for (int j = i + 1; j < m; ++j) jacTjac[(size_t)i * m + j] = jacTjac[(size_t)j * m + i];
    } else {  // large problem: blocked product in levmar; restated as a plain row-by-row accumulation
      for (double& v : jacTjac) v = 0.0;
      for (double& v : jacTe) v = 0.0;
      for (int l = 0; l < n; ++l) {
        const double* row = &jac[(size_t)l * m];
        for (int i = 0; i < m; ++i) {
          for (int j = 0; j <= i; ++j) jacTjac[(size_t)i * m + j] += row[i] * row[j];
          jacTe[i] += row[i] * e[l];
        }
      }
      for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j) jacTjac[(size_t)i * m + j] = jacTjac[(size_t)j * m + i];
    }
    p_L2 = jacTe_inf = 0.0;
    for (int i = 0; i < m; ++i) {
      const double t = std::fabs(jacTe[i]);
      if (jacTe_inf < t) jacTe_inf = t;
      diag[i] = jacTjac[(size_t)i * m + i];
      p_L2 += p[i] * p[i];
    }
    if (jacTe_inf <= eps1) {
      Dp_L2 = 0.0;
      stop = 1;
      break;
    }
    if (k == 0) {
      double t = -std::numeric_limits<double>::max();
      for (int i = 0; i < m; ++i)
        if (diag[i] > t) t = diag[i];
      mu = tau * t;
    }
    while (true) {
      for (int i = 0; i < m; ++i) jacTjac[(size_t)i * m + i] += mu;
      const bool issolved = ax_eq_b_lu(jacTjac, jacTe, Dp, m);
      ++nlss;
      if (issolved) {
        Dp_L2 = 0.0;
        for (int i = 0; i < m; ++i) {
          pDp[i] = p[i] + Dp[i];
          Dp_L2 += Dp[i] * Dp[i];
        }
        if (Dp_L2 <= eps2_sq * p_L2) {
          stop = 2;
          break;
        }
        if (Dp_L2 >= (p_L2 + eps2) / (EPSILON * EPSILON)) {
          stop = 4;
          break;
        }
        func(pDp, hx);
        ++nfev;
        pDp_eL2 = 0.0;
        for (int i = 0; i < n; ++i) {
          hx[i] = -hx[i];
          pDp_eL2 += hx[i] * hx[i];
        }
        if (!std::isfinite(pDp_eL2)) {
          stop = 7;
          break;
        }
        dL = 0.0;
        for (int i = 0; i < m; ++i) dL += Dp[i] * (mu * Dp[i] + jacTe[i]);
        dF = p_eL2 - pDp_eL2;
        if (dL > 0.0 && dF > 0.0) {
          double t = (2.0 * dF / dL - 1.0);
          t = 1.0 - t * t * t;
          mu = mu * ((t >= ONE_THIRD) ? t : ONE_THIRD);
          nu = 2;
          for (int i = 0; i < m; ++i) p[i] = pDp[i];
          for (int i = 0; i < n; ++i) e[i] = hx[i];
          p_eL2 = pDp_eL2;
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
      for (int i = 0; i < m; ++i) jacTjac[(size_t)i * m + i] = diag[i];
    }
  }
  if (k >= itmax) stop = 3;
  info.e = p_eL2;
  info.jte_inf = jacTe_inf;
  info.dp2 = Dp_L2;
  info.iters = k;
  info.stop = stop;
  info.nfev = nfev;
  info.njev = njev;
  info.nlss = nlss;
  return (stop != 4 && stop != 7) ? k : -1;
}

}  // namespace lm

// src/optimizers/LMSubspaceOptimizer.cpp:29-171 (the USE_LEVMAR branch)
class LMSubspaceOptimizer : public SubspaceOptimizer {
 public:
  explicit LMSubspaceOptimizer(OptimizableFunction& f_) : SubspaceOptimizer(f_) {}
  lm::Info lastInfo;

  Numeric optimize(const std::vector<Variable*>& vars, const std::vector<Factor*>& factors, std::vector<Numeric>& xval,
                   Numeric& deltaFval, bool /*printdbg*/) override {
    const int m = (int)vars.size();
    const int n = (int)std::max(factors.size(), (size_t)m);  // :48-49
    auto assign = [&](const std::vector<double>& p) {          // LMSSOpt::quickAssignVals: no clamping (:281-297)
      for (int i = 0; i < m; ++i) {
        vars[i]->assign(p[i]);
        f.onVarAssigned(vars[i]->getID(), p[i]);
      }
    };
    auto func = [&](const std::vector<double>& p, std::vector<double>& hx) {  // :176-204
      assign(p);
      for (int i = 0; i < n; ++i) hx[i] = ((size_t)i < factors.size()) ? std::sqrt(factors[i]->eval(f.counters) * 2.0) : 0.0;
    };
    PartialGradient pgt;
    auto jacf = [&](const std::vector<double>& p, std::vector<double>& jac) {  // :207-278
      assign(p);
      for (int j = 0; j < n; ++j) {
        Numeric feval = 0.0;
        pgt.clear();
        if ((size_t)j < factors.size()) {
          factors[j]->computeGradient(pgt);
          feval = std::sqrt(factors[j]->eval(f.counters) * 2.0);
        }
        for (int i = 0; i < m; ++i) {
          const Numeric* d = pgt.find(vars[i]->getID());
          jac[(size_t)j * m + i] = (d == nullptr) ? 0 : (*d / feval);
        }
      }
    };
    const Numeric ival = f.evalFactors(factors, true);  // :81
    const double opts[4] = {1e-3, 1e-15, 1e-15, ftol};  // :83-86
    lm::levmar_der(func, jacf, xval, m, n, (int)maxiters, opts, lastInfo);
    assign(xval);                                        // :102
    for (int i = 0; i < m; ++i) xval[i] = vars[i]->eval();
    const Numeric fval = f.evalFactors(factors, true);  // :110
    deltaFval = (fval - ival);
    lastIters = (size_t)lastInfo.iters;
    return fval;
  }
};

}  // namespace oracle
