// lm_dense.cuh — Levenberg-Marquardt for ONE component of any size (PARITY UNPINNED, like lm_kernels.cuh).
//
// LMSubspaceOptimizer::optimize (src/optimizers/LMSubspaceOptimizer.cpp:29-171) hands levmar a dense n x m Jacobian
// (:48-49, :207-278) whatever the component's size — 31 843 x 4 754 doubles for the block RDIS poses first on ladybug.
// Here the Jacobian stays what it is, block sparse (one row per factor, <= arity entries), and only the normal
// equations are dense:
//   lm_func_kernel      hx_j = sqrt(2 f_j), e = -hx, sum f, sum hx^2            one thread per factor   (LMSSOpt::evalFunc, :176-204)
//   lm_rows_kernel      row j of J: (d f_j / d x_i) / hx_j at the factor's slots (LMSSOpt::evalJacf, :207-278)
//   lm_assemble_kernel  A = J^T J (lower triangle, dense m x m in HBM) and g = J^T e: one warp per VARIABLE walks the
//                       variable's incident rows in ascending factor order, lanes over the row's entries — no atomics,
//                       a fixed accumulation order
//   (A + mu I) dp = g   blocked right-looking Cholesky, block width 64:
//       lm_potrf_kernel   the 64 x 64 diagonal block in shared memory, and its inverse
//       lm_tile_kernel    the two dense contractions of the factorisation on the FP64 tensor cores (mma.sync.m8n8k4.f64,
//                         SASS DMMA; 64 x 64 tiles, 4 warps x (4 x 4) accumulator fragments, panels staged in shared
//                         memory): the panel solve L21 = A21 L11^-T and the trailing update C -= L21 L21^T (2/3 m^3 of
//                         the factorisation's flops)
//       lm_trsv_kernel    forward and back substitution, one CTA
//   control flow        levmar's dlevmar_der (stop codes 1..7, mu0 = tau max diag, gain ratio, nu doubling) on the host:
//                       a handful of scalars come back per iteration, which is noise against a 4 754^3 / 3 factorisation
// levmar solves the augmented system by LU with partial pivoting; A + mu I is symmetric positive definite for mu > 0, so
// Cholesky returns the same step up to rounding (a non-positive pivot is reported as "singular": mu is increased, as
// levmar does when its solver fails).
#pragma once
#include "factors.cuh"
#include "lm_kernels.cuh"

namespace rdisgpu {

constexpr int kLmNB = 64;          // block width of the factorisation = tile edge of the trailing update
constexpr int kLmKH = 32;           // columns of a panel staged in shared memory at a time
constexpr int kLmTileLd = kLmKH + 4;  // shared-memory row stride of a staged panel (conflict-free 8-byte fragment loads)

struct LmDenseView {
  int m, nf;                 // variables, factors of the component
  const int32_t* fids;       // device: the component's factor list
  const int32_t* vids;       // device: its variables
  const int32_t* vloc;       // i32[V]: variable -> local index, -1 = not a variable of the component
  const int32_t* jptr;       // i32[nf + 1]: first entry of row j in jcol / jval (row length = the factor's arity)
  int32_t* jcol;             // local column of every entry, -1 for a frozen variable
  double* jval;
  double* e;                 // f64[nf]
  double* hx;                // f64[nf]
  const int32_t* voff;       // i32[m + 1]: incidence of local variable i: entries ient[voff[i] .. voff[i+1])
  const int32_t* ient;       // entry index (into jcol / jval) of every incidence, ascending factor position
  const int32_t* irow;       // the row (factor position) of that entry
};

// ---- func ------------------------------------------------------------------------------------------
__global__ void lm_assign_kernel(GraphView G, const int32_t* vids, int m, const double* q) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    G.xbd[vids[i]] = make_double2(q[i], qnan);  // LMSSOpt::quickAssignVals: no clamping (:281-297)
    G.xval[vids[i]] = q[i];
  }
}

template <class Ops>
__global__ void __launch_bounds__(256) lm_func_kernel(GraphView G, LmDenseView L, double* partials /* [gridDim.x][2] */) {
  double sf = 0.0, s2 = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L.nf; k += gridDim.x * blockDim.x) {
    const int32_t fid = L.fids[k];
    double sl;
    double fv = Ops::template value<false>(G, fid, 0.0, false, sl);
    if (G.fconst_on != nullptr && G.fconst_on[fid]) fv = G.fconst_val[fid];
    const double h = sqrt(fv * 2.0);
    L.hx[k] = h;
    sf += fv;
    s2 += h * h;
  }
  __shared__ double wa[8], wb[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sf += __shfl_xor_sync(0xffffffffu, sf, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    wa[threadIdx.x >> 5] = sf;
    wb[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = wa[0], b = wb[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a += wa[w];
      b += wb[w];
    }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

__global__ void lm_negate_kernel(const double* hx, double* e, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) e[k] = -hx[k];
}

// ---- jacf ------------------------------------------------------------------------------------------
template <class Ops>
__global__ void __launch_bounds__(128) lm_rows_kernel(GraphView G, LmDenseView L) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L.nf; k += gridDim.x * blockDim.x) {
    const int32_t fid = L.fids[k];
    double* ge = G.gedge + Ops::edge_base(G, fid);
    double fv = Ops::gradient(G, fid, ge);
    if (G.fconst_on != nullptr && G.fconst_on[fid]) fv = G.fconst_val[fid];
    const double feval = sqrt(fv * 2.0);
    const int ar = Ops::arity(G, fid);
    const int32_t base = L.jptr[k];
    for (int s = 0; s < ar; ++s) {
      const int32_t li = L.vloc[slot_var(G, (Ops*)nullptr, fid, s)];
      L.jcol[base + s] = li;
      L.jval[base + s] = (li >= 0) ? ge[s] / feval : 0.0;
    }
  }
}

// ---- A = J^T J (lower), g = J^T e ----------------------------------------------------------------------
__global__ void __launch_bounds__(128) lm_assemble_kernel(LmDenseView L, double* A /* m x m, zeroed */, double* g) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp0; i < L.m; i += nwarps) {
    double gi = 0.0;
    double* Ai = A + (size_t)i * L.m;
    for (int t = L.voff[i]; t < L.voff[i + 1]; ++t) {
      const int32_t ent = L.ient[t], row = L.irow[t];
      const double jki = L.jval[ent];
      gi += jki * L.e[row];
      const int32_t r0 = L.jptr[row], r1 = L.jptr[row + 1];
      for (int32_t u = r0 + lane; u < r1; u += 32) {  // the row's entries: distinct columns, so lanes never collide
        const int32_t col = L.jcol[u];
        if (col >= 0 && col <= i) Ai[col] += jki * L.jval[u];
      }
      __syncwarp();
    }
    if (lane == 0) g[i] = gi;
  }
}

// scalars of one iteration: out[0] = max |g_i|, out[1] = sum p_i^2, out[2] = max A_ii; also diag[i] = A_ii
__global__ void __launch_bounds__(256) lm_scalars_kernel(const double* A, const double* g, const double* p, int m, double* diag, double* out) {
  double mg = 0.0, sp = 0.0, md = -1.7976931348623157e308;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double d = A[(size_t)i * m + i];
    diag[i] = d;
    const double t = fabs(g[i]);
    mg = (t > mg) ? t : mg;
    sp += p[i] * p[i];
    md = (d > md) ? d : md;
  }
  __shared__ double sa[8], sb[8], sc[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mg, o), c = __shfl_xor_sync(0xffffffffu, md, o);
    mg = (a > mg) ? a : mg;
    md = (c > md) ? c : md;
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sa[threadIdx.x >> 5] = mg;
    sb[threadIdx.x >> 5] = sp;
    sc[threadIdx.x >> 5] = md;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = sa[0], b = sb[0], c = sc[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a = (sa[w] > a) ? sa[w] : a;
      b += sb[w];
      c = (sc[w] > c) ? sc[w] : c;
    }
    out[0] = a;
    out[1] = b;
    out[2] = c;
  }
}

// L := lower(A) with mu added on the diagonal (the factorisation works in place on L)
__global__ void lm_augment_kernel(const double* A, double* Lm, int m, double mu) {
  const size_t n = (size_t)m * m;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / m), c = (int)(t % m);
    Lm[t] = (c < r) ? A[t] : ((c == r) ? A[t] + mu : 0.0);
  }
}

// ---- blocked Cholesky ----------------------------------------------------------------------------------
// Diagonal block [kb, kb+nb) in shared memory: factor it, and invert the factor (Linv = L11^-1, 64 x 64 row-major,
// zero above the diagonal and past nb) so that the panel solve below becomes a dense contraction.  *flag is set when a
// pivot is not positive.
__global__ void __launch_bounds__(256) lm_potrf_kernel(double* Lm, int m, int kb, double* Linv, int* flag) {
  // one 64 x 65 array holds both triangles: a[r][c], c <= r, is the factor; the strictly lower part of its inverse is kept
  // transposed in the strictly upper part (inverse(r, c) at a[c][r], c < r), its diagonal in dinv
  __shared__ double a[kLmNB][kLmNB + 1];
  __shared__ double dinv[kLmNB];
  const int nb = min(kLmNB, m - kb), t = threadIdx.x;
  for (int p = t; p < kLmNB * kLmNB; p += blockDim.x) {
    const int r = p / kLmNB, c = p % kLmNB;
    a[r][c] = (r < nb && c <= r) ? Lm[(size_t)(kb + r) * m + kb + c] : 0.0;
  }
  if (t < kLmNB) dinv[t] = 0.0;
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (t == 0) {
      const double d = a[j][j];
      if (!(d > 0.0)) *flag = 1;
      a[j][j] = sqrt(d);
    }
    __syncthreads();
    if (t > j && t < nb) a[t][j] = a[t][j] / a[j][j];
    __syncthreads();
    // the trailing (r, c) pairs, j < c <= r < nb: a 16 x 16 thread grid strides over them (no integer division)
    for (int r = j + 1 + (t >> 4); r < nb; r += 16) {
      const double arj = a[r][j];
      for (int c = j + 1 + (t & 15); c <= r; c += 16) a[r][c] -= arj * a[c][j];
    }
    __syncthreads();
  }
  for (int p = t; p < nb * nb; p += blockDim.x) {
    const int r = p / nb, c = p % nb;
    if (c <= r) Lm[(size_t)(kb + r) * m + kb + c] = a[r][c];
  }
  // inverse of the lower-triangular factor, column c by forward substitution, four lanes (one warp quarter) per column:
  // x_c = 1 / l_cc, x_r = -(sum_{k=c}^{r-1} l_rk x_k) / l_rr.  Column c is touched by its own four lanes only.
  {
    const int c = t >> 2, q = t & 3;
    if (q == 0 && c < nb) dinv[c] = 1.0 / a[c][c];
    __syncwarp();
    for (int r = 1; r < nb; ++r) {
      double s = 0.0;
      if (c < r && c < nb)
        for (int k = c + q; k < r; k += 4) s += a[r][k] * ((k == c) ? dinv[c] : a[c][k]);  // inverse(k, c)
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0 && c < r && c < nb) a[c][r] = -s / a[r][r];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int p = t; p < kLmNB * kLmNB; p += blockDim.x) {
    const int r = p / kLmNB, c = p % kLmNB;
    Linv[p] = (r >= nb || c > r) ? 0.0 : ((c == r) ? dinv[r] : a[c][r]);
  }
}

// C = P Q^T on the FP64 tensor cores, one CTA per 64 x 64 tile, 4 warps, warp (wi, wj) owns a 32 x 32 quarter = 4 x 4
// fragments of mma.sync.m8n8k4.f64 (SASS DMMA): A fragment = P rows (thread T: row T/4, k T%4), B fragment (col-major
// 4 x 8) = Q rows read the same way, C fragment: row T/4, columns 2 (T%4), 2 (T%4) + 1.  Panels are staged in shared memory
// 32 columns at a time.  Two uses, the two dense contractions of the factorisation:
//   kPanelSolve   L21 = A21 L11^-T: P = the 64 rows of A21 of this tile (read completely before anything is written),
//                 Q = Linv; the tile OVERWRITES A21.                                  grid (row tiles, 1)
//   kTrailing     C -= L21 L21^T on the lower triangle of the trailing matrix: P, Q = row tiles ti >= tj of L21.
//                                                                                     grid (row tiles, row tiles)
enum LmTileOp : int { kPanelSolve = 0, kTrailing = 1 };

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int kOp>
__global__ void __launch_bounds__(128) lm_tile_kernel(double* Lm, int m, int kb, int nb, const double* __restrict__ Linv) {
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (kOp == kTrailing && tj > ti) return;
  __shared__ double Ps[kLmNB][kLmTileLd], Qs[kLmNB][kLmTileLd];  // half a panel (32 columns) at a time: 36 KB
  const int r0 = kb + nb + ti * kLmNB;                             // first row of this tile's P rows in L
  const int c0 = (kOp == kTrailing) ? kb + nb + tj * kLmNB : kb;   // kTrailing: first row of the Q rows; kPanelSolve: first output column
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = warp >> 1, wj = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int kh = 0; kh < kLmNB; kh += kLmKH) {
    __syncthreads();
    for (int t = threadIdx.x; t < kLmNB * kLmKH; t += blockDim.x) {
      const int r = t / kLmKH, k = t % kLmKH;
      Ps[r][k] = (r0 + r < m && kh + k < nb) ? Lm[(size_t)(r0 + r) * m + kb + kh + k] : 0.0;
      if (kOp == kTrailing) Qs[r][k] = (c0 + r < m && kh + k < nb) ? Lm[(size_t)(c0 + r) * m + kb + kh + k] : 0.0;
      else Qs[r][k] = Linv[r * kLmNB + kh + k];
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < kLmKH; k0 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = Ps[wi * 32 + a * 8 + fr][k0 + fk];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = Qs[wj * 32 + b * 8 + fr][k0 + fk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = r0 + wi * 32 + a * 8 + fr;
    if (row >= m) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cl = wj * 32 + b * 8 + 2 * fk;  // column inside the tile
      double* dst = Lm + (size_t)row * m + c0 + cl;
      if (kOp == kTrailing) {
        if (c0 + cl <= row && c0 + cl < m) dst[0] -= acc[a][b][0];
        if (c0 + cl + 1 <= row && c0 + cl + 1 < m) dst[1] -= acc[a][b][1];
      } else {
        if (cl < nb) dst[0] = acc[a][b][0];
        if (cl + 1 < nb) dst[1] = acc[a][b][1];
      }
    }
  }
}

// L y = g, then L^T dp = y; one CTA.  Also pDp = p + dp, and out[0] = sum dp^2, out[1] = sum dp (mu dp + g).
__global__ void __launch_bounds__(1024) lm_trsv_kernel(const double* Lm, int m, const double* g, const double* p, double mu, double* y /* scratch m */,
                                                       double* dp, double* pDp, double* out) {
  __shared__ double xb[kLmNB];
  const int t = threadIdx.x;
  for (int i = t; i < m; i += blockDim.x) y[i] = g[i];
  __syncthreads();
  for (int kb = 0; kb < m; kb += kLmNB) {  // forward
    const int nb = min(kLmNB, m - kb);
    if (t < 32) {
      for (int j = 0; j < nb; ++j) {
        double yj = y[kb + j] / Lm[(size_t)(kb + j) * m + kb + j];
        if (t == 0) {
          y[kb + j] = yj;
          xb[j] = yj;
        }
        __syncwarp();
        for (int r = j + 1 + t; r < nb; r += 32) y[kb + r] -= Lm[(size_t)(kb + r) * m + kb + j] * yj;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = kb + nb + t; r < m; r += blockDim.x) {
      const double* row = Lm + (size_t)r * m + kb;
      double s = y[r];
      for (int c = 0; c < nb; ++c) s -= row[c] * xb[c];
      y[r] = s;
    }
    __syncthreads();
  }
  const int nblk = (m + kLmNB - 1) / kLmNB;
  for (int b = nblk - 1; b >= 0; --b) {  // backward: L^T dp = y
    const int kb = b * kLmNB, nb = min(kLmNB, m - kb);
    if (t < 32) {
      for (int j = nb - 1; j >= 0; --j) {
        double xj = y[kb + j] / Lm[(size_t)(kb + j) * m + kb + j];
        if (t == 0) {
          y[kb + j] = xj;
          xb[j] = xj;
        }
        __syncwarp();
        for (int r = t; r < j; r += 32) y[kb + r] -= Lm[(size_t)(kb + j) * m + kb + r] * xj;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = t; r < kb; r += blockDim.x) {
      double s = y[r];
      for (int c = 0; c < nb; ++c) s -= Lm[(size_t)(kb + c) * m + r] * xb[c];
      y[r] = s;
    }
    __syncthreads();
  }
  double s2 = 0.0, dl = 0.0;
  for (int i = t; i < m; i += blockDim.x) {
    const double d = y[i];
    dp[i] = d;
    pDp[i] = p[i] + d;
    s2 += d * d;
    dl += d * (mu * d + g[i]);
  }
  __shared__ double wa[32], wb[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    dl += __shfl_xor_sync(0xffffffffu, dl, o);
  }
  if ((t & 31) == 0) {
    wa[t >> 5] = s2;
    wb[t >> 5] = dl;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += wa[w];
      b += wb[w];
    }
    out[0] = a;
    out[1] = b;
  }
}

}  // namespace rdisgpu
