// solve_kernels.cuh — batched CGDSubspaceOptimizer::optimize on the device.
//
// One thread *group* owns one subspace problem for the whole solve (no host round trips):
//   Tile<G>   G = 1..32 lanes of a warp   — components with <= 32 factors (BA point blocks: 3 vars,
//                                           2..29 observations); 32/G problems share a warp
//   Block     one CTA                     — components with up to a few thousand factors
//                                           (BA camera blocks: 9 vars, 361..906 observations)
//   Grid      one cooperative grid        — a single large component (top-level blocks)
// The three differ only in how they synchronise and all-reduce; the solve itself
// (solve_problem) is written once.  Control flow is the CgdMachine of cgd_machine.cuh.
//
// Reference semantics implemented here:
//   CGDSubspaceOptimizer::optimize       src/optimizers/CGDSubspaceOptimizer.cpp:19-98
//   SubfunctionFD::operator() / df       CGD.cpp:124-157 (value = evalFactors over the list,
//                                        gradient = computeGradientOfSum restricted to vars)
//   Df1dim::operator() / df              external/include/minimize_nrc.h:432-447
//   vector updates of Frprmn / linmin    minimize_nrc.h:508-511, 637-641, 668-685
#pragma once
#include <cooperative_groups.h>

#include "cgd_machine.cuh"
#include "factors.cuh"

namespace rdisgpu {

namespace cg = cooperative_groups;

struct ProblemDesc {
  int64_t var_off;  // into vids / x0 / xout
  int64_t fac_off;  // into fids
  int32_t nv;
  int32_t nf;
  // resident NonlinearProductFactor class only (nlpf_resident.cuh), 0 otherwise: edges, distinct terms of the
  // problem's own variables, edges on frozen variables
  int32_t nE, nT, nFz, goff;  // goff: first slot of the problem's slice of BatchView::gscr
};

struct ResultRec {
  double f_init;
  double f_end;
  int32_t iters;
  int32_t status;
  int32_t n_value;
  int32_t n_slope;
};

struct BatchView {
  const ProblemDesc* probs;
  const int32_t* vids;
  const int32_t* fids;
  const double* x0;  // nullable
  double* xout;
  ResultRec* res;
  double* gscr;     // resident NLPF class: per-edge partial scratch, one slice per problem (nullable)
  uint16_t* gvinc;  // ... and the variable-major incidence lists, same slices
};

// ------------------------------------------------------------------------------------------
// groups
// ------------------------------------------------------------------------------------------
template <int G>
struct Tile {
  unsigned mask;
  int r;
  __device__ Tile() {
    const int lane = threadIdx.x & 31;
    r = lane & (G - 1);
    mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  }
  __device__ __forceinline__ int rank() const { return r; }
  __device__ __forceinline__ int size() const { return G; }
  __device__ __forceinline__ void sync() { __syncwarp(mask); }
  __device__ __forceinline__ void sum2(double& a, double& b) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      a += __shfl_xor_sync(mask, a, o);
      b += __shfl_xor_sync(mask, b, o);
    }
  }
  __device__ __forceinline__ void sum3max(double& a, double& b, double& c, double& mx) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      a += __shfl_xor_sync(mask, a, o);
      b += __shfl_xor_sync(mask, b, o);
      c += __shfl_xor_sync(mask, c, o);
      const double om = __shfl_xor_sync(mask, mx, o);
      mx = (om > mx) ? om : mx;
    }
  }
};

__device__ __forceinline__ void warp_sum4max(double& a, double& b, double& c, double& mx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    const double om = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (om > mx) ? om : mx;
  }
}

// One CTA.  smem: double[2][32][4] reduction scratch (double-buffered so one barrier per reduce).
struct Block {
  double* scratch;
  int flip;
  __device__ Block(double* s) : scratch(s), flip(0) {}
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void reduce(double& a, double& b, double& c, double& mx) {
    warp_sum4max(a, b, c, mx);
    double* buf = scratch + flip * 128;
    flip ^= 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) {
      buf[warp * 4 + 0] = a;
      buf[warp * 4 + 1] = b;
      buf[warp * 4 + 2] = c;
      buf[warp * 4 + 3] = mx;
    }
    __syncthreads();
    a = buf[0]; b = buf[1]; c = buf[2]; mx = buf[3];
    for (int wv = 1; wv < nw; ++wv) {  // fixed order: identical in every thread
      a += buf[wv * 4 + 0];
      b += buf[wv * 4 + 1];
      c += buf[wv * 4 + 2];
      const double om = buf[wv * 4 + 3];
      mx = (om > mx) ? om : mx;
    }
  }
  __device__ __forceinline__ void sum2(double& a, double& b) {
    double c = 0.0, mx = 0.0;
    reduce(a, b, c, mx);
  }
  __device__ __forceinline__ void sum3max(double& a, double& b, double& c, double& mx) { reduce(a, b, c, mx); }
};

// One cooperative grid.  partials: double[2][gridDim.x][4] in global memory.
struct Grid {
  cg::grid_group grid;
  Block blk;
  double* partials;
  int flip;
  __device__ Grid(double* smem_scratch, double* gpartials)
      : grid(cg::this_grid()), blk(smem_scratch), partials(gpartials), flip(0) {}
  __device__ __forceinline__ int rank() const { return blockIdx.x * blockDim.x + threadIdx.x; }
  __device__ __forceinline__ int size() const { return gridDim.x * blockDim.x; }
  __device__ __forceinline__ void sync() { grid.sync(); }
  __device__ __forceinline__ void reduce(double& a, double& b, double& c, double& mx) {
    blk.reduce(a, b, c, mx);  // every thread of the block now holds the block total
    double* buf = partials + (size_t)flip * gridDim.x * 4;
    flip ^= 1;
    if (threadIdx.x == 0) {
      buf[blockIdx.x * 4 + 0] = a;
      buf[blockIdx.x * 4 + 1] = b;
      buf[blockIdx.x * 4 + 2] = c;
      buf[blockIdx.x * 4 + 3] = mx;
    }
    grid.sync();
    // warp 0 of every block folds the per-block partials in the same fixed order
    double ta = 0.0, tb = 0.0, tc = 0.0, tm = 0.0;
    if (threadIdx.x < 32) {
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) {
        ta += buf[i * 4 + 0];
        tb += buf[i * 4 + 1];
        tc += buf[i * 4 + 2];
        const double om = buf[i * 4 + 3];
        tm = (om > tm) ? om : tm;
      }
      warp_sum4max(ta, tb, tc, tm);
      if (threadIdx.x == 0) {
        double* s = blk.scratch + 256;  // 4 doubles past the block scratch
        s[0] = ta; s[1] = tb; s[2] = tc; s[3] = tm;
      }
    }
    __syncthreads();
    const double* s = blk.scratch + 256;
    a = s[0]; b = s[1]; c = s[2]; mx = s[3];
    __syncthreads();
  }
  __device__ __forceinline__ void sum2(double& a, double& b) {
    double c = 0.0, mx = 0.0;
    reduce(a, b, c, mx);
  }
  __device__ __forceinline__ void sum3max(double& a, double& b, double& c, double& mx) { reduce(a, b, c, mx); }
};

// ------------------------------------------------------------------------------------------
// the solve
// ------------------------------------------------------------------------------------------
template <class Ops>
__device__ __forceinline__ double factor_value_or_const(const GraphView& G, int32_t fid, double alpha, bool along,
                                                        bool want_slope, double& slope) {
  const double fv = along ? Ops::template value<true>(G, fid, alpha, want_slope, slope)
                          : Ops::template value<false>(G, fid, alpha, want_slope, slope);
  if (G.fconst_on != nullptr && G.fconst_on[fid]) return G.fconst_val[fid];  // Factor::eval, src/Factor.cpp:110-119
  return fv;
}

// SubfunctionFD::operator() (+ Df1dim::df when want_slope) over the problem's factor list.
template <class Ops, class Grp>
__device__ __forceinline__ void objective_along_line(const GraphView& G, Grp& grp, const int32_t* fids, int nf,
                                                     double alpha, bool want_slope, double& f, double& slope) {
  double fs = 0.0, ss = 0.0;
  for (int k = grp.rank(); k < nf; k += grp.size()) {
    double sl;
    fs += factor_value_or_const<Ops>(G, fids[k], alpha, true, want_slope, sl);
    ss += sl;
  }
  grp.sum2(fs, ss);
  f = fs;
  slope = ss;
}

// Phase A of a full-gradient evaluation: every factor writes its partials to gedge.
template <class Ops, class Grp>
__device__ __forceinline__ double write_factor_partials(const GraphView& G, Grp& grp, const int32_t* fids, int nf) {
  double fs = 0.0;
  for (int k = grp.rank(); k < nf; k += grp.size()) {
    const int32_t fid = fids[k];
    double fv = Ops::gradient(G, fid, G.gedge + Ops::edge_base(G, fid));
    if (G.fconst_on != nullptr && G.fconst_on[fid]) fv = G.fconst_val[fid];
    fs += fv;
  }
  return fs;
}

// pidx < 0: this group has no problem (tail of a tile launch); it only answers the warp votes.
template <class Ops, class Grp>
__device__ void solve_problem(const GraphView& G, Grp& grp, const BatchView& B, int pidx, int maxiters, double ftol) {
  const bool active = (pidx >= 0);
  ProblemDesc P;
  P.var_off = 0; P.fac_off = 0; P.nv = 0; P.nf = 0;
  if (active) P = B.probs[pidx];
  const int32_t* vids = B.vids + P.var_off;
  const int32_t* fids = B.fids + P.fac_off;
  const int nv = P.nv, nf = P.nf;
  const int32_t stamp = pidx;
  const bool empty = active && (nf == 0);

  if (empty) {  // CGD.cpp:26-29: nothing to optimise, xval untouched
    for (int j = grp.rank(); j < nv; j += grp.size()) {
      const double xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : G.xbd[vids[j]].x;
      B.xout[P.var_off + j] = xv;
    }
    if (grp.rank() == 0) B.res[pidx] = ResultRec{0.0, 0.0, 0, ST_EMPTY, 0, 0};
  }
  const bool run = active && !empty;

  // ---- claim variables and factors ------------------------------------------------------
  if (run) {
    for (int j = grp.rank(); j < nv; j += grp.size()) {
      const int32_t vid = vids[j];
      const double xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : G.xbd[vid].x;
      G.xsave[vid] = xv;
      G.xbd[vid] = make_double2(xv, 0.0);  // direction 0: unfrozen, evaluated clamped (quickAssignVals, CGD.cpp:33)
    }
    for (int k = grp.rank(); k < nf; k += grp.size()) G.fstamp[fids[k]] = stamp;
    grp.sync();
  }

  CgdMachine m;
  m.start(maxiters, ftol);
  if (!run) m.req = REQ_DONE;
  double f_init = 0.0;

  while (true) {
    const bool fin = m.done();
    if (__all_sync(0xffffffffu, fin)) break;
    if (fin) continue;

    if (m.req == REQ_INIT_GRAD) {
      // initialFval = sfd(xval) (CGD.cpp:37) and Frprmn's fp / first gradient (:634-641)
      double fs = write_factor_partials<Ops>(G, grp, fids, nf);
      double zero = 0.0;
      grp.sum2(fs, zero);  // also orders the gedge writes before the gather below
      grp.sync();
      for (int j = grp.rank(); j < nv; j += grp.size()) {
        const int32_t vid = vids[j];
        const double gr = Ops::gather_var(G, vid, stamp, true);
        const double gneg = -gr;
        G.gvec[vid] = gneg;
        G.hvec[vid] = gneg;
        G.xbd[vid].y = gneg;
      }
      grp.sync();
      f_init = fs;
      m.on_init(fs);
    } else {
      double f, sl;
      objective_along_line<Ops>(G, grp, fids, nf, m.alpha, m.req == REQ_VALUE_SLOPE, f, sl);
      m.on_eval(f, sl);
    }

    // bookkeeping requests between two line evaluations
    while (m.req == REQ_MOVE) {
      const double step = m.alpha;
      for (int j = grp.rank(); j < nv; j += grp.size()) {  // minimize_nrc.h:508-511
        const int32_t vid = vids[j];
        double2 xb = G.xbd[vid];
        xb.y *= step;
        xb.x += xb.y;
        G.xbd[vid] = xb;
      }
      grp.sync();
      m.on_moved();
      if (m.req != REQ_GRADIENT) break;

      double fs = write_factor_partials<Ops>(G, grp, fids, nf);
      (void)fs;
      grp.sync();
      double gg = 0.0, dgg = 0.0, dummy = 0.0, tnum = 0.0;
      for (int j = grp.rank(); j < nv; j += grp.size()) {
        const int32_t vid = vids[j];
        const double gr = Ops::gather_var(G, vid, stamp, true);
        double2 xb = G.xbd[vid];
        xb.y = gr;  // func.df(p, xi), :654
        G.xbd[vid] = xb;
        const double pj = fabs(xb.x);
        const double t = fabs(gr) * ((pj < 1.0) ? 1.0 : pj);  // :659 numerator
        tnum = (t > tnum) ? t : tnum;
        const double gj = G.gvec[vid];
        gg += gj * gj;
        dgg += (gr + gj) * gr;
      }
      grp.sum3max(gg, dgg, dummy, tnum);
      m.on_gradient(tnum, gg, dgg);
      if (m.req != REQ_DIRECTION) break;
      const double gam = m.gam;
      for (int j = grp.rank(); j < nv; j += grp.size()) {  // :681-685
        const int32_t vid = vids[j];
        const double gj = -G.xbd[vid].y;
        const double hj = gj + gam * G.hvec[vid];
        G.gvec[vid] = gj;
        G.hvec[vid] = hj;
        G.xbd[vid].y = hj;
      }
      grp.sync();
      m.on_directed();
    }
  }

  if (!run) return;
  // ---- commit (CGD.cpp:61-89) --------------------------------------------------------------
  double fret = m.fret;
  bool restore = (fret > f_init);
  // the safety exits keep p / fret of the last completed line search, like the reference's own throws (CGD.cpp:41-61)
  if (restore) fret = f_init;
  for (int j = grp.rank(); j < nv; j += grp.size()) {
    const int32_t vid = vids[j];
    const double raw = restore ? G.xsave[vid] : G.xbd[vid].x;
    const double val = clamp_to_domain(raw, G.dom[vid]);
    G.xbd[vid] = make_double2(val, __longlong_as_double(0x7ff8000000000000LL));  // freeze again
    G.xval[vid] = val;
    B.xout[P.var_off + j] = val;
  }
  for (int k = grp.rank(); k < nf; k += grp.size()) G.fstamp[fids[k]] = -1;
  if (grp.rank() == 0) {
    ResultRec r;
    r.f_init = f_init;
    r.f_end = fret;
    r.iters = m.iter;
    r.status = m.status;
    r.n_value = m.n_value;
    r.n_slope = m.n_slope;
    B.res[pidx] = r;
  }
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// `order` lists the problem indices of this size class; 32/G problems per warp.
template <class Ops, int G>
__global__ void __launch_bounds__(128) solve_tile_kernel(GraphView Gv, BatchView B, const int32_t* order, int count,
                                                         int maxiters, double ftol) {
  constexpr int per_block = 128 / G;
  const int slot = blockIdx.x * per_block + threadIdx.x / G;
  Tile<G> grp;
  solve_problem<Ops>(Gv, grp, B, (slot < count) ? order[slot] : -1, maxiters, ftol);
}

// *out += sum of f_end over the batch (one block, fixed order): the per-GPU partial of the global
// objective that is all-reduced across ranks.
__global__ void sum_f_end_kernel(const ResultRec* res, int64_t n, double* out) {
  __shared__ double wsum[8];
  double t = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) t += res[i].f_end;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt = wsum[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) tt += wsum[w];
    *out += tt;
  }
}

#ifndef RDIS_BLOCK_MIN_CTAS
#define RDIS_BLOCK_MIN_CTAS 1
#endif
template <class Ops>
__global__ void __launch_bounds__(256, RDIS_BLOCK_MIN_CTAS) solve_block_kernel(GraphView Gv, BatchView B, const int32_t* order, int count,
                                                             int maxiters, double ftol) {
  __shared__ double scratch[260];
  if ((int)blockIdx.x >= count) return;
  Block grp(scratch);
  solve_problem<Ops>(Gv, grp, B, order[blockIdx.x], maxiters, ftol);
}

template <class Ops>
__global__ void solve_grid_kernel(GraphView Gv, BatchView B, int pidx, double* partials, int maxiters, double ftol) {
  __shared__ double scratch[260];
  Grid grp(scratch, partials);
  solve_problem<Ops>(Gv, grp, B, pidx, maxiters, ftol);
}

}  // namespace rdisgpu
