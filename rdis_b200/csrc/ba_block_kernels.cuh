// ba_block_kernels.cuh — register / shared-memory resident subspace solves for the two component
// shapes the recursive decomposer produces on a bundle-adjustment graph, where whole variable
// blocks are chosen together (src/RDISOptimizer.cpp:464-487):
//
//   point block    vars = the 3 coordinates of ONE point, factors = (some of) its observations,
//                  at most 32 of them (ladybug: 2..29).          -> solve_ba_points_kernel
//   camera block   vars = the 9 parameters of ONE camera, factors = (some of) its observations
//                  (ladybug: 361..906).                          -> solve_ba_cameras_kernel
//
// What is staged where for the life of a solve (nothing but the final commit touches HBM):
//   point block    one lane per observation (G = 1..32 lanes per problem, 32/G problems per warp).
//                  The lane keeps its observation's FROZEN variable block in registers — the
//                  camera's rotation already reduced to axis / angle / sin / cos, translation,
//                  intrinsics, the pixel — and every lane keeps a replica of the problem's own
//                  state (p, xi, g, h, domain: 3 each).  Reductions are warp shuffles.
//   camera block   a thread-block CLUSTER of C CTAs on C SMs per problem (C*T threads >= #observations
//                  where possible): a thread keeps its observation's frozen point + pixel in
//                  registers, every CTA keeps a replica of the camera state (p, xi, g, h, domain:
//                  9 each) in shared memory, and the per-evaluation all-reduce goes warp shuffle ->
//                  distributed-shared-memory stores into every CTA of the cluster -> one cluster
//                  barrier -> fixed-order fold.  The FP64 pipe of one SM (measured 58 DFMA/clk)
//                  is the per-evaluation bound of a camera block, hence the spread over C SMs.
//
// Semantics are those of solve_problem (solve_kernels.cuh) — same CgdMachine, same factor
// arithmetic (BaOps), same commit rules.  The point kernel folds objective sums with the same shuffle
// trees as the generic tile path and the gradient in ascending factor id; the camera kernel folds its
// sums in a different (tree) order.  Both agree with the generic path to rounding (FMA contraction
// is decided per compilation context, so not to the bit).
// Eligibility is decided on the host (rdis_gpu.cu: classify_ba_blocks); everything else goes
// through the generic kernels.
#pragma once
#include <cooperative_groups.h>

#include "ptx_async.cuh"
#include "solve_kernels.cuh"

namespace rdisgpu {

__device__ __forceinline__ double qnan_f64() { return __longlong_as_double(0x7ff8000000000000LL); }

// ------------------------------------------------------------------------------------------
// point blocks
// ------------------------------------------------------------------------------------------
// G (lanes per problem) is a run-time value: every warp of the launch runs the same code whatever its
// size class, which keeps the instruction working set to one copy of the solve loop.
struct TileRt {
  unsigned mask;
  int r, G;
  __device__ TileRt(int lg) {
    const int lane = threadIdx.x & 31;
    G = 1 << lg;
    r = lane & (G - 1);
    mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  }
  __device__ __forceinline__ void sum2(double& a, double& b) const {
    for (int o = G >> 1; o > 0; o >>= 1) {
      a += __shfl_xor_sync(mask, a, o);
      b += __shfl_xor_sync(mask, b, o);
    }
  }
  __device__ __forceinline__ void sum2max(double& a, double& b, double& mx) const {
    for (int o = G >> 1; o > 0; o >>= 1) {
      a += __shfl_xor_sync(mask, a, o);
      b += __shfl_xor_sync(mask, b, o);
      const double om = __shfl_xor_sync(mask, mx, o);
      mx = (om > mx) ? om : mx;
    }
  }
};

// The reference's value cache as seen by ONE block problem: every factor of the problem depends on all of the
// problem's variables, so Variable::assign's change filter (a move below 1e-12 notifies nobody, src/Variable.cpp:66-88)
// and Factor::eval's cached value (src/Factor.cpp:110-119) collapse to one dirty flag and one cached SUM per problem
// (the factors are all recomputed at the same events, and their values are summed in the same order every time).
// A solve starts with everything dirty (the strict kernels also carry the cache across calls).
template <int NV>
struct BlockCache {
  double last[NV];  // Variable::eval(): the value of the last assign
  double cached;    // sum of the factors' cached values
  bool dirty, assigned;
  __device__ __forceinline__ void reset() {
    dirty = true;
    assigned = false;
    cached = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) last[j] = 0.0;
  }
  // Variable::assign of the clamped point
  __device__ __forceinline__ void assign(const double* x) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (!assigned || !(fabs(x[j] - last[j]) < 1e-12)) dirty = true;
      last[j] = x[j];
    }
    assigned = true;
  }
  // evalFactors over the list: `fresh` is the sum just computed at the assigned point
  __device__ __forceinline__ double eval(double fresh) {
    if (dirty) {
      cached = fresh;
      dirty = false;
    }
    return cached;
  }
};

// Returns the number of evaluation passes the WARP ran (the slowest block of the warp decides).
__device__ __forceinline__ int solve_ba_point_tile(const GraphView& Gv, const BatchView& B, int lg, int pidx, int maxiters,
                                                   double ftol) {
  const TileRt grp(lg);
  const int r = grp.r;
  const int G = grp.G;
  const bool run = (pidx >= 0);
  ProblemDesc P;
  P.var_off = 0; P.fac_off = 0; P.nv = 0; P.nf = 0;
  if (run) P = B.probs[pidx];
  const int nf = P.nf;
  const int32_t v0 = run ? B.vids[P.var_off] : 0;

  // ---- the problem's own state, replicated in every lane of the tile ----
  double p[3], xi[3], g[3], h[3], xs[3];
  double2 dom[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double xv = 0.0;
    if (run) xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : Gv.xbd[v0 + j].x;
    p[j] = xs[j] = xv;
    xi[j] = 0.0;  // direction 0: evaluated clamped (quickAssignVals, CGD.cpp:33)
    g[j] = h[j] = 0.0;
    dom[j] = run ? __ldg(&Gv.dom[v0 + j]) : make_double2(0.0, 0.0);
  }

  // ---- this lane's observation: frozen camera block staged once ----
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
    for (int s = 0; s < 9; ++s) x[s] = Gv.xbd[9 * cam + s].x;  // frozen: read as stored, never clamped
    BaOps::rotation(x[0], x[1], x[2], m);
    if (Gv.fconst_on != nullptr && Gv.fconst_on[fid]) {
      fc_on = true;
      fc_val = Gv.fconst_val[fid];
    }
  }

  CgdMachine mc;
  mc.start(maxiters, ftol, /*faithful=*/true);
  if (!run) mc.req = REQ_DONE;
  double f_init = 0.0;
  BlockCache<3> cache;
  cache.reset();
  double f_at_p = 0.0;  // objective at clamp(p), recomputed by every gradient pass: serves fa = func(0) (minimize_nrc.h:88)

  // Every iteration of this loop is one objective evaluation for every unfinished problem of the
  // warp.  The evaluation and all shuffles run CONVERGED with the full-warp mask (indexed shuffles below width G
  // never leave a tile), whatever phase each problem's state machine is in; only the scalar state-machine step at
  // the end diverges between the tiles of a warp.
  //
  // Every sum is folded in the REFERENCE's order — the factor values left to right in list order
  // (OptimizableFunction.cpp:108-132), each variable's derivative over its factors in ascending factor id, the first
  // one copied (State.h:157-194), the directional derivative left to right over the variables (minimize_nrc.h:444) —
  // so, with the library built without FMA contraction and the correctly rounded quotients of BaOps::partials, a point
  // block is solved to the same BITS as the reference arithmetic produces (tests: the full real ladybug wave).
  const unsigned kFull = 0xffffffffu;
#ifdef RDIS_PT_PROFILE  // cycle accounting of one evaluation (clock64; a tuning build, never shipped)
  long long pprof[4] = {0, 0, 0, 0}, plast = clock64();
  int pnev = 0;
#define PTPROF(i) do { const long long t_ = clock64(); pprof[i] += t_ - plast; plast = t_; } while (0)
#else
#define PTPROF(i) do { } while (0)
#endif
  int passes = 0;
  while (true) {
    const bool fin = mc.done();
    if (__all_sync(kFull, fin)) break;
    ++passes;
#ifdef RDIS_PT_PROFILE
    ++pnev;
#endif
    const int kind = fin ? (int)REQ_DONE : mc.req;  // REQ_INIT_GRAD, REQ_VALUE, REQ_VALUE_SLOPE or REQ_GRADIENT
    const bool along = (kind == REQ_VALUE) || (kind == REQ_VALUE_SLOPE);
    const bool grad_kind = (kind == REQ_INIT_GRAD) || (kind == REQ_GRADIENT);
    const bool want_g = (kind == REQ_VALUE_SLOPE) || grad_kind;
    const bool live = have && !fin;
    const double alpha = mc.alpha;
    double fv = 0.0, g9 = 0.0, g10 = 0.0, g11 = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) x[9 + j] = clamp_to_domain(along ? (p[j] + alpha * xi[j]) : p[j], dom[j]);
    if (live) {
      fv = BaOps::project(x, ob, m);
      if (want_g) {
        double gq[12];
        BaOps::partials(x, m, gq);
        g9 = gq[9]; g10 = gq[10]; g11 = gq[11];
      }
      if (fc_on) fv = fc_val;  // Factor::eval of an assigned-constant factor, src/Factor.cpp:110-119
    }

    PTPROF(0);  // vote, clamp, the observation's value / partials
    // left-to-right folds over the tile's lanes (= the problem's factor list), replicated in every lane
    double fs = 0.0;
    double gr[3] = {0.0, 0.0, 0.0};
    const bool any_g = __any_sync(kFull, want_g);
    // The broadcasts do not depend on the running sums: the loops are unrolled so that the shuffles of the next lanes
    // are in flight while the additions of the previous ones retire (the chain is then the additions alone).
    if (any_g) {
#pragma unroll 4
      for (int k = 0; k < G; ++k) {
        const double bf = __shfl_sync(kFull, fv, k, G);
        const double b0 = __shfl_sync(kFull, g9, k, G), b1 = __shfl_sync(kFull, g10, k, G), b2 = __shfl_sync(kFull, g11, k, G);
        if (k < nf) fs = fs + bf;
        if (k == 0) {
          gr[0] = b0; gr[1] = b1; gr[2] = b2;
        } else if (k < nf) {
          gr[0] = gr[0] + b0; gr[1] = gr[1] + b1; gr[2] = gr[2] + b2;
        }
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < G; ++k) {
        const double bf = __shfl_sync(kFull, fv, k, G);
        if (k < nf) fs = fs + bf;
      }
    }

    PTPROF(1);  // ordered folds
    // ---- the state machine's step: scalar work, the only part where tiles of a warp diverge ----
    if (!fin) {
      cache.assign(&x[9]);  // SubfunctionFD::quickAssignVals of this evaluation's point
      if (along) {
        const double fcached = cache.eval(fs);
        double ss = 0.0;  // Df1dim::df: df1 += dft[j] * xi[j]
        if (kind == REQ_VALUE_SLOPE) ss = ((0.0 + gr[0] * xi[0]) + gr[1] * xi[1]) + gr[2] * xi[2];
        mc.on_eval(fcached, ss);
        if (mc.req == REQ_MOVE) {  // minimize_nrc.h:508-511
          const double step = mc.alpha;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            xi[j] *= step;
            p[j] += xi[j];
          }
          mc.on_moved();
        }
      } else if (kind == REQ_INIT_GRAD) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double gneg = -gr[j];
          g[j] = gneg; h[j] = gneg; xi[j] = gneg;
        }
        f_init = cache.eval(fs);
        f_at_p = fs;
        mc.on_init(f_init);
      } else {  // REQ_GRADIENT: func.df(p, xi) — assigns, no Factor::eval
        f_at_p = fs;
        double gg = 0.0, dgg = 0.0, tnum = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          xi[j] = gr[j];
          const double pj = fabs(p[j]);
          const double t = fabs(gr[j]) * ((pj < 1.0) ? 1.0 : pj);
          tnum = (t > tnum) ? t : tnum;
          gg += g[j] * g[j];
          dgg += (gr[j] + g[j]) * gr[j];
        }
        mc.on_gradient(tnum, gg, dgg);
        if (mc.req == REQ_DIRECTION) {  // :681-685
          const double gam = mc.gam;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double gj = -xi[j];
            const double hj = gj + gam * h[j];
            g[j] = gj; h[j] = hj; xi[j] = hj;
          }
          mc.on_directed();
        }
      }
      // fa = func(ax = 0) at the top of a line search: the point is clamp(p), already assigned by the gradient
      // pass that preceded it, and f_at_p is its objective — answered without another pass
      if (mc.req == REQ_VALUE && mc.phase == CgdMachine::PH_BR_FA) mc.on_eval(cache.eval(f_at_p), 0.0);
    }
    __syncwarp();
    PTPROF(2);  // machine step (tiles diverge)
  }
#ifdef RDIS_PT_PROFILE
  if ((threadIdx.x & 31) == 0 && pnev >= 300)
    printf("ptprof G %d passes %d | eval %lld fold %lld machine %lld cycles per pass\n", G, pnev, pprof[0] / pnev, pprof[1] / pnev, pprof[2] / pnev);
#endif

  if (!run) return passes;
  // ---- commit (CGD.cpp:61-89): quickAssignVals(gdmin.p); if worse than the start, the start is re-assigned and
  //      re-evaluated (through the cache).  The safety exits (non-finite abscissa, bracket cap) keep p and fret of the
  //      last completed line search like the reference's own exceptions do. ----
  double xfin[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) xfin[j] = clamp_to_domain(p[j], dom[j]);
  cache.assign(xfin);
  double fret = mc.fret;
  const bool restore = (fret > f_init);
  if (restore) {
#pragma unroll
    for (int j = 0; j < 3; ++j) xfin[j] = clamp_to_domain(xs[j], dom[j]);
    cache.assign(xfin);
    fret = cache.eval(f_init);  // a recomputation at the start point reproduces f_init's bits
  }
  if (r == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Gv.xbd[v0 + j] = make_double2(xfin[j], qnan_f64());
      Gv.xval[v0 + j] = xfin[j];
      B.xout[P.var_off + j] = xfin[j];
    }
    ResultRec res;
    res.f_init = f_init;
    res.f_end = fret;
    res.iters = mc.iter;
    res.status = mc.status;
    res.n_value = mc.n_value;
    res.n_slope = mc.n_slope;
    B.res[pidx] = res;
  }
  return passes;
}

// One warp per CTA.  warp_task[w] = {log2 G, first slot in `order`, number of problems of that
// class still to hand out from that slot on}: every size class runs in the SAME launch, so the
// launch lasts as long as the slowest single problem, not the sum of the per-class tails.
struct PointWarpTask {
  int32_t lg;
  int32_t first;
  int32_t count;
  float key;  // written by the solve: predicted duration of this warp at the next visit (pt_reorder_kernel sorts by it)
};

// ------------------------------------------------------------------------------------------
// Launch order from the previous visit.  Both block kernels are ONE wave of long dependent chains whose length (the
// number of line-search evaluations) differs 8x between problems, and neither fits the machine entirely: 49 clusters of
// 8 CTAs place 6 per GPC, so the last one starts when the first one ends; the point kernel has 1.2x more warps than are
// resident.  What is dispatched last must therefore be SHORT.  A sibling set is re-posed many times by the tree search
// (alternating minimisation, src/RDISOptimizer.cpp:1148-1181) with start values that move little, so the evaluation
// counts of the previous solve (ResultRec; a point warp records its own pass count in its task) predict the next one:
// after every solve the clusters / warp tasks are re-sorted, longest first (rank sort in one CTA, stream-ordered, no
// host involvement).  Results do not depend on the
// order (every problem is solved by its own cluster / tile).  Measured on ladybug's 49 cameras: 3.14 ms in camera-id
// order, 2.77 ms longest-first.
// ------------------------------------------------------------------------------------------
constexpr int kReorderMax = 2048;

__global__ void __launch_bounds__(1024) cam_reorder_kernel(int32_t* __restrict__ order, int n, const ResultRec* __restrict__ res) {
  __shared__ int32_t item[kReorderMax];
  __shared__ float key[kReorderMax];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int32_t p = order[i];
    item[i] = p;
    // a value-only evaluation costs ~3600 cycles, one with the gradient ~5500 (cycle accounting in DESIGN.md)
    key[i] = 36.0f * (float)res[p].n_value + 55.0f * (float)res[p].n_slope;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float k = key[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (key[j] > k || (key[j] == k && j < i)) ? 1 : 0;
    order[rank] = item[i];
  }
}

// Rank sort over several CTAs: a CTA ranks kReorderItems tasks against all n keys (four threads per task, each a quarter
// of the keys, interleaved so that a warp's four shared-memory addresses are neighbours) and writes them to their places
// in the OTHER list (the host swaps the two lists after the launch).
constexpr int kReorderItems = 64;
__global__ void __launch_bounds__(4 * kReorderItems) pt_reorder_kernel(const PointWarpTask* __restrict__ src, PointWarpTask* __restrict__ dst,
                                                                       int n) {
  __shared__ float key[kReorderMax + 4];
  for (int i = threadIdx.x; i < n; i += blockDim.x) key[i] = src[i].key;
  __syncthreads();
  const int i = blockIdx.x * kReorderItems + ((int)threadIdx.x >> 2), q = (int)threadIdx.x & 3;
  const float k = (i < n) ? key[i] : 0.0f;
  int rank = 0;
#pragma unroll 4
  for (int j = q; j < n; j += 4) {
    const float kj = key[j];
    rank += (kj > k || (kj == k && j < i)) ? 1 : 0;
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  if (q == 0 && i < n) dst[rank] = src[i];
}

#ifndef RDIS_PT_MIN_CTAS
#define RDIS_PT_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(32, RDIS_PT_MIN_CTAS) solve_ba_points_kernel(GraphView Gv, BatchView B, const int32_t* order,
                                                             PointWarpTask* tasks, int maxiters, double ftol) {
  const PointWarpTask t = tasks[blockIdx.x];
  const int slot = (int)threadIdx.x >> t.lg;
  const int passes = solve_ba_point_tile(Gv, B, t.lg, (slot < t.count) ? order[t.first + slot] : -1, maxiters, ftol);
  // the machine steps of the blocks of a warp serialise where they diverge: more blocks, longer passes
  if (threadIdx.x == 0) tasks[blockIdx.x].key = (float)passes * (1.0f + 0.05f * (float)t.count);
}

// ------------------------------------------------------------------------------------------
// camera blocks
// ------------------------------------------------------------------------------------------
#ifndef RDIS_CAM_THREADS
#define RDIS_CAM_THREADS 192
#endif
#ifndef RDIS_CAM_MAXREG
#define RDIS_CAM_MAXREG 128  // 3 CTAs of 128 + 32 threads per SM (measured: 3.28 ms per ladybug camera wave against 3.64 at 168)
#endif
constexpr int kCamMaxThreads = RDIS_CAM_THREADS;          // WORKER threads of a CTA (one more warp runs the solve's scalar side)
constexpr int kCamMaxCluster = 8;
constexpr int kCamMaxWorkerWarps = kCamMaxThreads / 32;
constexpr int kCamMaxRows = kCamMaxCluster * kCamMaxWorkerWarps;  // warp partials of a whole cluster
constexpr int kCamRedWidth = 10;  // f + 9 partials (gradient evaluations); f + slope use the first two

// What the scalar warp publishes for one evaluation and the worker warps read.
struct CamRequest {
  int kind, pad;
  double x[9];                      // the evaluation's point, clamped into the domain (quickAssignVals)
  double a0, a1, a2, theta, s, c;   // BaOps::rotation of it: computed once per evaluation, not once per observation
  double xi[9];                     // the search direction (directional derivative of slope evaluations)
};

struct CamShared {
  CamRequest rq;
  alignas(16) double red[2][kCamMaxRows][kCamRedWidth];  // (16-byte pairs travel by st.async.v2) warp partials of the whole cluster, indexed by the warp's rank in the cluster; double-buffered
  unsigned long long mbar[2];                // one transaction barrier per buffer (the peers' st.async complete on it)
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// grid = nprobs * C CTAs, cluster = C CTAs (set at launch), blockDim = T worker threads + one scalar warp, C*T >= the
// largest factor list of the batch (one observation per worker thread); `order` lists the camera-class problems.
//
// WARP-SPECIALISED.  The solve is a chain of <= ~1200 dependent objective evaluations; what can be shortened is what
// happens between two of them, so the scalar side never shares a thread with the observation arithmetic:
//   scalar warp (the last warp of every CTA; identical in every CTA of the cluster on identical inputs)
//       lane 0 keeps the CgdMachine and the value cache (BlockCache semantics), lane j < 9 variable j of the camera
//       (p, xi, g, h, domain) in REGISTERS for the whole solve; per evaluation it forms the clamped point lane-parallel
//       and its rotation (sqrt, 3 divisions, sincos: once, not once per observation), publishes the request through
//       shared memory and releases the workers with a named barrier; then waits on the buffer's transaction barrier,
//       folds the cluster's warp partials with a lane-parallel fixed tree and steps.
//   worker warps: read the request, evaluate their observation (BaOps::project / partials, the operands of the frozen
//       point block staged in registers for the life of the solve), butterfly-reduce inside the warp, and lanes < C
//       push the warp's partial into the slot [warp's rank in the cluster] of every CTA with ASYNCHRONOUS stores over
//       distributed shared memory that complete the receiving CTA's transaction barrier (st.async ... complete_tx):
//       one one-way trip, no CTA barrier and no cluster barrier on the way.
// The fold order is a function of the observation index only (warp partial = butterfly over 32 consecutive
// observations; partials folded by rank mod 4 into four chains, chains combined pairwise), so the result does not
// depend on the cluster shape the launch picked.
__global__ void __maxnreg__(RDIS_CAM_MAXREG) solve_ba_cameras_kernel(GraphView Gv, BatchView B, const int32_t* order,
                                                                               int C, int maxiters, double ftol) {
  namespace cgn = cooperative_groups;
  cgn::cluster_group cluster = cgn::this_cluster();
  __shared__ CamShared sh;
  const int cta = (C > 1) ? (int)cluster.block_rank() : 0;
  const int pidx = order[blockIdx.x / C];
  const ProblemDesc P = B.probs[pidx];
  const int nf = P.nf;
  const int32_t v0 = B.vids[P.var_off];
  const int T = (int)blockDim.x - 32;        // worker threads
  const int W = T >> 5;                       // worker warps per CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool scalar_warp = (warp == W);
  const int nthreads = (int)blockDim.x;
  if (threadIdx.x == 0) {
    mbar_init(&sh.mbar[0], 1);
    mbar_init(&sh.mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (C > 1) cluster.sync();  // every CTA's barriers are initialised before a peer's st.async can reach them

  if (!scalar_warp) {
    // ------------------------------------------------------------------ workers
    const int rank = cta * T + threadIdx.x;   // = index of this thread's observation in the problem's factor list
    const int grow = cta * W + warp;          // this warp's rank in the cluster = its row of the reduction buffers
    const bool have = (rank < nf);
    double x[12];
    double2 ob = make_double2(0.0, 0.0);
    bool fc_on = false;
    double fc_val = 0.0;
    x[9] = x[10] = x[11] = 0.0;
    if (have) {  // the frozen point block + pixel: registers for the life of the solve
      const int32_t pbase = 9 * Gv.ncams;
      const int32_t fid = B.fids[P.fac_off + rank];
      const int32_t pt = __ldg(&Gv.pt[fid]);
      ob = __ldg(&Gv.obs[fid]);
      x[9] = Gv.xbd[pbase + 3 * pt].x; x[10] = Gv.xbd[pbase + 3 * pt + 1].x; x[11] = Gv.xbd[pbase + 3 * pt + 2].x;
      if (Gv.fconst_on != nullptr && Gv.fconst_on[fid]) {
        fc_on = true;
        fc_val = Gv.fconst_val[fid];
      }
    }
    int flip = 0;
#ifdef RDIS_CAM_PROFILE
    long long wp[4] = {0, 0, 0, 0}, wl = clock64();
    int wn = 0;
#define WKPROF(i) do { const long long t_ = clock64(); wp[i] += t_ - wl; wl = t_; } while (0)
#else
#define WKPROF(i) do { } while (0)
#endif
    while (true) {
      named_bar_sync(1, nthreads);  // the request is published
      WKPROF(0);  // waiting for the request
      const int kind = sh.rq.kind;
      if (kind == REQ_DONE) break;
#ifdef RDIS_CAM_PROFILE
      ++wn;
#endif
      BaOps::Fwd m;
#pragma unroll
      for (int j = 0; j < 9; ++j) x[j] = sh.rq.x[j];
      m.a0 = sh.rq.a0; m.a1 = sh.rq.a1; m.a2 = sh.rq.a2; m.theta = sh.rq.theta; m.s = sh.rq.s; m.c = sh.rq.c;
      const uint32_t bar = smem_addr_u32(&sh.mbar[flip]);
      if (kind == REQ_VALUE || kind == REQ_VALUE_SLOPE) {
        double v0s = 0.0, v1s = 0.0;
        if (have) {
          double fv = BaOps::project(x, ob, m);
          if (kind == REQ_VALUE_SLOPE) {
            double gq[12];
            BaOps::partials(x, m, gq);
            double sl = 0.0;
#pragma unroll
            for (int s = 0; s < 9; ++s) sl += gq[s] * sh.rq.xi[s];
            v1s = sl;
          }
          if (fc_on) fv = fc_val;  // Factor::eval of an assigned-constant factor, src/Factor.cpp:110-119
          v0s = fv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v0s += __shfl_xor_sync(0xffffffffu, v0s, o);
          v1s += __shfl_xor_sync(0xffffffffu, v1s, o);
        }
        WKPROF(1);  // evaluation + butterfly
        if (lane < C) {
          const uint32_t dst = (C > 1) ? map_to_cta(smem_addr_u32(&sh.red[flip][grow][0]), (uint32_t)lane) : smem_addr_u32(&sh.red[flip][grow][0]);
          const uint32_t rbar = (C > 1) ? map_to_cta(bar, (uint32_t)lane) : bar;
          st_async_v2(dst, v0s, v1s, rbar);
        }
        WKPROF(2);  // send
      } else {
        double acc[kCamRedWidth];
#pragma unroll
        for (int i = 0; i < kCamRedWidth; ++i) acc[i] = 0.0;
        if (have) {
          double fv = BaOps::project(x, ob, m);
          double gq[12];
          BaOps::partials(x, m, gq);
#pragma unroll
          for (int s = 0; s < 9; ++s) acc[1 + s] = gq[s];
          if (fc_on) fv = fc_val;
          acc[0] = fv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < kCamRedWidth; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        if (lane < C) {
          const uint32_t dst = (C > 1) ? map_to_cta(smem_addr_u32(&sh.red[flip][grow][0]), (uint32_t)lane) : smem_addr_u32(&sh.red[flip][grow][0]);
          const uint32_t rbar = (C > 1) ? map_to_cta(bar, (uint32_t)lane) : bar;
#pragma unroll
          for (int i = 0; i < kCamRedWidth; i += 2) st_async_v2(dst + 8u * i, acc[i], acc[i + 1], rbar);
        }
      }
      flip ^= 1;
    }
#ifdef RDIS_CAM_PROFILE
    if (threadIdx.x == 0 && cta == 0 && (blockIdx.x / C) < 4)
      printf("camprof worker prob %d evals %d | wait-for-request %lld eval+butterfly %lld send %lld\n", pidx, wn, wp[0] / (wn ? wn : 1), wp[1] / (wn ? wn : 1), wp[2] / (wn ? wn : 1));
#endif
  } else {
    // ------------------------------------------------------------------ the scalar warp
    // lane j < 9 owns variable j of the camera (p, xi, g, h, the start value, the last assigned value, the domain: one
    // register each); lane 0 also owns the state machine and the value cache.  Scalars travel by shuffles.
    const unsigned kFull = 0xffffffffu;
    const int rows = C * W;
    const bool vl = (lane < 9);
    CgdMachine mc;
    double pj = 0.0, xij = 0.0, gj = 0.0, hj = 0.0, xsj = 0.0, lastj = 0.0;
    double2 domj = make_double2(0.0, 0.0);
    if (vl) {
      pj = xsj = (B.x0 != nullptr) ? B.x0[P.var_off + lane] : Gv.xbd[v0 + lane].x;
      domj = __ldg(&Gv.dom[v0 + lane]);
    }
    double f_init = 0.0, c_sum = 0.0, f_at_p = 0.0;
    bool c_dirty = true, c_assigned = false;
    auto cache_eval = [&](double fresh) {  // Factor::eval over the list, src/Factor.cpp:110-119
      if (c_dirty) {
        c_sum = fresh;
        c_dirty = false;
      }
      return c_sum;
    };
    mc.start(maxiters, ftol, /*faithful=*/true);
    int flip = 0;
    uint32_t phase = 0;
#ifdef RDIS_CAM_PROFILE
    long long prof[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
    const long long tstart = tlast;
    int nev = 0;
#define SCPROF(i) do { const long long t_ = clock64(); prof[i] += t_ - tlast; tlast = t_; } while (0)
#else
#define SCPROF(i) do { } while (0)
#endif
    // Publishes the machine's request — the clamped point (lane-parallel), its rotation (lane 0), the direction — does
    // Variable::assign's bookkeeping, releases the workers.  Returns the request kind (uniform).
    auto publish = [&]() -> int {
      const int kind = __shfl_sync(kFull, mc.req, 0);
      const double alpha = __shfl_sync(kFull, mc.alpha, 0);
      if (lane == 0) sh.rq.kind = kind;
      if (kind != REQ_DONE) {
        const bool along = (kind == REQ_VALUE) || (kind == REQ_VALUE_SLOPE);
        double xj = 0.0;
        bool chg = false;
        if (vl) {
          xj = clamp_to_domain(along ? (pj + alpha * xij) : pj, domj);
          chg = !c_assigned || !(fabs(xj - lastj) < 1e-12);  // Variable::assign, src/Variable.cpp:66-88
          lastj = xj;
          sh.rq.x[lane] = xj;
          sh.rq.xi[lane] = xij;
        }
        if (__any_sync(kFull, chg)) c_dirty = true;
        c_assigned = true;
        const double r0 = __shfl_sync(kFull, xj, 0), r1 = __shfl_sync(kFull, xj, 1), r2 = __shfl_sync(kFull, xj, 2);
        if (lane == 0) {
          BaOps::Fwd m;
          BaOps::rotation(r0, r1, r2, m);
          sh.rq.a0 = m.a0; sh.rq.a1 = m.a1; sh.rq.a2 = m.a2; sh.rq.theta = m.theta; sh.rq.s = m.s; sh.rq.c = m.c;
        }
      }
      __syncwarp();
      named_bar_arrive(1, nthreads);
      return kind;
    };
    int kind = publish();
    SCPROF(0);
    while (kind != REQ_DONE) {
#ifdef RDIS_CAM_PROFILE
      ++nev;
#endif
      const bool along = (kind == REQ_VALUE) || (kind == REQ_VALUE_SLOPE);
      const int N = along ? 2 : kCamRedWidth;
      if (lane == 0) mbar_arrive_expect_tx(&sh.mbar[flip], (uint32_t)(rows * N * 8));
#ifdef RDIS_CAM_PLAIN_WAIT
      mbar_wait(&sh.mbar[flip], (phase >> flip) & 1u);
#else
      mbar_wait_cluster(&sh.mbar[flip], (phase >> flip) & 1u);
#endif
      SCPROF(1);  // workers' evaluation + flight
      phase ^= (1u << flip);
      // Fold of the cluster's warp partials in an order that is a function of the row index (= observation index / 32)
      // only: rows are dealt to lane groups by index modulo a constant, each group adds its rows in ascending order, the
      // groups are combined by a fixed tree.  Rows past C*W do not exist (they would hold exact zeros).
      double t0 = 0.0, t1 = 0.0, tot = 0.0;  // f, slope (line evaluations); column sums in lanes 0..9 (gradient evaluations)
      if (along) {
        const int col = lane & 1, grp = lane >> 1;  // 16 groups x 2 columns
        double sacc = 0.0;
        for (int r = grp; r < rows; r += 16) sacc += sh.red[flip][r][col];
#pragma unroll
        for (int o = 2; o < 32; o <<= 1) sacc += __shfl_xor_sync(kFull, sacc, o);
        t0 = __shfl_sync(kFull, sacc, 0);
        t1 = __shfl_sync(kFull, sacc, 1);
      } else {
        const int col = lane % kCamRedWidth, grp = lane / kCamRedWidth;  // 3 groups x 10 columns (lanes 30, 31 idle)
        double sacc = 0.0;
        if (grp < 3)
          for (int r = grp; r < rows; r += 3) sacc += sh.red[flip][r][col];
        const double s1 = __shfl_sync(kFull, sacc, (lane + 10) & 31), s2 = __shfl_sync(kFull, sacc, (lane + 20) & 31);
        tot = (sacc + s1) + s2;  // meaningful in lanes 0..9
        t0 = __shfl_sync(kFull, tot, 0);
      }
      const double gr_own = __shfl_sync(kFull, tot, (lane + 1) & 31);  // lane j < 9: d f / d x_j
      flip ^= 1;
      SCPROF(2);  // fold
      // ---- the step: lane 0 advances the machine; the 9-element vector updates it asks for run lane-parallel ----
      if (along) {
        if (lane == 0) mc.on_eval(cache_eval(t0), (kind == REQ_VALUE_SLOPE) ? t1 : 0.0);
        if (__shfl_sync(kFull, mc.req, 0) == REQ_MOVE) {  // minimize_nrc.h:508-511
          const double step = __shfl_sync(kFull, mc.alpha, 0);
          xij *= step;
          pj += xij;
          if (lane == 0) mc.on_moved();
        }
      } else if (kind == REQ_INIT_GRAD) {
        if (vl) {
          const double gneg = -gr_own;
          gj = gneg; hj = gneg; xij = gneg;
        }
        if (lane == 0) {
          f_at_p = t0;
          f_init = cache_eval(f_at_p);
          mc.on_init(f_init);
        }
      } else {  // REQ_GRADIENT: func.df(p, xi), then :656-685 (left-to-right sums over the variables, by lane 0)
        if (vl) xij = gr_own;
        double gg = 0.0, dgg = 0.0, tnum = 0.0;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const double pv = __shfl_sync(kFull, pj, j), gv = __shfl_sync(kFull, gj, j), grv = __shfl_sync(kFull, tot, 1 + j);
          const double pa = fabs(pv);
          const double tt = fabs(grv) * ((pa < 1.0) ? 1.0 : pa);
          tnum = (tt > tnum) ? tt : tnum;
          gg += gv * gv;
          dgg += (grv + gv) * grv;
        }
        if (lane == 0) {
          f_at_p = t0;
          mc.on_gradient(tnum, gg, dgg);
        }
        if (__shfl_sync(kFull, mc.req, 0) == REQ_DIRECTION) {
          const double gam = __shfl_sync(kFull, mc.gam, 0);
          if (vl) {
            const double gn = -xij;
            const double hn = gn + gam * hj;
            gj = gn; hj = hn; xij = hn;
          }
          if (lane == 0) mc.on_directed();
        }
      }
      // fa = func(ax = 0) at the top of a line search (minimize_nrc.h:88): clamp(p) is the point the gradient pass
      // assigned and f_at_p its objective — answered from the cache rules without another pass
      if (lane == 0 && mc.req == REQ_VALUE && mc.phase == CgdMachine::PH_BR_FA) mc.on_eval(cache_eval(f_at_p), 0.0);
      SCPROF(3);  // machine step
      kind = publish();
      SCPROF(0);  // point + rotation + publish
    }
#ifdef RDIS_CAM_PROFILE
    if (lane == 0 && cta == 0 && (blockIdx.x / C) < 4)
      printf("camprof prob %d nf %d evals %d cycles/eval %lld | publish %lld workers+flight %lld fold %lld machine %lld\n", pidx, nf, nev,
             (clock64() - tstart) / (nev > 0 ? nev : 1), prof[0] / nev, prof[1] / nev, prof[2] / nev, prof[3] / nev);
#endif
    // ---- commit (CGD.cpp:61-89): quickAssignVals(gdmin.p); if worse than the start the start point is re-assigned and
    //      re-evaluated through the cache.  The scalar warp of CTA 0 writes. ----
    {
      const double fret0 = __shfl_sync(kFull, mc.fret, 0);
      const double finit0 = __shfl_sync(kFull, f_init, 0);
      const bool restore = (fret0 > finit0);
      const double pf = clamp_to_domain(pj, domj), ps = clamp_to_domain(xsj, domj);
      const bool chgj = vl && (!(fabs(pf - lastj) < 1e-12) || !(fabs(ps - pf) < 1e-12));  // the two assigns: gdmin.p, then the start
      const bool chg = __any_sync(kFull, chgj);
      if (cta == 0) {
        if (vl) {
          const double val = restore ? ps : pf;
          Gv.xbd[v0 + lane] = make_double2(val, qnan_f64());
          Gv.xval[v0 + lane] = val;
          B.xout[P.var_off + lane] = val;
        }
        if (lane == 0) {
          double fret = mc.fret;
          if (restore) fret = (chg || c_dirty) ? f_init : c_sum;  // a recomputation at the start point reproduces f_init's bits
          ResultRec res;
          res.f_init = f_init;
          res.f_end = fret;
          res.iters = mc.iter;
          res.status = mc.status;
          res.n_value = mc.n_value;
          res.n_slope = mc.n_slope;
          B.res[pidx] = res;
        }
      }
    }
  }
  if (C > 1) cluster.sync();  // no CTA leaves while a peer could still address its shared memory
}

}  // namespace rdisgpu
