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

__device__ __forceinline__ void solve_ba_point_tile(const GraphView& Gv, const BatchView& B, int lg, int pidx, int maxiters,
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
  while (true) {
    const bool fin = mc.done();
    if (__all_sync(kFull, fin)) break;
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

    // left-to-right folds over the tile's lanes (= the problem's factor list), replicated in every lane
    double fs = 0.0;
    double gr[3] = {0.0, 0.0, 0.0};
    const bool any_g = __any_sync(kFull, want_g);
    for (int k = 0; k < G; ++k) {
      const double bf = __shfl_sync(kFull, fv, k, G);
      if (k < nf) fs = fs + bf;
      if (any_g) {
        const double b0 = __shfl_sync(kFull, g9, k, G), b1 = __shfl_sync(kFull, g10, k, G), b2 = __shfl_sync(kFull, g11, k, G);
        if (k == 0) {
          gr[0] = b0; gr[1] = b1; gr[2] = b2;
        } else if (k < nf) {
          gr[0] = gr[0] + b0; gr[1] = gr[1] + b1; gr[2] = gr[2] + b2;
        }
      }
    }

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
  }

  if (!run) return;
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
}

// One warp per CTA.  warp_task[w] = {log2 G, first slot in `order`, number of problems of that
// class still to hand out from that slot on}: every size class runs in the SAME launch, so the
// launch lasts as long as the slowest single problem, not the sum of the per-class tails.
struct PointWarpTask {
  int32_t lg;
  int32_t first;
  int32_t count;
};

__global__ void __launch_bounds__(32) solve_ba_points_kernel(GraphView Gv, BatchView B, const int32_t* order,
                                                             const PointWarpTask* tasks, int maxiters, double ftol) {
  const PointWarpTask t = tasks[blockIdx.x];
  const int slot = (int)threadIdx.x >> t.lg;
  solve_ba_point_tile(Gv, B, t.lg, (slot < t.count) ? order[t.first + slot] : -1, maxiters, ftol);
}

// ------------------------------------------------------------------------------------------
// camera blocks
// ------------------------------------------------------------------------------------------
#ifndef RDIS_CAM_THREADS
#define RDIS_CAM_THREADS 192
#endif
#ifndef RDIS_CAM_MIN_CTAS
#define RDIS_CAM_MIN_CTAS 2
#endif
// 2 CTAs per SM at <= 168 registers: clusters can be wide enough for one observation per thread
constexpr int kCamMaxThreads = RDIS_CAM_THREADS;
constexpr int kCamMaxCluster = 8;
constexpr int kCamMaxWarps = kCamMaxThreads / 32;
constexpr int kCamRedWidth = 10;  // f + 9 partials (gradient mode); f + slope use the first two

struct CamShared {
  double p[9], xi[9], g[9], h[9], xs[9];
  double last[9];  // Variable::eval() of the camera's variables: the point of the previous evaluation (thread 0 only)
  double2 dom[9];
  double wred[kCamMaxWarps][kCamRedWidth];          // warp partials of this CTA (first reduction level)
  double red[2][kCamMaxCluster][kCamRedWidth];      // CTA partials of the whole cluster (second level, double-buffered)
  unsigned long long mbar[2];  // one transaction barrier per reduction buffer (remote st.async completes on it)
  CgdMachine m;                // advanced by thread 0 of every CTA of the cluster on identical inputs
  double tot[kCamRedWidth];    // the all-reduced sums of the current evaluation (thread 0 -> nobody else needs them)
};

// Sum of N doubles per thread over the whole cluster into sh.tot of EVERY CTA (read by thread 0 only), fixed order,
// two levels:
//   (1) warp butterfly, warp partials folded per CTA through shared memory (one __syncthreads);
//   (2) lane d of warp 0 pushes the CTA's partial into slot `cta` of CTA d's buffer with an ASYNCHRONOUS
//       store over distributed shared memory that completes a transaction barrier in the receiving CTA
//       (st.async ... mbarrier::complete_tx); thread 0 waits for its own barrier to have received C*N*8
//       bytes and folds the C partials in CTA order: one one-way trip per evaluation, no cluster barrier.
// The cluster buffers (and their barriers) alternate; re-use two rounds later is safe because a CTA can only
// send round r+1 after its thread 0 folded round r, and nobody finishes round r+1 before every CTA of the cluster has
// sent it.  `phase` holds the two barriers' parities.  The caller's __syncthreads after thread 0's scalar step
// orders the next evaluation's wred writes after this fold.
#ifdef RDIS_CAM_PROFILE
#define CAMPROF(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); prof[i] += t_ - tlast; tlast = t_; } } while (0)
#define CAMPROF_ARGS , long long* prof, long long& tlast
#define CAMPROF_PASS , prof, tlast
#else
#define CAMPROF(i) do { } while (0)
#define CAMPROF_ARGS
#define CAMPROF_PASS
#endif

template <int N>
__device__ __forceinline__ void cluster_reduce_to_thread0(CamShared& sh, int& flip, uint32_t& phase, double (&v)[N], int C, int cta CAMPROF_ARGS) {
  static_assert(N % 2 == 0, "partials travel as 16-byte pairs");
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) sh.wred[warp][i] = v[i];
  }
  CAMPROF(1);  // evaluation + warp butterfly (thread 0's view)
  __syncthreads();
  CAMPROF(2);  // waiting for the CTA's other warps
  if (warp != 0) return;
  if (C == 1) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = sh.wred[0][i];
      for (int w = 1; w < nw; ++w) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += sh.wred[w][i];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) sh.tot[i] = v[i];
    }
    return;
  }
  if (lane == 0) mbar_arrive_expect_tx(&sh.mbar[flip], (uint32_t)(C * N * 8));
  if (lane < C) {  // lane d delivers this CTA's partial to CTA d
    double c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) c[i] = sh.wred[0][i];
    for (int w = 1; w < nw; ++w) {
#pragma unroll
      for (int i = 0; i < N; ++i) c[i] += sh.wred[w][i];
    }
    const uint32_t dst = map_to_cta(smem_addr_u32(&sh.red[flip][cta][0]), (uint32_t)lane);
    const uint32_t bar = map_to_cta(smem_addr_u32(&sh.mbar[flip]), (uint32_t)lane);
#pragma unroll
    for (int i = 0; i < N; i += 2) st_async_v2(dst + 8u * i, c[i], c[i + 1], bar);
  }
  CAMPROF(3);  // fold of the warp partials + st.async issue
  if (lane == 0) {
    mbar_wait_cluster(&sh.mbar[flip], (phase >> flip) & 1u);
    CAMPROF(4);  // waiting for the peers' partials
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = sh.red[flip][0][i];
    for (int s = 1; s < C; ++s) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] += sh.red[flip][s][i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) sh.tot[i] = v[i];
  }
  phase ^= (1u << flip);
  flip ^= 1;
}

// grid = nprobs * C CTAs, cluster = C CTAs (set at launch), `order` lists the camera-class problems.
//
// One evaluation = [all threads] read the request (kind, alpha) and the camera state from shared memory, form the
// point, the rotation, their observation's value (+ partials), reduce over the cluster; [thread 0 of every CTA] the
// scalar step: the value cache (BlockCache semantics), the CgdMachine, and the 9-element vector updates it asks for
// (move, gradient bookkeeping, new direction); one CTA barrier.  Keeping the machine out of the other threads'
// registers is what lets the observation arithmetic run without spills at 2-3 CTAs per SM.
__global__ void __launch_bounds__(kCamMaxThreads, RDIS_CAM_MIN_CTAS) solve_ba_cameras_kernel(GraphView Gv, BatchView B, const int32_t* order,
                                                                          int C, int maxiters, double ftol) {
  namespace cgn = cooperative_groups;
  cgn::cluster_group cluster = cgn::this_cluster();
  __shared__ CamShared sh;
  const int cta = (C > 1) ? (int)cluster.block_rank() : 0;
  const int pidx = order[blockIdx.x / C];
  const ProblemDesc P = B.probs[pidx];
  const int nf = P.nf;
  const int32_t v0 = B.vids[P.var_off];
  const int T = blockDim.x;
  const int rank = cta * T + threadIdx.x;
  const int size = C * T;
  int flip = 0;
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    mbar_init(&sh.mbar[0], 1);
    mbar_init(&sh.mbar[1], 1);
    mbar_fence_init();
    sh.m.start(maxiters, ftol, /*faithful=*/true);
  }
  if (threadIdx.x < 9) {
    const int j = threadIdx.x;
    const double xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : Gv.xbd[v0 + j].x;
    sh.p[j] = xv; sh.xs[j] = xv; sh.xi[j] = 0.0; sh.g[j] = 0.0; sh.h[j] = 0.0;
    sh.last[j] = 0.0;
    sh.dom[j] = __ldg(&Gv.dom[v0 + j]);
  }
  // this thread's first observation: frozen point block + pixel staged in registers
  const bool have = (rank < nf);
  double q0 = 0.0, q1 = 0.0, q2 = 0.0;
  double2 ob0 = make_double2(0.0, 0.0);
  bool fc_on0 = false;
  double fc_val0 = 0.0;
  const int32_t pbase = 9 * Gv.ncams;
  if (have) {
    const int32_t fid = B.fids[P.fac_off + rank];
    const int32_t pt = __ldg(&Gv.pt[fid]);
    ob0 = __ldg(&Gv.obs[fid]);
    q0 = Gv.xbd[pbase + 3 * pt].x; q1 = Gv.xbd[pbase + 3 * pt + 1].x; q2 = Gv.xbd[pbase + 3 * pt + 2].x;
    if (Gv.fconst_on != nullptr && Gv.fconst_on[fid]) {
      fc_on0 = true;
      fc_val0 = Gv.fconst_val[fid];
    }
  }
  __syncthreads();
  if (C > 1) cluster.sync();  // every CTA's barriers are initialised before a peer's st.async can reach them

  // thread 0's private scalars: the value cache of the block (BlockCache semantics) and the solve's bookkeeping
  double f_init = 0.0, c_sum = 0.0, f_at_p = 0.0;
  bool c_dirty = true, c_assigned = false;

#ifdef RDIS_CAM_PROFILE
  long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long profx[3] = {0, 0, 0};
  int nslope = 0;
  long long tlast = clock64();
  const long long tstart = tlast;
  int nev = 0;
#endif
  while (true) {
    const int kind = sh.m.req;
    if (kind == REQ_DONE) break;
    CAMPROF(0);  // barrier release -> request read
#ifdef RDIS_CAM_PROFILE
    ++nev;
#endif
    const bool along = (kind == REQ_VALUE) || (kind == REQ_VALUE_SLOPE);
    const bool want_g = (kind != REQ_VALUE);
    const double alpha = sh.m.alpha;
    double x[12];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const double pj = sh.p[j];
      x[j] = clamp_to_domain(along ? (pj + alpha * sh.xi[j]) : pj, sh.dom[j]);
    }
    BaOps::Fwd m;
    BaOps::rotation(x[0], x[1], x[2], m);
#ifdef RDIS_CAM_PROFILE
    if (threadIdx.x == 0 && m.s != 2.0) { const long long t_ = clock64(); profx[0] += t_ - tlast; tlast = t_; }
#endif

    // this thread's observations: round 0 from registers, later rounds (more observations than
    // threads in the cluster) re-read their frozen block through L1/L2
    auto stage = [&](int k, double2& ob, bool& fc_on, double& fc_val) {
      if (k == rank) {
        x[9] = q0; x[10] = q1; x[11] = q2; ob = ob0; fc_on = fc_on0; fc_val = fc_val0;
      } else {
        const int32_t fid = B.fids[P.fac_off + k];
        const int32_t pt = __ldg(&Gv.pt[fid]);
        ob = __ldg(&Gv.obs[fid]);
        x[9] = Gv.xbd[pbase + 3 * pt].x; x[10] = Gv.xbd[pbase + 3 * pt + 1].x; x[11] = Gv.xbd[pbase + 3 * pt + 2].x;
        fc_on = (Gv.fconst_on != nullptr) && Gv.fconst_on[fid];
        fc_val = fc_on ? Gv.fconst_val[fid] : 0.0;
      }
    };

    if (along) {
      double v2[2] = {0.0, 0.0};
      for (int k = rank; k < nf; k += size) {
        double2 ob;
        bool fc_on;
        double fc_val;
        stage(k, ob, fc_on, fc_val);
        double fv = BaOps::project(x, ob, m);
        if (want_g) {
          double gq[12];
          BaOps::partials(x, m, gq);
          double sl = 0.0;
#pragma unroll
          for (int s = 0; s < 9; ++s) sl += gq[s] * sh.xi[s];
          v2[1] += sl;
        }
        if (fc_on) fv = fc_val;
        v2[0] += fv;
      }
#ifdef RDIS_CAM_PROFILE
      if (threadIdx.x == 0 && v2[0] != -1.0) { const long long t_ = clock64(); profx[want_g ? 2 : 1] += t_ - tlast; tlast = t_; if (want_g) ++nslope; }
#endif
      cluster_reduce_to_thread0<2>(sh, flip, phase, v2, C, cta CAMPROF_PASS);
    } else {
      double acc[kCamRedWidth];
#pragma unroll
      for (int i = 0; i < kCamRedWidth; ++i) acc[i] = 0.0;
      for (int k = rank; k < nf; k += size) {
        double2 ob;
        bool fc_on;
        double fc_val;
        stage(k, ob, fc_on, fc_val);
        double fv = BaOps::project(x, ob, m);
        double gq[12];
        BaOps::partials(x, m, gq);
#pragma unroll
        for (int s = 0; s < 9; ++s) acc[1 + s] += gq[s];
        if (fc_on) fv = fc_val;
        acc[0] += fv;
      }
      cluster_reduce_to_thread0<kCamRedWidth>(sh, flip, phase, acc, C, cta CAMPROF_PASS);
    }

    // ---- the scalar step (thread 0 of every CTA, identical inputs in every CTA of the cluster) ----
    CAMPROF(5);  // fold of the cluster's partials
    if (threadIdx.x == 0) {
      CgdMachine mc = sh.m;  // registers for the duration of the step only (one burst of loads instead of dependent ones)
      // Variable::assign of this evaluation's point (src/Variable.cpp:66-88): a move below 1e-12 notifies nobody
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        if (!c_assigned || !(fabs(x[j] - sh.last[j]) < 1e-12)) c_dirty = true;
        sh.last[j] = x[j];
      }
      c_assigned = true;
      CAMPROF(6);  // machine copy-in + Variable::assign bookkeeping
      auto cache_eval = [&](double fresh) {  // Factor::eval over the list, src/Factor.cpp:110-119
        if (c_dirty) {
          c_sum = fresh;
          c_dirty = false;
        }
        return c_sum;
      };
      if (along) {
        mc.on_eval(cache_eval(sh.tot[0]), (kind == REQ_VALUE_SLOPE) ? sh.tot[1] : 0.0);
        if (mc.req == REQ_MOVE) {  // minimize_nrc.h:508-511
          const double step = mc.alpha;
          for (int j = 0; j < 9; ++j) {
            const double d = sh.xi[j] * step;
            sh.xi[j] = d;
            sh.p[j] += d;
          }
          mc.on_moved();
        }
      } else if (kind == REQ_INIT_GRAD) {
        for (int j = 0; j < 9; ++j) {
          const double gneg = -sh.tot[1 + j];
          sh.g[j] = gneg; sh.h[j] = gneg; sh.xi[j] = gneg;
        }
        f_at_p = sh.tot[0];
        f_init = cache_eval(f_at_p);
        mc.on_init(f_init);
      } else {  // REQ_GRADIENT: func.df(p, xi), then :656-685
        f_at_p = sh.tot[0];
        double gg = 0.0, dgg = 0.0, tnum = 0.0;
        for (int j = 0; j < 9; ++j) {
          const double gr = sh.tot[1 + j];
          const double pj = fabs(sh.p[j]);
          const double t = fabs(gr) * ((pj < 1.0) ? 1.0 : pj);
          tnum = (t > tnum) ? t : tnum;
          const double gj = sh.g[j];
          gg += gj * gj;
          dgg += (gr + gj) * gr;
        }
        mc.on_gradient(tnum, gg, dgg);
        if (mc.req == REQ_DIRECTION) {
          const double gam = mc.gam;
          for (int j = 0; j < 9; ++j) {
            const double gj = -sh.tot[1 + j];
            const double hj = gj + gam * sh.h[j];
            sh.g[j] = gj; sh.h[j] = hj; sh.xi[j] = hj;
          }
          mc.on_directed();
        } else {
          for (int j = 0; j < 9; ++j) sh.xi[j] = sh.tot[1 + j];
        }
      }
      // fa = func(ax = 0) at the top of a line search (minimize_nrc.h:88): clamp(p) is the point the gradient pass
      // assigned and f_at_p its objective — answered from the cache rules without another pass
      if (mc.req == REQ_VALUE && mc.phase == CgdMachine::PH_BR_FA) mc.on_eval(cache_eval(f_at_p), 0.0);
      CAMPROF(7);  // machine step + vector updates
      sh.m = mc;
      CAMPROF(8);  // machine copy-out
    }
    __syncthreads();
    CAMPROF(9);  // closing barrier
  }
#ifdef RDIS_CAM_PROFILE
  if (threadIdx.x == 0 && cta == 0 && (blockIdx.x / C) < 4)
    printf("camprof prob %d: x+rotation %lld per eval; value-only obs pass %lld per value eval (%d); value+slope obs pass %lld per slope eval (%d); eval includes butterfly\n",
           pidx, profx[0] / nev, profx[1] / (nev - nslope > 0 ? nev - nslope : 1), nev - nslope, profx[2] / (nslope > 0 ? nslope : 1), nslope);
  if (threadIdx.x == 0 && cta == 0 && (blockIdx.x / C) < 4)
    printf("camprof prob %d nf %d evals %d cycles/eval %lld | req %lld eval %lld bar1 %lld wfold+send %lld peerwait %lld cfold %lld copyin+assign %lld machine %lld copyout %lld bar2 %lld\n",
           pidx, nf, nev, (clock64() - tstart) / (nev > 0 ? nev : 1), prof[0] / nev, prof[1] / nev, prof[2] / nev, prof[3] / nev, prof[4] / nev,
           prof[5] / nev, prof[6] / nev, prof[7] / nev, prof[8] / nev, prof[9] / nev);
#endif

  // ---- commit (CGD.cpp:61-89): quickAssignVals(gdmin.p); if worse than the start the start point is re-assigned and
  //      re-evaluated through the cache.  CTA 0 of the cluster writes. ----
  if (cta == 0 && threadIdx.x == 0) {
    const CgdMachine& mc = sh.m;
    double fret = mc.fret;
    const bool restore = (fret > f_init);
    if (restore) {
      bool chg = c_dirty;
      for (int j = 0; j < 9; ++j) {  // the two assigns: gdmin.p, then the start point
        const double pf = clamp_to_domain(sh.p[j], sh.dom[j]), ps = clamp_to_domain(sh.xs[j], sh.dom[j]);
        if (!(fabs(pf - sh.last[j]) < 1e-12) || !(fabs(ps - pf) < 1e-12)) chg = true;
      }
      fret = chg ? f_init : c_sum;  // a recomputation at the start point reproduces f_init's bits
    }
    for (int j = 0; j < 9; ++j) {
      const double val = clamp_to_domain(restore ? sh.xs[j] : sh.p[j], sh.dom[j]);
      Gv.xbd[v0 + j] = make_double2(val, qnan_f64());
      Gv.xval[v0 + j] = val;
      B.xout[P.var_off + j] = val;
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
  if (C > 1) cluster.sync();  // no CTA leaves while a peer could still address its shared memory
}

}  // namespace rdisgpu
