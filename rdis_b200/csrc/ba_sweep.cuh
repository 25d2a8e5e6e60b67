// ba_sweep.cuh — the all-factor residual sweep of a bundle-adjustment graph
//   OptimizableFunction::evalFactors                       src/OptimizableFunction.cpp:95-135
//   BundleAdjustmentFactor::evalFactor / evalPixelVals     src/bundleadjust/BundleAdjustmentFactor.cpp:55-64,160-185
// in one launch (two when the camera table does not fit shared memory):
//   camera table             one thread per CAMERA (inside every CTA, or ba_camera_table_kernel): the camera-only part of the forward model — |r|, the unit
//                            axis, sin / cos of the angle (BaOps::rotation), translation and intrinsics — into a
//                            96-byte table row.  Every observation of a camera re-uses it, so the square root,
//                            the three divisions and the sincos leave the per-factor path (ladybug: 650
//                            observations per camera on average).
//   ba_sweep_kernel          one thread per observation (kBaSweepUnroll in flight per thread): cam / pt / pixel streamed
//                            coalesced (24 B), the camera row from L1/L2, the point from the dense value mirror,
//                            BaOps::project, value out (8 B), fixed-order block / grid sum.
// Algorithmic bytes 32 B/factor + 8 B/variable.  Per-factor values are bit-identical to
// BaOps::value (same expressions in the same order; only WHERE the camera part is computed changes).
// Precondition as for the NLPF streaming sweep: no solve in flight, every variable frozen, xval mirrors xbd.x.
#pragma once
#include "factors.cuh"
#include "sweep_kernels.cuh"

namespace rdisgpu {

struct __align__(16) CameraRow {
  double a0, a1, a2, theta, s, c;  // BaOps::rotation
  double t0, t1, t2;               // translation
  double f, k1, k2;                // focal length, radial distortion
};

__global__ void ba_camera_table_kernel(GraphView G, CameraRow* __restrict__ table) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G.ncams) return;
  double x[9];
#pragma unroll
  for (int s = 0; s < 9; ++s) x[s] = G.xval[9 * c + s];
  BaOps::Fwd m;
  BaOps::rotation(x[0], x[1], x[2], m);
  CameraRow r;
  r.a0 = m.a0; r.a1 = m.a1; r.a2 = m.a2; r.theta = m.theta; r.s = m.s; r.c = m.c;
  r.t0 = x[3]; r.t1 = x[4]; r.t2 = x[5];
  r.f = x[6]; r.k1 = x[7]; r.k2 = x[8];
  table[c] = r;
}

// kSmemTable: every CTA first computes the whole camera table into shared memory itself (it fits whenever the graph has
// at most kBaSmemCams cameras — 49 on ladybug; 4 rows per thread at the limit): no separate launch, no table round trip
// through HBM, and lanes of a warp that look at up to 32 different cameras per observation batch read shared memory
// instead of issuing 6 x 32 sector requests through L1.  Larger graphs read the table ba_camera_table_kernel wrote.
constexpr int kBaSmemCams = 1024;
constexpr int kBaSmemRow = 13;  // doubles per shared-memory table row (12 + 1 pad)
constexpr int kBaSweepUnroll = 1;  // observations per thread and pipeline stage

#ifndef RDIS_BA_STAGES
#define RDIS_BA_STAGES 3
#endif
#ifndef RDIS_BA_SWEEP_CTAS
#define RDIS_BA_SWEEP_CTAS 3
#endif

// The sweep is balanced between HBM (32 B + a 24 B gather per observation) and the FP64 pipe (~92 instructions per
// observation without FMA contraction: 17 us of pipe time for 3.2 M observations against 18 us of HBM time), so the
// loop is SOFTWARE-PIPELINED three batches deep in registers: while batch i is being projected, the point gather of
// batch i+1 (whose indices landed during the previous iteration) and the index / pixel stream of batch i+2 are in
// flight.  A thread's chain is then one load round trip + compute per batch instead of two dependent round trips.
// kRows: also the 12 partial derivatives of every observation (the Levenberg-Marquardt path's Jacobian rows,
// LMSSOpt::evalJacf, src/optimizers/LMSubspaceOptimizer.cpp:207-278): 128 B/factor + 8 B/variable.
struct BaStage {
  int32_t c, p;        // camera, point of the observation
  double2 o;           // pixel
  double q0, q1, q2;   // the point (gathered one stage after the indices)
};

template <bool kSmemTable, bool kRows>
__global__ void __launch_bounds__(256, kRows ? 2 : RDIS_BA_SWEEP_CTAS) ba_sweep_kernel(GraphView G, const CameraRow* __restrict__ gtable, double* __restrict__ per_factor,
                                                       double* __restrict__ rows, double* partials, unsigned int* counter, double* sum_out) {
  extern __shared__ __align__(16) unsigned char ba_smem_raw[];
  // shared-memory rows are padded to 13 doubles: with 12 (48 words) the rows of a half-warp's 16 lanes fall into 4 bank
  // classes (8-byte reads, 16 lanes per wavefront), with 13 into 16 — lanes of a warp look at unrelated cameras
  double* const st = reinterpret_cast<double*>(ba_smem_raw);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double* qbase = G.xval + 9 * (int64_t)G.ncams;
  auto load_stream = [&](int64_t j, BaStage& S) {
    const bool in = j < G.F;
    S.c = in ? __ldg(&G.cam[j]) : 0;
    S.p = in ? __ldg(&G.pt[j]) : 0;
    S.o = in ? __ldg(&G.obs[j]) : make_double2(0.0, 0.0);
  };
  auto gather = [&](BaStage& S) {
    const double* qp = qbase + 3 * (int64_t)S.p;
    S.q0 = qp[0]; S.q1 = qp[1]; S.q2 = qp[2];
  };
  auto build_table = [&]() {
    if (kSmemTable) {
      for (int c = threadIdx.x; c < G.ncams; c += blockDim.x) {
        double x[9];
#pragma unroll
        for (int s = 0; s < 9; ++s) x[s] = G.xval[9 * c + s];
        BaOps::Fwd m;
        BaOps::rotation(x[0], x[1], x[2], m);
        double* r = st + kBaSmemRow * c;
        r[0] = m.a0; r[1] = m.a1; r[2] = m.a2; r[3] = m.theta; r[4] = m.s; r[5] = m.c;
#pragma unroll
        for (int s = 3; s < 9; ++s) r[3 + s] = x[s];
      }
      __syncthreads();
    }
  };
  double acc = 0.0;
  auto compute = [&](int64_t j, const BaStage& S) {
    CameraRow r;
    if (kSmemTable) {
      const double* t = st + kBaSmemRow * S.c;
      r.a0 = t[0]; r.a1 = t[1]; r.a2 = t[2]; r.theta = t[3]; r.s = t[4]; r.c = t[5];
      r.t0 = t[6]; r.t1 = t[7]; r.t2 = t[8]; r.f = t[9]; r.k1 = t[10]; r.k2 = t[11];
    } else {
      r = gtable[S.c];
    }
    double x[12];
    x[0] = 0.0; x[1] = 0.0; x[2] = 0.0;  // project() / partials() read the rotation from m, not from x[0..2]
    x[3] = r.t0; x[4] = r.t1; x[5] = r.t2; x[6] = r.f; x[7] = r.k1; x[8] = r.k2;
    x[9] = S.q0; x[10] = S.q1; x[11] = S.q2;
    BaOps::Fwd m;
    m.a0 = r.a0; m.a1 = r.a1; m.a2 = r.a2; m.theta = r.theta; m.s = r.s; m.c = r.c;
    double fv = BaOps::project(x, S.o, m);
    if (kRows) {
      double g[12];
      BaOps::partials(x, m, g);
      double2* dst = reinterpret_cast<double2*>(rows + 12 * j);  // 96 B per observation, 16-byte aligned
#pragma unroll
      for (int s = 0; s < 12; s += 2) __stcs(dst + (s >> 1), make_double2(g[s], g[s + 1]));
    }
    if (G.fconst_on != nullptr && G.fconst_on[j]) fv = G.fconst_val[j];  // Factor::eval, src/Factor.cpp:110-119
    if (per_factor) __stcs(&per_factor[j], fv);
    acc += fv;
  };
  // three stages in NAMED registers, the loop unrolled by three so that no stage is ever copied (a rotating array of
  // stages is spilled to local memory by the compiler, and the spill store waits for the very load it should overlap)
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
#if RDIS_BA_STAGES == 4
  // four stages: indices / pixel three observations ahead, the point gather two ahead
  BaStage A, Bs, Cs, Ds;
  load_stream(j, A);
  load_stream(j + stride, Bs);
  load_stream(j + 2 * stride, Cs);
  build_table();
  gather(A);
  gather(Bs);
#define RDIS_BA_STEP(CUR, NEXT, NEXT2, AFTER)  \
  if (j >= G.F) break;                         \
  gather(NEXT2);                               \
  load_stream(j + 3 * stride, AFTER);          \
  compute(j, CUR);                             \
  j += stride;
  for (;;) {
    RDIS_BA_STEP(A, Bs, Cs, Ds)
    RDIS_BA_STEP(Bs, Cs, Ds, A)
    RDIS_BA_STEP(Cs, Ds, A, Bs)
    RDIS_BA_STEP(Ds, A, Bs, Cs)
  }
#undef RDIS_BA_STEP
#else
  BaStage A, Bs, Cs;
  load_stream(j, A);
  load_stream(j + stride, Bs);
  build_table();  // the first index / pixel loads are in flight while the CTA computes its camera table
  gather(A);
#define RDIS_BA_STEP(CUR, NEXT, AFTER)  \
  if (j >= G.F) break;                  \
  gather(NEXT);                         \
  load_stream(j + 2 * stride, AFTER);   \
  compute(j, CUR);                      \
  j += stride;
  for (;;) {
    RDIS_BA_STEP(A, Bs, Cs)
    RDIS_BA_STEP(Bs, Cs, A)
    RDIS_BA_STEP(Cs, A, Bs)
  }
#undef RDIS_BA_STEP
#endif
  block_then_grid_sum(acc, partials, counter, sum_out);
}

}  // namespace rdisgpu
