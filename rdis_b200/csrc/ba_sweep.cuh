// ba_sweep.cuh — the all-factor residual sweep of a bundle-adjustment graph
//   OptimizableFunction::evalFactors                       src/OptimizableFunction.cpp:95-135
//   BundleAdjustmentFactor::evalFactor / evalPixelVals     src/bundleadjust/BundleAdjustmentFactor.cpp:55-64,160-185
// in two launches:
//   ba_camera_table_kernel   one thread per CAMERA: the camera-only part of the forward model — |r|, the unit
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

// kSmemTable: the whole camera table is first copied into shared memory (it fits whenever the graph has at most
// kBaSmemCams cameras — 49 on ladybug): lanes of a warp look at up to 32 different cameras per observation batch,
// which as global loads is 6 x 32 sector requests through L1 per warp, and as shared-memory reads a few wavefronts.
constexpr int kBaSmemCams = 1024;
#ifndef RDIS_BA_SWEEP_UNROLL
#define RDIS_BA_SWEEP_UNROLL 4
#endif
constexpr int kBaSweepUnroll = RDIS_BA_SWEEP_UNROLL;

#ifndef RDIS_BA_SWEEP_CTAS
#define RDIS_BA_SWEEP_CTAS 3
#endif
template <bool kSmemTable>
__global__ void __launch_bounds__(256, RDIS_BA_SWEEP_CTAS) ba_sweep_kernel(GraphView G, const CameraRow* __restrict__ gtable, double* __restrict__ per_factor,
                                                       double* partials, unsigned int* counter, double* sum_out) {
  extern __shared__ __align__(16) unsigned char ba_smem_raw[];
  const CameraRow* table = gtable;
  if (kSmemTable) {
    CameraRow* st = reinterpret_cast<CameraRow*>(ba_smem_raw);
    const double2* src = reinterpret_cast<const double2*>(gtable);
    double2* dst = reinterpret_cast<double2*>(st);
    for (int i = threadIdx.x; i < G.ncams * 6; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    table = st;
  }
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = kBaSweepUnroll;  // observations in flight per thread: all their loads are issued before any is used
  for (int64_t j0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j0 < G.F; j0 += U * stride) {
    int32_t c[U], p[U];
    double2 o[U];
    double q[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = j0 + u * stride;
      const bool in = j < G.F;
      c[u] = in ? __ldg(&G.cam[j]) : 0;
      p[u] = in ? __ldg(&G.pt[j]) : 0;
      o[u] = in ? __ldg(&G.obs[j]) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* qp = G.xval + 9 * (int64_t)G.ncams + 3 * (int64_t)p[u];
      q[u][0] = qp[0]; q[u][1] = qp[1]; q[u][2] = qp[2];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = j0 + u * stride;
      if (j < G.F) {
        const CameraRow r = table[c[u]];
        double x[12];
        x[3] = r.t0; x[4] = r.t1; x[5] = r.t2; x[6] = r.f; x[7] = r.k1; x[8] = r.k2;
        x[9] = q[u][0]; x[10] = q[u][1]; x[11] = q[u][2];
        BaOps::Fwd m;
        m.a0 = r.a0; m.a1 = r.a1; m.a2 = r.a2; m.theta = r.theta; m.s = r.s; m.c = r.c;
        double fv = BaOps::project(x, o[u], m);
        if (G.fconst_on != nullptr && G.fconst_on[j]) fv = G.fconst_val[j];  // Factor::eval, src/Factor.cpp:110-119
        if (per_factor) __stcs(&per_factor[j], fv);
        acc += fv;
      }
    }
  }
  block_then_grid_sum(acc, partials, counter, sum_out);
}

}  // namespace rdisgpu
