// sweep_kernels.cuh — the factor-graph sweeps outside a solve:
//   eval_sweep_kernel        OptimizableFunction::evalFactors        src/OptimizableFunction.cpp:95-135
//   factor_partials_kernel   Factor::computeGradient per factor      src/Factor.cpp:142-151 and overrides
//   gather_grad_kernel       computeGradientOfSum / productGradient  src/OptimizableFunction.cpp:248-262, src/State.h:157-194
// plus the Variable::assign / eval plumbing (scatter_x / gather_x).
// Sums are reproducible run to run: fixed grid, warp butterflies, block partials folded in index
// order by the last block to finish.
#pragma once
#include "factors.cuh"

namespace rdisgpu {

__global__ void scatter_x_kernel(GraphView G, int64_t n, const int32_t* vid, const double* x) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = vid ? vid[i] : i;
    G.xbd[v] = make_double2(x[i], qnan);
    G.xval[v] = x[i];
  }
}

__global__ void gather_x_kernel(GraphView G, int64_t n, const int32_t* vid, double* x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = vid ? vid[i] : i;
    x[i] = G.xbd[v].x;
  }
}

__global__ void set_fconst_kernel(uint8_t* on, double* val, int64_t n, const int32_t* fid, const double* v,
                                  const uint8_t* o) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    on[fid[i]] = o[i] ? 1 : 0;
    val[fid[i]] = v[i];
  }
}

__device__ __forceinline__ double warp_sum(double a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// Block total in thread 0 (fixed order), then the last block folds all block partials.
__device__ __forceinline__ void block_then_grid_sum(double v, double* partials, unsigned int* counter, double* out) {
  __shared__ double wsum[32];
  __shared__ bool is_last;
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) wsum[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = wsum[0];
    for (int w = 1; w < nw; ++w) t += wsum[w];
    partials[blockIdx.x] = t;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += ((volatile double*)partials)[i];
    t = warp_sum(t);
    if (lane == 0) wsum[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = wsum[0];
      for (int w = 1; w < nw; ++w) tt += wsum[w];
      *out = tt;
      *counter = 0u;  // ready for the next launch
    }
  }
}

template <class Ops>
__global__ void __launch_bounds__(256) eval_sweep_kernel(GraphView G, const int32_t* fids, int64_t nf, double* per_factor,
                                                         double* partials, unsigned int* counter, double* sum_out) {
  double acc = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nf; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fid = fids ? fids[k] : k;
    double sl;
    double fv = Ops::template value<false>(G, fid, 0.0, false, sl);
    if (G.fconst_on != nullptr && G.fconst_on[fid]) fv = G.fconst_val[fid];
    if (per_factor) per_factor[k] = fv;
    acc += fv;
  }
  block_then_grid_sum(acc, partials, counter, sum_out);
}

template <class Ops>
__global__ void __launch_bounds__(256) factor_partials_kernel(GraphView G, const int32_t* fids, int64_t nf, int32_t stamp) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nf; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fid = fids ? fids[k] : k;
    Ops::gradient(G, fid, G.gedge + Ops::edge_base(G, fid));
    if (stamp >= 0) G.fstamp[fid] = stamp;
  }
}

template <class Ops>
__global__ void __launch_bounds__(256) gather_grad_kernel(GraphView G, const int32_t* vids, int64_t nv, int32_t stamp,
                                                          bool filter, double* gout) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t vid = vids ? vids[i] : (int32_t)i;
    gout[i] = Ops::gather_var(G, vid, stamp, filter);
  }
}

__global__ void unstamp_kernel(GraphView G, const int32_t* fids, int64_t nf) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nf; k += (int64_t)gridDim.x * blockDim.x)
    G.fstamp[fids[k]] = -1;
}

template <class Ops>
__global__ void __launch_bounds__(128) factor_rows_kernel(GraphView G, const int32_t* fids, int64_t nf, int arity_max,
                                                          double* rows) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nf; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fid = fids ? fids[k] : k;
    const int ar = Ops::arity(G, fid);
    double* scratch = G.gedge + Ops::edge_base(G, fid);
    Ops::gradient(G, fid, scratch);
    for (int s = 0; s < ar && s < arity_max; ++s) rows[k * arity_max + s] = scratch[s];
  }
}

}  // namespace rdisgpu
