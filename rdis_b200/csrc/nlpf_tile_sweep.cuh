// nlpf_tile_sweep.cuh — the HBM-streaming form of the NonlinearProductFactor sweeps over ALL factors:
//   nlpf_tile_sweep_kernel<false>   OptimizableFunction::evalFactors          src/OptimizableFunction.cpp:95-135
//   nlpf_tile_sweep_kernel<true>    + Factor::computeGradient of every factor  src/NonlinearProductFactor.cpp:57-117
//
// The factor CSR is cut (on the host, at finalize) into TILES of consecutive factors whose edge rows
// are one contiguous slice of the edge arrays: at most kTileFactors factors and kTileEdges edges.
// The kernel is PERSISTENT (a few CTAs per SM, tiles dealt round-robin) and software-pipelined:
//
//   producer   a dedicated warp (one elected lane) streams the tile's six array slices — evid i32, expo f64, konst f64,
//              sine u8 (21 B/edge), rowptr i32, coeff f64 (12 B/factor) — from HBM into a shared-memory
//              stage with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), kTileStages tiles
//              ahead of the math (full / empty mbarrier ring).  No registers are spent on data in flight, so the bytes in flight per
//              SM are set by the stage size (~28 KB x stages x CTAs/SM), not by occupancy.
//   phase 1    (edge-parallel) thread t owns edges t, t+T, ... of the stage: the variable gather
//              (L2-resident, 16 B/var) was put in flight one tile earlier; the thread parks the value term t_e = [sin]((x-k)^e) — with kGrad also
//              its own-slot derivative — in shared memory.  The transcendental work is spread evenly
//              over the threads whatever the arity mix of the tile.
//   phase 2    (factor-parallel) thread t owns factors t, t+T, ...: folds the product from shared
//              memory IN SLOT ORDER (the reference's loop order, NonlinearProductFactor.cpp:186-209),
//              writes the value (coalesced) and accumulates it.
// HBM sees each algorithmic byte once, as full-line bulk traffic.
// Per-factor values are bit-identical to NlpfOps::value / NlpfOps::gradient (same expressions);
// the grand total is folded in a fixed order (reproducible run to run on one device).
#pragma once
#include "factors.cuh"
#include "ptx_async.cuh"
#include "sweep_kernels.cuh"

namespace rdisgpu {

// Tuning constants (the -D overrides exist for the sweeps recorded in profiles/; the defaults ship).
#ifndef RDIS_TILE_THREADS
#define RDIS_TILE_THREADS 256
#endif
#ifndef RDIS_TILE_EDGES
#define RDIS_TILE_EDGES 512
#endif
#ifndef RDIS_TILE_FACTORS
#define RDIS_TILE_FACTORS 512
#endif
#ifndef RDIS_TILE_STAGES
#define RDIS_TILE_STAGES 3
#endif
#ifndef RDIS_TILE_CTAS
#define RDIS_TILE_CTAS 3
#endif
constexpr int kTileThreads = RDIS_TILE_THREADS;   // consumer threads
constexpr int kTileBlock = kTileThreads + 64;     // + the bulk-copy warp + the gather warp
constexpr int kTileEdges = RDIS_TILE_EDGES;
constexpr int kTileFactors = RDIS_TILE_FACTORS;
constexpr int kTileStages = RDIS_TILE_STAGES;
#ifndef RDIS_TILE_CTAS_GRAD
#define RDIS_TILE_CTAS_GRAD 2
#endif
constexpr int kTileCtasPerSm = RDIS_TILE_CTAS;
constexpr int kEdgeBatch = kTileEdges / kTileThreads;  // gathers per thread per tile, issued one tile ahead

// One tile: factors [f0, f1), edges [e0, e1).  e1 - e0 > kTileEdges marks a single over-wide factor
// (folded serially from global memory; never produced by the reference's generators).
struct TileDesc {
  int32_t f0, f1, e0, e1;
};

// Shared-memory stage.  Bulk copies need 16-byte aligned source, destination and size, so a slice is
// fetched from its start rounded down to 16 elements (edges) / 4 elements (factors) — the slack is
// the +32 / +8 below — and the device arrays carry the same padding at their ends.
struct TileStage {
  double expo[kTileEdges + 32];
  double konst[kTileEdges + 32];
  double coeff[kTileFactors + 8];
  int32_t evid[kTileEdges + 32];
  int32_t rowptr[kTileFactors + 8];
  double xs[kTileEdges + 32];   // gathered variable values (cp.async by the producer warp)
  uint8_t sine[kTileEdges + 32];
  TileDesc desc;  // written by the producer before it arms the barrier: consumers never touch the descriptor array
};
static_assert(sizeof(TileStage) % 16 == 0, "stages must keep 16-byte alignment");

template <bool kGrad>
struct TileSmem {
  TileStage stage[kTileStages];
  // kGrad: the tile's per-edge partials, parked here by phase 2 and written to gedge by ONE bulk store per tile (two
  // buffers: the store of tile i reads its buffer while phase 2 of tile i+1 fills the other).  Slot parity = parity of
  // the global edge index, so that the even-aligned body of the slice is 16-byte aligned on both sides.
  double gout[kGrad ? 2 : 1][kGrad ? kTileEdges + 2 : 2];
  unsigned long long full[kTileStages];   // producer -> consumers: the stage's bytes have landed
  unsigned long long xready[kTileStages]; // producer -> consumers: ... and the gathered variable values too
  unsigned long long empty[kTileStages];  // consumers -> producer: every consumer warp is done with the stage
};

// gedge[e0 .. e0+ne) <- buf[(e0 & 1) ...]: the 16-byte aligned body as one bulk store, an odd first / last edge directly.
// Called by one consumer thread after a consumer barrier that follows the writes to buf.
__device__ __forceinline__ void tile_store_partials(double* __restrict__ gedge, const double* buf, int e0, int ne) {
  const int ao = e0 & 1;
  const int body = (ne - ao) & ~1;
  if (body > 0) tma_bulk_s2g(gedge + e0 + ao, buf + 2 * ao, (uint32_t)body * 8u);
  if (ao) gedge[e0] = buf[1];
  if (ao + body < ne) gedge[e0 + ne - 1] = buf[ao + ne - 1];
  bulk_commit();
}

// value term exactly as NlpfOps::value (no-slope path) computes it
__device__ __forceinline__ double nlpf_term_value(double xv, double k, double ex, bool sn) {
  double val = xv;
  if (k != 0) val -= k;
  if (ex != 1) val = rdis_power(val, ex);
#ifndef RDIS_EXP_NOSIN
  if (sn) val = rdis_sin(val);
#endif
  return val;
}

// value term + own-slot derivative exactly as NlpfOps::term computes them (plain slot: derivative 1,
// and x * 1.0 == x to the bit, so the reference's "skip the multiplication" needs no flag)
__device__ __forceinline__ void nlpf_term_grad(double xv, double k, double ex, bool sn, double& t, double& dt) {
  double val = xv;
  if (k != 0) val -= k;
  if (ex == 1) {
    // Exponent 1 (every term of the sinusoid family): the general path below reduces, bit for bit, to
    //   plain:  t = x - k, dt = 1;   sine:  t = sin(x - k), dt = 1.0 * 1.0 * cos(x - k) = cos(x - k)
    // (power(inner, 1) = inner, power(inner, 0) = 1, and inner == val because x - 0 == x), so the three power() calls
    // and their branches are skipped — a third of the instructions of a sine term.
    if (!sn) {
      t = val;
      dt = 1.0;
    } else {
      double sv, cv;
      rdis_sincos(val, sv, cv);
      t = sv;
      dt = cv;
    }
    return;
  }
  val = rdis_power(val, ex);
  const double inner = xv - k;
  const double innerexp = rdis_power(inner, ex);
  double dv = rdis_power(inner, ex - 1.0);
  dv *= ex;
  if (sn) {
    double sv, cv;
    rdis_sincos(val, sv, cv);
    dv *= (innerexp == val) ? cv : rdis_cos(innerexp);
    t = sv;
  } else {
    t = val;
  }
  dt = dv;
}

// Producer: arm the stage's barrier with the byte count and launch the six bulk copies of tile `d`.
__device__ __forceinline__ void tile_issue(const GraphView& G, const TileDesc d, TileStage& st, unsigned long long* bar,
                                           uint64_t pol) {
  const int ne = d.e1 - d.e0;
  if (ne > kTileEdges) {  // over-wide factor: nothing to stage
    mbar_arrive(bar);
    return;
  }
  const int e_lo = d.e0 & ~15;
  const uint32_t n_e = (uint32_t)((d.e1 - e_lo + 15) & ~15);
  const int f_lo = d.f0 & ~3;
  const uint32_t n_rp = (uint32_t)((d.f1 + 1 - f_lo + 3) & ~3);
  const uint32_t n_cf = (uint32_t)((d.f1 - f_lo + 1) & ~1);
#ifdef RDIS_EXP_NOPARAMS  // timing experiment: what the sweep costs when the per-edge parameters do not cross HBM
  mbar_arrive_expect_tx(bar, n_e * 5u + n_rp * 4u + n_cf * 8u);
#else
  mbar_arrive_expect_tx(bar, n_e * 21u + n_rp * 4u + n_cf * 8u);
  tma_bulk_g2s(st.expo, G.expo + e_lo, n_e * 8u, bar, pol);
  tma_bulk_g2s(st.konst, G.konst + e_lo, n_e * 8u, bar, pol);
#endif
  tma_bulk_g2s(st.evid, G.evid + e_lo, n_e * 4u, bar, pol);
  tma_bulk_g2s(st.sine, G.sine + e_lo, n_e, bar, pol);
  tma_bulk_g2s(st.rowptr, G.rowptr + f_lo, n_rp * 4u, bar, pol);
  tma_bulk_g2s(st.coeff, G.coeff + f_lo, n_cf * 8u, bar, pol);
}

// Named barrier over the consumer warps only (the producer warp never joins it).
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTileThreads) : "memory"); }

// Block layout: warps 0..kTileThreads/32-1 = consumers, then the bulk-copy warp, then the gather warp.
//   bulk-copy warp   lane 0 launches the six TMA bulk copies of a tile as soon as its stage has been released;
//   gather warp      once a slice has landed (its `full` barrier), all 32 lanes gather the tile's variable values into
//                    the stage with 8-byte cp.async copies from the dense value mirror (xval) and signal `xready`.
// The two run on separate warps because each blocks on a different event: one warp doing both issues the gathers of
// tile j only after the consumers have released the stage of tile j+1's copy, which caps a CTA at one tile per gather
// latency (ncu of that form: 39 % of all warp samples on the consumers' `xready` wait).
// Consumers find everything in shared memory: no register staging, no exposed gather latency.  Terms are written IN
// PLACE over the stage's exponent slots (derivatives over the constant slots): the thread that consumes expo[e] /
// konst[e] is the one that produces term[e], so a tile needs exactly one CTA-wide barrier.
// Precondition (holds for every stream-ordered caller): no subspace solve is running on this context,
// i.e. every variable is frozen and xval mirrors xbd.x.
template <bool kGrad>
__global__ void __launch_bounds__(kTileBlock, kGrad ? RDIS_TILE_CTAS_GRAD : kTileCtasPerSm)
    nlpf_tile_sweep_kernel(GraphView G, const TileDesc* __restrict__ tiles, int ntiles, double* __restrict__ per_factor,
                           double* partials, unsigned int* counter, double* sum_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<kGrad>& S = *reinterpret_cast<TileSmem<kGrad>*>(smem_raw);
  const int tid = threadIdx.x;
  const int stride = gridDim.x;
  const int my_tiles = (ntiles - (int)blockIdx.x + stride - 1) / stride;

  if (tid == 0) {
    for (int s = 0; s < kTileStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.xready[s], 32);                // one arrival per producer lane
      mbar_init(&S.empty[s], kTileThreads / 32);  // one arrival per consumer warp
    }
    mbar_fence_init();
  }
  __syncthreads();

  double acc = 0.0;
  if (tid >= kTileThreads + 32) {
    // ---- gather warp ----  waits only for a slice to land (never for the consumers), so the gathers of tile j+1 are
    // issued while those of tile j may still be in flight
    const int lane = tid & 31;
    const uint64_t keep = l2_policy_evict_last();
    for (int g = 0; g < my_tiles; ++g) {
      const int sg = g % kTileStages;
      TileStage& st = S.stage[sg];
      mbar_wait(&S.full[sg], (uint32_t)(g / kTileStages) & 1u);
      const TileDesc dg = st.desc;
      const int ne = dg.e1 - dg.e0, eo = dg.e0 & 15;
      if (ne <= kTileEdges) {
#ifndef RDIS_EXP_NOGATHER
        for (int le = lane; le < ne; le += 32) cp_async_gather8(&st.xs[eo + le], G.xval + st.evid[eo + le], keep);
#endif
      }
      cp_async_arrive_noinc(&S.xready[sg]);  // fires when this lane's copies have landed
    }
  } else if (tid >= kTileThreads) {
    // ---- bulk-copy warp ----
    // Descriptors are fetched 32 at a time (one per lane, the next batch already in flight), so the
    // issue loop never waits on HBM for a descriptor.
    const int lane = tid & 31;
    auto fetch = [&](int batch) {
      const int i = batch * 32 + lane;
      return tiles[(i < my_tiles) ? (blockIdx.x + i * stride) : 0];
    };
    TileDesc mine = fetch(0), ahead = fetch(1);
    const uint64_t pol = l2_policy_evict_first();
    for (int j = 0; j < my_tiles; ++j) {
      if (j > 0 && (j & 31) == 0) {
        mine = ahead;
        ahead = fetch((j >> 5) + 1);
      }
      TileDesc d;
      d.f0 = __shfl_sync(0xffffffffu, mine.f0, j & 31);
      d.f1 = __shfl_sync(0xffffffffu, mine.f1, j & 31);
      d.e0 = __shfl_sync(0xffffffffu, mine.e0, j & 31);
      d.e1 = __shfl_sync(0xffffffffu, mine.e1, j & 31);
      if (lane == 0) {
        const int s = j % kTileStages;
        if (j >= kTileStages) mbar_wait_backoff(&S.empty[s], (uint32_t)(j / kTileStages - 1) & 1u);  // stage released
        S.stage[s].desc = d;
        tile_issue(G, d, S.stage[s], &S.full[s], pol);
      }
      __syncwarp();
    }
  } else {
    // ---- consumers ----
    const int lane = tid & 31;
    // kGrad: the previous tile's partials sit in gout[(it - 1) & 1]; thread 0 stores them once a consumer barrier has
    // ordered every thread's phase 2 before it — the barrier of the NEXT tile, so no extra barrier is paid.
    int pend_e0 = 0, pend_ne = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % kTileStages;
      const uint32_t parity = (uint32_t)(it / kTileStages) & 1u;
      TileStage& st = S.stage[s];
      mbar_wait(&S.xready[s], parity);
      mbar_wait(&S.full[s], parity);  // already complete (the producer waited on it); makes the TMA writes visible here
      const TileDesc d = st.desc;
      const int ne = d.e1 - d.e0;

      if (ne > kTileEdges) {  // one over-wide factor, folded serially from global memory
        if (kGrad) {
          consumer_barrier();
          if (tid == 0 && pend_ne > 0) tile_store_partials(G.gedge, S.gout[(it - 1) & 1], pend_e0, pend_ne);
          pend_ne = 0;
        }
        if (tid == 0) {
          double fv;
          if (kGrad) {
            fv = NlpfOps::gradient(G, d.f0, G.gedge + d.e0);
          } else {
            double sl;
            fv = NlpfOps::value<false>(G, d.f0, 0.0, false, sl);
          }
          if (G.fconst_on != nullptr && G.fconst_on[d.f0]) fv = G.fconst_val[d.f0];
          if (per_factor) per_factor[d.f0] = fv;
          acc += fv;
        }
      } else {
        // ---- phase 1: staged edge slices -> terms (in place) ----
        double* term = st.expo + (d.e0 & 15);   // term[le] overwrites expo[le]
        double* dterm = st.konst + (d.e0 & 15);  // dterm[le] overwrites konst[le]
        const double* xs = st.xs + (d.e0 & 15);
        const uint8_t* sine = st.sine + (d.e0 & 15);
#pragma unroll
        for (int i = 0; i < kEdgeBatch; ++i) {
          const int le = i * kTileThreads + tid;
          if (le < ne) {
#ifdef RDIS_EXP_NOGATHER
            const double xv = (double)st.evid[(d.e0 & 15) + le];
#else
            const double xv = xs[le];
#endif
#ifdef RDIS_EXP_NOPARAMS
            const double ex = 1.0, kk = 0.0;
#else
            const double ex = term[le], kk = dterm[le];
#endif
            const bool sn = sine[le] != 0;
            if (kGrad) {
              double tv, dt;
              nlpf_term_grad(xv, kk, ex, sn, tv, dt);
              term[le] = tv;
              dterm[le] = dt;
            } else {
              term[le] = nlpf_term_value(xv, kk, ex, sn);
            }
          }
        }
        if (kGrad && tid == 0) bulk_wait_read_all();  // the store issued one tile ago has left gout[it & 1]
        consumer_barrier();  // the only CTA-wide wait of a tile
        if (kGrad) {
          if (tid == 0 && pend_ne > 0) tile_store_partials(G.gedge, S.gout[(it - 1) & 1], pend_e0, pend_ne);
          pend_e0 = d.e0;
          pend_ne = ne;
        }
        double* const gst = S.gout[kGrad ? (it & 1) : 0] + (d.e0 & 1);
        // ---- phase 2: factors ----
        const int fo = d.f0 & 3;
        const int nfac = d.f1 - d.f0;
        for (int lf = tid; lf < nfac; lf += kTileThreads) {
          const int32_t f = d.f0 + lf;
          const int r0 = st.rowptr[fo + lf] - d.e0;
          const int n = st.rowptr[fo + lf + 1] - d.e0 - r0;
          const double c = st.coeff[fo + lf];
          double prod = 1.0;
          if (n <= 4) {  // slot-order product, branch-free: a missing slot multiplies by 1.0 (exact)
            const double t0 = (n > 0) ? term[r0] : 1.0, t1 = (n > 1) ? term[r0 + 1] : 1.0;
            const double t2 = (n > 2) ? term[r0 + 2] : 1.0, t3 = (n > 3) ? term[r0 + 3] : 1.0;
            prod = (((prod * t0) * t1) * t2) * t3;
            if (kGrad) {  // getDerivative: product in slot order, own slot replaced by its derivative
              const double d0 = (n > 0) ? dterm[r0] : 1.0, d1 = (n > 1) ? dterm[r0 + 1] : 1.0;
              const double d2 = (n > 2) ? dterm[r0 + 2] : 1.0, d3 = (n > 3) ? dterm[r0 + 3] : 1.0;
              double* ge = gst + r0;
              if (n > 0) ge[0] = ((((1.0 * d0) * t1) * t2) * t3) * c;
              if (n > 1) ge[1] = ((((1.0 * t0) * d1) * t2) * t3) * c;
              if (n > 2) ge[2] = ((((1.0 * t0) * t1) * d2) * t3) * c;
              if (n > 3) ge[3] = ((((1.0 * t0) * t1) * t2) * d3) * c;
            }
          } else {
            for (int r = r0; r < r0 + n; ++r) prod *= term[r];
            if (kGrad) {
              for (int i = r0; i < r0 + n; ++i) {
                double pe = 1.0;
                for (int j = r0; j < r0 + n; ++j) pe *= (j == i) ? dterm[j] : term[j];
                gst[i] = pe * c;
              }
            }
          }
          double fv = prod * c;
          if (G.fconst_on != nullptr && G.fconst_on[f]) fv = G.fconst_val[f];  // Factor::eval, src/Factor.cpp:110-119
          if (per_factor) __stcs(&per_factor[f], fv);  // streaming store: written once, not re-read by the sweep
          acc += fv;
        }
      }
      if (kGrad) fence_proxy_async_smem();  // this thread's partials, before the barrier that precedes their bulk store
      // this warp is done with stage s: release it to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[s]);
    }
    if (kGrad) {
      consumer_barrier();
      if (tid == 0) {
        if (pend_ne > 0) tile_store_partials(G.gedge, S.gout[(my_tiles - 1) & 1], pend_e0, pend_ne);
        bulk_wait_all();  // performed before the kernel ends: the gather launch reads gedge
      }
    }
  }
  block_then_grid_sum(acc, partials, counter, sum_out);
}

}  // namespace rdisgpu
