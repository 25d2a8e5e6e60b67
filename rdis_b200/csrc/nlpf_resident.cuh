// nlpf_resident.cuh — CGDSubspaceOptimizer::optimize for a mid-size NonlinearProductFactor component that is
// RESIDENT in one CTA's shared memory for the whole solve (BASELINE config 4's sibling wave: 1024 components of
// 1023 variables / 4092 factors each).
//
// The generic block kernel (solve_kernels.cuh) evaluates every line-search point from HBM/L2 state: per edge it
// gathers the variable (16 B), its domain (16 B), the exponent, the constant and the sine flag, clamps, and
// evaluates [sin]((x-k)^e) — 434 KB of L2 traffic and 8184 transcendental terms per evaluation of such a
// component, ~445 evaluations per solve.  Here the component is flattened once, at the start of the solve, into
//   variables   p, xi, lb, ub                                                (4 x 8 B per variable)
//   TERMS       the distinct (variable, k, e, sine) tuples of the component  (value, derivative, e, k, flag, owner)
//   edges       u16 index of their term + u16 slot of their variable         (4 B per edge)
//   factors     coefficient + first edge                                     (12 B per factor)
// and a line evaluation is two shared-memory passes: (1) one thread per TERM clamps p + alpha*xi and evaluates the
// term (and its derivative when the slope is wanted), (2) one thread per FACTOR folds the product in slot order.
// A variable that enters many factors through the same expression (sin(x_v) appears in every factor of the
// sinusoid family that touches v) is evaluated once per point instead of once per edge: the term table is built
// on the host at finalize (rdisgpu_finalize: tvrow / t_* / eterm) by exact comparison of (k, e, flag), so a shared
// term is the SAME number the per-edge evaluation would produce, and the per-factor arithmetic is expression for
// expression NlpfOps::value's.  With the same thread count the factor -> thread mapping and the reduction order
// are the generic kernel's too, so the two kernels publish bit-identical objective / slope sequences
// (tests/test_gpu_parity.py demands equality of the whole result, not a tolerance).
// Edges on frozen variables (assigned ancestors) become constant terms evaluated once.
// Terms are laid out by class — all sine terms first, then the rest — so a warp of pass 1 runs ONE of the two
// bodies (the sincos chain or the few-instruction polynomial one) instead of both under divergence, and pass 2
// dispatches on the factor's arity to straight-line code (the generators emit factors arity-major, so warps are
// uniform there too).
// Full gradients (one per CG iteration) take the same two passes with the per-edge partials parked in an
// L2-resident scratch slice of the problem, then a third pass, one thread per variable, folds the variable's
// incident partials in ascending factor id (productGradient's order, src/State.h:157-194) through a
// shared-memory incidence list built once at the start of the solve — no index chasing per gradient.
//
// Reference semantics: CGDSubspaceOptimizer::optimize src/optimizers/CGDSubspaceOptimizer.cpp:19-98,
// SubfunctionFD::operator()/df :124-157, quickAssignVals :160-184, NonlinearProductFactor::evalFactor /
// getDerivative src/NonlinearProductFactor.cpp:186-209,149-178.
#pragma once
#include "nlpf_tile_sweep.cuh"
#include "solve_kernels.cuh"

namespace rdisgpu {

// CTA width.  The passes are issue- and barrier-bound, so the CTA is as wide as the register file allows: with the
// state machine in shared memory the kernel needs 64 registers and runs 1024 threads (cfg4 wave: 256 threads 36 ms,
// 512 threads 16.2 ms, 768 15.3 ms, 1024 15.1 ms).  256 threads is the generic block kernel's widest CTA — the same
// factor -> thread mapping and reduction tree, hence bit-identical results — and stays selectable
// (rdisgpu_set_option "resident_threads") for the equality test.
#ifndef RDIS_RES_THREADS
#define RDIS_RES_THREADS 1024
#endif
constexpr int kResThreads = RDIS_RES_THREADS;
constexpr int kResThreadsExact = 256;
// Components whose layout leaves room for at least two CTAs on the SM run 256 threads wide, up to four CTAs to the SM
// (64 registers): half of an evaluation is a serial path (barriers, the scalar line-search step), and a second
// resident CTA fills it.  (Level-13 subtrees of config 4, 127 variables: 18.0 ms per 8192 solves at 128 threads,
// 16.4 ms at 256.)
#ifndef RDIS_RES_SMALL_THREADS
#define RDIS_RES_SMALL_THREADS 256
#endif
constexpr int kResThreadsSmall = RDIS_RES_SMALL_THREADS;
constexpr int kResSmallCtas = 4;
constexpr int kResSmallFactors = 4096;       // at most this many factors ...
constexpr int kResSmallSmem = 110 * 1024;    // ... and this much shared memory (two CTAs per SM)

// Shared-memory carve-up, computed identically on the host (fits? how many bytes to ask for) and in the kernel.
struct ResLayout {
  int xs, ds, lb, ub, tval, tdt, texpo, tkonst, fcoef;  // double arrays (byte offsets)
  int toff, frow;                                        // int32
  int tlv, elt, elv;                                     // uint16
  int total;
};
__host__ __device__ inline ResLayout res_layout(int nv, int nf, int nE, int nT, int nFz) {
  ResLayout L;
  int o = 0;
  auto take = [&](int bytes) {
    const int at = o;
    o += (bytes + 15) & ~15;
    return at;
  };
  L.xs = take(8 * nv); L.ds = take(8 * (nv + 1)); L.lb = take(8 * nv); L.ub = take(8 * nv);
  L.tval = take(8 * (nT + nFz)); L.tdt = take(8 * (nT + nFz));
  L.texpo = take(8 * nT); L.tkonst = take(8 * nT);
  L.fcoef = take(8 * nf);
  L.toff = take(4 * (nv + 1)); L.frow = take(4 * (nf + 1));
  L.tlv = take(2 * nT); L.elt = take(2 * nE); L.elv = take(2 * nE);
  L.total = o;
  return L;
}

// In-place exclusive scan of a[0..n) by the whole CTA; a[n] receives the total.  `wtot` = 32 ints of shared scratch.
__device__ __forceinline__ void block_exclusive_scan(int32_t* a, int n, int32_t* wtot) {
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (T + 31) >> 5;
  const int chunk = (n + T - 1) / T;
  const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
  int32_t mine = 0;
  for (int i = lo; i < hi; ++i) mine += a[i];
  int32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  int32_t base = 0;
  for (int w = 0; w < warp; ++w) base += wtot[w];
  int32_t run = base + incl - mine;
  for (int i = lo; i < hi; ++i) {
    const int32_t v = a[i];
    a[i] = run;
    run += v;
  }
  if (tid == T - 1) {
    int32_t tot = 0;
    for (int w = 0; w < nw; ++w) tot += wtot[w];
    a[n] = tot;
  }
  __syncthreads();
}

struct ResView {
  double *xs, *ds, *lb, *ub, *tval, *tdt, *texpo, *tkonst, *fcoef;
  int32_t *toff, *frow;
  uint16_t *tlv, *elt, *elv;
  int nT, nS;            // own terms; the first nS of them are the sine class
  double* gscr;          // this problem's slice of the per-edge partial scratch (HBM/L2), indexed by local edge
  const uint16_t* vinc;  // ... and of the variable-major incidence list (local edge ids, ascending factor id)
};

// Sum of two doubles per thread over the CTA, delivered to WARP 0 only (the state machine's thread is its consumer).
// kExact: Block::reduce's order — warp butterfly, warp partials folded 0..nw-1 — so a 256-thread CTA reproduces the
// generic kernel's totals to the bit.  Otherwise the warp partials are folded by a second butterfly (5 shuffle steps
// instead of a serial walk over up to 32 partials on the critical path of every evaluation).  buf = 2 x 64 doubles.
template <bool kExact>
__device__ __forceinline__ void resident_sum2(double* buf2, int& flip, double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  double* buf = buf2 + flip * 64;
  flip ^= 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    buf[2 * warp] = a;
    buf[2 * warp + 1] = b;
  }
  __syncthreads();
  if (warp != 0) return;
  if (kExact) {
    a = buf[0];
    b = buf[1];
    for (int w = 1; w < nw; ++w) {
      a += buf[2 * w];
      b += buf[2 * w + 1];
    }
  } else {
    a = (lane < nw) ? buf[2 * lane] : 0.0;
    b = (lane < nw) ? buf[2 * lane + 1] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}

// One factor of arity N (compile-time): value term product in slot order and, with kSlope, the directional
// derivative sum_i (d f/d x_i) xi_i — NlpfOps::value's expressions (1.0 * t == t, so the leading 1.0 is dropped).
template <int N, bool kSlope>
__device__ __forceinline__ void resident_fold(const ResView& R, int r0, double c, double& prod, double& s) {
  double t[N], dt[N], dir[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int lt = R.elt[r0 + i];
    t[i] = R.tval[lt];
    if (kSlope) {
      dt[i] = R.tdt[lt];
      dir[i] = R.ds[R.elv[r0 + i]];  // slot nv holds 0.0: frozen variable
    }
  }
  prod = t[0];
#pragma unroll
  for (int i = 1; i < N; ++i) prod *= t[i];
  s = 0.0;
  if (kSlope) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (dir[i] != 0.0) {
        double pe = (i == 0) ? dt[0] : t[0];  // getDerivative: own slot replaced by its derivative (1.0 when plain)
#pragma unroll
        for (int j = 1; j < N; ++j) pe *= (j == i) ? dt[j] : t[j];
        s = s + (pe * c) * dir[i];
      }
    }
  }
}

// Pass 1 of an evaluation: every own term at x = clamp(p + alpha*xi).  kGrad: also the own-slot derivative.
template <bool kGrad>
__device__ __forceinline__ void resident_terms(const ResView& R, double alpha, bool along) {
  const int T = blockDim.x, tid = threadIdx.x;
  for (int t = tid; t < R.nS; t += T) {  // sine class
    const int lv = R.tlv[t];
    const double raw = along ? __dadd_rn(R.xs[lv], __dmul_rn(alpha, R.ds[lv])) : R.xs[lv];  // Df1dim: p + x*xi, product rounded first (load_var)
    const double xv = clamp_to_domain(raw, make_double2(R.lb[lv], R.ub[lv]));
    if (kGrad) {
      double tv, dt;
      nlpf_term_grad(xv, R.tkonst[t], R.texpo[t], true, tv, dt);
      R.tval[t] = tv;
      R.tdt[t] = dt;
    } else {
      R.tval[t] = nlpf_term_value(xv, R.tkonst[t], R.texpo[t], true);
    }
  }
  for (int t = R.nS + tid; t < R.nT; t += T) {  // polynomial class
    const int lv = R.tlv[t];
    const double raw = along ? __dadd_rn(R.xs[lv], __dmul_rn(alpha, R.ds[lv])) : R.xs[lv];  // Df1dim: p + x*xi, product rounded first (load_var)
    const double xv = clamp_to_domain(raw, make_double2(R.lb[lv], R.ub[lv]));
    if (kGrad) {
      double tv, dt;
      nlpf_term_grad(xv, R.tkonst[t], R.texpo[t], false, tv, dt);
      R.tval[t] = tv;
      R.tdt[t] = dt;
    } else {
      R.tval[t] = nlpf_term_value(xv, R.tkonst[t], R.texpo[t], false);
    }
  }
  __syncthreads();
}

// One line evaluation: f(p + alpha*xi) and, with kSlope, d/dalpha — the shared-memory form of objective_along_line.
template <bool kSlope, bool kExact>
__device__ __forceinline__ void resident_line_eval(const GraphView& G, const ResView& R, double* buf2, int& flip,
                                                   const int32_t* fids, int nf, double alpha, double& f, double& slope) {
  const int T = blockDim.x, tid = threadIdx.x;
  resident_terms<kSlope>(R, alpha, true);
  // ---- pass 2: factors, thread k owns factors k, k+T, ... (objective_along_line's mapping) ----
  double fs = 0.0, ss = 0.0;
  for (int k = tid; k < nf; k += T) {
    const int r0 = R.frow[k] & 0x7fffffff;
    const int n = (R.frow[k + 1] & 0x7fffffff) - r0;
    const bool is_const = R.frow[k] < 0;  // Factor::eval of an assigned constant (src/Factor.cpp:110-119)
    const double c = R.fcoef[k];
    double prod = 1.0, s = 0.0;
    switch (n) {
      case 0: break;
      case 1: resident_fold<1, kSlope>(R, r0, c, prod, s); break;
      case 2: resident_fold<2, kSlope>(R, r0, c, prod, s); break;
      case 3: resident_fold<3, kSlope>(R, r0, c, prod, s); break;
      case 4: resident_fold<4, kSlope>(R, r0, c, prod, s); break;
      default: {
        for (int i = 0; i < n; ++i) prod *= R.tval[R.elt[r0 + i]];
        if (kSlope) {
          for (int i = 0; i < n; ++i) {
            const double diri = R.ds[R.elv[r0 + i]];
            if (diri != 0.0) {
              double pe = 1.0;
              for (int j = 0; j < n; ++j) {
                const int lj = R.elt[r0 + j];
                pe *= (j == i) ? R.tdt[lj] : R.tval[lj];
              }
              s = s + (pe * c) * diri;
            }
          }
        }
      }
    }
    double fv = prod * c;
    if (is_const) fv = G.fconst_val[fids[k]];
    fs += fv;
    ss += s;
  }
  resident_sum2<kExact>(buf2, flip, fs, ss);  // totals in warp 0; its barrier also orders this evaluation's reads of tval before the next pass 1
  f = fs;
  slope = ss;
}

// Per-edge partials of one factor of arity N into the scratch slice (NlpfOps::gradient's expressions).
template <int N>
__device__ __forceinline__ double resident_partials(const ResView& R, int r0, double c) {
  double t[N], dt[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int lt = R.elt[r0 + i];
    t[i] = R.tval[lt];
    dt[i] = R.tdt[lt];
  }
  double prod = t[0];
#pragma unroll
  for (int i = 1; i < N; ++i) prod *= t[i];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double pe = (i == 0) ? dt[0] : t[0];
#pragma unroll
    for (int j = 1; j < N; ++j) pe *= (j == i) ? dt[j] : t[j];
    R.gscr[r0 + i] = pe * c;
  }
  return prod;
}

// Value and full gradient at p (Factor::computeGradient of every factor + computeGradientOfSum restricted to the
// problem's variables): returns this thread's partial of the objective (caller reduces it when it needs f);
// thread j-strided gradient entries are handed to `sink(j, dF/dx_j)`.  R.toff holds the incidence offsets here.
template <class Sink>
__device__ __forceinline__ double resident_gradient(const GraphView& G, const ResView& R, const int32_t* fids, int nv, int nf,
                                                    Sink sink) {
  const int T = blockDim.x, tid = threadIdx.x;
  resident_terms<true>(R, 0.0, false);  // load_var<false>: x = clamp(p)
  double fs = 0.0;
  for (int k = tid; k < nf; k += T) {
    const int r0 = R.frow[k] & 0x7fffffff;
    const int n = (R.frow[k + 1] & 0x7fffffff) - r0;
    const bool is_const = R.frow[k] < 0;
    const double c = R.fcoef[k];
    double prod = 1.0;
    switch (n) {
      case 0: break;
      case 1: prod = resident_partials<1>(R, r0, c); break;
      case 2: prod = resident_partials<2>(R, r0, c); break;
      case 3: prod = resident_partials<3>(R, r0, c); break;
      case 4: prod = resident_partials<4>(R, r0, c); break;
      default:
        for (int i = 0; i < n; ++i) {
          prod *= R.tval[R.elt[r0 + i]];
          double pe = 1.0;
          for (int j = 0; j < n; ++j) {
            const int lj = R.elt[r0 + j];
            pe *= (j == i) ? R.tdt[lj] : R.tval[lj];
          }
          R.gscr[r0 + i] = pe * c;
        }
    }
    double fv = prod * c;
    if (is_const) fv = G.fconst_val[fids[k]];
    fs += fv;
  }
  __syncthreads();  // the partials of every factor are visible to the CTA
  for (int j = tid; j < nv; j += T) {
    const int a0 = R.toff[j], a1 = R.toff[j + 1];
    double acc = 0.0;
    for (int a = a0; a < a1; a += 4) {  // four partials in flight, folded in list order (gather_var's filter path)
      double ge[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ge[i] = (a + i < a1) ? R.gscr[R.vinc[a + i]] : 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a + i < a1) acc = (a + i == a0) ? ge[i] : acc + ge[i];
      }
    }
    sink(j, acc);
  }
  return fs;
}

// grid = number of resident-class problems, one CTA each; dynamic shared memory = the largest layout of the class.
template <int kThreads, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas)
    solve_nlpf_resident_kernel(GraphView G, BatchView B, const int32_t* order, int count, int maxiters, double ftol) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  __shared__ double scratch[260];
  __shared__ double buf2[128];
  __shared__ int32_t wtot[32];
  __shared__ int32_t nfz_counter;
  if ((int)blockIdx.x >= count) return;
  const int pidx = order[blockIdx.x];
  const ProblemDesc P = B.probs[pidx];
  const int32_t* vids = B.vids + P.var_off;
  const int32_t* fids = B.fids + P.fac_off;
  const int nv = P.nv, nf = P.nf;
  const int32_t stamp = pidx;
  const int T = blockDim.x, tid = threadIdx.x;
  Block grp(scratch);
  int flip2 = 0;

  const ResLayout L = res_layout(nv, nf, P.nE, P.nT, P.nFz);
  ResView R;
  R.xs = reinterpret_cast<double*>(res_smem + L.xs); R.ds = reinterpret_cast<double*>(res_smem + L.ds);
  R.lb = reinterpret_cast<double*>(res_smem + L.lb); R.ub = reinterpret_cast<double*>(res_smem + L.ub);
  R.tval = reinterpret_cast<double*>(res_smem + L.tval); R.tdt = reinterpret_cast<double*>(res_smem + L.tdt);
  R.texpo = reinterpret_cast<double*>(res_smem + L.texpo); R.tkonst = reinterpret_cast<double*>(res_smem + L.tkonst);
  R.fcoef = reinterpret_cast<double*>(res_smem + L.fcoef);
  R.toff = reinterpret_cast<int32_t*>(res_smem + L.toff); R.frow = reinterpret_cast<int32_t*>(res_smem + L.frow);
  R.tlv = reinterpret_cast<uint16_t*>(res_smem + L.tlv); R.elt = reinterpret_cast<uint16_t*>(res_smem + L.elt);
  R.elv = reinterpret_cast<uint16_t*>(res_smem + L.elv);
  R.gscr = B.gscr + P.goff;
  uint16_t* vinc = B.gvinc + P.goff;
  R.vinc = vinc;
  R.nT = P.nT;

  // ---- claim (CGD.cpp:33: quickAssignVals of the start point) + flatten ----
  if (tid == 0) {
    nfz_counter = 0;
    R.ds[nv] = 0.0;  // the "frozen variable" direction slot
  }
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    const double xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : G.xbd[vid].x;
    G.xsave[vid] = xv;
    G.xbd[vid] = make_double2(xv, 0.0);
    G.vloc[vid] = j;
    const double2 dm = __ldg(&G.dom[vid]);
    R.xs[j] = xv; R.ds[j] = 0.0; R.lb[j] = dm.x; R.ub[j] = dm.y;
    int ns = 0, np = 0;  // sine-class / polynomial-class terms of this variable, packed (each total < 65536)
    for (int32_t g = __ldg(&G.tvrow[vid]); g < __ldg(&G.tvrow[vid + 1]); ++g) {
      if (__ldg(&G.t_sine[g])) ++ns; else ++np;
    }
    R.toff[j] = ns | (np << 16);
  }
  for (int k = tid; k < nf; k += T) {
    const int32_t fid = fids[k];
    G.fstamp[fid] = stamp;
    G.floc[fid] = k;
    R.frow[k] = __ldg(&G.rowptr[fid + 1]) - __ldg(&G.rowptr[fid]);
    R.fcoef[k] = __ldg(&G.coeff[fid]);
  }
  __syncthreads();
  block_exclusive_scan(R.toff, nv, wtot);
  block_exclusive_scan(R.frow, nf, wtot);
  const int nS = R.toff[nv] & 0xffff;
  R.nS = nS;
  // term descriptors of the component's own variables: sine class in [0, nS), the rest in [nS, nT)
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    int a = R.toff[j] & 0xffff, b = nS + (int)((uint32_t)R.toff[j] >> 16);
    for (int32_t g = __ldg(&G.tvrow[vid]); g < __ldg(&G.tvrow[vid + 1]); ++g) {
      const int slot = __ldg(&G.t_sine[g]) ? a++ : b++;
      R.tlv[slot] = (uint16_t)j;
      R.texpo[slot] = __ldg(&G.t_expo[g]);
      R.tkonst[slot] = __ldg(&G.t_konst[g]);
    }
  }
  // edges: own variables -> their term; frozen variables -> a constant term evaluated here, once
  for (int k = tid; k < nf; k += T) {
    const int32_t fid = fids[k];
    const int32_t e0 = __ldg(&G.rowptr[fid]);
    const int r0 = R.frow[k], n = R.frow[k + 1] - r0;
    for (int i = 0; i < n; ++i) {
      const int32_t e = e0 + i;
      const int32_t vid = __ldg(&G.evid[e]);
      const double2 xb = G.xbd[vid];
      if (xb.y != xb.y) {  // frozen (load_var)
        const int slot = P.nT + atomicAdd(&nfz_counter, 1);
        R.tval[slot] = nlpf_term_value(xb.x, __ldg(&G.konst[e]), __ldg(&G.expo[e]), __ldg(&G.sine[e]) != 0);
        R.tdt[slot] = 0.0;
        R.elt[r0 + i] = (uint16_t)slot;
        R.elv[r0 + i] = (uint16_t)nv;
      } else {
        const int lv = G.vloc[vid];
        const int32_t g0 = __ldg(&G.tvrow[vid]), gt = __ldg(&G.eterm[e]);
        const bool sn = __ldg(&G.t_sine[gt]) != 0;
        int rank = 0;  // terms of the same class that precede this one in the variable's list
        for (int32_t g = g0; g < gt; ++g) rank += ((__ldg(&G.t_sine[g]) != 0) == sn);
        const int base = sn ? (R.toff[lv] & 0xffff) : nS + (int)((uint32_t)R.toff[lv] >> 16);
        R.elt[r0 + i] = (uint16_t)(base + rank);
        R.elv[r0 + i] = (uint16_t)lv;
      }
    }
  }
  __syncthreads();
  // variable-major incidence of the component (local edge ids, ascending factor id = the order of the global
  // incidence list) into the problem's scratch slice: toff is free now and becomes its offsets
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    int cnt = 0;
    for (int32_t r = __ldg(&G.vrow[vid]); r < __ldg(&G.vrow[vid + 1]); ++r)
      cnt += (G.fstamp[__ldg(&G.efac[__ldg(&G.vedge[r])])] == stamp);
    R.toff[j] = cnt;
  }
  __syncthreads();
  block_exclusive_scan(R.toff, nv, wtot);
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    int pos = R.toff[j];
    for (int32_t r = __ldg(&G.vrow[vid]); r < __ldg(&G.vrow[vid + 1]); ++r) {
      const int32_t e = __ldg(&G.vedge[r]);
      const int32_t f = __ldg(&G.efac[e]);
      if (G.fstamp[f] != stamp) continue;
      vinc[pos++] = (uint16_t)(R.frow[G.floc[f]] + (e - __ldg(&G.rowptr[f])));
    }
  }
  __syncthreads();
  // assigned-constant factors: flag in the sign bit of the row start (values fetched from HBM when it is set)
  if (G.fconst_on != nullptr) {
    for (int k = tid; k < nf; k += T) {
      if (G.fconst_on[fids[k]]) R.frow[k] |= (int32_t)0x80000000;
    }
    __syncthreads();
  }

  // The CG / line-search state machine lives in shared memory and is advanced by ONE thread; everybody else reads the
  // published request.  (Replicating it in every thread, as the generic kernels do, costs 13 % of this kernel's
  // instructions and ~70 registers per thread — the difference between 16 and 32 resident warps.)  Every branch below
  // has a CTA barrier between the reads at the top of the loop and thread 0's update, and the loop's own barrier
  // publishes the update.
  __shared__ CgdMachine m;
  __shared__ double s_f_init;
  if (tid == 0) m.start(maxiters, ftol);
  for (;;) {
    __syncthreads();
    const int req = m.req;
    if (req == REQ_DONE) break;
    const double alpha = m.alpha, gam = m.gam;
    if (req == REQ_INIT_GRAD) {
      double fs = resident_gradient(G, R, fids, nv, nf, [&](int j, double gr) {
        const int32_t vid = vids[j];
        const double gneg = -gr;
        G.gvec[vid] = gneg;
        G.hvec[vid] = gneg;
        R.ds[j] = gneg;
      });
      double zero = 0.0;
      grp.sum2(fs, zero);
      if (tid == 0) {
        s_f_init = fs;
        m.on_init(fs);
      }
    } else if (req == REQ_VALUE || req == REQ_VALUE_SLOPE) {
      double f, sl;
      constexpr bool kExact = (kThreads == kResThreadsExact);
      if (req == REQ_VALUE_SLOPE)
        resident_line_eval<true, kExact>(G, R, buf2, flip2, fids, nf, alpha, f, sl);
      else
        resident_line_eval<false, kExact>(G, R, buf2, flip2, fids, nf, alpha, f, sl);
      if (tid == 0) m.on_eval(f, sl);
    } else if (req == REQ_MOVE) {
      for (int j = tid; j < nv; j += T) {  // minimize_nrc.h:508-511
        double2 xb = make_double2(R.xs[j], R.ds[j]);
        xb.y *= alpha;
        xb.x += xb.y;
        R.xs[j] = xb.x; R.ds[j] = xb.y;
      }
      __syncthreads();
      if (tid == 0) m.on_moved();
    } else if (req == REQ_GRADIENT) {
      double gg = 0.0, dgg = 0.0, dummy = 0.0, tnum = 0.0;
      (void)resident_gradient(G, R, fids, nv, nf, [&](int j, double gr) {
        R.ds[j] = gr;  // func.df(p, xi), :654
        const double pj = fabs(R.xs[j]);
        const double t = fabs(gr) * ((pj < 1.0) ? 1.0 : pj);  // :659 numerator
        tnum = (t > tnum) ? t : tnum;
        const double gj = G.gvec[vids[j]];
        gg += gj * gj;
        dgg += (gr + gj) * gr;
      });
      grp.sum3max(gg, dgg, dummy, tnum);
      if (tid == 0) m.on_gradient(tnum, gg, dgg);
    } else {  // REQ_DIRECTION, :681-685
      for (int j = tid; j < nv; j += T) {
        const int32_t vid = vids[j];
        const double gj = -R.ds[j];
        const double hj = gj + gam * G.hvec[vid];
        G.gvec[vid] = gj;
        G.hvec[vid] = hj;
        R.ds[j] = hj;
      }
      __syncthreads();
      if (tid == 0) m.on_directed();
    }
  }
  const double f_init = s_f_init;

  // ---- commit (CGD.cpp:61-89) ----
  double fret = m.fret;
  bool restore = (fret > f_init);
  // the safety exits keep p / fret of the last completed line search, like the reference's own throws (CGD.cpp:41-61)
  if (restore) fret = f_init;
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    const double raw = restore ? G.xsave[vid] : R.xs[j];
    const double val = clamp_to_domain(raw, make_double2(R.lb[j], R.ub[j]));
    G.xbd[vid] = make_double2(val, qnan_f64());
    G.xval[vid] = val;
    B.xout[P.var_off + j] = val;
  }
  for (int k = tid; k < nf; k += T) G.fstamp[fids[k]] = -1;
  if (tid == 0) {
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

}  // namespace rdisgpu
