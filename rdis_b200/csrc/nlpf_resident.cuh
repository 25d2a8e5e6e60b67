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
//   edges       u16 index of their term                                      (2 B per edge)
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

// CTA width.  512 threads (128 registers, 16 warps on the SM) is the default: the passes are latency-bound and twice
// the warps hide more of it (cfg4 wave 36 -> 27 ms).  256 threads is the generic block kernel's widest CTA — the same
// factor -> thread mapping and reduction tree, hence bit-identical results — and stays selectable
// (rdisgpu_set_option "resident_threads") for the equality test.
constexpr int kResThreads = 512;
constexpr int kResThreadsExact = 256;

// Shared-memory carve-up, computed identically on the host (fits? how many bytes to ask for) and in the kernel.
struct ResLayout {
  int xs, ds, lb, ub, tval, tdt, texpo, tkonst, fcoef;  // double arrays (byte offsets)
  int toff, frow;                                        // int32
  int tlv, elt, vinc;                                    // uint16
  int tsine;                                             // uint8
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
  L.xs = take(8 * nv); L.ds = take(8 * nv); L.lb = take(8 * nv); L.ub = take(8 * nv);
  L.tval = take(8 * (nT + nFz)); L.tdt = take(8 * (nT + nFz));
  L.texpo = take(8 * nT); L.tkonst = take(8 * nT);
  L.fcoef = take(8 * nf);
  L.toff = take(4 * (nv + 1)); L.frow = take(4 * (nf + 1));
  L.tlv = take(2 * (nT + nFz)); L.elt = take(2 * nE); L.vinc = take(2 * nE);
  L.tsine = take(nT);
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
  uint16_t *tlv, *elt, *vinc;
  uint8_t* tsine;
  int nT;
  double* gscr;  // this problem's slice of the per-edge partial scratch (HBM/L2), indexed by local edge
};

// One line evaluation: f(p + alpha*xi) and, if want_slope, d/dalpha — the shared-memory form of objective_along_line.
__device__ __forceinline__ void resident_line_eval(const GraphView& G, const ResView& R, Block& grp, const int32_t* fids, int nf,
                                                   double alpha, bool want_slope, double& f, double& slope) {
  const int T = blockDim.x, tid = threadIdx.x;
  // ---- pass 1: terms, two per trip: every input of both is loaded before either result is stored, so the two
  // dependency chains (clamp, power, sincos) interleave ----
  for (int t = tid; t < R.nT; t += 2 * T) {
    const int u = t + T;
    const bool two = u < R.nT;
    const int ub_ = two ? u : t;
    const int lv0 = R.tlv[t], lv1 = R.tlv[ub_];
    const double raw0 = R.xs[lv0] + alpha * R.ds[lv0], raw1 = R.xs[lv1] + alpha * R.ds[lv1];
    const double xv0 = clamp_to_domain(raw0, make_double2(R.lb[lv0], R.ub[lv0]));
    const double xv1 = clamp_to_domain(raw1, make_double2(R.lb[lv1], R.ub[lv1]));
    const double ex0 = R.texpo[t], kk0 = R.tkonst[t], ex1 = R.texpo[ub_], kk1 = R.tkonst[ub_];
    const bool sn0 = R.tsine[t] != 0, sn1 = R.tsine[ub_] != 0;
    if (want_slope) {
      double tv0, dt0, tv1, dt1;
      nlpf_term_grad(xv0, kk0, ex0, sn0, tv0, dt0);
      nlpf_term_grad(xv1, kk1, ex1, sn1, tv1, dt1);
      R.tval[t] = tv0;
      R.tdt[t] = dt0;
      if (two) {
        R.tval[u] = tv1;
        R.tdt[u] = dt1;
      }
    } else {
      const double tv0 = nlpf_term_value(xv0, kk0, ex0, sn0), tv1 = nlpf_term_value(xv1, kk1, ex1, sn1);
      R.tval[t] = tv0;
      if (two) R.tval[u] = tv1;
    }
  }
  __syncthreads();
  // ---- pass 2: factors, thread k owns factors k, k+T, ... (objective_along_line's mapping) ----
  double fs = 0.0, ss = 0.0;
#pragma unroll 2
  for (int k = tid; k < nf; k += T) {
    const int r0 = R.frow[k] & 0x7fffffff;
    const int n = (R.frow[k + 1] & 0x7fffffff) - r0;
    const bool is_const = R.frow[k] < 0;  // Factor::eval of an assigned constant (src/Factor.cpp:110-119)
    const double c = R.fcoef[k];
    double prod = 1.0, s = 0.0;
    if (n <= NlpfOps::kMaxArityFast) {
      double t[NlpfOps::kMaxArityFast], dt[NlpfOps::kMaxArityFast], dir[NlpfOps::kMaxArityFast];
#pragma unroll
      for (int i = 0; i < NlpfOps::kMaxArityFast; ++i) {
        if (i < n) {
          const int lt = R.elt[r0 + i];
          t[i] = R.tval[lt];
          if (want_slope) {
            dt[i] = R.tdt[lt];
            dir[i] = (lt < R.nT) ? R.ds[R.tlv[lt]] : 0.0;
          }
          prod *= t[i];
        }
      }
      if (want_slope) {
#pragma unroll
        for (int i = 0; i < NlpfOps::kMaxArityFast; ++i) {
          if (i < n && dir[i] != 0.0) {
            double pe = 1.0;  // getDerivative: product in slot order, own slot replaced by its derivative (1.0 when plain)
#pragma unroll
            for (int j = 0; j < NlpfOps::kMaxArityFast; ++j) {
              if (j < n) pe *= (j == i) ? dt[j] : t[j];
            }
            s += (pe * c) * dir[i];
          }
        }
      }
    } else {
      for (int i = 0; i < n; ++i) prod *= R.tval[R.elt[r0 + i]];
      if (want_slope) {
        for (int i = 0; i < n; ++i) {
          const int lt = R.elt[r0 + i];
          const double diri = (lt < R.nT) ? R.ds[R.tlv[lt]] : 0.0;
          if (diri != 0.0) {
            double pe = 1.0;
            for (int j = 0; j < n; ++j) {
              const int lj = R.elt[r0 + j];
              pe *= (j == i) ? R.tdt[lj] : R.tval[lj];
            }
            s += pe * c * diri;
          }
        }
      }
    }
    double fv = prod * c;
    if (is_const) fv = G.fconst_val[fids[k]];
    fs += fv;
    ss += s;
  }
  grp.sum2(fs, ss);  // one __syncthreads inside: also orders this evaluation's reads of tval before the next pass 1
  f = fs;
  slope = ss;
}

// Value and full gradient at p (Factor::computeGradient of every factor + computeGradientOfSum restricted to the
// problem's variables): returns this thread's partial of the objective (caller reduces it when it needs f);
// thread j-strided gradient entries are handed to `sink(j, dF/dx_j)`.  R.toff holds the incidence offsets here.
template <class Sink>
__device__ __forceinline__ double resident_gradient(const GraphView& G, const ResView& R, const int32_t* fids, int nv, int nf,
                                                    Sink sink) {
  const int T = blockDim.x, tid = threadIdx.x;
  for (int t = tid; t < R.nT; t += 2 * T) {  // two terms per trip (see resident_line_eval)
    const int u = t + T;
    const bool two = u < R.nT;
    const int ub_ = two ? u : t;
    const int lv0 = R.tlv[t], lv1 = R.tlv[ub_];
    const double xv0 = clamp_to_domain(R.xs[lv0], make_double2(R.lb[lv0], R.ub[lv0]));  // load_var<false>
    const double xv1 = clamp_to_domain(R.xs[lv1], make_double2(R.lb[lv1], R.ub[lv1]));
    const double ex0 = R.texpo[t], kk0 = R.tkonst[t], ex1 = R.texpo[ub_], kk1 = R.tkonst[ub_];
    const bool sn0 = R.tsine[t] != 0, sn1 = R.tsine[ub_] != 0;
    double tv0, dt0, tv1, dt1;
    nlpf_term_grad(xv0, kk0, ex0, sn0, tv0, dt0);
    nlpf_term_grad(xv1, kk1, ex1, sn1, tv1, dt1);
    R.tval[t] = tv0;
    R.tdt[t] = dt0;
    if (two) {
      R.tval[u] = tv1;
      R.tdt[u] = dt1;
    }
  }
  __syncthreads();
  double fs = 0.0;
  for (int k = tid; k < nf; k += T) {  // NlpfOps::gradient, expression for expression
    const int r0 = R.frow[k] & 0x7fffffff;
    const int n = (R.frow[k + 1] & 0x7fffffff) - r0;
    const bool is_const = R.frow[k] < 0;
    const double c = R.fcoef[k];
    double prod = 1.0;
    if (n <= NlpfOps::kMaxArityFast) {
      double t[NlpfOps::kMaxArityFast], dt[NlpfOps::kMaxArityFast];
#pragma unroll
      for (int i = 0; i < NlpfOps::kMaxArityFast; ++i) {
        if (i < n) {
          const int lt = R.elt[r0 + i];
          t[i] = R.tval[lt];
          dt[i] = R.tdt[lt];
          prod *= t[i];
        }
      }
#pragma unroll
      for (int i = 0; i < NlpfOps::kMaxArityFast; ++i) {
        if (i < n) {
          double pe = 1.0;
#pragma unroll
          for (int j = 0; j < NlpfOps::kMaxArityFast; ++j) {
            if (j < n) pe *= (j == i) ? dt[j] : t[j];
          }
          R.gscr[r0 + i] = pe * c;
        }
      }
    } else {
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
template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1)
    solve_nlpf_resident_kernel(GraphView G, BatchView B, const int32_t* order, int count, int maxiters, double ftol) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  __shared__ double scratch[260];
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

  const ResLayout L = res_layout(nv, nf, P.nE, P.nT, P.nFz);
  ResView R;
  R.xs = reinterpret_cast<double*>(res_smem + L.xs); R.ds = reinterpret_cast<double*>(res_smem + L.ds);
  R.lb = reinterpret_cast<double*>(res_smem + L.lb); R.ub = reinterpret_cast<double*>(res_smem + L.ub);
  R.tval = reinterpret_cast<double*>(res_smem + L.tval); R.tdt = reinterpret_cast<double*>(res_smem + L.tdt);
  R.texpo = reinterpret_cast<double*>(res_smem + L.texpo); R.tkonst = reinterpret_cast<double*>(res_smem + L.tkonst);
  R.fcoef = reinterpret_cast<double*>(res_smem + L.fcoef);
  R.toff = reinterpret_cast<int32_t*>(res_smem + L.toff); R.frow = reinterpret_cast<int32_t*>(res_smem + L.frow);
  R.tlv = reinterpret_cast<uint16_t*>(res_smem + L.tlv); R.elt = reinterpret_cast<uint16_t*>(res_smem + L.elt);
  R.vinc = reinterpret_cast<uint16_t*>(res_smem + L.vinc);
  R.gscr = B.gscr + P.goff;
  R.tsine = res_smem + L.tsine;
  R.nT = P.nT;

  // ---- claim (CGD.cpp:33: quickAssignVals of the start point) + flatten ----
  if (tid == 0) nfz_counter = 0;
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    const double xv = (B.x0 != nullptr) ? B.x0[P.var_off + j] : G.xbd[vid].x;
    G.xsave[vid] = xv;
    G.xbd[vid] = make_double2(xv, 0.0);
    G.vloc[vid] = j;
    const double2 dm = __ldg(&G.dom[vid]);
    R.xs[j] = xv; R.ds[j] = 0.0; R.lb[j] = dm.x; R.ub[j] = dm.y;
    R.toff[j] = __ldg(&G.tvrow[vid + 1]) - __ldg(&G.tvrow[vid]);
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
  // term descriptors of the component's own variables
  for (int j = tid; j < nv; j += T) {
    const int32_t vid = vids[j];
    const int32_t g0 = __ldg(&G.tvrow[vid]);
    const int t0 = R.toff[j], n = R.toff[j + 1] - t0;
    for (int u = 0; u < n; ++u) {
      R.tlv[t0 + u] = (uint16_t)j;
      R.texpo[t0 + u] = __ldg(&G.t_expo[g0 + u]);
      R.tkonst[t0 + u] = __ldg(&G.t_konst[g0 + u]);
      R.tsine[t0 + u] = __ldg(&G.t_sine[g0 + u]);
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
        R.tlv[slot] = 0;
        R.elt[r0 + i] = (uint16_t)slot;
      } else {
        const int lv = G.vloc[vid];
        R.elt[r0 + i] = (uint16_t)(R.toff[lv] + (__ldg(&G.eterm[e]) - __ldg(&G.tvrow[vid])));
      }
    }
  }
  __syncthreads();
  // variable-major incidence of the component (local edge ids, ascending factor id = the order of the global
  // incidence list): toff is free now and becomes its offsets
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
      R.vinc[pos++] = (uint16_t)(R.frow[G.floc[f]] + (e - __ldg(&G.rowptr[f])));
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

  CgdMachine m;
  m.start(maxiters, ftol);
  double f_init = 0.0;

  while (!m.done()) {
    if (m.req == REQ_INIT_GRAD) {
      double fs = resident_gradient(G, R, fids, nv, nf, [&](int j, double gr) {
        const int32_t vid = vids[j];
        const double gneg = -gr;
        G.gvec[vid] = gneg;
        G.hvec[vid] = gneg;
        R.ds[j] = gneg;
      });
      double zero = 0.0;
      grp.sum2(fs, zero);
      grp.sync();
      f_init = fs;
      m.on_init(fs);
    } else {
      double f, sl;
      resident_line_eval(G, R, grp, fids, nf, m.alpha, m.req == REQ_VALUE_SLOPE, f, sl);
      m.on_eval(f, sl);
    }

    while (m.req == REQ_MOVE) {
      const double step = m.alpha;
      for (int j = tid; j < nv; j += T) {  // minimize_nrc.h:508-511
        double2 xb = make_double2(R.xs[j], R.ds[j]);
        xb.y *= step;
        xb.x += xb.y;
        R.xs[j] = xb.x; R.ds[j] = xb.y;
      }
      grp.sync();
      m.on_moved();
      if (m.req != REQ_GRADIENT) break;

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
      m.on_gradient(tnum, gg, dgg);
      if (m.req != REQ_DIRECTION) break;
      const double gam = m.gam;
      for (int j = tid; j < nv; j += T) {  // :681-685
        const int32_t vid = vids[j];
        const double gj = -R.ds[j];
        const double hj = gj + gam * G.hvec[vid];
        G.gvec[vid] = gj;
        G.hvec[vid] = hj;
        R.ds[j] = hj;
      }
      grp.sync();
      m.on_directed();
    }
  }

  // ---- commit (CGD.cpp:61-89) ----
  double fret = m.fret;
  bool restore = (fret > f_init);
  if (m.status == ST_NONFINITE || m.status == ST_BRACKET_CAP) restore = true;
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
