/*
 * rdis_gpu.h — C-ABI of the B200-native subspace-solve / factor-sweep path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no C++ or
 * torch types.  Each entry point names the reference interface it replaces
 * (paths relative to the afriesen/rdis tree).  The C++ adapter that keeps the
 * reference's plugin surface (class CudaSubspaceOptimizer : SubspaceOptimizer) lives in
 * rdis_b200/host/ and calls only the functions declared here; INTEGRATION.md shows the
 * reference-side binding.
 *
 * All arithmetic is IEEE fp64 (reference: typedef double Numeric, src/common.h:25).
 * There is no CPU fallback: every call fails with RDISGPU_ERR_CUDA if no sm_100 device
 * is usable.
 *
 * Threading: a context is single-threaded, like the reference (src/IntrusivePtrPool.h:53).
 */
#ifndef RDIS_GPU_H_
#define RDIS_GPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RDISGPU_API __attribute__((visibility("default")))
#else
#define RDISGPU_API
#endif

typedef struct rdisgpu_ctx rdisgpu_ctx;     /* one OptimizableFunction resident on one GPU */
typedef struct rdisgpu_batch rdisgpu_batch; /* a packed set of sibling subspace problems    */

enum {
  RDISGPU_OK = 0,
  RDISGPU_ERR_CUDA = 1,      /* CUDA runtime error; text in rdisgpu_last_error            */
  RDISGPU_ERR_ARG = 2,       /* bad argument (null pointer, id out of range, wrong state)  */
  RDISGPU_ERR_STATE = 3,     /* call order violated (e.g. solve before finalize)           */
  RDISGPU_ERR_OVERLAP = 4    /* problems of one batch share a variable or a factor         */
};

/* Per-problem termination status (rdisgpu_result.status). */
enum {
  RDISGPU_DONE_FTOL = 0,       /* Frprmn ftol return,   external/include/minimize_nrc.h:648-652 */
  RDISGPU_DONE_GTOL = 1,       /* zero-gradient return, minimize_nrc.h:663-666                  */
  RDISGPU_DONE_GG_ZERO = 2,    /* gg == 0 return,       minimize_nrc.h:676-679                  */
  RDISGPU_DONE_MAXITERS = 3,   /* "Too many iterations in frprmn", minimize_nrc.h:690 (normal)  */
  RDISGPU_DONE_DBRENT_ITMAX = 4, /* "Too many iterations in routine dbrent", minimize_nrc.h:403 */
  RDISGPU_DONE_EMPTY = 5,      /* no factors: returns 0, delta 0, x untouched (CGDSubspaceOptimizer.cpp:26-29) */
  RDISGPU_DONE_NONFINITE = 6,  /* a non-finite line-search abscissa was produced (the reference asserts, CGD.cpp:172): p and
                                  fret of the last completed line search are kept, as after the reference's own throws */
  RDISGPU_DONE_BRACKET_CAP = 7 /* bracket expansion exceeded the device safety cap of 2000 rounds (the reference's loop is
                                  unbounded, minimize_nrc.h:101); same commit rule                                  */
};

/* One SubspaceOptimizer::optimize call (src/SubspaceOptimizer.h:37-39):
 * minimise sum_{j in fid} f_j over the variables vid, everything else held fixed. */
typedef struct rdisgpu_problem {
  int64_t nv;          /* |vars|                                                              */
  const int32_t* vid;  /* variable ids, order = the reference's `vars` vector                 */
  int64_t nf;          /* |factors|                                                           */
  const int64_t* fid;  /* factor ids, order = the reference's `factors` vector (sorted by id
                          when built by RDISOptimizer.cpp:1049-1059)                          */
  const double* x0;    /* nv start values (`xval` on entry); NULL = current device values     */
} rdisgpu_problem;

typedef struct rdisgpu_result {
  double* x;        /* out, nv values: domain-clamped final values (`xval` on exit); may be NULL */
  double f_init;    /* f at the (clamped) start point                                            */
  double f_end;     /* return value of optimize(): f at the returned point                       */
  int32_t iters;    /* Frprmn::iter at exit                                                      */
  int32_t status;   /* RDISGPU_DONE_*                                                            */
  int64_t n_feval;  /* objective evaluations the device performed (each = nf factor evals)       */
  int64_t n_geval;  /* of those, evaluations that also produced derivatives                      */
} rdisgpu_result;

/* ---- lifetime ------------------------------------------------------------------------- */
RDISGPU_API int rdisgpu_create(rdisgpu_ctx** out, int device);
RDISGPU_API void rdisgpu_destroy(rdisgpu_ctx* ctx);
RDISGPU_API const char* rdisgpu_last_error(const rdisgpu_ctx* ctx); /* ctx may be NULL: last create error */
/* All work of this context is enqueued on `cuda_stream` (a cudaStream_t; NULL = default). */
RDISGPU_API int rdisgpu_set_stream(rdisgpu_ctx* ctx, void* cuda_stream);
RDISGPU_API int rdisgpu_synchronize(rdisgpu_ctx* ctx);
/* Tuning / test switches.  "generic_only" != 0: batches created afterwards bypass the bundle-adjustment
 * block kernels and run every problem through the generic tile / CTA / grid kernels.  "resident_threads" = 256 | 1024:
 * CTA width of the shared-memory resident NonlinearProductFactor component kernel (1024 = default; 256 reproduces the
 * generic CTA kernel's reduction order and therefore its results to the bit — used by the equality test).
 * "strict" != 0 (needs a finalized context): every subspace solve runs through the strict kernel, which reproduces
 * CGDSubspaceOptimizer::optimize (src/optimizers/CGDSubspaceOptimizer.cpp:19-98) operation for operation — the reference's
 * summation orders (src/OptimizableFunction.cpp:108-132, src/State.h:157-194, minimize_nrc.h:440-446,668-673), true
 * divisions, the per-factor value cache with Variable::assign's 1e-12 change filter (src/Variable.cpp:66-88,
 * src/Factor.cpp:110-119; the cache persists across calls and rdisgpu_set_x applies the filter) and the reference's
 * evaluation sequence — and is bit-identical to the CPU restatement built with the device's sin / cos.
 * "camera_cluster" = 0..8: pins the thread-block-cluster width of the camera-block kernel (0 = choose; results do not
 * depend on it).  "adaptive_order" = 0 | 1 (default 1): after a batch solve the
 * block kernels' launch order is re-sorted on the device by the evaluation counts just observed, longest first, for the next
 * visit of the same batch (results do not depend on it).  "point_tiles_per_warp" = 0..32: at most this many point blocks share a warp (0 = choose from the batch size;
 * results do not depend on it). */
RDISGPU_API int rdisgpu_set_option(rdisgpu_ctx* ctx, const char* name, int64_t value);

/* ---- function definition (replaces the OptimizableFunction / Factor object graph) ------ */
/* Variables 0..V-1 with their single-interval domains (VariableDomain, src/VariableDomain.cpp:130-163). */
RDISGPU_API int rdisgpu_set_vars(rdisgpu_ctx* ctx, int64_t V, const double* lb, const double* ub);
/* NonlinearProductFactor list in CSR form by factor id (src/NonlinearProductFactor.h:27-55,100-113):
 * factor j = coeff[j] * prod_{e in [rowptr[j],rowptr[j+1])} t_e,
 * t_e = [sin]( (x[vid[e]] - konst[e]) ^ expo[e] ).                                              */
RDISGPU_API int rdisgpu_add_nlpf(rdisgpu_ctx* ctx, int64_t F, const int64_t* rowptr, const int32_t* vid,
                                 const double* expo, const double* konst, const uint8_t* use_sine,
                                 const double* coeff);
/* BundleAdjustmentFactor list (src/bundleadjust/BundleAdjustmentFactor.h:129-136): observation j of
 * point pt[j] by camera cam[j] at pixel obs_xy[2j..2j+1].  Variable ids are implied by
 * BundleAdjustmentFunction.h:88-96: camera c parameter p -> 9c+p, point i coordinate d -> 9*ncams+3i+d. */
RDISGPU_API int rdisgpu_add_ba(rdisgpu_ctx* ctx, int64_t F, const int32_t* cam, const int32_t* pt,
                               const double* obs_xy, int32_t ncams, int32_t npts);
/* Builds the variable-major incidence and uploads everything (OptimizableFunction::init). */
RDISGPU_API int rdisgpu_finalize(rdisgpu_ctx* ctx);

/* ---- state ---------------------------------------------------------------------------- */
/* Variable::assign for n variables (vid NULL = 0..n-1), src/Variable.cpp:66-88.  x (and vid) may be
 * device pointers: then nothing is copied and the call is asynchronous on the context's stream. */
RDISGPU_API int rdisgpu_set_x(rdisgpu_ctx* ctx, int64_t n, const int32_t* vid, const double* x);
RDISGPU_API int rdisgpu_get_x(rdisgpu_ctx* ctx, int64_t n, const int32_t* vid, double* x);
/* Factors the tree search simplified to a constant (Factor::isAssigned, src/Factor.cpp:110-119):
 * their value is val[i] while on[i] != 0; their gradient is still the full analytic one
 * (src/NonlinearProductFactor.cpp:57-117, BundleAdjustmentFactor.cpp:338-348).                   */
RDISGPU_API int rdisgpu_set_factor_const(rdisgpu_ctx* ctx, int64_t n, const int64_t* fid, const double* val,
                                         const uint8_t* on);

/* ---- sweeps --------------------------------------------------------------------------- */
/* OptimizableFunction::evalFactors (src/OptimizableFunction.cpp:95-135) over a factor-id list
 * (fid NULL = all F factors).  *sum = sum_j f_j; per_factor (nullable) receives nf values. */
RDISGPU_API int rdisgpu_eval(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, double* sum, double* per_factor);
/* OptimizableFunction::computeGradient(facs, pg) + the scatter of SubfunctionFD::df
 * (src/OptimizableFunction.cpp:234-262, CGDSubspaceOptimizer.cpp:135-157): g[i] = d/dx_{vid[i]} sum_j f_j. */
RDISGPU_API int rdisgpu_grad(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, int64_t nv, const int32_t* vid,
                             double* g);
/* The two sweeps with every pointer in DEVICE memory (int32 factor / variable ids, nullable = all):
 * nothing is copied and the host does not wait — the call only enqueues kernels on the context's
 * stream.  This is the form the roofline is measured on and the one a device-resident caller uses. */
RDISGPU_API int rdisgpu_eval_device(rdisgpu_ctx* ctx, int64_t nf, const int32_t* fid_dev, double* sum_dev,
                                    double* per_factor_dev);
RDISGPU_API int rdisgpu_grad_device(rdisgpu_ctx* ctx, int64_t nf, const int32_t* fid_dev, int64_t nv,
                                    const int32_t* vid_dev, double* g_dev);
/* Factor::computeGradient for one factor list: rows[k*arity_max + s] = d f_{fid[k]} / d slot s. */
RDISGPU_API int rdisgpu_factor_grad(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, int32_t arity_max,
                                    double* rows);

/* ---- subspace solves ------------------------------------------------------------------- */
/* CGDSubspaceOptimizer::optimize (src/optimizers/CGDSubspaceOptimizer.cpp:19-98) for nprobs
 * problems in one call.  nprobs > 1 is the sibling-component batch (RDISOptimizer.cpp:184-211,
 * 291-314): the problems must not share variables or factors, and no factor of one problem may
 * touch a variable of another.  Host buffers in, host buffers out; final values are also
 * committed to the device state (Variable::assign post-condition, CGD.cpp:61,84-86).            */
RDISGPU_API int rdisgpu_solve_cgd(rdisgpu_ctx* ctx, const rdisgpu_problem* probs, int64_t nprobs, int maxiters,
                                  double ftol, rdisgpu_result* out);

/* The same with the batch described by packed (CSR) lists — what a caller that holds a whole wave of
 * sibling components naturally has: problem p owns vids[var_off[p]..var_off[p+1]) and
 * fids[fac_off[p]..fac_off[p+1]).  x0 (nullable = current device values) and x_out (nullable) are
 * concatenated in problem order; the per-problem outputs (each nullable) have nprobs entries.
 * All buffers are host memory; index lists, start values and results cross PCIe once each.        */
RDISGPU_API int rdisgpu_solve_cgd_csr(rdisgpu_ctx* ctx, int64_t nprobs, const int64_t* var_off, const int32_t* vids,
                                      const int64_t* fac_off, const int64_t* fids, const double* x0, int maxiters,
                                      double ftol, double* x_out, double* f_init, double* f_end, int32_t* iters,
                                      int32_t* status, int64_t* n_feval, int64_t* n_geval);

/* LMSubspaceOptimizer::optimize (src/optimizers/LMSubspaceOptimizer.cpp:29-171) for nprobs sibling problems,
 * packed like rdisgpu_solve_cgd_csr: residual sqrt(2 f_j) per factor, dense Jacobian, levmar's dlevmar_der
 * (x = NULL) with opts4 = {tau, eps1, eps2, eps3} (NULL = the reference's {1e-3, 1e-15, 1e-15, 3e-8}, :83-86),
 * no domain clamping (:281-297), no revert.  PARITY UNPINNED: levmar is not vendored with the reference and
 * no reference test runs this optimizer; the arithmetic follows oracle/lm_oracle.hpp.  Components of more
 * than 32 variables are rejected (RDISGPU_ERR_ARG) in this version.  stop[p] is levmar's termination code:
 * 1 small gradient, 2 small step, 3 itmax, 4 singular, 5 no further reduction, 6 small ||e||^2, 7 non-finite;
 * 0 for a problem without factors (returns 0, x untouched).  n_feval / n_jeval count func / jacf calls. */
RDISGPU_API int rdisgpu_solve_lm_csr(rdisgpu_ctx* ctx, int64_t nprobs, const int64_t* var_off, const int32_t* vids,
                                     const int64_t* fac_off, const int64_t* fids, const double* x0, int maxiters,
                                     const double* opts4, double* x_out, double* f_init, double* f_end, int32_t* iters,
                                     int32_t* stop, int64_t* n_feval, int64_t* n_jeval);

/* The same in three steps, for callers that revisit one component structure many times
 * (alternating minimisation, RDISOptimizer.cpp:1148-1181): index lists stay resident in HBM. */
RDISGPU_API int rdisgpu_batch_create(rdisgpu_ctx* ctx, const rdisgpu_problem* probs, int64_t nprobs,
                                     rdisgpu_batch** out);
RDISGPU_API int rdisgpu_batch_create_csr(rdisgpu_ctx* ctx, int64_t nprobs, const int64_t* var_off, const int32_t* vids,
                                         const int64_t* fac_off, const int64_t* fids, rdisgpu_batch** out);
/* x0 (nullable): concatenated start values in problem order, host (pinned for a truly asynchronous
 * copy) or device memory; NULL = current device values.  Asynchronous on the context's stream. */
RDISGPU_API int rdisgpu_batch_solve_cgd(rdisgpu_batch* b, const double* x0_host, int maxiters, double ftol);
/* Waits, then copies results out (out[i].x may be NULL).  sum_f_end (nullable) = sum of f_end. */
RDISGPU_API int rdisgpu_batch_fetch(rdisgpu_batch* b, rdisgpu_result* out, double* sum_f_end);
/* The same with packed outputs, laid out like rdisgpu_solve_cgd_csr's (every pointer nullable): what the C++ adapter's
 * batch cache uses when the tree search revisits a sibling set (alternating minimisation, RDISOptimizer.cpp:1148-1181). */
RDISGPU_API int rdisgpu_batch_fetch_csr(rdisgpu_batch* b, double* x_out, double* f_init, double* f_end, int32_t* iters, int32_t* status,
                                        int64_t* n_feval, int64_t* n_geval);
/* *sum_dev += sum_i f_end[i], computed on the device (sum_dev is device memory): the per-GPU partial of
 * the global objective, ready for an NCCL all-reduce on the same stream.  Asynchronous. */
RDISGPU_API int rdisgpu_batch_objective_device(rdisgpu_batch* b, double* sum_dev);
RDISGPU_API void rdisgpu_batch_destroy(rdisgpu_batch* b);
/* How the batch was mapped: out = {nprobs, point-block warps, camera blocks, cluster size, threads per
 * CTA of the camera kernel, problems on the generic kernels, max observations of a camera block,
 * launches of the last solve}. */
RDISGPU_API int rdisgpu_batch_info(const rdisgpu_batch* b, int32_t out[8]);
/* NonlinearProductFactor components solved by the shared-memory resident kernel (a component of more than 32
 * factors whose flattened form — variables, distinct terms, edges, factors — fits one CTA's shared memory):
 * out = {number of such problems in the batch, dynamic shared memory bytes of the launch}.  They are not counted
 * in rdisgpu_batch_info's "problems on the generic kernels". */
RDISGPU_API int rdisgpu_batch_resident_info(const rdisgpu_batch* b, int32_t out[2]);
/* Kernel launches the last rdisgpu_batch_solve_cgd enqueued (bench.py's gpu_launches). */
RDISGPU_API int rdisgpu_batch_last_launches(const rdisgpu_batch* b);

/* ---- component membership ------------------------------------------------- */
/* Sibling components as Component::createChildren produces them (src/Component.cpp:508-549; connectivity rule
 * src/ConnectivityGraph.cpp:211-285): connected components of the bipartite variable / factor graph with an edge
 * (v, f) iff variable v is unassigned (assigned[v] == 0) and factor f is not an assigned constant
 * (rdisgpu_set_factor_const).  var_label[v] = smallest variable id of v's component, -1 for assigned variables;
 * fac_label[f] (nullable) = the label of the component f belongs to, -1 if it has none.  Exact integer bookkeeping
 * (min-label propagation with pointer jumping on the device); host buffers in and out.  n_components / n_rounds
 * (nullable) receive the number of components and of propagation rounds. */
RDISGPU_API int rdisgpu_components(rdisgpu_ctx* ctx, const uint8_t* assigned, int32_t* var_label, int32_t* fac_label,
                                   int32_t* n_components, int32_t* n_rounds);

/* ---- interval bounds for branch & bound ------------------------------------ */
/* Factor::computeBounds (src/Factor.cpp:122-139) for every listed factor and their interval sum as
 * OptimizableFunction::computeBounds folds it (src/OptimizableFunction.cpp:186-216; the caller applies that function's
 * assigned-key filter by choosing the list).  A variable with assigned[v] != 0 enters as the point [x, x] (device
 * state, rdisgpu_set_x), any other as its domain hull [lb, ub]; a factor whose variables are all assigned, or that is
 * an assigned constant, contributes its point value.  NonlinearProductFactor::computeFactorBounds
 * (src/NonlinearProductFactor.cpp:120-145) and BundleAdjustmentFactor::evalFactor(IntervalVec)
 * (src/bundleadjust/BundleAdjustmentFactor.cpp:104-157,186-232) over the reference's no-rounding / no-checking
 * Boost.Interval policy (src/common.h:43-60).  fid NULL = all factors; lower / upper (nf doubles, nullable) receive
 * the per-factor bounds, sum (nullable) = {lower, upper} of the list, accumulated in list order. */
RDISGPU_API int rdisgpu_bounds(rdisgpu_ctx* ctx, const uint8_t* assigned, int64_t nf, const int64_t* fid, double* lower,
                               double* upper, double sum[2]);

/* The residual + Jacobian-rows sweep over ALL factors of a bundle-adjustment graph, device pointers, asynchronous on the
 * context's stream: per_factor_dev[F] (nullable) the factor values, rows_dev[12 F] the 12 partial derivatives of every
 * factor in slot order (rot xyz, trans xyz, focal, k1, k2, point xyz), *sum_dev (nullable) the total.  This is what
 * LMSSOpt::evalFunc / evalJacf produce one factor at a time (src/optimizers/LMSubspaceOptimizer.cpp:176-278, through
 * BundleAdjustmentFactor::computeGradient, src/bundleadjust/BundleAdjustmentFactor.cpp:351-554); 128 B/factor + 8 B/variable. */
RDISGPU_API int rdisgpu_factor_rows_device(rdisgpu_ctx* ctx, double* sum_dev, double* per_factor_dev, double* rows_dev);

/* Interval bounds of MANY factor lists in one call: list l = fid[list_off[l] .. list_off[l+1]), sums[2l] / sums[2l+1] = lower /
 * upper bound of the list's sum, folded on the device in list order.  This is Component::computeBounds for every child of a
 * decomposition at once (src/Component.cpp:240-241,592-599 -> src/OptimizableFunction.cpp:181-216), what the branch & bound
 * bookkeeping of the sibling loop reads (src/RDISOptimizer.cpp:291-314,894-934). */
RDISGPU_API int rdisgpu_bounds_lists(rdisgpu_ctx* ctx, const uint8_t* assigned, int64_t nlists, const int64_t* list_off, const int64_t* fid,
                                     double* sums);

/* ---- introspection (tests / bench) ------------------------------------------------------ */
RDISGPU_API int64_t rdisgpu_num_vars(const rdisgpu_ctx* ctx);
RDISGPU_API int64_t rdisgpu_num_factors(const rdisgpu_ctx* ctx);
/* Raw device pointer + element count of the committed state, for zero-copy interop
 * (torch.from_blob / __cuda_array_interface__): V pairs {value, direction-slot}. */
RDISGPU_API int rdisgpu_device_state(rdisgpu_ctx* ctx, void** dptr, int64_t* n_pairs);
/* Total kernel launches issued by this context so far. */
RDISGPU_API int64_t rdisgpu_launch_count(const rdisgpu_ctx* ctx);
RDISGPU_API const char* rdisgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RDIS_GPU_H_ */
