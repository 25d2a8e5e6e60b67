// rdis_gpu.cu — the C-ABI of include/rdis_gpu.h: context / HBM residency management, the
// sweep kernels (evalFactors, computeGradient) and the launch logic of the batched subspace
// solves (solve_kernels.cuh).  CUDA runtime only; no torch, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rdis_gpu.h"
#include "ba_block_kernels.cuh"
#include "solve_kernels.cuh"
#include "sweep_kernels.cuh"
#include "nlpf_tile_sweep.cuh"
#include "lm_kernels.cuh"
#include "ba_sweep.cuh"
#include "components.cuh"
#include "nlpf_resident.cuh"
#include "bounds_kernels.cuh"
#include "strict_kernels.cuh"
#include "lm_dense.cuh"

using namespace rdisgpu;

namespace {

std::string g_create_error;

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t ensure(size_t count) {  // grow-only
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t n = 0;
  PinnedBuf() {}
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    if (count == 0) count = 1;
    cudaError_t e = cudaMallocHost((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

}  // namespace

struct rdisgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 148;
  bool finalized = false;
  int kind = KIND_NONE;
  int64_t V = 0, F = 0, E = 0;
  int64_t launches = 0;

  // host staging of the definition (dropped at finalize)
  std::vector<double> h_lb, h_ub;
  std::vector<int64_t> h_rowptr;
  std::vector<int32_t> h_evid;
  std::vector<double> h_expo, h_konst, h_coeff;
  std::vector<uint8_t> h_sine;
  std::vector<int32_t> h_cam, h_pt;
  std::vector<double> h_obs;
  int32_t ncams = 0, npts = 0;

  // device residency
  DevBuf<double2> xbd, dom;
  DevBuf<int32_t> rowptr, evid, vrow, vedge, efac, cam, pt, crow, cfac, prow, pfac, fstamp;
  DevBuf<CameraRow> cam_table;  // BA: per-camera part of the forward model for the all-factor sweep (ba_sweep.cuh)
  DevBuf<TileDesc> tiles;  // NLPF: factor tiles of the streaming sweep (nlpf_tile_sweep.cuh)
  int ntiles = 0, tile_grid[2] = {0, 0};  // persistent grid size of the eval / grad instantiation
  DevBuf<double> expo, konst, coeff, gedge, gvec, hvec, xsave, fconst_val, xval;
  DevBuf<uint8_t> sine, fconst_on;
  DevBuf<double2> obs;
  bool has_fconst = false;
  std::vector<int32_t> vmark, fmark;  // sibling check: epoch stamps (no per-call clearing)
  int32_t mark_epoch = 0;
  rdisgpu_batch* scratch_batch = nullptr;  // reused by the one-shot solve entry points
  int live_batches = 0;                    // rdisgpu_batch_create'd and not yet destroyed: rdisgpu_destroy refuses while > 0
  bool generic_only = false;  // rdisgpu_set_option("generic_only"): bypass the BA block kernels (tests)
  bool adaptive_order = true; // rdisgpu_set_option("adaptive_order"): re-sort clusters / warp tasks by the previous visit's evaluation counts
  int pt_warps_per_sm = 0;    // resident warps of solve_ba_points_kernel per SM (occupancy query, once)
  int pt_tiles_cap = 0;       // rdisgpu_set_option("point_tiles_per_warp"): at most this many point blocks share a warp (0 = choose)
  int cam_cluster_opt = 0;    // rdisgpu_set_option("camera_cluster"): pin the cluster width of the camera-block kernel (0 = choose)
  bool strict = false;        // rdisgpu_set_option("strict"): every solve through strict_kernels.cuh (bit-exact parity instrument)
  DevBuf<double> fcache;      // strict mode: the reference's per-factor value cache ...
  DevBuf<uint8_t> fdirty, vchg;  // ... its dirty flags, and Variable::assign's change flags
  GraphView gv;

  // scratch for the sweep / state calls
  DevBuf<int32_t> s_i32a, s_i32b;
  DevBuf<double> s_f64a, s_f64b, s_partials;
  DevBuf<unsigned int> s_counter;
  // Levenberg-Marquardt path (lm_kernels.cuh)
  DevBuf<int32_t> lm_vloc;
  DevBuf<double> lm_scratch;
  DevBuf<int64_t> lm_off;
  bool lm_vloc_ready = false;
  // Levenberg-Marquardt for one component of any size (lm_dense.cuh): dense normal equations + their factor, the
  // block-sparse Jacobian, the component's incidence, vectors and scalars
  DevBuf<double> lmd_A, lmd_L, lmd_jval, lmd_e, lmd_hx, lmd_vec, lmd_scal;
  DevBuf<int32_t> lmd_idx;
  DevBuf<int> lmd_flag;
  bool ba_smem_optin = false;
  // resident NonlinearProductFactor components (nlpf_resident.cuh): term table + host copies for classification
  DevBuf<int32_t> tvrow, eterm, vloc, floc;
  DevBuf<double> t_expo, t_konst;
  DevBuf<uint8_t> t_sine;
  std::vector<int32_t> h_rp32, h_evid32, h_tvrow, vowner;
  int res_threads = kResThreads;  // rdisgpu_set_option("resident_threads", 256 | 1024)
  int res_smem_cap = -1;  // dynamic shared memory a resident CTA may ask for (queried at the first NLPF batch)
  DevBuf<uint8_t> cc_assigned;
  DevBuf<int32_t> cc_vlabel, cc_flabel, cc_flag;
  DevBuf<double> grid_partials;
  PinnedBuf<char> pin;

  int fail_cuda(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return RDISGPU_ERR_CUDA;
  }
  int fail(int code, const char* what) {
    err = what;
    return code;
  }
};

#define CK(call)                                            \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return ctx->fail_cuda(e_, #call); \
  } while (0)

struct rdisgpu_batch {
  rdisgpu_ctx* ctx = nullptr;
  int64_t nprobs = 0;
  int64_t total_nv = 0, total_nf = 0;
  std::vector<ProblemDesc> h_probs;
  // All index lists of the batch live in ONE device allocation, staged through one pinned buffer and
  // uploaded with one copy: [ProblemDesc probs | vids | fids | order | pt_order | cam_order | pt_tasks]
  DevBuf<char> blob;
  PinnedBuf<char> h_blob;
  ProblemDesc* d_probs = nullptr;
  int32_t *d_vids = nullptr, *d_fids = nullptr, *d_order = nullptr, *d_pt_order = nullptr, *d_cam_order = nullptr;
  PointWarpTask *d_pt_tasks = nullptr, *d_pt_tasks_alt = nullptr;  // current launch order / where the next one is written
  DevBuf<double> x0, xout;
  DevBuf<ResultRec> res;
  // size classes of the generic kernels
  std::vector<int32_t> h_order;      // problem indices grouped by class
  struct Class { int kind; int param; int64_t off; int64_t count; };  // kind 0 = tile(G), 1 = block(threads), 2 = grid, 3 / 4 = resident NLPF (wide / small CTAs)
  std::vector<Class> classes;
  PinnedBuf<ResultRec> h_res;
  PinnedBuf<double> h_x;  // staging for x0 upload and xout download
  // bundle-adjustment block fast paths (ba_block_kernels.cuh)
  int n_pt_warps = 0, n_cam = 0, cam_nf_max = 0;
  int cam_C = 0, cam_T = 0;  // chosen at the first solve (needs the occupancy query)
  DevBuf<uint16_t> res_gvinc;  // ... and their variable-major incidence lists
  DevBuf<double> res_gscr;  // resident NLPF class: per-edge partials, one slice per problem
  DevBuf<double> strict_dscr;  // strict mode: Df1dim's derivative vector, one slice per problem
  int res_small_smem = 0;
  int res_smem = 0;  // dynamic shared memory of the resident NLPF class (largest layout in the batch)
  int last_launches = 0;
  bool solved = false;
};

// The two ways a caller can describe a batch: an array of rdisgpu_problem, or packed CSR lists.
struct ProblemsView {
  int64_t n = 0;
  const rdisgpu_problem* arr = nullptr;
  const int64_t* var_off = nullptr;
  const int32_t* vids = nullptr;
  const int64_t* fac_off = nullptr;
  const int64_t* fids = nullptr;
  int64_t nv(int64_t p) const { return arr ? arr[p].nv : var_off[p + 1] - var_off[p]; }
  int64_t nf(int64_t p) const { return arr ? arr[p].nf : fac_off[p + 1] - fac_off[p]; }
  const int32_t* vid(int64_t p) const { return arr ? arr[p].vid : vids + var_off[p]; }
  const int64_t* fid(int64_t p) const { return arr ? arr[p].fid : fids + fac_off[p]; }
};

namespace {

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

void fill_view(rdisgpu_ctx* c) {
  GraphView& g = c->gv;
  std::memset(&g, 0, sizeof g);
  g.kind = c->kind;
  g.V = c->V; g.F = c->F; g.E = c->E;
  g.xbd = c->xbd.p; g.xval = c->xval.p; g.dom = c->dom.p;
  g.rowptr = c->rowptr.p; g.evid = c->evid.p; g.expo = c->expo.p; g.konst = c->konst.p;
  g.sine = c->sine.p; g.coeff = c->coeff.p; g.vrow = c->vrow.p; g.vedge = c->vedge.p; g.efac = c->efac.p;
  g.tvrow = c->tvrow.p; g.eterm = c->eterm.p; g.t_expo = c->t_expo.p; g.t_konst = c->t_konst.p; g.t_sine = c->t_sine.p;
  g.vloc = c->vloc.p; g.floc = c->floc.p;
  g.cam = c->cam.p; g.pt = c->pt.p; g.obs = c->obs.p; g.ncams = c->ncams; g.npts = c->npts;
  g.crow = c->crow.p; g.cfac = c->cfac.p; g.prow = c->prow.p; g.pfac = c->pfac.p;
  g.fconst_on = c->has_fconst ? c->fconst_on.p : nullptr;
  g.fconst_val = c->has_fconst ? c->fconst_val.p : nullptr;
  g.fcache = c->strict ? c->fcache.p : nullptr;
  g.fdirty = c->strict ? c->fdirty.p : nullptr;
  g.vchg = c->strict ? c->vchg.p : nullptr;
  g.gedge = c->gedge.p; g.gvec = c->gvec.p; g.hvec = c->hvec.p; g.xsave = c->xsave.p; g.fstamp = c->fstamp.p;
}

template <class T>
cudaError_t upload(DevBuf<T>& d, const T* h, size_t n, cudaStream_t s) {
  cudaError_t e = d.ensure(n);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  return cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, s);
}

// true if `p` is device (or managed) memory this process can read from a kernel
bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// page-locked (cudaHostAlloc / cudaHostRegister) host memory: a cudaMemcpyAsync from it is truly asynchronous, the caller
// may not touch the buffer until the stream has passed the copy.  From pageable memory the driver stages the source
// before the call returns.
bool is_pinned_host_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// counting-sort incidence: for each key in [0,K) the list of items with that key, ascending item id
void build_incidence(int64_t K, const std::vector<int32_t>& key_of_item, std::vector<int32_t>& row,
                     std::vector<int32_t>& items) {
  row.assign(K + 1, 0);
  for (int32_t k : key_of_item) ++row[k + 1];
  for (int64_t i = 0; i < K; ++i) row[i + 1] += row[i];
  items.resize(key_of_item.size());
  std::vector<int32_t> cur(row.begin(), row.end() - 1);
  for (size_t it = 0; it < key_of_item.size(); ++it) items[cur[key_of_item[it]]++] = (int32_t)it;
}

}  // namespace

// ==========================================================================================
// lifetime
// ==========================================================================================
extern "C" {

int rdisgpu_create(rdisgpu_ctx** out, int device) {
  if (!out) return RDISGPU_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return RDISGPU_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return RDISGPU_ERR_ARG;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return RDISGPU_ERR_CUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return RDISGPU_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "this library carries sm_100a code only; device is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor);
    return RDISGPU_ERR_CUDA;
  }
  rdisgpu_ctx* c = new rdisgpu_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return RDISGPU_OK;
}

void rdisgpu_destroy(rdisgpu_ctx* ctx) {
  if (!ctx) return;
  if (ctx->live_batches > 0) {  // a batch holds a pointer to its context: destroying it first would leave it dangling
    ctx->err = "rdisgpu_destroy: " + std::to_string(ctx->live_batches) + " batch(es) still alive; destroy them first (context kept)";
    std::fprintf(stderr, "%s\n", ctx->err.c_str());
    return;
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  delete ctx->scratch_batch;
  delete ctx;
}

const char* rdisgpu_last_error(const rdisgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rdisgpu_set_stream(rdisgpu_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RDISGPU_ERR_ARG;
  ctx->stream = (cudaStream_t)cuda_stream;
  return RDISGPU_OK;
}

int rdisgpu_set_option(rdisgpu_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return RDISGPU_ERR_ARG;
  if (std::strcmp(name, "generic_only") == 0) {
    ctx->generic_only = (value != 0);
    return RDISGPU_OK;
  }
  if (std::strcmp(name, "adaptive_order") == 0) {
    ctx->adaptive_order = (value != 0);
    return RDISGPU_OK;
  }
  if (std::strcmp(name, "point_tiles_per_warp") == 0) {
    if (value < 0 || value > 32) return ctx->fail(RDISGPU_ERR_ARG, "set_option: point_tiles_per_warp must be 0..32");
    ctx->pt_tiles_cap = (int)value;
    return RDISGPU_OK;
  }
  if (std::strcmp(name, "camera_cluster") == 0) {
    if (value < 0 || value > kCamMaxCluster) return ctx->fail(RDISGPU_ERR_ARG, "set_option: camera_cluster must be 0..8");
    ctx->cam_cluster_opt = (int)value;
    return RDISGPU_OK;
  }
  if (std::strcmp(name, "strict") == 0) {
    if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "set_option: strict needs a finalized context");
    if (value != 0 && !ctx->strict) {
      // the reference's Factor objects start dirty (nothing evaluated yet) and its Variables unassigned: the first
      // assign of every variable notifies its factors, whatever the value
      CK(cudaSetDevice(ctx->device));
      CK(ctx->fcache.ensure((size_t)ctx->F));
      CK(ctx->fdirty.ensure((size_t)ctx->F));
      CK(ctx->vchg.ensure((size_t)ctx->V));
      CK(cudaMemsetAsync(ctx->fcache.p, 0, (size_t)ctx->F * sizeof(double), ctx->stream));
      CK(cudaMemsetAsync(ctx->fdirty.p, 1, (size_t)ctx->F, ctx->stream));
      CK(cudaMemsetAsync(ctx->vchg.p, 0, (size_t)ctx->V, ctx->stream));
    }
    ctx->strict = (value != 0);
    fill_view(ctx);
    return RDISGPU_OK;
  }
  if (std::strcmp(name, "resident_threads") == 0) {
    if (value != kResThreads && value != kResThreadsExact) return ctx->fail(RDISGPU_ERR_ARG, "set_option: resident_threads must be 256 or 1024");
    ctx->res_threads = (int)value;
    return RDISGPU_OK;
  }
  return ctx->fail(RDISGPU_ERR_ARG, "set_option: unknown option");
}

int rdisgpu_synchronize(rdisgpu_ctx* ctx) {
  if (!ctx) return RDISGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return RDISGPU_OK;
}

// ==========================================================================================
// definition
// ==========================================================================================
int rdisgpu_set_vars(rdisgpu_ctx* ctx, int64_t V, const double* lb, const double* ub) {
  if (!ctx || V <= 0 || !lb || !ub) return ctx ? ctx->fail(RDISGPU_ERR_ARG, "set_vars: bad argument") : RDISGPU_ERR_ARG;
  if (ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "set_vars after finalize");
  if (V > 0x7fffffffLL) return ctx->fail(RDISGPU_ERR_ARG, "set_vars: V exceeds int32 ids");
  ctx->V = V;
  ctx->h_lb.assign(lb, lb + V);
  ctx->h_ub.assign(ub, ub + V);
  return RDISGPU_OK;
}

int rdisgpu_add_nlpf(rdisgpu_ctx* ctx, int64_t F, const int64_t* rowptr, const int32_t* vid, const double* expo,
                     const double* konst, const uint8_t* use_sine, const double* coeff) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (ctx->finalized || ctx->kind != KIND_NONE) return ctx->fail(RDISGPU_ERR_STATE, "add_nlpf: factors already defined");
  if (F <= 0 || !rowptr || !coeff) return ctx->fail(RDISGPU_ERR_ARG, "add_nlpf: bad argument");
  const int64_t E = rowptr[F];
  if (E > 0 && (!vid || !expo || !konst || !use_sine)) return ctx->fail(RDISGPU_ERR_ARG, "add_nlpf: null edge array");
  if (E > 0x7fffffffLL || F > 0x7fffffffLL) return ctx->fail(RDISGPU_ERR_ARG, "add_nlpf: exceeds int32 ids");
  for (int64_t j = 0; j < F; ++j)
    if (rowptr[j + 1] < rowptr[j]) return ctx->fail(RDISGPU_ERR_ARG, "add_nlpf: rowptr not monotone");
  for (int64_t e = 0; e < E; ++e)
    if (vid[e] < 0 || vid[e] >= ctx->V) return ctx->fail(RDISGPU_ERR_ARG, "add_nlpf: variable id out of range");
  ctx->kind = KIND_NLPF;
  ctx->F = F;
  ctx->E = E;
  ctx->h_rowptr.assign(rowptr, rowptr + F + 1);
  ctx->h_evid.assign(vid, vid + E);
  ctx->h_expo.assign(expo, expo + E);
  ctx->h_konst.assign(konst, konst + E);
  ctx->h_sine.assign(use_sine, use_sine + E);
  ctx->h_coeff.assign(coeff, coeff + F);
  return RDISGPU_OK;
}

int rdisgpu_add_ba(rdisgpu_ctx* ctx, int64_t F, const int32_t* cam, const int32_t* pt, const double* obs_xy,
                   int32_t ncams, int32_t npts) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (ctx->finalized || ctx->kind != KIND_NONE) return ctx->fail(RDISGPU_ERR_STATE, "add_ba: factors already defined");
  if (F <= 0 || !cam || !pt || !obs_xy || ncams <= 0 || npts <= 0) return ctx->fail(RDISGPU_ERR_ARG, "add_ba: bad argument");
  if (ctx->V != 9LL * ncams + 3LL * npts) return ctx->fail(RDISGPU_ERR_ARG, "add_ba: V != 9*ncams + 3*npts");
  if (F * 12 > 0x7fffffffLL) return ctx->fail(RDISGPU_ERR_ARG, "add_ba: exceeds int32 edge ids");
  for (int64_t j = 0; j < F; ++j)
    if (cam[j] < 0 || cam[j] >= ncams || pt[j] < 0 || pt[j] >= npts)
      return ctx->fail(RDISGPU_ERR_ARG, "add_ba: camera / point id out of range");
  ctx->kind = KIND_BA;
  ctx->F = F;
  ctx->E = 12 * F;
  ctx->ncams = ncams;
  ctx->npts = npts;
  ctx->h_cam.assign(cam, cam + F);
  ctx->h_pt.assign(pt, pt + F);
  ctx->h_obs.assign(obs_xy, obs_xy + 2 * F);
  return RDISGPU_OK;
}

int rdisgpu_finalize(rdisgpu_ctx* ctx) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "finalize called twice");
  if (ctx->V <= 0 || ctx->kind == KIND_NONE) return ctx->fail(RDISGPU_ERR_STATE, "finalize: variables / factors missing");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t V = ctx->V, F = ctx->F, E = ctx->E;

  std::vector<double2> dom(V), xbd(V);
  const double qnan = std::nan("");
  for (int64_t v = 0; v < V; ++v) {
    dom[v] = make_double2(ctx->h_lb[v], ctx->h_ub[v]);
    xbd[v] = make_double2(0.0, qnan);
  }
  CK(upload(ctx->dom, dom.data(), (size_t)V, s));
  CK(upload(ctx->xbd, xbd.data(), (size_t)V, s));
  CK(ctx->xval.ensure((size_t)V));
  CK(cudaMemsetAsync(ctx->xval.p, 0, (size_t)V * sizeof(double), s));
  CK(cudaStreamSynchronize(s));

  if (ctx->kind == KIND_NLPF) {
    std::vector<int32_t> rp(F + 1), efac(E);
    for (int64_t j = 0; j <= F; ++j) rp[j] = (int32_t)ctx->h_rowptr[j];
    for (int64_t j = 0; j < F; ++j)
      for (int32_t e = rp[j]; e < rp[j + 1]; ++e) efac[e] = (int32_t)j;
    std::vector<int32_t> vrow, vedge;
    build_incidence(V, ctx->h_evid, vrow, vedge);  // items = edge ids, ascending = ascending factor id
    // the streamed arrays carry 32 elements of tail padding: bulk copies fetch 16-byte-aligned slices
    CK(ctx->rowptr.ensure(rp.size() + 32));
    CK(ctx->evid.ensure((size_t)E + 32));
    CK(ctx->expo.ensure((size_t)E + 32));
    CK(ctx->konst.ensure((size_t)E + 32));
    CK(ctx->sine.ensure((size_t)E + 32));
    CK(ctx->coeff.ensure((size_t)F + 32));
    CK(cudaMemsetAsync(ctx->rowptr.p, 0, (rp.size() + 32) * sizeof(int32_t), s));
    CK(cudaMemsetAsync(ctx->evid.p, 0, ((size_t)E + 32) * sizeof(int32_t), s));
    CK(cudaMemsetAsync(ctx->expo.p, 0, ((size_t)E + 32) * sizeof(double), s));
    CK(cudaMemsetAsync(ctx->konst.p, 0, ((size_t)E + 32) * sizeof(double), s));
    CK(cudaMemsetAsync(ctx->sine.p, 0, (size_t)E + 32, s));
    CK(cudaMemsetAsync(ctx->coeff.p, 0, ((size_t)F + 32) * sizeof(double), s));
    CK(upload(ctx->rowptr, rp.data(), rp.size(), s));
    CK(upload(ctx->evid, ctx->h_evid.data(), (size_t)E, s));
    CK(upload(ctx->expo, ctx->h_expo.data(), (size_t)E, s));
    CK(upload(ctx->konst, ctx->h_konst.data(), (size_t)E, s));
    CK(upload(ctx->sine, ctx->h_sine.data(), (size_t)E, s));
    CK(upload(ctx->coeff, ctx->h_coeff.data(), (size_t)F, s));
    CK(upload(ctx->vrow, vrow.data(), vrow.size(), s));
    CK(upload(ctx->vedge, vedge.data(), vedge.size(), s));
    CK(upload(ctx->efac, efac.data(), efac.size(), s));
    // term table: the distinct (k, e, sine) expressions each variable enters factors through, variable-major, in
    // order of first appearance (ascending factor id); compared bit for bit, so a shared term is the same number
    {
      std::vector<int32_t> tvrow((size_t)V + 1, 0), eterm((size_t)E, 0);
      std::vector<double> t_expo, t_konst;
      std::vector<uint8_t> t_sine;
      t_expo.reserve((size_t)V * 2); t_konst.reserve((size_t)V * 2); t_sine.reserve((size_t)V * 2);
      auto bits = [](double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; };
      for (int64_t v = 0; v < V; ++v) {
        const size_t base = t_expo.size();
        tvrow[v] = (int32_t)base;
        for (int32_t r = vrow[v]; r < vrow[v + 1]; ++r) {
          const int32_t e = vedge[r];
          const uint64_t kb = bits(ctx->h_konst[e]), eb = bits(ctx->h_expo[e]);
          const uint8_t sn = ctx->h_sine[e] ? 1 : 0;
          size_t u = base;
          for (; u < t_expo.size(); ++u)
            if (bits(t_konst[u]) == kb && bits(t_expo[u]) == eb && t_sine[u] == sn) break;
          if (u == t_expo.size()) {
            if (u >= 0x7fffffffULL) return ctx->fail(RDISGPU_ERR_ARG, "finalize: too many terms");
            t_expo.push_back(ctx->h_expo[e]); t_konst.push_back(ctx->h_konst[e]); t_sine.push_back(sn);
          }
          eterm[e] = (int32_t)u;
        }
      }
      tvrow[V] = (int32_t)t_expo.size();
      if (t_expo.empty()) { t_expo.push_back(0.0); t_konst.push_back(0.0); t_sine.push_back(0); }
      CK(upload(ctx->tvrow, tvrow.data(), tvrow.size(), s));
      CK(upload(ctx->eterm, eterm.data(), eterm.size(), s));
      CK(upload(ctx->t_expo, t_expo.data(), t_expo.size(), s));
      CK(upload(ctx->t_konst, t_konst.data(), t_konst.size(), s));
      CK(upload(ctx->t_sine, t_sine.data(), t_sine.size(), s));
      CK(ctx->vloc.ensure((size_t)V));
      CK(ctx->floc.ensure((size_t)std::max<int64_t>(F, 1)));
      CK(cudaStreamSynchronize(s));  // the staging vectors die with this scope
      ctx->h_tvrow.swap(tvrow);
      ctx->h_rp32 = rp;
      ctx->h_evid32 = ctx->h_evid;
    }
    // factor tiles of the streaming sweep: consecutive factors, <= kTileFactors of them, <= kTileEdges edges
    std::vector<TileDesc> tiles;
    for (int64_t f = 0; f < F;) {
      const int64_t eb = rp[f];
      int64_t g = f;
      while (g < F && g - f < kTileFactors && rp[g + 1] - eb <= kTileEdges) ++g;
      if (g == f) g = f + 1;  // a single factor wider than a tile: folded serially
      tiles.push_back(TileDesc{(int32_t)f, (int32_t)g, rp[f], rp[g]});
      f = g;
    }
    ctx->ntiles = (int)tiles.size();
    CK(upload(ctx->tiles, tiles.data(), tiles.size(), s));
    // persistent grids: as many CTAs as fit, from the occupancy calculator (dynamic shared memory opt-in)
    {
      const int smem_e = (int)sizeof(TileSmem<false>), smem_g = (int)sizeof(TileSmem<true>);
      CK(cudaFuncSetAttribute(nlpf_tile_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_e));
      CK(cudaFuncSetAttribute(nlpf_tile_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_g));
      int occ_e = 0, occ_g = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, nlpf_tile_sweep_kernel<false>, kTileBlock, smem_e));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, nlpf_tile_sweep_kernel<true>, kTileBlock, smem_g));
      if (occ_e < 1 || occ_g < 1) return ctx->fail(RDISGPU_ERR_CUDA, "streaming sweep kernel cannot be resident");
      ctx->tile_grid[0] = std::min(ctx->ntiles, occ_e * ctx->sm_count);
      ctx->tile_grid[1] = std::min(ctx->ntiles, occ_g * ctx->sm_count);
    }
    CK(cudaStreamSynchronize(s));
  } else {
    std::vector<int32_t> crow, cfac, prow, pfac;
    build_incidence(ctx->ncams, ctx->h_cam, crow, cfac);
    build_incidence(ctx->npts, ctx->h_pt, prow, pfac);
    std::vector<double2> obs(F);
    for (int64_t j = 0; j < F; ++j) obs[j] = make_double2(ctx->h_obs[2 * j], ctx->h_obs[2 * j + 1]);
    CK(upload(ctx->cam, ctx->h_cam.data(), (size_t)F, s));
    CK(upload(ctx->pt, ctx->h_pt.data(), (size_t)F, s));
    CK(upload(ctx->obs, obs.data(), (size_t)F, s));
    CK(upload(ctx->crow, crow.data(), crow.size(), s));
    CK(upload(ctx->cfac, cfac.data(), cfac.size(), s));
    CK(upload(ctx->prow, prow.data(), prow.size(), s));
    CK(upload(ctx->pfac, pfac.data(), pfac.size(), s));
    CK(cudaStreamSynchronize(s));
  }
  CK(ctx->gedge.ensure((size_t)std::max<int64_t>(E, 1)));
  CK(ctx->gvec.ensure((size_t)V));
  CK(ctx->hvec.ensure((size_t)V));
  CK(ctx->xsave.ensure((size_t)V));
  CK(ctx->fstamp.ensure((size_t)F));
  CK(cudaMemsetAsync(ctx->fstamp.p, 0xff, (size_t)F * sizeof(int32_t), s));  // -1
  CK(ctx->s_counter.ensure(4));
  CK(cudaMemsetAsync(ctx->s_counter.p, 0, 4 * sizeof(unsigned int), s));
  CK(cudaStreamSynchronize(s));

  // the host copy of the definition is no longer needed
  std::vector<double>().swap(ctx->h_lb);
  std::vector<double>().swap(ctx->h_ub);
  std::vector<int64_t>().swap(ctx->h_rowptr);
  std::vector<int32_t>().swap(ctx->h_evid);
  std::vector<double>().swap(ctx->h_expo);
  std::vector<double>().swap(ctx->h_konst);
  std::vector<double>().swap(ctx->h_coeff);
  std::vector<uint8_t>().swap(ctx->h_sine);
  // h_cam / h_pt stay: batch_create classifies point / camera blocks on the host with them
  std::vector<double>().swap(ctx->h_obs);

  ctx->finalized = true;
  fill_view(ctx);
  return RDISGPU_OK;
}

// ==========================================================================================
// state
// ==========================================================================================
int rdisgpu_set_x(rdisgpu_ctx* ctx, int64_t n, const int32_t* vid, const double* x) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "set_x before finalize");
  if (n < 0 || (n > 0 && !x) || (!vid && n > ctx->V)) return ctx->fail(RDISGPU_ERR_ARG, "set_x: bad argument");
  if (n == 0) return RDISGPU_OK;
  CK(cudaSetDevice(ctx->device));
  if (is_device_ptr(x)) {
    // device-resident values (and ids, if given): no copy, no host sync — stays asynchronous
    if (vid && !is_device_ptr(vid)) return ctx->fail(RDISGPU_ERR_ARG, "set_x: device x needs device (or null) vid");
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 65535);
    if (ctx->strict)
      scatter_x_strict_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->gv, n, vid, x);
    else
      scatter_x_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->gv, n, vid, x);
    ++ctx->launches;
    CK(cudaGetLastError());
    return RDISGPU_OK;
  }
  if (vid)
    for (int64_t i = 0; i < n; ++i)
      if (vid[i] < 0 || vid[i] >= ctx->V) return ctx->fail(RDISGPU_ERR_ARG, "set_x: variable id out of range");
  cudaStream_t s = ctx->stream;
  CK(ctx->s_f64a.ensure((size_t)n));
  CK(cudaMemcpyAsync(ctx->s_f64a.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  if (vid) {
    CK(ctx->s_i32a.ensure((size_t)n));
    CK(cudaMemcpyAsync(ctx->s_i32a.p, vid, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  }
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 65535);
  if (ctx->strict)
    scatter_x_strict_kernel<<<blocks, threads, 0, s>>>(ctx->gv, n, vid ? ctx->s_i32a.p : nullptr, ctx->s_f64a.p);
  else
    scatter_x_kernel<<<blocks, threads, 0, s>>>(ctx->gv, n, vid ? ctx->s_i32a.p : nullptr, ctx->s_f64a.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  // caller-owned buffers: pageable sources have been staged by the driver when cudaMemcpyAsync returns; page-locked ones
  // are read by the DMA engine later, so the call waits for them
  if (is_pinned_host_ptr(x) || is_pinned_host_ptr(vid)) CK(cudaStreamSynchronize(s));
  return RDISGPU_OK;
}

int rdisgpu_get_x(rdisgpu_ctx* ctx, int64_t n, const int32_t* vid, double* x) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "get_x before finalize");
  if (n < 0 || (n > 0 && !x) || (!vid && n > ctx->V)) return ctx->fail(RDISGPU_ERR_ARG, "get_x: bad argument");
  if (n == 0) return RDISGPU_OK;
  if (vid)
    for (int64_t i = 0; i < n; ++i)
      if (vid[i] < 0 || vid[i] >= ctx->V) return ctx->fail(RDISGPU_ERR_ARG, "get_x: variable id out of range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(ctx->s_f64a.ensure((size_t)n));
  if (vid) {
    CK(ctx->s_i32a.ensure((size_t)n));
    CK(cudaMemcpyAsync(ctx->s_i32a.p, vid, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  }
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 65535);
  gather_x_kernel<<<blocks, threads, 0, s>>>(ctx->gv, n, vid ? ctx->s_i32a.p : nullptr, ctx->s_f64a.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(x, ctx->s_f64a.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return RDISGPU_OK;
}

int rdisgpu_set_factor_const(rdisgpu_ctx* ctx, int64_t n, const int64_t* fid, const double* val, const uint8_t* on) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "set_factor_const before finalize");
  if (n < 0 || (n > 0 && (!fid || !val || !on))) return ctx->fail(RDISGPU_ERR_ARG, "set_factor_const: bad argument");
  for (int64_t i = 0; i < n; ++i)
    if (fid[i] < 0 || fid[i] >= ctx->F) return ctx->fail(RDISGPU_ERR_ARG, "set_factor_const: factor id out of range");
  if (n == 0) return RDISGPU_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (!ctx->has_fconst) {
    CK(ctx->fconst_on.ensure((size_t)ctx->F));
    CK(ctx->fconst_val.ensure((size_t)ctx->F));
    CK(cudaMemsetAsync(ctx->fconst_on.p, 0, (size_t)ctx->F, s));
    CK(cudaMemsetAsync(ctx->fconst_val.p, 0, (size_t)ctx->F * sizeof(double), s));
    ctx->has_fconst = true;
    fill_view(ctx);
  }
  std::vector<int32_t> f32(n);
  for (int64_t i = 0; i < n; ++i) f32[i] = (int32_t)fid[i];
  CK(ctx->s_i32a.ensure((size_t)n));
  CK(ctx->s_f64a.ensure((size_t)n));
  CK(ctx->s_i32b.ensure((size_t)(n + 3) / 4 + 1));
  CK(cudaMemcpyAsync(ctx->s_i32a.p, f32.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->s_f64a.p, val, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->s_i32b.p, on, (size_t)n, cudaMemcpyHostToDevice, s));
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 65535);
  set_fconst_kernel<<<blocks, threads, 0, s>>>(ctx->fconst_on.p, ctx->fconst_val.p, n, ctx->s_i32a.p, ctx->s_f64a.p,
                                               (const uint8_t*)ctx->s_i32b.p);
  ++ctx->launches;
  if (ctx->strict) {  // Factor::setAssignedConstant(false): the factor must be recomputed at its next eval()
    unset_const_dirty_kernel<<<blocks, threads, 0, s>>>(ctx->fdirty.p, n, ctx->s_i32a.p, (const uint8_t*)ctx->s_i32b.p);
    ++ctx->launches;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  return RDISGPU_OK;
}

// ==========================================================================================
// sweeps
// ==========================================================================================
static int upload_fids(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, DevBuf<int32_t>& dst) {
  for (int64_t i = 0; i < nf; ++i)
    if (fid[i] < 0 || fid[i] >= ctx->F) return ctx->fail(RDISGPU_ERR_ARG, "factor id out of range");
  std::vector<int32_t> f32(nf);
  for (int64_t i = 0; i < nf; ++i) f32[i] = (int32_t)fid[i];
  CK(dst.ensure((size_t)nf));
  CK(cudaMemcpyAsync(dst.p, f32.data(), (size_t)nf * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return RDISGPU_OK;
}

// The all-factor bundle-adjustment sweep (ba_sweep.cuh): values, optionally the 12 partials of every observation.
static int launch_ba_sweep(rdisgpu_ctx* ctx, int blocks, double* per_factor_dev, double* rows_dev, double* dsum) {
  cudaStream_t s = ctx->stream;
  const int threads = 256;
  if (ctx->ncams <= kBaSmemCams) {
    const size_t smem = (size_t)ctx->ncams * kBaSmemRow * sizeof(double);
    if (smem > 48 * 1024 && !ctx->ba_smem_optin) {
      CK(cudaFuncSetAttribute(ba_sweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBaSmemCams * kBaSmemRow * sizeof(double))));
      CK(cudaFuncSetAttribute(ba_sweep_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBaSmemCams * kBaSmemRow * sizeof(double))));
      ctx->ba_smem_optin = true;
    }
    if (rows_dev)
      ba_sweep_kernel<true, true><<<blocks, threads, smem, s>>>(ctx->gv, nullptr, per_factor_dev, rows_dev, ctx->s_partials.p, ctx->s_counter.p, dsum);
    else
      ba_sweep_kernel<true, false><<<blocks, threads, smem, s>>>(ctx->gv, nullptr, per_factor_dev, nullptr, ctx->s_partials.p, ctx->s_counter.p, dsum);
  } else {
    CK(ctx->cam_table.ensure((size_t)ctx->ncams));
    ba_camera_table_kernel<<<(ctx->ncams + 127) / 128, 128, 0, s>>>(ctx->gv, ctx->cam_table.p);
    ++ctx->launches;
    if (rows_dev)
      ba_sweep_kernel<false, true><<<blocks, threads, 0, s>>>(ctx->gv, ctx->cam_table.p, per_factor_dev, rows_dev, ctx->s_partials.p, ctx->s_counter.p, dsum);
    else
      ba_sweep_kernel<false, false><<<blocks, threads, 0, s>>>(ctx->gv, ctx->cam_table.p, per_factor_dev, nullptr, ctx->s_partials.p, ctx->s_counter.p, dsum);
  }
  return RDISGPU_OK;
}

// Enqueues one residual sweep: fid_dev (device, nullable = all factors), per_factor_dev (device, nullable);
// the total lands in *dsum_out (device scratch of the context).  No copies, no waiting.
static int enqueue_eval(rdisgpu_ctx* ctx, int64_t nf, const int32_t* fid_dev, double* per_factor_dev, double** dsum_out,
                        double* sum_dst = nullptr /* device: write the total here instead of the context scratch */) {
  cudaStream_t s = ctx->stream;
  const int threads = 256;
  const bool tiled = (ctx->kind == KIND_NLPF && fid_dev == nullptr);
  const bool ba_all = (ctx->kind == KIND_BA && fid_dev == nullptr);
  const int blocks = tiled ? ctx->tile_grid[0]
                           : (int)std::min<int64_t>(((ba_all ? (nf + kBaSweepUnroll - 1) / kBaSweepUnroll : nf) + threads - 1) / threads,
                                                    (int64_t)ctx->sm_count * (ba_all ? RDIS_BA_SWEEP_CTAS : 8));
  CK(ctx->s_partials.ensure((size_t)blocks + 1));
  double* dsum = sum_dst ? sum_dst : ctx->s_partials.p + blocks;
  if (tiled)
    nlpf_tile_sweep_kernel<false><<<blocks, kTileBlock, sizeof(TileSmem<false>), s>>>(
        ctx->gv, ctx->tiles.p, ctx->ntiles, per_factor_dev, ctx->s_partials.p, ctx->s_counter.p, dsum);
  else if (ba_all) {
    int rc = launch_ba_sweep(ctx, blocks, per_factor_dev, nullptr, dsum);
    if (rc) return rc;
  } else if (ctx->kind == KIND_NLPF)
    eval_sweep_kernel<NlpfOps><<<blocks, threads, 0, s>>>(ctx->gv, fid_dev, nf, per_factor_dev, ctx->s_partials.p,
                                                          ctx->s_counter.p, dsum);
  else
    eval_sweep_kernel<BaOps><<<blocks, threads, 0, s>>>(ctx->gv, fid_dev, nf, per_factor_dev, ctx->s_partials.p,
                                                        ctx->s_counter.p, dsum);
  ++ctx->launches;
  CK(cudaGetLastError());
  *dsum_out = dsum;
  return RDISGPU_OK;
}

int rdisgpu_eval(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, double* sum, double* per_factor) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "eval before finalize");
  if (!fid) nf = ctx->F;
  if (nf < 0) return ctx->fail(RDISGPU_ERR_ARG, "eval: bad argument");
  if (nf == 0) {
    if (sum) *sum = 0.0;
    return RDISGPU_OK;
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (fid) {
    int rc = upload_fids(ctx, nf, fid, ctx->s_i32a);
    if (rc) return rc;
  }
  if (per_factor) CK(ctx->s_f64a.ensure((size_t)nf));
  double* dsum = nullptr;
  int rc = enqueue_eval(ctx, nf, fid ? ctx->s_i32a.p : nullptr, per_factor ? ctx->s_f64a.p : nullptr, &dsum);
  if (rc) return rc;
  double hsum = 0.0;
  CK(cudaMemcpyAsync(&hsum, dsum, sizeof(double), cudaMemcpyDeviceToHost, s));
  if (per_factor) CK(cudaMemcpyAsync(per_factor, ctx->s_f64a.p, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (sum) *sum = hsum;
  return RDISGPU_OK;
}

int rdisgpu_eval_device(rdisgpu_ctx* ctx, int64_t nf, const int32_t* fid_dev, double* sum_dev, double* per_factor_dev) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "eval_device before finalize");
  if (!fid_dev) nf = ctx->F;
  if (nf <= 0) return ctx->fail(RDISGPU_ERR_ARG, "eval_device: bad argument");
  if ((fid_dev && !is_device_ptr(fid_dev)) || (sum_dev && !is_device_ptr(sum_dev)) ||
      (per_factor_dev && !is_device_ptr(per_factor_dev)))
    return ctx->fail(RDISGPU_ERR_ARG, "eval_device: every pointer must be device memory");
  CK(cudaSetDevice(ctx->device));
  double* dsum = nullptr;
  return enqueue_eval(ctx, nf, fid_dev, per_factor_dev, &dsum, sum_dev);  // the kernel writes the total straight to sum_dev
}

// Enqueues computeGradient over a factor list: phase A writes every listed factor's partials to gedge
// (the streaming tile kernel when the list is "all NLPF factors"), phase B gathers them per variable in
// ascending factor id.  Device pointers only; no copies, no waiting.
static int enqueue_grad(rdisgpu_ctx* ctx, int64_t nf, const int32_t* fid_dev, int64_t nv, const int32_t* vid_dev,
                        double* g_dev) {
  cudaStream_t s = ctx->stream;
  const bool filter = (fid_dev != nullptr);
  const int32_t stamp = 0x40000000;  // outside the range of batch problem indices
  const int threads = 256;
  if (nf > 0) {
    if (ctx->kind == KIND_NLPF && !filter) {
      const int tg = ctx->tile_grid[1];
      CK(ctx->s_partials.ensure((size_t)tg + 1));
      nlpf_tile_sweep_kernel<true><<<tg, kTileBlock, sizeof(TileSmem<true>), s>>>(
          ctx->gv, ctx->tiles.p, ctx->ntiles, nullptr, ctx->s_partials.p, ctx->s_counter.p, ctx->s_partials.p + tg);
    } else {
      const int fb = (int)std::min<int64_t>((nf + threads - 1) / threads, (int64_t)ctx->sm_count * 8);
      if (ctx->kind == KIND_NLPF)
        factor_partials_kernel<NlpfOps><<<fb, threads, 0, s>>>(ctx->gv, fid_dev, nf, filter ? stamp : -1);
      else
        factor_partials_kernel<BaOps><<<fb, threads, 0, s>>>(ctx->gv, fid_dev, nf, filter ? stamp : -1);
    }
    ++ctx->launches;
    CK(cudaGetLastError());
  }
#ifndef RDIS_GATHER_BLOCKS_PER_SM
#define RDIS_GATHER_BLOCKS_PER_SM 16
#endif
  const int vb = (int)std::min<int64_t>((nv + threads - 1) / threads, (int64_t)ctx->sm_count * RDIS_GATHER_BLOCKS_PER_SM);
  // with an explicit list an unlisted factor must not contribute; with nf == 0 nothing does
  const bool eff_filter = filter || nf == 0;
  if (ctx->kind == KIND_NLPF)
    gather_grad_kernel<NlpfOps><<<vb, threads, 0, s>>>(ctx->gv, vid_dev, nv, stamp, eff_filter, g_dev);
  else
    gather_grad_kernel<BaOps><<<vb, threads, 0, s>>>(ctx->gv, vid_dev, nv, stamp, eff_filter, g_dev);
  ++ctx->launches;
  CK(cudaGetLastError());
  if (filter && nf > 0) {
    const int fb = (int)std::min<int64_t>((nf + threads - 1) / threads, 65535);
    unstamp_kernel<<<fb, threads, 0, s>>>(ctx->gv, fid_dev, nf);
    ++ctx->launches;
    CK(cudaGetLastError());
  }
  return RDISGPU_OK;
}

int rdisgpu_grad(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, int64_t nv, const int32_t* vid, double* g) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "grad before finalize");
  if (!fid) nf = ctx->F;
  if (!vid) nv = ctx->V;
  if (nf < 0 || nv < 0 || (nv > 0 && !g)) return ctx->fail(RDISGPU_ERR_ARG, "grad: bad argument");
  if (nv == 0) return RDISGPU_OK;
  if (vid)
    for (int64_t i = 0; i < nv; ++i)
      if (vid[i] < 0 || vid[i] >= ctx->V) return ctx->fail(RDISGPU_ERR_ARG, "grad: variable id out of range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (fid) {
    int rc = upload_fids(ctx, nf, fid, ctx->s_i32a);
    if (rc) return rc;
  }
  if (vid) {
    CK(ctx->s_i32b.ensure((size_t)nv));
    CK(cudaMemcpyAsync(ctx->s_i32b.p, vid, (size_t)nv * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  }
  CK(ctx->s_f64a.ensure((size_t)nv));
  int rc = enqueue_grad(ctx, nf, fid ? ctx->s_i32a.p : nullptr, nv, vid ? ctx->s_i32b.p : nullptr, ctx->s_f64a.p);
  if (rc) return rc;
  CK(cudaMemcpyAsync(g, ctx->s_f64a.p, (size_t)nv * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return RDISGPU_OK;
}

int rdisgpu_grad_device(rdisgpu_ctx* ctx, int64_t nf, const int32_t* fid_dev, int64_t nv, const int32_t* vid_dev,
                        double* g_dev) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "grad_device before finalize");
  if (!fid_dev) nf = ctx->F;
  if (!vid_dev) nv = ctx->V;
  if (nf < 0 || nv <= 0 || !g_dev) return ctx->fail(RDISGPU_ERR_ARG, "grad_device: bad argument");
  if ((fid_dev && !is_device_ptr(fid_dev)) || (vid_dev && !is_device_ptr(vid_dev)) || !is_device_ptr(g_dev))
    return ctx->fail(RDISGPU_ERR_ARG, "grad_device: every pointer must be device memory");
  CK(cudaSetDevice(ctx->device));
  return enqueue_grad(ctx, nf, fid_dev, nv, vid_dev, g_dev);
}

int rdisgpu_factor_grad(rdisgpu_ctx* ctx, int64_t nf, const int64_t* fid, int32_t arity_max, double* rows) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "factor_grad before finalize");
  if (!fid) nf = ctx->F;
  if (nf <= 0 || arity_max <= 0 || !rows) return ctx->fail(RDISGPU_ERR_ARG, "factor_grad: bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (fid) {
    int rc = upload_fids(ctx, nf, fid, ctx->s_i32a);
    if (rc) return rc;
  }
  CK(ctx->s_f64b.ensure((size_t)nf * arity_max));
  CK(cudaMemsetAsync(ctx->s_f64b.p, 0, (size_t)nf * arity_max * sizeof(double), s));
  const int threads = 128;
  const int blocks = (int)std::min<int64_t>((nf + threads - 1) / threads, (int64_t)ctx->sm_count * 8);
  if (ctx->kind == KIND_NLPF)
    factor_rows_kernel<NlpfOps><<<blocks, threads, 0, s>>>(ctx->gv, fid ? ctx->s_i32a.p : nullptr, nf, arity_max, ctx->s_f64b.p);
  else
    factor_rows_kernel<BaOps><<<blocks, threads, 0, s>>>(ctx->gv, fid ? ctx->s_i32a.p : nullptr, nf, arity_max, ctx->s_f64b.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(rows, ctx->s_f64b.p, (size_t)nf * arity_max * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return RDISGPU_OK;
}

int rdisgpu_factor_rows_device(rdisgpu_ctx* ctx, double* sum_dev, double* per_factor_dev, double* rows_dev) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "factor_rows_device before finalize");
  if (ctx->kind != KIND_BA) return ctx->fail(RDISGPU_ERR_ARG, "factor_rows_device: bundle-adjustment graphs only (use rdisgpu_grad_device)");
  if (!is_device_ptr(rows_dev) || (per_factor_dev && !is_device_ptr(per_factor_dev)) || (sum_dev && !is_device_ptr(sum_dev)))
    return ctx->fail(RDISGPU_ERR_ARG, "factor_rows_device: device pointers required");
  CK(cudaSetDevice(ctx->device));
  const int blocks = (int)std::min<int64_t>(((ctx->F + kBaSweepUnroll - 1) / kBaSweepUnroll + 255) / 256, (int64_t)ctx->sm_count * 2);
  CK(ctx->s_partials.ensure((size_t)blocks + 1));
  int rc = launch_ba_sweep(ctx, blocks, per_factor_dev, rows_dev, sum_dev ? sum_dev : ctx->s_partials.p + blocks);
  if (rc) return rc;
  ++ctx->launches;
  CK(cudaGetLastError());
  return RDISGPU_OK;
}

// ==========================================================================================
// subspace solves
// ==========================================================================================
// (Re)builds `b` in place from a problem description: validation, sibling check, size classes,
// block-shape classification, and ONE upload of all index lists.  Device / pinned buffers only grow,
// so a batch object that is rebuilt call after call (the context's scratch batch) allocates nothing
// in steady state.
static int batch_build(rdisgpu_batch* b, const ProblemsView& pv) {
  rdisgpu_ctx* ctx = b->ctx;
  const int64_t nprobs = pv.n;
  b->nprobs = nprobs;
  b->h_probs.resize(nprobs);
  b->classes.clear();
  b->h_order.clear();
  b->n_pt_warps = b->n_cam = b->cam_nf_max = 0;
  b->cam_C = b->cam_T = 0;
  b->solved = false;
  int64_t tv = 0, tf = 0;
  for (int64_t p = 0; p < nprobs; ++p) {
    const int64_t nv = pv.nv(p), nf = pv.nf(p);
    if (nv < 0 || nf < 0 || (nv > 0 && !pv.vid(p)) || (nf > 0 && !pv.fid(p)) || nv > 0x7fffffffLL || nf > 0x7fffffffLL)
      return ctx->fail(RDISGPU_ERR_ARG, "batch: malformed problem");
    b->h_probs[p] = ProblemDesc{tv, tf, (int32_t)nv, (int32_t)nf, 0, 0, 0, 0};
    tv += nv;
    tf += nf;
  }
  if (tv > 0x7fffffffLL || tf > 0x7fffffffLL) return ctx->fail(RDISGPU_ERR_ARG, "batch: too many variables / factors");
  b->total_nv = tv;
  b->total_nf = tf;

  // upper bounds of every list, then carve the pinned staging blob
  const size_t npt_tasks_max = (size_t)nprobs + 6;  // one task per problem at worst (point_tiles_per_warp = 1)
  auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t o_probs = 0;
  const size_t o_vids = align16(o_probs + sizeof(ProblemDesc) * (size_t)nprobs);
  const size_t o_fids = align16(o_vids + 4 * (size_t)tv);
  const size_t o_order = align16(o_fids + 4 * (size_t)tf);
  const size_t o_pt = align16(o_order + 4 * (size_t)nprobs);
  const size_t o_cam = align16(o_pt + 4 * (size_t)nprobs);
  const size_t o_tasks = align16(o_cam + 4 * (size_t)nprobs);
  const size_t o_tasks2 = align16(o_tasks + sizeof(PointWarpTask) * npt_tasks_max);  // the re-sorted list of the next visit
  const size_t total = align16(o_tasks2 + sizeof(PointWarpTask) * npt_tasks_max);
  CK(b->h_blob.ensure(total));
  CK(b->blob.ensure(total));
  char* hb = b->h_blob.p;
  ProblemDesc* h_probs = reinterpret_cast<ProblemDesc*>(hb + o_probs);
  int32_t* vids = reinterpret_cast<int32_t*>(hb + o_vids);
  int32_t* fids = reinterpret_cast<int32_t*>(hb + o_fids);
  int32_t* h_order = reinterpret_cast<int32_t*>(hb + o_order);
  int32_t* h_pt = reinterpret_cast<int32_t*>(hb + o_pt);
  int32_t* h_cam = reinterpret_cast<int32_t*>(hb + o_cam);
  PointWarpTask* h_tasks = reinterpret_cast<PointWarpTask*>(hb + o_tasks);
  std::memcpy(h_probs, b->h_probs.data(), sizeof(ProblemDesc) * (size_t)nprobs);

  // sibling check: no variable and no factor may belong to two problems of one batch
  if ((int64_t)ctx->vmark.size() != ctx->V) ctx->vmark.assign((size_t)ctx->V, 0);
  if ((int64_t)ctx->fmark.size() != ctx->F) ctx->fmark.assign((size_t)ctx->F, 0);
  if ((int64_t)ctx->vowner.size() != ctx->V) ctx->vowner.assign((size_t)ctx->V, 0);
  if (++ctx->mark_epoch == 0x7fffffff) {
    std::fill(ctx->vmark.begin(), ctx->vmark.end(), 0);
    std::fill(ctx->fmark.begin(), ctx->fmark.end(), 0);
    ctx->mark_epoch = 1;
  }
  const int32_t epoch = ctx->mark_epoch;
  const int32_t first_point_var = 9 * ctx->ncams;  // bundle adjustment: which sides own variables in this batch
  bool owns_cam_vars = false, owns_pt_vars = false;
  for (int64_t p = 0; p < nprobs; ++p) {
    const ProblemDesc& D = b->h_probs[p];
    const int32_t* pvid = pv.vid(p);
    const int64_t* pfid = pv.fid(p);
    if (D.nv > 0 && ctx->kind == KIND_BA) {
      owns_cam_vars = owns_cam_vars || pvid[0] < first_point_var;       // ids are checked below; a mixed problem sets both
      owns_pt_vars = owns_pt_vars || pvid[D.nv - 1] >= first_point_var;
      if (!(owns_cam_vars && owns_pt_vars))
        for (int64_t j = 0; j < D.nv; ++j) (pvid[j] < first_point_var ? owns_cam_vars : owns_pt_vars) = true;
    }
    for (int64_t j = 0; j < D.nv; ++j) {
      const int32_t v = pvid[j];
      if (v < 0 || v >= ctx->V) return ctx->fail(RDISGPU_ERR_ARG, "batch: variable id out of range");
      if (ctx->vmark[v] == epoch) return ctx->fail(RDISGPU_ERR_OVERLAP, "batch: a variable appears twice in the batch");
      ctx->vmark[v] = epoch;
      ctx->vowner[v] = (int32_t)p;
      vids[D.var_off + j] = v;
    }
    for (int64_t k = 0; k < D.nf; ++k) {
      const int64_t f = pfid[k];
      if (f < 0 || f >= ctx->F) return ctx->fail(RDISGPU_ERR_ARG, "batch: factor id out of range");
      if (ctx->fmark[f] == epoch) return ctx->fail(RDISGPU_ERR_OVERLAP, "batch: a factor appears twice in the batch");
      ctx->fmark[f] = epoch;
      fids[D.fac_off + k] = (int32_t)f;
    }
  }

  // sibling check, second half: a factor of problem p may read variables of p and FROZEN variables only.  A variable
  // owned by another problem of the batch would be rewritten concurrently by that problem's thread group (the
  // reference sees it as an assigned, fixed value): not a sibling set -> RDISGPU_ERR_OVERLAP.
  {
    const bool ba = (ctx->kind == KIND_BA);
    const int32_t pbase = 9 * ctx->ncams;
    std::atomic<int> foreign{0};
    auto check_range = [&](int64_t p0, int64_t p1) {
      for (int64_t p = p0; p < p1 && !foreign.load(std::memory_order_relaxed); ++p) {
        const ProblemDesc& D = b->h_probs[p];
        const int32_t* pf = fids + D.fac_off;
        for (int k = 0; k < D.nf; ++k) {
          if (ba) {  // only the sides that own variables in this batch can hold a foreign owner
            const int32_t cb = 9 * ctx->h_cam[pf[k]], qb = pbase + 3 * ctx->h_pt[pf[k]];
            for (int sl = owns_cam_vars ? 0 : 9; sl < (owns_pt_vars ? 12 : 9); ++sl) {
              const int32_t v = (sl < 9) ? cb + sl : qb + (sl - 9);
              if (ctx->vmark[v] == epoch && ctx->vowner[v] != (int32_t)p) foreign.store(1, std::memory_order_relaxed);
            }
          } else {
            for (int32_t e = ctx->h_rp32[pf[k]]; e < ctx->h_rp32[pf[k] + 1]; ++e) {
              const int32_t v = ctx->h_evid32[e];
              if (ctx->vmark[v] == epoch && ctx->vowner[v] != (int32_t)p) foreign.store(1, std::memory_order_relaxed);
            }
          }
        }
      }
    };
    const unsigned hwc = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (tf >= 200000 && nprobs >= 2 * (int64_t)hwc && hwc > 1) {
      std::vector<std::thread> pool;
      const int64_t per = (nprobs + hwc - 1) / hwc;
      for (unsigned t = 0; t < hwc; ++t) {
        const int64_t p0 = std::min<int64_t>(nprobs, (int64_t)t * per), p1 = std::min<int64_t>(nprobs, p0 + per);
        if (p0 < p1) pool.emplace_back(check_range, p0, p1);
      }
      for (std::thread& th : pool) th.join();
    } else {
      check_range(0, nprobs);
    }
    if (foreign.load())
      return ctx->fail(RDISGPU_ERR_OVERLAP, "batch: a factor of one problem reads a variable owned by another problem of the batch");
  }

  // size classes: tiles of 1..32 lanes, CTAs of 64..256 threads, cooperative grid
  const int kTileMax = 32, kBlockMax = 4096;
  std::vector<std::vector<int32_t>> tile_lists(6), block_lists(3);
  std::vector<int32_t> grid_list;
  // bundle-adjustment block shapes get the register / cluster resident kernels
  std::vector<uint8_t> fast((size_t)nprobs, 0);  // 1 = point block, 2 = camera block
  if (ctx->kind == KIND_BA && !ctx->generic_only) {
    const int32_t pbase = 9 * ctx->ncams;
    std::vector<int32_t> pt_lists[6];
    for (int64_t p = 0; p < nprobs; ++p) {
      const ProblemDesc& D = b->h_probs[p];
      const int32_t* pvv = vids + D.var_off;
      const int32_t* pf = fids + D.fac_off;
      if (D.nf < 1) continue;
      bool asc = true;
      for (int k = 1; k < D.nf && asc; ++k) asc = pf[k] > pf[k - 1];
      if (!asc) continue;
      if (D.nv == 3 && D.nf <= 32 && pvv[0] >= pbase && (pvv[0] - pbase) % 3 == 0 && pvv[1] == pvv[0] + 1 && pvv[2] == pvv[0] + 2) {
        const int32_t pt = (pvv[0] - pbase) / 3;
        bool ok = true;
        for (int k = 0; k < D.nf && ok; ++k) ok = (ctx->h_pt[pf[k]] == pt);
        if (!ok) continue;
        int lg = 0;
        while ((1 << lg) < D.nf) ++lg;
        pt_lists[lg].push_back((int32_t)p);
        fast[p] = 1;
      } else if (D.nv == 9 && D.nf <= kCamMaxCluster * kCamMaxThreads && pvv[0] < pbase && pvv[0] % 9 == 0) {
        bool ok = true;
        for (int j = 1; j < 9 && ok; ++j) ok = (pvv[j] == pvv[0] + j);
        const int32_t cam = pvv[0] / 9;
        for (int k = 0; k < D.nf && ok; ++k) ok = (ctx->h_cam[pf[k]] == cam);
        if (!ok) continue;
        h_cam[b->n_cam++] = (int32_t)p;
        b->cam_nf_max = std::max(b->cam_nf_max, D.nf);
        fast[p] = 2;
      }
    }
    // point blocks: one warp per task, 32/G problems per warp; big classes first so that the
    // longest-running warps are scheduled first
    // How many blocks share a warp: the state-machine steps of the blocks of one warp serialise where they diverge
    // (1800 of the 3100 cycles of a pass with 16 two-observation blocks per warp), but fewer blocks per warp means
    // more warps, and only 8 warps per SM are resident (234 registers; occupancy query).  The smallest cap whose warp count
    // stays within 1.3 x the resident capacity is taken (measured: cap 8 for the 7776 blocks of ladybug on one GPU,
    // 4 / 2 / 2 for the 3888 / 1944 / 972 a rank holds at 2 / 4 / 8 GPUs); results do not depend on it.
    int tiles_cap = ctx->pt_tiles_cap;
    if (tiles_cap <= 0) {
      if (ctx->pt_warps_per_sm <= 0) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, solve_ba_points_kernel, 32, 0));
        ctx->pt_warps_per_sm = std::max(occ, 1);
      }
      const int64_t room = (int64_t)ctx->sm_count * ctx->pt_warps_per_sm * 13 / 10;
      tiles_cap = 32;
      for (int cap = 2; cap < 32; cap <<= 1) {
        int64_t warps = 0;
        for (int lg = 0; lg <= 5; ++lg) {
          const int64_t per = std::min(32 >> lg, cap);
          warps += ((int64_t)pt_lists[lg].size() + per - 1) / per;
        }
        if (warps <= room) {
          tiles_cap = cap;
          break;
        }
      }
    }
    int32_t npt = 0;
    for (int lg = 5; lg >= 0; --lg) {
      const int per_warp = std::min(32 >> lg, tiles_cap);
      const int32_t base = npt;
      const int32_t cnt = (int32_t)pt_lists[lg].size();
      if (cnt) std::memcpy(h_pt + npt, pt_lists[lg].data(), 4 * (size_t)cnt);
      npt += cnt;
      for (int32_t o = 0; o < cnt; o += per_warp) h_tasks[b->n_pt_warps++] = PointWarpTask{lg, base + o, std::min(per_warp, cnt - o), 0.0f};
    }
  }
  // NonlinearProductFactor components that fit one CTA's shared memory: the resident kernel (nlpf_resident.cuh)
  std::vector<int32_t> res_list, res_small_list;
  int64_t res_edges = 0;
  b->res_smem = b->res_small_smem = 0;
  if (ctx->kind == KIND_NLPF && !ctx->generic_only) {
    if (ctx->res_smem_cap < 0) {
      int optin = 0;
      cudaFuncAttributes fa;
      CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
      CK(cudaFuncGetAttributes(&fa, solve_nlpf_resident_kernel<kResThreads, 1>));
      ctx->res_smem_cap = std::max(0, optin - (int)fa.sharedSizeBytes);
      CK(cudaFuncSetAttribute(solve_nlpf_resident_kernel<kResThreads, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->res_smem_cap));
      CK(cudaFuncSetAttribute(solve_nlpf_resident_kernel<kResThreadsExact, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->res_smem_cap));
      CK(cudaFuncSetAttribute(solve_nlpf_resident_kernel<kResThreadsSmall, kResSmallCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, kResSmallSmem));
    }
    // pass 1 (host threads when the batch is large): per problem the flattened sizes — edges, distinct terms of its
    // own variables, edges on frozen variables — and whether it is a proper sibling set
    struct ResCount { int64_t nE, nT, nFz; bool ok; };
    std::vector<ResCount> cnt((size_t)nprobs);
    auto count_range = [&](int64_t p0, int64_t p1) {
      for (int64_t p = p0; p < p1; ++p) {
        const ProblemDesc& D = b->h_probs[p];
        ResCount c{0, 0, 0, true};
        if (D.nf <= kTileMax || D.nv > 0xffff) {
          c.ok = false;
          cnt[(size_t)p] = c;
          continue;
        }
        const int32_t* pvv = vids + D.var_off;
        const int32_t* pf = fids + D.fac_off;
        for (int j = 0; j < D.nv; ++j) c.nT += ctx->h_tvrow[pvv[j] + 1] - ctx->h_tvrow[pvv[j]];
        for (int k = 0; k < D.nf && c.ok; ++k) {
          for (int32_t e = ctx->h_rp32[pf[k]]; e < ctx->h_rp32[pf[k] + 1]; ++e) {
            const int32_t v = ctx->h_evid32[e];
            if (ctx->vmark[v] != epoch) ++c.nFz;                       // frozen variable: a constant term
            ++c.nE;                                                     // (foreign owners were rejected above)
          }
        }
        cnt[(size_t)p] = c;
      }
    };
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (tf >= 200000 && nprobs >= 2 * (int64_t)hw && hw > 1) {
      std::vector<std::thread> pool;
      const int64_t per = (nprobs + hw - 1) / hw;
      for (unsigned t = 0; t < hw; ++t) {
        const int64_t p0 = std::min<int64_t>(nprobs, (int64_t)t * per), p1 = std::min<int64_t>(nprobs, p0 + per);
        if (p0 < p1) pool.emplace_back(count_range, p0, p1);
      }
      for (std::thread& th : pool) th.join();
    } else {
      count_range(0, nprobs);
    }
    // pass 2: admission in problem order (scratch slices are handed out in that order)
    for (int64_t p = 0; p < nprobs; ++p) {
      ProblemDesc& D = b->h_probs[p];
      const ResCount& c = cnt[(size_t)p];
      const int64_t nE = c.nE, nT = c.nT, nFz = c.nFz;
      if (!c.ok || nT + nFz > 0xffff || nE > 0xffff || res_edges + nE > 0x7fffffffLL) continue;
      const ResLayout L = res_layout(D.nv, D.nf, (int)nE, (int)nT, (int)nFz);
      if (L.total > ctx->res_smem_cap) continue;
      D.nE = (int32_t)nE; D.nT = (int32_t)nT; D.nFz = (int32_t)nFz; D.goff = (int32_t)res_edges;
      res_edges += nE;
      h_probs[p] = D;
      if (ctx->res_threads != kResThreadsExact && D.nf <= kResSmallFactors && L.total <= kResSmallSmem) {
        b->res_small_smem = std::max(b->res_small_smem, L.total);
        res_small_list.push_back((int32_t)p);
      } else {
        b->res_smem = std::max(b->res_smem, L.total);
        res_list.push_back((int32_t)p);
      }
      fast[p] = 3;
    }
  }
  for (int64_t p = 0; p < nprobs; ++p) {
    if (fast[p]) continue;
    const int nf = b->h_probs[p].nf;
    if (nf <= kTileMax) {
      int g = next_pow2(std::max(nf, 1)), lg = 0;
      while ((1 << lg) < g) ++lg;
      tile_lists[lg].push_back((int32_t)p);
    } else if (nf <= kBlockMax) {
      const int t = nf <= 64 ? 0 : (nf <= 128 ? 1 : 2);
      block_lists[t].push_back((int32_t)p);
    } else {
      grid_list.push_back((int32_t)p);
    }
  }
  for (int lg = 0; lg < 6; ++lg) {
    if (tile_lists[lg].empty()) continue;
    b->classes.push_back({0, 1 << lg, (int64_t)b->h_order.size(), (int64_t)tile_lists[lg].size()});
    b->h_order.insert(b->h_order.end(), tile_lists[lg].begin(), tile_lists[lg].end());
  }
  for (int t = 0; t < 3; ++t) {
    if (block_lists[t].empty()) continue;
    b->classes.push_back({1, 64 << t, (int64_t)b->h_order.size(), (int64_t)block_lists[t].size()});
    b->h_order.insert(b->h_order.end(), block_lists[t].begin(), block_lists[t].end());
  }
  if (res_edges > 0) {
    CK(b->res_gscr.ensure((size_t)res_edges));
    CK(b->res_gvinc.ensure((size_t)res_edges));
  }
  if (!res_list.empty()) {
    b->classes.push_back({3, ctx->res_threads, (int64_t)b->h_order.size(), (int64_t)res_list.size()});
    b->h_order.insert(b->h_order.end(), res_list.begin(), res_list.end());
  }
  if (!res_small_list.empty()) {
    b->classes.push_back({4, kResThreadsSmall, (int64_t)b->h_order.size(), (int64_t)res_small_list.size()});
    b->h_order.insert(b->h_order.end(), res_small_list.begin(), res_small_list.end());
  }
  if (!grid_list.empty()) {
    b->classes.push_back({2, 256, (int64_t)b->h_order.size(), (int64_t)grid_list.size()});
    b->h_order.insert(b->h_order.end(), grid_list.begin(), grid_list.end());
  }
  if (!b->h_order.empty()) std::memcpy(h_order, b->h_order.data(), 4 * b->h_order.size());

  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(b->blob.p, hb, total, cudaMemcpyHostToDevice, s));
  char* db = b->blob.p;
  b->d_probs = reinterpret_cast<ProblemDesc*>(db + o_probs);
  b->d_vids = reinterpret_cast<int32_t*>(db + o_vids);
  b->d_fids = reinterpret_cast<int32_t*>(db + o_fids);
  b->d_order = reinterpret_cast<int32_t*>(db + o_order);
  b->d_pt_order = reinterpret_cast<int32_t*>(db + o_pt);
  b->d_cam_order = reinterpret_cast<int32_t*>(db + o_cam);
  b->d_pt_tasks = reinterpret_cast<PointWarpTask*>(db + o_tasks);
  b->d_pt_tasks_alt = reinterpret_cast<PointWarpTask*>(db + o_tasks2);
  CK(b->x0.ensure((size_t)tv));
  CK(b->xout.ensure((size_t)tv));
  CK(b->res.ensure((size_t)nprobs));
  CK(b->h_res.ensure((size_t)nprobs));
  CK(b->h_x.ensure((size_t)tv));
  // The pinned blob is only rewritten by the next batch_build on this object, which runs after
  // the caller has fetched (= synchronised) this batch's results, so no sync is needed here.
  return RDISGPU_OK;
}

int rdisgpu_batch_create(rdisgpu_ctx* ctx, const rdisgpu_problem* probs, int64_t nprobs, rdisgpu_batch** out) {
  if (!ctx || !out) return RDISGPU_ERR_ARG;
  *out = nullptr;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "batch_create before finalize");
  if (nprobs <= 0 || !probs) return ctx->fail(RDISGPU_ERR_ARG, "batch_create: no problems");
  if (nprobs >= 0x40000000LL) return ctx->fail(RDISGPU_ERR_ARG, "batch_create: too many problems");
  CK(cudaSetDevice(ctx->device));
  std::unique_ptr<rdisgpu_batch> b(new rdisgpu_batch());
  b->ctx = ctx;
  ProblemsView pv;
  pv.n = nprobs;
  pv.arr = probs;
  const int rc = batch_build(b.get(), pv);
  if (rc) return rc;
  *out = b.release();
  ++ctx->live_batches;
  return RDISGPU_OK;
}

int rdisgpu_batch_create_csr(rdisgpu_ctx* ctx, int64_t nprobs, const int64_t* var_off, const int32_t* vids,
                             const int64_t* fac_off, const int64_t* fids, rdisgpu_batch** out) {
  if (!ctx || !out) return RDISGPU_ERR_ARG;
  *out = nullptr;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "batch_create_csr before finalize");
  if (nprobs <= 0 || !var_off || !fac_off) return ctx->fail(RDISGPU_ERR_ARG, "batch_create_csr: no problems");
  if (nprobs >= 0x40000000LL) return ctx->fail(RDISGPU_ERR_ARG, "batch_create_csr: too many problems");
  CK(cudaSetDevice(ctx->device));
  std::unique_ptr<rdisgpu_batch> b(new rdisgpu_batch());
  b->ctx = ctx;
  ProblemsView pv;
  pv.n = nprobs;
  pv.var_off = var_off; pv.vids = vids; pv.fac_off = fac_off; pv.fids = fids;
  const int rc = batch_build(b.get(), pv);
  if (rc) return rc;
  *out = b.release();
  ++ctx->live_batches;
  return RDISGPU_OK;
}

int rdisgpu_batch_solve_cgd(rdisgpu_batch* b, const double* x0_host, int maxiters, double ftol) {
  if (!b) return RDISGPU_ERR_ARG;
  rdisgpu_ctx* ctx = b->ctx;
  if (maxiters <= 0) return ctx->fail(RDISGPU_ERR_ARG, "solve: maxiters must be positive");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const bool x0_on_device = is_device_ptr(x0_host);
  if (x0_host && !x0_on_device && b->total_nv > 0)  // pageable source: staged by the driver before the call returns; pinned: truly async
    CK(cudaMemcpyAsync(b->x0.p, x0_host, (size_t)b->total_nv * sizeof(double), cudaMemcpyHostToDevice, s));
  BatchView bv;
  bv.probs = b->d_probs;
  bv.vids = b->d_vids;
  bv.fids = b->d_fids;
  bv.x0 = x0_host ? (x0_on_device ? x0_host : b->x0.p) : nullptr;
  bv.xout = b->xout.p;
  bv.res = b->res.p;
  bv.gscr = b->res_gscr.p;
  bv.gvinc = b->res_gvinc.p;
  GraphView gv = ctx->gv;
  int launches = 0;
  if (ctx->strict) {  // the parity instrument: one kernel for every component shape, the reference's own operation order
    CK(b->strict_dscr.ensure((size_t)std::max<int64_t>(b->total_nv, 1)));
    if (ctx->kind == KIND_NLPF)
      solve_strict_kernel<NlpfOps><<<(unsigned)b->nprobs, kStrictThreads, 0, s>>>(gv, bv, b->strict_dscr.p, maxiters, ftol);
    else
      solve_strict_kernel<BaOps><<<(unsigned)b->nprobs, kStrictThreads, 0, s>>>(gv, bv, b->strict_dscr.p, maxiters, ftol);
    CK(cudaGetLastError());
    b->last_launches = 1;
    ctx->launches += 1;
    b->solved = true;
    return RDISGPU_OK;
  }
  if (b->n_pt_warps > 0) {
    solve_ba_points_kernel<<<b->n_pt_warps, 32, 0, s>>>(gv, bv, b->d_pt_order, b->d_pt_tasks, maxiters, ftol);
    ++launches;
    if (ctx->adaptive_order && b->n_pt_warps > 1 && b->n_pt_warps <= kReorderMax) {  // longest warps first at the next visit
      pt_reorder_kernel<<<(b->n_pt_warps + kReorderItems - 1) / kReorderItems, 4 * kReorderItems, 0, s>>>(b->d_pt_tasks, b->d_pt_tasks_alt,
                                                                                                       b->n_pt_warps);
      std::swap(b->d_pt_tasks, b->d_pt_tasks_alt);
      ++launches;
    }
    CK(cudaGetLastError());
  }
  if (b->n_cam > 0) {
    if (b->cam_C == 0) {
      // one observation per worker thread (C*T >= the longest factor list): the widest cluster — fewest observations per
      // SM, shortest evaluation — such that every camera block of the batch is resident at once; rdisgpu_set_option
      // "camera_cluster" pins C (experiments)
      for (int C = (ctx->cam_cluster_opt > 0 ? ctx->cam_cluster_opt : kCamMaxCluster); C >= 1 && b->cam_C == 0; --C) {
        int T = ((b->cam_nf_max + C - 1) / C + 31) / 32 * 32;
        T = std::max(T, 32);
        if (T > kCamMaxThreads) continue;  // cannot cover the list with this few CTAs (classification guarantees C = 8 can)
        if (C == 1 || ctx->cam_cluster_opt > 0) {
          b->cam_C = C;
          b->cam_T = T;
          break;
        }
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3((unsigned)(b->n_cam * C));
        qc.blockDim = dim3((unsigned)(T + 32));
        qc.stream = s;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = (unsigned)C;
        qa[0].val.clusterDim.y = 1;
        qa[0].val.clusterDim.z = 1;
        qc.attrs = qa;
        qc.numAttrs = 1;
        int nclusters = 0;
        cudaError_t qe = cudaOccupancyMaxActiveClusters(&nclusters, solve_ba_cameras_kernel, &qc);
        if (qe != cudaSuccess) {
          cudaGetLastError();
          continue;
        }
        // all clusters resident at once, give or take 3 % of them: GPC granularity leaves 48 slots for ladybug's 49
        // clusters of 8, and the straggler costs less than a narrower cluster for everybody (measured)
        if ((int64_t)nclusters * 32 >= (int64_t)b->n_cam * 31) {
          b->cam_C = C;
          b->cam_T = T;
        }
      }
      if (b->cam_C == 0) {  // more camera blocks than the chip holds clusters of any covering shape: the narrowest covering one, in waves
        int C = 1;
        while ((b->cam_nf_max + C - 1) / C > kCamMaxThreads) ++C;
        b->cam_C = C;
        b->cam_T = std::max(32, ((b->cam_nf_max + C - 1) / C + 31) / 32 * 32);
      }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b->n_cam * b->cam_C));
    cfg.blockDim = dim3((unsigned)(b->cam_T + 32));  // + the scalar warp
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)b->cam_C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int32_t* cord = b->d_cam_order;
    int Cc = b->cam_C;
    CK(cudaLaunchKernelEx(&cfg, solve_ba_cameras_kernel, gv, bv, cord, Cc, maxiters, ftol));
    ++launches;
    if (ctx->adaptive_order && b->n_cam > 1 && b->n_cam <= kReorderMax) {  // longest clusters first at the next visit
      cam_reorder_kernel<<<1, 1024, 0, s>>>(b->d_cam_order, b->n_cam, b->res.p);
      ++launches;
      CK(cudaGetLastError());
    }
  }
  for (const auto& c : b->classes) {
    const int32_t* ord = b->d_order + c.off;
    const int cnt = (int)c.count;
    if (c.kind == 0) {
      const int per_block = 128 / c.param;
      const int blocks = (cnt + per_block - 1) / per_block;
#define LAUNCH_TILE(OPS, GG) solve_tile_kernel<OPS, GG><<<blocks, 128, 0, s>>>(gv, bv, ord, cnt, maxiters, ftol)
#define LAUNCH_TILE_G(OPS)                 \
  switch (c.param) {                       \
    case 1: LAUNCH_TILE(OPS, 1); break;    \
    case 2: LAUNCH_TILE(OPS, 2); break;    \
    case 4: LAUNCH_TILE(OPS, 4); break;    \
    case 8: LAUNCH_TILE(OPS, 8); break;    \
    case 16: LAUNCH_TILE(OPS, 16); break;  \
    default: LAUNCH_TILE(OPS, 32); break;  \
  }
      if (ctx->kind == KIND_NLPF) { LAUNCH_TILE_G(NlpfOps) } else { LAUNCH_TILE_G(BaOps) }
#undef LAUNCH_TILE_G
#undef LAUNCH_TILE
      ++launches;
    } else if (c.kind == 1) {
      if (ctx->kind == KIND_NLPF)
        solve_block_kernel<NlpfOps><<<cnt, c.param, 0, s>>>(gv, bv, ord, cnt, maxiters, ftol);
      else
        solve_block_kernel<BaOps><<<cnt, c.param, 0, s>>>(gv, bv, ord, cnt, maxiters, ftol);
      ++launches;
    } else if (c.kind == 4) {
      solve_nlpf_resident_kernel<kResThreadsSmall, kResSmallCtas><<<cnt, kResThreadsSmall, (size_t)b->res_small_smem, s>>>(gv, bv, ord, cnt, maxiters, ftol);
      ++launches;
    } else if (c.kind == 3) {
      if (c.param == kResThreadsExact)
        solve_nlpf_resident_kernel<kResThreadsExact, 1><<<cnt, kResThreadsExact, (size_t)b->res_smem, s>>>(gv, bv, ord, cnt, maxiters, ftol);
      else
        solve_nlpf_resident_kernel<kResThreads, 1><<<cnt, kResThreads, (size_t)b->res_smem, s>>>(gv, bv, ord, cnt, maxiters, ftol);
      ++launches;
    } else {
      const int threads = 256;
      int occ = 0;
      const void* fn = (ctx->kind == KIND_NLPF) ? (const void*)solve_grid_kernel<NlpfOps> : (const void*)solve_grid_kernel<BaOps>;
      if (ctx->kind == KIND_NLPF)
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, solve_grid_kernel<NlpfOps>, threads, 0));
      else
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, solve_grid_kernel<BaOps>, threads, 0));
      if (occ < 1) return ctx->fail(RDISGPU_ERR_CUDA, "cooperative solve kernel cannot be resident");
      const int max_blocks = occ * ctx->sm_count;
      CK(ctx->grid_partials.ensure((size_t)max_blocks * 8));
      for (int i = 0; i < cnt; ++i) {
        int pidx = b->h_order[c.off + i];
        const ProblemDesc& D = b->h_probs[pidx];
        const int work = std::max(D.nf, D.nv);
        int blocks = std::min(max_blocks, std::max(1, (work + threads - 1) / threads));
        double* partials = ctx->grid_partials.p;
        void* args[] = {(void*)&gv, (void*)&bv, (void*)&pidx, (void*)&partials, (void*)&maxiters, (void*)&ftol};
        CK(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, 0, s));
        ++launches;
      }
    }
    CK(cudaGetLastError());
  }
  b->last_launches = launches;
  ctx->launches += launches;
  b->solved = true;
  return RDISGPU_OK;
}

int rdisgpu_batch_fetch(rdisgpu_batch* b, rdisgpu_result* out, double* sum_f_end) {
  if (!b) return RDISGPU_ERR_ARG;
  rdisgpu_ctx* ctx = b->ctx;
  if (!b->solved) return ctx->fail(RDISGPU_ERR_STATE, "batch_fetch before batch_solve");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(b->h_res.p, b->res.p, (size_t)b->nprobs * sizeof(ResultRec), cudaMemcpyDeviceToHost, s));
  bool want_x = false;
  if (out)
    for (int64_t p = 0; p < b->nprobs && !want_x; ++p) want_x = (out[p].x != nullptr);
  if (want_x && b->total_nv > 0)
    CK(cudaMemcpyAsync(b->h_x.p, b->xout.p, (size_t)b->total_nv * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double tot = 0.0;
  for (int64_t p = 0; p < b->nprobs; ++p) {
    const ResultRec& r = b->h_res.p[p];
    tot += r.f_end;
    if (!out) continue;
    out[p].f_init = r.f_init;
    out[p].f_end = r.f_end;
    out[p].iters = r.iters;
    out[p].status = r.status;
    out[p].n_feval = (int64_t)r.n_value + r.n_slope;
    out[p].n_geval = r.n_slope;
    if (out[p].x) std::memcpy(out[p].x, b->h_x.p + b->h_probs[p].var_off, (size_t)b->h_probs[p].nv * sizeof(double));
  }
  if (sum_f_end) *sum_f_end = tot;
  return RDISGPU_OK;
}

int rdisgpu_batch_fetch_csr(rdisgpu_batch* b, double* x_out, double* f_init, double* f_end, int32_t* iters, int32_t* status,
                            int64_t* n_feval, int64_t* n_geval) {
  if (!b) return RDISGPU_ERR_ARG;
  rdisgpu_ctx* ctx = b->ctx;
  if (!b->solved) return ctx->fail(RDISGPU_ERR_STATE, "batch_fetch_csr before batch_solve");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(b->h_res.p, b->res.p, (size_t)b->nprobs * sizeof(ResultRec), cudaMemcpyDeviceToHost, s));
  if (x_out && b->total_nv > 0)
    CK(cudaMemcpyAsync(b->h_x.p, b->xout.p, (size_t)b->total_nv * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (x_out && b->total_nv > 0) std::memcpy(x_out, b->h_x.p, (size_t)b->total_nv * sizeof(double));
  for (int64_t p = 0; p < b->nprobs; ++p) {
    const ResultRec& r = b->h_res.p[p];
    if (f_init) f_init[p] = r.f_init;
    if (f_end) f_end[p] = r.f_end;
    if (iters) iters[p] = r.iters;
    if (status) status[p] = r.status;
    if (n_feval) n_feval[p] = (int64_t)r.n_value + r.n_slope;
    if (n_geval) n_geval[p] = r.n_slope;
  }
  return RDISGPU_OK;
}

int rdisgpu_batch_objective_device(rdisgpu_batch* b, double* sum_dev) {
  if (!b) return RDISGPU_ERR_ARG;
  rdisgpu_ctx* ctx = b->ctx;
  if (!b->solved) return ctx->fail(RDISGPU_ERR_STATE, "batch_objective_device before batch_solve");
  if (!is_device_ptr(sum_dev)) return ctx->fail(RDISGPU_ERR_ARG, "batch_objective_device: sum_dev must be device memory");
  CK(cudaSetDevice(ctx->device));
  sum_f_end_kernel<<<1, 256, 0, ctx->stream>>>(b->res.p, b->nprobs, sum_dev);
  ++ctx->launches;
  CK(cudaGetLastError());
  return RDISGPU_OK;
}

void rdisgpu_batch_destroy(rdisgpu_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  --b->ctx->live_batches;
  delete b;
}

int rdisgpu_batch_last_launches(const rdisgpu_batch* b) { return b ? b->last_launches : 0; }

static rdisgpu_batch* scratch_batch(rdisgpu_ctx* ctx) {
  if (!ctx->scratch_batch) {
    ctx->scratch_batch = new rdisgpu_batch();
    ctx->scratch_batch->ctx = ctx;
  }
  return ctx->scratch_batch;
}

int rdisgpu_solve_cgd(rdisgpu_ctx* ctx, const rdisgpu_problem* probs, int64_t nprobs, int maxiters, double ftol,
                      rdisgpu_result* out) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!out) return ctx->fail(RDISGPU_ERR_ARG, "solve_cgd: null result array");
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "solve_cgd before finalize");
  if (nprobs <= 0 || !probs) return ctx->fail(RDISGPU_ERR_ARG, "solve_cgd: no problems");
  if (nprobs >= 0x40000000LL) return ctx->fail(RDISGPU_ERR_ARG, "solve_cgd: too many problems");
  CK(cudaSetDevice(ctx->device));
  rdisgpu_batch* b = scratch_batch(ctx);
  ProblemsView pv;
  pv.n = nprobs;
  pv.arr = probs;
  int rc = batch_build(b, pv);
  if (rc) return rc;
  // x0: either every problem brings start values or none does
  bool any = false, all = true;
  for (int64_t p = 0; p < nprobs; ++p) {
    if (probs[p].x0) any = true;
    else if (probs[p].nv > 0) all = false;
  }
  double* x0 = b->h_x.p;  // pinned staging (re-used for the download after the solve)
  if (any && !all) {
    // mixed: fill the gaps from the device state
    for (int64_t p = 0; p < nprobs && rc == RDISGPU_OK; ++p) {
      double* dst = x0 + b->h_probs[p].var_off;
      if (probs[p].x0) std::memcpy(dst, probs[p].x0, (size_t)probs[p].nv * sizeof(double));
      else rc = rdisgpu_get_x(ctx, probs[p].nv, probs[p].vid, dst);
    }
  } else if (any) {
    for (int64_t p = 0; p < nprobs; ++p)
      if (probs[p].nv > 0) std::memcpy(x0 + b->h_probs[p].var_off, probs[p].x0, (size_t)probs[p].nv * sizeof(double));
  }
  if (rc == RDISGPU_OK) rc = rdisgpu_batch_solve_cgd(b, any ? x0 : nullptr, maxiters, ftol);
  if (rc == RDISGPU_OK) rc = rdisgpu_batch_fetch(b, out, nullptr);
  return rc;
}

int rdisgpu_solve_cgd_csr(rdisgpu_ctx* ctx, int64_t nprobs, const int64_t* var_off, const int32_t* vids,
                          const int64_t* fac_off, const int64_t* fids, const double* x0, int maxiters, double ftol,
                          double* x_out, double* f_init, double* f_end, int32_t* iters, int32_t* status,
                          int64_t* n_feval, int64_t* n_geval) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "solve_cgd_csr before finalize");
  if (nprobs <= 0 || !var_off || !fac_off) return ctx->fail(RDISGPU_ERR_ARG, "solve_cgd_csr: no problems");
  if (nprobs >= 0x40000000LL) return ctx->fail(RDISGPU_ERR_ARG, "solve_cgd_csr: too many problems");
  CK(cudaSetDevice(ctx->device));
  rdisgpu_batch* b = scratch_batch(ctx);
  ProblemsView pv;
  pv.n = nprobs;
  pv.var_off = var_off; pv.vids = vids; pv.fac_off = fac_off; pv.fids = fids;
  int rc = batch_build(b, pv);
  if (rc) return rc;
  rc = rdisgpu_batch_solve_cgd(b, x0, maxiters, ftol);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(b->h_res.p, b->res.p, (size_t)nprobs * sizeof(ResultRec), cudaMemcpyDeviceToHost, s));
  if (x_out && b->total_nv > 0)
    CK(cudaMemcpyAsync(b->h_x.p, b->xout.p, (size_t)b->total_nv * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (x_out && b->total_nv > 0) std::memcpy(x_out, b->h_x.p, (size_t)b->total_nv * sizeof(double));
  for (int64_t p = 0; p < nprobs; ++p) {
    const ResultRec& r = b->h_res.p[p];
    if (f_init) f_init[p] = r.f_init;
    if (f_end) f_end[p] = r.f_end;
    if (iters) iters[p] = r.iters;
    if (status) status[p] = r.status;
    if (n_feval) n_feval[p] = (int64_t)r.n_value + r.n_slope;
    if (n_geval) n_geval[p] = r.n_slope;
  }
  return RDISGPU_OK;
}

// LMSubspaceOptimizer::optimize for ONE component of more than kLmMaxVars variables (lm_dense.cuh): levmar's dlevmar_der
// control flow on the host, everything else on the device.  x0_dev: start values (device, m doubles) or null = device state.
static int solve_lm_dense(rdisgpu_ctx* ctx, rdisgpu_batch* b, int64_t pidx, const double* x0_dev, int itmax, double tau, double eps1,
                          double eps2, double eps3, ResultRec* res_host, double* xout_dev) {
  cudaStream_t s = ctx->stream;
  const ProblemDesc& D = b->h_probs[(size_t)pidx];
  const int m = D.nv, nf = D.nf;
  const int32_t* hv = reinterpret_cast<const int32_t*>(b->h_blob.p + ((const char*)b->d_vids - b->blob.p)) + D.var_off;
  const int32_t* hf = reinterpret_cast<const int32_t*>(b->h_blob.p + ((const char*)b->d_fids - b->blob.p)) + D.fac_off;
  // ---- the component's row pointers and variable-major incidence (host, once per solve) ----
  std::vector<int32_t> loc((size_t)ctx->V, -1);
  for (int i = 0; i < m; ++i) loc[(size_t)hv[i]] = i;
  std::vector<int32_t> jptr((size_t)nf + 1, 0);
  auto arity = [&](int32_t f) { return ctx->kind == KIND_BA ? 12 : ctx->h_rp32[(size_t)f + 1] - ctx->h_rp32[(size_t)f]; };
  auto slot_vid = [&](int32_t f, int sl) -> int32_t {
    if (ctx->kind == KIND_BA) return sl < 9 ? 9 * ctx->h_cam[(size_t)f] + sl : 9 * ctx->ncams + 3 * ctx->h_pt[(size_t)f] + (sl - 9);
    return ctx->h_evid32[(size_t)ctx->h_rp32[(size_t)f] + sl];
  };
  for (int k = 0; k < nf; ++k) jptr[(size_t)k + 1] = jptr[(size_t)k] + arity(hf[k]);
  const int64_t nent = jptr[(size_t)nf];
  std::vector<int32_t> voff((size_t)m + 1, 0);
  for (int k = 0; k < nf; ++k)
    for (int sl = 0; sl < arity(hf[k]); ++sl) {
      const int32_t li = loc[(size_t)slot_vid(hf[k], sl)];
      if (li >= 0) ++voff[(size_t)li + 1];
    }
  for (int i = 0; i < m; ++i) voff[(size_t)i + 1] += voff[(size_t)i];
  std::vector<int32_t> ient((size_t)voff[(size_t)m]), irow((size_t)voff[(size_t)m]), cur(voff.begin(), voff.end() - 1);
  for (int k = 0; k < nf; ++k)
    for (int sl = 0; sl < arity(hf[k]); ++sl) {
      const int32_t li = loc[(size_t)slot_vid(hf[k], sl)];
      if (li < 0) continue;
      ient[(size_t)cur[(size_t)li]] = jptr[(size_t)k] + sl;
      irow[(size_t)cur[(size_t)li]++] = k;
    }
  // ---- device buffers ----
  const size_t mm = (size_t)m * m;
  CK(ctx->lmd_A.ensure(mm));
  CK(ctx->lmd_L.ensure(mm));
  CK(ctx->lmd_jval.ensure((size_t)std::max<int64_t>(nent, 1)));
  CK(ctx->lmd_e.ensure((size_t)nf));
  CK(ctx->lmd_hx.ensure((size_t)nf));
  CK(ctx->lmd_vec.ensure((size_t)m * 7));   // g | p | pDp | dp | y | diag | spare
  CK(ctx->lmd_scal.ensure(4096 + kLmNB * kLmNB));
  CK(ctx->lmd_flag.ensure(1));
  const size_t nidx = (size_t)nf + 1 + (size_t)nent + (size_t)m + 1 + 2 * ient.size() + 8;
  CK(ctx->lmd_idx.ensure(nidx));
  int32_t* d_jptr = ctx->lmd_idx.p;
  int32_t* d_jcol = d_jptr + nf + 1;
  int32_t* d_voff = d_jcol + nent;
  int32_t* d_ient = d_voff + m + 1;
  int32_t* d_irow = d_ient + ient.size();
  CK(cudaMemcpyAsync(d_jptr, jptr.data(), jptr.size() * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_voff, voff.data(), voff.size() * 4, cudaMemcpyHostToDevice, s));
  if (!ient.empty()) {
    CK(cudaMemcpyAsync(d_ient, ient.data(), ient.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_irow, irow.data(), irow.size() * 4, cudaMemcpyHostToDevice, s));
  }
  CK(cudaMemcpyAsync(ctx->lm_vloc.p, loc.data(), loc.size() * 4, cudaMemcpyHostToDevice, s));  // whole map: restored below
  CK(cudaStreamSynchronize(s));  // the staging vectors above may now die
  double* g = ctx->lmd_vec.p;
  double* p = g + m;
  double* pDp = p + m;
  double* dp = pDp + m;
  double* y = dp + m;
  double* diag = y + m;
  double* scal = ctx->lmd_scal.p;
  double* linv = scal + 4096;  // inverse of the current diagonal block
  LmDenseView L;
  L.m = m; L.nf = nf;
  L.fids = b->d_fids + D.fac_off;
  L.vids = b->d_vids + D.var_off;
  L.vloc = ctx->lm_vloc.p;
  L.jptr = d_jptr; L.jcol = d_jcol; L.jval = ctx->lmd_jval.p;
  L.e = ctx->lmd_e.p; L.hx = ctx->lmd_hx.p;
  L.voff = d_voff; L.ient = d_ient; L.irow = d_irow;
  const bool nlpf = (ctx->kind == KIND_NLPF);
  const int fblocks = std::min(512, (nf + 255) / 256);
  int launches = 0;
  if (x0_dev) CK(cudaMemcpyAsync(p, x0_dev, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, s));
  else {
    gather_x_kernel<<<(m + 255) / 256, 256, 0, s>>>(ctx->gv, m, L.vids, p);
    ++launches;
  }
  // func(q): assign, residuals; returns {sum f, sum hx^2}
  auto func = [&](const double* q, double& sumf, double& e2) -> int {
    lm_assign_kernel<<<(m + 255) / 256, 256, 0, s>>>(ctx->gv, L.vids, m, q);
    if (nlpf) lm_func_kernel<NlpfOps><<<fblocks, 256, 0, s>>>(ctx->gv, L, scal + 8);
    else lm_func_kernel<BaOps><<<fblocks, 256, 0, s>>>(ctx->gv, L, scal + 8);
    launches += 2;
    std::vector<double> part((size_t)fblocks * 2);
    CK(cudaMemcpyAsync(part.data(), scal + 8, part.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    sumf = 0.0; e2 = 0.0;
    for (int i = 0; i < fblocks; ++i) { sumf += part[2 * (size_t)i]; e2 += part[2 * (size_t)i + 1]; }
    return RDISGPU_OK;
  };
  auto finite = [](double v) { return std::fabs(v) <= 1.7976931348623157e308; };
  double ival = 0, p_eL2 = 0;
  int rc = func(p, ival, p_eL2);
  if (rc) return rc;
  lm_negate_kernel<<<(nf + 255) / 256, 256, 0, s>>>(L.hx, L.e, nf);
  ++launches;
  double mu = 0.0;
  int nu = 2, stop = 0, nfev = 1, njev = 0, k_it = 0;
  if (!finite(p_eL2)) stop = 7;
  for (k_it = 0; k_it < itmax && !stop; ++k_it) {
    if (p_eL2 <= eps3) { stop = 6; break; }
    if (nlpf) lm_rows_kernel<NlpfOps><<<std::min(1024, (nf + 127) / 128), 128, 0, s>>>(ctx->gv, L);
    else lm_rows_kernel<BaOps><<<std::min(1024, (nf + 127) / 128), 128, 0, s>>>(ctx->gv, L);
    ++njev;
    CK(cudaMemsetAsync(ctx->lmd_A.p, 0, mm * sizeof(double), s));
    lm_assemble_kernel<<<std::min(2048, (m + 3) / 4), 128, 0, s>>>(L, ctx->lmd_A.p, g);
    lm_scalars_kernel<<<1, 256, 0, s>>>(ctx->lmd_A.p, g, p, m, diag, scal);
    launches += 3;
    double h3[3];
    CK(cudaMemcpyAsync(h3, scal, sizeof h3, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const double jacTe_inf = h3[0], p_L2 = h3[1];
    if (jacTe_inf <= eps1) { stop = 1; break; }
    if (k_it == 0) mu = tau * h3[2];
    while (true) {  // adaptive damping
      lm_augment_kernel<<<std::min<int64_t>(4096, (int64_t)((mm + 255) / 256)), 256, 0, s>>>(ctx->lmd_A.p, ctx->lmd_L.p, m, mu);
      CK(cudaMemsetAsync(ctx->lmd_flag.p, 0, sizeof(int), s));
      ++launches;
      for (int kb = 0; kb < m; kb += kLmNB) {
        const int nb = std::min(kLmNB, m - kb), rest = m - kb - nb;
        lm_potrf_kernel<<<1, 256, 0, s>>>(ctx->lmd_L.p, m, kb, linv, ctx->lmd_flag.p);
        ++launches;
        if (rest > 0) {
          const int nt = (rest + kLmNB - 1) / kLmNB;
          lm_tile_kernel<kPanelSolve><<<dim3((unsigned)nt, 1), 128, 0, s>>>(ctx->lmd_L.p, m, kb, nb, linv);
          lm_tile_kernel<kTrailing><<<dim3((unsigned)nt, (unsigned)nt), 128, 0, s>>>(ctx->lmd_L.p, m, kb, nb, linv);
          launches += 2;
        }
      }
      lm_trsv_kernel<<<1, 1024, 0, s>>>(ctx->lmd_L.p, m, g, p, mu, y, dp, pDp, scal + 4);
      ++launches;
      int hflag = 0;
      double h2[2];
      CK(cudaMemcpyAsync(&hflag, ctx->lmd_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(h2, scal + 4, sizeof h2, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (hflag == 0 && finite(h2[0])) {
        const double Dp_L2 = h2[0], dL = h2[1];
        if (Dp_L2 <= eps2 * eps2 * p_L2) { stop = 2; break; }
        if (Dp_L2 >= (p_L2 + eps2) / (1e-12 * 1e-12)) { stop = 4; break; }
        double sumf, pDp_eL2;
        rc = func(pDp, sumf, pDp_eL2);
        if (rc) return rc;
        ++nfev;
        if (!finite(pDp_eL2)) { stop = 7; break; }
        const double dF = p_eL2 - pDp_eL2;
        if (dL > 0.0 && dF > 0.0) {
          double t = (2.0 * dF / dL - 1.0);
          t = 1.0 - t * t * t;
          mu = mu * ((t >= 0.3333333334) ? t : 0.3333333334);
          nu = 2;
          CK(cudaMemcpyAsync(p, pDp, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, s));
          lm_negate_kernel<<<(nf + 255) / 256, 256, 0, s>>>(L.hx, L.e, nf);
          ++launches;
          p_eL2 = pDp_eL2;
          break;
        }
      }
      mu *= nu;
      const int nu2 = nu << 1;
      if (nu2 <= nu) { stop = 5; break; }
      nu = nu2;
    }
  }
  if (k_it >= itmax) stop = 3;
  double fval, e2;
  rc = func(p, fval, e2);  // commit: quickAssignVals(xval) without clamping, fval = evalFactors (:102-110)
  if (rc) return rc;
  CK(cudaMemcpyAsync(xout_dev, p, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(ctx->lm_vloc.p, 0xff, (size_t)ctx->V * sizeof(int32_t), s));  // -1 everywhere again
  CK(cudaStreamSynchronize(s));
  res_host->f_init = ival;
  res_host->f_end = fval;
  res_host->iters = k_it;
  res_host->status = stop;
  res_host->n_value = nfev;
  res_host->n_slope = njev;
  ctx->launches += launches;
  b->last_launches += launches;
  return RDISGPU_OK;
}

int rdisgpu_solve_lm_csr(rdisgpu_ctx* ctx, int64_t nprobs, const int64_t* var_off, const int32_t* vids, const int64_t* fac_off,
                         const int64_t* fids, const double* x0, int maxiters, const double* opts4, double* x_out,
                         double* f_init, double* f_end, int32_t* iters, int32_t* stop, int64_t* n_feval, int64_t* n_jeval) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "solve_lm_csr before finalize");
  if (nprobs <= 0 || !var_off || !fac_off) return ctx->fail(RDISGPU_ERR_ARG, "solve_lm_csr: no problems");
  if (nprobs >= 0x40000000LL) return ctx->fail(RDISGPU_ERR_ARG, "solve_lm_csr: too many problems");
  if (maxiters <= 0) return ctx->fail(RDISGPU_ERR_ARG, "solve_lm_csr: maxiters must be positive");
  CK(cudaSetDevice(ctx->device));
  rdisgpu_batch* b = scratch_batch(ctx);
  ProblemsView pv;
  pv.n = nprobs;
  pv.var_off = var_off; pv.vids = vids; pv.fac_off = fac_off; pv.fids = fids;
  int rc = batch_build(b, pv);
  if (rc) return rc;
  // dense-Jacobian scratch: per problem jac[n*m] | e[n] | hx[n], n = max(|factors|, m) (LMSubspaceOptimizer.cpp:48-49)
  std::vector<int64_t> off((size_t)nprobs + 1, 0);
  for (int64_t p = 0; p < nprobs; ++p) {
    const int64_t m = b->h_probs[p].nv, nf = b->h_probs[p].nf;
    const int64_t n = std::max(nf, m);
    off[(size_t)p + 1] = off[(size_t)p] + ((m > kLmMaxVars) ? 0 : n * m + 2 * n);  // larger components: lm_dense.cuh, own buffers
  }
  cudaStream_t s = ctx->stream;
  CK(ctx->lm_off.ensure(off.size()));
  CK(ctx->lm_scratch.ensure((size_t)std::max<int64_t>(off.back(), 1)));
  CK(cudaMemcpyAsync(ctx->lm_off.p, off.data(), off.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  if (!ctx->lm_vloc_ready) {
    CK(ctx->lm_vloc.ensure((size_t)ctx->V));
    CK(cudaMemsetAsync(ctx->lm_vloc.p, 0xff, (size_t)ctx->V * sizeof(int32_t), s));  // -1
    ctx->lm_vloc_ready = true;
  }
  const bool x0_on_device = is_device_ptr(x0);
  if (x0 && !x0_on_device && b->total_nv > 0)
    CK(cudaMemcpyAsync(b->x0.p, x0, (size_t)b->total_nv * sizeof(double), cudaMemcpyHostToDevice, s));
  BatchView bv;
  bv.probs = b->d_probs;
  bv.vids = b->d_vids;
  bv.fids = b->d_fids;
  bv.x0 = x0 ? (x0_on_device ? x0 : b->x0.p) : nullptr;
  bv.xout = b->xout.p;
  bv.res = b->res.p;
  bv.gscr = nullptr;
  bv.gvinc = nullptr;
  LmView lv;
  lv.vloc = ctx->lm_vloc.p;
  lv.scratch = ctx->lm_scratch.p;
  lv.scr_off = ctx->lm_off.p;
  const double tau = opts4 ? opts4[0] : 1e-3, eps1 = opts4 ? opts4[1] : 1e-15, eps2 = opts4 ? opts4[2] : 1e-15,
               eps3 = opts4 ? opts4[3] : 3e-8;
  // bundle-adjustment point blocks run register-resident (one lane per observation); camera blocks and
  // every other shape take the generic one-CTA-per-component kernel
  int launches = 0;
  if (b->n_pt_warps > 0) {
    solve_lm_ba_points_kernel<<<b->n_pt_warps, 32, 0, s>>>(ctx->gv, bv, b->d_pt_order, b->d_pt_tasks, maxiters, tau, eps1, eps2, eps3);
    ++launches;
    CK(cudaGetLastError());
  }
  auto launch_generic = [&](const int32_t* order, int count) -> cudaError_t {
    if (count <= 0) return cudaSuccess;
    if (ctx->kind == KIND_NLPF)
      solve_lm_block_kernel<NlpfOps><<<(unsigned)count, kLmThreads, 0, s>>>(ctx->gv, bv, lv, order, maxiters, tau, eps1, eps2, eps3);
    else
      solve_lm_block_kernel<BaOps><<<(unsigned)count, kLmThreads, 0, s>>>(ctx->gv, bv, lv, order, maxiters, tau, eps1, eps2, eps3);
    ++launches;
    return cudaGetLastError();
  };
  CK(launch_generic(b->d_cam_order, b->n_cam));
  // generic shapes: up to kLmMaxVars variables -> one CTA per component; larger -> the dense solver, one at a time
  std::vector<int32_t> small_list, big_list;
  for (int32_t pi : b->h_order) (b->h_probs[(size_t)pi].nv > kLmMaxVars ? big_list : small_list).push_back(pi);
  if (!small_list.empty()) {
    CK(ctx->s_i32b.ensure(small_list.size()));
    CK(cudaMemcpyAsync(ctx->s_i32b.p, small_list.data(), small_list.size() * 4, cudaMemcpyHostToDevice, s));
    CK(launch_generic(ctx->s_i32b.p, (int)small_list.size()));
    CK(cudaStreamSynchronize(s));  // small_list is a local
  }
  ctx->launches += launches;
  b->last_launches = launches;
  std::vector<std::pair<int32_t, ResultRec>> big_res;
  for (int32_t pi : big_list) {
    ResultRec r;
    const ProblemDesc& D = b->h_probs[(size_t)pi];
    rc = solve_lm_dense(ctx, b, pi, bv.x0 ? bv.x0 + D.var_off : nullptr, maxiters, tau, eps1, eps2, eps3, &r, b->xout.p + D.var_off);
    if (rc) return rc;
    big_res.push_back({pi, r});
  }
  CK(cudaMemcpyAsync(b->h_res.p, b->res.p, (size_t)nprobs * sizeof(ResultRec), cudaMemcpyDeviceToHost, s));
  if (x_out && b->total_nv > 0)
    CK(cudaMemcpyAsync(b->h_x.p, b->xout.p, (size_t)b->total_nv * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));  // also keeps `off` alive until the copy is done
  for (const auto& br : big_res) b->h_res.p[br.first] = br.second;
  if (x_out && b->total_nv > 0) std::memcpy(x_out, b->h_x.p, (size_t)b->total_nv * sizeof(double));
  for (int64_t p = 0; p < nprobs; ++p) {
    const ResultRec& r = b->h_res.p[p];
    if (f_init) f_init[p] = r.f_init;
    if (f_end) f_end[p] = r.f_end;
    if (iters) iters[p] = r.iters;
    if (stop) stop[p] = r.status;
    if (n_feval) n_feval[p] = r.n_value;
    if (n_jeval) n_jeval[p] = r.n_slope;
  }
  return RDISGPU_OK;
}

int rdisgpu_components(rdisgpu_ctx* ctx, const uint8_t* assigned, int32_t* var_label, int32_t* fac_label, int32_t* n_components,
                       int32_t* n_rounds) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "components before finalize");
  if (!assigned || !var_label) return ctx->fail(RDISGPU_ERR_ARG, "components: null argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int64_t V = ctx->V, F = ctx->F;
  CK(ctx->cc_assigned.ensure((size_t)V));
  CK(ctx->cc_vlabel.ensure((size_t)V));
  CK(ctx->cc_flabel.ensure((size_t)F));
  CK(cudaMemcpyAsync(ctx->cc_assigned.p, assigned, (size_t)V, cudaMemcpyHostToDevice, s));
  ComponentsView cv;
  cv.assigned = ctx->cc_assigned.p;
  cv.vlabel = ctx->cc_vlabel.p;
  cv.flabel = ctx->cc_flabel.p;
  cv.changed = ctx->cc_flag.p;
  const int threads = 256;
  const int vb = (int)std::min<int64_t>((V + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
  const int fb = (int)std::min<int64_t>((F + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
  cc_init_kernel<<<vb, threads, 0, s>>>(ctx->gv, cv);
  ++ctx->launches;
  // Rounds are enqueued kRoundsPerCheck at a time, each with its own change flag; the host looks at the flags once per
  // group (one copy + one synchronisation instead of one per round): a round that changed nothing is a fixed point, and
  // the rounds enqueued after it are no-ops.  Long chains (the 1000-variable sinusoid chain) need tens of rounds.
  constexpr int kRoundsPerCheck = 8;
  CK(ctx->cc_flag.ensure(kRoundsPerCheck));
  int rounds = 0;
  for (bool converged = false; !converged;) {
    if (rounds > (1 << 20)) return ctx->fail(RDISGPU_ERR_CUDA, "components: label propagation did not converge");
    CK(cudaMemsetAsync(ctx->cc_flag.p, 0, kRoundsPerCheck * sizeof(int32_t), s));
    for (int r = 0; r < kRoundsPerCheck; ++r) {
      cv.changed = ctx->cc_flag.p + r;
      cc_hook_kernel<<<fb, threads, 0, s>>>(ctx->gv, cv);
      cc_jump_kernel<<<vb, threads, 0, s>>>(ctx->gv, cv);
    }
    ctx->launches += 2 * kRoundsPerCheck;
    CK(cudaGetLastError());
    int32_t changed[kRoundsPerCheck];
    CK(cudaMemcpyAsync(changed, ctx->cc_flag.p, sizeof changed, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int r = 0; r < kRoundsPerCheck; ++r) {
      if (!changed[r]) {
        converged = true;
        break;
      }
      ++rounds;
    }
  }
  cc_factor_labels_kernel<<<fb, threads, 0, s>>>(ctx->gv, cv);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(var_label, ctx->cc_vlabel.p, (size_t)V * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (fac_label) CK(cudaMemcpyAsync(fac_label, ctx->cc_flabel.p, (size_t)F * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (n_components) {
    int32_t n = 0;
    for (int64_t v = 0; v < V; ++v) n += (var_label[v] == (int32_t)v);
    *n_components = n;
  }
  if (n_rounds) *n_rounds = rounds + 1;
  return RDISGPU_OK;
}

int rdisgpu_bounds(rdisgpu_ctx* ctx, const uint8_t* assigned, int64_t nf, const int64_t* fid, double* lower, double* upper,
                   double sum[2]) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "bounds before finalize");
  if (!assigned) return ctx->fail(RDISGPU_ERR_ARG, "bounds: null argument");
  if (!fid) nf = ctx->F;
  if (nf < 0) return ctx->fail(RDISGPU_ERR_ARG, "bounds: bad argument");
  if (sum) sum[0] = sum[1] = 0.0;  // semiring Product identity (MinSum: +, 0), src/OptimizableFunction.cpp:192
  if (nf == 0) return RDISGPU_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (fid) {
    int rc = upload_fids(ctx, nf, fid, ctx->s_i32a);
    if (rc) return rc;
  }
  CK(ctx->cc_assigned.ensure((size_t)ctx->V));
  CK(cudaMemcpyAsync(ctx->cc_assigned.p, assigned, (size_t)ctx->V, cudaMemcpyHostToDevice, s));
  CK(ctx->s_f64a.ensure((size_t)nf));
  CK(ctx->s_f64b.ensure((size_t)nf));
  const int threads = 128;
  const int blocks = (int)std::min<int64_t>((nf + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
  if (ctx->kind == KIND_NLPF)
    factor_bounds_kernel<NlpfOps><<<blocks, threads, 0, s>>>(ctx->gv, ctx->cc_assigned.p, fid ? ctx->s_i32a.p : nullptr, nf, ctx->s_f64a.p, ctx->s_f64b.p);
  else
    factor_bounds_kernel<BaOps><<<blocks, threads, 0, s>>>(ctx->gv, ctx->cc_assigned.p, fid ? ctx->s_i32a.p : nullptr, nf, ctx->s_f64a.p, ctx->s_f64b.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  std::vector<double> lo_tmp, hi_tmp;
  double* lo = lower;
  double* hi = upper;
  if (!lo) { lo_tmp.resize((size_t)nf); lo = lo_tmp.data(); }
  if (!hi) { hi_tmp.resize((size_t)nf); hi = hi_tmp.data(); }
  CK(cudaMemcpyAsync(lo, ctx->s_f64a.p, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(hi, ctx->s_f64b.p, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (sum) {  // bounds = Product(bounds, fb), factor by factor in list order (OptimizableFunction.cpp:194-211)
    double a = 0.0, b = 0.0;
    for (int64_t k = 0; k < nf; ++k) {
      a += lo[k];
      b += hi[k];
    }
    sum[0] = a;
    sum[1] = b;
  }
  return RDISGPU_OK;
}

int rdisgpu_bounds_lists(rdisgpu_ctx* ctx, const uint8_t* assigned, int64_t nlists, const int64_t* list_off, const int64_t* fid,
                         double* sums) {
  if (!ctx) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "bounds_lists before finalize");
  if (!assigned || nlists < 0 || (nlists > 0 && (!list_off || !sums))) return ctx->fail(RDISGPU_ERR_ARG, "bounds_lists: bad argument");
  if (nlists == 0) return RDISGPU_OK;
  const int64_t nf = list_off[nlists];
  if (list_off[0] != 0 || nf < 0 || (nf > 0 && !fid)) return ctx->fail(RDISGPU_ERR_ARG, "bounds_lists: malformed offsets");
  for (int64_t l = 0; l < nlists; ++l)
    if (list_off[l + 1] < list_off[l]) return ctx->fail(RDISGPU_ERR_ARG, "bounds_lists: offsets must be non-decreasing");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (nf > 0) {
    int rc = upload_fids(ctx, nf, fid, ctx->s_i32a);
    if (rc) return rc;
  }
  CK(ctx->cc_assigned.ensure((size_t)ctx->V));
  CK(cudaMemcpyAsync(ctx->cc_assigned.p, assigned, (size_t)ctx->V, cudaMemcpyHostToDevice, s));
  CK(ctx->lm_off.ensure((size_t)nlists + 1));
  CK(cudaMemcpyAsync(ctx->lm_off.p, list_off, (size_t)(nlists + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  CK(ctx->s_f64a.ensure((size_t)nlists * 2));
  const int threads = 128;
  const int blocks = (int)std::min<int64_t>((nlists * 32 + threads - 1) / threads, (int64_t)ctx->sm_count * 16);
  if (ctx->kind == KIND_NLPF)
    list_bounds_kernel<NlpfOps><<<blocks, threads, 0, s>>>(ctx->gv, ctx->cc_assigned.p, ctx->lm_off.p, ctx->s_i32a.p, nlists, ctx->s_f64a.p);
  else
    list_bounds_kernel<BaOps><<<blocks, threads, 0, s>>>(ctx->gv, ctx->cc_assigned.p, ctx->lm_off.p, ctx->s_i32a.p, nlists, ctx->s_f64a.p);
  ++ctx->launches;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(sums, ctx->s_f64a.p, (size_t)nlists * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return RDISGPU_OK;
}

int rdisgpu_batch_info(const rdisgpu_batch* b, int32_t out[8]) {
  if (!b || !out) return RDISGPU_ERR_ARG;
  int n_generic = 0;
  for (const auto& c : b->classes)
    if (c.kind != 3 && c.kind != 4) n_generic += (int)c.count;
  out[0] = (int32_t)b->nprobs;
  out[1] = b->n_pt_warps;
  out[2] = b->n_cam;
  out[3] = b->cam_C;
  out[4] = b->cam_T;
  out[5] = n_generic;
  out[6] = b->cam_nf_max;
  out[7] = b->last_launches;
  return RDISGPU_OK;
}

int rdisgpu_batch_resident_info(const rdisgpu_batch* b, int32_t out[2]) {
  if (!b || !out) return RDISGPU_ERR_ARG;
  out[0] = 0;
  for (const auto& c : b->classes)
    if (c.kind == 3 || c.kind == 4) out[0] += (int32_t)c.count;
  out[1] = std::max(b->res_smem, b->res_small_smem);
  return RDISGPU_OK;
}

// ==========================================================================================
// introspection
// ==========================================================================================
int64_t rdisgpu_num_vars(const rdisgpu_ctx* ctx) { return ctx ? ctx->V : 0; }
int64_t rdisgpu_num_factors(const rdisgpu_ctx* ctx) { return ctx ? ctx->F : 0; }
int rdisgpu_device_state(rdisgpu_ctx* ctx, void** dptr, int64_t* n_pairs) {
  if (!ctx || !dptr || !n_pairs) return RDISGPU_ERR_ARG;
  if (!ctx->finalized) return ctx->fail(RDISGPU_ERR_STATE, "device_state before finalize");
  *dptr = ctx->xbd.p;
  *n_pairs = ctx->V;
  return RDISGPU_OK;
}
int64_t rdisgpu_launch_count(const rdisgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* rdisgpu_version(void) { return "rdis_b200 0.1 (sm_100a, fp64)"; }

}  // extern "C"
