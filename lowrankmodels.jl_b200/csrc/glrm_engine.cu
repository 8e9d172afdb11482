// glrm_engine.cu — host side of the B200 GLRM prox-grad engine and the C ABI (include/glrm_b200.h).
//
// Replaces, behind a C boundary, fit!(glrm::GLRM, params::ProxGradParams) of
// /root/reference/src/algorithms/proxgrad.jl:34-220: problem encoding + upload (create), the whole
// alternating loop with the objective record and the stopping rule (fit), objective() of
// src/evaluate_fit.jl:57-83.  No CPU compute path exists here: without a CUDA device every entry
// point fails with GLRMB200_E_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "glrm_launch.cuh"
#include "glrm_small.cuh"
#include "glrm_vec.cuh"
#include "glrm_dense_host.h"
#include "glrm_eval.cuh"

using namespace glrm;

// ------------------------------------------------------------------------------------------------
// errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_OK(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(GLRMB200_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency: single-GPU users never load it)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) return fail(GLRMB200_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(GLRMB200_E_NCCL, "libnccl lacks %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Broadcast, "ncclBroadcast");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = h;
  return 0;
}
#define NCCL_OK(call)                                                                              \
  do {                                                                                             \
    int r__ = (call);                                                                              \
    if (r__ != 0) return fail(GLRMB200_E_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__)); \
  } while (0)
static const int kNcclDouble = 8;  // ncclFloat64 (nccl.h ncclDataType_t)

// ------------------------------------------------------------------------------------------------
// engine state
struct Side {
  int64_t units = 0;          // total units on this side (m or n)
  int64_t begin = 0, end = 0; // shard owned by this rank
  int64_t full_len = 0;
  int64_t nnz_local = 0;
  int64_t* d_ptr = nullptr;
  int32_t* d_idx = nullptr;
  double* d_val = nullptr;
  int32_t* d_order = nullptr;
  int64_t n_cluster16 = 0, n_cluster = 0, n_heavy = 0, n_light = 0;   // schedule = [8-CTA clusters | 4-CTA clusters | CTA tier | warp tier]
  int32_t* d_order_vec = nullptr;   // units that go through vec_sweep_kernel (block columns / all rows of a problem with them)
  int64_t n_vec = 0;
  int32_t* d_reg_code = nullptr;
  double* d_reg_param = nullptr;
  int64_t* d_reg_payload_ptr = nullptr;   // vector payloads of the regularizers (fixed_latent_features.y, RemQuadReg.m) or nullptr
  double* d_reg_payload = nullptr;
  std::vector<double> h_reg_param;  // for set_reg_scale
  std::vector<int32_t> h_reg_code;
  int reg_uniform = 1;
  double* d_alpha = nullptr;
  double* d_obj = nullptr;    // obj_by_row / obj_by_col
  std::vector<int64_t> bounds;  // [nranks+1] shard boundaries (all ranks)
};

// fully observed problems (csrc/glrm_dense.cuh): A as Julia stores it, the static chunk plan (all features), the plan of the
// features still searching, and the Y-sweep work buffers
struct DenseHost {
  bool on = false;
  int kt = 0;
  int sms = 148;
  int nbuf = 2;                             // tile buffers of A in shared memory
  int ctas_per_sm = 1;
  int64_t lda = 0;                          // leading dimension of the device copy of A (multiple of the 64-row tile)
  double* d_A = nullptr;
  int32_t *d_chunk_ptr = nullptr, *d_feat_list = nullptr, *d_feat_off = nullptr, *d_nchunks = nullptr;
  int32_t *d_chunk_ptr2 = nullptr, *d_feat_list2 = nullptr, *d_feat_off2 = nullptr, *d_nchunks2 = nullptr;
  int32_t *d_nactive = nullptr, *d_active = nullptr;
  int32_t* d_diag = nullptr;                // [8] watchdog record of the dense kernels (glrm_dense.cuh: dn_give_up)
  int max_chunks = 1;
  double *d_gscratch = nullptr, *d_gpart = nullptr, *d_objpart = nullptr, *d_G = nullptr, *d_Ynew = nullptr;
  double *d_colobj = nullptr, *d_objold = nullptr, *d_regnew = nullptr;
  int n_blocks = 1;                         // DN_GROUPS * bg row blocks (fixed by m alone)
  int bg = 1;                               // row blocks per group
  int64_t rows_per_block = 0;
  int64_t row0 = 0, row1 = 0;               // rows this rank owns (multi-GPU: whole groups of row blocks; else all rows)
  int g0 = 0, g1 = DN_GROUPS;               // its groups
  double *d_gsum_G = nullptr, *d_gsum_o = nullptr, *d_gsum_x = nullptr;   // per-group partial sums [DN_GROUPS][...]
  int x_grid = 1;
  // tensor-core kernels (glrm_dense_mma.cuh): factor width in n-tiles, pipeline stages, per-unit column tables of both plans
  bool mma = false;
  int nt = 0, nst = 0;
  int unit_cols = DN_TN;
  int32_t *d_ucol_feat = nullptr, *d_ucol_y = nullptr, *d_ucol_feat2 = nullptr, *d_ucol_y2 = nullptr;
  unsigned long long* d_phase = nullptr;    // GLRMB200_PHASE_TIMERS=1: phase clocks of the X sweep (printed when the handle goes)
  volatile int32_t* h_nactive = nullptr;    // mapped pinned [2]: (features still searching, sweep sequence number)
  int32_t seq = 0;
};

constexpr int EVENT_CHUNK = 64;   // iterations whose timing events are in flight before the host harvests them
constexpr int MAX_RANKS = 64;     // slots of the peer flag array

struct glrmb200_engine {
  int device = 0, rank = 0, nranks = 1;
  int64_t m = 0, n = 0, k = 0, d = 0;
  int kp = 0;
  int stride = 0;           // doubles between factor columns on the device
  int tile_g = 0, tile_r = 0;
  int loss_template = 0;   // 0 generic, else uniform loss code instantiated at compile time
  double uparam[3] = {1, 0, 0};
  int64_t heavy_threshold = 1024;
  int64_t cluster_threshold = 8192;
  int64_t cluster16_threshold = 16384;
  int64_t nnz_rows_total = 0;
  bool obs_full = false;
  bool has_vec = false;              // some column has a vector-valued loss
  int64_t* d_ystart = nullptr;       // [n+1]
  std::vector<int64_t> ystart;
  std::vector<char> col_is_vec, all_rows_vec;   // schedule routing (vector-loss units)
  std::vector<int32_t> h_loss_code;             // host copy (error messages)
  Side rows, cols;
  int32_t* d_loss_code = nullptr;
  double* d_loss_param = nullptr;
  double* d_xchg = nullptr;                  // ONE allocation [X | Y | obj_by_col | peer flags] (2 MiB granules): one IPC handle
  size_t xchg_bytes = 0;
  double* d_X = nullptr;
  double* d_Y = nullptr;
  unsigned long long* d_flags = nullptr;     // [MAX_RANKS] epochs published by the peers (inside d_xchg)
  double* d_scalars = nullptr;              // [4]
  unsigned long long* d_trials = nullptr;   // [2]
  int* d_stop = nullptr;                    // [2]: stop iteration (0 = running), barrier time-out flag
  double* d_objs = nullptr;                 // [objs_cap] objective record of the running fit
  int64_t objs_cap = 0;
  double* d_stage = nullptr;                // staging buffer for contiguous factor transfers [(m + d) * k]
  double* h_pinned = nullptr;               // [16] pinned, device-mapped scratch (pinned_get / pinned_put)
  volatile int* h_stop = nullptr;           // mapped flag written by record_kernel (inside h_pinned)
  cudaStream_t stream = nullptr;
  cudaStream_t tier_stream[3] = {nullptr, nullptr, nullptr};   // cluster-8 / cluster-4 / warp tier of a sweep (the CTA tier runs on `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_base = nullptr, ev_aux = nullptr;
  std::vector<cudaEvent_t> evpool;           // EVENT_CHUNK x 5 per-iteration timing events
  ncclComm_t comm = nullptr;
  bool factors_resident = false;
  // fused exchange: peers' replicas opened through CUDA IPC (index = position among the other ranks)
  bool peer_ready = false;
  std::vector<uint8_t> peer_blobs;           // the IPC blobs the current mappings were opened from
  std::vector<void*> opened;                 // every pointer returned by cudaIpcOpenMemHandle
  double** d_peer_X = nullptr;
  double** d_peer_Y = nullptr;
  double** d_peer_objc = nullptr;
  unsigned long long** d_peer_flags = nullptr;
  int* d_peer_rank = nullptr;
  unsigned long long epoch = 0;              // barriers executed since the flags were last zeroed (glrmb200_ipc_export)
  double* d_barrier = nullptr;
  DenseHost dn;
};

static int g_device_checked = -1;
static int check_device() {
  if (g_device_checked > 0) return 0;
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) {
    cudaGetLastError();
    return fail(GLRMB200_E_NO_DEVICE, "no CUDA device is visible (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  g_device_checked = cnt;
  return 0;
}

// Pinned, device-mapped scratch of a handle: [0..7] doubles for scalar read-backs, [8] the stop flag, [10..13] the dense path's
// ring of 4 (features still searching, plan key) pairs.  cudaHostAlloc / cudaFreeHost are slow and synchronising, so blocks are
// recycled through a process-level free list (distinct handles never share a block: they may run on different threads).
static std::mutex g_pinned_mu;
static std::vector<double*> g_pinned_free;
static int pinned_get(double** out) {
  {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    if (!g_pinned_free.empty()) { *out = g_pinned_free.back(); g_pinned_free.pop_back(); }
    else *out = nullptr;
  }
  if (!*out) CUDA_OK(cudaHostAlloc((void**)out, 16 * sizeof(double), cudaHostAllocMapped | cudaHostAllocPortable));
  memset(*out, 0, 16 * sizeof(double));
  return 0;
}
static void pinned_put(double* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned_free.push_back(p);
}

// ------------------------------------------------------------------------------------------------
// host-side helpers
static int loss_dim(int code, const double* p) {
  switch (code) {
    case GLRMB200_LOSS_MULTINOMIAL: case GLRMB200_LOSS_OVA: case GLRMB200_LOSS_ORDISTIC: return (int)p[2];
    case GLRMB200_LOSS_BVS: case GLRMB200_LOSS_MULTINOMIAL_ORDINAL: return (int)p[2] - 1;
    default: return 1;
  }
}

// Predicted cost of updating one unit, in entry slots: a pass walks the list in chunks of 32 entries (a partial chunk
// costs a whole one) and every unit pays a fixed overhead (pipeline start-up, line search, write-back) of about one chunk.
static inline int64_t unit_cost(int64_t deg) { return ((deg + 31) / 32) * 32 + 32; }

extern "C" int glrmb200_plan_shards(const int64_t* ptr, int64_t count, int32_t nranks, int64_t* bounds) {
  // contiguous shards balanced by predicted cost (ptr == NULL: by unit count)
  if (nranks < 1 || count < 0 || !bounds) return fail(GLRMB200_E_INVALID, "plan_shards: bad arguments");
  bounds[0] = 0;
  bounds[nranks] = count;
  if (!ptr || ptr[count] - ptr[0] == 0) {
    for (int r = 1; r < nranks; ++r) bounds[r] = count * r / nranks;
    return 0;
  }
  std::vector<int64_t> cum((size_t)count + 1);
  cum[0] = 0;
  for (int64_t u = 0; u < count; ++u) cum[(size_t)u + 1] = cum[(size_t)u] + unit_cost(ptr[u + 1] - ptr[u]);
  const int64_t total = cum[(size_t)count];
  for (int r = 1; r < nranks; ++r) {
    const int64_t target = (total / nranks) * r + std::min<int64_t>(r, total % nranks);
    auto it = std::lower_bound(cum.begin(), cum.end(), target);
    if (it != cum.begin() && it != cum.end() && target - *(it - 1) < *it - target) --it;   // nearest boundary
    int64_t b = it - cum.begin();
    if (b > count) b = count;
    if (b < bounds[r - 1]) b = bounds[r - 1];
    bounds[r] = b;
  }
  return 0;
}

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync, release threshold = keep everything):
// allocation and free cost microseconds instead of the 0.1-1 ms of cudaMalloc / cudaFree — and pool memory is not
// mapped into the peers, so freeing it does not pay the cross-device unmap that made destroy cost 0.22 s at 8 GPUs
// once CUDA IPC had enabled peer access (round-1 SCALE record).  Only the exchange allocation, which must be
// exportable through CUDA IPC, is a plain cudaMalloc (and is cached per process, see XchgCache).
static int pool_setup(int device) {
  static std::vector<char> done;
  if ((int)done.size() <= device) done.resize((size_t)device + 1, 0);
  if (done[(size_t)device]) return 0;
  cudaMemPool_t pool;
  CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = UINT64_MAX;
  CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  done[(size_t)device] = 1;
  return 0;
}
template <class T>
static int dalloc(T** dst, size_t count, cudaStream_t st) {
  *dst = nullptr;
  CUDA_OK(cudaMallocAsync((void**)dst, std::max<size_t>(1, count) * sizeof(T), st));
  return 0;
}
template <class T>
static void dfree(T*& p, cudaStream_t st) {
  if (p) cudaFreeAsync((void*)p, st);
  p = nullptr;
}

// `pad` extra zeroed elements follow the payload: the entry passes read idx[start] / val[start] of a unit even when it
// is empty, and an empty unit at the end of a shard has start == nnz_local.  The copy is stream-ordered: from pinned
// host memory it is a true asynchronous DMA (the caller synchronises before handing the host buffer back), from
// pageable memory cudaMemcpyAsync returns once the source has been staged.
template <class T>
static int upload(T** dst, const T* src, size_t count, cudaStream_t st, size_t pad = 0) {
  if (count == 0) src = nullptr;
  const size_t alloc = std::max<size_t>(1, count + pad);
  int rc = dalloc(dst, alloc, st);
  if (rc) return rc;
  if (alloc > count) CUDA_OK(cudaMemsetAsync(*dst + count, 0, (alloc - count) * sizeof(T), st));
  if (src) CUDA_OK(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

// degree-sorted schedule (heaviest first): LPT order for the tail, and neighbouring warps of a CTA get
// units of similar length.  `is_vec` (optional) routes units to the vector-loss kernel instead.
// NB: the tier of a unit (warp / CTA / cluster) fixes its reduction tree, so the thresholds are constants of the
// unit's degree — never of the rank count — or a sharded fit would stop being bit-identical to the 1-GPU fit.
static int build_schedule(glrmb200_engine* E, Side& S, const int64_t* ptr_global, const std::vector<char>* is_vec = nullptr) {
  const int64_t cnt = S.end - S.begin;
  std::vector<int32_t> order, vec;
  order.reserve((size_t)cnt);
  for (int64_t u = S.begin; u < S.end; ++u) {
    if (is_vec && (*is_vec)[(size_t)u]) vec.push_back((int32_t)u); else order.push_back((int32_t)u);
  }
  auto deg = [&](int32_t u) { return ptr_global ? ptr_global[u + 1] - ptr_global[u] : S.full_len; };
  if (ptr_global) {
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return deg(a) > deg(b); });
    std::stable_sort(vec.begin(), vec.end(), [&](int32_t a, int32_t b) { return deg(a) > deg(b); });
  }
  int64_t pos = 0;
  const int64_t total = (int64_t)order.size();
  auto take = [&](int64_t threshold) { const int64_t p0 = pos; while (pos < total && deg(order[(size_t)pos]) >= threshold) ++pos; return pos - p0; };
  S.n_cluster16 = take(E->cluster16_threshold);
  S.n_cluster = take(E->cluster_threshold);
  S.n_heavy = take(E->heavy_threshold);
  S.n_light = total - pos;
  S.n_vec = (int64_t)vec.size();
  int rc = upload(&S.d_order, order.data(), order.size(), E->stream);
  if (rc) return rc;
  return upload(&S.d_order_vec, vec.data(), vec.size(), E->stream);
}

static int setup_regs(glrmb200_engine* E, Side& S, int64_t count, const int32_t* code, const double* param, bool allow_ordinal,
                      const int64_t* payload_ptr, const double* payload) {
  if (count != 1 && count != S.units) return fail(GLRMB200_E_INVALID, "regularizer count must be 1 or the number of columns");
  S.reg_uniform = (count == 1);
  S.h_reg_code.assign(code, code + count);
  S.h_reg_param.assign(param, param + count * GLRMB200_REG_NPARAM);
  constexpr int kFixed = GLRMB200_REG_FIXED_FIRST | GLRMB200_REG_FIXED_LAST;
  constexpr int kOffset = GLRMB200_REG_LASTENTRY1 | GLRMB200_REG_LASTENTRY_UNPENALIZED;
  constexpr int kBlock = GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL;
  for (int64_t i = 0; i < count; ++i) {
    const int base = code[i] & GLRMB200_REG_BASE_MASK;
    const int ord = code[i] & kBlock;
    if (ord && (!allow_ordinal || (code[i] & (kOffset | kFixed))))
      return fail(GLRMB200_E_UNSUPPORTED, "OrdinalReg / MNLOrdinalReg are column (ry) regularizers and are not combined with other wrappers");
    if (base > GLRMB200_REG_REM_QUAD || (code[i] & ~(GLRMB200_REG_BASE_MASK | kOffset | kBlock | kFixed)))
      return fail(GLRMB200_E_UNSUPPORTED, "regularizer code %d has no device implementation", code[i]);
    const bool fixed = code[i] & kFixed;
    if (fixed && ((code[i] & kOffset) || base == GLRMB200_REG_REM_QUAD || (code[i] & kFixed) == kFixed))
      return fail(GLRMB200_E_UNSUPPORTED, "fixed_latent_features is not combined with the offset wrappers, RemQuadReg or its twin");
    if (base == GLRMB200_REG_REM_QUAD && (code[i] & kOffset))
      return fail(GLRMB200_E_UNSUPPORTED, "RemQuadReg inside an offset wrapper has no device implementation");
    if (fixed || base == GLRMB200_REG_REM_QUAD) {
      if (!payload_ptr || !payload) return fail(GLRMB200_E_INVALID, "regularizer %lld needs a vector payload (rx_payload / ry_payload)", (long long)i);
      const int64_t len = payload_ptr[i + 1] - payload_ptr[i];
      if (len < 0 || len > E->k || (base == GLRMB200_REG_REM_QUAD && !fixed && len != E->k))
        return fail(GLRMB200_E_INVALID, "regularizer %lld: payload of length %lld with k = %lld", (long long)i, (long long)len, (long long)E->k);
    }
  }
  int rc = upload(&S.d_reg_code, code, (size_t)count, E->stream);
  if (rc) return rc;
  if ((rc = upload(&S.d_reg_param, param, (size_t)count * GLRMB200_REG_NPARAM, E->stream))) return rc;
  if (payload_ptr && payload) {
    if ((rc = upload(&S.d_reg_payload_ptr, payload_ptr, (size_t)count + 1, E->stream))) return rc;
    if ((rc = upload(&S.d_reg_payload, payload, (size_t)std::max<int64_t>(1, payload_ptr[count]), E->stream))) return rc;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch over the (G, R) tile and the loss template: the instantiations live in sweep_inst.cu, compiled once
// per (loss template, tile half) so the build parallelises (csrc/Makefile)
struct Tile { int g, r; };
static const Tile kTiles[] = {{4, 1}, {8, 1}, {8, 2}, {8, 3}, {8, 4}, {16, 2}, {16, 3}, {16, 4}, {32, 2}, {32, 3}, {32, 4}};

static cudaError_t launch_sweep(const glrmb200_engine* E, const SweepArgs& A, const Side& S, int64_t* launches) {
  const TierCounts tc{S.n_cluster16, S.n_cluster, S.n_heavy, S.n_light};
  const bool two_streams = !(getenv("GLRMB200_ONE_STREAM") && atoi(getenv("GLRMB200_ONE_STREAM")));
  Streams st;
  st.main = E->stream;
  st.fork = E->ev_fork;
  const bool warp_only = getenv("GLRMB200_TIER_STREAMS") && atoi(getenv("GLRMB200_TIER_STREAMS")) == 1;   // A/B hook: round-1 layout
  for (int i = 0; i < 3; ++i) { st.tier[i] = (two_streams && !(warp_only && i < 2)) ? E->tier_stream[i] : nullptr; st.join[i] = E->ev_join[i]; }
  const bool wide = E->tile_g >= 16;
  switch (E->loss_template) {
    case GLRMB200_LOSS_QUAD: return (wide ? launch_quad_wide : launch_quad_narrow)(E->tile_g, E->tile_r, A, tc, st, launches);
    case GLRMB200_LOSS_LOGISTIC: return (wide ? launch_logistic_wide : launch_logistic_narrow)(E->tile_g, E->tile_r, A, tc, st, launches);
    default: return (wide ? launch_generic_wide : launch_generic_narrow)(E->tile_g, E->tile_r, A, tc, st, launches);
  }
}

// units that involve vector-valued losses (csrc/glrm_vec.cuh, instantiated in vec_inst.cu): one warp per unit
namespace glrm { cudaError_t launch_vec_inst(int g, int r, const VecArgs& V, bool x_side, int64_t n_vec, cudaStream_t stream, int64_t* launches); }
static cudaError_t launch_vec(const glrmb200_engine* E, const SweepArgs& A, bool x_side, const Side& S, int64_t* launches) {
  if (S.n_vec == 0) return cudaSuccess;
  VecArgs V;
  V.s = A;
  V.s.order = S.d_order_vec;
  V.s.n_units = S.n_vec;
  V.s.own_col = nullptr;
  V.ystart = E->d_ystart;
  V.x_side = x_side ? 1 : 0;
  return launch_vec_inst(E->tile_g, E->tile_r, V, x_side, S.n_vec, E->stream, launches);
}

static cudaError_t launch_reg_eval(const glrmb200_engine* E, const double* own, const Side& S, double* out) {
  const int g = E->tile_g, r = E->tile_r;
  const int64_t per_cta = 4 * (32 / g);
  const unsigned grid = (unsigned)((S.units + per_cta - 1) / per_cta);
#define T(GG, RR) if (g == GG && r == RR) { reg_eval_kernel<GG, RR><<<grid, 128, 0, E->stream>>>(own, S.units, E->stride, (int)E->k, S.d_reg_code, S.d_reg_param, S.reg_uniform, out, S.d_reg_payload_ptr, S.d_reg_payload); return cudaGetLastError(); }
  T(4, 1) T(8, 1) T(8, 2) T(8, 3) T(8, 4) T(16, 2) T(16, 3) T(16, 4) T(32, 2) T(32, 3) T(32, 4)
#undef T
  return cudaErrorInvalidValue;
}

static SweepArgs make_args(const glrmb200_engine* E, bool x_side, int flags, double min_stepsize, bool honour_stop = false) {
  const Side& S = x_side ? E->rows : E->cols;
  SweepArgs A;
  A.ptr = S.d_ptr;
  A.idx = S.d_idx;
  A.val = S.d_val;
  A.full_len = S.full_len;
  A.unit_base = S.begin;
  A.order = S.d_order;
  A.n_units = 0;
  A.own = x_side ? E->d_X : E->d_Y;
  A.own_col = (!x_side && E->has_vec) ? E->d_ystart : nullptr;
  A.opp = x_side ? E->d_Y : E->d_X;
  A.stride = E->stride;
  A.last_lanes = E->kp / 2 - E->tile_g * (E->tile_r - 1);
  A.k = (int)E->k;
  A.loss_code = E->d_loss_code;
  A.loss_param = E->d_loss_param;
  A.uparam[0] = E->uparam[0]; A.uparam[1] = E->uparam[1]; A.uparam[2] = E->uparam[2];
  A.reg_code = S.d_reg_code;
  A.reg_param = S.d_reg_param;
  A.reg_payload_ptr = S.d_reg_payload_ptr;
  A.reg_payload = S.d_reg_payload;
  A.reg_uniform = S.reg_uniform;
  A.flags = flags | ((x_side && E->loss_template == 0) ? FLAG_LOSS_BY_ENTRY : 0);
  A.alpha = S.d_alpha;
  A.min_stepsize = min_stepsize;
  A.global_alpha = 0.0;
  A.obj_out = S.d_obj;
  A.trial_counter = E->d_trials + (x_side ? 0 : 1);
  A.peer_own = E->peer_ready ? (x_side ? E->d_peer_X : E->d_peer_Y) : nullptr;
  A.peer_obj = (E->peer_ready && !x_side) ? E->d_peer_objc : nullptr;
  A.n_peers = E->peer_ready ? E->nranks - 1 : 0;
  A.stop = honour_stop ? E->d_stop : nullptr;
  return A;
}

// ------------------------------------------------------------------------------------------------
// fully observed path (csrc/glrm_dense.cuh)
static void dense_free(glrmb200_engine* E) {
  DenseHost& D = E->dn;
  if (D.d_phase) {
    unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyAsync(ph, D.d_phase, sizeof(ph), cudaMemcpyDeviceToHost, E->stream) == cudaSuccess && cudaStreamSynchronize(E->stream) == cudaSuccess && ph[7])
      fprintf(stderr, "[glrmb200] X sweep phases, clocks per tile (thread 0): wait %.0f, gradient pass %.0f, search set-up %.0f, trial points %.0f, "
                      "trial losses %.0f, accept/compact %.0f; rounds per tile %.2f; tiles %llu\n",
              (double)ph[0] / ph[7], (double)ph[1] / ph[7], (double)ph[2] / ph[7], (double)ph[3] / ph[7], (double)ph[4] / ph[7], (double)ph[5] / ph[7],
              (double)ph[6] / ph[7], ph[7]);
    dfree(D.d_phase, E->stream);
  }
  dfree(D.d_A, E->stream);
  dfree(D.d_chunk_ptr, E->stream); dfree(D.d_feat_list, E->stream); dfree(D.d_feat_off, E->stream); dfree(D.d_nchunks, E->stream);
  dfree(D.d_chunk_ptr2, E->stream); dfree(D.d_feat_list2, E->stream); dfree(D.d_feat_off2, E->stream); dfree(D.d_nchunks2, E->stream);
  dfree(D.d_nactive, E->stream); dfree(D.d_active, E->stream); dfree(D.d_diag, E->stream);
  dfree(D.d_gscratch, E->stream); dfree(D.d_gpart, E->stream); dfree(D.d_objpart, E->stream); dfree(D.d_G, E->stream);
  dfree(D.d_Ynew, E->stream); dfree(D.d_colobj, E->stream); dfree(D.d_objold, E->stream); dfree(D.d_regnew, E->stream);
  dfree(D.d_gsum_G, E->stream); dfree(D.d_gsum_o, E->stream); dfree(D.d_gsum_x, E->stream);
  dfree(D.d_ucol_feat, E->stream); dfree(D.d_ucol_y, E->stream); dfree(D.d_ucol_feat2, E->stream); dfree(D.d_ucol_y2, E->stream);
  D.on = false;
}

// can this problem take the dense kernels?  fully observed, rank within the register-tile range, regularizers that act
// element-wise on block columns (the ordinal block regularizers stay on the gather path), one rank
static bool dense_eligible(const glrmb200_engine* E, const glrmb200_problem* P) {
  if (!E->obs_full || E->k > 104) return false;
  // several GPUs (mode B): rows are sharded by whole groups of row blocks, Y is replicated; needs the
  // group count to divide evenly and enough rows for every rank to own some
  if (E->nranks != 1 && !((E->nranks == 2 || E->nranks == 4 || E->nranks == 8) && E->m >= 512 * (int64_t)E->nranks)) return false;
  if (const char* t = getenv("GLRMB200_DENSE")) { if (atoi(t) == 0) return false; }
  if (getenv("GLRMB200_TILE")) return false;
  for (int64_t i = 0; i < P->ry_count; ++i)
    if (P->ry_code[i] & (GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL)) return false;
  if (P->rx_payload_ptr || P->ry_payload_ptr) return false;      // regularizers with vector payloads run on the gather kernels
  return true;
}

// row blocks of the Y sweep: DN_GROUPS groups of up to 37 blocks (two CTA waves of 148), fixed by m alone so that neither
// the reduction order nor the multi-GPU row shards depend on the launch
static void dense_blocks(int64_t m, int* bg, int64_t* rows_per_block) {
  // t 64-row tiles per block so that 8 x 37 blocks cover m; then just as many blocks per group as those tiles need (a
  // mid-size m would otherwise leave the last groups — and, sharded, the last GPUs — without rows)
  const int64_t tiles = (m + DN_TM - 1) / DN_TM;
  const int64_t t = std::max<int64_t>(1, (tiles + DN_GROUPS * 37 - 1) / (DN_GROUPS * 37));
  *bg = (int)std::max<int64_t>(1, (tiles + DN_GROUPS * t - 1) / (DN_GROUPS * t));
  *rows_per_block = t * DN_TM;
}

// rows of a fully observed problem on `nranks` GPUs: rank r owns the row-block groups [r G / N, (r + 1) G / N)
static void dense_row_bounds(int64_t m, int nranks, int64_t* bounds) {
  int bg;
  int64_t rpb;
  dense_blocks(m, &bg, &rpb);
  for (int r = 0; r <= nranks; ++r) bounds[r] = std::min<int64_t>(m, (int64_t)(r * DN_GROUPS / nranks) * bg * rpb);
}
extern "C" int glrmb200_plan_dense_rows(int64_t m, int32_t nranks, int64_t* bounds) {
  if (!bounds || m <= 0) return fail(GLRMB200_E_INVALID, "bad argument");
  if (nranks == 1) { bounds[0] = 0; bounds[1] = m; return 0; }
  if (!((nranks == 2 || nranks == 4 || nranks == 8) && m >= 512 * (int64_t)nranks))
    return fail(GLRMB200_E_UNSUPPORTED, "row sharding of a fully observed problem needs 2, 4 or 8 ranks and m >= 512 per rank "
                                        "(otherwise the handle shards observation lists: glrmb200_plan_shards)");
  dense_row_bounds(m, nranks, bounds);
  return 0;
}

static int dense_setup(glrmb200_engine* E, const glrmb200_problem* P) {
  DenseHost& D = E->dn;
  const int64_t m = E->m, n = E->n, d = E->d;
  int rc;
  D.kt = (int)((E->k + 15) / 16);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, E->device);
  D.sms = sms;
  // row blocks / groups, and the rows this rank owns (whole groups)
  dense_blocks(m, &D.bg, &D.rows_per_block);
  D.n_blocks = DN_GROUPS * D.bg;
  D.g0 = E->rank * DN_GROUPS / E->nranks;
  D.g1 = (E->rank + 1) * DN_GROUPS / E->nranks;
  D.row0 = std::min<int64_t>(m, (int64_t)D.g0 * D.bg * D.rows_per_block);
  D.row1 = std::min<int64_t>(m, (int64_t)D.g1 * D.bg * D.rows_per_block);
  const int64_t mloc = D.row1 - D.row0;
  // A as Julia stores it (column-major), this rank's rows only, columns padded to whole 64-row tiles: every tile column is
  // one aligned 512-byte bulk copy, and the rows past the end read as zero.  (row0 is a multiple of the tile height.)
  D.lda = std::max<int64_t>(DN_TM, (mloc + DN_TM - 1) / DN_TM * DN_TM);
  if ((rc = dalloc(&D.d_A, (size_t)(D.lda * n), E->stream))) return rc;
  if (D.lda == m) {
    CUDA_OK(cudaMemcpyAsync(D.d_A, P->dense_A, (size_t)(m * n) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
  } else {
    CUDA_OK(cudaMemsetAsync(D.d_A, 0, (size_t)(D.lda * n) * sizeof(double), E->stream));
    if (mloc > 0)
      CUDA_OK(cudaMemcpy2DAsync(D.d_A, (size_t)D.lda * sizeof(double), P->dense_A + D.row0, (size_t)m * sizeof(double),
                                (size_t)mloc * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, E->stream));
  }
  // tensor-core kernels (default): as many pipeline stages as fit next to the own tile; GLRMB200_DENSE_MMA=0 keeps the
  // FP64-FMA kernels of glrm_dense.cuh
  {
    const bool generic = E->loss_template != GLRMB200_LOSS_QUAD;
    D.nt = mm_nt_for_k((int)E->k);
    D.mma = D.nt > 0 && 8 * D.nt <= E->stride && !(getenv("GLRMB200_DENSE_MMA") && atoi(getenv("GLRMB200_DENSE_MMA")) == 0);
    D.nst = 0;
    if (D.mma) {
      for (int s = MM_MAX_STAGES; s >= 2 && D.nst == 0; --s)
        if (mm_smem_bytes(D.nt, s, generic) <= (size_t)227 * 1024) D.nst = s;
      if (const char* t = getenv("GLRMB200_MMA_STAGES")) { const int v = atoi(t); if (v >= 2 && v <= D.nst) D.nst = v; }
      if (D.nst == 0) D.mma = false;
    }
    D.unit_cols = D.mma ? MM_UNIT : DN_TN;
  }
  // shared memory: two tile buffers of A when a CTA fills the SM anyway, one when that lets two CTAs share the SM
  if (!D.mma) {
    const size_t cap = 227 * 1024;
    const size_t s1 = dense_smem_needed((int)E->k, D.kt, 1) + 1024, s2 = dense_smem_needed((int)E->k, D.kt, 2) + 1024;
    if (s1 > cap) return fail(GLRMB200_E_UNSUPPORTED, "dense path: k = %lld needs too much shared memory", (long long)E->k);
    if (2 * s2 <= cap) { D.nbuf = 2; D.ctas_per_sm = 2; }
    else if (2 * s1 <= cap) { D.nbuf = 1; D.ctas_per_sm = 2; }
    else if (s2 <= cap) { D.nbuf = 2; D.ctas_per_sm = 1; }
    else { D.nbuf = 1; D.ctas_per_sm = 1; }
    if (const char* t = getenv("GLRMB200_DENSE_NBUF")) { const int v = atoi(t); if (v == 1 || (v == 2 && s2 <= cap)) { D.nbuf = v; D.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, cap / (v == 1 ? s1 : s2))); } }
    if (const char* t = getenv("GLRMB200_DENSE_CTAS")) { const int v = atoi(t); if (v == 1) D.ctas_per_sm = 1; }
  }
  // static plan: all features in order, chunks of <= unit_cols columns made of whole features
  const int tn = D.unit_cols;
  std::vector<int32_t> chunk_ptr{0}, feat_list, feat_off, ucol_feat, ucol_y;
  int used = 0;
  for (int64_t f = 0; f < n; ++f) {
    const int dim = (int)(E->ystart[(size_t)f + 1] - E->ystart[(size_t)f]);
    if (used + dim > tn) {
      chunk_ptr.push_back((int32_t)feat_list.size());
      ucol_feat.resize(ucol_feat.size() + (size_t)(tn - used), -1); ucol_y.resize(ucol_y.size() + (size_t)(tn - used), -1);
      used = 0;
    }
    feat_list.push_back((int32_t)f);
    feat_off.push_back(used);
    for (int c = 0; c < dim; ++c) { ucol_feat.push_back((int32_t)f); ucol_y.push_back((int32_t)(E->ystart[(size_t)f] + c)); }
    used += dim;
  }
  chunk_ptr.push_back((int32_t)feat_list.size());
  ucol_feat.resize(ucol_feat.size() + (size_t)(tn - used), -1); ucol_y.resize(ucol_y.size() + (size_t)(tn - used), -1);
  const int32_t nchunks = (int32_t)chunk_ptr.size() - 1;
  D.max_chunks = 2 * (int)((d + tn - 1) / tn) + 1;                 // next-fit bound for any subset of the features
  if (D.max_chunks < nchunks) D.max_chunks = nchunks;
  if (D.mma) {
    if ((rc = upload(&D.d_ucol_feat, ucol_feat.data(), ucol_feat.size(), E->stream))) return rc;
    if ((rc = upload(&D.d_ucol_y, ucol_y.data(), ucol_y.size(), E->stream))) return rc;
    if ((rc = dalloc(&D.d_ucol_feat2, (size_t)D.max_chunks * MM_UNIT, E->stream))) return rc;
    if ((rc = dalloc(&D.d_ucol_y2, (size_t)D.max_chunks * MM_UNIT, E->stream))) return rc;
  }
  if ((rc = upload(&D.d_chunk_ptr, chunk_ptr.data(), chunk_ptr.size(), E->stream))) return rc;
  if ((rc = upload(&D.d_feat_list, feat_list.data(), feat_list.size(), E->stream))) return rc;
  if ((rc = upload(&D.d_feat_off, feat_off.data(), feat_off.size(), E->stream))) return rc;
  if ((rc = upload(&D.d_nchunks, &nchunks, 1, E->stream))) return rc;
  if ((rc = dalloc(&D.d_chunk_ptr2, (size_t)n + 2, E->stream))) return rc;
  if ((rc = dalloc(&D.d_feat_list2, (size_t)n + 1, E->stream))) return rc;
  if ((rc = dalloc(&D.d_feat_off2, (size_t)n + 1, E->stream))) return rc;
  if ((rc = dalloc(&D.d_nchunks2, 1, E->stream))) return rc;
  if ((rc = dalloc(&D.d_nactive, 1, E->stream))) return rc;
  if ((rc = dalloc(&D.d_active, (size_t)n, E->stream))) return rc;
  if ((rc = dalloc(&D.d_diag, 8, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(D.d_diag, 0, 8 * sizeof(int32_t), E->stream));
  CUDA_OK(cudaMemsetAsync(D.d_active, 0, (size_t)n * sizeof(int32_t), E->stream));
  CUDA_OK(cudaMemsetAsync(D.d_nactive, 0, sizeof(int32_t), E->stream));
  CUDA_OK(cudaMemsetAsync(D.d_nchunks2, 0, sizeof(int32_t), E->stream));
  const int tile_rows = D.mma ? MM_TM : DN_TM;
  const int64_t tiles = std::max<int64_t>(1, (mloc + tile_rows - 1) / tile_rows);
  D.x_grid = (int)std::min<int64_t>(tiles, D.mma ? (int64_t)sms : (int64_t)sms * D.ctas_per_sm);
  const size_t gs_doubles = (size_t)D.x_grid * tile_rows * E->stride * (D.mma ? 2 : 1);   // (tensor-core kernels: + the rows' last trial points)
  if ((rc = dalloc(&D.d_gscratch, gs_doubles, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(D.d_gscratch, 0, gs_doubles * sizeof(double), E->stream));   // lanes past the register tiles stay zero
  if ((rc = dalloc(&D.d_gpart, (size_t)D.n_blocks * (size_t)d * E->stride, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(D.d_gpart, 0, (size_t)D.n_blocks * (size_t)d * E->stride * sizeof(double), E->stream));   // (the same)
  if ((rc = dalloc(&D.d_objpart, (size_t)D.n_blocks * (size_t)n, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(D.d_objpart, 0, (size_t)D.n_blocks * (size_t)n * sizeof(double), E->stream));
  if ((rc = dalloc(&D.d_G, (size_t)d * E->stride, E->stream))) return rc;
  if ((rc = dalloc(&D.d_Ynew, (size_t)d * E->stride, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(D.d_Ynew, 0, (size_t)d * E->stride * sizeof(double), E->stream));
  if ((rc = dalloc(&D.d_gsum_G, (size_t)DN_GROUPS * (size_t)d * E->stride, E->stream))) return rc;
  if ((rc = dalloc(&D.d_gsum_o, (size_t)DN_GROUPS * (size_t)n, E->stream))) return rc;
  if ((rc = dalloc(&D.d_gsum_x, DN_GROUPS, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(D.d_gsum_G, 0, (size_t)DN_GROUPS * (size_t)d * E->stride * sizeof(double), E->stream));
  CUDA_OK(cudaMemsetAsync(D.d_gsum_o, 0, (size_t)DN_GROUPS * (size_t)n * sizeof(double), E->stream));
  CUDA_OK(cudaMemsetAsync(D.d_gsum_x, 0, DN_GROUPS * sizeof(double), E->stream));
  if (D.mma && getenv("GLRMB200_PHASE_TIMERS") && atoi(getenv("GLRMB200_PHASE_TIMERS"))) {
    if ((rc = dalloc(&D.d_phase, 8, E->stream))) return rc;
    CUDA_OK(cudaMemsetAsync(D.d_phase, 0, 8 * sizeof(unsigned long long), E->stream));
  }
  if ((rc = dalloc(&D.d_colobj, (size_t)n, E->stream))) return rc;
  if ((rc = dalloc(&D.d_objold, (size_t)n, E->stream))) return rc;
  if ((rc = dalloc(&D.d_regnew, (size_t)n, E->stream))) return rc;
  D.h_nactive = reinterpret_cast<volatile int32_t*>(E->h_pinned + 10);
  D.on = true;
  return 0;
}

static DenseArgs dense_args(const glrmb200_engine* E, bool static_plan, const double* Ymat, int flags, double min_stepsize, bool honour_stop) {
  const DenseHost& D = E->dn;
  DenseArgs P;
  memset(&P, 0, sizeof(P));
  P.A = D.d_A - D.row0;                         // indexed by the global row: A[f * lda + e], e in [row0, row1)
  P.m = E->m; P.n = E->n; P.lda = D.lda; P.nbuf = D.nbuf; P.diag = D.d_diag;
  P.row0 = D.row0; P.row1 = D.row1;
  P.block0 = D.g0 * D.bg;
  P.X = E->d_X; P.Ymat = Ymat;
  P.stride = E->stride; P.k = (int)E->k; P.kp = E->kp;
  P.ystart = E->d_ystart;
  P.loss_code = E->d_loss_code; P.loss_param = E->d_loss_param;
  P.chunk_ptr = static_plan ? D.d_chunk_ptr : D.d_chunk_ptr2;
  P.feat_list = static_plan ? D.d_feat_list : D.d_feat_list2;
  P.feat_off = static_plan ? D.d_feat_off : D.d_feat_off2;
  P.nchunks = static_plan ? D.d_nchunks : D.d_nchunks2;
  P.reg_code = E->rows.d_reg_code; P.reg_param = E->rows.d_reg_param; P.reg_uniform = E->rows.reg_uniform;
  P.alpha = E->rows.d_alpha; P.min_stepsize = min_stepsize; P.obj_out = E->rows.d_obj;
  P.gscratch = D.d_gscratch;
  P.trial_counter = E->d_trials;
  P.stop = honour_stop ? E->d_stop : nullptr;
  P.flags = flags;
  P.gpart = D.d_gpart; P.objpart = D.d_objpart;
  P.rows_per_block = D.rows_per_block; P.n_blocks = D.n_blocks;
  P.ucol_feat = static_plan ? D.d_ucol_feat : D.d_ucol_feat2;
  P.ucol_y = static_plan ? D.d_ucol_y : D.d_ucol_y2;
  P.nst = D.nst;
  P.phase = D.d_phase;
  P.uparam[0] = E->uparam[0]; P.uparam[1] = E->uparam[1]; P.uparam[2] = E->uparam[2];
  return P;
}
static DenseYState dense_ystate(glrmb200_engine* E, int flags, double min_stepsize, bool honour_stop) {
  DenseHost& D = E->dn;
  DenseYState Q;
  memset(&Q, 0, sizeof(Q));
  Q.Y = E->d_Y; Q.Ynew = D.d_Ynew; Q.G = D.d_G;
  Q.ystart = E->d_ystart;
  Q.colobj = D.d_colobj; Q.objold = D.d_objold; Q.regnew = D.d_regnew;
  Q.alpha = E->cols.d_alpha; Q.obj_out = E->cols.d_obj;
  Q.active = D.d_active; Q.nactive = D.d_nactive; Q.h_nactive = D.h_nactive;
  Q.chunk_ptr = D.d_chunk_ptr2; Q.feat_list = D.d_feat_list2; Q.feat_off = D.d_feat_off2; Q.nchunks = D.d_nchunks2;
  Q.reg_code = E->cols.d_reg_code; Q.reg_param = E->cols.d_reg_param; Q.reg_uniform = E->cols.reg_uniform;
  Q.n = E->n; Q.d = E->d; Q.m = E->m;
  Q.stride = E->stride; Q.k = (int)E->k;
  Q.min_stepsize = min_stepsize;
  Q.trial_counter = E->d_trials + 1;
  Q.stop = honour_stop ? E->d_stop : nullptr;
  Q.flags = flags;
  Q.seq = D.seq;
  Q.unit_cols = D.unit_cols;
  Q.ucol_feat = D.d_ucol_feat2; Q.ucol_y = D.d_ucol_y2;
  return Q;
}
#define DN_OK(call)                                                                                                   \
  do {                                                                                                                \
    cudaError_t ce__ = (call);                                                                                        \
    if (ce__ != cudaSuccess) return fail(GLRMB200_E_CUDA, "%s: %s", #call, cudaGetErrorString(ce__));                  \
  } while (0)

// the dense kernels' watchdog record (synchronises the stream)
static int dense_check_diag(glrmb200_engine* E) {
  if (!E->dn.on || !E->dn.d_diag) return 0;
  int32_t dg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CUDA_OK(cudaMemcpyAsync(dg, E->dn.d_diag, sizeof(dg), cudaMemcpyDeviceToHost, E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  if (dg[0] == 0) return 0;
  cudaMemsetAsync(E->dn.d_diag, 0, sizeof(dg), E->stream);
  return fail(GLRMB200_E_STATE, "dense kernel watchdog: %s (block %d,%d thread %d; %d %d %d)",
              dg[0] == 1 ? "a tile of A never arrived" : "the line search did not end", dg[1], dg[2], dg[3], dg[4], dg[5], dg[6]);
}
// GLRMB200_DENSE_DEBUG=1: synchronise after every dense launch and say which one it was (hang hunting)
static int dense_debug(glrmb200_engine* E, const char* what) {
  static const bool on = getenv("GLRMB200_DENSE_DEBUG") && atoi(getenv("GLRMB200_DENSE_DEBUG"));
  if (!on) return 0;
  fprintf(stderr, "[dense] %s ...", what); fflush(stderr);
  const cudaError_t ce = cudaStreamSynchronize(E->stream);
  fprintf(stderr, " %s\n", ce == cudaSuccess ? "done" : cudaGetErrorString(ce)); fflush(stderr);
  if (ce != cudaSuccess) return fail(GLRMB200_E_CUDA, "%s: %s", what, cudaGetErrorString(ce));
  return dense_check_diag(E);
}
#define DN_DBG(what) do { int drc__ = dense_debug(E, what); if (drc__) return drc__; } while (0)

static int dense_sweep_x(glrmb200_engine* E, double min_stepsize, bool honour_stop, int64_t* launches) {
  const DenseHost& D = E->dn;
  const int loss = E->loss_template == GLRMB200_LOSS_QUAD ? GLRMB200_LOSS_QUAD : 0;
  DenseArgs P = dense_args(E, true, E->d_Y, 0, min_stepsize, honour_stop);
  if (D.mma) DN_OK(dense_mma_launch_x(D.nt, E->tile_g, E->tile_r, loss, P, D.x_grid, E->stream));
  else DN_OK(dense_launch_x(D.kt, E->tile_g, E->tile_r, loss, P, D.x_grid, E->stream));
  DN_DBG("dense_x_kernel");
  ++*launches;
  return 0;
}

// all-gather of per-group partial sums ([DN_GROUPS][len], every rank owns the groups [g0, g1)); no-op on one rank
static int dense_gather_groups(glrmb200_engine* E, double* gsum, int64_t len);

static double dense_wait_limit_s() {
  if (const char* t = getenv("GLRMB200_WAIT_LIMIT_S")) return std::max(1.0, atof(t));
  return 120.0;
}

// losses of every feature at the current Y -> colobj (mode 0: and the gradient G_Y); then obj_by_col = loss + ry and the
// search state (dense_y_begin_kernel)
static int dense_eval_cols(glrmb200_engine* E, int flags, double min_stepsize, bool honour_stop, int mode, int64_t* launches) {
  DenseHost& D = E->dn;
  const int loss = E->loss_template == GLRMB200_LOSS_QUAD ? GLRMB200_LOSS_QUAD : 0;
  DenseArgs P = dense_args(E, true, E->d_Y, flags, min_stepsize, honour_stop);
  const int local_blocks = (D.g1 - D.g0) * D.bg;
  int rc;
  if (D.mma) DN_OK(dense_mma_launch_y(D.nt, loss, mode, P, local_blocks, D.max_chunks, E->stream));
  else DN_OK(dense_launch_y_pass(D.kt, loss, mode, P, local_blocks, D.max_chunks, E->stream));
  DN_DBG(mode == 0 ? "dense_y_pass_kernel (gradient)" : "dense_y_pass_kernel (evaluation)");
  // fixed-order two-level reduction over the row blocks: this rank's groups, the other ranks' groups (all-gather), total
  const int64_t glen = E->d * (int64_t)E->stride;
  if (mode == 0) DN_OK(dense_launch_reduce_groups(D.d_gpart, D.bg, D.g0, D.g1, glen, D.d_gsum_G, nullptr, P.stop, E->stream));
  DN_OK(dense_launch_reduce_groups(D.d_objpart, D.bg, D.g0, D.g1, E->n, D.d_gsum_o, nullptr, P.stop, E->stream));
  if (mode == 0 && (rc = dense_gather_groups(E, D.d_gsum_G, glen))) return rc;
  if ((rc = dense_gather_groups(E, D.d_gsum_o, E->n))) return rc;
  if (mode == 0) DN_OK(dense_launch_reduce_total(D.d_gsum_G, glen, D.d_G, nullptr, P.stop, E->stream));
  DN_OK(dense_launch_reduce_total(D.d_gsum_o, E->n, D.d_colobj, nullptr, P.stop, E->stream));
  DN_DBG("dense_reduce kernels");
  DenseYState Q = dense_ystate(E, flags, min_stepsize, honour_stop);
  DN_OK(dense_launch_begin(E->tile_g, E->tile_r, Q, E->stream));
  DN_DBG("dense_y_begin_kernel");
  *launches += mode == 0 ? 6 : 4;
  return 0;
}

// the Y sweep (proxgrad.jl:160-203): gradient pass, then line-search rounds over the features still searching.  Every round
// ends with the plan of the next one (compacted feature list + its size, published to mapped host memory).  The host runs
// one round ahead of the device: it enqueues round r, then waits for the plan of round r (complete when round r-1 is) and
// stops if nothing is left — the round it enqueued in vain returns at once.
static int dense_sweep_y(glrmb200_engine* E, double min_stepsize, bool honour_stop, int64_t* launches) {
  DenseHost& D = E->dn;
  const int loss = E->loss_template == GLRMB200_LOSS_QUAD ? GLRMB200_LOSS_QUAD : 0;
  int rc;
  D.seq = (D.seq + 1) & 0x3ffff;
  auto key = [&](int round) { return (int32_t)((D.seq << 12) | round); };
  if ((rc = dense_eval_cols(E, 0, min_stepsize, honour_stop, 0, launches))) return rc;
  DenseYState Q = dense_ystate(E, 0, min_stepsize, honour_stop);
  DenseArgs P = dense_args(E, false, D.d_Ynew, 0, min_stepsize, honour_stop);
  Q.seq = key(0);
  DN_OK(dense_launch_plan(Q, E->stream));
  ++*launches;
  for (int round = 0;; ++round) {
    if (round >= 4095) return fail(GLRMB200_E_STATE, "dense Y line search did not terminate");
    DN_OK(dense_launch_step(E->tile_g, E->tile_r, Q, E->stream));
    DN_DBG("dense_y_step_kernel");
    if (D.mma) DN_OK(dense_mma_launch_y(D.nt, loss, 1, P, (D.g1 - D.g0) * D.bg, D.max_chunks, E->stream));
    else DN_OK(dense_launch_y_pass(D.kt, loss, 1, P, (D.g1 - D.g0) * D.bg, D.max_chunks, E->stream));
    DN_DBG("dense_y_pass_kernel (trial)");
    DN_OK(dense_launch_reduce_groups(D.d_objpart, D.bg, D.g0, D.g1, E->n, D.d_gsum_o, D.d_nactive, P.stop, E->stream));
    if ((rc = dense_gather_groups(E, D.d_gsum_o, E->n))) return rc;
    DN_OK(dense_launch_reduce_total(D.d_gsum_o, E->n, D.d_colobj, D.d_nactive, P.stop, E->stream));
    DN_OK(dense_launch_decide(Q, E->stream));
    DN_DBG("dense_y_decide_kernel");
    Q.seq = key(round + 1);
    DN_OK(dense_launch_plan(Q, E->stream));
    *launches += 6;
    if (round == 0) continue;
    // plan(round) has been published?  (slot round % 4 of the ring; at most the plans of rounds round .. round + 2 are
    // outstanding, so the slot cannot have been overwritten)
    volatile int32_t* slot = D.h_nactive + 2 * (round & 3);
    const auto t0 = std::chrono::steady_clock::now();
    for (int spin = 0;; ++spin) {
      if (slot[1] == key(round)) break;
      if ((spin & 1023) == 1023) {
        const cudaError_t q = cudaStreamQuery(E->stream);
        if (q == cudaSuccess) break;                                       // everything enqueued has run: the plan is there
        if (q != cudaErrorNotReady) return fail(GLRMB200_E_CUDA, "dense Y sweep: %s", cudaGetErrorString(q));
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > dense_wait_limit_s())
          return fail(GLRMB200_E_STATE, "dense Y line search: the device did not publish its plan within %.0f s", dense_wait_limit_s());
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (slot[0] == 0) break;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
extern "C" int glrmb200_version(void) { return GLRMB200_VERSION; }
extern "C" const char* glrmb200_last_error(void) { return g_err; }
extern "C" int glrmb200_device_count(int32_t* count) {
  if (count) *count = 0;
  int rc = check_device();
  if (rc) return rc;
  if (count) *count = g_device_checked;
  return 0;
}

static void free_lists(glrmb200_engine* E, Side& S) {
  dfree(S.d_ptr, E->stream); dfree(S.d_idx, E->stream); dfree(S.d_val, E->stream);
  dfree(S.d_order, E->stream); dfree(S.d_order_vec, E->stream);
}
static void free_side(glrmb200_engine* E, Side& S) {
  free_lists(E, S);
  dfree(S.d_reg_code, E->stream); dfree(S.d_reg_param, E->stream); dfree(S.d_alpha, E->stream); dfree(S.d_obj, E->stream);
  dfree(S.d_reg_payload_ptr, E->stream); dfree(S.d_reg_payload, E->stream);
}

// Process-level cache of the exchange allocation and its peer mappings.  cudaIpcCloseMemHandle + cudaFree of an
// exported allocation cost ~0.13 s per handle (measured at N=4): callers that re-create handles of the same shape
// (cross-validation folds, regularization paths, bench.py's end-to-end leg) get the allocation and the mappings of the
// previous handle back instead (same allocation => same IPC handle bytes => the peers' mappings stay valid too).
struct XchgCache {
  int device = -1, rank = -1, nranks = 0;
  size_t bytes = 0;
  double* d_xchg = nullptr;
  std::vector<uint8_t> blobs;
  std::vector<void*> opened;
  double** dpx = nullptr; double** dpy = nullptr; double** dpo = nullptr;
  unsigned long long** dpf = nullptr; int* dpr = nullptr;
};
static XchgCache g_xc;
static void xc_release(XchgCache& c) {
  if (!c.d_xchg) return;
  cudaSetDevice(c.device);
  for (void* p : c.opened) cudaIpcCloseMemHandle(p);
  cudaFree(c.dpx); cudaFree(c.dpy); cudaFree(c.dpo); cudaFree(c.dpf); cudaFree(c.dpr);
  cudaFree(c.d_xchg);
  c = XchgCache();
}

extern "C" int glrmb200_destroy(glrmb200_handle E) {
  if (!E) return 0;
  cudaSetDevice(E->device);
  if (E->stream) cudaStreamSynchronize(E->stream);
  // E->comm is the process-wide cached communicator (glrmb200_comm_init): not destroyed with the handle
  if (E->d_xchg) {                            // park the exchange allocation (with its peer mappings, if any)
    xc_release(g_xc);
    g_xc.device = E->device; g_xc.rank = E->rank; g_xc.nranks = E->nranks; g_xc.bytes = E->xchg_bytes;
    g_xc.d_xchg = E->d_xchg; g_xc.blobs = E->peer_blobs; g_xc.opened = E->opened;
    g_xc.dpx = E->d_peer_X; g_xc.dpy = E->d_peer_Y; g_xc.dpo = E->d_peer_objc; g_xc.dpf = E->d_peer_flags; g_xc.dpr = E->d_peer_rank;
    E->d_xchg = nullptr;
  }
  if (E->stream) {
    E->cols.d_obj = nullptr;   // lives inside d_xchg
    free_side(E, E->rows); free_side(E, E->cols);
    dfree(E->d_loss_code, E->stream); dfree(E->d_loss_param, E->stream); dfree(E->d_ystart, E->stream);
    dfree(E->d_scalars, E->stream); dfree(E->d_trials, E->stream); dfree(E->d_stop, E->stream);
    dfree(E->d_objs, E->stream); dfree(E->d_stage, E->stream); dfree(E->d_barrier, E->stream);
    dense_free(E);
  }
  for (auto& e : E->evpool) if (e) cudaEventDestroy(e);
  if (E->ev_base) cudaEventDestroy(E->ev_base);
  if (E->ev_aux) cudaEventDestroy(E->ev_aux);
  if (E->ev_fork) cudaEventDestroy(E->ev_fork);
  for (int i = 0; i < 3; ++i) {
    if (E->ev_join[i]) cudaEventDestroy(E->ev_join[i]);
    if (E->tier_stream[i]) cudaStreamDestroy(E->tier_stream[i]);
  }
  if (E->stream) cudaStreamDestroy(E->stream);   // work already enqueued (the frees) still completes
  pinned_put(E->h_pinned);                       // the stream was synchronised above: no kernel still writes the flags
  delete E;
  return 0;
}

// Observation lists (list mode): host checks, cost-balanced shards, upload of this rank's shard, device-side
// validation (index bounds, NaN, label domains), degree-sorted schedules.  Used by glrmb200_create and by
// glrmb200_set_obs (new lists on a live handle: cross-validation folds, src/cross_validate.jl:31-33).
// The four big copies are enqueued first; the host-side work (monotonicity checks, local ptr arrays, the two degree
// sorts) runs while they are in flight, and one synchronisation at the end hands the host buffers back.
static int load_lists(glrmb200_engine* E, const int64_t* row_ptr, const int32_t* row_idx, const double* row_val,
                      const int64_t* col_ptr, const int32_t* col_idx, const double* col_val) {
  const int64_t m = E->m, n = E->n;
  Side& R = E->rows;
  Side& C = E->cols;
  if (!row_ptr || !col_ptr) return fail(GLRMB200_E_INVALID, "observation lists missing");
  const int64_t nr = row_ptr[m], nc = col_ptr[n];
  if (row_ptr[0] != 0 || col_ptr[0] != 0) return fail(GLRMB200_E_INVALID, "ptr arrays must start at 0");
  if ((nr && (!row_idx || !row_val)) || (nc && (!col_idx || !col_val))) return fail(GLRMB200_E_INVALID, "observation arrays missing");
  for (int64_t e = 0; e < m; ++e)
    if (row_ptr[e + 1] < row_ptr[e]) return fail(GLRMB200_E_INVALID, "row_ptr not monotone");
  for (int64_t f = 0; f < n; ++f)
    if (col_ptr[f + 1] < col_ptr[f]) return fail(GLRMB200_E_INVALID, "col_ptr not monotone");
  E->obs_full = false;
  R.full_len = C.full_len = 0;
  E->nnz_rows_total = nr;
  R.bounds.assign((size_t)E->nranks + 1, 0);
  C.bounds.assign((size_t)E->nranks + 1, 0);
  glrmb200_plan_shards(row_ptr, m, E->nranks, R.bounds.data());
  glrmb200_plan_shards(col_ptr, n, E->nranks, C.bounds.data());
  R.begin = R.bounds[E->rank]; R.end = R.bounds[E->rank + 1];
  C.begin = C.bounds[E->rank]; C.end = C.bounds[E->rank + 1];
  free_lists(E, R);
  free_lists(E, C);
  int rc;
  // 1. the big streams first (asynchronous from pinned memory)
  auto up_lists = [&](Side& S, const int64_t* ptr, const int32_t* idx, const double* val) -> int {
    const int64_t q0 = ptr[S.begin], q1 = ptr[S.end];
    S.nnz_local = q1 - q0;
    int r2;
    if ((r2 = upload(&S.d_idx, idx ? idx + q0 : nullptr, (size_t)S.nnz_local, E->stream, 32))) return r2;
    return upload(&S.d_val, val ? val + q0 : nullptr, (size_t)S.nnz_local, E->stream, 32);
  };
  if ((rc = up_lists(R, row_ptr, row_idx, row_val))) return rc;
  if ((rc = up_lists(C, col_ptr, col_idx, col_val))) return rc;
  // 2. host work under the copies: shard-local ptr arrays and the degree-sorted schedules
  auto up_sched = [&](Side& S, const int64_t* ptr, const std::vector<char>* vecflags) -> int {
    const int64_t cnt = S.end - S.begin, q0 = ptr[S.begin];
    std::vector<int64_t> local((size_t)cnt + 1);
    for (int64_t i = 0; i <= cnt; ++i) local[(size_t)i] = ptr[S.begin + i] - q0;
    int r2 = upload(&S.d_ptr, local.data(), local.size(), E->stream);
    if (r2) return r2;
    return build_schedule(E, S, ptr, E->has_vec ? vecflags : nullptr);
  };
  if ((rc = up_sched(R, row_ptr, &E->all_rows_vec))) return rc;
  if ((rc = up_sched(C, col_ptr, &E->col_is_vec))) return rc;
  // 3. validation on the device
  unsigned long long* d_bad = nullptr;
  if ((rc = dalloc(&d_bad, 2, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(d_bad, 0xff, 2 * sizeof(unsigned long long), E->stream));
  if (R.nnz_local > 0)
    validate_rows_kernel<<<(unsigned)((R.nnz_local + 255) / 256), 256, 0, E->stream>>>(R.d_idx, R.d_val, R.nnz_local, n, E->d_loss_code, E->d_loss_param, d_bad);
  if (C.end > C.begin)
    validate_cols_kernel<<<(unsigned)((C.end - C.begin + 7) / 8), 256, 0, E->stream>>>(C.d_ptr, C.d_idx, C.d_val, C.end - C.begin, C.begin, m, E->d_loss_code, E->d_loss_param, d_bad + 1);
  CUDA_OK(cudaGetLastError());
  unsigned long long bad[2] = {0, 0};
  CUDA_OK(cudaMemcpyAsync(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  dfree(d_bad, E->stream);
  for (int side = 0; side < 2; ++side) {
    if (bad[side] == ~0ULL) continue;
    const int kind = -(int)(bad[side] & 15ULL);
    const Side& S = side == 0 ? R : C;
    const int64_t* ptr = side == 0 ? row_ptr : col_ptr;
    const int64_t q = ptr[S.begin] + (int64_t)(bad[side] >> 4);
    const int64_t unit = (std::upper_bound(ptr, ptr + (side == 0 ? m : n) + 1, q) - ptr) - 1;
    const int64_t other = side == 0 ? row_idx[q] : col_idx[q];
    const int64_t e = side == 0 ? unit : other, f = side == 0 ? other : unit;
    const double v = side == 0 ? row_val[q] : col_val[q];
    if (kind == GLRMB200_E_INVALID) return fail(kind, "%s[%lld] = %lld out of range", side == 0 ? "row_idx" : "col_idx", (long long)q, (long long)other);
    if (kind == GLRMB200_E_NAN) return fail(kind, "Observed value in entry (%lld, %lld) is NaN.", (long long)e + 1, (long long)f + 1);
    return fail(kind, "entry (%lld, %lld): label %g is outside the domain of loss code %d", (long long)e + 1, (long long)f + 1, v, (int)E->h_loss_code[(size_t)f]);
  }
  return 0;
}

static int create_impl(glrmb200_engine* E, const glrmb200_problem* P, int flags) {
  const int64_t m = P->m, n = P->n, k = P->k;
  if (m <= 0 || n <= 0 || k <= 0) return fail(GLRMB200_E_INVALID, "m, n, k must be positive");
  if (m >= INT32_MAX || n >= INT32_MAX) return fail(GLRMB200_E_INVALID, "m, n must fit int32");
  if (E->nranks > MAX_RANKS) return fail(GLRMB200_E_INVALID, "at most %d ranks", MAX_RANKS);
  if (!P->loss_code || !P->loss_param || !P->rx_code || !P->ry_code) return fail(GLRMB200_E_INVALID, "null descriptor table");
  E->m = m; E->n = n; E->k = k; E->d = P->d;
  E->obs_full = P->obs_full != 0;

  // ---- losses: embedding dims must sum to d (get_yidxs, losses.jl:76-93) ----------------------
  int64_t dsum = 0;
  bool uniform = true;
  E->ystart.assign((size_t)n + 1, 0);
  E->col_is_vec.assign((size_t)n, 0);
  E->h_loss_code.assign(P->loss_code, P->loss_code + n);
  std::vector<char>& col_is_vec = E->col_is_vec;
  for (int64_t f = 0; f < n; ++f) {
    const int code = P->loss_code[f];
    const double* p = P->loss_param + f * GLRMB200_LOSS_NPARAM;
    if (code < GLRMB200_LOSS_QUAD || code > GLRMB200_LOSS_MULTINOMIAL_ORDINAL)
      return fail(GLRMB200_E_UNSUPPORTED, "loss code %d (column %lld) is unknown", code, (long long)f);
    const int dim = loss_dim(code, p);
    if (code >= GLRMB200_LOSS_MULTINOMIAL) {
      E->has_vec = true;
      col_is_vec[(size_t)f] = 1;
      if (dim < 1 || dim > VEC_DMAX)
        return fail(GLRMB200_E_UNSUPPORTED, "column %lld: embedding dimension %d is outside 1..%d supported on the device", (long long)f, dim, VEC_DMAX);
      if ((code == GLRMB200_LOSS_OVA || code == GLRMB200_LOSS_BVS) &&
          ((int)p[3] < GLRMB200_LOSS_QUAD || (int)p[3] > GLRMB200_LOSS_WEIGHTED_HINGE))
        return fail(GLRMB200_E_UNSUPPORTED, "column %lld: bin_loss code %d must be a scalar loss", (long long)f, (int)p[3]);
    }
    E->ystart[(size_t)f] = dsum;
    dsum += dim;
    if (code != P->loss_code[0] || memcmp(p, P->loss_param, 3 * sizeof(double)) != 0) uniform = false;
  }
  E->ystart[(size_t)n] = dsum;
  for (int64_t f = 0; f < n; ++f) {          // OrdinalReg / MNLOrdinalReg columns take the block path whatever their width
    if (P->ry_code[P->ry_count == 1 ? 0 : f] & (GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL)) { col_is_vec[(size_t)f] = 1; E->has_vec = true; }
  }
  if (dsum != P->d) return fail(GLRMB200_E_INVALID, "d = %lld but the losses' embedding dimensions sum to %lld (proxgrad.jl:55-63)", (long long)P->d, (long long)dsum);
  E->loss_template = 0;
  if (uniform && !E->has_vec && (P->loss_code[0] == GLRMB200_LOSS_QUAD || P->loss_code[0] == GLRMB200_LOSS_LOGISTIC)) {
    E->loss_template = P->loss_code[0];
    E->uparam[0] = P->loss_param[0]; E->uparam[1] = P->loss_param[1]; E->uparam[2] = P->loss_param[2];
  }

  // ---- tile selection ---------------------------------------------------------------------------
  // tile = lane-group width G x slots per lane R; a factor column is stored as 2*G*R doubles (zero past k)
  // so that every gather is R unconditional 16-byte loads per lane with immediate offsets.  Pick the
  // narrowest group that holds k in at most 4 slots per lane (fewest reduction shuffles, least padding).
  E->kp = (int)((k + 1) / 2 * 2);
  if (E->kp > 256) return fail(GLRMB200_E_UNSUPPORTED, "k = %lld: ranks above 256 are not supported", (long long)k);
  if (E->kp <= 8) { E->tile_g = 4; E->tile_r = 1; }
  else if (E->kp <= 64) { E->tile_g = 8; E->tile_r = (E->kp + 15) / 16; }
  else if (E->kp <= 128) { E->tile_g = 16; E->tile_r = (E->kp + 31) / 32; }
  else { E->tile_g = 32; E->tile_r = (E->kp + 63) / 64; }
  if (const char* t = getenv("GLRMB200_TILE")) {   // tuning hook: "G,R"
    int g = 0, r = 0;
    if (sscanf(t, "%d,%d", &g, &r) == 2 && 2 * g * r >= E->kp) {
      for (const Tile& tl : kTiles) if (tl.g == g && tl.r == r) { E->tile_g = g; E->tile_r = r; }
    }
  }
  E->stride = 2 * E->tile_g * E->tile_r;
  const bool want_dense = dense_eligible(E, P) && !(flags & GLRMB200_CREATE_GATHER_ONLY);
  if (E->has_vec && E->tile_r > 2 && !want_dense)
    return fail(GLRMB200_E_UNSUPPORTED, "vector-valued losses are supported on the device for k <= 32 this round (k = %lld)", (long long)k);
  if (const char* t = getenv("GLRMB200_HEAVY")) E->heavy_threshold = std::max<long long>(1, atoll(t));
  if (const char* t = getenv("GLRMB200_CLUSTER")) E->cluster_threshold = std::max<long long>(1, atoll(t));
  if (const char* t = getenv("GLRMB200_CLUSTER16")) E->cluster16_threshold = std::max<long long>(1, atoll(t));
  if (E->cluster_threshold < E->heavy_threshold) E->cluster_threshold = E->heavy_threshold;
  if (E->cluster16_threshold < E->cluster_threshold) E->cluster16_threshold = E->cluster_threshold;

  Side& R = E->rows;
  Side& C = E->cols;
  R.units = m; C.units = n;
  R.bounds.assign((size_t)E->nranks + 1, 0);
  C.bounds.assign((size_t)E->nranks + 1, 0);
  if (E->obs_full) {
    if (!P->dense_A) return fail(GLRMB200_E_INVALID, "obs_full needs dense_A");
    R.full_len = n; C.full_len = m;
    E->nnz_rows_total = m * n;
    glrmb200_plan_shards(nullptr, m, E->nranks, R.bounds.data());
    glrmb200_plan_shards(nullptr, n, E->nranks, C.bounds.data());
  }   // (observation lists: checked, sharded and uploaded by load_lists below)
  if (E->obs_full && want_dense && E->nranks > 1) {
    // mode B: rank r owns the row-block groups [r * G / N, (r + 1) * G / N); every rank updates all of Y
    dense_row_bounds(m, E->nranks, R.bounds.data());
    C.bounds[0] = 0;
    for (int r = 1; r <= E->nranks; ++r) C.bounds[(size_t)r] = n;
  }
  R.begin = R.bounds[E->rank]; R.end = R.bounds[E->rank + 1];
  C.begin = C.bounds[E->rank]; C.end = C.bounds[E->rank + 1];

  // ---- device ---------------------------------------------------------------------------------------
  int rc = check_device();
  if (rc) return rc;
  if (E->device < 0 || E->device >= g_device_checked) return fail(GLRMB200_E_INVALID, "device %d out of range (%d visible)", E->device, g_device_checked);
  CUDA_OK(cudaSetDevice(E->device));
  if ((rc = pool_setup(E->device))) return rc;
  if ((rc = pinned_get(&E->h_pinned))) return rc;
  E->h_stop = reinterpret_cast<volatile int*>(E->h_pinned + 8);
  {
    // stream priorities (lower number = higher priority): cluster tiers above the CTA tier above the warp tier, so the
    // units on a sweep's critical path get their SMs first
    int lo = 0, hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const bool prio = !(getenv("GLRMB200_NO_PRIORITY") && atoi(getenv("GLRMB200_NO_PRIORITY"))) && hi < lo;
    const int p_cluster = prio ? hi : 0, p_cta = prio ? std::min(lo, hi + 1) : 0, p_warp = prio ? std::min(lo, hi + 2) : 0;
    CUDA_OK(cudaStreamCreateWithPriority(&E->stream, cudaStreamNonBlocking, p_cta));
    CUDA_OK(cudaStreamCreateWithPriority(&E->tier_stream[0], cudaStreamNonBlocking, p_cluster));
    CUDA_OK(cudaStreamCreateWithPriority(&E->tier_stream[1], cudaStreamNonBlocking, p_cluster));
    CUDA_OK(cudaStreamCreateWithPriority(&E->tier_stream[2], cudaStreamNonBlocking, p_warp));
  }
  CUDA_OK(cudaEventCreateWithFlags(&E->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 3; ++i) CUDA_OK(cudaEventCreateWithFlags(&E->ev_join[i], cudaEventDisableTiming));
  CUDA_OK(cudaEventCreate(&E->ev_base));
  CUDA_OK(cudaEventCreate(&E->ev_aux));
  if ((rc = dalloc(&E->d_scalars, 4, E->stream))) return rc;
  if ((rc = dalloc(&E->d_trials, 2, E->stream))) return rc;
  if ((rc = dalloc(&E->d_stop, 2, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(E->d_trials, 0, 2 * sizeof(unsigned long long), E->stream));
  CUDA_OK(cudaMemsetAsync(E->d_stop, 0, 2 * sizeof(int), E->stream));

  if ((rc = upload(&E->d_loss_code, P->loss_code, (size_t)n, E->stream))) return rc;
  if ((rc = upload(&E->d_loss_param, P->loss_param, (size_t)n * GLRMB200_LOSS_NPARAM, E->stream))) return rc;
  if ((rc = setup_regs(E, R, P->rx_count, P->rx_code, P->rx_param, false, P->rx_payload_ptr, P->rx_payload))) return rc;
  if ((rc = setup_regs(E, C, P->ry_count, P->ry_code, P->ry_param, true, P->ry_payload_ptr, P->ry_payload))) return rc;
  if (E->has_vec && (R.d_reg_payload_ptr || C.d_reg_payload_ptr))
    return fail(GLRMB200_E_UNSUPPORTED, "fixed_latent_features / RemQuadReg together with vector-valued losses have no device implementation");
  E->all_rows_vec.assign(E->has_vec ? (size_t)m : 0, 1);   // with block columns every row takes the vector path
  std::vector<char>& all_rows_vec = E->all_rows_vec;
  if (E->has_vec || want_dense) {
    if ((rc = upload(&E->d_ystart, E->ystart.data(), E->ystart.size(), E->stream))) return rc;
  }
  if (E->has_vec) {
    for (int64_t f = 0; f < n; ++f) {
      if (!col_is_vec[(size_t)f]) continue;
      const int rcf = P->ry_code[P->ry_count == 1 ? 0 : f];
      const int base = rcf & GLRMB200_REG_BASE_MASK;
      if (rcf & (GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL)) continue;   // inner prox acts on one (k-1)-vector: any base
      if (!(base == GLRMB200_REG_ZERO || base == GLRMB200_REG_QUAD || base == GLRMB200_REG_ONE ||
            base == GLRMB200_REG_NONNEG || base == GLRMB200_REG_NONNEG_ONE))
        return fail(GLRMB200_E_UNSUPPORTED, "column %lld: regularizer code %d on a block column (only element-wise regularizers decompose over the k x d_f block)", (long long)f, base);
    }
  }

  if (E->obs_full && want_dense) {
    // dense kernels: A stays exactly as Julia stores it (one copy, no transposed twin, no schedules)
    if ((rc = dense_setup(E, P))) return rc;
    R.nnz_local = C.nnz_local = m * n;
    unsigned long long* d_bad = nullptr;
    if ((rc = dalloc(&d_bad, 1, E->stream))) return rc;
    CUDA_OK(cudaMemsetAsync(d_bad, 0xff, sizeof(unsigned long long), E->stream));
    const int64_t mloc = E->dn.row1 - E->dn.row0;   // this rank's rows (all of them on one GPU)
    if (mloc > 0) validate_dense_kernel<<<(unsigned)((E->dn.lda * n + 255) / 256), 256, 0, E->stream>>>(E->dn.d_A, E->dn.lda * n, mloc, E->dn.lda, E->d_loss_code, E->d_loss_param, d_bad);
    CUDA_OK(cudaGetLastError());
    unsigned long long bad = 0;
    CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
    CUDA_OK(cudaStreamSynchronize(E->stream));
    dfree(d_bad, E->stream);
    if (bad != ~0ULL) {
      const int kind = -(int)(bad & 15ULL);
      const int64_t pos = (int64_t)(bad >> 4);
      const int64_t f = pos / mloc, e = E->dn.row0 + pos % mloc;
      if (kind == GLRMB200_E_NAN) return fail(kind, "Observed value in entry (%lld, %lld) is NaN.", (long long)e + 1, (long long)f + 1);
      return fail(kind, "entry (%lld, %lld): label %g is outside the domain of loss code %d", (long long)e + 1, (long long)f + 1, P->dense_A[f * m + e], P->loss_code[f]);
    }
  } else if (E->obs_full) {
    // column side streams the Julia (column-major) A as is; the row side gets a row-major copy
    const int64_t cb = C.begin, ce = C.end;
    C.nnz_local = (ce - cb) * m;
    if ((rc = upload(&C.d_val, P->dense_A + cb * m, (size_t)C.nnz_local, E->stream, 32))) return rc;
    // row-major copy of rows [R.begin, R.end): transpose on the device from a temporary full upload
    R.nnz_local = (R.end - R.begin) * n;
    double* d_full = nullptr;
    if (E->nranks == 1) d_full = C.d_val;
    else if ((rc = upload(&d_full, P->dense_A, (size_t)(m * n), E->stream))) return rc;
    double* d_rowmajor = nullptr;
    if ((rc = dalloc(&d_rowmajor, (size_t)(m * n) + 32, E->stream))) return rc;
    CUDA_OK(cudaMemsetAsync(d_rowmajor + m * n, 0, 32 * sizeof(double), E->stream));
    dim3 blk(32, 8), grd((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
    transpose_kernel<<<grd, blk, 0, E->stream>>>(d_full, d_rowmajor, n, m);   // src = n x m row-major view of A
    CUDA_OK(cudaGetLastError());
    if (E->nranks == 1) {
      R.d_val = d_rowmajor;
    } else {
      if ((rc = dalloc(&R.d_val, (size_t)R.nnz_local + 32, E->stream))) return rc;
      CUDA_OK(cudaMemsetAsync(R.d_val + R.nnz_local, 0, 32 * sizeof(double), E->stream));
      CUDA_OK(cudaMemcpyAsync(R.d_val, d_rowmajor + R.begin * n, (size_t)R.nnz_local * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));
      dfree(d_rowmajor, E->stream);
      dfree(d_full, E->stream);
    }
    if ((rc = build_schedule(E, R, nullptr, E->has_vec ? &all_rows_vec : nullptr))) return rc;
    if ((rc = build_schedule(E, C, nullptr, E->has_vec ? &col_is_vec : nullptr))) return rc;
    {
      unsigned long long* d_bad = nullptr;
      if ((rc = dalloc(&d_bad, 1, E->stream))) return rc;
      CUDA_OK(cudaMemsetAsync(d_bad, 0xff, sizeof(unsigned long long), E->stream));
      const int64_t total = C.nnz_local;
      if (total > 0) validate_dense_kernel<<<(unsigned)((total + 255) / 256), 256, 0, E->stream>>>(C.d_val, total, m, m, E->d_loss_code + C.begin, E->d_loss_param + C.begin * GLRMB200_LOSS_NPARAM, d_bad);
      CUDA_OK(cudaGetLastError());
      unsigned long long bad = 0;
      CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
      CUDA_OK(cudaStreamSynchronize(E->stream));
      dfree(d_bad, E->stream);
      if (bad != ~0ULL) {
        const int kind = -(int)(bad & 15ULL);
        const int64_t pos = (int64_t)(bad >> 4);
        const int64_t f = C.begin + pos / m, e = pos % m;
        if (kind == GLRMB200_E_NAN) return fail(kind, "Observed value in entry (%lld, %lld) is NaN.", (long long)e + 1, (long long)f + 1);
        return fail(kind, "entry (%lld, %lld): label %g is outside the domain of loss code %d", (long long)e + 1, (long long)f + 1, P->dense_A[f * m + e], P->loss_code[f]);
      }
    }
  } else {
    if ((rc = load_lists(E, P->row_ptr, P->row_idx, P->row_val, P->col_ptr, P->col_idx, P->col_val))) return rc;
  }

  {
    // factors, obj_by_col and the peer flags share one allocation rounded to the 2 MiB allocation granule, so that a
    // single CUDA IPC handle maps exactly this memory in the peers (small cudaMallocs share granules; handles would alias)
    const size_t nx = (size_t)m * E->stride, ny = (size_t)E->d * E->stride, no = ((size_t)n + 1) / 2 * 2, nf = MAX_RANKS;
    const size_t granule = (size_t)2 << 20;
    E->xchg_bytes = ((nx + ny + no + nf) * sizeof(double) + granule - 1) / granule * granule;
    bool fresh = true;
    if (g_xc.d_xchg && g_xc.device == E->device && g_xc.rank == E->rank && g_xc.nranks == E->nranks && g_xc.bytes == E->xchg_bytes) {
      E->d_xchg = g_xc.d_xchg; E->peer_blobs = g_xc.blobs; E->opened = g_xc.opened;
      E->d_peer_X = g_xc.dpx; E->d_peer_Y = g_xc.dpy; E->d_peer_objc = g_xc.dpo; E->d_peer_flags = g_xc.dpf; E->d_peer_rank = g_xc.dpr;
      g_xc = XchgCache();
      fresh = false;
    } else {
      CUDA_OK(cudaMalloc((void**)&E->d_xchg, E->xchg_bytes));
    }
    E->d_X = E->d_xchg;
    E->d_Y = E->d_X + nx;
    C.d_obj = E->d_Y + ny;
    E->d_flags = reinterpret_cast<unsigned long long*>(C.d_obj + no);
    // a recycled allocation keeps its content: factors are fully rewritten by upload_factors (padding included) and
    // obj_by_col by the first sweep; the flags are zeroed by glrmb200_ipc_export before any peer can write them
    if (fresh) CUDA_OK(cudaMemsetAsync(E->d_xchg, 0, E->xchg_bytes, E->stream));
  }
  if ((rc = dalloc(&R.d_alpha, (size_t)m, E->stream))) return rc;
  if ((rc = dalloc(&C.d_alpha, (size_t)n, E->stream))) return rc;
  if ((rc = dalloc(&R.d_obj, (size_t)m, E->stream))) return rc;
  CUDA_OK(cudaMemsetAsync(R.d_obj, 0, (size_t)m * sizeof(double), E->stream));
  CUDA_OK(cudaMemsetAsync(R.d_alpha, 0, (size_t)m * sizeof(double), E->stream));
  CUDA_OK(cudaMemsetAsync(C.d_alpha, 0, (size_t)n * sizeof(double), E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  return 0;
}

extern "C" int glrmb200_create(glrmb200_handle* out, const glrmb200_problem* problem, int32_t device,
                               int32_t rank, int32_t nranks) {
  return glrmb200_create_ex(out, problem, device, rank, nranks, 0);
}

extern "C" int glrmb200_create_ex(glrmb200_handle* out, const glrmb200_problem* problem, int32_t device,
                                  int32_t rank, int32_t nranks, int32_t flags) {
  if (!out || !problem) return fail(GLRMB200_E_INVALID, "null argument");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GLRMB200_E_INVALID, "bad rank %d / nranks %d", rank, nranks);
  glrmb200_engine* E = new glrmb200_engine();
  E->device = device; E->rank = rank; E->nranks = nranks;
  const int rc = create_impl(E, problem, flags);
  if (rc) { glrmb200_destroy(E); return rc; }
  *out = E;
  return 0;
}

extern "C" int glrmb200_set_obs(glrmb200_handle E, const int64_t* row_ptr, const int32_t* row_idx, const double* row_val,
                                const int64_t* col_ptr, const int32_t* col_idx, const double* col_val) {
  if (!E) return fail(GLRMB200_E_STATE, "null handle");
  CUDA_OK(cudaSetDevice(E->device));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  if (E->dn.on) {
    if (E->has_vec && E->tile_r > 2)
      return fail(GLRMB200_E_UNSUPPORTED, "observation lists with vector-valued losses need k <= 32 (k = %lld)", (long long)E->k);
    dense_free(E);                             // a fully observed handle becomes list mode
  }
  return load_lists(E, row_ptr, row_idx, row_val, col_ptr, col_idx, col_val);
}

extern "C" int glrmb200_comm_unique_id(uint8_t id[128]) {
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t u;
  NCCL_OK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return 0;
}

// The communicator is created once per process and cached ("NCCL communicators are created
// lazily and cached per process"): a second handle of the same (rank, nranks, device) may pass id == NULL to reuse it,
// so re-fitting callers (cross-validation folds, bench.py's end-to-end leg) do not pay ncclCommInitRank again.
static ncclComm_t g_comm = nullptr;
static int g_comm_rank = -1, g_comm_nranks = 0, g_comm_device = -1;

extern "C" int glrmb200_comm_init(glrmb200_handle E, const uint8_t id[128]) {
  if (!E) return fail(GLRMB200_E_STATE, "null handle");
  if (E->nranks == 1) return 0;
  int rc = load_nccl();
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(E->device));
  if (!id) {
    if (!g_comm || g_comm_rank != E->rank || g_comm_nranks != E->nranks || g_comm_device != E->device)
      return fail(GLRMB200_E_STATE, "no cached communicator for rank %d/%d on device %d", E->rank, E->nranks, E->device);
    E->comm = g_comm;
    return 0;
  }
  ncclUniqueId_t u;
  memcpy(u.internal, id, 128);
  ncclComm_t c = nullptr;
  NCCL_OK(g_nccl.CommInitRank(&c, E->nranks, u, E->rank));
  if (g_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comm);
  g_comm = c; g_comm_rank = E->rank; g_comm_nranks = E->nranks; g_comm_device = E->device;
  E->comm = c;
  return 0;
}

extern "C" int glrmb200_shard(glrmb200_handle E, int64_t* rb, int64_t* re, int64_t* cb, int64_t* ce) {
  if (!E) return fail(GLRMB200_E_STATE, "null handle");
  if (rb) *rb = E->rows.begin;
  if (re) *re = E->rows.end;
  if (cb) *cb = E->cols.begin;
  if (ce) *ce = E->cols.end;
  return 0;
}

// Barrier between the ranks on the engine's stream: with the fused exchange a flag exchange over peer memory
// (peer_barrier_kernel), otherwise a 1-element NCCL all-reduce.
static int comm_barrier(glrmb200_engine* E) {
  if (E->nranks == 1) return 0;
  static const bool no_exchange = getenv("GLRMB200_NO_EXCHANGE") && atoi(getenv("GLRMB200_NO_EXCHANGE"));
  if (no_exchange) return 0;      // tuning hook (tools/tune.py): time one shard of an N-rank fit on a single GPU; results are wrong
  if (E->peer_ready) {
    ++E->epoch;
    peer_barrier_kernel<<<1, 64, 0, E->stream>>>(E->d_peer_flags, E->d_flags, E->d_peer_rank, E->rank, E->nranks - 1, E->epoch, E->d_stop + 1);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (!E->comm) return fail(GLRMB200_E_STATE, "glrmb200_comm_init was not called");
  if (!E->d_barrier) {
    int rc = dalloc(&E->d_barrier, 1, E->stream);
    if (rc) return rc;
    CUDA_OK(cudaMemsetAsync(E->d_barrier, 0, sizeof(double), E->stream));
  }
  NCCL_OK(g_nccl.AllReduce(E->d_barrier, E->d_barrier, 1, kNcclDouble, /*ncclSum*/ 0, E->comm, E->stream));
  return 0;
}
static int check_barrier_timeout(glrmb200_engine* E) {     // after a synchronisation
  if (!E->peer_ready) return 0;
  int flag = 0;
  CUDA_OK(cudaMemcpy(&flag, E->d_stop + 1, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) return fail(GLRMB200_E_NCCL, "peer barrier timed out: a rank of the fused exchange stopped responding");
  return 0;
}

extern "C" int glrmb200_comm_barrier(glrmb200_handle E) {
  if (!E) return fail(GLRMB200_E_STATE, "null handle");
  CUDA_OK(cudaSetDevice(E->device));
  int rc = comm_barrier(E);
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(E->stream));
  return check_barrier_timeout(E);
}

extern "C" int glrmb200_ipc_export(glrmb200_handle E, uint8_t out[GLRMB200_IPC_BYTES]) {
  if (!E || !out) return fail(GLRMB200_E_INVALID, "null argument");
  CUDA_OK(cudaSetDevice(E->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  // the host all-gathers the blobs between export and open: zeroing the flag array here means every rank's flags are
  // clean before any peer can publish an epoch into them
  CUDA_OK(cudaMemsetAsync(E->d_flags, 0, MAX_RANKS * sizeof(unsigned long long), E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  E->epoch = 0;
  cudaIpcMemHandle_t h;
  CUDA_OK(cudaIpcGetMemHandle(&h, E->d_xchg));
  memcpy(out, &h, sizeof(h));
  return 0;
}

extern "C" int glrmb200_ipc_open(glrmb200_handle E, const uint8_t* blobs) {
  if (!E || !blobs) return fail(GLRMB200_E_INVALID, "null argument");
  if (E->nranks == 1) return 0;
  if (E->has_vec) return 0;   // block columns keep the NCCL exchange this round
  if (E->dn.on) return 0;     // fully observed path: no factor exchange at all (rows sharded, Y replicated)
  CUDA_OK(cudaSetDevice(E->device));
  const size_t nb = (size_t)E->nranks * GLRMB200_IPC_BYTES;
  if (!E->opened.empty() && E->peer_blobs.size() == nb && memcmp(E->peer_blobs.data(), blobs, nb) == 0) {
    E->peer_ready = true;       // every peer re-used its allocation too: the cached mappings are still the right ones
    return 0;
  }
  for (void* p : E->opened) cudaIpcCloseMemHandle(p);
  E->opened.clear();
  cudaFree(E->d_peer_X); cudaFree(E->d_peer_Y); cudaFree(E->d_peer_objc); cudaFree(E->d_peer_flags); cudaFree(E->d_peer_rank);
  E->d_peer_X = E->d_peer_Y = E->d_peer_objc = nullptr; E->d_peer_flags = nullptr; E->d_peer_rank = nullptr;
  E->peer_blobs.assign(blobs, blobs + nb);
  std::vector<double*> px, py, po;
  std::vector<unsigned long long*> pf;
  std::vector<int> pr;
  for (int r = 0; r < E->nranks; ++r) {
    if (r == E->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, blobs + (size_t)r * GLRMB200_IPC_BYTES, sizeof(h));
    void* base = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    E->opened.push_back(base);
    double* bx = (double*)base;                       // same layout in every rank: [X | Y | obj_by_col | flags]
    px.push_back(bx);
    py.push_back(bx + (E->d_Y - E->d_xchg));
    po.push_back(bx + (E->cols.d_obj - E->d_xchg));
    pf.push_back(reinterpret_cast<unsigned long long*>(bx + (reinterpret_cast<double*>(E->d_flags) - E->d_xchg)));
    pr.push_back(r);
  }
  // these tables live in cudaMalloc memory (not the pool): they travel with the cached exchange allocation
  auto up = [&](auto** dst, const auto* src, size_t cnt) -> int {
    CUDA_OK(cudaMalloc((void**)dst, cnt * sizeof(**dst)));
    CUDA_OK(cudaMemcpy(*dst, src, cnt * sizeof(**dst), cudaMemcpyHostToDevice));
    return 0;
  };
  int rc;
  if ((rc = up(&E->d_peer_X, px.data(), px.size()))) return rc;
  if ((rc = up(&E->d_peer_Y, py.data(), py.size()))) return rc;
  if ((rc = up(&E->d_peer_objc, po.data(), po.size()))) return rc;
  if ((rc = up(&E->d_peer_flags, pf.data(), pf.size()))) return rc;
  if ((rc = up(&E->d_peer_rank, pr.data(), pr.size()))) return rc;
  E->peer_ready = true;
  return 0;
}

// ---- factor transfers: one contiguous PCIe copy + a repack kernel -------------------------------------------------------
static int ensure_stage(glrmb200_engine* E) {
  if (E->d_stage) return 0;
  return dalloc(&E->d_stage, (size_t)(E->m + E->d) * (size_t)E->k, E->stream);
}
static void pack_launch(glrmb200_engine* E, const double* src, double* dst, int64_t col0, int64_t ncols, double* const* peers, int n_peers) {
  if (ncols <= 0) return;
  const int64_t total = ncols * E->stride;
  pack_factor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, E->stream>>>(src, dst, col0, ncols, (int)E->k, E->stride, peers, n_peers);
}

extern "C" int glrmb200_upload_factors(glrmb200_handle E, const double* X, const double* Y) {
  if (!E || !X || !Y) return fail(GLRMB200_E_INVALID, "null argument");
  CUDA_OK(cudaSetDevice(E->device));
  int rc = ensure_stage(E);
  if (rc) return rc;
  const int64_t k = E->k;
  double* sx = E->d_stage;
  double* sy = E->d_stage + E->m * k;
  if (E->peer_ready) {
    // sharded upload: this rank moves only the columns it owns across PCIe and stores them into every replica over
    // NVLink; the barrier makes all shards visible everywhere before the first sweep reads them
    const int64_t rb = E->rows.begin, re = E->rows.end;
    const int64_t yb = E->has_vec ? E->ystart[(size_t)E->cols.begin] : E->cols.begin, ye = E->has_vec ? E->ystart[(size_t)E->cols.end] : E->cols.end;
    if ((rc = comm_barrier(E))) return rc;      // every rank is done reading its replicas (e.g. a download after the last fit)
    if (re > rb) CUDA_OK(cudaMemcpyAsync(sx + rb * k, X + rb * k, (size_t)((re - rb) * k) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    if (ye > yb) CUDA_OK(cudaMemcpyAsync(sy + yb * k, Y + yb * k, (size_t)((ye - yb) * k) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    pack_launch(E, sx + rb * k, E->d_X, rb, re - rb, E->d_peer_X, E->nranks - 1);
    pack_launch(E, sy + yb * k, E->d_Y, yb, ye - yb, E->d_peer_Y, E->nranks - 1);
    CUDA_OK(cudaGetLastError());
    if ((rc = comm_barrier(E))) return rc;
  } else if (E->dn.on && E->nranks > 1) {
    // fully observed path on several GPUs: a rank only ever touches its own rows of X (and the whole of Y)
    const int64_t rb = E->rows.begin, re = E->rows.end;
    if (re > rb) CUDA_OK(cudaMemcpyAsync(sx + rb * k, X + rb * k, (size_t)((re - rb) * k) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CUDA_OK(cudaMemcpyAsync(sy, Y, (size_t)(E->d * k) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    pack_launch(E, sx + rb * k, E->d_X, rb, re - rb, nullptr, 0);
    pack_launch(E, sy, E->d_Y, 0, E->d, nullptr, 0);
    CUDA_OK(cudaGetLastError());
  } else {
    CUDA_OK(cudaMemcpyAsync(sx, X, (size_t)(E->m * k) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CUDA_OK(cudaMemcpyAsync(sy, Y, (size_t)(E->d * k) * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    pack_launch(E, sx, E->d_X, 0, E->m, nullptr, 0);
    pack_launch(E, sy, E->d_Y, 0, E->d, nullptr, 0);
    CUDA_OK(cudaGetLastError());
  }
  CUDA_OK(cudaStreamSynchronize(E->stream));     // the host buffers are the caller's again
  E->factors_resident = true;
  return 0;
}

extern "C" int glrmb200_download_factors(glrmb200_handle E, double* X, double* Y) {
  if (!E || !X || !Y) return fail(GLRMB200_E_INVALID, "null argument");
  if (!E->factors_resident) return fail(GLRMB200_E_STATE, "no factors on the device");
  CUDA_OK(cudaSetDevice(E->device));
  int rc = ensure_stage(E);
  if (rc) return rc;
  const int64_t k = E->k;
  double* sx = E->d_stage;
  double* sy = E->d_stage + E->m * k;
  // fully observed path on several GPUs: this rank holds (and returns) only its own rows of X; the caller's other rows
  // are left as they are (each process of a multi-process fit gets its shard back)
  const bool own_rows = E->dn.on && E->nranks > 1;
  const int64_t rb = own_rows ? E->rows.begin : 0, re = own_rows ? E->rows.end : E->m;
  if (re > rb) unpack_factor_kernel<<<(unsigned)(((re - rb) * k + 255) / 256), 256, 0, E->stream>>>(E->d_X + rb * E->stride, sx + rb * k, re - rb, (int)k, E->stride);
  unpack_factor_kernel<<<(unsigned)((E->d * k + 255) / 256), 256, 0, E->stream>>>(E->d_Y, sy, E->d, (int)k, E->stride);
  CUDA_OK(cudaGetLastError());
  if (re > rb) CUDA_OK(cudaMemcpyAsync(X + rb * k, sx + rb * k, (size_t)((re - rb) * k) * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CUDA_OK(cudaMemcpyAsync(Y, sy, (size_t)(E->d * k) * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  return 0;
}

// all-gather of per-unit arrays whose shards are contiguous (`elems` doubles per unit)
static int allgather_units(glrmb200_engine* E, double* buf, const Side& S, int64_t elems, const int64_t* colmap = nullptr) {
  if (E->nranks == 1) return 0;
  if (getenv("GLRMB200_NO_EXCHANGE") && atoi(getenv("GLRMB200_NO_EXCHANGE"))) return 0;   // tuning hook, see comm_barrier
  if (!E->comm) return fail(GLRMB200_E_STATE, "glrmb200_comm_init was not called");
  NCCL_OK(g_nccl.GroupStart());
  for (int r = 0; r < E->nranks; ++r) {
    const int64_t b = colmap ? colmap[S.bounds[r]] : S.bounds[r];
    const int64_t e = colmap ? colmap[S.bounds[r + 1]] : S.bounds[r + 1];
    const int64_t cnt = (e - b) * elems;
    if (cnt == 0) continue;
    NCCL_OK(g_nccl.Broadcast(buf + b * elems, buf + b * elems, (size_t)cnt, kNcclDouble, r, E->comm, E->stream));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}

static int dense_gather_groups(glrmb200_engine* E, double* gsum, int64_t len) {
  if (E->nranks == 1 || len <= 0) return 0;
  if (!E->comm) return fail(GLRMB200_E_STATE, "glrmb200_comm_init was not called");
  NCCL_OK(g_nccl.GroupStart());
  for (int r = 0; r < E->nranks; ++r) {
    const int g0 = r * DN_GROUPS / E->nranks, g1 = (r + 1) * DN_GROUPS / E->nranks;
    NCCL_OK(g_nccl.Broadcast(gsum + (int64_t)g0 * len, gsum + (int64_t)g0 * len, (size_t)((g1 - g0) * len), kNcclDouble, r, E->comm, E->stream));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}

// objective(glrm, X, Y) on the resident factors: losses over observed_examples + penalties.  Every read of the factors
// (the evaluation sweep over Y's columns, the row penalties over X) is enqueued BEFORE the barrier: once a rank has
// passed it, its peers may start storing the next sweep's columns into this rank's replicas.
static int objective_resident(glrmb200_engine* E, bool include_reg, double* out, int64_t* launches) {
  cudaError_t ce = cudaSuccess;
  if (E->dn.on) {
    int rc0 = dense_eval_cols(E, FLAG_EVAL_ONLY | (include_reg ? 0 : FLAG_NO_REG), INFINITY, false, 1, launches);
    if (rc0) return rc0;
  } else {
    SweepArgs A = make_args(E, /*x_side=*/false, FLAG_EVAL_ONLY | (include_reg ? 0 : FLAG_NO_REG), INFINITY);
    ce = launch_sweep(E, A, E->cols, launches);
    if (ce == cudaSuccess) ce = launch_vec(E, A, false, E->cols, launches);
    if (ce != cudaSuccess) return fail(GLRMB200_E_CUDA, "objective sweep launch: %s", cudaGetErrorString(ce));
  }
  if (include_reg) {
    ce = launch_reg_eval(E, E->d_X, E->rows, E->rows.d_obj);
    if (ce != cudaSuccess) return fail(GLRMB200_E_CUDA, "penalty launch: %s", cudaGetErrorString(ce));
    ++*launches;
  }
  if (E->dn.on) {
    // fully observed path: Y (and obj_by_col) are replicated, rows are owned by whole groups of row blocks; the row penalties
    // are summed per group (by the owner), gathered, and totalled in group order — the same tree on 1, 2, 4 or 8 ranks
    sum_kernel<<<1, 1024, 0, E->stream>>>(E->cols.d_obj, E->n, E->d_scalars);
    ++*launches;
    if (include_reg) {
      DenseHost& D = E->dn;
      const int64_t grows = (int64_t)D.bg * D.rows_per_block;
      for (int g = D.g0; g < D.g1; ++g) {
        const int64_t b = std::min<int64_t>(E->m, g * grows), e = std::min<int64_t>(E->m, (g + 1) * grows);
        sum_kernel<<<1, 1024, 0, E->stream>>>(E->rows.d_obj + b, e - b, D.d_gsum_x + g);
        ++*launches;
      }
      int rcg = dense_gather_groups(E, D.d_gsum_x, 1);
      if (rcg) return rcg;
      sum_kernel<<<1, 1024, 0, E->stream>>>(D.d_gsum_x, DN_GROUPS, E->d_scalars + 1);
      ++*launches;
    }
  } else {
  int rc = E->peer_ready ? comm_barrier(E) : allgather_units(E, E->cols.d_obj, E->cols, 1);
  if (rc) return rc;
  sum_kernel<<<1, 1024, 0, E->stream>>>(E->cols.d_obj, E->n, E->d_scalars);
  ++*launches;
  if (include_reg) {
    sum_kernel<<<1, 1024, 0, E->stream>>>(E->rows.d_obj, E->m, E->d_scalars + 1);
    ++*launches;
  }
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(E->h_pinned, E->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  double total = E->h_pinned[0];
  if (include_reg) total += E->h_pinned[1];
  *out = total;
  return 0;
}

static int fill(glrmb200_engine* E, double* p, int64_t n, double v, const int* stop = nullptr) {
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, E->stream>>>(p, n, v, stop);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int glrmb200_fit_resident(glrmb200_handle E, const glrmb200_params* prm, double* ch_objective,
                                     double* ch_seconds, int32_t cap, int32_t* n_recorded,
                                     glrmb200_profile* profile) {
  if (!E || !prm || !ch_objective || !ch_seconds || !n_recorded) return fail(GLRMB200_E_INVALID, "null argument");
  if (!E->factors_resident) return fail(GLRMB200_E_STATE, "upload factors first");
  if (cap < prm->max_iter + 1) return fail(GLRMB200_E_INVALID, "cap %d < max_iter+1", cap);
  if (prm->inner_iter_X < 1 || prm->inner_iter_Y < 1) return fail(GLRMB200_E_INVALID, "inner iteration counts must be >= 1");
  CUDA_OK(cudaSetDevice(E->device));
  glrmb200_profile prof;
  memset(&prof, 0, sizeof(prof));
  int rc;
  using clk = std::chrono::steady_clock;
  const int max_iter = prm->max_iter;

  // ---- setup (proxgrad.jl:69-76) ---------------------------------------------------------------------
  CUDA_OK(cudaEventRecord(E->ev_base, E->stream));
  if ((rc = fill(E, E->rows.d_alpha, E->m, prm->stepsize))) return rc;       // :69
  if ((rc = fill(E, E->cols.d_alpha, E->n, prm->stepsize))) return rc;       // :70
  CUDA_OK(cudaMemsetAsync(E->d_trials, 0, 2 * sizeof(unsigned long long), E->stream));
  CUDA_OK(cudaMemsetAsync(E->d_stop, 0, sizeof(int), E->stream));
  *E->h_stop = 0;
  if (E->objs_cap < max_iter + 1) {
    dfree(E->d_objs, E->stream);
    if ((rc = dalloc(&E->d_objs, (size_t)max_iter + 1, E->stream))) return rc;
    E->objs_cap = max_iter + 1;
  }
  if (E->evpool.empty()) {
    E->evpool.resize((size_t)EVENT_CHUNK * 5, nullptr);
    for (auto& e : E->evpool) CUDA_OK(cudaEventCreate(&e));
  }
  const double scaled_abs_tol = prm->abs_tol * (double)E->nnz_rows_total;     // :72
  double obj0 = 0.0;
  if ((rc = objective_resident(E, true, &obj0, &prof.other_launches))) return rc;   // :76
  E->h_pinned[2] = obj0;
  CUDA_OK(cudaMemcpyAsync(E->d_objs, E->h_pinned + 2, sizeof(double), cudaMemcpyHostToDevice, E->stream));
  CUDA_OK(cudaEventRecord(E->ev_aux, E->stream));
  CUDA_OK(cudaEventSynchronize(E->ev_aux));
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, E->ev_base, E->ev_aux));
  prof.setup_ms = ms;

  // ---- the loop (proxgrad.jl:107-217), enqueued ahead of the device -----------------------------------------------------
  // The record and the stopping rule run on the device (record_kernel); the host never waits for an objective value.  It
  // polls the mapped stop flag between iterations and stops enqueuing once the device has stopped; launches that were
  // already enqueued return immediately (SweepArgs::stop).  Timing events are harvested in chunks of EVENT_CHUNK iterations.
  std::vector<double> t_end((size_t)max_iter + 1, 0.0);
  const auto t_loop = clk::now();
  CUDA_OK(cudaEventRecord(E->ev_base, E->stream));
  int enqueued = 0, harvested = 0;
  auto harvest = [&](int upto) -> int {       // iterations (harvested, upto]; the stream has been synchronised up to `upto`
    for (int it = harvested + 1; it <= upto; ++it) {
      cudaEvent_t* e = &E->evpool[(size_t)((it - 1) % EVENT_CHUNK) * 5];
      float a = 0, b = 0, c = 0, d2 = 0, te = 0;
      cudaEventElapsedTime(&a, e[0], e[1]);
      cudaEventElapsedTime(&b, e[1], e[2]);
      cudaEventElapsedTime(&c, e[2], e[3]);
      cudaEventElapsedTime(&d2, e[3], e[4]);
      cudaEventElapsedTime(&te, E->ev_base, e[4]);
      prof.update_x_ms += a; prof.update_y_ms += c; prof.comm_ms += b + d2;
      t_end[(size_t)it] = te * 1e-3;
    }
    harvested = upto;
    return 0;
  };
  const int sync_every = E->nranks == 1 ? 0 : (E->dn.on ? 1 : 8);
  for (int it = 1; it <= max_iter; ++it) {                                // :107
    if (it > 1 && (it - 1) % EVENT_CHUNK == 0) {                           // the event pool wraps: harvest the chunk first
      CUDA_OK(cudaStreamSynchronize(E->stream));
      harvest(it - 1);
    }
    // the device has stopped: nothing more to enqueue.  Several ranks must leave the loop at the SAME iteration (they
    // enqueue collectives / peer barriers in lockstep), so they only look at the flag after a synchronisation, at
    // iterations fixed in advance; the iterations enqueued in between return at once on every rank.
    if (sync_every == 0) { if (*E->h_stop != 0) break; }
    else if (it > 1 && (it - 1) % sync_every == 0) {
      CUDA_OK(cudaStreamSynchronize(E->stream));
      if (*E->h_stop != 0) break;
    }
    cudaEvent_t* ev = &E->evpool[(size_t)((it - 1) % EVENT_CHUNK) * 5];
    if (prm->inner_iter_X > 1 || prm->inner_iter_Y > 1) {                  // :112-115
      if ((rc = fill(E, E->rows.d_alpha, E->m, prm->stepsize, E->d_stop))) return rc;
      if ((rc = fill(E, E->cols.d_alpha, E->n, prm->stepsize, E->d_stop))) return rc;
    }
    CUDA_OK(cudaEventRecord(ev[0], E->stream));
    for (int inner = 0; inner < prm->inner_iter_X; ++inner) {              // :117-158
      if (E->dn.on) { if ((rc = dense_sweep_x(E, prm->min_stepsize, true, &prof.x_launches))) return rc; continue; }
      SweepArgs A = make_args(E, true, 0, prm->min_stepsize, true);
      cudaError_t ce = launch_sweep(E, A, E->rows, &prof.x_launches);
      if (ce == cudaSuccess) ce = launch_vec(E, A, true, E->rows, &prof.x_launches);
      if (ce != cudaSuccess) return fail(GLRMB200_E_CUDA, "update-X launch: %s", cudaGetErrorString(ce));
    }
    CUDA_OK(cudaEventRecord(ev[1], E->stream));
    if (E->dn.on) {}                                                       // fully observed path: X never leaves its owner
    else if (E->peer_ready) { if ((rc = comm_barrier(E))) return rc; }     // columns already stored into the peers
    else if ((rc = allgather_units(E, E->d_X, E->rows, E->stride))) return rc;
    CUDA_OK(cudaEventRecord(ev[2], E->stream));
    for (int inner = 0; inner < prm->inner_iter_Y; ++inner) {              // :160-203
      if (E->dn.on) { if ((rc = dense_sweep_y(E, prm->min_stepsize, true, &prof.y_launches))) return rc; continue; }
      SweepArgs A = make_args(E, false, 0, prm->min_stepsize, true);
      cudaError_t ce = launch_sweep(E, A, E->cols, &prof.y_launches);
      if (ce == cudaSuccess) ce = launch_vec(E, A, false, E->cols, &prof.y_launches);
      if (ce != cudaSuccess) return fail(GLRMB200_E_CUDA, "update-Y launch: %s", cudaGetErrorString(ce));
    }
    CUDA_OK(cudaEventRecord(ev[3], E->stream));
    if (E->dn.on) {}                                                       // ... and every rank holds the whole of Y
    else if (E->peer_ready) { if ((rc = comm_barrier(E))) return rc; }
    else {
      if ((rc = allgather_units(E, E->d_Y, E->cols, E->stride, E->has_vec ? E->ystart.data() : nullptr))) return rc;
      if ((rc = allgather_units(E, E->cols.d_obj, E->cols, 1))) return rc;
    }
    record_kernel<<<1, 1024, 0, E->stream>>>(E->cols.d_obj, E->n, E->d_objs, it, scaled_abs_tol, prm->rel_tol, E->d_stop,
                                             const_cast<int*>(E->h_stop));                               // :204-213
    CUDA_OK(cudaGetLastError());
    ++prof.other_launches;
    CUDA_OK(cudaEventRecord(ev[4], E->stream));
    enqueued = it;
  }
  CUDA_OK(cudaStreamSynchronize(E->stream));
  prof.loop_ms = std::chrono::duration<double, std::milli>(clk::now() - t_loop).count();
  harvest(enqueued);
  if ((rc = check_barrier_timeout(E))) return rc;
  if ((rc = dense_check_diag(E))) return rc;
  int stop_it = 0;
  CUDA_OK(cudaMemcpy(&stop_it, E->d_stop, sizeof(int), cudaMemcpyDeviceToHost));
  const int last = stop_it > 0 ? stop_it : enqueued;          // iterations actually executed
  CUDA_OK(cudaMemcpy(ch_objective, E->d_objs, (size_t)(last + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  ch_seconds[0] = 0.0;
  for (int it = 1; it <= last; ++it) ch_seconds[it] = t_end[(size_t)it] - t_end[(size_t)it - 1];           // :206-207
  prof.iterations = last;
  prof.reduce_ms = prof.loop_ms - prof.update_x_ms - prof.update_y_ms - prof.comm_ms;
  unsigned long long tr[2] = {0, 0};
  CUDA_OK(cudaMemcpy(tr, E->d_trials, sizeof(tr), cudaMemcpyDeviceToHost));
  prof.x_trials = (int64_t)tr[0];
  prof.y_trials = (int64_t)tr[1];
  *n_recorded = last + 1;
  if (profile) *profile = prof;
  return 0;
}

extern "C" int glrmb200_fit(glrmb200_handle E, const glrmb200_params* prm, double* X, double* Y,
                            double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
                            glrmb200_profile* profile) {
  if (!E || !X || !Y) return fail(GLRMB200_E_INVALID, "null argument");
  // the reference's norm(Y)==0 branch is an UndefVarError (proxgrad.jl:45-47): refuse
  bool allzero = true;
  for (int64_t i = 0; i < E->k * E->d && allzero; ++i) allzero = (Y[i] == 0.0);
  if (allzero) return fail(GLRMB200_E_INVALID, "norm(Y) == 0: the solver would never move (proxgrad.jl:44-47)");
  int rc = glrmb200_upload_factors(E, X, Y);
  if (rc) return rc;
  if ((rc = glrmb200_fit_resident(E, prm, ch_objective, ch_seconds, cap, n_recorded, profile))) return rc;
  return glrmb200_download_factors(E, X, Y);
}

extern "C" int glrmb200_fit_sparse(glrmb200_handle E, const glrmb200_sparse_params* prm, double* X, double* Y,
                                   double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
                                   glrmb200_profile* profile) {
  if (!E || !prm || !X || !Y || !ch_objective || !ch_seconds || !n_recorded) return fail(GLRMB200_E_INVALID, "null argument");
  if (E->has_vec) return fail(GLRMB200_E_UNSUPPORTED, "SparseProxGradParams handles scalar-embedding losses only (sparse_proxgrad.jl:70 uses dot(x_e, y_f))");
  if (E->dn.on) return fail(GLRMB200_E_UNSUPPORTED, "SparseProxGradParams on a fully observed handle: create it with glrmb200_create_ex(..., GLRMB200_CREATE_GATHER_ONLY)");
  if (cap < prm->max_iter + 2) return fail(GLRMB200_E_INVALID, "cap %d < max_iter+2", cap);
  if (prm->inner_iter < 1) return fail(GLRMB200_E_INVALID, "inner_iter must be >= 1");
  bool allzero = true;
  for (int64_t i = 0; i < E->k * E->d && allzero; ++i) allzero = (Y[i] == 0.0);
  if (allzero) return fail(GLRMB200_E_INVALID, "norm(Y) == 0: the solver would never move (sparse_proxgrad.jl:37-40)");
  int rc = glrmb200_upload_factors(E, X, Y);                      // working copies X, Y (:33)
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(E->device));
  using clk = std::chrono::steady_clock;
  glrmb200_profile prof;
  memset(&prof, 0, sizeof(prof));
  // best-so-far factors (glrm.X / glrm.Y in the reference) live next to the working copies on the device
  const size_t fac_doubles = (size_t)(E->m + E->d) * E->stride;   // X and Y are contiguous inside d_xchg
  double* d_best = nullptr;
  if ((rc = dalloc(&d_best, fac_doubles, E->stream))) return rc;
  CUDA_OK(cudaMemcpyAsync(d_best, E->d_X, fac_doubles * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));
  auto cleanup = [&]() { cudaStreamSynchronize(E->stream); dfree(d_best, E->stream); };

  double alpha = prm->stepsize;                                                   // :44
  const double tol = prm->abs_tol * (double)E->nnz_rows_total;                    // :46
  int nrec = 0;
  double obj0 = 0.0;
  if ((rc = objective_resident(E, true, &obj0, &prof.other_launches))) { cleanup(); return rc; }   // :50
  ch_objective[nrec] = obj0; ch_seconds[nrec++] = 0.0;
  auto t = clk::now();
  const auto t_loop = t;
  int steps_in_a_row = 0;
  for (int it = 1; it <= prm->max_iter; ++it) {                                   // :60
    for (int side = 0; side < 2; ++side) {                                        // X update :62-79, Y update :83-99
      const bool xs = side == 0;
      Side& S = xs ? E->rows : E->cols;
      for (int inner = 0; inner < prm->inner_iter; ++inner) {
        SweepArgs A = make_args(E, xs, FLAG_UNCONDITIONAL, 0.0);
        A.global_alpha = alpha;
        cudaError_t ce = launch_sweep(E, A, S, xs ? &prof.x_launches : &prof.y_launches);
        if (ce != cudaSuccess) { cleanup(); return fail(GLRMB200_E_CUDA, "sparse sweep launch: %s", cudaGetErrorString(ce)); }
      }
      if (E->peer_ready) rc = comm_barrier(E);
      else rc = allgather_units(E, xs ? E->d_X : E->d_Y, S, E->stride);
      if (rc) { cleanup(); return rc; }
    }
    double obj = 0.0;
    if ((rc = objective_resident(E, true, &obj, &prof.other_launches))) { cleanup(); return rc; }   // :102
    prof.iterations = it;
    if (obj < ch_objective[nrec - 1]) {                                           // :104
      const auto now = clk::now();
      ch_objective[nrec] = obj;
      ch_seconds[nrec++] = std::chrono::duration<double>(now - t).count();        // :105-106
      CUDA_OK(cudaMemcpyAsync(d_best, E->d_X, fac_doubles * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));   // :107
      // with the fused exchange the peers store into this replica during their next sweep: the snapshot must be taken first
      if (E->peer_ready && (rc = comm_barrier(E))) { cleanup(); return rc; }
      alpha = alpha * 1.05;                                                       // :108
      steps_in_a_row = std::max(1, steps_in_a_row + 1);                           // :109
      t = clk::now();
    } else {
      alpha = alpha / std::max(1.5, (double)(-steps_in_a_row));                   // :113
      CUDA_OK(cudaMemcpyAsync(E->d_X, d_best, fac_doubles * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));   // :115
      // ... and they must not start before the revert has landed
      if (E->peer_ready && (rc = comm_barrier(E))) { cleanup(); return rc; }
      steps_in_a_row = std::min(0, steps_in_a_row - 1);                           // :116
    }
    // :119  i>10 && (steps_in_a_row > 3 && ch.objective[end-1] - obj < tol) || alpha <= min_stepsize
    const double prev = nrec >= 2 ? ch_objective[nrec - 2] : INFINITY;
    if ((it > 10 && (steps_in_a_row > 3 && prev - obj < tol)) || alpha <= prm->min_stepsize) break;
  }
  ch_objective[nrec] = ch_objective[nrec - 1];                                    // :125-126
  ch_seconds[nrec] = std::chrono::duration<double>(clk::now() - t).count();
  ++nrec;
  prof.loop_ms = std::chrono::duration<double, std::milli>(clk::now() - t_loop).count();
  // hand back the best model: working copies <- best, then the usual download
  CUDA_OK(cudaMemcpyAsync(E->d_X, d_best, fac_doubles * sizeof(double), cudaMemcpyDeviceToDevice, E->stream));
  cleanup();
  if ((rc = check_barrier_timeout(E))) return rc;
  *n_recorded = nrec;
  if (profile) *profile = prof;
  return glrmb200_download_factors(E, X, Y);
}

extern "C" int glrmb200_objective(glrmb200_handle E, const double* X, const double* Y,
                                  int32_t include_regularization, double* out) {
  if (!E || !X || !Y || !out) return fail(GLRMB200_E_INVALID, "null argument");
  int rc = glrmb200_upload_factors(E, X, Y);
  if (rc) return rc;
  int64_t launches = 0;
  return objective_resident(E, include_regularization != 0, out, &launches);
}

extern "C" int glrmb200_set_reg_scale(glrmb200_handle E, double newscale) {
  // scale_regularizer! -> mul!(r, newscale) sets r.scale (regularizers.jl:38); regularizers whose
  // mul! is a no-op (constraints, ZeroReg: :76,97,114,138,255,318,348) are left untouched
  if (!E) return fail(GLRMB200_E_STATE, "null handle");
  CUDA_OK(cudaSetDevice(E->device));
  for (Side* S : {&E->rows, &E->cols}) {
    for (size_t i = 0; i < S->h_reg_code.size(); ++i) {
      const int base = S->h_reg_code[i] & GLRMB200_REG_BASE_MASK;
      if (base == GLRMB200_REG_QUAD || base == GLRMB200_REG_ONE || base == GLRMB200_REG_REM_QUAD) S->h_reg_param[i * GLRMB200_REG_NPARAM] = newscale;
    }
    CUDA_OK(cudaMemcpyAsync(S->d_reg_param, S->h_reg_param.data(), S->h_reg_param.size() * sizeof(double), cudaMemcpyHostToDevice, E->stream));
  }
  CUDA_OK(cudaStreamSynchronize(E->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// evaluation of a fitted model (csrc/glrm_eval.cuh): impute / error_metric
static int eval_args(glrmb200_engine* E, const double* X, const double* Y, const int32_t* dom_code, const double* dom_param,
                     EvalArgs* P, int32_t** d_dc, double** d_dp) {
  if (!E || !X || !Y || !dom_code || !dom_param) return fail(GLRMB200_E_INVALID, "null argument");
  if (E->nranks != 1) return fail(GLRMB200_E_UNSUPPORTED, "impute / error_metric run on one rank");
  int rc = glrmb200_upload_factors(E, X, Y);
  if (rc) return rc;
  for (int64_t f = 0; f < E->n; ++f)
    if (dom_code[f] < GLRMB200_DOMAIN_REAL || dom_code[f] > GLRMB200_DOMAIN_COUNT) return fail(GLRMB200_E_INVALID, "domain code %d (column %lld) is unknown", dom_code[f], (long long)f);
  if ((rc = upload(d_dc, dom_code, (size_t)E->n, E->stream))) return rc;
  if ((rc = upload(d_dp, dom_param, (size_t)(2 * E->n), E->stream))) return rc;
  memset(P, 0, sizeof(*P));
  P->X = E->d_X; P->Y = E->d_Y; P->stride = E->stride; P->k = (int32_t)E->k; P->m = E->m; P->n = E->n;
  P->ystart = E->d_ystart;                                   // nullptr: every feature has one column
  P->loss_code = E->d_loss_code; P->loss_param = E->d_loss_param;
  P->dom_code = *d_dc; P->dom_param = *d_dp;
  if (E->dn.on) { P->dense_A = E->dn.d_A; P->lda = E->dn.lda; }
  else if (E->obs_full) { P->dense_A = E->cols.d_val; P->lda = E->m; }
  else { P->col_ptr = E->cols.d_ptr; P->col_idx = E->cols.d_idx; P->col_val = E->cols.d_val; }
  return 0;
}

extern "C" int glrmb200_impute(glrmb200_handle E, const double* X, const double* Y, const int32_t* dom_code,
                               const double* dom_param, double* A_imputed) {
  if (!A_imputed) return fail(GLRMB200_E_INVALID, "null argument");
  EvalArgs P;
  int32_t* d_dc = nullptr;
  double *d_dp = nullptr, *d_out = nullptr;
  int rc = eval_args(E, X, Y, dom_code, dom_param, &P, &d_dc, &d_dp);
  if (!rc) rc = dalloc(&d_out, (size_t)(E->m * E->n), E->stream);
  if (!rc) {
    P.out_imputed = d_out;
    const dim3 grid((unsigned)E->n, (unsigned)std::min<int64_t>(64, (E->m + 255) / 256), 1);
    eval_kernel<0><<<grid, 256, (size_t)VEC_DMAX * E->k * sizeof(double), E->stream>>>(P);
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(A_imputed, d_out, (size_t)(E->m * E->n) * sizeof(double), cudaMemcpyDeviceToHost, E->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(E->stream);
    if (ce != cudaSuccess) rc = fail(GLRMB200_E_CUDA, "impute: %s", cudaGetErrorString(ce));
  }
  if (E) { dfree(d_dc, E->stream); dfree(d_dp, E->stream); dfree(d_out, E->stream); }
  return rc;
}

extern "C" int glrmb200_error_metric(glrmb200_handle E, const double* X, const double* Y, const int32_t* dom_code,
                                     const double* dom_param, int32_t standardize, double* out) {
  if (!out) return fail(GLRMB200_E_INVALID, "null argument");
  EvalArgs P;
  int32_t* d_dc = nullptr;
  double *d_dp = nullptr, *d_col = nullptr;
  int rc = eval_args(E, X, Y, dom_code, dom_param, &P, &d_dc, &d_dp);
  if (!rc) rc = dalloc(&d_col, (size_t)(3 * E->n), E->stream);
  if (!rc) {
    P.col_err = d_col; P.col_sq = d_col + E->n; P.col_cnt = d_col + 2 * E->n;
    eval_kernel<1><<<(unsigned)E->n, 256, (size_t)VEC_DMAX * E->k * sizeof(double), E->stream>>>(P);
    eval_total_kernel<<<1, 32, 0, E->stream>>>(P.col_err, P.col_sq, P.col_cnt, E->n, standardize, E->d_scalars);
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(E->h_pinned, E->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, E->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(E->stream);
    if (ce != cudaSuccess) rc = fail(GLRMB200_E_CUDA, "error_metric: %s", cudaGetErrorString(ce));
    else *out = E->h_pinned[0];
  }
  if (E) { dfree(d_dc, E->stream); dfree(d_dp, E->stream); dfree(d_col, E->stream); }
  return rc;
}

extern "C" int glrmb200_get_stepsizes(glrmb200_handle E, double* alpharow, double* alphacol) {
  if (!E) return fail(GLRMB200_E_STATE, "null handle");
  CUDA_OK(cudaSetDevice(E->device));
  int rc;
  if ((rc = allgather_units(E, E->rows.d_alpha, E->rows, 1))) return rc;
  if ((rc = allgather_units(E, E->cols.d_alpha, E->cols, 1))) return rc;
  if (alpharow) CUDA_OK(cudaMemcpyAsync(alpharow, E->rows.d_alpha, (size_t)E->m * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  if (alphacol) CUDA_OK(cudaMemcpyAsync(alphacol, E->cols.d_alpha, (size_t)E->n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CUDA_OK(cudaStreamSynchronize(E->stream));
  return 0;
}
