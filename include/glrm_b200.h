/*
 * glrm_b200.h — C ABI of the B200-native GLRM proximal-gradient fitting engine.
 *
 * This is the drop-in boundary for ONE hot path of madeleineudell/LowRankModels.jl:
 *     fit!(glrm::GLRM, params::ProxGradParams; ch, verbose)      src/algorithms/proxgrad.jl:34-220
 * (threaded twin src/algorithms/proxgrad_multithread.jl:34-222).  The reference has no FFI; its
 * extension point is Julia multiple dispatch on `T <: AbstractParams` (src/fit.jl:4,8-21), exactly
 * how SparseProxGradParams plugs in (src/algorithms/sparse_proxgrad.jl:4-24).  A Julia maintainer
 * adds `struct B200ProxGradParams <: AbstractParams` and a `fit!` method that `ccall`s the entry
 * points below (the binding is shown in INTEGRATION.md and shipped, unexecuted, in
 * lowrankmodels.jl_b200/julia/LowRankModelsB200.jl).
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success or a negative
 *     GLRMB200_E_* code; glrmb200_last_error() gives the thread-local message.
 *   - the caller owns every host pointer; the library copies what it needs during the call and
 *     never keeps a host pointer after returning.
 *   - factor matrices are column-major Float64 exactly as Julia stores glrm.X (k x m) and
 *     glrm.Y (k x d): element (r, c) at [c*k + r].
 *   - all indices crossing the ABI are 0-based (the Julia shim subtracts 1).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     GLRMB200_E_NO_DEVICE.
 */
#ifndef GLRM_B200_H
#define GLRM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLRMB200_VERSION 102          /* major*100 + minor */
#define GLRMB200_LOSS_NPARAM 8        /* doubles per loss descriptor */
#define GLRMB200_REG_NPARAM 4         /* doubles per regularizer descriptor */

/* ---- status codes ------------------------------------------------------------------------- */
enum {
  GLRMB200_OK = 0,
  GLRMB200_E_INVALID = -1,      /* bad argument / shape (reference: error(...) in src/glrm.jl:38-43) */
  GLRMB200_E_UNSUPPORTED = -2,  /* loss / regularizer without a device implementation        */
  GLRMB200_E_NO_DEVICE = -3,    /* no CUDA device: the engine never falls back to the CPU     */
  GLRMB200_E_CUDA = -4,         /* CUDA runtime error (message has the call site)             */
  GLRMB200_E_NCCL = -5,         /* NCCL error / libnccl not loadable                          */
  GLRMB200_E_LABEL = -6,        /* label outside the loss's domain (reference: InexactError from
                                   myBool, src/losses.jl:104; BoundsError for u[a])           */
  GLRMB200_E_NAN = -7,          /* NaN among observed values (reference: src/glrm.jl:63-71)   */
  GLRMB200_E_STATE = -8         /* call order (e.g. fit on a destroyed handle)                */
};

/* ---- loss codes (src/losses.jl) ----------------------------------------------------------- *
 * loss_param[f*8 + ...]:  [0]=scale  [1]=p1  [2]=p2  [3]=bin_loss code  [4]=bin scale  [5]=bin p1  [6]=bin p2
 *   QUAD           losses.jl:138-146   -
 *   L1             losses.jl:152-160   -
 *   HUBER          losses.jl:166-177   p1=crossover
 *   QUANTILE       losses.jl:186-201   p1=quantile
 *   PERIODIC       losses.jl:209-218   p1=T
 *   POISSON        losses.jl:231-241   -
 *   ORDINAL_HINGE  losses.jl:247-292   p1=min p2=max
 *   LOGISTIC       losses.jl:298-306   -            labels: 1 -> true; 0,-1 -> false (losses.jl:104)
 *   WEIGHTED_HINGE losses.jl:317-341   p1=case_weight_ratio   (HingeLoss == ratio 1, losses.jl:324)
 *   MULTINOMIAL    losses.jl:360-398   p2=max  (embedding dim = max)       labels 1..max
 *   OVA            losses.jl:413-438   p2=max  [3..5]=bin loss             labels 1..max
 *   BVS            losses.jl:450-475   p2=max  (embedding dim = max-1)     labels 1..max
 *   ORDISTIC       losses.jl:490-519   p2=max
 *   MULTINOMIAL_ORDINAL losses.jl:562-608  p2=max (embedding dim = max-1)
 */
enum {
  GLRMB200_LOSS_QUAD = 1,
  GLRMB200_LOSS_L1 = 2,
  GLRMB200_LOSS_HUBER = 3,
  GLRMB200_LOSS_QUANTILE = 4,
  GLRMB200_LOSS_PERIODIC = 5,
  GLRMB200_LOSS_POISSON = 6,
  GLRMB200_LOSS_ORDINAL_HINGE = 7,
  GLRMB200_LOSS_LOGISTIC = 8,
  GLRMB200_LOSS_WEIGHTED_HINGE = 9,
  GLRMB200_LOSS_MULTINOMIAL = 10,
  GLRMB200_LOSS_OVA = 11,
  GLRMB200_LOSS_BVS = 12,
  GLRMB200_LOSS_ORDISTIC = 13,
  GLRMB200_LOSS_MULTINOMIAL_ORDINAL = 14
};

/* ---- regularizer codes (src/regularizers.jl) ---------------------------------------------- *
 * reg_code = base | wrapper flags;  reg_param[i*4 + ...]: [0]=scale / max_2norm / k
 *   ZERO            regularizers.jl:91-97
 *   QUAD            regularizers.jl:52-58     [0]=scale
 *   QUAD_CONSTRAINT regularizers.jl:68-76     [0]=max_2norm
 *   ONE             regularizers.jl:79-88     [0]=scale
 *   NONNEG          regularizers.jl:101-114
 *   NONNEG_ONE      regularizers.jl:118-138   [0]=scale (prox ignores it, as in the reference)
 *   ONE_SPARSE      regularizers.jl:235-255
 *   KSPARSE         regularizers.jl:258-291   [0]=k
 *   UNIT_ONE_SPARSE regularizers.jl:295-318
 *   SIMPLEX         regularizers.jl:323-348
 * wrappers (the offset path, src/modify_glrm.jl:21-24):
 *   LASTENTRY1            regularizers.jl:163-174   last factor entry pinned to 1
 *   LASTENTRY_UNPENALIZED regularizers.jl:178-189   last factor entry skipped by the inner reg
 *   REM_QUAD        regularizers.jl:412-423   [0]=scale, payload = the mean m (k values)
 * wrappers with a vector payload (rx_payload / ry_payload below; not combined with the offset wrappers, REM_QUAD or the
 * block regularizers):
 *   FIXED_FIRST     regularizers.jl:193-210   fixed_latent_features: the first n entries are pinned to the payload y
 *                                             (n = payload length), the inner regularizer sees the other k-n;
 *                                             evaluate = Inf unless a[1:n] == y
 *   FIXED_LAST      regularizers.jl:214-231   fixed_last_latent_features: the last n entries pinned to y.  Restated
 *                                             literally: prox = [prox(inner, u[n+1:end]); y] (:223 feeds the LAST k-n
 *                                             entries to the inner prox, not the first); evaluate uses a[1:k-n] (:230)
 * block regularizers of the ordinal losses (ry only; applied to the whole k x d_f block of a column):
 *   ORDINAL_REG           regularizers.jl:356-380   first k-1 rows: mean over the block's columns, inner prox, copied
 *                                                    back to every column; evaluate = inner reg on a[1:k-1, 1]
 *   MNL_ORDINAL_REG       regularizers.jl:385-407   the same + last row made negative and decreasing (TOL 1e-3)
 */
enum {
  GLRMB200_REG_ZERO = 0,
  GLRMB200_REG_QUAD = 1,
  GLRMB200_REG_QUAD_CONSTRAINT = 2,
  GLRMB200_REG_ONE = 3,
  GLRMB200_REG_NONNEG = 4,
  GLRMB200_REG_NONNEG_ONE = 5,
  GLRMB200_REG_ONE_SPARSE = 6,
  GLRMB200_REG_KSPARSE = 7,
  GLRMB200_REG_UNIT_ONE_SPARSE = 8,
  GLRMB200_REG_SIMPLEX = 9,
  GLRMB200_REG_REM_QUAD = 10,
  GLRMB200_REG_BASE_MASK = 0xff,
  GLRMB200_REG_LASTENTRY1 = 0x100,
  GLRMB200_REG_LASTENTRY_UNPENALIZED = 0x200,
  GLRMB200_REG_ORDINAL = 0x400,
  GLRMB200_REG_MNL_ORDINAL = 0x800,
  GLRMB200_REG_FIXED_FIRST = 0x1000,
  GLRMB200_REG_FIXED_LAST = 0x2000
};

/* ---- the problem: what `GLRM(A, losses, rx, ry, k; ...)` holds (src/glrm.jl:12-22) ---------- *
 * Observations replace glrm.observed_features / glrm.observed_examples (src/glrm.jl:9,17-18;
 * built by sort_observations, src/modify_glrm.jl:5-18).  Both adjacency lists are passed because
 * the reference keeps both and they may legitimately differ (order, duplicates, even content:
 * src/cross_validate.jl:255-257).  Values of A are co-located with each list so the device
 * never performs an A[e,f] lookup (src/algorithms/proxgrad.jl:125,168).
 */
typedef struct glrmb200_problem {
  int64_t m;                 /* rows of A      (examples)                                      */
  int64_t n;                 /* columns of A   (features)                                      */
  int64_t k;                 /* rank                                                           */
  int64_t d;                 /* sum of embedding dims == size(Y,2)   (losses.jl:72-93)         */

  const int32_t* loss_code;  /* [n]                                                            */
  const double*  loss_param; /* [n * GLRMB200_LOSS_NPARAM]                                     */

  int64_t        rx_count;   /* 1 (one regularizer shared by all rows) or m                    */
  const int32_t* rx_code;    /* [rx_count]                                                     */
  const double*  rx_param;   /* [rx_count * GLRMB200_REG_NPARAM]                               */
  int64_t        ry_count;   /* 1 or n                                                         */
  const int32_t* ry_code;    /* [ry_count]                                                     */
  const double*  ry_param;   /* [ry_count * GLRMB200_REG_NPARAM]                               */

  int32_t obs_full;          /* 1: every entry observed (the UnitRange default,
                                src/glrm.jl:33-34); `dense_A` is used, the lists are ignored   */
  const double*  dense_A;    /* [m*n] column-major, as Julia stores A; labels as Float64       */

  /* observed_features: for row e, entries row_ptr[e] .. row_ptr[e+1]-1 in list order           */
  const int64_t* row_ptr;    /* [m+1]                                                          */
  const int32_t* row_idx;    /* [nnz_rows] feature index f (0-based)                           */
  const double*  row_val;    /* [nnz_rows] A[e,f]                                              */
  /* observed_examples: for column f, entries col_ptr[f] .. col_ptr[f+1]-1 in list order        */
  const int64_t* col_ptr;    /* [n+1]                                                          */
  const int32_t* col_idx;    /* [nnz_cols] example index e (0-based)                           */
  const double*  col_val;    /* [nnz_cols] A[e,f]                                              */

  /* optional vector payloads of the regularizers (fixed_latent_features.y, fixed_last_latent_features.y,
   * RemQuadReg.m: src/regularizers.jl:193-231,412-423).  Regularizer i of the side owns
   * payload[payload_ptr[i] .. payload_ptr[i+1]); NULL pointers = no regularizer carries a payload.    */
  const int64_t* rx_payload_ptr;  /* [rx_count + 1]                                            */
  const double*  rx_payload;
  const int64_t* ry_payload_ptr;  /* [ry_count + 1]                                            */
  const double*  ry_payload;
} glrmb200_problem;

/* ---- ProxGradParams (src/algorithms/proxgrad.jl:4-31), same seven fields -------------------- */
typedef struct glrmb200_params {
  double  stepsize;      /* initial step size                     (default 1.0)                */
  int32_t max_iter;      /* outer iterations                      (default 100)                */
  int32_t inner_iter_X;  /* prox-grad steps on X per outer iter   (default 1)                  */
  int32_t inner_iter_Y;  /* prox-grad steps on Y per outer iter   (default 1)                  */
  double  abs_tol;       /* stop if decrease < abs_tol * |Omega|  (default 1e-5)               */
  double  rel_tol;       /* stop if decrease/obj < rel_tol        (default 1e-4)               */
  double  min_stepsize;  /* line-search floor                     (default 0.01*stepsize)      */
} glrmb200_params;

/* ---- SparseProxGradParams (src/algorithms/sparse_proxgrad.jl:4-18), same five fields ------------- */
typedef struct glrmb200_sparse_params {
  double  stepsize;      /* initial (global) step size                  (default 1.0)              */
  int32_t max_iter;      /* outer iterations                            (default 100)              */
  int32_t inner_iter;    /* prox-grad steps on X, then on Y, per iter   (default 1)                */
  double  abs_tol;       /* stop if decrease < abs_tol * |Omega|        (default 1e-5)             */
  double  min_stepsize;  /* stop when the step size falls to this       (default 0.01*stepsize)    */
} glrmb200_sparse_params;

/* per-fit device timings (CUDA events on the engine's own stream), all in milliseconds         */
typedef struct glrmb200_profile {
  double  setup_ms;        /* H2D of X,Y + initial objective                                   */
  double  update_x_ms;     /* sum over iterations of the update-X launches                     */
  double  update_y_ms;     /* sum over iterations of the update-Y launches                     */
  double  reduce_ms;       /* objective reduction + D2H of the scalar                          */
  double  comm_ms;         /* all-gathers (multi-GPU only)                                     */
  double  loop_ms;         /* first update-X launch .. last iteration end                      */
  int64_t x_launches;      /* kernels launched for X sweeps                                    */
  int64_t y_launches;      /* kernels launched for Y sweeps                                    */
  int64_t other_launches;  /* objective / reduction kernels                                    */
  int64_t x_trials;        /* line-search trial passes summed over rows and iterations         */
  int64_t y_trials;        /* line-search trial passes summed over columns and iterations      */
  int32_t iterations;      /* outer iterations executed                                        */
  int32_t reserved;
} glrmb200_profile;

typedef struct glrmb200_engine* glrmb200_handle;

/* Library / device discovery. */
int         glrmb200_version(void);
const char* glrmb200_last_error(void);
int         glrmb200_device_count(int32_t* count);          /* 0 devices => GLRMB200_E_NO_DEVICE */

/* glrmb200_create: validate the problem (labels, NaN, shapes: src/glrm.jl:38-43,63-71),
 * encode it for the device (int32 indices, k padded to a 32-byte multiple, degree-sorted
 * schedules) and upload it to `device`.  Replaces the setup block src/algorithms/proxgrad.jl:38-105.
 * rank/nranks describe the shard this handle owns (1-GPU: rank=0, nranks=1); with nranks>1 call
 * glrmb200_comm_init before glrmb200_fit. */
int glrmb200_create(glrmb200_handle* out, const glrmb200_problem* problem,
                    int32_t device, int32_t rank, int32_t nranks);

/* glrmb200_create_ex: the same with creation flags.
 *   GLRMB200_CREATE_GATHER_ONLY  a fully observed problem is encoded with (implicit) observation lists for the gather
 *                                kernels instead of the streaming kernels of the fully observed path — what
 *                                fit!(glrm, ::SparseProxGradParams) on a dense A needs (glrmb200_fit_sparse runs on the
 *                                gather kernels only). */
enum { GLRMB200_CREATE_GATHER_ONLY = 1 };
int glrmb200_create_ex(glrmb200_handle* out, const glrmb200_problem* problem,
                       int32_t device, int32_t rank, int32_t nranks, int32_t flags);

/* Multi-GPU plumbing (one process per GPU).  Rank 0 calls glrmb200_comm_unique_id and the host
 * side (torch.distributed / Julia Distributed) broadcasts the 128 bytes; every rank then calls
 * glrmb200_comm_init.  The communicator is cached per process: later handles of the same (rank, nranks, device)
 * may pass id == NULL to reuse it.  The data-path exchange is one all-gather of the freshly updated factor per
 * half-iteration, or the fused peer-store exchange below. */
int glrmb200_comm_unique_id(uint8_t id[128]);
int glrmb200_comm_init(glrmb200_handle h, const uint8_t id[128]);

/* Host-only shard planner (no device needed): contiguous ranges balanced by observation count.
 * ptr = row_ptr / col_ptr ([count+1]) or NULL (balance by unit count); bounds receives nranks+1 entries. */
int glrmb200_plan_shards(const int64_t* ptr, int64_t count, int32_t nranks, int64_t* bounds);
/* ... and the row ranges of a FULLY OBSERVED problem on nranks GPUs (whole groups of row blocks, multiples of 64 rows; see
 * glrmb200_shard): bounds receives nranks+1 entries.  GLRMB200_E_UNSUPPORTED when such a problem would not be row-sharded
 * (nranks not in {1, 2, 4, 8} or fewer than 512 rows per rank). */
int glrmb200_plan_dense_rows(int64_t m, int32_t nranks, int64_t* bounds);

/* Fused exchange over NVLink peer memory (optional, after glrmb200_comm_init).  Every rank exports CUDA IPC
 * handles of its factor replicas (glrmb200_ipc_export, GLRMB200_IPC_BYTES bytes), the host all-gathers the blobs
 * (rank order) and hands them to glrmb200_ipc_open; from then on the update kernels store every accepted factor
 * column (and the unit's objective) directly into all peers while the sweep runs, and the per-half-iteration NCCL
 * all-gather shrinks to a barrier.  glrmb200_comm_barrier must be called by all ranks before glrmb200_destroy. */
#define GLRMB200_IPC_BYTES 64
int glrmb200_ipc_export(glrmb200_handle h, uint8_t out[GLRMB200_IPC_BYTES]);
int glrmb200_ipc_open(glrmb200_handle h, const uint8_t* all_blobs /* nranks * GLRMB200_IPC_BYTES */);
int glrmb200_comm_barrier(glrmb200_handle h);

/* Row range [row_begin,row_end) and column range [col_begin,col_end) this handle updates
 * (cost-balanced contiguous shards; whole range when nranks==1).
 *
 * Fully observed problems on several GPUs (nranks in {2, 4, 8}, m >= 512 * nranks; SURVEY.md section 8e, second mode):
 * only the ROWS of A and X are sharded — by whole groups of row blocks, [row_begin, row_end) — and every rank updates
 * all of Y (the partial gradients / per-feature objectives of the 8 row-block groups are all-gathered per line-search
 * round and summed in group order, so every rank holds identical Y, step sizes and objective, bit-identical to the
 * one-GPU fit).  Such a handle reads and returns ONLY ITS OWN ROWS of X (glrmb200_fit, glrmb200_upload_factors,
 * glrmb200_download_factors leave the caller's other columns of X[k x m] untouched); Y is complete on every rank.
 * [col_begin, col_end) is [0, n) on rank 0 and empty elsewhere (bookkeeping only). */
int glrmb200_shard(glrmb200_handle h, int64_t* row_begin, int64_t* row_end,
                   int64_t* col_begin, int64_t* col_end);

/* glrmb200_fit: the whole outer loop of proxgrad.jl:107-217 on the device.
 *   X [k*m], Y [k*d]   in/out, column-major (aliases of glrm.X / glrm.Y: mutated in place,
 *                      proxgrad.jl:43,219; warm start = call again)
 *   ch_objective[cap]  out: objective series exactly as the reference records it — entry 0 is the
 *                      full objective (proxgrad.jl:76), later entries are sum(obj_by_col)
 *                      = losses + ry only (proxgrad.jl:205)
 *   ch_seconds[cap]    out: per-entry elapsed seconds (0 for entry 0), NOT cumulative; the shim
 *                      feeds them to update_ch! (src/convergence.jl:16-27) which accumulates
 *   n_recorded         out: entries written (<= max_iter+1; cap must be >= max_iter+1)
 *   profile            optional (may be NULL) */
int glrmb200_fit(glrmb200_handle h, const glrmb200_params* params,
                 double* X, double* Y,
                 double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
                 glrmb200_profile* profile);

/* glrmb200_fit_sparse: fit!(glrm, ::SparseProxGradParams) of src/algorithms/sparse_proxgrad.jl:21-130 — what plain
 * `fit!(glrm)` selects for a SparseMatrixCSC (src/fit.jl:13-15) — on the same kernels: unconditional prox-gradient
 * sweeps with one global step size (:62-99), the full sparse objective (:102), accept (alpha*1.05, keep) or revert
 * (alpha / max(1.5, -steps_in_a_row)) (:104-117), stop rule (:119).  Scalar-embedding losses only, as in the
 * reference.  X, Y in/out: on return they hold the best model found (glrm.X / glrm.Y).  ch_* as the reference
 * records them: entry 0 the initial objective, one entry per ACCEPTED iteration, and the final duplicate (:126);
 * cap >= max_iter + 2. */
int glrmb200_fit_sparse(glrmb200_handle h, const glrmb200_sparse_params* params, double* X, double* Y,
                        double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
                        glrmb200_profile* profile);

/* glrmb200_objective: objective(glrm, X, Y; include_regularization) of src/evaluate_fit.jl:57-83
 * (loss summed over observed_examples, then calc_penalty :91-104), evaluated on observed entries
 * only. */
int glrmb200_objective(glrmb200_handle h, const double* X, const double* Y,
                       int32_t include_regularization, double* out);

/* Residency helpers for the callers that refit the same data (src/cross_validate.jl:141-240):
 * rescale all regularizers (scale_regularizer!, src/glrm.jl:84-88) without re-uploading A ... */
int glrmb200_set_reg_scale(glrmb200_handle h, double newscale);
/* ... and swap the observation lists of a live handle (a training fold: cross_validate.jl:31-33 replaces
 * glrm.observed_features / observed_examples and refits): same array meaning as in glrmb200_problem; losses,
 * regularizers, shapes and the device-resident factors are kept.  A fully observed handle becomes list mode. */
int glrmb200_set_obs(glrmb200_handle h, const int64_t* row_ptr, const int32_t* row_idx, const double* row_val,
                     const int64_t* col_ptr, const int32_t* col_idx, const double* col_val);

/* Device-resident benchmarking hooks: keep factors on the device between calls so the timed
 * region of bench.py's `value` leg contains no host<->device copies.
 *   glrmb200_upload_factors  copies host X,Y to the device buffers
 *   glrmb200_fit_resident    runs the loop on the resident factors (no X/Y copies; objective
 *                            scalars still come back once per iteration)
 *   glrmb200_download_factors copies the device factors to host X,Y */
int glrmb200_upload_factors(glrmb200_handle h, const double* X, const double* Y);
int glrmb200_fit_resident(glrmb200_handle h, const glrmb200_params* params,
                          double* ch_objective, double* ch_seconds, int32_t cap,
                          int32_t* n_recorded, glrmb200_profile* profile);
int glrmb200_download_factors(glrmb200_handle h, double* X, double* Y);

/* Evaluation of a fitted model on the device (callers of the path: cross-validation scores a fold with them,
 * src/cross_validate.jl:34-44).  Domains (src/domains.jl) say how a feature is imputed and scored; domain_code[n] holds
 * GLRMB200_DOMAIN_* and domain_param[2n] the pairs (min, max) for ORDINAL / CATEGORICAL, (T, -) for PERIODIC,
 * (max_count, -) for COUNT.  The host side passes each loss's own domain (l.domain) unless the caller overrides it.
 *   glrmb200_impute        impute(glrm) = impute(losses, X'Y) (src/evaluate_fit.jl:150, src/impute_and_err.jl:147-168):
 *                          A_imputed[m*n] column-major, a_u = argmin_a loss(u, a) over the domain (impute_and_err.jl:40-120);
 *                          NaN where the reference has no method for the (domain, loss) pair or raises
 *   glrmb200_error_metric  error_metric(glrm, X, Y, domains; standardize) (src/evaluate_fit.jl:106-143) over the observed
 *                          entries: squared error or misclassification of the imputed value, per the domain
 * One rank only (GLRMB200_E_UNSUPPORTED on a sharded handle). */
enum { GLRMB200_DOMAIN_REAL = 1, GLRMB200_DOMAIN_BOOL = 2, GLRMB200_DOMAIN_ORDINAL = 3, GLRMB200_DOMAIN_CATEGORICAL = 4,
       GLRMB200_DOMAIN_PERIODIC = 5, GLRMB200_DOMAIN_COUNT = 6 };
int glrmb200_impute(glrmb200_handle h, const double* X, const double* Y, const int32_t* domain_code,
                    const double* domain_param, double* A_imputed);
int glrmb200_error_metric(glrmb200_handle h, const double* X, const double* Y, const int32_t* domain_code,
                          const double* domain_param, int32_t standardize, double* out);

/* Step-size state (alpharow / alphacol, proxgrad.jl:69-70) for tests: n_row = m, n_col = n. */
int glrmb200_get_stepsizes(glrmb200_handle h, double* alpharow, double* alphacol);

int glrmb200_destroy(glrmb200_handle h);

#ifdef __cplusplus
}
#endif
#endif /* GLRM_B200_H */
