// glrm_dense_host.h — argument blocks and launch wrappers of the fully observed path, shared by the engine
// (glrm_engine.cu) and the kernel unit (dense_inst.cu, glrm_dense.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glrm {

constexpr int DN_TM = 64;            // rows per tile
constexpr int DN_TN = 64;            // Y columns per chunk
constexpr int DN_THREADS = 128;
constexpr int DN_RP = 66;            // pitch (doubles) of [index][row] tiles: 16-byte stores from the GEMM layout are conflict-free
constexpr int DN_YP = DN_TN + 1;     // pitch of the Y chunk [i][column]: column reads and row reads are both conflict-free
constexpr int DN_CT = DN_TN / 16;    // columns per thread in the U contraction
constexpr int DN_GROUPS = 8;         // row-block groups: the unit of the Y sweep's reduction tree and of multi-GPU row sharding

// tensor-core kernels (glrm_dense_mma.cuh): FP64 mma.sync m8n8k4 (SASS DMMA)
constexpr int MM_WARPS = 8;                        // consumer warps of a CTA (each owns 16 rows / columns of the CTA's tile)
constexpr int MM_THREADS = 32 * MM_WARPS;
constexpr int MM_TM = 16 * MM_WARPS;               // rows of X (X sweep) / columns of Y (Y sweep) a CTA owns
constexpr int MM_SC = 32;                          // columns / rows of the other factor per pipeline stage
constexpr int MM_UNIT = 16;                        // plan granularity: whole features are packed into units of <= 16 columns
constexpr int MM_SW = 33;                          // pitch (doubles) of a warp's 16 x 32 scratch tile (heterogeneous losses)
constexpr int MM_MAX_STAGES = 4;
// the rank k is served by the smallest instantiated NT (n-tiles of 8 factor indices) with 8 NT >= k
inline int mm_nt_for_k(int k) {
  const int need = (k + 7) / 8;
  const int inst[] = {1, 2, 3, 4, 6, 8, 10, 12, 13};
  for (int v : inst) if (v >= need) return v;
  return 0;
}
// shared memory of the tensor-core kernels (mirrors mm_carve)
inline size_t mm_smem_bytes(int nt, int nst, bool generic) {
  const size_t P = 8 * (size_t)nt + 4;
  size_t b = (size_t)MM_TM * P * 8 + (size_t)nst * MM_SC * P * 8;
  if (generic) b += (size_t)MM_WARPS * 16 * MM_SW * 8 + (size_t)MM_WARPS * 16 * 32 * 8;
  b += 8 * (size_t)MM_TM * 8;                      // rowv[4][MM_TM] + psum[4][MM_TM]
  b += 16 * 8;                                     // mbarriers
  b += (size_t)nst * (generic ? 136 : 32) * 4;     // stage meta
  b += (size_t)(generic ? 544 : 256) * 4;          // row state (X sweep) / own-column meta (Y sweep)
  b += 16 * 4;                                     // counters
  return b;
}

struct DenseArgs {
  const double* A;            // column-major m x n as Julia stores it, leading dimension lda (multiple of 64, zero rows past m)
  int64_t m, n, lda;
  int32_t nbuf;               // tile buffers of A in shared memory (1 or 2)
  int32_t* diag;              // [8] device-side watchdog record: [0] != 0 means a kernel gave up waiting (see dn_give_up)
  int64_t row0, row1;         // rows this launch covers
  double* X;                  // k x m factor (device layout: `stride` doubles per column)
  const double* Ymat;         // k x d matrix the pass contracts with (Y, or the trial blocks Ynew)
  int32_t stride, k, kp;
  const int64_t* ystart;      // [n+1] first Y column of each feature
  const int32_t* loss_code;   // [n]
  const double* loss_param;   // [n * 8]
  // chunk plan: chunk c holds the features feat_list[chunk_ptr[c] .. chunk_ptr[c+1]); feat_off = first local column
  const int32_t* chunk_ptr;
  const int32_t* feat_list;
  const int32_t* feat_off;
  const int32_t* nchunks;     // device scalar
  // X side
  const int32_t* reg_code; const double* reg_param; int32_t reg_uniform;
  double* alpha; double min_stepsize; double* obj_out;
  double* gscratch;           // [grid][64][stride] gradient of the tile (kept out of shared memory during the trials)
  unsigned long long* trial_counter;
  const int* stop;
  int32_t flags;
  // Y side
  double* gpart;              // [n_blocks][d][stride] partial G_Y per row block
  double* objpart;            // [n_blocks][n] partial per-feature loss sums
  int64_t rows_per_block;
  int32_t n_blocks;
  int32_t block0;             // first row block of this rank (multi-GPU: rows are sharded by groups of row blocks)
  // tensor-core kernels: per-unit column tables of the plan in use ([unit * MM_UNIT + column]: feature / Y column, -1 = unused),
  // pipeline stages in shared memory, loss parameters of a uniform-loss problem
  const int32_t* ucol_feat;
  const int32_t* ucol_y;
  int32_t nst;
  double uparam[3];
  unsigned long long* phase;  // [8] or nullptr: clocks of thread 0 per phase of the X sweep (GLRMB200_PHASE_TIMERS=1): waiting for
                              // the tile, gradient pass, search set-up, trial points, trial losses, accept / compact; rounds; tiles
};

struct DenseYState {
  double* Y; double* Ynew; double* G;          // [d][stride]
  const int64_t* ystart;
  double* colobj;                              // [n] loss sums of the last pass
  double* objold; double* regnew;              // [n]
  double* alpha; double* obj_out;              // [n] alphacol, obj_by_col
  int32_t* active;                             // [n] 1 = still searching
  int32_t* nactive;                            // device scalar
  volatile int32_t* h_nactive;                 // mapped host copy
  int32_t* chunk_ptr; int32_t* feat_list; int32_t* feat_off; int32_t* nchunks;   // plan of the features still searching
  const int32_t* reg_code; const double* reg_param; int32_t reg_uniform;
  int64_t n, d, m;
  int32_t stride, k;
  double min_stepsize;
  unsigned long long* trial_counter;
  const int* stop;
  int32_t flags;
  int32_t seq;                                 // sweep sequence number published next to h_nactive
  int32_t unit_cols;                           // columns per chunk of the plan (DN_TN, or MM_UNIT for the tensor-core kernels)
  int32_t* ucol_feat; int32_t* ucol_y;         // per-unit column tables of the plan (tensor-core kernels; else nullptr)
};

cudaError_t dense_launch_x(int kt, int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st);
cudaError_t dense_launch_y_pass(int kt, int loss, int mode, const DenseArgs& P, int n_blocks, int max_chunks, cudaStream_t st);
// two-level fixed-order reduction over the row blocks: DN_GROUPS groups of `bg` consecutive blocks each
//   groups:  gsum[g][x] = sum over the blocks of group g (for g0 <= g < g1: the groups this rank owns)
//   total:   out[x] = sum over all DN_GROUPS groups (after the groups of the other ranks have been gathered)
cudaError_t dense_launch_reduce_groups(const double* part, int bg, int g0, int g1, int64_t len, double* gsum, const int32_t* nactive, const int* stop, cudaStream_t st);
cudaError_t dense_launch_reduce_total(const double* gsum, int64_t len, double* out, const int32_t* nactive, const int* stop, cudaStream_t st);
cudaError_t dense_launch_plan(const DenseYState& Q, cudaStream_t st);
cudaError_t dense_launch_begin(int tg, int tr, const DenseYState& Q, cudaStream_t st);
cudaError_t dense_launch_step(int tg, int tr, const DenseYState& Q, cudaStream_t st);
cudaError_t dense_launch_decide(const DenseYState& Q, cudaStream_t st);
size_t dense_smem_needed(int k, int kt, int nbuf);
// tensor-core kernels: nt = mm_nt_for_k(k); P.nst stages; smem = mm_smem_bytes(nt, P.nst, loss == 0)
cudaError_t dense_mma_launch_x(int nt, int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st);
cudaError_t dense_mma_launch_y(int nt, int loss, int mode, const DenseArgs& P, int n_blocks, int max_units, cudaStream_t st);

}  // namespace glrm
