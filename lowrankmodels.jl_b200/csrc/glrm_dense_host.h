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
};

cudaError_t dense_launch_x(int kt, int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st);
cudaError_t dense_launch_y_pass(int kt, int loss, int mode, const DenseArgs& P, int n_blocks, int max_chunks, cudaStream_t st);
cudaError_t dense_launch_reduce(const double* part, int n_blocks, int64_t len, double* out, const int32_t* nactive, const int* stop, cudaStream_t st);
cudaError_t dense_launch_plan(const DenseYState& Q, cudaStream_t st);
cudaError_t dense_launch_begin(int tg, int tr, const DenseYState& Q, cudaStream_t st);
cudaError_t dense_launch_step(int tg, int tr, const DenseYState& Q, cudaStream_t st);
cudaError_t dense_launch_decide(const DenseYState& Q, cudaStream_t st);
size_t dense_smem_needed(int k, int kt, int nbuf);

}  // namespace glrm
