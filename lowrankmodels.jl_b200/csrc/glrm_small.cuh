// glrm_small.cuh — the non-template helper kernels of the engine (included by glrm_engine.cu only): objective
// reduction, input validation, dense transposition.  All < 1 % of a step.
#pragma once
#include "glrm_device.cuh"

namespace glrm {

// out[0] = sum(v[0..n)) in a fixed order (obj = sum(obj_by_col), proxgrad.jl:205)
__global__ void __launch_bounds__(1024) sum_kernel(const double* __restrict__ v, int64_t n, double* out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// One outer iteration's record and stopping rule on the device (proxgrad.jl:204-213): objs[it] = sum(obj_by_col) in the
// same fixed order as sum_kernel; stop iff it > 10 and (decrease < scaled_abs_tol or decrease / obj < rel_tol).  The host
// enqueues iterations without waiting for the objective; once `stop` is set every later launch of the fit returns at once.
__global__ void __launch_bounds__(1024) record_kernel(const double* __restrict__ v, int64_t n, double* objs, int it,
                                                      double scaled_abs_tol, double rel_tol, int* stop, volatile int* host_stop) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && *stop == 0) {
    const double obj = sh[0];
    objs[it] = obj;                                                      // :205-207
    const double obj_decrease = objs[it - 1] - obj;                      // :210
    if (it > 10 && (obj_decrease < scaled_abs_tol || obj_decrease / obj < rel_tol)) {   // :211
      *stop = it;
      *host_stop = it;
      __threadfence_system();
    }
  }
}

// p[i] = v unless the fit has stopped (step sizes are reset per outer iteration when inner_iter > 1, proxgrad.jl:112-115)
__global__ void fill_kernel(double* p, int64_t n, double v, const int* stop) {
  if (stop != nullptr && *reinterpret_cast<const volatile int*>(stop) != 0) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Barrier between the ranks of a fused-exchange fit, over NVLink peer memory: thread p publishes this rank's epoch in
// peer p's flag array and waits until peer p's epoch shows up in ours.  It follows the sweep kernels on the same stream,
// so the peer stores of the sweep have been performed before the flag is written (kernel boundary + system fence), and it
// precedes the kernels that read what the peers stored.  ~3 us instead of the 30-50 us of a 1-element NCCL all-reduce.
__global__ void peer_barrier_kernel(unsigned long long* const* peer_flags, volatile unsigned long long* my_flags,
                                    const int* peer_rank, int my_rank, int n_peers, unsigned long long epoch, int* timed_out) {
  const int p = threadIdx.x;
  if (p >= n_peers) return;
  __threadfence_system();
  volatile unsigned long long* dst = peer_flags[p] + my_rank;
  *dst = epoch;
  __threadfence_system();
  const int pr = peer_rank[p];
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (my_flags[pr] < epoch) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) { *timed_out = 1; break; }          // 20 s: a peer died; the host reports it
    __nanosleep(100);
  }
  __threadfence_system();
}

// Factor matrices cross PCIe as one contiguous copy (k doubles per column) into / out of a staging buffer; these kernels
// move them to / from the padded device layout (stride doubles per column, zero past k).  A 2-D cudaMemcpy of 400-byte rows
// into a 512-byte pitch runs at 3.5 GB/s (measured, profiles/r2_microbench.json) against 55 GB/s for the contiguous copy.
// With peers, the columns are stored into every replica (sharded upload: each rank uploads only its own columns).
__global__ void pack_factor_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t col0, int64_t ncols,
                                   int k, int stride, double* const* peers, int n_peers) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols * stride) return;
  const int64_t c = i / stride;
  const int e = (int)(i - c * stride);
  const double v = e < k ? src[c * k + e] : 0.0;
  const int64_t o = (col0 + c) * stride + e;
  dst[o] = v;
  for (int p = 0; p < n_peers; ++p) peers[p][o] = v;
}
__global__ void unpack_factor_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t ncols, int k, int stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols * k) return;
  const int64_t c = i / k;
  const int e = (int)(i - c * k);
  dst[i] = src[c * stride + e];
}

// ---- input validation on the device (glrm.jl:63-71 NaN check; myBool / level bounds, losses.jl:104) ----
// bad[0] = smallest offending entry position (or ~0), bad[1] = error kind of some offender at that position
__device__ __forceinline__ int label_error(int code, const double* __restrict__ lp, double a) {
  if (a != a) return GLRMB200_E_NAN;
  switch (code) {
    case GLRMB200_LOSS_LOGISTIC: case GLRMB200_LOSS_WEIGHTED_HINGE:
      return (a == 1.0 || a == 0.0 || a == -1.0) ? 0 : GLRMB200_E_LABEL;
    case GLRMB200_LOSS_MULTINOMIAL: case GLRMB200_LOSS_OVA: case GLRMB200_LOSS_BVS:
    case GLRMB200_LOSS_ORDISTIC: case GLRMB200_LOSS_MULTINOMIAL_ORDINAL:
      return (a == floor(a) && a >= 1.0 && a <= lp[2]) ? 0 : GLRMB200_E_LABEL;
    default: return 0;
  }
}
__device__ __forceinline__ void report_bad(unsigned long long* bad, unsigned long long pos, int kind) {
  const unsigned long long key = (pos << 4) | (unsigned long long)(-kind);   // kinds are small negatives
  atomicMin(bad, key);
}
// side lists where the entry's feature is idx[q] (row side) — one thread per entry
__global__ void validate_rows_kernel(const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t nnz,
                                     int64_t n, const int32_t* __restrict__ loss_code,
                                     const double* __restrict__ loss_param, unsigned long long* bad) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nnz) return;
  const int32_t f = idx[q];
  if (f < 0 || f >= n) { report_bad(bad, (unsigned long long)q, GLRMB200_E_INVALID); return; }
  const int e = label_error(loss_code[f], loss_param + (int64_t)f * GLRMB200_LOSS_NPARAM, val[q]);
  if (e) report_bad(bad, (unsigned long long)q, e);
}
// column side: the feature is the unit owning the entry — one warp per unit
__global__ void validate_cols_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                     const double* __restrict__ val, int64_t units, int64_t unit_base, int64_t m,
                                     const int32_t* __restrict__ loss_code, const double* __restrict__ loss_param,
                                     unsigned long long* bad) {
  const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= units) return;
  const int64_t f = unit_base + u;
  const int code = loss_code[f];
  const double* lp = loss_param + f * GLRMB200_LOSS_NPARAM;
  for (int64_t q = ptr[u] + (threadIdx.x & 31); q < ptr[u + 1]; q += 32) {
    const int32_t e = idx[q];
    if (e < 0 || e >= m) { report_bad(bad, (unsigned long long)q, GLRMB200_E_INVALID); continue; }
    const int err = label_error(code, lp, val[q]);
    if (err) report_bad(bad, (unsigned long long)q, err);
  }
}
// fully observed: A is column-major m x n, element i belongs to feature i / m
__global__ void validate_dense_kernel(const double* __restrict__ A, int64_t total, int64_t m, int64_t lda,
                                      const int32_t* __restrict__ loss_code, const double* __restrict__ loss_param,
                                      unsigned long long* bad) {
  // `total` = lda * columns positions of a column-major array with leading dimension lda >= m; reports f * m + e
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t f = i / lda, e = i - f * lda;
  if (e >= m) return;
  const int err = label_error(loss_code[f], loss_param + f * GLRMB200_LOSS_NPARAM, A[i]);
  if (err) report_bad(bad, (unsigned long long)(f * m + e), err);
}

// dst[c*rows + r] = src[r*cols + c]   (row-major copy of the Julia column-major A for the X sweep)
__global__ void transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t rows, int64_t cols) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][i];
  }
}

}  // namespace glrm
