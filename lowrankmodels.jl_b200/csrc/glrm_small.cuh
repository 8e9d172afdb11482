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
__global__ void validate_dense_kernel(const double* __restrict__ A, int64_t total, int64_t m,
                                      const int32_t* __restrict__ loss_code, const double* __restrict__ loss_param,
                                      unsigned long long* bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t f = i / m;
  const int err = label_error(loss_code[f], loss_param + f * GLRMB200_LOSS_NPARAM, A[i]);
  if (err) report_bad(bad, (unsigned long long)i, err);
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
