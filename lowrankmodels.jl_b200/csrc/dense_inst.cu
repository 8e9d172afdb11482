// dense_inst.cu — instantiations and launch wrappers of the fully observed path (csrc/glrm_dense.cuh)
#include "glrm_dense.cuh"
#include "glrm_dense_host.h"

namespace glrm {

template <class K>
static cudaError_t set_smem(K kern, size_t smem) {
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int KT, int TG, int TR>
static cudaError_t launch_x_tile(int loss, const DenseArgs& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t ce;
  if (loss == GLRMB200_LOSS_QUAD) {
    if ((ce = set_smem(dense_x_kernel<KT, TG, TR, GLRMB200_LOSS_QUAD>, smem)) != cudaSuccess) return ce;
    dense_x_kernel<KT, TG, TR, GLRMB200_LOSS_QUAD><<<grid, DN_THREADS, smem, st>>>(P);
  } else {
    if ((ce = set_smem(dense_x_kernel<KT, TG, TR, 0>, smem)) != cudaSuccess) return ce;
    dense_x_kernel<KT, TG, TR, 0><<<grid, DN_THREADS, smem, st>>>(P);
  }
  return cudaGetLastError();
}

cudaError_t dense_launch_x(int kt, int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st) {
  const size_t smem = dense_smem_bytes(P.k, kt, P.nbuf);
#define T(KT, TG, TR) if (kt == KT && tg == TG && tr == TR) return launch_x_tile<KT, TG, TR>(loss, P, grid, smem, st)
  T(1, 4, 1); T(1, 8, 1); T(2, 8, 2); T(3, 8, 3); T(4, 8, 4); T(5, 16, 3); T(6, 16, 3); T(7, 16, 4);
#undef T
  return cudaErrorInvalidValue;
}

template <int KT, int LOSS>
static cudaError_t launch_y_mode(int mode, const DenseArgs& P, dim3 grid, size_t smem, cudaStream_t st) {
  cudaError_t ce;
  if (mode == 0) {
    if ((ce = set_smem(dense_y_pass_kernel<KT, LOSS, 0>, smem)) != cudaSuccess) return ce;
    dense_y_pass_kernel<KT, LOSS, 0><<<grid, DN_THREADS, smem, st>>>(P);
  } else {
    if ((ce = set_smem(dense_y_pass_kernel<KT, LOSS, 1>, smem)) != cudaSuccess) return ce;
    dense_y_pass_kernel<KT, LOSS, 1><<<grid, DN_THREADS, smem, st>>>(P);
  }
  return cudaGetLastError();
}

cudaError_t dense_launch_y_pass(int kt, int loss, int mode, const DenseArgs& P, int n_blocks, int max_chunks, cudaStream_t st) {
  const size_t smem = dense_smem_bytes(P.k, kt, P.nbuf);
  const dim3 grid((unsigned)n_blocks, (unsigned)max_chunks, 1);
#define T(KT)                                                                                                   \
  if (kt == KT) {                                                                                               \
    if (loss == GLRMB200_LOSS_QUAD) return launch_y_mode<KT, GLRMB200_LOSS_QUAD>(mode, P, grid, smem, st);      \
    return launch_y_mode<KT, 0>(mode, P, grid, smem, st);                                                       \
  }
  T(1) T(2) T(3) T(4) T(5) T(6) T(7)
#undef T
  return cudaErrorInvalidValue;
}

cudaError_t dense_launch_reduce_groups(const double* part, int bg, int g0, int g1, int64_t len, double* gsum, const int32_t* nactive,
                                       const int* stop, cudaStream_t st) {
  if (len <= 0 || g1 <= g0) return cudaSuccess;
  const dim3 grid((unsigned)((len + 255) / 256), (unsigned)(g1 - g0), 1);
  dense_reduce_groups_kernel<<<grid, 256, 0, st>>>(part, bg, g0, g1, len, gsum, nactive, stop);
  return cudaGetLastError();
}

cudaError_t dense_launch_reduce_total(const double* gsum, int64_t len, double* out, const int32_t* nactive, const int* stop, cudaStream_t st) {
  if (len <= 0) return cudaSuccess;
  dense_reduce_total_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(gsum, len, out, nactive, stop);
  return cudaGetLastError();
}

cudaError_t dense_launch_plan(const DenseYState& Q, cudaStream_t st) {
  dense_y_plan_kernel<<<1, 32, 0, st>>>(Q);
  return cudaGetLastError();
}

cudaError_t dense_launch_begin(int tg, int tr, const DenseYState& Q, cudaStream_t st) {
  const int64_t per_cta = 4 * (32 / tg);
  const unsigned grid = (unsigned)((Q.n + per_cta - 1) / per_cta);
#define T(TG, TR) if (tg == TG && tr == TR) { dense_y_begin_kernel<TG, TR><<<grid, 128, 0, st>>>(Q); return cudaGetLastError(); }
  T(4, 1) T(8, 1) T(8, 2) T(8, 3) T(8, 4) T(16, 3) T(16, 4)
#undef T
  return cudaErrorInvalidValue;
}

cudaError_t dense_launch_step(int tg, int tr, const DenseYState& Q, cudaStream_t st) {
  const int64_t per_cta = 4 * (32 / tg);
  const unsigned grid = (unsigned)((Q.n + per_cta - 1) / per_cta);
#define T(TG, TR) if (tg == TG && tr == TR) { dense_y_step_kernel<TG, TR><<<grid, 128, 0, st>>>(Q); return cudaGetLastError(); }
  T(4, 1) T(8, 1) T(8, 2) T(8, 3) T(8, 4) T(16, 3) T(16, 4)
#undef T
  return cudaErrorInvalidValue;
}

cudaError_t dense_launch_decide(const DenseYState& Q, cudaStream_t st) {
  dense_y_decide_kernel<<<(unsigned)((Q.n + 127) / 128), 128, 0, st>>>(Q);
  return cudaGetLastError();
}

size_t dense_smem_needed(int k, int kt, int nbuf) { return dense_smem_bytes(k, kt, nbuf); }

// tensor-core kernels: one translation unit per factor width (dense_mma_inst.cu)
#define MM_DECL(NT)                                                                                                  \
  cudaError_t dense_mma_x_nt##NT(int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st);           \
  cudaError_t dense_mma_y_nt##NT(int loss, int mode, const DenseArgs& P, int n_blocks, int max_units, cudaStream_t st);
MM_DECL(1) MM_DECL(2) MM_DECL(3) MM_DECL(4) MM_DECL(6) MM_DECL(8) MM_DECL(10) MM_DECL(12) MM_DECL(13)
#undef MM_DECL

cudaError_t dense_mma_launch_x(int nt, int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st) {
#define T(NT) if (nt == NT) return dense_mma_x_nt##NT(tg, tr, loss, P, grid, st);
  T(1) T(2) T(3) T(4) T(6) T(8) T(10) T(12) T(13)
#undef T
  return cudaErrorInvalidValue;
}
cudaError_t dense_mma_launch_y(int nt, int loss, int mode, const DenseArgs& P, int n_blocks, int max_units, cudaStream_t st) {
#define T(NT) if (nt == NT) return dense_mma_y_nt##NT(loss, mode, P, n_blocks, max_units, st);
  T(1) T(2) T(3) T(4) T(6) T(8) T(10) T(12) T(13)
#undef T
  return cudaErrorInvalidValue;
}

}  // namespace glrm
