// dense_mma_inst.cu — one translation unit per instantiated factor width NT (compiled with -DMM_INST_NT=<nt>, see the
// Makefile): the tensor-core kernels of the fully observed path (csrc/glrm_dense_mma.cuh) for ranks 8 (NT - 1) < k <= 8 NT.
#define GLRM_DENSE_HELPERS_ONLY
#include "glrm_dense_mma.cuh"

#ifndef MM_INST_NT
#error "compile with -DMM_INST_NT=<nt>"
#endif

namespace glrm {

// lane-group tile (G lanes x R slots) of the rank range this NT serves — the same choice as create_impl's tile selection
#if MM_INST_NT == 1
#define MM_TG 4
#define MM_TR 1
#elif MM_INST_NT == 2
#define MM_TG 8
#define MM_TR 1
#elif MM_INST_NT <= 4
#define MM_TG 8
#define MM_TR 2
#elif MM_INST_NT == 6
#define MM_TG 8
#define MM_TR 3
#elif MM_INST_NT == 8
#define MM_TG 8
#define MM_TR 4
#elif MM_INST_NT <= 12
#define MM_TG 16
#define MM_TR 3
#else
#define MM_TG 16
#define MM_TR 4
#endif

#define MM_CAT2(a, b) a##b
#define MM_CAT(a, b) MM_CAT2(a, b)

template <class K>
static cudaError_t mm_set_smem(K kern, size_t smem) {
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t MM_CAT(dense_mma_x_nt, MM_INST_NT)(int tg, int tr, int loss, const DenseArgs& P, int grid, cudaStream_t st) {
  if (tg != MM_TG || tr != MM_TR) return cudaErrorInvalidValue;
  const size_t smem = mm_smem_bytes(MM_INST_NT, P.nst, loss == 0);
  cudaError_t ce;
  if (loss == GLRMB200_LOSS_QUAD) {
    if ((ce = mm_set_smem(dense_mma_x_kernel<MM_INST_NT, MM_TG, MM_TR, GLRMB200_LOSS_QUAD>, smem)) != cudaSuccess) return ce;
    dense_mma_x_kernel<MM_INST_NT, MM_TG, MM_TR, GLRMB200_LOSS_QUAD><<<grid, MM_THREADS, smem, st>>>(P);
  } else {
    if ((ce = mm_set_smem(dense_mma_x_kernel<MM_INST_NT, MM_TG, MM_TR, 0>, smem)) != cudaSuccess) return ce;
    dense_mma_x_kernel<MM_INST_NT, MM_TG, MM_TR, 0><<<grid, MM_THREADS, smem, st>>>(P);
  }
  return cudaGetLastError();
}

template <int LOSS>
static cudaError_t mm_launch_y_mode(int mode, const DenseArgs& P, dim3 grid, size_t smem, cudaStream_t st) {
  cudaError_t ce;
  if (mode == 0) {
    if ((ce = mm_set_smem(dense_mma_y_kernel<MM_INST_NT, LOSS, 0>, smem)) != cudaSuccess) return ce;
    dense_mma_y_kernel<MM_INST_NT, LOSS, 0><<<grid, MM_THREADS, smem, st>>>(P);
  } else {
    if ((ce = mm_set_smem(dense_mma_y_kernel<MM_INST_NT, LOSS, 1>, smem)) != cudaSuccess) return ce;
    dense_mma_y_kernel<MM_INST_NT, LOSS, 1><<<grid, MM_THREADS, smem, st>>>(P);
  }
  return cudaGetLastError();
}

cudaError_t MM_CAT(dense_mma_y_nt, MM_INST_NT)(int loss, int mode, const DenseArgs& P, int n_blocks, int max_units, cudaStream_t st) {
  const size_t smem = mm_smem_bytes(MM_INST_NT, P.nst, loss == 0);
  const dim3 grid((unsigned)n_blocks, (unsigned)((max_units + MM_WARPS - 1) / MM_WARPS), 1);
  if (loss == GLRMB200_LOSS_QUAD) return mm_launch_y_mode<GLRMB200_LOSS_QUAD>(mode, P, grid, smem, st);
  return mm_launch_y_mode<0>(mode, P, grid, smem, st);
}

}  // namespace glrm
