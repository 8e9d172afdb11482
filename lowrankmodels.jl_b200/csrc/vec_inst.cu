// vec_inst.cu — instantiations of vec_sweep_kernel (units that involve vector-valued losses), csrc/glrm_vec.cuh
#include "glrm_vec.cuh"

namespace glrm {

cudaError_t launch_vec_inst(int g, int r, const VecArgs& V, bool x_side, int64_t n_vec, cudaStream_t stream, int64_t* launches) {
  const unsigned grid = (unsigned)((n_vec + 3) / 4);
#define T(GG, RR)                                                                                    \
  if (g == GG && r == RR) {                                                                          \
    if (x_side) vec_sweep_kernel<GG, RR, 1><<<grid, 128, 0, stream>>>(V);                             \
    else vec_sweep_kernel<GG, RR, VEC_DMAX><<<grid, 128, 0, stream>>>(V);                             \
    ++*launches;                                                                                     \
    return cudaGetLastError();                                                                       \
  }
  T(4, 1) T(8, 1) T(8, 2)
#undef T
  return cudaErrorInvalidValue;
}

}  // namespace glrm
