// sweep_inst.cu — instantiations of the fused update kernels for ONE loss template and ONE half of the tile list.
// Compiled six times (csrc/Makefile): -DGLRM_INST_LOSS={1 Quad, 8 Logistic, 0 generic} -DGLRM_INST_WIDE={0: G<=8, 1: G>=16}
// -DGLRM_INST_NAME=launch_<loss>_<narrow|wide>.
#include "glrm_launch.cuh"

namespace glrm {

template <int G, int R, int LOSS, int CS>
static cudaError_t launch_cluster(const SweepArgs& K, int64_t n_units, cudaStream_t stream) {
  auto kern = sweep_cluster_kernel<G, R, LOSS, CS>;
  if (CS > 8) {                                // cluster sizes above 8 are opt-in (non-portable)
    static bool allowed = false;
    if (!allowed) {
      cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (ce != cudaSuccess) return ce;
      allowed = true;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_units * CS), 1, 1);
  cfg.blockDim = dim3(WARPS_PER_CTA_HEAVY * 32, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, K);
}

// the tiers of a sweep touch disjoint units: the warp tier is forked onto the side stream so it fills the SMs the
// CTA / cluster tiers leave idle in their tails, and joined back before anything else is enqueued
template <int G, int R, int LOSS>
static cudaError_t launch_tile(const SweepArgs& A0, const TierCounts& tc, const Streams& st, int64_t* launches) {
  const int64_t n_big = tc.n_cluster16 + tc.n_cluster4 + tc.n_heavy;
  const bool both = n_big > 0 && tc.n_light > 0 && st.side;
  if (both) {
    cudaEventRecord(st.fork, st.main);
    cudaStreamWaitEvent(st.side, st.fork, 0);
  }
  const int32_t* order = A0.order;
  if (tc.n_cluster16 > 0) {                  // heaviest units first (LPT): a cluster of 8 CTAs per unit
    SweepArgs K = A0;
    K.order = order; K.n_units = tc.n_cluster16;
    cudaError_t ce = launch_cluster<G, R, LOSS, CLUSTER_CTAS_BIG>(K, tc.n_cluster16, st.main);
    if (ce != cudaSuccess) return ce;
    ++*launches;
    order += tc.n_cluster16;
  }
  if (tc.n_cluster4 > 0) {
    SweepArgs K = A0;
    K.order = order; K.n_units = tc.n_cluster4;
    cudaError_t ce = launch_cluster<G, R, LOSS, CLUSTER_CTAS>(K, tc.n_cluster4, st.main);
    if (ce != cudaSuccess) return ce;
    ++*launches;
    order += tc.n_cluster4;
  }
  if (tc.n_heavy > 0) {
    SweepArgs H = A0;
    H.order = order; H.n_units = tc.n_heavy;
    sweep_cta_kernel<G, R, LOSS><<<(unsigned)tc.n_heavy, WARPS_PER_CTA_HEAVY * 32, 0, st.main>>>(H);
    ++*launches;
    order += tc.n_heavy;
  }
  if (tc.n_light > 0) {
    SweepArgs L = A0;
    L.order = order; L.n_units = tc.n_light;
    const int64_t grid = (tc.n_light + WARPS_PER_CTA_LIGHT - 1) / WARPS_PER_CTA_LIGHT;
    sweep_warp_kernel<G, R, LOSS><<<(unsigned)grid, WARPS_PER_CTA_LIGHT * 32, 0, both ? st.side : st.main>>>(L);
    ++*launches;
  }
  if (both) {
    cudaEventRecord(st.join, st.side);
    cudaStreamWaitEvent(st.main, st.join, 0);
  }
  return cudaGetLastError();
}

cudaError_t GLRM_INST_NAME(int g, int r, const SweepArgs& A, const TierCounts& tc, const Streams& st, int64_t* launches) {
#define T(GG, RR) if (g == GG && r == RR) return launch_tile<GG, RR, GLRM_INST_LOSS>(A, tc, st, launches)
#if GLRM_INST_WIDE
  T(16, 2); T(16, 3); T(16, 4); T(32, 2); T(32, 3); T(32, 4);
#else
  T(4, 1); T(8, 1); T(8, 2); T(8, 3); T(8, 4);
#endif
#undef T
  return cudaErrorInvalidValue;
}

}  // namespace glrm
