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

// The tiers of a sweep touch disjoint units, so each tier is launched on its own stream and they all start together: the
// heaviest units (cluster tiers, LPT) no longer delay the CTA tier behind them, and the warp tier fills the SMs the others
// leave idle.  The side streams carry descending priorities (cluster tiers first) so that, when CTAs of several tiers
// compete for an SM, the units on the critical path are placed first.  Everything is joined back into `main`.
template <int G, int R, int LOSS>
static cudaError_t launch_tile(const SweepArgs& A0, const TierCounts& tc, const Streams& st, int64_t* launches) {
  const int64_t counts[4] = {tc.n_cluster16, tc.n_cluster4, tc.n_heavy, tc.n_light};
  // stream of each tier: the CTA tier stays on main; a tier with no side stream, or alone in the sweep, runs on main too
  cudaStream_t where[4] = {st.tier[0], st.tier[1], st.main, st.tier[2]};
  int n_tiers = 0;
  for (int i = 0; i < 4; ++i) n_tiers += counts[i] > 0;
  bool forked[4] = {false, false, false, false};
  for (int i = 0; i < 4; ++i) {
    if (where[i] == nullptr || n_tiers < 2) where[i] = st.main;
    forked[i] = counts[i] > 0 && where[i] != st.main;
  }
  if (forked[0] || forked[1] || forked[3]) cudaEventRecord(st.fork, st.main);
  for (int i = 0; i < 4; ++i) if (forked[i]) cudaStreamWaitEvent(where[i], st.fork, 0);
  const int32_t* order = A0.order;
  if (tc.n_cluster16 > 0) {                  // heaviest units first (LPT): a cluster of 8 CTAs per unit
    SweepArgs K = A0;
    K.order = order; K.n_units = tc.n_cluster16;
    cudaError_t ce = launch_cluster<G, R, LOSS, CLUSTER_CTAS_BIG>(K, tc.n_cluster16, where[0]);
    if (ce != cudaSuccess) return ce;
    ++*launches;
    order += tc.n_cluster16;
  }
  if (tc.n_cluster4 > 0) {
    SweepArgs K = A0;
    K.order = order; K.n_units = tc.n_cluster4;
    cudaError_t ce = launch_cluster<G, R, LOSS, CLUSTER_CTAS>(K, tc.n_cluster4, where[1]);
    if (ce != cudaSuccess) return ce;
    ++*launches;
    order += tc.n_cluster4;
  }
  if (tc.n_heavy > 0) {
    SweepArgs H = A0;
    H.order = order; H.n_units = tc.n_heavy;
    sweep_cta_kernel<G, R, LOSS><<<(unsigned)tc.n_heavy, WARPS_PER_CTA_HEAVY * 32, 0, where[2]>>>(H);
    ++*launches;
    order += tc.n_heavy;
  }
  if (tc.n_light > 0) {
    SweepArgs L = A0;
    L.order = order; L.n_units = tc.n_light;
    const int64_t grid = (tc.n_light + WARPS_PER_CTA_LIGHT - 1) / WARPS_PER_CTA_LIGHT;
    sweep_warp_kernel<G, R, LOSS><<<(unsigned)grid, WARPS_PER_CTA_LIGHT * 32, 0, where[3]>>>(L);
    ++*launches;
  }
  const int jn[4] = {0, 1, -1, 2};
  for (int i = 0; i < 4; ++i) {
    if (!forked[i]) continue;
    cudaEventRecord(st.join[jn[i]], where[i]);
    cudaStreamWaitEvent(st.main, st.join[jn[i]], 0);
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
