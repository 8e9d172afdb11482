// glrm_launch.cuh — launch plumbing shared by the engine (glrm_engine.cu) and the per-(loss template, tile half)
// instantiation units (sweep_inst.cu).  The sweep kernels are instantiated for 11 tiles x 3 loss templates; splitting
// them over translation units lets `make -j` build them in parallel.
#pragma once
#include "glrm_device.cuh"

namespace glrm {

// main: where the sweep is ordered; tier[0..2]: streams of the cluster-8 / cluster-4 / warp tiers (nullptr: run on main);
// join[i] pairs with tier[i]
struct Streams { cudaStream_t main; cudaStream_t tier[3]; cudaEvent_t fork; cudaEvent_t join[3]; };

// tier sizes of one sweep: schedule = [cluster8 | cluster4 | CTA | warp] (degree-sorted, heaviest first)
struct TierCounts { int64_t n_cluster16, n_cluster4, n_heavy, n_light; };

#define GLRM_DECLARE_LAUNCH(NAME) \
  cudaError_t NAME(int g, int r, const SweepArgs& A, const TierCounts& tc, const Streams& st, int64_t* launches);
GLRM_DECLARE_LAUNCH(launch_quad_narrow)
GLRM_DECLARE_LAUNCH(launch_quad_wide)
GLRM_DECLARE_LAUNCH(launch_logistic_narrow)
GLRM_DECLARE_LAUNCH(launch_logistic_wide)
GLRM_DECLARE_LAUNCH(launch_generic_narrow)
GLRM_DECLARE_LAUNCH(launch_generic_wide)
#undef GLRM_DECLARE_LAUNCH

}  // namespace glrm
