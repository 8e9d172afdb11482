// microbench.cu — measured roofs for the GLRM sweep kernels on one B200 (run under gpurun; prints one JSON object).
//
//   l2_gather_ldg    random 400-byte factor columns (512-byte stride, as the engine stores k=50 columns) out of a table
//                    that fits L2 (13.7 MB = Y of config 2, 71 MB = X of config 2), fetched exactly like
//                    entry_pass does: a lane group of 8 owns one entry, 4 x LDG.128 per lane through the read-only
//                    path, index stream loaded coalesced.  This is the practical L2->SM gather ceiling the sweeps run against.
//   l2_gather_bulk   the same rows fetched by cp.async.bulk (UBLKCP, one 400-byte bulk copy per entry, mbarrier
//                    complete_tx) into a per-warp shared-memory ring and read back with LDS.128 in the same lane layout.
//   l2_gather_cpasync the same with cp.async 16-byte (LDGSTS) per lane.
//   dfma_peak        FP64 FMA issue peak (the roof of the dense path).
//   hbm_read / hbm_copy  streaming bandwidth (sanity check of MEASURED_PEAKS.json).
//   h2d_pinned       host->device bandwidth from pinned memory (what glrmb200_create is bound by).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ROWB = 400;      // bytes gathered per entry (k = 50 doubles)
constexpr int STRIDE = 512;    // bytes between columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- LDG gather: group of 8 lanes per entry, R = 4 slots, DEPTH steps in flight ------------------------------------
template <int DEPTH, bool BUTTERFLY>
__global__ void __launch_bounds__(128) gather_ldg(const char* __restrict__ table, const int32_t* __restrict__ idx,
                                                  int64_t n_entries, double* out) {
  const int lane = threadIdx.x & 31, lg = lane & 7, gq = lane >> 3;
  const int64_t warp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 4;
  const int64_t chunks = n_entries / 32;
  double acc0 = 0.0, acc1 = 0.0;
  const char* lane_base = table + 16 * lg;
  const char* last_base = table + 16 * ((lg < 1 ? lg : 0) + 24);
  for (int64_t c = warp; c < chunks; c += nwarps) {
    const int32_t j = __ldcs(idx + c * 32 + lane);
    double2 y[DEPTH][4];
    auto fetch = [&](int s, double2 (&yy)[4]) {
      const int32_t jj = __shfl_sync(0xffffffffu, j, (s & 7) * 4 + gq);
      const char* p = lane_base + (int64_t)jj * STRIDE;
#pragma unroll
      for (int r = 0; r < 3; ++r) yy[r] = __ldg(reinterpret_cast<const double2*>(p + r * 128));
      yy[3] = __ldg(reinterpret_cast<const double2*>(last_base + (int64_t)jj * STRIDE));
    };
#pragma unroll
    for (int u = 0; u < DEPTH - 1; ++u) fetch(u, y[u]);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s + DEPTH - 1 < 8) fetch(s + DEPTH - 1, y[(s + DEPTH - 1) % DEPTH]);
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) { d0 = fma(y[s % DEPTH][r].x, 1.0000001, d0); d1 = fma(y[s % DEPTH][r].y, 0.9999999, d1); }
      double d = d0 + d1;
      if (BUTTERFLY) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      }
      acc0 += d;
    }
  }
  if (acc0 + acc1 == 1.2345e300) out[0] = acc0;
}

// ---- bulk-copy gather: lane l issues one cp.async.bulk for entry l of the chunk -----------------------------------------
template <int NS, bool BUTTERFLY>
__global__ void __launch_bounds__(128) gather_bulk(const char* __restrict__ table, const int32_t* __restrict__ idx,
                                                   int64_t n_entries, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, lg = lane & 7, gq = lane >> 3, w = threadIdx.x >> 5;
  unsigned char* ring = smem + (size_t)w * NS * 32 * ROWB;                  // [NS][32][ROWB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)4 * NS * 32 * ROWB) + w * NS;
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(32));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = (int64_t)blockIdx.x * 4 + w;
  const int64_t nwarps = (int64_t)gridDim.x * 4;
  const int64_t chunks = n_entries / 32;
  const int64_t my = chunks > warp ? (chunks - warp + nwarps - 1) / nwarps : 0;
  double acc0 = 0.0;
  auto issue = [&](int64_t t) {              // t-th chunk of this warp -> stage t % NS
    const int s = (int)(t % NS);
    const int32_t j = __ldcs(idx + (warp + t * nwarps) * 32 + lane);
    const uint32_t dst = smem_u32(ring + ((size_t)s * 32 + lane) * ROWB);
    const uint32_t bar = smem_u32(bars + s);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWB) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(table + (int64_t)j * STRIDE), "r"(ROWB), "r"(bar) : "memory");
  };
  for (int t = 0; t < NS - 1 && t < my; ++t) issue(t);
  for (int64_t t = 0; t < my; ++t) {
    if (t + NS - 1 < my) issue(t + NS - 1);
    const int s = (int)(t % NS);
    const uint32_t bar = smem_u32(bars + s), parity = (uint32_t)((t / NS) & 1);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
    const unsigned char* st = ring + (size_t)s * 32 * ROWB;
#pragma unroll
    for (int step = 0; step < 8; ++step) {
      const unsigned char* row = st + (size_t)(step * 4 + gq) * ROWB;
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double2 v = *reinterpret_cast<const double2*>(row + 16 * (lg + 8 * r));
        d0 = fma(v.x, 1.0000001, d0); d1 = fma(v.y, 0.9999999, d1);
      }
      const double2 v = *reinterpret_cast<const double2*>(row + 16 * ((lg < 1 ? lg : 0) + 24));
      d0 = fma(v.x, 1.0000001, d0); d1 = fma(v.y, 0.9999999, d1);
      double d = d0 + d1;
      if (BUTTERFLY) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      }
      acc0 += d;
    }
    __syncwarp();                              // every lane is done reading the stage before it is refilled
  }
  if (acc0 == 1.2345e300) out[0] = acc0;
}

// ---- cp.async (LDGSTS) gather: group of 8 lanes per entry, 16 bytes per lane per slot ----------------------------------
template <int NS>
__global__ void __launch_bounds__(128) gather_cpasync(const char* __restrict__ table, const int32_t* __restrict__ idx,
                                                      int64_t n_entries, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, lg = lane & 7, gq = lane >> 3, w = threadIdx.x >> 5;
  unsigned char* ring = smem + (size_t)w * NS * 32 * ROWB;
  const int64_t warp = (int64_t)blockIdx.x * 4 + w;
  const int64_t nwarps = (int64_t)gridDim.x * 4;
  const int64_t chunks = n_entries / 32;
  const int64_t my = chunks > warp ? (chunks - warp + nwarps - 1) / nwarps : 0;
  double acc0 = 0.0;
  auto issue = [&](int64_t t) {
    const int s = (int)(t % NS);
    const int32_t j = __ldcs(idx + (warp + t * nwarps) * 32 + lane);
#pragma unroll
    for (int step = 0; step < 8; ++step) {
      const int32_t jj = __shfl_sync(0xffffffffu, j, step * 4 + gq);
      const char* src = table + (int64_t)jj * STRIDE;
      unsigned char* row = ring + ((size_t)s * 32 + step * 4 + gq) * ROWB;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(row + 16 * (lg + 8 * r))), "l"(src + 16 * (lg + 8 * r)) : "memory");
      if (lg < 1)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(row + 16 * 24)), "l"(src + 16 * 24) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int t = 0; t < NS - 1; ++t) { if (t < my) issue(t); else asm volatile("cp.async.commit_group;" ::: "memory"); }
  for (int64_t t = 0; t < my; ++t) {
    if (t + NS - 1 < my) issue(t + NS - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
    __syncwarp();
    const unsigned char* st = ring + (size_t)(t % NS) * 32 * ROWB;
#pragma unroll
    for (int step = 0; step < 8; ++step) {
      const unsigned char* row = st + (size_t)(step * 4 + gq) * ROWB;
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double2 v = *reinterpret_cast<const double2*>(row + 16 * (lg + 8 * r));
        d0 = fma(v.x, 1.0000001, d0); d1 = fma(v.y, 0.9999999, d1);
      }
      const double2 v = *reinterpret_cast<const double2*>(row + 16 * ((lg < 1 ? lg : 0) + 24));
      d0 = fma(v.x, 1.0000001, d0); d1 = fma(v.y, 0.9999999, d1);
      acc0 += d0 + d1;
    }
    __syncwarp();
  }
  if (acc0 == 1.2345e300) out[0] = acc0;
}

// ---- FP64 FMA peak ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double b = 1.0000000001, c = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345e300) out[0] = s;
}

// ---- FP64 tensor-core (DMMA m8n8k4) peak: 8 independent accumulator pairs per warp ---------------------------------------
__global__ void __launch_bounds__(256) dmma_peak(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = i; }
  const double a = 1.0000000001, b = 0.25 + 1e-12 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 1.2345e300) out[0] = s;
}

__global__ void stream_read(const double2* __restrict__ p, int64_t n, double* out) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = __ldcs(p + i);
    acc += v.x + v.y;
  }
  if (acc == 1.2345e300) out[0] = acc;
}
__global__ void stream_copy(const double2* __restrict__ p, double2* __restrict__ q, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) q[i] = __ldcs(p + i);
}

template <class F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();                                          // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  const bool only_fp64 = argc > 1 && strcmp(argv[1], "fp64") == 0;   // `microbench fp64`: just the FP64 pipe sections
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* d_out;
  CK(cudaMalloc(&d_out, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

  const int64_t n_entries = 20000256;          // ~ nnz of config 2, multiple of 32
  std::vector<int32_t> h_idx((size_t)n_entries);
  int32_t* d_idx;
  CK(cudaMalloc(&d_idx, n_entries * 4));
  for (int which = 0; which < (only_fp64 ? 0 : 2); ++which) {
    const int64_t ncols = which == 0 ? 26744 : 138493;
    uint64_t seed = 12345 + which;
    for (auto& v : h_idx) v = (int32_t)(splitmix(seed) % (uint64_t)ncols);
    CK(cudaMemcpy(d_idx, h_idx.data(), n_entries * 4, cudaMemcpyHostToDevice));
    char* d_table;
    CK(cudaMalloc(&d_table, ncols * STRIDE));
    CK(cudaMemset(d_table, 0, ncols * STRIDE));
    const double gbytes = (double)n_entries * (ROWB + 4) / 1e9;
    const char* tag = which == 0 ? "y13MB" : "x71MB";
    auto report = [&](const char* name, float ms) {
      printf(", \"%s_%s\": {\"ms\": %.4f, \"GBps\": %.1f, \"Gentries_s\": %.2f}", name, tag, ms, gbytes / (ms * 1e-3), n_entries / (ms * 1e-3) / 1e9);
      fflush(stdout);
    };
    const int grids[] = {sms * 4, sms * 8, 34174};
    float best;
    best = 1e30f;
    for (int g : grids) best = fminf(best, time_ms([&] { gather_ldg<2, false><<<g, 128>>>(d_table, d_idx, n_entries, d_out); }));
    report("ldg_d2", best);
    best = 1e30f;
    for (int g : grids) best = fminf(best, time_ms([&] { gather_ldg<4, false><<<g, 128>>>(d_table, d_idx, n_entries, d_out); }));
    report("ldg_d4", best);
    best = 1e30f;
    for (int g : grids) best = fminf(best, time_ms([&] { gather_ldg<8, false><<<g, 128>>>(d_table, d_idx, n_entries, d_out); }));
    report("ldg_d8", best);
    best = 1e30f;
    for (int g : grids) best = fminf(best, time_ms([&] { gather_ldg<2, true><<<g, 128>>>(d_table, d_idx, n_entries, d_out); }));
    report("ldg_d2_butterfly", best);
    best = 1e30f;
    for (int g : grids) best = fminf(best, time_ms([&] { gather_ldg<4, true><<<g, 128>>>(d_table, d_idx, n_entries, d_out); }));
    report("ldg_d4_butterfly", best);
    {
      const size_t sm2 = 4 * 2 * 32 * ROWB + 4 * 2 * 8, sm3 = 4 * 3 * 32 * ROWB + 4 * 3 * 8;
      CK(cudaFuncSetAttribute(gather_bulk<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      CK(cudaFuncSetAttribute(gather_bulk<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      CK(cudaFuncSetAttribute(gather_bulk<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
      CK(cudaFuncSetAttribute(gather_cpasync<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      CK(cudaFuncSetAttribute(gather_cpasync<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
      const int g2[] = {sms * 2, sms * 4}, g3[] = {sms, sms * 2};
      best = 1e30f;
      for (int g : g2) best = fminf(best, time_ms([&] { gather_bulk<2, false><<<g, 128, sm2>>>(d_table, d_idx, n_entries, d_out); }));
      report("bulk_ns2", best);
      best = 1e30f;
      for (int g : g2) best = fminf(best, time_ms([&] { gather_bulk<2, true><<<g, 128, sm2>>>(d_table, d_idx, n_entries, d_out); }));
      report("bulk_ns2_butterfly", best);
      best = 1e30f;
      for (int g : g3) best = fminf(best, time_ms([&] { gather_bulk<3, false><<<g, 128, sm3>>>(d_table, d_idx, n_entries, d_out); }));
      report("bulk_ns3", best);
      best = 1e30f;
      for (int g : g2) best = fminf(best, time_ms([&] { gather_cpasync<2><<<g, 128, sm2>>>(d_table, d_idx, n_entries, d_out); }));
      report("cpasync_ns2", best);
      best = 1e30f;
      for (int g : g3) best = fminf(best, time_ms([&] { gather_cpasync<3><<<g, 128, sm3>>>(d_table, d_idx, n_entries, d_out); }));
      report("cpasync_ns3", best);
    }
    CK(cudaFree(d_table));
  }

  {  // FP64 FMA peak
    const int iters = 4000;
    const float ms = time_ms([&] { dfma_peak<<<sms * 8, 256>>>(d_out, iters); });
    const double flops = 2.0 * 64.0 * iters * 256.0 * sms * 8;
    printf(", \"dfma_peak\": {\"ms\": %.4f, \"TFLOPs\": %.2f}", ms, flops / (ms * 1e-3) / 1e12);
  }
  {  // FP64 FMA rate with one warp per scheduler (128 threads per SM) and with two
    const int iters = 4000;
    float ms = time_ms([&] { dfma_peak<<<sms, 128>>>(d_out, iters); });
    printf(", \"dfma_4warps_per_sm\": {\"ms\": %.4f, \"TFLOPs\": %.2f}", ms, 2.0 * 64.0 * iters * 128.0 * sms / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { dfma_peak<<<sms, 256>>>(d_out, iters); });
    printf(", \"dfma_8warps_per_sm\": {\"ms\": %.4f, \"TFLOPs\": %.2f}", ms, 2.0 * 64.0 * iters * 256.0 * sms / (ms * 1e-3) / 1e12);
  }
  {  // FP64 tensor-core rate (mma.sync m8n8k4 = 256 FMA per warp instruction)
    const int iters = 2000;
    const double fl = 2.0 * 256.0 * 32.0 * iters;   // per warp
    float ms = time_ms([&] { dmma_peak<<<sms * 8, 256>>>(d_out, iters); });
    printf(", \"dmma_peak\": {\"ms\": %.4f, \"TFLOPs\": %.2f}", ms, fl * 8.0 * sms * 8 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { dmma_peak<<<sms, 128>>>(d_out, iters); });
    printf(", \"dmma_4warps_per_sm\": {\"ms\": %.4f, \"TFLOPs\": %.2f}", ms, fl * 4.0 * sms / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { dmma_peak<<<sms, 256>>>(d_out, iters); });
    printf(", \"dmma_8warps_per_sm\": {\"ms\": %.4f, \"TFLOPs\": %.2f}", ms, fl * 8.0 * sms / (ms * 1e-3) / 1e12);
  }
  if (!only_fp64) {  // HBM streaming
    const int64_t n = (int64_t)4 << 30;        // 4 GiB
    double2 *p, *q;
    CK(cudaMalloc(&p, n)); CK(cudaMalloc(&q, n));
    CK(cudaMemset(p, 0, n)); CK(cudaMemset(q, 0, n));
    float ms = time_ms([&] { stream_read<<<sms * 16, 512>>>(p, n / 16, d_out); });
    printf(", \"hbm_read\": {\"ms\": %.4f, \"GBps\": %.1f}", ms, n / 1e9 / (ms * 1e-3));
    ms = time_ms([&] { stream_copy<<<sms * 16, 512>>>(p, q, n / 16); });
    printf(", \"hbm_copy\": {\"ms\": %.4f, \"GBps_rw\": %.1f}", ms, 2.0 * n / 1e9 / (ms * 1e-3));
    // H2D from pinned memory, one 512 MiB copy and 8 concurrent 64 MiB copies on separate streams
    void* h;
    const size_t hb = (size_t)512 << 20;
    CK(cudaMallocHost(&h, hb));
    ms = time_ms([&] { CK(cudaMemcpyAsync(p, h, hb, cudaMemcpyHostToDevice, 0)); });
    printf(", \"h2d_pinned_512MiB\": {\"ms\": %.3f, \"GBps\": %.1f}", ms, hb / 1e9 / (ms * 1e-3));
    ms = time_ms([&] { CK(cudaMemcpyAsync(h, p, hb, cudaMemcpyDeviceToHost, 0)); });
    printf(", \"d2h_pinned_512MiB\": {\"ms\": %.3f, \"GBps\": %.1f}", ms, hb / 1e9 / (ms * 1e-3));
    // 2-D copy of 400-byte rows into a 512-byte pitch (what upload_factors did in round 1)
    ms = time_ms([&] { CK(cudaMemcpy2DAsync(p, 512, h, 400, 400, 138493, cudaMemcpyHostToDevice, 0)); });
    printf(", \"h2d_2d_400B_rows_55MB\": {\"ms\": %.3f, \"GBps\": %.1f}", ms, 400.0 * 138493 / 1e9 / (ms * 1e-3));
    CK(cudaFreeHost(h));
    CK(cudaFree(p)); CK(cudaFree(q));
  }
  printf("}\n");
  return 0;
}
