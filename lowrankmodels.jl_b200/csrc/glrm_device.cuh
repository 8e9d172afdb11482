// glrm_device.cuh — device side of the B200 GLRM prox-grad engine (sm_100a).
//
// One *unit* = one factor column (a row's x_e in the X sweep, a feature's y_f in the Y sweep).
// The fused update kernel does, for its unit, everything the reference does between
// proxgrad.jl:119-155 (X) / :163-200 (Y):
//   gradient pass  : gather the opposite factor's columns at the observed entries, u = x.y,
//                    loss + dloss/du (losses.jl), g += dloss * y, and the unit's current objective
//                    (row_objective / col_objective, evaluate_fit.jl:24-55) in the same pass
//   line search    : x_new = prox(x - (alpha/l) g, alpha/l) in registers (regularizers.jl), trial
//                    objective by a second gather pass, accept iff new < old (strict), alpha*1.05 /
//                    alpha*0.7 with the min_stepsize floor (proxgrad.jl:136-155)
//   write-back     : the accepted column, alpha, and the unit's recorded objective.
//
// Thread mapping.  A *lane group* of G lanes owns one observed entry at a time; lane `lg` of the group
// holds R double2 slices of the k-vector (elements 2*(lg+G*r), +1), so a group reads one factor column
// with R coalesced 16-byte loads per lane and reduces the dot product with log2(G) shuffles.  A warp
// holds 32/G groups; a unit is processed by W warps (W=1: one warp per unit, no block barrier;
// W=8: one CTA per heavy unit, cross-warp reduction through shared memory in a fixed order).
// Every reduction tree depends only on (G, R, W, degree): results are independent of which SM / GPU
// the unit lands on.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <math.h>
#include "../../include/glrm_b200.h"

namespace glrm {

constexpr unsigned FULLMASK = 0xffffffffu;
#define TRIAL_TILE_DOUBLES(G) ((G) <= 8 ? 2 * 32 * ((G) + 1) : 1)   /* two [32][G+1] partial-dot tiles per warp (narrow groups only) */
enum : int { FLAG_EVAL_ONLY = 1, FLAG_NO_REG = 2, FLAG_LOSS_BY_ENTRY = 4, FLAG_UNCONDITIONAL = 8 };

struct SweepArgs {
  // observation lists of this side (shard-local), see glrm_b200.h
  const int64_t* ptr;     // [units+1] rebased to the shard, or nullptr when fully observed
  const int32_t* idx;     // opposite index per entry, or nullptr (fully observed: idx = position)
  const double* val;      // A values aligned with the lists
  int64_t full_len;       // list length when fully observed
  int64_t unit_base;      // first unit of the shard (ptr is indexed by unit - unit_base)
  const int32_t* order;   // schedule: unit ids, heaviest first
  int64_t n_units;        // units this launch covers: order[0 .. n_units)
  double* own;            // factor being updated   [units_total * stride]
  const int64_t* own_col; // unit -> column of `own` (nullptr: identity); set when block columns shift the numbering
  const double* opp;      // factor being gathered  [opp_total * kp]
  int32_t stride;         // doubles between factor columns == 2*G*R of the tile (zero-padded past k)
  int32_t last_lanes;     // lanes of a group whose LAST slot holds real data (ceil(k/2) - G*(R-1)); the others
                          // re-read lane 0's 16 bytes there (same sector: no traffic) against a zero x slot
  int32_t k;              // rank
  const int32_t* loss_code;   // [n] per feature
  const double* loss_param;   // [n * 8]
  double uparam[3];           // uniform loss parameters (scale, p1, p2) when LOSS != 0
  const int32_t* reg_code;    // [1] or [units_total]
  const double* reg_param;    // [1*4] or [units_total*4]
  const int64_t* reg_payload_ptr;  // [1+1] or [units_total+1] offsets of the regularizers' vector payloads, or nullptr
  const double* reg_payload;
  int32_t reg_uniform;
  int32_t flags;
  double* alpha;          // [units_total] step sizes (alpharow / alphacol, proxgrad.jl:69-70)
  double min_stepsize;
  double global_alpha;    // FLAG_UNCONDITIONAL (SparseProxGradParams, sparse_proxgrad.jl:62-99): one shared step size,
                          // the prox-gradient step is taken without a line search
  double* obj_out;        // [units_total] recorded objective of the unit (obj_by_col, proxgrad.jl:178,190)
  unsigned long long* trial_counter;  // total line-search trials (profile)
  // fused exchange (multi-GPU): replicas of `own` / `obj_out` on the peer GPUs, mapped through CUDA IPC.  The
  // accepted column and the unit's objective are stored straight into every peer over NVLink from the update
  // kernel, so the all-gather overlaps the sweep and only a barrier remains between half-iterations.
  double* const* peer_own;   // [n_peers] or nullptr
  double* const* peer_obj;   // [n_peers] or nullptr
  int32_t n_peers;
  const int* stop;           // device flag set by record_kernel when the stopping rule fired (proxgrad.jl:209-213): the host
                             // enqueues iterations ahead of the device, launches after the stop are no-ops
};
__device__ __forceinline__ bool sweep_stopped(const SweepArgs& A) {
  return A.stop != nullptr && *reinterpret_cast<const volatile int*>(A.stop) != 0;
}

// ------------------------------------------------------------------------------------------------
// small helpers
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ int group_sum_i(int v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
// sum over the 32/G groups of a warp (every lane of a group holds the same value on entry)
template <int G>
__device__ __forceinline__ double cross_group_sum(double v) {
#pragma unroll
  for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ double jl_max0(double x) { return (x != x) ? x : (x > 0.0 ? x : 0.0); }   // max(x,0)
__device__ __forceinline__ double jl_min0(double x) { return (x != x) ? x : (x < 0.0 ? x : 0.0); }   // min(x,0)
__device__ __forceinline__ double jl_maxd(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
__device__ __forceinline__ double jl_mind(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
__device__ __forceinline__ double jl_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }

// ------------------------------------------------------------------------------------------------
// scalar losses (losses.jl:136-341): value and d/du in one call.  LOSS != 0 fixes the code at
// compile time (uniform-loss problems); LOSS == 0 switches on `code` (warp-uniform in the Y sweep,
// per-entry in the X sweep of heterogeneous problems).
template <int LOSS, bool WANT_GRAD>
__device__ __forceinline__ void loss_eval(int code, double s, double p1, double p2, double u, double a,
                                          double& l, double& c) {
  const int cd = LOSS ? LOSS : code;
  c = 0.0;
  switch (cd) {
    case GLRMB200_LOSS_QUAD: {                       // losses.jl:144,146
      const double d = u - a;
      l = s * (d * d);
      if (WANT_GRAD) c = 2.0 * d * s;
    } break;
    case GLRMB200_LOSS_L1: {                         // :158,160
      const double d = u - a;
      l = s * fabs(d);
      if (WANT_GRAD) c = jl_sign(d) * s;
    } break;
    case GLRMB200_LOSS_HUBER: {                      // :173-177 (grad has no factor 2, as in the reference)
      const double d = u - a, ad = fabs(d);
      l = ad > p1 ? (ad - p1 + p1 * p1) * s : d * d * s;
      if (WANT_GRAD) c = ad > p1 ? jl_sign(d) * s : d * s;
    } break;
    case GLRMB200_LOSS_QUANTILE: {                   // :193-201
      const double diff = a - u;
      l = diff > 0.0 ? s * p1 * diff : -s * (1.0 - p1) * diff;
      if (WANT_GRAD) c = diff > 0.0 ? -s * p1 : s * (1.0 - p1);
    } break;
    case GLRMB200_LOSS_PERIODIC: {                   // :216,218
      const double w = (a - u) * (2.0 * M_PI) / p1;
      l = s * (1.0 - cos(w));
      if (WANT_GRAD) c = -s * ((2.0 * M_PI) / p1) * sin(w);
    } break;
    case GLRMB200_LOSS_POISSON: {                    // :237-241
      const double eu = exp(u);
      l = s * (eu - a * u + (a == 0.0 ? 0.0 : a * (log(a) - 1.0)));
      if (WANT_GRAD) c = s * (eu - a);
    } break;
    case GLRMB200_LOSS_ORDINAL_HINGE: {              // :258-292  (p1 = min, p2 = max)
      double n, loss;
      if (u > p2 - 1.0) {
        n = jl_mind(floor(u), p2 - 1.0) - a;
        loss = n * (n + 1.0) / 2.0 + (n + 1.0) * (u - p2 + 1.0);
      } else if (u > a) {
        n = jl_mind(floor(u), p2) - a;
        loss = n * (n + 1.0) / 2.0 + (n + 1.0) * (u - floor(u));
      } else if (u > p1 + 1.0) {
        n = a - jl_maxd(ceil(u), p1 + 1.0);
        loss = n * (n + 1.0) / 2.0 + (n + 1.0) * (ceil(u) - u);
      } else {
        n = a - jl_maxd(ceil(u), p1 + 1.0);
        loss = n * (n + 1.0) / 2.0 + (n + 1.0) * (p1 + 1.0 - u);
      }
      l = s * loss;
      if (WANT_GRAD) c = s * (u > a ? (jl_mind(ceil(u), p2) - a) : -(a - jl_maxd(floor(u), p1)));
    } break;
    case GLRMB200_LOSS_LOGISTIC: {                   // :304,306 (labels: 1 -> +1, 0/-1 -> -1, :104)
      const double aa = (a == 1.0) ? 1.0 : -1.0;
      const double e = exp(-aa * u);                 // exp(-(2a-1)u)
      l = s * log(1.0 + e);
      if (WANT_GRAD) c = -aa * s / (1.0 + exp(aa * u));
    } break;
    case GLRMB200_LOSS_WEIGHTED_HINGE: {             // :326-341 (p1 = case_weight_ratio)
      const bool pos = (a == 1.0);
      const double an = pos ? 1.0 : -1.0;
      l = s * jl_max0(1.0 - an * u);
      if (WANT_GRAD) c = (an * u >= 1.0) ? 0.0 : -an * s;
      if (p1 != 1.0 && pos) { l *= p1; c *= p1; }
    } break;
    default: l = NAN; break;
  }
}

// ------------------------------------------------------------------------------------------------
// regularizers on a k-vector distributed over a lane group (regularizers.jl:52-348).
// Lane `lg` holds v[r] = elements (i0, i0+1), i0 = 2*(lg + G*r).  `kin` = number of elements the inner
// regularizer sees (k, or k-1 under the offset wrappers lastentry1 / lastentry_unpenalized :163-189).
struct ArgMax { double v; int i; };
__device__ __forceinline__ bool am_better(double av, int ai, double bv, int bi) {   // Julia argmax: first max, NaN wins
  // branch-free (predicate logic only): this sits on the critical path of every one-hot / k-sparse / simplex prox
  const bool an = av != av, bn = bv != bv, lower = ai < bi;
  const bool ordered = (av > bv) | ((av == bv) & lower);
  return an ? (!bn | lower) : (!bn & ordered);
}
template <int G>
__device__ __forceinline__ ArgMax group_argmax(ArgMax m) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(FULLMASK, m.v, o);
    const int oi = __shfl_xor_sync(FULLMASK, m.i, o);
    if (am_better(ov, oi, m.v, m.i)) { m.v = ov; m.i = oi; }
  }
  return m;
}

// `pay` / `npay`: the regularizer's vector payload (fixed_latent_features.y / fixed_last_latent_features.y: the pinned
// values, RemQuadReg.m: the mean; regularizers.jl:193-231,412-423).  The inner regularizer sees the index range [klo, khi).
template <int G, int R>
__device__ __forceinline__ double reg_eval(int code, const double* __restrict__ rp, const double2 (&v)[R], int lg, int k,
                                           const double* __restrict__ pay = nullptr, int npay = 0) {
  const int base = code & GLRMB200_REG_BASE_MASK;
  const bool wrapped = code & (GLRMB200_REG_LASTENTRY1 | GLRMB200_REG_LASTENTRY_UNPENALIZED);
  const int klo = (code & GLRMB200_REG_FIXED_FIRST) ? npay : 0;                                 // :208 evaluate(r.r, a[n+1:end])
  const int khi = wrapped ? k - 1 : ((code & GLRMB200_REG_FIXED_LAST) ? k - npay : k);          // :230 evaluate(r.r, a[1:k-n])
  // BODY is applied to every element the inner regularizer sees (klo <= index < khi); padded slots are skipped
#define GLRM_EACH(BODY)                                                                   \
  _Pragma("unroll") for (int r = 0; r < R; ++r) {                                         \
    const int i0 = 2 * (lg + G * r);                                                      \
    { const double e = v[r].x; const int i = i0; (void)i; if (i0 >= klo && i0 < khi) { BODY; } }              \
    { const double e = v[r].y; const int i = i0 + 1; (void)i; if (i0 + 1 >= klo && i0 + 1 < khi) { BODY; } }  \
  }
  double res;
  switch (base) {
    case GLRMB200_REG_REM_QUAD: {                                                               // :423
      double s2 = 0.0;
      GLRM_EACH(const double t = e - pay[i]; s2 = fma(t, t, s2))
      res = rp[0] * group_sum<G>(s2);
    } break;
    case GLRMB200_REG_ZERO: res = 0.0; break;                                                   // :95
    case GLRMB200_REG_QUAD: {                                                                   // :58
      double s2 = 0.0;
      GLRM_EACH(s2 = fma(e, e, s2))
      res = rp[0] * group_sum<G>(s2);
    } break;
    case GLRMB200_REG_QUAD_CONSTRAINT: {                                                        // :74
      double s2 = 0.0;
      GLRM_EACH(s2 = fma(e, e, s2))
      res = sqrt(group_sum<G>(s2)) > rp[0] + 1e-12 ? INFINITY : 0.0;
    } break;
    case GLRMB200_REG_ONE: {                                                                    // :88
      double sabs = 0.0;
      GLRM_EACH(sabs += fabs(e))
      res = rp[0] * group_sum<G>(sabs);
    } break;
    case GLRMB200_REG_NONNEG: {                                                                 // :105-112
      int neg = 0;
      GLRM_EACH(neg |= (e < 0.0))
      res = group_sum_i<G>(neg) ? INFINITY : 0.0;
    } break;
    case GLRMB200_REG_NONNEG_ONE: {                                                             // :129-136
      int neg = 0;
      double s1 = 0.0;
      GLRM_EACH(neg |= (e < 0.0); s1 += e)
      const int n = group_sum_i<G>(neg);
      const double t = group_sum<G>(s1);
      res = n ? INFINITY : rp[0] * t;
    } break;
    case GLRMB200_REG_ONE_SPARSE:                                                               // :239-253
    case GLRMB200_REG_KSPARSE: {                                                                // :261-276
      int nz = 0;
      GLRM_EACH(nz += (e != 0.0))
      const int lim = base == GLRMB200_REG_ONE_SPARSE ? 1 : (int)rp[0];
      res = group_sum_i<G>(nz) > lim ? INFINITY : 0.0;
    } break;
    case GLRMB200_REG_UNIT_ONE_SPARSE: {                                                        // :300-316
      int ones = 0, other = 0;
      GLRM_EACH(ones += (e == 1.0); other |= (e != 0.0 && e != 1.0))
      const int o = group_sum_i<G>(ones), x = group_sum_i<G>(other);
      res = (x || o > 1) ? INFINITY : 0.0;
    } break;
    case GLRMB200_REG_SIMPLEX: {                                                                // :338-346
      int neg = 0;
      double s1 = 0.0;
      GLRM_EACH(neg |= (e < 0.0); s1 += e)
      const double t = group_sum<G>(s1);
      const int n = group_sum_i<G>(neg);
      res = (fabs(t - 1.0) > 1e-12 || n) ? INFINITY : 0.0;
    } break;
    default: res = NAN; break;
  }
#undef GLRM_EACH
  if (code & GLRMB200_REG_LASTENTRY1) {                                                         // :171
    int badlast = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = 2 * (lg + G * r);
      if (i0 == k - 1 && v[r].x != 1.0) badlast = 1;
      if (i0 + 1 == k - 1 && v[r].y != 1.0) badlast = 1;
    }
    if (group_sum_i<G>(badlast)) res = INFINITY;
  }
  if (code & (GLRMB200_REG_FIXED_FIRST | GLRMB200_REG_FIXED_LAST)) {                             // :208 / :230: pinned entries == y, else Inf
    const int p0 = (code & GLRMB200_REG_FIXED_FIRST) ? 0 : k - npay;
    int bad = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = 2 * (lg + G * r);
      if (i0 >= p0 && i0 < p0 + npay && v[r].x != pay[i0 - p0]) bad = 1;
      if (i0 + 1 >= p0 && i0 + 1 < p0 + npay && v[r].y != pay[i0 + 1 - p0]) bad = 1;
    }
    if (group_sum_i<G>(bad)) res = INFINITY;
  }
  return res;
}

// v'[j] = v[j + n] on a lane-distributed vector (elements shifted in past the end are zero): what
// fixed_last_latent_features' prox feeds its inner regularizer (regularizers.jl:223, u[(r.n+1):end]).  Element j lives in
// flattened slot p = j / 2 = lg + G * r; every source lane works out which of its slots its reader needs.
template <int G, int R>
__device__ __forceinline__ void shift_left(double2 (&v)[R], int n, int lg) {
  const int q = n >> 1, o = n & 1;
  const int gbase = (threadIdx.x & 31) - lg;
  double2 out[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int dq = q + (h ? o : 0);                 // slots to the right of the reader
      const int hs = h ? 1 - o : o;                   // half of the source slot
      const int src_lg = (lg + dq) % G;               // lane (within the group) this reader fetches from
      const int lg_t = ((lg - dq) % G + G) % G;       // the reader that fetches from THIS lane ...
      const int rs = r + (lg_t + dq) / G;             // ... wants this slot of it
      double supply = 0.0;
#pragma unroll
      for (int rr = 0; rr < R; ++rr) if (rr == rs) supply = hs ? v[rr].y : v[rr].x;
      const double got = __shfl_sync(FULLMASK, supply, gbase + src_lg);
      if (h) out[r].y = got; else out[r].x = got;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = out[r];
}

template <int G, int R>
__device__ __forceinline__ void reg_prox(int code, const double* __restrict__ rp, double2 (&v)[R], int lg, int k, double alpha,
                                         const double* __restrict__ pay = nullptr, int npay = 0) {
  const int base = code & GLRMB200_REG_BASE_MASK;
  const bool wrapped = code & (GLRMB200_REG_LASTENTRY1 | GLRMB200_REG_LASTENTRY_UNPENALIZED);
  // fixed_last_latent_features, literally (:223): [prox(r.r, u[(n+1):end], alpha); y] — the inner prox sees the LAST k-n
  // entries, its result becomes the FIRST k-n
  if (code & GLRMB200_REG_FIXED_LAST) shift_left<G, R>(v, npay, lg);
  const int klo = (code & GLRMB200_REG_FIXED_FIRST) ? npay : 0;                                 // :203 prox(r.r, u[(n+1):end])
  const int khi = wrapped ? k - 1 : ((code & GLRMB200_REG_FIXED_LAST) ? k - npay : k);
  const int kin = khi - klo;                                                                    // length the inner regularizer sees
#define GLRM_FOREACH(BODY)                                              \
  _Pragma("unroll") for (int r = 0; r < R; ++r) {                       \
    const int i0 = 2 * (lg + G * r);                                    \
    { double& e = v[r].x; const int i = i0; if (i >= klo && i < khi) { BODY; } }     \
    { double& e = v[r].y; const int i = i0 + 1; if (i >= klo && i < khi) { BODY; } } \
  }
  switch (base) {
    case GLRMB200_REG_REM_QUAD: {                                                               // :417-418
      const double c2 = 2.0 * alpha * rp[0];
      GLRM_FOREACH(e = (e + c2 * pay[i]) / (1.0 + c2))
    } break;
    case GLRMB200_REG_ZERO: break;                                                              // :93
    case GLRMB200_REG_QUAD: {                                                                   // :56
      const double c = 1.0 / (1.0 + 2.0 * alpha * rp[0]);
      GLRM_FOREACH(e = c * e; (void)i)
    } break;
    case GLRMB200_REG_QUAD_CONSTRAINT: {                                                        // :72
      double s2 = 0.0;
      GLRM_FOREACH(s2 += e * e; (void)i)
      const double c = rp[0] / sqrt(group_sum<G>(s2));
      GLRM_FOREACH(e = c * e; (void)i)
    } break;
    case GLRMB200_REG_ONE: {                                                                    // :83-87
      const double t = rp[0] * alpha;
      GLRM_FOREACH(e = jl_max0(e - t) + jl_min0(e + t); (void)i)
    } break;
    case GLRMB200_REG_NONNEG: GLRM_FOREACH(e = jl_max0(e); (void)i) break;                      // :103
    case GLRMB200_REG_NONNEG_ONE: GLRM_FOREACH(e = jl_max0(e - alpha); (void)i) break;          // :122
    case GLRMB200_REG_ONE_SPARSE:                                                               // :237
    case GLRMB200_REG_UNIT_ONE_SPARSE: {                                                        // :297
      ArgMax m{-INFINITY, 0x7fffffff};
      GLRM_FOREACH(if (am_better(e, i, m.v, m.i)) { m.v = e; m.i = i; })
      m = group_argmax<G>(m);
      const bool unit = base == GLRMB200_REG_UNIT_ONE_SPARSE;
      GLRM_FOREACH(e = (i == m.i) ? (unit ? 1.0 : e) : 0.0)
    } break;
    case GLRMB200_REG_KSPARSE: {                                                                // :277-283
      // keep the kk entries of largest |v| (ties: lowest index first): kk rounds of group arg-max
      const int kk = (int)rp[0];
      unsigned keep = 0;  // bit 2r+h of this lane
      for (int round = 0; round < kk && round < kin; ++round) {
        ArgMax m{-INFINITY, 0x7fffffff};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i0 = 2 * (lg + G * r);
          if (i0 >= klo && i0 < khi && !(keep >> (2 * r) & 1u) && am_better(fabs(v[r].x), i0, m.v, m.i)) { m.v = fabs(v[r].x); m.i = i0; }
          if (i0 + 1 >= klo && i0 + 1 < khi && !(keep >> (2 * r + 1) & 1u) && am_better(fabs(v[r].y), i0 + 1, m.v, m.i)) { m.v = fabs(v[r].y); m.i = i0 + 1; }
        }
        m = group_argmax<G>(m);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i0 = 2 * (lg + G * r);
          if (m.i == i0) keep |= 1u << (2 * r);
          if (m.i == i0 + 1) keep |= 1u << (2 * r + 1);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i0 = 2 * (lg + G * r);
        if (i0 >= klo && i0 < khi && !(keep >> (2 * r) & 1u)) v[r].x = 0.0;
        if (i0 + 1 >= klo && i0 + 1 < khi && !(keep >> (2 * r + 1) & 1u)) v[r].y = 0.0;
      }
    } break;
    case GLRMB200_REG_SIMPLEX: {                                                                // :325-337
      // walk the entries in descending order (group arg-max per step) instead of sorting:
      // t = (ysum[i]-1)/i at the first i with (ysum[i]-1)/i >= y[i+1], else (ysum[n]-1)/n
      unsigned used = 0;
      double ysum = 0.0, t = 0.0;
      bool found = false;
      for (int step = 0; step < kin; ++step) {
        ArgMax m{-INFINITY, 0x7fffffff};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i0 = 2 * (lg + G * r);
          if (i0 >= klo && i0 < khi && !(used >> (2 * r) & 1u) && am_better(v[r].x, i0, m.v, m.i)) { m.v = v[r].x; m.i = i0; }
          if (i0 + 1 >= klo && i0 + 1 < khi && !(used >> (2 * r + 1) & 1u) && am_better(v[r].y, i0 + 1, m.v, m.i)) { m.v = v[r].y; m.i = i0 + 1; }
        }
        m = group_argmax<G>(m);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i0 = 2 * (lg + G * r);
          if (m.i == i0) used |= 1u << (2 * r);
          if (m.i == i0 + 1) used |= 1u << (2 * r + 1);
        }
        // m.v is y[step+1] (1-based); test the candidate built from the first `step` entries
        if (step > 0 && !found && (ysum - 1.0) / (double)step >= m.v) { t = (ysum - 1.0) / (double)step; found = true; }
        ysum += m.v;
      }
      if (!found) t = (ysum - 1.0) / (double)kin;
      GLRM_FOREACH(e = jl_max0(e - t); (void)i)
    } break;
    default: break;
  }
#undef GLRM_FOREACH
  if (code & GLRMB200_REG_LASTENTRY1) {                                                         // :168
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = 2 * (lg + G * r);
      if (i0 == k - 1) v[r].x = 1.0;
      if (i0 + 1 == k - 1) v[r].y = 1.0;
    }
  }
  if (code & (GLRMB200_REG_FIXED_FIRST | GLRMB200_REG_FIXED_LAST)) {                             // :203 / :223: the pinned entries
    const int p0 = (code & GLRMB200_REG_FIXED_FIRST) ? 0 : k - npay;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = 2 * (lg + G * r);
      if (i0 >= p0 && i0 < p0 + npay) v[r].x = pay[i0 - p0];
      if (i0 + 1 >= p0 && i0 + 1 < p0 + npay) v[r].y = pay[i0 + 1 - p0];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One pass over the unit's observed entries.  GRAD: accumulate g and the objective; otherwise the
// objective only (line-search trial).
//
// Memory pipeline.  Each warp walks its share of the list in chunks of 32 entries: lane l loads
// (idx, val) of entry l with one coalesced streaming load, one chunk ahead of use, and the groups pick
// their entry up by shuffle — the index load never sits in front of the gather it feeds.  The gathers
// themselves are double-buffered: the R 16-byte loads of step s+1 are in flight while step s is reduced,
// so every lane group always has one factor column on its way from L2.  The loop body is branch-free.
template <int G, int R, int W, int LOSS, bool GRAD, int PIPE_DEPTH>
__device__ __forceinline__ void entry_pass(const SweepArgs& A, int64_t start, int64_t len, int warp, int lane,
                                           const double2 (&x)[R], int ucode, double us,
                                           double up1, double up2, double2 (&g)[R], double& obj) {
  constexpr int NGW = 32 / G;
  const int lg = lane % G, gq = lane / G;
  const bool by_entry = (LOSS == 0) && (A.flags & FLAG_LOSS_BY_ENTRY);
  const uint32_t ulen = (uint32_t)len;
  const uint32_t nchunks = (ulen + 31u) >> 5;
  const uint32_t my_chunks = (nchunks + W - 1) / W;   // same trip count in every warp of the unit
  const uint32_t nsteps = my_chunks * G;
  const char* opp_lane = reinterpret_cast<const char*>(A.opp + 2 * lg);
  const char* opp_last = reinterpret_cast<const char*>(A.opp + 2 * ((lg < A.last_lanes ? lg : 0) + G * (R - 1)));
  const int stride_bytes = A.stride * 8;
  obj = 0.0;
  if (GRAD) {
#pragma unroll
    for (int r = 0; r < R; ++r) g[r] = make_double2(0.0, 0.0);
  }
  auto load_chunk = [&](uint32_t ci, int32_t& j, double& a) {
    const uint32_t t = ((ci * W + warp) << 5) + lane;
    const bool ok = ci < my_chunks && t < ulen;
    const int64_t q = start + (ok ? t : 0u);
    const int32_t jr = A.idx ? __ldcs(A.idx + q) : (int32_t)t;
    const double ar = __ldcs(A.val + q);
    j = ok ? jr : -1;
    a = ok ? ar : 0.0;
  };
  struct Ent { int32_t j; double a; int code; double s, p1, p2; uint32_t cs; };
  auto fetch = [&](uint32_t s, int32_t cj, double ca, double2 (&y)[R], Ent& e) {
    e.cs = s & (G - 1);
    const int src = (int)(s & (G - 1)) * NGW + gq;
    e.j = __shfl_sync(FULLMASK, cj, src);
    e.a = __shfl_sync(FULLMASK, ca, src);
    const int32_t jj = e.j < 0 ? 0 : e.j;          // inactive slots read column 0 and are masked below
    // R unconditional 16-byte loads per lane with immediate offsets; in the last slot only the lanes that
    // hold real elements read their own 16 bytes
    const char* yp = opp_lane + (int64_t)jj * stride_bytes;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) y[r] = __ldg(reinterpret_cast<const double2*>(yp + r * G * 16));
    y[R - 1] = __ldg(reinterpret_cast<const double2*>(opp_last + (int64_t)jj * stride_bytes));
    e.code = ucode; e.s = us; e.p1 = up1; e.p2 = up2;
    if (by_entry) {
      e.code = __ldg(A.loss_code + jj);
      const double* lp = A.loss_param + (int64_t)jj * GLRMB200_LOSS_NPARAM;
      e.s = __ldg(lp); e.p1 = __ldg(lp + 1); e.p2 = __ldg(lp + 2);
    }
  };
  auto consume = [&](const double2 (&y)[R], const Ent& e) {
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { d0 = fma(y[r].x, x[r].x, d0); d1 = fma(y[r].y, x[r].y, d1); }
    const double dot = group_sum<G>(d0 + d1);
    double l, c;
    loss_eval<LOSS, GRAD>(e.code, e.s, e.p1, e.p2, dot, e.a, l, c);
    // objective bookkeeping mirrors trial_pass bit for bit: one accumulator per chunk slot (here it lives
    // in lane lg == step-in-chunk of the group), summed over chunks in order, then one fixed warp tree.
    const bool act = e.j >= 0;
    obj += (act && lg == (int)e.cs) ? l : 0.0;
    if (GRAD) {
      c = act ? c : 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) { g[r].x = fma(c, y[r].x, g[r].x); g[r].y = fma(c, y[r].y, g[r].y); }
    }
  };
  // software pipeline of depth D: D-1 steps of gathers are always in flight per lane group
  // (in-flight entries per SM = warps x groups x (D-1); with ~600-cycle loaded L2 latency this, not
  // instruction issue, bounds the sweep — see DESIGN.md "Little's law").
  constexpr int D = PIPE_DEPTH;
  int32_t fj, nj, mj;        // (idx, val) of the chunk the fetch stream is in and of the two chunks after it:
  double fa, na, ma;         // the index stream comes from DRAM, so it is requested two chunks (64 entries) early
  load_chunk(0, fj, fa);
  load_chunk(1, nj, na);
  load_chunk(2, mj, ma);
  double2 y[D][R];
  Ent e[D];
  auto fetch_step = [&](uint32_t f, double2 (&yy)[R], Ent& ee) {
    if (f != 0 && (f & (G - 1)) == 0) {           // the fetch stream enters the next chunk
      fj = nj; fa = na;
      nj = mj; na = ma;
      load_chunk(f / G + 2, mj, ma);
    }
    fetch(f, fj, fa, yy, ee);
  };
#pragma unroll
  for (int u = 0; u < D - 1; ++u) fetch_step(u, y[u], e[u]);
  for (uint32_t s = 0; s < nsteps; s += D) {      // nsteps is a multiple of G, G of D
#pragma unroll
    for (int u = 0; u < D; ++u) {
      fetch_step(s + u + D - 1, y[(u + D - 1) % D], e[(u + D - 1) % D]);
      consume(y[u], e[u]);
    }
  }
  // slot t = cs*NGW + gq was accumulated in lane gq*G + cs: bring it to lane t, then the same tree as trial_pass
  obj = __shfl_sync(FULLMASK, obj, (lane % NGW) * G + lane / NGW);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) obj += __shfl_xor_sync(FULLMASK, obj, o);
}

// Line-search trial pass: objective only.  No gradient means the gathered column is dead once its
// partial dot product exists, so the cross-lane reduction is taken off the critical path: every lane
// stores its partial (one STS per step) into a per-warp [32 entries][G lanes] tile, and once the 32
// entries of a chunk are in, lane l finishes entry l — sums its G partials in a fixed order, evaluates
// the loss with the A value it already holds from the coalesced chunk load, and accumulates.  No
// shuffles, 32 losses per instruction stream (vs 32/G), and the gathers of the next chunk are already
// in flight while a chunk is being finished.
template <int G, int R, int W, int LOSS, int PIPE_DEPTH>
__device__ __forceinline__ double trial_pass(const SweepArgs& A, int64_t start, int64_t len, int warp, int lane,
                                             const double2 (&x)[R], int ucode, double us, double up1,
                                             double up2, double* part) {
  constexpr int NGW = 32 / G;
  constexpr int ROW = G + 1;                       // odd row pitch: conflict-free transposed reads
  constexpr int D = PIPE_DEPTH;
  const int lg = lane % G, gq = lane / G;
  const bool by_entry = (LOSS == 0) && (A.flags & FLAG_LOSS_BY_ENTRY);
  const uint32_t ulen = (uint32_t)len;
  const uint32_t nchunks = (ulen + 31u) >> 5;
  const uint32_t my_chunks = (nchunks + W - 1) / W;
  const uint32_t nsteps = my_chunks * G;
  const char* opp_lane = reinterpret_cast<const char*>(A.opp + 2 * lg);
  const char* opp_last = reinterpret_cast<const char*>(A.opp + 2 * ((lg < A.last_lanes ? lg : 0) + G * (R - 1)));
  const int stride_bytes = A.stride * 8;
  double obj = 0.0;
  auto load_chunk = [&](uint32_t ci, int32_t& j, double& a) {
    const uint32_t t = ((ci * W + warp) << 5) + lane;
    const bool ok = ci < my_chunks && t < ulen;
    const int64_t q = start + (ok ? t : 0u);
    const int32_t jr = A.idx ? __ldcs(A.idx + q) : (int32_t)t;
    const double ar = __ldcs(A.val + q);
    j = ok ? jr : -1;
    a = ok ? ar : 0.0;
  };
  int32_t cj, nj, mj;   // chunk being consumed and the two after it: lane l holds entry l of the chunk
  double ca, na, ma;
  load_chunk(0, cj, ca);
  load_chunk(1, nj, na);
  load_chunk(2, mj, ma);
  uint32_t cons_chunk = 0;
  auto fetch = [&](uint32_t f, double2 (&y)[R]) {
    const int32_t sj = (f / G == cons_chunk) ? cj : nj;       // the fetch stream runs < G steps ahead
    const int src = (int)(f & (G - 1)) * NGW + gq;
    const int32_t j = __shfl_sync(FULLMASK, sj, src);
    const int64_t joff = (int64_t)(j < 0 ? 0 : j) * stride_bytes;
    const char* yp = opp_lane + joff;
#pragma unroll
    for (int r = 0; r < R - 1; ++r) y[r] = __ldg(reinterpret_cast<const double2*>(yp + r * G * 16));
    y[R - 1] = __ldg(reinterpret_cast<const double2*>(opp_last + joff));
  };
  auto consume = [&](uint32_t c, const double2 (&y)[R]) {
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { d0 = fma(y[r].x, x[r].x, d0); d1 = fma(y[r].y, x[r].y, d1); }
    const uint32_t cs = c & (G - 1);
    double* buf = part + ((c / G) & 1u) * (32 * ROW);
    buf[(cs * NGW + gq) * ROW + lg] = d0 + d1;
    if (cs == G - 1) {                              // chunk complete: lane l finishes entry l
      __syncwarp();
      // same association as group_sum's xor butterfly, so a point evaluated by the gradient pass and by a
      // trial pass yields identical bits (the reference's strict `<` relies on f(x) == f(x))
      const double* row = buf + lane * ROW;
      double v[G];
#pragma unroll
      for (int i = 0; i < G; ++i) v[i] = row[i];
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < o; ++i) v[i] += v[i + o];
      }
      const double u = v[0];
      int code = ucode;
      double ls = us, p1 = up1, p2 = up2;
      if (by_entry) {
        const int32_t jj = cj < 0 ? 0 : cj;
        code = __ldg(A.loss_code + jj);
        const double* lp = A.loss_param + (int64_t)jj * GLRMB200_LOSS_NPARAM;
        ls = __ldg(lp); p1 = __ldg(lp + 1); p2 = __ldg(lp + 2);
      }
      double l, cdummy;
      loss_eval<LOSS, false>(code, ls, p1, p2, u, ca, l, cdummy);
      obj += (cj >= 0) ? l : 0.0;
      cj = nj; ca = na;                             // consumption moves to the next chunk
      nj = mj; na = ma;
      ++cons_chunk;
      load_chunk(cons_chunk + 2, mj, ma);
    }
  };
  double2 y[D][R];
#pragma unroll
  for (int u = 0; u < D - 1; ++u) fetch(u, y[u]);
  for (uint32_t s = 0; s < nsteps; s += D) {
#pragma unroll
    for (int u = 0; u < D; ++u) {
      fetch(s + u + D - 1, y[(u + D - 1) % D]);
      consume(s + u, y[u]);
    }
  }
  // warp total in a fixed order (all lanes end with the same bits)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) obj += __shfl_xor_sync(FULLMASK, obj, o);
  return obj;
}

// sum of the per-warp objective totals in a fixed order (W > 1); every thread gets the same bits
template <int W>
__device__ __forceinline__ double block_sum_obj(double obj, double* red, int lane, int warp) {
  if (W == 1) return obj;
  __syncthreads();
  if (lane == 0) red[warp] = obj;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < W; ++w) t += red[w];
  return t;
}

// reduce g over all groups of the unit: shuffles inside the warp, shared memory across warps
template <int G, int R, int W>
__device__ __forceinline__ void unit_reduce_g(double2 (&g)[R], double* red, int lane, int warp, int lg) {
#pragma unroll
  for (int r = 0; r < R; ++r) { g[r].x = cross_group_sum<G>(g[r].x); g[r].y = cross_group_sum<G>(g[r].y); }
  if (W > 1) {
    constexpr int STRIDE = G * 2 * R;
    __syncthreads();
    if (lane < G) {
#pragma unroll
      for (int r = 0; r < R; ++r) { red[warp * STRIDE + (lg * R + r) * 2] = g[r].x; red[warp * STRIDE + (lg * R + r) * 2 + 1] = g[r].y; }
    }
    __syncthreads();
    double2 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_double2(0.0, 0.0);
    for (int w = 0; w < W; ++w) {   // fixed order
#pragma unroll
      for (int r = 0; r < R; ++r) { acc[r].x += red[w * STRIDE + (lg * R + r) * 2]; acc[r].y += red[w * STRIDE + (lg * R + r) * 2 + 1]; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) g[r] = acc[r];
  }
}

// ---- thread-block-cluster tier: CS CTAs share one super-heavy unit ---------------------------------------------
// Partial results are exchanged through distributed shared memory and summed in cluster-rank order, so the
// result does not depend on which CTA finishes first (nor on the GPU the unit is scheduled on).
template <int CS>
__device__ __forceinline__ double cluster_sum_obj(double v, double* slot) {
  if (CS == 1) return v;
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) *slot = v;
  cl.sync();
  double t = 0.0;
  for (int r = 0; r < CS; ++r) t += *cl.map_shared_rank(slot, r);
  cl.sync();
  return t;
}
template <int G, int R, int CS>
__device__ __forceinline__ void cluster_sum_g(double2 (&g)[R], double2* buf, int gid, int lg) {
  if (CS == 1) return;
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  if (gid == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) buf[r * G + lg] = g[r];
  }
  cl.sync();
  if (gid == 0) {
    double2 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_double2(0.0, 0.0);
    for (int c = 0; c < CS; ++c) {
      const double2* rb = cl.map_shared_rank(buf, c);
#pragma unroll
      for (int r = 0; r < R; ++r) { const double2 v = rb[r * G + lg]; acc[r].x += v.x; acc[r].y += v.y; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) g[r] = acc[r];
  }
  cl.sync();
}

// pipeline depth / residency per tile: DEPTH gather buffers (R double2 each) in the gradient pass, TRIAL_DEPTH
// in the trial passes (x and g live in shared memory there, so the registers go to in-flight gathers)
#ifndef GLRM_PIPE_DEPTH
#define GLRM_PIPE_DEPTH 2
#endif
#ifndef GLRM_TRIAL_DEPTH
#define GLRM_TRIAL_DEPTH 4
#endif
#ifndef GLRM_LIGHT_CTAS
#define GLRM_LIGHT_CTAS 4      /* resident 4-warp CTAs per SM the warp-tier kernel is compiled for (register cap) */
#endif
// the CTA / cluster tiers run the units whose latency bounds a sharded sweep: their depth / residency are separate knobs
#ifndef GLRM_HEAVY_DEPTH
#define GLRM_HEAVY_DEPTH 2
#endif
#ifndef GLRM_HEAVY_TRIAL_DEPTH
#define GLRM_HEAVY_TRIAL_DEPTH 4
#endif
#ifndef GLRM_HEAVY_CTAS
#define GLRM_HEAVY_CTAS 2
#endif
template <int R> struct TileCfg {
  static constexpr int DEPTH = GLRM_PIPE_DEPTH;
  static constexpr int TRIAL_DEPTH = GLRM_TRIAL_DEPTH;
  static constexpr int LIGHT_CTAS = GLRM_LIGHT_CTAS;
  static constexpr int HEAVY_DEPTH = GLRM_HEAVY_DEPTH;
  static constexpr int HEAVY_TRIAL_DEPTH = GLRM_HEAVY_TRIAL_DEPTH;
  static constexpr int HEAVY_CTAS = GLRM_HEAVY_CTAS;
};

template <int G, int R, int W, int LOSS, int DEPTH, int CS = 1, int TDEPTH = TileCfg<R>::TRIAL_DEPTH>
__device__ __forceinline__ void process_unit(const SweepArgs& A, int64_t unit, double* red, double* part, double* xg, double* clbuf = nullptr) {
  constexpr int NGW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int warp = (W == 1) ? 0 : (threadIdx.x >> 5);
  const int lg = lane % G;
  const int gid = warp * NGW + lane / G;
  const int k = A.k;
  // with a cluster, the unit's chunks are dealt round-robin over all CS*W warps
  constexpr int WT = W * CS;
  int crank = 0;
  if (CS > 1) crank = (int)cooperative_groups::this_cluster().block_rank();
  const int gwarp = crank * W + warp;

  int64_t start, len;
  if (A.ptr) {
    start = A.ptr[unit - A.unit_base];
    len = A.ptr[unit - A.unit_base + 1] - start;
  } else {
    start = (unit - A.unit_base) * A.full_len;
    len = A.full_len;
  }
  double* own = A.own + (A.own_col ? A.own_col[unit] : unit) * (int64_t)A.stride;
  // columns are stored with stride 2*G*R doubles, zero past k: every lane owns R real slots
  double2 x[R];
#pragma unroll
  for (int r = 0; r < R; ++r) x[r] = *reinterpret_cast<const double2*>(own + 2 * (lg + G * r));
  // loss descriptor: uniform (template / by value), per unit (Y sweep), or per entry (X sweep)
  int ucode = LOSS;
  double us = A.uparam[0], up1 = A.uparam[1], up2 = A.uparam[2];
  if (LOSS == 0 && !(A.flags & FLAG_LOSS_BY_ENTRY)) {
    ucode = A.loss_code[unit];
    const double* lp = A.loss_param + unit * GLRMB200_LOSS_NPARAM;
    us = lp[0]; up1 = lp[1]; up2 = lp[2];
  }
  const int rcode = A.reg_code[A.reg_uniform ? 0 : unit];
  const double* rp = A.reg_param + (A.reg_uniform ? 0 : unit) * GLRMB200_REG_NPARAM;
  const double* pay = nullptr;
  int npay = 0;
  if (A.reg_payload_ptr) {
    const int64_t p0 = A.reg_payload_ptr[A.reg_uniform ? 0 : unit];
    pay = A.reg_payload + p0;
    npay = (int)(A.reg_payload_ptr[(A.reg_uniform ? 0 : unit) + 1] - p0);
  }
  const bool use_reg = !(A.flags & FLAG_NO_REG);

  // ---- gradient pass (proxgrad.jl:119-135 / :163-178) ----------------------------------------
  double2 g[R];
  double obj_old;
  entry_pass<G, R, WT, LOSS, true, DEPTH>(A, start, len, gwarp, lane, x, ucode, us, up1, up2, g, obj_old);
  obj_old = block_sum_obj<W>(obj_old, red, lane, warp);
  obj_old = cluster_sum_obj<CS>(obj_old, clbuf);
  unit_reduce_g<G, R, W>(g, red, lane, warp, lg);
  cluster_sum_g<G, R, CS>(g, reinterpret_cast<double2*>(clbuf + 2), gid, lg);
  if (use_reg) obj_old += reg_eval<G, R>(rcode, rp, x, lg, k, pay, npay);

  const bool uncond = A.flags & FLAG_UNCONDITIONAL;
  double alpha = uncond ? A.global_alpha : A.alpha[unit];
  double obj_rec = obj_old;
  int ntrials = 0;
  if (!(A.flags & FLAG_EVAL_ONLY) && (uncond || alpha > A.min_stepsize)) {
    // x and g are only needed to form trial points: park them in shared memory so the trial passes can
    // spend the registers on a deeper gather pipeline
    double2* xs = reinterpret_cast<double2*>(xg);
    double2* gs = xs + G * R;
    if (W > 1) __syncthreads();
    if (gid == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) { xs[r * G + lg] = x[r]; gs[r * G + lg] = g[r]; }
      if (lg >= A.last_lanes) gs[(R - 1) * G + lg] = make_double2(0.0, 0.0);   // duplicates gathered against a zero x slot
    }
    if (W > 1) __syncthreads(); else __syncwarp();
    const double l1 = (double)(len + 1);                                 // proxgrad.jl:134
    while (uncond || alpha > A.min_stepsize) {                           // :136
      const double stepsize = alpha / l1;                                // :137
      double2 xn[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double2 xr = xs[r * G + lg], gr = gs[r * G + lg];
        xn[r].x = fma(-stepsize, gr.x, xr.x); xn[r].y = fma(-stepsize, gr.y, xr.y);                   // :140
      }
      reg_prox<G, R>(rcode, rp, xn, lg, k, stepsize, pay, npay);         // :142
      if (uncond) {                                                      // sparse_proxgrad.jl:73-78 / :93-98: no trial
        if (gid == 0 && crank == 0) {
          const int64_t off = own - A.own;
#pragma unroll
          for (int r = 0; r < R; ++r) *reinterpret_cast<double2*>(own + 2 * (lg + G * r)) = xn[r];
          if (A.peer_own) {
            for (int p = 0; p < A.n_peers; ++p) {
              double* po = A.peer_own[p] + off;
#pragma unroll
              for (int r = 0; r < R; ++r) *reinterpret_cast<double2*>(po + 2 * (lg + G * r)) = xn[r];
            }
          }
        }
        break;
      }
      double obj_new;
      if constexpr (G <= 8) {                       // shared-memory transposed reduction (k <= 64)
        obj_new = trial_pass<G, R, WT, LOSS, (TDEPTH <= G ? TDEPTH : G)>(
            A, start, len, gwarp, lane, xn, ucode, us, up1, up2,
            part + (W == 1 ? (threadIdx.x >> 5) : warp) * TRIAL_TILE_DOUBLES(G));
        obj_new = block_sum_obj<W>(obj_new, red, lane, warp);
        obj_new = cluster_sum_obj<CS>(obj_new, clbuf);
      } else {                                      // wide groups: shuffle-reduced pass
        double2 dummy[R];
        entry_pass<G, R, WT, LOSS, false, DEPTH>(A, start, len, gwarp, lane, xn, ucode, us, up1, up2, dummy, obj_new);
        obj_new = block_sum_obj<W>(obj_new, red, lane, warp);
        obj_new = cluster_sum_obj<CS>(obj_new, clbuf);
      }
      obj_new += reg_eval<G, R>(rcode, rp, xn, lg, k, pay, npay);
      ++ntrials;
      if (obj_new < obj_old) {                                           // :143 (strict; NaN rejects)
        if (gid == 0 && crank == 0) {
#pragma unroll
          for (int r = 0; r < R; ++r) *reinterpret_cast<double2*>(own + 2 * (lg + G * r)) = xn[r];   // :144
          if (A.peer_own) {                                              // same column into every peer's replica
            const int64_t off = own - A.own;
            for (int p = 0; p < A.n_peers; ++p) {
              double* po = A.peer_own[p] + off;
#pragma unroll
              for (int r = 0; r < R; ++r) *reinterpret_cast<double2*>(po + 2 * (lg + G * r)) = xn[r];
            }
          }
        }
        alpha *= 1.05;                                                   // :145
        obj_rec = obj_new;                                               // :190
        break;
      } else {
        alpha *= .7;                                                     // :149
        if (alpha < A.min_stepsize) { alpha = A.min_stepsize * 1.1; break; }   // :150-153
      }
    }
  }
  if (gid == 0 && lg == 0 && crank == 0) {
    if (!uncond) A.alpha[unit] = alpha;
    A.obj_out[unit] = obj_rec;
    if (A.peer_obj) for (int p = 0; p < A.n_peers; ++p) A.peer_obj[p][unit] = obj_rec;
    if (ntrials && A.trial_counter) atomicAdd(A.trial_counter, (unsigned long long)ntrials);
  }
  // (no per-unit system fence: measured +20 % on the X sweep at N=2.  Consumers of the peer stores are kernels
  //  launched after a stream-ordered NCCL barrier that follows this kernel's completion on every rank.)
}

constexpr int WARPS_PER_CTA_LIGHT = 4;
constexpr int WARPS_PER_CTA_HEAVY = 8;

// light units: one warp per unit, no block-level synchronisation
template <int G, int R, int LOSS>
__global__ void __launch_bounds__(WARPS_PER_CTA_LIGHT * 32, TileCfg<R>::LIGHT_CTAS) sweep_warp_kernel(const SweepArgs A) {
  const int64_t slot = (int64_t)blockIdx.x * WARPS_PER_CTA_LIGHT + (threadIdx.x >> 5);
  __shared__ double part[WARPS_PER_CTA_LIGHT * TRIAL_TILE_DOUBLES(G)];
  __shared__ __align__(16) double xg[WARPS_PER_CTA_LIGHT * 4 * G * R];
  if (slot >= A.n_units || sweep_stopped(A)) return;
  process_unit<G, R, 1, LOSS, TileCfg<R>::DEPTH>(A, A.order[slot], nullptr, part, xg + (threadIdx.x >> 5) * 4 * G * R);
}

// heavy units: one CTA (8 warps) per unit
template <int G, int R, int LOSS>
__global__ void __launch_bounds__(WARPS_PER_CTA_HEAVY * 32, TileCfg<R>::HEAVY_CTAS) sweep_cta_kernel(const SweepArgs A) {
  __shared__ double red[WARPS_PER_CTA_HEAVY * (G * 2 * R + 1)];
  __shared__ double part[WARPS_PER_CTA_HEAVY * TRIAL_TILE_DOUBLES(G)];
  __shared__ __align__(16) double xg[4 * G * R];
  if (sweep_stopped(A)) return;
  process_unit<G, R, WARPS_PER_CTA_HEAVY, LOSS, (TileCfg<R>::HEAVY_DEPTH <= G ? TileCfg<R>::HEAVY_DEPTH : G), 1, TileCfg<R>::HEAVY_TRIAL_DEPTH>(A, A.order[blockIdx.x], red, part, xg);
}

// super-heavy units: a cluster of CS CTAs (8 warps each) per unit — CLUSTER_CTAS for degrees >= cluster_threshold,
// CLUSTER_CTAS_BIG for degrees >= cluster16_threshold (16-CTA clusters were measured slower: 128 warps leave each warp
// 4-8 chunks per pass, and clusters of a non-portable size schedule poorly next to the warp tier).  On one GPU these tiers cost
// a few % (per-pass pipeline start-up is amortised over fewer chunks per warp); sharded, they are what keeps the heaviest
// columns from becoming the critical path of the sweep.
constexpr int CLUSTER_CTAS = 4;
constexpr int CLUSTER_CTAS_BIG = 8;
template <int G, int R, int LOSS, int CS>
__global__ void __launch_bounds__(WARPS_PER_CTA_HEAVY * 32, TileCfg<R>::HEAVY_CTAS) sweep_cluster_kernel(const SweepArgs A) {
  __shared__ double red[WARPS_PER_CTA_HEAVY * (G * 2 * R + 1)];
  __shared__ double part[WARPS_PER_CTA_HEAVY * TRIAL_TILE_DOUBLES(G)];
  __shared__ __align__(16) double xg[4 * G * R];
  __shared__ __align__(16) double clbuf[2 + 2 * G * R];
  if (sweep_stopped(A)) return;                // the same value in every CTA of the cluster: the flag only changes between kernels
  process_unit<G, R, WARPS_PER_CTA_HEAVY, LOSS, (TileCfg<R>::HEAVY_DEPTH <= G ? TileCfg<R>::HEAVY_DEPTH : G), CS, TileCfg<R>::HEAVY_TRIAL_DEPTH>(A, A.order[blockIdx.x / CS], red, part, xg, clbuf);
}

// out[unit] = r(own[:, unit])  — the penalty terms of calc_penalty (evaluate_fit.jl:91-104).  With one regularizer per
// unit the codes of neighbouring units may differ, and the reductions inside reg_eval are warp-wide shuffles: the warp
// then walks its units one at a time (every lane group evaluates the same unit), so the code stays warp-uniform.
template <int G, int R>
__global__ void __launch_bounds__(128) reg_eval_kernel(const double* __restrict__ own, int64_t units, int stride, int k,
                                                       const int32_t* reg_code, const double* reg_param,
                                                       int reg_uniform, double* out,
                                                       const int64_t* reg_payload_ptr = nullptr, const double* reg_payload = nullptr) {
  const int lane = threadIdx.x & 31, lg = lane % G;
  const int64_t first = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * (32 / G);
  if (reg_uniform) {
    const int64_t unit = first + lane / G;
    const bool ok = unit < units;
    double2 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = 2 * (lg + G * r);
      x[r] = ok ? *reinterpret_cast<const double2*>(own + unit * (int64_t)stride + i0) : make_double2(0.0, 0.0);
    }
    const double* pay = reg_payload_ptr ? reg_payload + reg_payload_ptr[0] : nullptr;
    const int npay = reg_payload_ptr ? (int)(reg_payload_ptr[1] - reg_payload_ptr[0]) : 0;
    const double v = reg_eval<G, R>(reg_code[0], reg_param, x, lg, k, pay, npay);
    if (ok && lg == 0) out[unit] = v;
    return;
  }
  for (int j = 0; j < 32 / G; ++j) {
    const int64_t unit = first + j;
    if (unit >= units) break;
    double2 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = *reinterpret_cast<const double2*>(own + unit * (int64_t)stride + 2 * (lg + G * r));
    const double* pay = reg_payload_ptr ? reg_payload + reg_payload_ptr[unit] : nullptr;
    const int npay = reg_payload_ptr ? (int)(reg_payload_ptr[unit + 1] - reg_payload_ptr[unit]) : 0;
    const double v = reg_eval<G, R>(reg_code[unit], reg_param + unit * GLRMB200_REG_NPARAM, x, lg, k, pay, npay);
    if (lane == 0) out[unit] = v;
  }
}

}  // namespace glrm
