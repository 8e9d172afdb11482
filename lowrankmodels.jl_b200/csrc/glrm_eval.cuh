// glrm_eval.cuh — device-resident evaluation of a fitted model: impute / error_metric.
//
//   impute(glrm)        = impute(losses, X'Y)   src/evaluate_fit.jl:150, src/impute_and_err.jl:147-168 (every entry of the m x n table)
//   error_metric(glrm)  raw and standardised    src/evaluate_fit.jl:106-143 (over observed_examples)
// with the (domain, loss) rules of src/impute_and_err.jl:31-146.  These are callers of the hot path (cross-validation scores a
// fitted fold with them, src/cross_validate.jl:34-44), not part of it: straightforward kernels — one CTA per feature, the
// feature's block of Y in shared memory, a thread per example — that save the round trip of X, Y and an m x n product
// through the host.
#pragma once
#include "glrm_device.cuh"
#include "glrm_vec.cuh"

namespace glrm {

struct EvalArgs {
  const double* X; const double* Y;           // device factors [m][stride], [d][stride]
  int32_t stride, k;
  int64_t m, n;
  const int64_t* ystart;                      // [n+1] or nullptr (every feature has one column)
  const int32_t* loss_code; const double* loss_param;
  const int32_t* dom_code; const double* dom_param;   // [n], [2n]: GLRMB200_DOMAIN_* and (min, max) / (T, -) / (max_count, -)
  // observations by feature: lists (ptr/idx/val of this handle's columns) or the dense array
  const int64_t* col_ptr; const int32_t* col_idx; const double* col_val;
  const double* dense_A; int64_t lda;
  double* out_imputed;                        // [n][m] column-major (impute)
  double* col_err; double* col_sq; double* col_cnt;   // [n] per-feature sums (error_metric)
};

__device__ __forceinline__ double ev_roundcutoff(double x, double a, double b) { return fmin(fmax(rint(x), a), b); }   // impute_and_err.jl:31
__device__ __forceinline__ double ev_pos_mod(double T, double x) { return x > 0.0 ? fmod(x, T) : fmod(x, T) + T; }       // :112
__device__ __forceinline__ double ev_loss(int code, const double* lp, double u, double a) {
  double l, c;
  loss_eval<0, false>(code, lp[0], lp[1], lp[2], u, a, l, c);
  return l;
}

// a_u = argmin_a loss(u, a) over the domain (impute_and_err.jl:40-120); NaN where the reference has no method / raises
__device__ __forceinline__ double ev_impute(int dom, double d0, double d1, int code, const double* lp, double (&u)[VEC_DMAX], int D) {
  if (dom == GLRMB200_DOMAIN_COUNT) { dom = GLRMB200_DOMAIN_ORDINAL; d1 = d0; d0 = 0.0; }        // :117
  if (code < GLRMB200_LOSS_MULTINOMIAL) {
    const double x = u[0];
    const bool diff = code <= GLRMB200_LOSS_PERIODIC;                                             // DiffLoss (losses.jl:55)
    switch (dom) {
      case GLRMB200_DOMAIN_REAL:
      case GLRMB200_DOMAIN_PERIODIC:                                                              // :109
        if (diff) return x;                                                                       // :40
        if (code == GLRMB200_LOSS_POISSON) return exp(x);                                         // :41
        if (code == GLRMB200_LOSS_ORDINAL_HINGE) return ev_roundcutoff(x, lp[1], lp[2]);          // :42
        if (code == GLRMB200_LOSS_WEIGHTED_HINGE) return 1.0 / x;                                 // :44-47
        return NAN;                                                                               // :43 (LogisticLoss: error)
      case GLRMB200_DOMAIN_BOOL:
        if (code == GLRMB200_LOSS_LOGISTIC || code == GLRMB200_LOSS_WEIGHTED_HINGE) return x >= 0.0 ? 1.0 : 0.0;   // :60
        return ev_loss(code, lp, x, 0.0) < ev_loss(code, lp, x, 1.0) ? 0.0 : 1.0;                 // :63
      case GLRMB200_DOMAIN_ORDINAL:
        if (diff) return ev_roundcutoff(x, d0, d1);                                               // :75
        if (code == GLRMB200_LOSS_POISSON) return ev_roundcutoff(exp(x), d0, d1);                 // :76
        if (code == GLRMB200_LOSS_ORDINAL_HINGE) return ev_roundcutoff(x, d0, d1);                // :77
        if (code == GLRMB200_LOSS_LOGISTIC) return x > 0.0 ? d1 : d0;                             // :78
        return ev_roundcutoff(x > 0.0 ? ceil(1.0 / x) : floor(1.0 / x), d0, d1);                  // :79-83
      default: return NAN;
    }
  }
  if (dom == GLRMB200_DOMAIN_CATEGORICAL && (code == GLRMB200_LOSS_MULTINOMIAL || code == GLRMB200_LOSS_OVA)) {   // :101-102 argmax(u)
    int best = 0;
    for (int j = 1; j < D; ++j) if (am_better(u[j], j, u[best], best)) best = j;
    return (double)(best + 1);
  }
  if (dom == GLRMB200_DOMAIN_ORDINAL) {
    if (code == GLRMB200_LOSS_ORDISTIC) {                                                         // :84 argmin(u.^2)
      int best = 0;
      for (int j = 1; j < D; ++j) if (am_better(-(u[j] * u[j]), j, -(u[best] * u[best]), best)) best = j;
      return (double)(best + 1);
    }
    if (code == GLRMB200_LOSS_MULTINOMIAL_ORDINAL) {                                              // :85-90
      u[0] = jl_mind(-1e-3, u[0]);
      for (int j = 1; j < D; ++j) u[j] = jl_mind(u[j], u[j - 1] - 1e-3);
      int best = 0;
      double pbest = 1.0 - exp(u[0]);
      for (int j = 1; j <= D; ++j) {
        const double p = j < D ? exp(u[j - 1]) - exp(u[j]) : exp(u[D - 1]);
        if (am_better(p, j, pbest, best)) { pbest = p; best = j; }
      }
      return (double)(best + 1);
    }
    // :91-93  (D.min:D.max)[argmin([evaluate(l, u, i) for i in D.min:D.max])]
    double lbest = INFINITY, abest = d0;
    bool first = true;
    for (double a = d0; a <= d1; a += 1.0) {
      double uu[VEC_DMAX], gc[VEC_DMAX];
      for (int j = 0; j < VEC_DMAX; ++j) { uu[j] = u[j]; gc[j] = 0.0; }
      const double l = vec_loss<false>(code, lp, uu, D, a, gc);
      if (first || am_better(-l, (int)(a - d0), -lbest, (int)(abest - d0))) { lbest = l; abest = a; first = false; }
    }
    return abest;
  }
  return NAN;
}
__device__ __forceinline__ double ev_error(int dom, double d0, double a_imp, double a) {
  switch (dom) {
    case GLRMB200_DOMAIN_BOOL: case GLRMB200_DOMAIN_CATEGORICAL: return a_imp == a ? 0.0 : 1.0;   // :35,64-67,103-106
    case GLRMB200_DOMAIN_PERIODIC: { const double t = ev_pos_mod(d0, a_imp) - ev_pos_mod(d0, a); return t * t; }   // :113-116
    default: { const double t = a_imp - a; return t * t; }                                        // :34,49-52,94-97,118-121
  }
}

// u = x_e' Y_f (one dot product per column of the feature's block)
__device__ __forceinline__ void ev_dots(const EvalArgs& P, const double* __restrict__ ys, int D, int64_t e, double (&u)[VEC_DMAX]) {
  const double* x = P.X + e * P.stride;
#pragma unroll
  for (int c = 0; c < VEC_DMAX; ++c) u[c] = 0.0;
  for (int i = 0; i < P.k; ++i) {
    const double xi = x[i];
    for (int c = 0; c < D; ++c) u[c] = fma(xi, ys[c * P.k + i], u[c]);
  }
}

// MODE 0: A_imputed[:, f] for every example;  MODE 1: per-feature sums of error_metric, a^2 and the count over the observed examples
template <int MODE>
__global__ void __launch_bounds__(256) eval_kernel(const EvalArgs P) {
  extern __shared__ double ev_ys[];            // [D][k] the feature's block of Y
  __shared__ double red[3][256];
  const int64_t f = blockIdx.x;
  const int64_t y0 = P.ystart ? P.ystart[f] : f;
  const int D = P.ystart ? (int)(P.ystart[f + 1] - y0) : 1;
  for (int x = threadIdx.x; x < D * P.k; x += blockDim.x) ev_ys[x] = P.Y[(y0 + x / P.k) * P.stride + x % P.k];
  __syncthreads();
  const int code = P.loss_code[f];
  const double* lp = P.loss_param + f * GLRMB200_LOSS_NPARAM;
  const int dom = P.dom_code[f];
  const double d0 = P.dom_param[2 * f], d1 = P.dom_param[2 * f + 1];
  if (MODE == 0) {
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < P.m; e += (int64_t)gridDim.y * blockDim.x) {
      double u[VEC_DMAX];
      ev_dots(P, ev_ys, D, e, u);
      P.out_imputed[f * P.m + e] = ev_impute(dom, d0, d1, code, lp, u, D);
    }
    return;
  }
  const int64_t q0 = P.dense_A ? 0 : P.col_ptr[f], q1 = P.dense_A ? P.m : P.col_ptr[f + 1];
  double serr = 0.0, ssq = 0.0;
  for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {                 // (fixed assignment: deterministic sums)
    const int64_t e = P.dense_A ? q : P.col_idx[q];
    const double a = P.dense_A ? P.dense_A[f * P.lda + q] : P.col_val[q];
    double u[VEC_DMAX];
    ev_dots(P, ev_ys, D, e, u);
    serr += ev_error(dom, d0, ev_impute(dom, d0, d1, code, lp, u, D), a);
    ssq += a * a;
  }
  red[0][threadIdx.x] = serr; red[1][threadIdx.x] = ssq;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { P.col_err[f] = red[0][0]; P.col_sq[f] = red[1][0]; P.col_cnt[f] = (double)(q1 - q0); }
}

// err = sum_f col_err[f]   or, standardised (evaluate_fit.jl:119-137), sum_f col_err[f] / mean(a^2) where that mean is not 0
__global__ void eval_total_kernel(const double* col_err, const double* col_sq, const double* col_cnt, int64_t n, int standardize, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double err = 0.0;
  for (int64_t f = 0; f < n; ++f) {
    double ce = col_err[f];
    if (standardize) {
      const double mean = col_sq[f] / col_cnt[f];
      if (mean != 0.0) ce = ce / mean;
    }
    err += ce;
  }
  out[0] = err;
}

}  // namespace glrm
