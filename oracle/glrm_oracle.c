/*
 * glrm_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (C11 + OpenMP, Float64) of the one hot path this repo accelerates:
 *     fit!(glrm::GLRM, params::ProxGradParams)   /root/reference/src/algorithms/proxgrad.jl:34-220
 *     (threaded twin: src/algorithms/proxgrad_multithread.jl — identical maths, Threads.@threads
 *      over rows :118 and columns :163, which `#pragma omp parallel for schedule(static)` mirrors)
 * together with everything that path calls: row_objective / col_objective / objective /
 * calc_penalty (src/evaluate_fit.jl:4-55,91-104), grad/evaluate of every loss
 * (src/losses.jl:136-676) and evaluate/prox of the regularizers (src/regularizers.jl:24-348).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (lowrankmodels.jl_b200/csrc) never links or calls
 * it; there is no CPU fallback in the product.
 *
 * Two forms of the same algorithm (selected by `mode`):
 *   mode 0  "faithful":  materialises XY = X'Y (proxgrad.jl:65-66,157,202) and evaluates the line
 *           search through dense x'*Y / X'*y products exactly like row_objective / col_objective
 *           (evaluate_fit.jl:29,45).  Reproduces the reference's cost profile: this is the CPU
 *           baseline "reference algorithm as written" (BASELINE.md B1).
 *   mode 1  "sparse-evaluated": same arithmetic, but X'Y is only evaluated at observed entries.
 *           Parity oracle at scale and the "best-effort CPU" baseline.
 *
 * PARITY PINNING: the reference is Julia and cannot run in this image (no julia binary, no
 * network).  Scalar loss / regularizer arithmetic is pinned against the known answers held by the
 * reference's own tests (tests/golden/reference_known_answers.json cites each file:line).  The
 * end-to-end objective trajectory is NOT pinned by any reference fixture (the reference's tests
 * assert no trajectory and draw data from Julia's RNG): trajectory parity is "unpinned" and rests
 * on this restatement plus the independent line-by-line Python restatement oracle/proxgrad_ref.py.
 *
 * The problem encoding is the C ABI's (include/glrm_b200.h) so tests feed identical bytes to the
 * oracle and to the engine.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/glrm_b200.h"

#define ORACLE_MAX_D 64 /* largest embedding dimension handled (levels of a categorical) */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Julia semantics helpers ------------------------------------------------------------------- */
static inline double jl_max(double a, double b) { /* Base.max propagates NaN */
  if (isnan(a) || isnan(b)) return NAN;
  return a > b ? a : b;
}
static inline double jl_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a < b ? a : b;
}
static inline double jl_sign(double x) { return (x > 0) - (x < 0); }

/* myBool (losses.jl:104): 1 -> true; -1, 0 -> false; anything else is an InexactError.
 * Labels cross the ABI as Float64; Bool true/false arrive as 1.0/0.0. */
static inline int label_bool(double a, int* err) {
  if (a == 1.0) return 1;
  if (a == 0.0 || a == -1.0) return 0;
  *err = 1;
  return 0;
}

/* ---- embedding dims (losses.jl:72-73,366,421,458,496,569) ---------------------------------- */
static int embedding_dim(int code, const double* p) {
  switch (code) {
    case GLRMB200_LOSS_MULTINOMIAL:
    case GLRMB200_LOSS_OVA:
    case GLRMB200_LOSS_ORDISTIC: return (int)p[2];
    case GLRMB200_LOSS_BVS:
    case GLRMB200_LOSS_MULTINOMIAL_ORDINAL: return (int)p[2] - 1;
    default: return 1;
  }
}

/* ---- scalar losses: evaluate / grad on u::Real ------------------------------------------------ */
static double sl_eval(int code, const double* p, double u, double a, int* err) {
  const double s = p[0];
  switch (code) {
    case GLRMB200_LOSS_QUAD: return s * (u - a) * (u - a);                    /* losses.jl:144 */
    case GLRMB200_LOSS_L1: return s * fabs(u - a);                             /* :158 */
    case GLRMB200_LOSS_HUBER: {                                                /* :173-175 */
      const double c = p[1];
      return fabs(u - a) > c ? (fabs(u - a) - c + c * c) * s : (u - a) * (u - a) * s;
    }
    case GLRMB200_LOSS_QUANTILE: {                                             /* :193-196 */
      const double diff = a - u, q = p[1];
      return diff > 0 ? s * q * diff : -s * (1 - q) * diff;
    }
    case GLRMB200_LOSS_PERIODIC:                                               /* :216 */
      return s * (1 - cos((a - u) * (2 * M_PI) / p[1]));
    case GLRMB200_LOSS_POISSON:                                                /* :237-239 */
      return s * (exp(u) - a * u + (a == 0 ? 0 : a * (log(a) - 1)));
    case GLRMB200_LOSS_ORDINAL_HINGE: {                                        /* :258-278 */
      const double lmin = p[1], lmax = p[2];
      double n, loss;
      if (u > lmax - 1) {
        n = jl_min(floor(u), lmax - 1) - a;
        loss = n * (n + 1) / 2 + (n + 1) * (u - lmax + 1);
      } else if (u > a) {
        n = jl_min(floor(u), lmax) - a;
        loss = n * (n + 1) / 2 + (n + 1) * (u - floor(u));
      } else if (u > lmin + 1) {
        n = a - jl_max(ceil(u), lmin + 1);
        loss = n * (n + 1) / 2 + (n + 1) * (ceil(u) - u);
      } else {
        n = a - jl_max(ceil(u), lmin + 1);
        loss = n * (n + 1) / 2 + (n + 1) * (lmin + 1 - u);
      }
      return s * loss;
    }
    case GLRMB200_LOSS_LOGISTIC: {                                             /* :304 */
      const int b = label_bool(a, err);
      return s * log(1 + exp(-(2.0 * b - 1) * u));
    }
    case GLRMB200_LOSS_WEIGHTED_HINGE: {                                       /* :326-332 */
      const int b = label_bool(a, err);
      double loss = s * jl_max(1 - (2.0 * b - 1) * u, 0);
      if (p[1] != 1.0 && b) loss *= p[1];
      return loss;
    }
    default: *err = 2; return NAN;
  }
}

static double sl_grad(int code, const double* p, double u, double a, int* err) {
  const double s = p[0];
  switch (code) {
    case GLRMB200_LOSS_QUAD: return 2 * (u - a) * s;                           /* losses.jl:146 */
    case GLRMB200_LOSS_L1: return jl_sign(u - a) * s;                          /* :160 */
    case GLRMB200_LOSS_HUBER:                                                  /* :177 (sic: no 2) */
      return fabs(u - a) > p[1] ? jl_sign(u - a) * s : (u - a) * s;
    case GLRMB200_LOSS_QUANTILE: {                                             /* :198-201 */
      const double diff = a - u, q = p[1];
      return diff > 0 ? -s * q : s * (1 - q);
    }
    case GLRMB200_LOSS_PERIODIC:                                               /* :218 */
      return -s * ((2 * M_PI) / p[1]) * sin((a - u) * (2 * M_PI) / p[1]);
    case GLRMB200_LOSS_POISSON: return s * (exp(u) - a);                       /* :241 */
    case GLRMB200_LOSS_ORDINAL_HINGE: {                                        /* :280-292 */
      const double lmin = p[1], lmax = p[2];
      double g;
      if (u > a) g = jl_min(ceil(u), lmax) - a;
      else g = -(a - jl_max(floor(u), lmin));
      return s * g;
    }
    case GLRMB200_LOSS_LOGISTIC: {                                             /* :306 */
      const int b = label_bool(a, err);
      const double aa = 2.0 * b - 1;
      return -aa * s / (1 + exp(aa * u));
    }
    case GLRMB200_LOSS_WEIGHTED_HINGE: {                                       /* :334-341 */
      const int b = label_bool(a, err);
      const double an = 2.0 * b - 1;
      double g = (an * u >= 1) ? 0 : -an * s;
      if (p[1] != 1.0 && b) g *= p[1];
      return g;
    }
    default: *err = 2; return NAN;
  }
}

/* ---- vector losses: u is a d_f-vector, a an Integer level (1-based) ------------------------- */
static void enforce_mnl_ord_rules(double* u, int D) {                          /* :572-578 */
  const double TOL = 1e-3;
  u[0] = jl_min(-TOL, u[0]);
  for (int j = 1; j < D; ++j) u[j] = jl_min(u[j], u[j - 1] - TOL);
}

static double vl_eval(int code, const double* p, const double* uin, int D, double alab, int* err) {
  const double s = p[0];
  const int a = (int)alab; /* 1-based level */
  double u[ORACLE_MAX_D];
  if (D > ORACLE_MAX_D) { *err = 3; return NAN; }
  if ((double)a != alab || a < 1) { *err = 1; return NAN; }
  memcpy(u, uin, sizeof(double) * (size_t)D);
  switch (code) {
    case GLRMB200_LOSS_MULTINOMIAL: {                                          /* :369-380 */
      if (a > D) { *err = 1; return NAN; }
      double mx = u[0];
      for (int j = 1; j < D; ++j) mx = jl_max(mx, u[j]);
      const double M = mx - u[a - 1];
      double sumexp = 0;
      for (int j = 0; j < D; ++j) sumexp += exp(u[j] - u[a - 1] - M);
      return s * (log(sumexp) + M);
    }
    case GLRMB200_LOSS_OVA: {                                                  /* :424-430 */
      const int bcode = (int)p[3];
      const double bp[GLRMB200_LOSS_NPARAM] = {p[4], p[5], p[6], 0, 0, 0, 0, 0};
      double loss = 0;
      for (int j = 0; j < D; ++j) loss += sl_eval(bcode, bp, u[j], (a == j + 1) ? 1.0 : 0.0, err);
      return s * loss;
    }
    case GLRMB200_LOSS_BVS: {                                                  /* :461-467 */
      const int bcode = (int)p[3];
      const double bp[GLRMB200_LOSS_NPARAM] = {p[4], p[5], p[6], 0, 0, 0, 0, 0};
      double loss = 0;
      for (int j = 0; j < D; ++j) loss += sl_eval(bcode, bp, u[j], (a > j + 1) ? 1.0 : 0.0, err);
      return s * loss;
    }
    case GLRMB200_LOSS_ORDISTIC: {                                             /* :499-505 */
      if (a > D) { *err = 1; return NAN; }
      double diff[ORACLE_MAX_D], M = -INFINITY, invlik = 0;
      for (int j = 0; j < D; ++j) { diff[j] = u[a - 1] * u[a - 1] - u[j] * u[j]; M = jl_max(M, diff[j]); }
      for (int j = 0; j < D; ++j) invlik += exp(diff[j] - M);
      return s * (M + log(invlik));
    }
    case GLRMB200_LOSS_MULTINOMIAL_ORDINAL: {                                  /* :581-590 */
      const int lmax = (int)p[2];
      if (a > lmax) { *err = 1; return NAN; }
      enforce_mnl_ord_rules(u, D);
      if (a == 1) return -s * log(exp(0.0) - exp(u[0]));
      if (a == lmax) return -s * u[a - 2];
      return -s * log(exp(u[a - 2]) - exp(u[a - 1]));
    }
    default: *err = 2; return NAN;
  }
}

static void vl_grad(int code, const double* p, const double* uin, int D, double alab, double* g, int* err) {
  const double s = p[0];
  const int a = (int)alab;
  double u[ORACLE_MAX_D];
  if (D > ORACLE_MAX_D) { *err = 3; return; }
  if ((double)a != alab || a < 1) { *err = 1; return; }
  memcpy(u, uin, sizeof(double) * (size_t)D);
  for (int j = 0; j < D; ++j) g[j] = 0;
  switch (code) {
    case GLRMB200_LOSS_MULTINOMIAL: {                                          /* :382-398 */
      if (a > D) { *err = 1; return; }
      g[a - 1] = -1;
      double mx = u[0];
      for (int j = 1; j < D; ++j) mx = jl_max(mx, u[j]);
      for (int j = 0; j < D; ++j) {
        const double M = mx - u[j];
        double sumexp = 0;
        for (int jp = 0; jp < D; ++jp) sumexp += exp(u[jp] - u[j] - M);
        g[j] += exp(-M) / sumexp;
      }
      for (int j = 0; j < D; ++j) g[j] *= s;
      return;
    }
    case GLRMB200_LOSS_OVA: {                                                  /* :432-438 */
      const int bcode = (int)p[3];
      const double bp[GLRMB200_LOSS_NPARAM] = {p[4], p[5], p[6], 0, 0, 0, 0, 0};
      for (int j = 0; j < D; ++j) g[j] = s * sl_grad(bcode, bp, u[j], (a == j + 1) ? 1.0 : 0.0, err);
      return;
    }
    case GLRMB200_LOSS_BVS: {                                                  /* :469-475 */
      const int bcode = (int)p[3];
      const double bp[GLRMB200_LOSS_NPARAM] = {p[4], p[5], p[6], 0, 0, 0, 0, 0};
      for (int j = 0; j < D; ++j) g[j] = s * sl_grad(bcode, bp, u[j], (a > j + 1) ? 1.0 : 0.0, err);
      return;
    }
    case GLRMB200_LOSS_ORDISTIC: {                                             /* :507-519 */
      if (a > D) { *err = 1; return; }
      g[a - 1] = 2 * u[a - 1];
      for (int j = 0; j < D; ++j) {
        double M = -INFINITY, invlik = 0;
        for (int jp = 0; jp < D; ++jp) M = jl_max(M, u[j] * u[j] - u[jp] * u[jp]);
        for (int jp = 0; jp < D; ++jp) invlik += exp(u[j] * u[j] - u[jp] * u[jp] - M);
        g[j] -= 2 * u[j] * exp(-M) / invlik;
      }
      for (int j = 0; j < D; ++j) g[j] *= s;
      return;
    }
    case GLRMB200_LOSS_MULTINOMIAL_ORDINAL: {                                  /* :592-608 */
      const int lmax = (int)p[2];
      if (a > lmax) { *err = 1; return; }
      enforce_mnl_ord_rules(u, D);
      if (a == 1) {
        g[0] = -exp(u[0]) / (exp(0.0) - exp(u[0]));
      } else if (a == lmax) {
        g[a - 2] = 1;
      } else {
        g[a - 1] = -exp(u[a - 1]) / (exp(u[a - 2]) - exp(u[a - 1]));
        g[a - 2] = exp(u[a - 2]) / (exp(u[a - 2]) - exp(u[a - 1]));
      }
      for (int j = 0; j < D; ++j) g[j] = -s * g[j];
      return;
    }
    default: *err = 2; return;
  }
}

/* unified: D==1 scalar path, otherwise vector path */
static inline double loss_eval(int code, const double* p, const double* u, int D, double a, int* err) {
  return embedding_dim(code, p) == 1 && code < GLRMB200_LOSS_MULTINOMIAL ? sl_eval(code, p, u[0], a, err)
                                                                         : vl_eval(code, p, u, D, a, err);
}
static inline void loss_grad(int code, const double* p, const double* u, int D, double a, double* g, int* err) {
  if (code < GLRMB200_LOSS_MULTINOMIAL) g[0] = sl_grad(code, p, u[0], a, err);
  else vl_grad(code, p, u, D, a, g, err);
}

/* ---- regularizers on a k x D block stored column-major (D=1: a factor column) -------------- */
static const double REG_TOL = 1e-12; /* regularizers.jl:25 */

static double base_reg_eval(int base, const double* p, const double* v, int64_t L) {
  switch (base) {
    case GLRMB200_REG_ZERO: return 0;                                          /* :95 */
    case GLRMB200_REG_QUAD: {                                                  /* :58 */
      double s = 0;
      for (int64_t i = 0; i < L; ++i) s += v[i] * v[i];
      return p[0] * s;
    }
    case GLRMB200_REG_QUAD_CONSTRAINT: {                                       /* :74 */
      double s = 0;
      for (int64_t i = 0; i < L; ++i) s += v[i] * v[i];
      return sqrt(s) > p[0] + REG_TOL ? INFINITY : 0;
    }
    case GLRMB200_REG_ONE: {                                                   /* :88 */
      double s = 0;
      for (int64_t i = 0; i < L; ++i) s += fabs(v[i]);
      return p[0] * s;
    }
    case GLRMB200_REG_NONNEG:                                                  /* :105-112 */
      for (int64_t i = 0; i < L; ++i) if (v[i] < 0) return INFINITY;
      return 0;
    case GLRMB200_REG_NONNEG_ONE: {                                            /* :129-136 */
      double s = 0;
      for (int64_t i = 0; i < L; ++i) if (v[i] < 0) return INFINITY;
      for (int64_t i = 0; i < L; ++i) s += v[i];
      return p[0] * s;
    }
    case GLRMB200_REG_ONE_SPARSE: {                                            /* :239-253 */
      int oneflag = 0;
      for (int64_t i = 0; i < L; ++i) {
        if (oneflag) { if (v[i] != 0) return INFINITY; }
        else if (v[i] != 0) oneflag = 1;
      }
      return 0;
    }
    case GLRMB200_REG_KSPARSE: {                                               /* :261-276 */
      const int64_t kk = (int64_t)p[0];
      int64_t nonz = 0;
      for (int64_t i = 0; i < L; ++i) {
        if (nonz == kk) { if (v[i] != 0) return INFINITY; }
        else if (v[i] != 0) nonz++;
      }
      return 0;
    }
    case GLRMB200_REG_UNIT_ONE_SPARSE: {                                       /* :300-316 */
      int oneflag = 0;
      for (int64_t i = 0; i < L; ++i) {
        if (v[i] == 0) continue;
        else if (v[i] == 1) { if (oneflag) return INFINITY; else oneflag = 1; }
        else return INFINITY;
      }
      return 0;
    }
    case GLRMB200_REG_SIMPLEX: {                                               /* :338-346 */
      double s = 0;
      for (int64_t i = 0; i < L; ++i) s += v[i];
      if (fabs(s - 1) > REG_TOL) return INFINITY;
      for (int64_t i = 0; i < L; ++i) if (v[i] < 0) return INFINITY;
      return 0;
    }
    default: return NAN;
  }
}

static int64_t jl_argmax(const double* v, int64_t L) { /* first maximal element; NaN wins */
  int64_t idx = 0;
  for (int64_t i = 1; i < L; ++i) {
    if (isnan(v[idx])) break;
    if (isnan(v[i]) || v[i] > v[idx]) idx = i;
  }
  return idx;
}

static int cmp_desc(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x < y) - (x > y);
}

static void base_reg_prox(int base, const double* p, double* v, int64_t L, double alpha) {
  switch (base) {
    case GLRMB200_REG_ZERO: return;                                            /* :93 */
    case GLRMB200_REG_QUAD: {                                                  /* :56 */
      const double c = 1 / (1 + 2 * alpha * p[0]);
      for (int64_t i = 0; i < L; ++i) v[i] = c * v[i];
      return;
    }
    case GLRMB200_REG_QUAD_CONSTRAINT: {                                       /* :72 (always rescales) */
      double s = 0;
      for (int64_t i = 0; i < L; ++i) s += v[i] * v[i];
      const double c = p[0] / sqrt(s);
      for (int64_t i = 0; i < L; ++i) v[i] = c * v[i];
      return;
    }
    case GLRMB200_REG_ONE: {                                                   /* :83-87 */
      const double t = p[0] * alpha;
      for (int64_t i = 0; i < L; ++i) v[i] = jl_max(v[i] - t, 0) + jl_min(v[i] + t, 0);
      return;
    }
    case GLRMB200_REG_NONNEG:                                                  /* :103 */
      for (int64_t i = 0; i < L; ++i) v[i] = jl_max(v[i], 0);
      return;
    case GLRMB200_REG_NONNEG_ONE:                                              /* :122 (ignores scale) */
      for (int64_t i = 0; i < L; ++i) v[i] = jl_max(v[i] - alpha, 0);
      return;
    case GLRMB200_REG_ONE_SPARSE: {                                            /* :237 */
      const int64_t idx = jl_argmax(v, L);
      const double keep = v[idx];
      for (int64_t i = 0; i < L; ++i) v[i] = 0;
      v[idx] = keep;
      return;
    }
    case GLRMB200_REG_KSPARSE: {                                               /* :277-283 */
      /* keep the k entries of largest |v| (ties: lowest index first) */
      const int64_t kk = (int64_t)p[0];
      char* keep = (char*)calloc((size_t)L, 1);
      for (int64_t r = 0; r < kk && r < L; ++r) {
        int64_t best = -1;
        for (int64_t i = 0; i < L; ++i) {
          if (keep[i]) continue;
          if (best < 0 || fabs(v[i]) > fabs(v[best])) best = i;
        }
        keep[best] = 1;
      }
      for (int64_t i = 0; i < L; ++i) if (!keep[i]) v[i] = 0;
      free(keep);
      return;
    }
    case GLRMB200_REG_UNIT_ONE_SPARSE: {                                       /* :297 */
      const int64_t idx = jl_argmax(v, L);
      for (int64_t i = 0; i < L; ++i) v[i] = 0;
      v[idx] = 1;
      return;
    }
    case GLRMB200_REG_SIMPLEX: {                                               /* :325-337 */
      double* y = (double*)malloc(sizeof(double) * (size_t)L);
      memcpy(y, v, sizeof(double) * (size_t)L);
      qsort(y, (size_t)L, sizeof(double), cmp_desc);
      double* ysum = (double*)malloc(sizeof(double) * (size_t)L);
      double acc = 0;
      for (int64_t i = 0; i < L; ++i) { acc += y[i]; ysum[i] = acc; }
      double t = (ysum[L - 1] - 1) / (double)L;
      for (int64_t i = 0; i < L - 1; ++i) {
        if ((ysum[i] - 1) / (double)(i + 1) >= y[i + 1]) { t = (ysum[i] - 1) / (double)(i + 1); break; }
      }
      for (int64_t i = 0; i < L; ++i) v[i] = jl_max(v[i] - t, 0);
      free(y); free(ysum);
      return;
    }
    default: return;
  }
}

/* wrappers lastentry1 / lastentry_unpenalized (regularizers.jl:163-189): the inner regularizer
 * sees rows 1..k-1 of every column of the block.  pay / npay: the regularizer's vector payload (fixed_latent_features.y,
 * fixed_last_latent_features.y, RemQuadReg.m), NULL / 0 otherwise. */
static double reg_eval_p(int code, const double* p, const double* v, int64_t k, int64_t D, const double* pay, int64_t npay) {
  const int base = code & GLRMB200_REG_BASE_MASK;
  if (code & GLRMB200_REG_FIXED_FIRST) {                                        /* :208: a[1:n]==y ? evaluate(r.r, a[n+1:end]) : Inf */
    for (int64_t i = 0; i < npay; ++i) if (v[i] != pay[i]) return INFINITY;
    return base_reg_eval(base, p, v + npay, k - npay);
  }
  if (code & GLRMB200_REG_FIXED_LAST) {                                         /* :230: a[k-n+1:end]==y ? evaluate(r.r, a[1:k-n]) : Inf */
    for (int64_t i = 0; i < npay; ++i) if (v[k - npay + i] != pay[i]) return INFINITY;
    return base_reg_eval(base, p, v, k - npay);
  }
  if (base == GLRMB200_REG_REM_QUAD) {                                          /* :423: scale * sum(abs2, a - m) */
    double s = 0;
    for (int64_t i = 0; i < k * D; ++i) { const double t = v[i] - pay[i]; s += t * t; }
    return p[0] * s;
  }
  if (code & (GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL))                 /* regularizers.jl:378,405 */
    return base_reg_eval(base, p, v, k - 1);                                    /* evaluate(r.r, a[1:end-1,1]) */
  if (code & (GLRMB200_REG_LASTENTRY1 | GLRMB200_REG_LASTENTRY_UNPENALIZED)) {
    if (code & GLRMB200_REG_LASTENTRY1)
      for (int64_t c = 0; c < D; ++c) if (v[c * k + k - 1] != 1) return INFINITY;      /* :171-172 */
    double* tmp = (double*)malloc(sizeof(double) * (size_t)((k - 1) * D + 1));
    for (int64_t c = 0; c < D; ++c) memcpy(tmp + c * (k - 1), v + c * k, sizeof(double) * (size_t)(k - 1));
    const double r = base_reg_eval(base, p, tmp, (k - 1) * D);
    free(tmp);
    return r;
  }
  return base_reg_eval(base, p, v, k * D);
}

static void reg_prox_p(int code, const double* p, double* v, int64_t k, int64_t D, double alpha, const double* pay, int64_t npay) {
  const int base = code & GLRMB200_REG_BASE_MASK;
  if (code & GLRMB200_REG_FIXED_FIRST) {                                        /* :203: [r.y; prox(r.r, u[n+1:end], alpha)] */
    base_reg_prox(base, p, v + npay, k - npay, alpha);
    for (int64_t i = 0; i < npay; ++i) v[i] = pay[i];
    return;
  }
  if (code & GLRMB200_REG_FIXED_LAST) {                                         /* :223: [prox(r.r, u[n+1:end], alpha); r.y]  (sic) */
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(k - npay + 1));
    memcpy(tmp, v + npay, sizeof(double) * (size_t)(k - npay));
    base_reg_prox(base, p, tmp, k - npay, alpha);
    memcpy(v, tmp, sizeof(double) * (size_t)(k - npay));
    for (int64_t i = 0; i < npay; ++i) v[k - npay + i] = pay[i];
    free(tmp);
    return;
  }
  if (base == GLRMB200_REG_REM_QUAD) {                                          /* :417-418: (u + 2 alpha scale m) / (1 + 2 alpha scale) */
    for (int64_t i = 0; i < k * D; ++i) v[i] = (v[i] + 2 * alpha * p[0] * pay[i]) / (1 + 2 * alpha * p[0]);
    return;
  }
  if (code & (GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL)) {               /* regularizers.jl:361-377,390-404 */
    double* um = (double*)malloc(sizeof(double) * (size_t)k);
    for (int64_t i = 0; i < k - 1; ++i) {                                       /* um = mean(u[1:end-1,:], dims=2) */
      double acc = 0.0;
      for (int64_t c = 0; c < D; ++c) acc += v[c * k + i];
      um[i] = acc / (double)D;
    }
    base_reg_prox(base, p, um, k - 1, alpha);                                   /* prox!(r.r, um, alpha) */
    for (int64_t i = 0; i < k - 1; ++i)
      for (int64_t c = 0; c < D; ++c) v[c * k + i] = um[i];
    free(um);
    if (code & GLRMB200_REG_MNL_ORDINAL) {                                      /* :399-402 */
      const double TOL = 1e-3;
      v[k - 1] = jl_min(-TOL, v[k - 1]);
      for (int64_t c = 1; c < D; ++c) v[c * k + k - 1] = jl_min(v[c * k + k - 1], v[(c - 1) * k + k - 1] - TOL);
    }
    return;
  }
  if (code & (GLRMB200_REG_LASTENTRY1 | GLRMB200_REG_LASTENTRY_UNPENALIZED)) {
    double* tmp = (double*)malloc(sizeof(double) * (size_t)((k - 1) * D + 1));
    for (int64_t c = 0; c < D; ++c) memcpy(tmp + c * (k - 1), v + c * k, sizeof(double) * (size_t)(k - 1));
    base_reg_prox(base, p, tmp, (k - 1) * D, alpha);
    for (int64_t c = 0; c < D; ++c) memcpy(v + c * k, tmp + c * (k - 1), sizeof(double) * (size_t)(k - 1));
    free(tmp);
    if (code & GLRMB200_REG_LASTENTRY1) for (int64_t c = 0; c < D; ++c) v[c * k + k - 1] = 1; /* :168,170 */
    return;
  }
  base_reg_prox(base, p, v, k * D, alpha);
}

/* ---- exported scalar entry points (known-answer tests) -------------------------------------- */
int oracle_embedding_dim(int code, const double* p) { return embedding_dim(code, p); }
double oracle_loss_eval(int code, const double* p, const double* u, int D, double a, int* err) {
  int e = 0; const double r = loss_eval(code, p, u, D, a, &e); if (err) *err = e; return r;
}
void oracle_loss_grad(int code, const double* p, const double* u, int D, double a, double* g, int* err) {
  int e = 0; loss_grad(code, p, u, D, a, g, &e); if (err) *err = e;
}
double oracle_reg_eval_payload(int code, const double* p, const double* v, int64_t k, int64_t D, const double* pay, int64_t npay) {
  return reg_eval_p(code, p, v, k, D, pay, npay);
}
void oracle_reg_prox_payload(int code, const double* p, double* v, int64_t k, int64_t D, double alpha, const double* pay, int64_t npay) {
  reg_prox_p(code, p, v, k, D, alpha, pay, npay);
}
double oracle_reg_eval(int code, const double* p, const double* v, int64_t k, int64_t D) {
  return reg_eval_p(code, p, v, k, D, NULL, 0);
}
void oracle_reg_prox(int code, const double* p, double* v, int64_t k, int64_t D, double alpha) {
  reg_prox_p(code, p, v, k, D, alpha, NULL, 0);
}

/* ---- problem view --------------------------------------------------------------------------- */
typedef struct {
  const glrmb200_problem* P;
  int64_t m, n, k, d;
  int64_t* ystart; /* [n+1] get_yidxs (losses.jl:76-93), 0-based */
  int* dim;        /* [n]   */
  int64_t nnz_rows, nnz_cols;
} view_t;

static int view_init(view_t* V, const glrmb200_problem* P) {
  V->P = P; V->m = P->m; V->n = P->n; V->k = P->k; V->d = P->d;
  V->ystart = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P->n + 1));
  V->dim = (int*)malloc(sizeof(int) * (size_t)P->n);
  int64_t acc = 0;
  for (int64_t f = 0; f < P->n; ++f) {
    V->ystart[f] = acc;
    V->dim[f] = embedding_dim(P->loss_code[f], P->loss_param + f * GLRMB200_LOSS_NPARAM);
    acc += V->dim[f];
  }
  V->ystart[P->n] = acc;
  if (acc != P->d) return -1;
  V->nnz_rows = P->obs_full ? P->m * P->n : P->row_ptr[P->m];
  V->nnz_cols = P->obs_full ? P->m * P->n : P->col_ptr[P->n];
  return 0;
}
static void view_free(view_t* V) { free(V->ystart); free(V->dim); }

/* accessors over observed_features[e] / observed_examples[f] */
static inline int64_t row_len(const view_t* V, int64_t e) {
  return V->P->obs_full ? V->n : V->P->row_ptr[e + 1] - V->P->row_ptr[e];
}
static inline int64_t col_len(const view_t* V, int64_t f) {
  return V->P->obs_full ? V->m : V->P->col_ptr[f + 1] - V->P->col_ptr[f];
}
static inline void row_entry(const view_t* V, int64_t e, int64_t t, int64_t* f, double* a) {
  if (V->P->obs_full) { *f = t; *a = V->P->dense_A[t * V->m + e]; }
  else { const int64_t q = V->P->row_ptr[e] + t; *f = V->P->row_idx[q]; *a = V->P->row_val[q]; }
}
static inline void col_entry(const view_t* V, int64_t f, int64_t t, int64_t* e, double* a) {
  if (V->P->obs_full) { *e = t; *a = V->P->dense_A[f * V->m + t]; }
  else { const int64_t q = V->P->col_ptr[f] + t; *e = V->P->col_idx[q]; *a = V->P->col_val[q]; }
}
static inline const int32_t* rx_code(const view_t* V, int64_t e) { return V->P->rx_code + (V->P->rx_count == 1 ? 0 : e); }
static inline const double* rx_par(const view_t* V, int64_t e) { return V->P->rx_param + (V->P->rx_count == 1 ? 0 : e) * GLRMB200_REG_NPARAM; }
static inline const int32_t* ry_code(const view_t* V, int64_t f) { return V->P->ry_code + (V->P->ry_count == 1 ? 0 : f); }
static inline const double* ry_par(const view_t* V, int64_t f) { return V->P->ry_param + (V->P->ry_count == 1 ? 0 : f) * GLRMB200_REG_NPARAM; }
/* vector payloads (fixed_latent_features.y, RemQuadReg.m): regularizers.jl:193-231,412-423 */
static inline const double* rx_pay(const view_t* V, int64_t e) {
  return V->P->rx_payload_ptr ? V->P->rx_payload + V->P->rx_payload_ptr[V->P->rx_count == 1 ? 0 : e] : NULL;
}
static inline int64_t rx_npay(const view_t* V, int64_t e) {
  const int64_t i = V->P->rx_count == 1 ? 0 : e;
  return V->P->rx_payload_ptr ? V->P->rx_payload_ptr[i + 1] - V->P->rx_payload_ptr[i] : 0;
}
static inline const double* ry_pay(const view_t* V, int64_t f) {
  return V->P->ry_payload_ptr ? V->P->ry_payload + V->P->ry_payload_ptr[V->P->ry_count == 1 ? 0 : f] : NULL;
}
static inline int64_t ry_npay(const view_t* V, int64_t f) {
  const int64_t i = V->P->ry_count == 1 ? 0 : f;
  return V->P->ry_payload_ptr ? V->P->ry_payload_ptr[i + 1] - V->P->ry_payload_ptr[i] : 0;
}

static inline double dotk(const double* a, const double* b, int64_t k) {
  double s = 0;
#pragma omp simd reduction(+ : s)
  for (int64_t r = 0; r < k; ++r) s += a[r] * b[r];
  return s;
}

/* XY = X'Y, stored row-major per example: XY[e*d + c]  (gemm!('T','N',...), proxgrad.jl:66).
 * The reference calls BLAS dgemm (OpenBLAS ships with Julia).  When the host has handed over a cblas_dgemm entry point
 * (oracle_set_dgemm: bench.py passes the one of NumPy's bundled OpenBLAS, ILP64 interface) it is used here, so the
 * CPU baseline pays what the reference pays for this product; otherwise a plain dot-product loop computes it. */
typedef void (*cblas_dgemm64_fn)(int order, int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha,
                                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
static cblas_dgemm64_fn g_dgemm = NULL;
void oracle_set_dgemm(void* fn) { g_dgemm = (cblas_dgemm64_fn)fn; }
int oracle_has_dgemm(void) { return g_dgemm != NULL; }

/* out[(e - e0) * (c1 - c0) + (c - c0)] = x_e . y_c for rows e0 <= e < e1, embedding columns c0 <= c < c1 */
static void gemm_part(const view_t* V, const double* X, const double* Y, double* out, int64_t e0, int64_t e1, int64_t c0, int64_t c1) {
  const int64_t k = V->k, nc = c1 - c0, nr = e1 - e0;
  if (nc <= 0 || nr <= 0) return;
  if (g_dgemm) {
    /* out (row-major nr x nc) == column-major nc x nr == Y[:,c0:c1]' X[:,e0:e1]: CblasColMajor=102, CblasTrans=112, CblasNoTrans=111 */
    g_dgemm(102, 112, 111, nc, nr, k, 1.0, Y + c0 * k, k, X + e0 * k, k, 0.0, out, nc);
    return;
  }
#pragma omp parallel for schedule(static)
  for (int64_t eb = e0; eb < e1; eb += 8) {
    const int64_t ee = eb + 8 < e1 ? eb + 8 : e1;
    for (int64_t c = c0; c < c1; ++c) {
      const double* y = Y + c * k;
      for (int64_t e = eb; e < ee; ++e) out[(e - e0) * nc + (c - c0)] = dotk(X + e * k, y, k);
    }
  }
}

/* row_objective (evaluate_fit.jl:24-38).  faithful: dense x'*Y first (xy has room for d). */
static double row_objective(const view_t* V, int64_t e, const double* x, const double* Y, int faithful,
                            double* xy, int* err) {
  const int64_t k = V->k;
  double obj = 0.0;
  if (faithful) for (int64_t c = 0; c < V->d; ++c) xy[c] = dotk(x, Y + c * k, k);
  const int64_t len = row_len(V, e);
  for (int64_t t = 0; t < len; ++t) {
    int64_t f; double a;
    row_entry(V, e, t, &f, &a);
    const int D = V->dim[f];
    const double* lp = V->P->loss_param + f * GLRMB200_LOSS_NPARAM;
    if (faithful) {
      obj += loss_eval(V->P->loss_code[f], lp, xy + V->ystart[f], D, a, err);
    } else {
      double u[ORACLE_MAX_D];
      for (int c = 0; c < D; ++c) u[c] = dotk(x, Y + (V->ystart[f] + c) * k, k);
      obj += loss_eval(V->P->loss_code[f], lp, u, D, a, err);
    }
  }
  obj += reg_eval_p(*rx_code(V, e), rx_par(V, e), x, k, 1, rx_pay(V, e), rx_npay(V, e));
  return obj;
}

/* col_objective (evaluate_fit.jl:39-55).  faithful: dense X'*y first (xy has room for m*D). */
static double col_objective(const view_t* V, int64_t f, const double* yblk, const double* X, int faithful,
                            double* xy, int* err) {
  const int64_t k = V->k, m = V->m;
  const int D = V->dim[f];
  const double* lp = V->P->loss_param + f * GLRMB200_LOSS_NPARAM;
  double obj = 0.0;
  if (faithful)
    for (int64_t e = 0; e < m; ++e)
      for (int c = 0; c < D; ++c) xy[e * D + c] = dotk(X + e * k, yblk + c * k, k);
  const int64_t len = col_len(V, f);
  for (int64_t t = 0; t < len; ++t) {
    int64_t e; double a;
    col_entry(V, f, t, &e, &a);
    if (faithful) {
      obj += loss_eval(V->P->loss_code[f], lp, xy + e * D, D, a, err);
    } else {
      double u[ORACLE_MAX_D];
      for (int c = 0; c < D; ++c) u[c] = dotk(X + e * k, yblk + c * k, k);
      obj += loss_eval(V->P->loss_code[f], lp, u, D, a, err);
    }
  }
  obj += reg_eval_p(*ry_code(V, f), ry_par(V, f), yblk, k, D, ry_pay(V, f), ry_npay(V, f));
  return obj;
}

/* objective(glrm, X, Y[, XY]) (evaluate_fit.jl:4-23,57-81) + calc_penalty (:91-104) */
static double full_objective(const view_t* V, const double* X, const double* Y, int include_reg, int* err) {
  const int64_t k = V->k;
  double total = 0.0;
  int errs = 0;
#pragma omp parallel for schedule(static) reduction(+ : total) reduction(| : errs)
  for (int64_t f = 0; f < V->n; ++f) {
    const int D = V->dim[f];
    const double* lp = V->P->loss_param + f * GLRMB200_LOSS_NPARAM;
    const double* yblk = Y + V->ystart[f] * k;
    double acc = 0.0;
    int le = 0;
    const int64_t len = col_len(V, f);
    for (int64_t t = 0; t < len; ++t) {
      int64_t e; double a;
      col_entry(V, f, t, &e, &a);
      double u[ORACLE_MAX_D];
      for (int c = 0; c < D; ++c) u[c] = dotk(X + e * k, yblk + c * k, k);
      acc += loss_eval(V->P->loss_code[f], lp, u, D, a, &le);
    }
    if (include_reg) acc += reg_eval_p(*ry_code(V, f), ry_par(V, f), yblk, k, D, ry_pay(V, f), ry_npay(V, f));
    total += acc;
    errs |= le;
  }
  if (include_reg) {
    double pen = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pen)
    for (int64_t e = 0; e < V->m; ++e) pen += reg_eval_p(*rx_code(V, e), rx_par(V, e), X + e * k, k, 1, rx_pay(V, e), rx_npay(V, e));
    total += pen;
  }
  if (errs) *err = errs;
  return total;
}

double oracle_objective(const glrmb200_problem* P, const double* X, const double* Y, int include_reg, int* err) {
  view_t V;
  int e = 0;
  if (view_init(&V, P)) { if (err) *err = 4; return NAN; }
  const double r = full_objective(&V, X, Y, include_reg, &e);
  view_free(&V);
  if (err) *err = e;
  return r;
}

/* ---- the fit loop (proxgrad.jl:34-220) -------------------------------------------------------- *
 * mode: 0 faithful (dense XY), 1 sparse-evaluated.   nthreads<=0: OpenMP default.
 * alpharow/alphacol (optional, may be NULL) receive the final step sizes; trials[2] (optional)
 * the number of line-search trial evaluations in X and Y sweeps.
 * returns 0, or 1 label error, 2 unknown loss, 3 dim too large, 4 bad problem, 5 cap too small. */
static int fit_impl(const glrmb200_problem* P, const glrmb200_params* prm, double* X, double* Y,
                    double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
                    int32_t mode, int32_t nthreads, double* alpharow_out, double* alphacol_out,
                    int64_t* trials, int64_t rb, int64_t re, int64_t cb, int64_t ce) {
  view_t V;
  if (view_init(&V, P)) return 4;
  if (rb < 0) { rb = 0; re = V.m; cb = 0; ce = V.n; }
  if (rb > re || re > V.m || cb > ce || ce > V.n) { view_free(&V); return 4; }
  const int full = (rb == 0 && re == V.m && cb == 0 && ce == V.n);
  if (cap < prm->max_iter + 1) { view_free(&V); return 5; }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const int nth = omp_get_max_threads();
#else
  const int nth = 1;
#endif
  const int faithful = (mode == 0);
  const int64_t m = V.m, n = V.n, k = V.k, d = V.d;
  int err = 0;
  int64_t tx = 0, ty = 0;

  /* the product X'Y (:65-66).  Unit sample (bench.py's bounded CPU step): only the rows [rb,re) are swept in X and the
   * columns [cb,ce) in Y, so only its rows rb:re (XYr) and its columns of cb:ce (XYc) are ever read; each is refreshed
   * where the reference refreshes the whole product (:157 after the X sweep feeds the Y sweep, :202 the next X sweep). */
  double* XYr = NULL; double* XYc = NULL;
  const int64_t coff = full ? 0 : V.ystart[cb];
  const int64_t ldc = full ? d : (V.ystart[ce] - V.ystart[cb]);
  if (faithful) {
    XYr = (double*)malloc(sizeof(double) * (size_t)(re - rb > 0 ? re - rb : 1) * (size_t)d);
    XYc = full ? XYr : (double*)malloc(sizeof(double) * (size_t)m * (size_t)(ldc > 0 ? ldc : 1));
    if (!XYr || !XYc) { view_free(&V); return 4; }
    gemm_part(&V, X, Y, XYr, rb, re, 0, d);
    if (!full) gemm_part(&V, X, Y, XYc, 0, m, coff, coff + ldc);
  }
#define XR(e, col) XYr[((e) - rb) * d + (col)]
#define XC(e, col) XYc[(e) * ldc + (col) - coff]
  double* alpharow = (double*)malloc(sizeof(double) * (size_t)m);
  double* alphacol = (double*)malloc(sizeof(double) * (size_t)n);
  for (int64_t e = 0; e < m; ++e) alpharow[e] = prm->stepsize;                 /* :69 */
  for (int64_t f = 0; f < n; ++f) alphacol[f] = prm->stepsize;                 /* :70 */
  const double scaled_abs_tol = prm->abs_tol * (double)V.nnz_rows;             /* :72 */
  double* obj_by_col = (double*)calloc((size_t)n, sizeof(double));
  int maxD = 1;
  for (int64_t f = 0; f < n; ++f) if (V.dim[f] > maxD) maxD = V.dim[f];

  int nrec = 0;
  ch_objective[nrec] = full ? full_objective(&V, X, Y, 1, &err) : 0.0;         /* :76 */
  ch_seconds[nrec++] = 0.0;
  double t0 = now_s();

  /* per-thread scratch: g (k*maxD), newv (k*maxD), xy (max(d, m*maxD) if faithful) */
  const size_t xy_len = faithful ? (size_t)((d > m * maxD) ? d : m * maxD) : 1;
  double** scr_g = (double**)malloc(sizeof(double*) * (size_t)nth);
  double** scr_new = (double**)malloc(sizeof(double*) * (size_t)nth);
  double** scr_xy = (double**)malloc(sizeof(double*) * (size_t)nth);
  for (int t = 0; t < nth; ++t) {
    scr_g[t] = (double*)malloc(sizeof(double) * (size_t)(k * maxD));
    scr_new[t] = (double*)malloc(sizeof(double) * (size_t)(k * maxD));
    scr_xy[t] = (double*)malloc(sizeof(double) * xy_len);
  }

  for (int it = 1; it <= prm->max_iter; ++it) {                                /* :107 */
    if (prm->inner_iter_X > 1 || prm->inner_iter_Y > 1) {                      /* :112-115 */
      for (int64_t e = 0; e < m; ++e) alpharow[e] = prm->stepsize;
      for (int64_t f = 0; f < n; ++f) alphacol[f] = prm->stepsize;
    }
    /* STEP 1: X update ------------------------------------------------------------- :117-158 */
    for (int inner = 0; inner < prm->inner_iter_X; ++inner) {
#pragma omp parallel for schedule(static) reduction(+ : tx) reduction(| : err)
      for (int64_t e = rb; e < re; ++e) {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* g = scr_g[tid];
        double* newx = scr_new[tid];
        double* xe = X + e * k;
        int le = 0;
        for (int64_t r = 0; r < k; ++r) g[r] = 0.0;                           /* :119 */
        const int64_t len = row_len(&V, e);
        for (int64_t t = 0; t < len; ++t) {                                    /* :122-132 */
          int64_t f; double a;
          row_entry(&V, e, t, &f, &a);
          const int D = V.dim[f];
          const double* lp = P->loss_param + f * GLRMB200_LOSS_NPARAM;
          double u[ORACLE_MAX_D], cg[ORACLE_MAX_D];
          if (faithful) for (int c = 0; c < D; ++c) u[c] = XR(e, V.ystart[f] + c);
          else for (int c = 0; c < D; ++c) u[c] = dotk(xe, Y + (V.ystart[f] + c) * k, k);
          loss_grad(P->loss_code[f], lp, u, D, a, cg, &le);                    /* :125 */
          for (int c = 0; c < D; ++c) {                                        /* :127 / :130 */
            const double* yc = Y + (V.ystart[f] + c) * k;
            const double cc = cg[c];
            for (int64_t r = 0; r < k; ++r) g[r] += cc * yc[r];
          }
        }
        const double l = (double)(len + 1);                                    /* :134 */
        const double obj_old = row_objective(&V, e, xe, Y, faithful, scr_xy[tid], &le); /* :135 */
        memcpy(newx, xe, sizeof(double) * (size_t)k);
        while (alpharow[e] > prm->min_stepsize) {                              /* :136 */
          const double stepsize = alpharow[e] / l;                             /* :137 */
          for (int64_t r = 0; r < k; ++r) newx[r] += -stepsize * g[r];         /* :140 */
          reg_prox_p(*rx_code(&V, e), rx_par(&V, e), newx, k, 1, stepsize, rx_pay(&V, e), rx_npay(&V, e));      /* :142 */
          tx++;
          if (row_objective(&V, e, newx, Y, faithful, scr_xy[tid], &le) < obj_old) { /* :143 */
            memcpy(xe, newx, sizeof(double) * (size_t)k);                      /* :144 */
            alpharow[e] *= 1.05;                                               /* :145 */
            break;
          } else {
            memcpy(newx, xe, sizeof(double) * (size_t)k);                      /* :148 */
            alpharow[e] *= .7;                                                 /* :149 */
            if (alpharow[e] < prm->min_stepsize) {                             /* :150-153 */
              alpharow[e] = prm->min_stepsize * 1.1;
              break;
            }
          }
        }
        err |= le;
      }
      if (faithful) { if (full) gemm_part(&V, X, Y, XYr, 0, m, 0, d); else gemm_part(&V, X, Y, XYc, 0, m, coff, coff + ldc); }   /* :157 */
    }
    /* STEP 2: Y update ------------------------------------------------------------- :160-203 */
    for (int inner = 0; inner < prm->inner_iter_Y; ++inner) {
#pragma omp parallel for schedule(static) reduction(+ : ty) reduction(| : err)
      for (int64_t f = cb; f < ce; ++f) {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        const int D = V.dim[f];
        double* G = scr_g[tid];
        double* newy = scr_new[tid];
        double* yf = Y + V.ystart[f] * k;
        const double* lp = P->loss_param + f * GLRMB200_LOSS_NPARAM;
        int le = 0;
        for (int64_t r = 0; r < k * D; ++r) G[r] = 0.0;                        /* :161 */
        const int64_t len = col_len(&V, f);
        for (int64_t t = 0; t < len; ++t) {                                    /* :165-175 */
          int64_t e; double a;
          col_entry(&V, f, t, &e, &a);
          const double* xe = X + e * k;
          double u[ORACLE_MAX_D], cg[ORACLE_MAX_D];
          if (faithful) for (int c = 0; c < D; ++c) u[c] = XC(e, V.ystart[f] + c);
          else for (int c = 0; c < D; ++c) u[c] = dotk(xe, yf + c * k, k);
          loss_grad(P->loss_code[f], lp, u, D, a, cg, &le);                    /* :168 */
          for (int c = 0; c < D; ++c) {                                        /* :170 / :173 */
            const double cc = cg[c];
            for (int64_t r = 0; r < k; ++r) G[c * k + r] += cc * xe[r];
          }
        }
        const double l = (double)(len + 1);                                    /* :177 */
        obj_by_col[f] = col_objective(&V, f, yf, X, faithful, scr_xy[tid], &le); /* :178 */
        memcpy(newy, yf, sizeof(double) * (size_t)(k * D));
        while (alphacol[f] > prm->min_stepsize) {                              /* :179 */
          const double stepsize = alphacol[f] / l;                             /* :180 */
          for (int64_t r = 0; r < k * D; ++r) newy[r] += -stepsize * G[r];     /* :183 */
          reg_prox_p(*ry_code(&V, f), ry_par(&V, f), newy, k, D, stepsize, ry_pay(&V, f), ry_npay(&V, f));      /* :185 */
          ty++;
          const double new_obj = col_objective(&V, f, newy, X, faithful, scr_xy[tid], &le); /* :186 */
          if (new_obj < obj_by_col[f]) {                                       /* :187 */
            memcpy(yf, newy, sizeof(double) * (size_t)(k * D));                /* :188 */
            alphacol[f] *= 1.05;                                               /* :189 */
            obj_by_col[f] = new_obj;                                           /* :190 */
            break;
          } else {
            memcpy(newy, yf, sizeof(double) * (size_t)(k * D));                /* :193 */
            alphacol[f] *= .7;                                                 /* :194 */
            if (alphacol[f] < prm->min_stepsize) {                             /* :195-198 */
              alphacol[f] = prm->min_stepsize * 1.1;
              break;
            }
          }
        }
        err |= le;
      }
      if (faithful) gemm_part(&V, X, Y, XYr, rb, re, 0, d);                    /* :202 */
    }
    /* STEP 3: record objective ------------------------------------------------------ :204-208 */
    double obj = 0.0;
    for (int64_t f = cb; f < ce; ++f) obj += obj_by_col[f];                    /* :205 */
    const double t1 = now_s();
    ch_objective[nrec] = obj;
    ch_seconds[nrec++] = t1 - t0;                                              /* :206-207 */
    t0 = now_s();
    /* STEP 4: stopping criterion ---------------------------------------------------- :209-213 */
    const double obj_decrease = ch_objective[nrec - 2] - obj;
    if (it > 10 && (obj_decrease < scaled_abs_tol || obj_decrease / obj < prm->rel_tol)) break;
  }

  *n_recorded = nrec;
  if (alpharow_out) memcpy(alpharow_out, alpharow, sizeof(double) * (size_t)m);
  if (alphacol_out) memcpy(alphacol_out, alphacol, sizeof(double) * (size_t)n);
  if (trials) { trials[0] = tx; trials[1] = ty; }
  for (int t = 0; t < nth; ++t) { free(scr_g[t]); free(scr_new[t]); free(scr_xy[t]); }
  free(scr_g); free(scr_new); free(scr_xy);
  free(alpharow); free(alphacol); free(obj_by_col);
  if (XYc != XYr) free(XYc);
  free(XYr);
#undef XR
#undef XC
  view_free(&V);
  return err;
}

int oracle_fit(const glrmb200_problem* P, const glrmb200_params* prm, double* X, double* Y,
               double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
               int32_t mode, int32_t nthreads, double* alpharow_out, double* alphacol_out,
               int64_t* trials) {
  return fit_impl(P, prm, X, Y, ch_objective, ch_seconds, cap, n_recorded, mode, nthreads, alpharow_out, alphacol_out,
                  trials, -1, -1, -1, -1);
}

/* The same loop over a sample of the units only: rows [rb, re) in the X sweeps, columns [cb, ce) in the Y sweeps; all
 * other columns of X and Y stay frozen.  bench.py's reference arm times this on the full-size problem (a bounded sample of
 * the workload); the recorded "objective" is the sampled columns' share and ch_objective[0] is 0. */
int oracle_fit_units(const glrmb200_problem* P, const glrmb200_params* prm, double* X, double* Y,
                     double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded,
                     int32_t mode, int32_t nthreads, int64_t rb, int64_t re, int64_t cb, int64_t ce, int64_t* trials) {
  return fit_impl(P, prm, X, Y, ch_objective, ch_seconds, cap, n_recorded, mode, nthreads, NULL, NULL, trials, rb, re, cb, ce);
}

/* One half-sweep over the units [begin, end) only (sparse-evaluated form): which = 0 updates rows of X
 * (proxgrad.jl:118-156), which = 1 updates columns of Y (:162-201).  alpha is alpharow / alphacol (in/out,
 * full length), obj_by_unit receives obj_by_row / obj_by_col for the touched units.  Used by the world_size-2
 * CPU tests to show that "each rank sweeps its shard, then all-gather" reproduces the unsharded sweep. */
int oracle_half_sweep(const glrmb200_problem* P, const glrmb200_params* prm, double* X, double* Y,
                      double* alpha, int32_t which, int64_t begin, int64_t end, double* obj_by_unit) {
  view_t V;
  if (view_init(&V, P)) return 4;
  const int64_t k = V.k;
  int err = 0;
  int maxD = 1;
  for (int64_t f = 0; f < V.n; ++f) if (V.dim[f] > maxD) maxD = V.dim[f];
  double* g = (double*)malloc(sizeof(double) * (size_t)(k * maxD));
  double* nw = (double*)malloc(sizeof(double) * (size_t)(k * maxD));
  double xy_dummy = 0;
  for (int64_t u = begin; u < end; ++u) {
    if (which == 0) {
      const int64_t e = u;
      double* xe = X + e * k;
      for (int64_t r = 0; r < k; ++r) g[r] = 0.0;
      const int64_t len = row_len(&V, e);
      for (int64_t t = 0; t < len; ++t) {
        int64_t f; double a;
        row_entry(&V, e, t, &f, &a);
        const int D = V.dim[f];
        double uu[ORACLE_MAX_D], cg[ORACLE_MAX_D];
        for (int c = 0; c < D; ++c) uu[c] = dotk(xe, Y + (V.ystart[f] + c) * k, k);
        loss_grad(P->loss_code[f], P->loss_param + f * GLRMB200_LOSS_NPARAM, uu, D, a, cg, &err);
        for (int c = 0; c < D; ++c) for (int64_t r = 0; r < k; ++r) g[r] += cg[c] * Y[(V.ystart[f] + c) * k + r];
      }
      const double l = (double)(len + 1);
      const double obj_old = row_objective(&V, e, xe, Y, 0, &xy_dummy, &err);
      obj_by_unit[e] = obj_old;
      memcpy(nw, xe, sizeof(double) * (size_t)k);
      while (alpha[e] > prm->min_stepsize) {
        const double stepsize = alpha[e] / l;
        for (int64_t r = 0; r < k; ++r) nw[r] += -stepsize * g[r];
        reg_prox_p(*rx_code(&V, e), rx_par(&V, e), nw, k, 1, stepsize, rx_pay(&V, e), rx_npay(&V, e));
        if (row_objective(&V, e, nw, Y, 0, &xy_dummy, &err) < obj_old) {
          memcpy(xe, nw, sizeof(double) * (size_t)k);
          alpha[e] *= 1.05;
          break;
        } else {
          memcpy(nw, xe, sizeof(double) * (size_t)k);
          alpha[e] *= .7;
          if (alpha[e] < prm->min_stepsize) { alpha[e] = prm->min_stepsize * 1.1; break; }
        }
      }
    } else {
      const int64_t f = u;
      const int D = V.dim[f];
      double* yf = Y + V.ystart[f] * k;
      for (int64_t r = 0; r < k * D; ++r) g[r] = 0.0;
      const int64_t len = col_len(&V, f);
      for (int64_t t = 0; t < len; ++t) {
        int64_t e; double a;
        col_entry(&V, f, t, &e, &a);
        const double* xe = X + e * k;
        double uu[ORACLE_MAX_D], cg[ORACLE_MAX_D];
        for (int c = 0; c < D; ++c) uu[c] = dotk(xe, yf + c * k, k);
        loss_grad(P->loss_code[f], P->loss_param + f * GLRMB200_LOSS_NPARAM, uu, D, a, cg, &err);
        for (int c = 0; c < D; ++c) for (int64_t r = 0; r < k; ++r) g[c * k + r] += cg[c] * xe[r];
      }
      const double l = (double)(len + 1);
      obj_by_unit[f] = col_objective(&V, f, yf, X, 0, &xy_dummy, &err);
      memcpy(nw, yf, sizeof(double) * (size_t)(k * D));
      while (alpha[f] > prm->min_stepsize) {
        const double stepsize = alpha[f] / l;
        for (int64_t r = 0; r < k * D; ++r) nw[r] += -stepsize * g[r];
        reg_prox_p(*ry_code(&V, f), ry_par(&V, f), nw, k, D, stepsize, ry_pay(&V, f), ry_npay(&V, f));
        const double new_obj = col_objective(&V, f, nw, X, 0, &xy_dummy, &err);
        if (new_obj < obj_by_unit[f]) {
          memcpy(yf, nw, sizeof(double) * (size_t)(k * D));
          alpha[f] *= 1.05;
          obj_by_unit[f] = new_obj;
          break;
        } else {
          memcpy(nw, yf, sizeof(double) * (size_t)(k * D));
          alpha[f] *= .7;
          if (alpha[f] < prm->min_stepsize) { alpha[f] = prm->min_stepsize * 1.1; break; }
        }
      }
    }
  }
  free(g); free(nw);
  view_free(&V);
  return err;
}

/* fit!(glrm, ::SparseProxGradParams) — /root/reference/src/algorithms/sparse_proxgrad.jl:21-130.
 * sp = {stepsize, max_iter, inner_iter, abs_tol, min_stepsize} (the struct of include/glrm_b200.h).
 * X, Y in/out (the best model, like glrm.X / glrm.Y); scalar-embedding losses only (the reference uses dot(x_e,y_f)).
 * cap >= max_iter + 2.  returns 0 or an error code as oracle_fit. */
int oracle_fit_sparse(const glrmb200_problem* P, const glrmb200_sparse_params* sp, double* Xbest, double* Ybest,
                      double* ch_objective, double* ch_seconds, int32_t cap, int32_t* n_recorded, double* alpha_out) {
  view_t V;
  if (view_init(&V, P)) return 4;
  if (cap < sp->max_iter + 2 || V.d != V.n) { view_free(&V); return 5; }
  const int64_t m = V.m, n = V.n, k = V.k;
  int err = 0;
  double* X = (double*)malloc(sizeof(double) * (size_t)(k * m));
  double* Y = (double*)malloc(sizeof(double) * (size_t)(k * n));
  memcpy(X, Xbest, sizeof(double) * (size_t)(k * m));                          /* :33 */
  memcpy(Y, Ybest, sizeof(double) * (size_t)(k * n));
  double alpha = sp->stepsize;                                                 /* :44 */
  const double tol = sp->abs_tol * (double)V.nnz_rows;                         /* :46 */
  int nrec = 0;
  ch_objective[nrec] = full_objective(&V, Xbest, Ybest, 1, &err);             /* :50 */
  ch_seconds[nrec++] = 0.0;
  double t = now_s();
  int steps_in_a_row = 0;
  double* g = (double*)malloc(sizeof(double) * (size_t)k);
  for (int i = 1; i <= sp->max_iter; ++i) {                                    /* :60 */
    for (int inner = 0; inner < sp->inner_iter; ++inner) {                     /* :62 */
      for (int64_t e = 0; e < m; ++e) {                                        /* :63 */
        double* xe = X + e * k;
        for (int64_t r = 0; r < k; ++r) g[r] *= 0;                             /* :64 */
        for (int64_t r = 0; r < k; ++r) g[r] = 0.0;
        const int64_t len = row_len(&V, e);
        for (int64_t q = 0; q < len; ++q) {                                    /* :67-72 */
          int64_t f; double a;
          row_entry(&V, e, q, &f, &a);
          const double* yf = Y + f * k;
          const double c = sl_grad(P->loss_code[f], P->loss_param + f * GLRMB200_LOSS_NPARAM, dotk(xe, yf, k), a, &err);
          for (int64_t r = 0; r < k; ++r) g[r] += c * yf[r];
        }
        const double l = (double)(len + 1);                                    /* :74 */
        for (int64_t r = 0; r < k; ++r) g[r] *= -alpha / l;                    /* :75 */
        for (int64_t r = 0; r < k; ++r) xe[r] += g[r];                         /* :77 */
        reg_prox_p(*rx_code(&V, e), rx_par(&V, e), xe, k, 1, alpha / l, rx_pay(&V, e), rx_npay(&V, e));         /* :79 */
      }
    }
    for (int inner = 0; inner < sp->inner_iter; ++inner) {                     /* :83 */
      for (int64_t f = 0; f < n; ++f) {                                        /* :84 */
        double* yf = Y + f * k;
        for (int64_t r = 0; r < k; ++r) g[r] = 0.0;                            /* :85 */
        const int64_t len = col_len(&V, f);
        for (int64_t q = 0; q < len; ++q) {                                    /* :88-92 */
          int64_t e; double a;
          col_entry(&V, f, q, &e, &a);
          const double* xe = X + e * k;
          const double c = sl_grad(P->loss_code[f], P->loss_param + f * GLRMB200_LOSS_NPARAM, dotk(xe, yf, k), a, &err);
          for (int64_t r = 0; r < k; ++r) g[r] += c * xe[r];
        }
        const double l = (double)(len + 1);                                    /* :94 */
        for (int64_t r = 0; r < k; ++r) g[r] *= -alpha / l;                    /* :95 */
        for (int64_t r = 0; r < k; ++r) yf[r] += g[r];                         /* :97 */
        reg_prox_p(*ry_code(&V, f), ry_par(&V, f), yf, k, 1, alpha / l, ry_pay(&V, f), ry_npay(&V, f));         /* :99 */
      }
    }
    const double obj = full_objective(&V, X, Y, 1, &err);                      /* :102 */
    if (obj < ch_objective[nrec - 1]) {                                        /* :104 */
      const double now = now_s();
      ch_objective[nrec] = obj;
      ch_seconds[nrec++] = now - t;                                            /* :105-106 */
      memcpy(Xbest, X, sizeof(double) * (size_t)(k * m));                      /* :107 */
      memcpy(Ybest, Y, sizeof(double) * (size_t)(k * n));
      alpha = alpha * 1.05;                                                    /* :108 */
      steps_in_a_row = steps_in_a_row + 1 > 1 ? steps_in_a_row + 1 : 1;        /* :109 */
      t = now_s();
    } else {
      const double div = (double)(-steps_in_a_row) > 1.5 ? (double)(-steps_in_a_row) : 1.5;
      alpha = alpha / div;                                                     /* :113 */
      memcpy(X, Xbest, sizeof(double) * (size_t)(k * m));                      /* :115 */
      memcpy(Y, Ybest, sizeof(double) * (size_t)(k * n));
      steps_in_a_row = steps_in_a_row - 1 < 0 ? steps_in_a_row - 1 : 0;        /* :116 */
    }
    const double prev = nrec >= 2 ? ch_objective[nrec - 2] : INFINITY;
    if ((i > 10 && (steps_in_a_row > 3 && prev - obj < tol)) || alpha <= sp->min_stepsize) break;   /* :119 */
  }
  ch_objective[nrec] = ch_objective[nrec - 1];                                 /* :125-126 */
  ch_seconds[nrec] = now_s() - t;
  ++nrec;
  *n_recorded = nrec;
  if (alpha_out) *alpha_out = alpha;
  free(g); free(X); free(Y);
  view_free(&V);
  return err;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
