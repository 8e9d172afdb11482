/*
 * synth_pattern.c — host-side synthetic sparsity-pattern generator for bench.py / tests
 * (data generation only; not on the accelerated path).  Same counter-based SplitMix64 stream as
 * lowrankmodels.jl_b200/synth.py::uniform, so Python can reproduce any draw.
 *
 * For every row e it draws columns from the popularity CDF until deg[e] DISTINCT columns are found,
 * then emits all pairs in CSC order (columns ascending, rows ascending inside a column) — the order
 * `findall(!iszero, A)` gives for a SparseMatrixCSC (reference src/glrm.jl:46-48).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t key, uint64_t idx) {
  const uint64_t h = splitmix64(idx ^ key);
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* key = synth.py::_key(seed, 13).  Row e uses counters e*2^26 + t. Returns 0, or -1 on failure. */
int glrm_synth_pattern(int64_t m, int64_t n, const int64_t* deg, const double* cdf,
                       const int64_t* relabel, uint64_t key, int64_t* out_rows, int64_t* out_cols) {
  int64_t* rptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m + 1));
  rptr[0] = 0;
  for (int64_t e = 0; e < m; ++e) rptr[e + 1] = rptr[e] + deg[e];
  const int64_t nnz = rptr[m];
  int32_t* rcols = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  int fail = 0;
#pragma omp parallel
  {
    uint8_t* seen = (uint8_t*)calloc((size_t)n, 1);
#pragma omp for schedule(dynamic, 64)
    for (int64_t e = 0; e < m; ++e) {
      int32_t* dst = rcols + rptr[e];
      int64_t have = 0;
      const int64_t want = deg[e];
      uint64_t t = 0;
      const uint64_t base = (uint64_t)e << 26;
      while (have < want) {
        if (t >= (1ULL << 26)) { fail = 1; break; }
        const double u = u01(key, base + t++);
        int64_t lo = 0, hi = n - 1; /* first index with cdf[idx] >= u */
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (cdf[mid] < u) lo = mid + 1; else hi = mid;
        }
        const int64_t c = relabel[lo];
        if (!seen[c]) { seen[c] = 1; dst[have++] = (int32_t)c; }
      }
      for (int64_t i = 0; i < have; ++i) seen[dst[i]] = 0;
    }
    free(seen);
  }
  if (fail) { free(rptr); free(rcols); return -1; }
  /* counting sort by column; rows visited ascending => rows ascending inside each column */
  int64_t* cptr = (int64_t*)calloc((size_t)(n + 1), sizeof(int64_t));
  for (int64_t q = 0; q < nnz; ++q) cptr[rcols[q] + 1]++;
  for (int64_t c = 0; c < n; ++c) cptr[c + 1] += cptr[c];
  for (int64_t e = 0; e < m; ++e)
    for (int64_t q = rptr[e]; q < rptr[e + 1]; ++q) {
      const int64_t pos = cptr[rcols[q]]++;
      out_rows[pos] = e;
      out_cols[pos] = rcols[q];
    }
  free(cptr); free(rptr); free(rcols);
  return 0;
}

/* ---- dense synthetic matrices for configs 4 and 5 (synth.py::config4 / config5), filled in parallel ------------------ */
#include <math.h>
static inline double n01(uint64_t key, uint64_t idx) {       /* synth.py::normal: Box-Muller on counters 2 idx, 2 idx + 1 */
  const double u1 = u01(key, idx * 2ULL), u2 = u01(key, idx * 2ULL + 1ULL);
  return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

/* C5: A[i + m j] = Cn[z_i + centroids j] + 0.1 N(0,1)(counter j m + i); Cn is (centroids x n) column-major */
void glrm_synth_c5(int64_t m, int64_t n, int64_t centroids, const double* Cn, const int64_t* z, uint64_t key, double* A) {
#pragma omp parallel for schedule(static) collapse(1)
  for (int64_t j = 0; j < n; ++j) {
    const double* c = Cn + j * centroids;
    double* a = A + j * m;
    for (int64_t i = 0; i < m; ++i) a[i] = c[z[i]] + 0.1 * n01(key, (uint64_t)(j * m + i));
  }
}

/* C4: base = P Q / 2 (P m x 4 column-major, Q 4 x n column-major);
 *     columns [0, nq): base + 0.3 N(key_n); [nq, nq + nh): sign(base + 0.3 N(key_n)) in {-1, +1};
 *     the rest: clip(floor(U(key_u) levels) + 1 + clip(round(base), -2, 2), 1, levels) */
void glrm_synth_c4(int64_t m, int64_t n, int64_t nq, int64_t nh, int64_t levels, const double* P, const double* Q,
                   uint64_t key_n, uint64_t key_u, double* A) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    const double q0 = Q[4 * j], q1 = Q[4 * j + 1], q2 = Q[4 * j + 2], q3 = Q[4 * j + 3];
    double* a = A + j * m;
    for (int64_t i = 0; i < m; ++i) {
      const double base = (P[i] * q0 + P[i + m] * q1 + P[i + 2 * m] * q2 + P[i + 3 * m] * q3) / 2.0;
      const uint64_t idx = (uint64_t)(j * m + i);
      if (j < nq) a[i] = base + 0.3 * n01(key_n, idx);
      else if (j < nq + nh) a[i] = (base + 0.3 * n01(key_n, idx)) >= 0.0 ? 1.0 : -1.0;
      else {
        double shift = rint(base);
        shift = shift < -2.0 ? -2.0 : (shift > 2.0 ? 2.0 : shift);
        double v = floor(u01(key_u, idx) * (double)levels) + 1.0 + shift;
        a[i] = v < 1.0 ? 1.0 : (v > (double)levels ? (double)levels : v);
      }
    }
  }
}

/* out[i] = N(0,1)(counter start + i) — synth.py::normal_matrix for the big factor initialisations */
void glrm_synth_normal_fill(uint64_t key, int64_t start, int64_t count, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < count; ++i) out[i] = n01(key, (uint64_t)(start + i));
}
