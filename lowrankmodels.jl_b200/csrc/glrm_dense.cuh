// glrm_dense.cuh — the fully observed path (glrm.observed_features == fill(1:n, m), src/glrm.jl:33-34): configs 1, 4, 5.
//
// With every entry observed the sweeps of proxgrad.jl:117-203 are dense contractions, not gathers:
//     U = X'Y                 (m x d, inner k)        what the reference's gemm! computes (proxgrad.jl:66,157,202)
//     R = dL/dU(U, A)         element-wise (per feature; vector-valued losses see the feature's block of U)
//     G_X = Y R'              (k x m, inner d)        gradient of every row          (proxgrad.jl:122-132)
//     G_Y = X R               (k x d, inner m)        gradient of every column block (proxgrad.jl:165-175)
// These kernels stream A exactly once per pass, straight from the column-major array Julia hands over (no index lists,
// no transposed copy), and never materialise U, R or a dense G_X: a CTA owns a tile of 64 rows, keeps the tile of X, a
// chunk of <= 64 columns of Y and the 64 x 64 tile of U / R in shared memory, and runs the two contractions as
// register-tiled FP64 FMA loops (8 x 4 accumulators per thread).  Float64 throughout: the line search's strict `<`
// (proxgrad.jl:143,186) needs the same arithmetic as the reference; the roof is the FP64 pipe (36.6 TFLOP/s measured),
// not tensor cores — B200's FP64 tensor rate equals its FP64 FMA rate, and an error-compensated bf16 split needs
// >= 28 partial products to carry 53 bits, i.e. less than the FMA pipe delivers directly.
//
//   dense_x_kernel      the whole X sweep for a tile: gradient pass over all chunks, then the per-row backtracking line
//                       search (each trial = one more pass over the chunks with the trial points), write-back.
//   dense_y_pass_kernel one pass of the Y sweep for (row block, chunk): partial G_Y and partial per-feature objectives
//                       (MODE 0), or objectives only for a list of features evaluated at trial blocks (MODE 1).
//   dense_y_reduce_kernel / dense_y_step_kernel / dense_y_decide_kernel / dense_y_plan_kernel
//                       fixed-order reduction over the row blocks, trial blocks prox(y - (alpha/l) g), accept / reject
//                       per feature (proxgrad.jl:179-200), and the compacted list of features still searching.
// Rows of a tile and features of a chunk keep fixed positions in every reduction, so results do not depend on the grid.
#pragma once
#include "glrm_device.cuh"
#include "glrm_vec.cuh"
#include "glrm_dense_host.h"

namespace glrm {

struct DenseSmem {
  double* Ys;    // [KT*16][DN_YP]
  double* Xs;    // [k][DN_RP]   current rows of X, [i][r]
  double* Xn;    // [k][DN_RP]   trial points
  double* Rs;    // [DN_TN][DN_RP]  U, then R, [column][row]
  double* rowv;  // [6][64] per-row scalars
  double* red;   // [4][DN_TN] per-warp partials
  int* s_col;    // [DN_TN] global Y column of local column jj (-1: unused)
  int* s_feat;   // [DN_TN] feature of chunk position p
  int* s_foff;   // [DN_TN + 1] first local column of chunk position p
  int* s_state;  // [64]
};
__host__ __device__ inline size_t dense_smem_bytes(int k, int kt) {
  return ((size_t)kt * 16 * DN_YP + 2 * (size_t)k * DN_RP + (size_t)DN_TN * DN_RP + 6 * 64 + 4 * DN_TN) * sizeof(double) +
         (3 * DN_TN + 1 + 64 + 3) * sizeof(int) + 64;
}
__device__ __forceinline__ DenseSmem dense_carve(unsigned char* base, int k, int kt) {
  DenseSmem S;
  double* p = reinterpret_cast<double*>(base);
  S.Ys = p; p += (size_t)kt * 16 * DN_YP;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) p += 1;       // 16-byte alignment for the LDS.128 tiles
  S.Xs = p; p += (size_t)k * DN_RP;
  S.Xn = p; p += (size_t)k * DN_RP;
  S.Rs = p; p += (size_t)DN_TN * DN_RP;
  S.rowv = p; p += 6 * 64;
  S.red = p; p += 4 * DN_TN;
  int* q = reinterpret_cast<int*>(p);
  S.s_col = q; q += DN_TN;
  S.s_feat = q; q += DN_TN;
  S.s_foff = q; q += DN_TN + 1;
  S.s_state = q;
  return S;
}

// ---- chunk set-up: feature list, local column map, the chunk of Y ([i][jj], rows >= k zero) ------------------------------
template <int KT>
__device__ __forceinline__ int dense_load_chunk(const DenseArgs& P, const DenseSmem& S, int c) {
  const int t = threadIdx.x;
  const int p0 = P.chunk_ptr[c], nf = P.chunk_ptr[c + 1] - p0;
  for (int jj = t; jj < DN_TN; jj += DN_THREADS) S.s_col[jj] = -1;
  __syncthreads();
  int ncols = 0;
  for (int p = t; p < nf; p += DN_THREADS) {
    const int f = P.feat_list[p0 + p], off = P.feat_off[p0 + p];
    const int64_t y0 = P.ystart[f];
    const int D = (int)(P.ystart[f + 1] - y0);
    S.s_feat[p] = f;
    S.s_foff[p] = off;
    if (p == nf - 1) S.s_foff[nf] = off + D;
    for (int cc = 0; cc < D; ++cc) S.s_col[off + cc] = (int)(y0 + cc);
  }
  __syncthreads();
  ncols = S.s_foff[nf];
  // Y chunk: lanes run over i (coalesced in global memory), one column per iteration
  for (int idx = t; idx < KT * 16 * DN_TN; idx += DN_THREADS) {
    const int jj = idx / (KT * 16), i = idx - jj * (KT * 16);
    const int col = S.s_col[jj];
    S.Ys[i * DN_YP + jj] = (col >= 0 && i < P.k) ? P.Ymat[(int64_t)col * P.stride + i] : 0.0;
  }
  __syncthreads();
  return ncols;
}

// tile of X -> Xs[i][r] (zero rows past the end)
__device__ __forceinline__ void dense_load_x(const DenseArgs& P, double* Xs, int64_t e0, int nrows) {
  const int k = P.k;
  for (int idx = threadIdx.x; idx < DN_TM * k; idx += DN_THREADS) {
    const int r = idx / k, i = idx - r * k;
    Xs[i * DN_RP + r] = r < nrows ? P.X[(e0 + r) * P.stride + i] : 0.0;
  }
}

// U[r][jj] = sum_i Xt[i][r] * Ys[i][jj]; thread (trow = t/16, tcol = t%16) owns rows trow*8..+7, columns tcol + 16 q.
// The result goes to Rs[jj][r] (4 conflict-free 16-byte stores per column).
__device__ __forceinline__ void dense_gemm_u(const double* __restrict__ Xt, const double* __restrict__ Ys, double* __restrict__ Rs, int k) {
  const int t = threadIdx.x, trow = t >> 4, tcol = t & 15;
  double acc[8][DN_CT];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < DN_CT; ++b) acc[a][b] = 0.0;
  const double* xp = Xt + trow * 8;
  const double* yp = Ys + tcol;
#pragma unroll 2
  for (int i = 0; i < k; ++i) {
    const double2 x01 = *reinterpret_cast<const double2*>(xp + i * DN_RP);
    const double2 x23 = *reinterpret_cast<const double2*>(xp + i * DN_RP + 2);
    const double2 x45 = *reinterpret_cast<const double2*>(xp + i * DN_RP + 4);
    const double2 x67 = *reinterpret_cast<const double2*>(xp + i * DN_RP + 6);
    const double x[8] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y, x67.x, x67.y};
    double y[DN_CT];
#pragma unroll
    for (int b = 0; b < DN_CT; ++b) y[b] = yp[i * DN_YP + 16 * b];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < DN_CT; ++b) acc[a][b] = fma(x[a], y[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < DN_CT; ++b) {
    double* dst = Rs + (tcol + 16 * b) * DN_RP + trow * 8;
    *reinterpret_cast<double2*>(dst) = make_double2(acc[0][b], acc[1][b]);
    *reinterpret_cast<double2*>(dst + 2) = make_double2(acc[2][b], acc[3][b]);
    *reinterpret_cast<double2*>(dst + 4) = make_double2(acc[4][b], acc[5][b]);
    *reinterpret_cast<double2*>(dst + 6) = make_double2(acc[6][b], acc[7][b]);
  }
}

// Element-wise phase on the U tile in shared memory: thread <-> (row r = t % 64, features p = t/64, t/64 + 2, ...) so the
// reads of A are coalesced down the column.  GRAD: U is replaced by dL/dU in place.  Returns this thread's share of the
// row's loss; COLSUM additionally reduces every feature's loss over the 64 rows into red[warp][p] (fixed shuffle tree).
template <int LOSS, bool GRAD, bool COLSUM>
__device__ __forceinline__ double dense_elementwise(const DenseArgs& P, const DenseSmem& S, int nf, int64_t e0, int nrows) {
  const int t = threadIdx.x, r = t & 63, half = t >> 6, warp = t >> 5;
  const bool valid = r < nrows;
  const int64_t e = e0 + (valid ? r : 0);
  double rowsum = 0.0;
  for (int p = half; p < nf; p += 2) {
    const int f = S.s_feat[p], off = S.s_foff[p], D = S.s_foff[p + 1] - off;
    const double a = P.A[(int64_t)f * P.m + e];
    const int code = LOSS ? LOSS : P.loss_code[f];
    const double* lp = P.loss_param + (int64_t)f * GLRMB200_LOSS_NPARAM;
    double l;
    if (LOSS != 0 || code < GLRMB200_LOSS_MULTINOMIAL) {
      const double u = S.Rs[off * DN_RP + r];
      double c;
      loss_eval<LOSS, GRAD>(code, lp[0], lp[1], lp[2], u, a, l, c);
      if (GRAD) S.Rs[off * DN_RP + r] = valid ? c : 0.0;
    } else {
      double u[VEC_DMAX], gc[VEC_DMAX];
#pragma unroll
      for (int cc = 0; cc < VEC_DMAX; ++cc) { u[cc] = cc < D ? S.Rs[(off + cc) * DN_RP + r] : 0.0; gc[cc] = 0.0; }
      l = vec_loss<GRAD>(code, lp, u, D, valid ? a : 1.0, gc);
      if (GRAD) {
#pragma unroll
        for (int cc = 0; cc < VEC_DMAX; ++cc) if (cc < D) S.Rs[(off + cc) * DN_RP + r] = valid ? gc[cc] : 0.0;
      }
    }
    l = valid ? l : 0.0;
    rowsum += l;
    if (COLSUM) {
      double cs = l;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(FULLMASK, cs, o);
      if ((t & 31) == 0) S.red[warp * DN_TN + p] = cs;
    }
  }
  return rowsum;
}

// G[i][r] += sum_jj Ys[i][jj] * Rs[jj][r]; thread (ti = t%16, tr = t/16) owns i = ti + 16 q (q < KT), rows tr*8..+7
template <int KT>
__device__ __forceinline__ void dense_gemm_gx(const double* __restrict__ Ys, const double* __restrict__ Rs, int ncols, double (&G)[KT][8]) {
  const int t = threadIdx.x, ti = t & 15, tr = t >> 4;
  const double* rp = Rs + tr * 8;
  const double* yp = Ys + ti * DN_YP;
#pragma unroll 2
  for (int jj = 0; jj < ncols; ++jj) {
    const double2 r01 = *reinterpret_cast<const double2*>(rp + jj * DN_RP);
    const double2 r23 = *reinterpret_cast<const double2*>(rp + jj * DN_RP + 2);
    const double2 r45 = *reinterpret_cast<const double2*>(rp + jj * DN_RP + 4);
    const double2 r67 = *reinterpret_cast<const double2*>(rp + jj * DN_RP + 6);
    const double rr[8] = {r01.x, r01.y, r23.x, r23.y, r45.x, r45.y, r67.x, r67.y};
#pragma unroll
    for (int q = 0; q < KT; ++q) {
      const double y = yp[q * 16 * DN_YP + jj];
#pragma unroll
      for (int a = 0; a < 8; ++a) G[q][a] = fma(y, rr[a], G[q][a]);
    }
  }
}

// ---- X sweep ------------------------------------------------------------------------------------------------------------
// rowv slots: 0 obj_old, 1 obj_new (trial), 2 reg of the trial point, 3 alpha, 4 recorded objective, 5 partial
template <int KT, int TG, int TR, int LOSS>
__global__ void __launch_bounds__(DN_THREADS, 1) dense_x_kernel(const DenseArgs P) {
  extern __shared__ __align__(16) unsigned char dn_smem[];
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  const DenseSmem S = dense_carve(dn_smem, P.k, KT);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k = P.k;
  const int nchunks = *P.nchunks;
  const int64_t ntiles = (P.row1 - P.row0 + DN_TM - 1) / DN_TM;
  double* Gg = P.gscratch + (int64_t)blockIdx.x * DN_TM * P.stride;
  const double l1 = (double)(P.n + 1);                                   // proxgrad.jl:134: length(observed_features[e]) + 1
  double* objold = S.rowv, *objnew = S.rowv + 64, *regnew = S.rowv + 128, *alpha = S.rowv + 192, *objrec = S.rowv + 256, *part = S.rowv + 320;
  int chunk_loaded = -1;
  constexpr int NGW = 32 / TG;
  const int lg = lane % TG, gq = lane / TG;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t e0 = P.row0 + tile * DN_TM;
    const int nrows = (int)((P.row1 - e0) < DN_TM ? (P.row1 - e0) : DN_TM);
    __syncthreads();
    dense_load_x(P, S.Xs, e0, nrows);
    dense_load_x(P, S.Xn, e0, nrows);              // rows that are not searching keep a finite point in the trial tile
    double G[KT][8];
#pragma unroll
    for (int q = 0; q < KT; ++q)
#pragma unroll
      for (int a = 0; a < 8; ++a) G[q][a] = 0.0;
    double rowsum = 0.0;
    // ---- gradient pass (proxgrad.jl:119-135) ----
    for (int c = 0; c < nchunks; ++c) {
      int ncols;
      if (chunk_loaded != c) { ncols = dense_load_chunk<KT>(P, S, c); chunk_loaded = c; }
      else { __syncthreads(); ncols = S.s_foff[P.chunk_ptr[c + 1] - P.chunk_ptr[c]]; }
      const int nf = P.chunk_ptr[c + 1] - P.chunk_ptr[c];
      dense_gemm_u(S.Xs, S.Ys, S.Rs, k);
      __syncthreads();
      rowsum += dense_elementwise<LOSS, true, false>(P, S, nf, e0, nrows);
      __syncthreads();
      dense_gemm_gx<KT>(S.Ys, S.Rs, ncols, G);
    }
    // gradient -> scratch [r][i] (read back lane-group-wise when trial points are formed)
    {
      const int ti = t & 15, tr = t >> 4;
#pragma unroll
      for (int q = 0; q < KT; ++q) {
        const int i = ti + 16 * q;
        if (i < P.stride) {
#pragma unroll
          for (int a = 0; a < 8; ++a) Gg[(int64_t)(tr * 8 + a) * P.stride + i] = i < k ? G[q][a] : 0.0;
        }
      }
    }
    if (t >= 64) part[t - 64] = rowsum;
    __syncthreads();
    if (t < 64) part[t] = rowsum + part[t];                               // loss of row t over all features (fixed order)
    __syncthreads();
    // regularizer of the current rows + line-search state: a lane group per row
    for (int step = 0; step < 16 / NGW; ++step) {
      const int r = warp * 16 + step * NGW + gq;
      const int64_t e = e0 + (r < nrows ? r : 0);
      const int rcode = P.reg_code[P.reg_uniform ? 0 : e];
      const double* rp = P.reg_param + (P.reg_uniform ? 0 : e) * GLRMB200_REG_NPARAM;
      double2 x[TR];
#pragma unroll
      for (int rr = 0; rr < TR; ++rr) {
        const int i0 = 2 * (lg + TG * rr);
        x[rr].x = i0 < k ? S.Xs[i0 * DN_RP + r] : 0.0;
        x[rr].y = i0 + 1 < k ? S.Xs[(i0 + 1) * DN_RP + r] : 0.0;
      }
      const double rv = (P.flags & FLAG_NO_REG) ? 0.0 : reg_eval<TG, TR>(rcode, rp, x, lg, k);
      if (lg == 0) {
        const double a0 = r < nrows ? P.alpha[e] : 0.0;
        objold[r] = part[r] + rv;
        objrec[r] = part[r] + rv;
        alpha[r] = a0;
        S.s_state[r] = (r < nrows && !(P.flags & FLAG_EVAL_ONLY) && a0 > P.min_stepsize) ? 0 : 1;
      }
    }
    __syncthreads();
    int ntrials = 0;
    // ---- line search (proxgrad.jl:136-155): all searching rows of the tile try their step together ----
    while (true) {
      const int any = __syncthreads_or(t < 64 && S.s_state[t] == 0);
      if (!any) break;
      for (int step = 0; step < 16 / NGW; ++step) {
        // every lane group runs the step (the shuffles inside reg_prox / reg_eval are warp-wide); only searching rows store
        const int r = warp * 16 + step * NGW + gq;
        const bool searching = S.s_state[r] == 0;
        const int64_t e = e0 + (r < nrows ? r : 0);
        const int rcode = P.reg_code[P.reg_uniform ? 0 : e];
        const double* rp = P.reg_param + (P.reg_uniform ? 0 : e) * GLRMB200_REG_NPARAM;
        const double stepsize = alpha[r] / l1;                            // :137
        double2 xn[TR];
#pragma unroll
        for (int rr = 0; rr < TR; ++rr) {
          const int i0 = 2 * (lg + TG * rr);
          const double x0 = i0 < k ? S.Xs[i0 * DN_RP + r] : 0.0, x1 = i0 + 1 < k ? S.Xs[(i0 + 1) * DN_RP + r] : 0.0;
          const double2 g = *reinterpret_cast<const double2*>(Gg + (int64_t)r * P.stride + i0);
          xn[rr].x = fma(-stepsize, g.x, x0); xn[rr].y = fma(-stepsize, g.y, x1);          // :140
        }
        reg_prox<TG, TR>(rcode, rp, xn, lg, k, stepsize);                  // :142
        const double rv = reg_eval<TG, TR>(rcode, rp, xn, lg, k);
        if (searching) {
#pragma unroll
          for (int rr = 0; rr < TR; ++rr) {
            const int i0 = 2 * (lg + TG * rr);
            if (i0 < k) S.Xn[i0 * DN_RP + r] = xn[rr].x;
            if (i0 + 1 < k) S.Xn[(i0 + 1) * DN_RP + r] = xn[rr].y;
          }
          if (lg == 0) regnew[r] = rv;
        }
      }
      __syncthreads();
      double trialsum = 0.0;
      for (int c = 0; c < nchunks; ++c) {
        if (chunk_loaded != c) { dense_load_chunk<KT>(P, S, c); chunk_loaded = c; }
        const int nf = P.chunk_ptr[c + 1] - P.chunk_ptr[c];
        dense_gemm_u(S.Xn, S.Ys, S.Rs, k);
        __syncthreads();
        trialsum += dense_elementwise<LOSS, false, false>(P, S, nf, e0, nrows);
        __syncthreads();
      }
      if (t >= 64) part[t - 64] = trialsum;
      __syncthreads();
      if (t < 64 && S.s_state[t] == 0) {
        const double on = (trialsum + part[t]) + regnew[t];
        objnew[t] = on;
        ++ntrials;
        if (on < objold[t]) {                                              // :143 (strict; NaN rejects)
          S.s_state[t] = 2;                                                // accepted: written back below
          alpha[t] *= 1.05;                                                // :145
          objrec[t] = on;
        } else {
          alpha[t] *= .7;                                                  // :149
          if (alpha[t] < P.min_stepsize) { alpha[t] = P.min_stepsize * 1.1; S.s_state[t] = 1; }   // :150-153
        }
      }
      __syncthreads();
      // accepted rows: the trial point becomes the row of X (:144)
      for (int idx = t; idx < DN_TM * k; idx += DN_THREADS) {
        const int r = idx / k, i = idx - r * k;
        if (S.s_state[r] == 2) P.X[(e0 + r) * P.stride + i] = S.Xn[i * DN_RP + r];
      }
      __syncthreads();
      if (t < 64 && S.s_state[t] == 2) S.s_state[t] = 1;
    }
    if (t < 64 && t < nrows) {
      if (!(P.flags & FLAG_EVAL_ONLY)) P.alpha[e0 + t] = alpha[t];
      if (P.obj_out) P.obj_out[e0 + t] = objrec[t];
    }
    if (ntrials && P.trial_counter) atomicAdd(P.trial_counter, (unsigned long long)ntrials);
  }
}

// ---- Y sweep: one pass for (row block, chunk) ------------------------------------------------------------------------------
// MODE 0: gradient pass — partial G_Y (k x columns of the chunk) and partial loss sums per feature.
// MODE 1: losses only (trial blocks / objective evaluation).
template <int KT, int LOSS, int MODE>
__global__ void __launch_bounds__(DN_THREADS, 1) dense_y_pass_kernel(const DenseArgs P) {
  extern __shared__ __align__(16) unsigned char dn_smem[];
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  const int c = blockIdx.y;
  if (c >= *P.nchunks) return;
  const DenseSmem S = dense_carve(dn_smem, P.k, KT);
  const int t = threadIdx.x;
  const int k = P.k;
  const int b = blockIdx.x;
  const int64_t rb0 = P.row0 + (int64_t)b * P.rows_per_block;
  const int64_t rb1 = (rb0 + P.rows_per_block) < P.row1 ? (rb0 + P.rows_per_block) : P.row1;
  const int ncols = dense_load_chunk<KT>(P, S, c);
  const int nf = P.chunk_ptr[c + 1] - P.chunk_ptr[c];
  // accumulators: thread (ti = t%16, tj = t/16) owns i = ti + 16 q, columns tj + 8 b2
  constexpr int CJ = DN_TN / 8;
  double GY[MODE == 0 ? KT : 1][MODE == 0 ? CJ : 1];
  if (MODE == 0) {
#pragma unroll
    for (int q = 0; q < KT; ++q)
#pragma unroll
      for (int b2 = 0; b2 < CJ; ++b2) GY[q][b2] = 0.0;
  }
  double* colacc = S.rowv;                 // [DN_TN] running per-feature sums (first 64 slots of rowv are enough: nf <= 64)
  if (t < DN_TN) colacc[t] = 0.0;
  for (int64_t e0 = rb0; e0 < rb1; e0 += DN_TM) {
    const int nrows = (int)((rb1 - e0) < DN_TM ? (rb1 - e0) : DN_TM);
    __syncthreads();
    dense_load_x(P, S.Xs, e0, nrows);
    __syncthreads();
    dense_gemm_u(S.Xs, S.Ys, S.Rs, k);
    __syncthreads();
    dense_elementwise<LOSS, MODE == 0, true>(P, S, nf, e0, nrows);
    __syncthreads();
    // per-feature sums of the tile, in a fixed order: rows 0-31 + rows 32-63 (feature p was handled by half p % 2)
    if (t < nf) {
      const int h = t & 1;
      colacc[t] += S.red[(2 * h) * DN_TN + t] + S.red[(2 * h + 1) * DN_TN + t];
    }
    if (MODE == 0) {
      // G_Y[i][jj] += sum_r Xs[i][r] * Rs[jj][r], two rows per step (16-byte loads of both operands)
      const int ti = t & 15, tj = t >> 4;
#pragma unroll 2
      for (int r = 0; r < DN_TM; r += 2) {
        double2 x[KT];
#pragma unroll
        for (int q = 0; q < KT; ++q) {
          const int i = ti + 16 * q;
          x[q] = i < k ? *reinterpret_cast<const double2*>(S.Xs + i * DN_RP + r) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int b2 = 0; b2 < CJ; ++b2) {
          const double2 rv = *reinterpret_cast<const double2*>(S.Rs + (tj + 8 * b2) * DN_RP + r);
#pragma unroll
          for (int q = 0; q < KT; ++q) GY[q][b2] = fma(x[q].y, rv.y, fma(x[q].x, rv.x, GY[q][b2]));
        }
      }
    }
  }
  __syncthreads();
  if (t < nf) P.objpart[(int64_t)b * P.n + S.s_feat[t]] = colacc[t];
  if (MODE == 0) {
    const int ti = t & 15, tj = t >> 4;
    double* gp = P.gpart + (int64_t)b * ((int64_t)P.ystart[P.n] * P.stride);
#pragma unroll
    for (int b2 = 0; b2 < CJ; ++b2) {
      const int jj = tj + 8 * b2;
      const int col = jj < ncols ? S.s_col[jj] : -1;
      if (col < 0) continue;
#pragma unroll
      for (int q = 0; q < KT; ++q) {
        const int i = ti + 16 * q;
        if (i < P.stride) gp[(int64_t)col * P.stride + i] = i < k ? GY[q][b2] : 0.0;
      }
    }
  }
}

// ---- Y sweep: small kernels -----------------------------------------------------------------------------------------------
// out[x] = sum over row blocks (fixed order) of part[b][x]; `feat_only`: only features of the current plan matter, but
// summing everything is cheap (n_blocks * n doubles)
__global__ void dense_reduce_kernel(const double* __restrict__ part, int32_t n_blocks, int64_t len, double* __restrict__ out,
                                    const int32_t* nactive) {
  if (nactive != nullptr && *nactive == 0) return;
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += part[(int64_t)b * len + x];
  out[x] = s;
}

// plan of the features still searching: compacted list, chunks of <= DN_TN columns made of whole features
__global__ void dense_y_plan_kernel(DenseYState Q) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int nact = 0, nch = 0, used = 0;
  Q.chunk_ptr[0] = 0;
  for (int64_t f = 0; f < Q.n; ++f) {
    if (!Q.active[f]) continue;
    const int D = (int)(Q.ystart[f + 1] - Q.ystart[f]);
    if (used + D > DN_TN) { ++nch; Q.chunk_ptr[nch] = nact; used = 0; }
    Q.feat_list[nact] = (int32_t)f;
    Q.feat_off[nact] = used;
    used += D;
    ++nact;
  }
  if (nact > 0) { ++nch; Q.chunk_ptr[nch] = nact; }
  *Q.nchunks = nch;
  *Q.nactive = nact;
  Q.h_nactive[0] = nact;               // the host stops enqueuing line-search rounds once it reads (0, this sweep's number)
  __threadfence_system();
  Q.h_nactive[1] = Q.seq;
  __threadfence_system();
}

// after the gradient pass: obj_old = loss + ry(y_f), search state (proxgrad.jl:177-179); one lane group per feature column
template <int TG, int TR>
__global__ void __launch_bounds__(128) dense_y_begin_kernel(DenseYState Q) {
  const int lane = threadIdx.x & 31, lg = lane % TG;
  const int64_t f = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * (32 / TG) + lane / TG;
  const bool ok = f < Q.n;
  const int64_t ff = ok ? f : 0;
  const int rcode = Q.reg_code[Q.reg_uniform ? 0 : ff];
  const double* rp = Q.reg_param + (Q.reg_uniform ? 0 : ff) * GLRMB200_REG_NPARAM;
  const int64_t y0 = Q.ystart[ff];
  const int D = (int)(Q.ystart[ff + 1] - y0);
  double rv = 0.0;
  int Dw = D;                                           // warp-uniform trip count: the shuffles in reg_eval are warp-wide
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) Dw = max(Dw, __shfl_xor_sync(FULLMASK, Dw, o));
  for (int c = 0; c < Dw; ++c) {                        // element-wise regularizers decompose over the block's columns
    const int cc = c < D ? c : 0;
    double2 v[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) v[r] = *reinterpret_cast<const double2*>(Q.Y + (y0 + cc) * Q.stride + 2 * (lg + TG * r));
    const double rc1 = reg_eval<TG, TR>(rcode, rp, v, lg, Q.k);
    if (c < D) rv += rc1;
  }
  if (Q.flags & FLAG_NO_REG) rv = 0.0;
  if (ok && lg == 0) {
    const double o = Q.colobj[f] + rv;
    Q.objold[f] = o;
    Q.obj_out[f] = o;
    Q.active[f] = (!(Q.flags & FLAG_EVAL_ONLY) && Q.alpha[f] > Q.min_stepsize) ? 1 : 0;
  }
}

// trial blocks of the features still searching: Ynew_f = prox(y_f - (alpha/l) G_f)   (proxgrad.jl:180-185)
template <int TG, int TR>
__global__ void __launch_bounds__(128) dense_y_step_kernel(DenseYState Q) {
  const int nact = *Q.nactive;
  const int lane = threadIdx.x & 31, lg = lane % TG;
  const int64_t p = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * (32 / TG) + lane / TG;
  if (nact == 0) return;
  const bool ok = p < nact;
  const int64_t f = Q.feat_list[ok ? p : 0];
  const int rcode = Q.reg_code[Q.reg_uniform ? 0 : f];
  const double* rp = Q.reg_param + (Q.reg_uniform ? 0 : f) * GLRMB200_REG_NPARAM;
  const int64_t y0 = Q.ystart[f];
  const int D = (int)(Q.ystart[f + 1] - y0);
  const double stepsize = Q.alpha[f] / (double)(Q.m + 1);                 // :179-180: l = length(observed_examples[f]) + 1
  double rv = 0.0;
  int Dw = D;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) Dw = max(Dw, __shfl_xor_sync(FULLMASK, Dw, o));
  for (int c = 0; c < Dw; ++c) {
    const int cc = c < D ? c : 0;
    double2 v[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int64_t o = (y0 + cc) * Q.stride + 2 * (lg + TG * r);
      const double2 y = *reinterpret_cast<const double2*>(Q.Y + o), g = *reinterpret_cast<const double2*>(Q.G + o);
      v[r].x = fma(-stepsize, g.x, y.x); v[r].y = fma(-stepsize, g.y, y.y);                  // :183
    }
    reg_prox<TG, TR>(rcode, rp, v, lg, Q.k, stepsize);                    // :185
    const double rc1 = reg_eval<TG, TR>(rcode, rp, v, lg, Q.k);
    if (c < D) {
      rv += rc1;
      if (ok) {
#pragma unroll
        for (int r = 0; r < TR; ++r) *reinterpret_cast<double2*>(Q.Ynew + (y0 + c) * Q.stride + 2 * (lg + TG * r)) = v[r];
      }
    }
  }
  if (ok && lg == 0) Q.regnew[f] = rv;
}

// accept / reject per feature (proxgrad.jl:186-199)
__global__ void dense_y_decide_kernel(DenseYState Q) {
  const int nact = *Q.nactive;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nact) return;
  const int64_t f = Q.feat_list[p];
  const double on = Q.colobj[f] + Q.regnew[f];
  double a = Q.alpha[f];
  if (on < Q.objold[f]) {                                                  // :186
    const int64_t y0 = Q.ystart[f], y1 = Q.ystart[f + 1];
    for (int64_t o = y0 * Q.stride; o < y1 * Q.stride; ++o) Q.Y[o] = Q.Ynew[o];   // :187
    a *= 1.05;                                                             // :188
    Q.obj_out[f] = on;                                                     // :190
    Q.active[f] = 0;
  } else {
    a *= .7;                                                               // :192
    if (a < Q.min_stepsize) { a = Q.min_stepsize * 1.1; Q.active[f] = 0; }  // :193-196
  }
  Q.alpha[f] = a;
  if (Q.trial_counter) atomicAdd(Q.trial_counter, 1ull);
}

}  // namespace glrm
