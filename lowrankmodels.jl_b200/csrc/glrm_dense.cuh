// glrm_dense.cuh — the fully observed path (glrm.observed_features == fill(1:n, m), src/glrm.jl:33-34): configs 1, 4, 5.
//
// With every entry observed the sweeps of proxgrad.jl:117-203 are dense contractions, not gathers:
//     U = X'Y                 (m x d, inner k)        what the reference's gemm! computes (proxgrad.jl:66,157,202)
//     R = dL/dU(U, A)         element-wise (per feature; vector-valued losses see the feature's block of U)
//     G_X = Y R'              (k x m, inner d)        gradient of every row          (proxgrad.jl:122-132)
//     G_Y = X R               (k x d, inner m)        gradient of every column block (proxgrad.jl:165-175)
// These kernels stream A exactly once per pass, straight from the column-major array Julia hands over (device copy with
// the leading dimension padded to the 64-row tile; no index lists, no transposed copy), and never materialise U, R or a
// dense G_X.  A CTA owns a tile of 64 rows, keeps the tile of X, a chunk of <= 64 columns of Y and the 64 x 64 tile of
// U / R in shared memory, and runs the two contractions as register-tiled FP64 FMA loops (8 x 4 accumulators per thread).
//
// The 64 x 64 tile of A belonging to a step lands in shared memory through the TMA unit: one 512-byte bulk copy
// (cp.async.bulk.shared::cluster.global, SASS UBLKCP) per feature column, completion counted on an mbarrier, issued one
// step ahead of its use (two tile buffers when shared memory allows, else one buffer refilled as soon as the element-wise
// phase has consumed it) — the loads of A are never on the critical path of the FMA loops.
//
// Float64 throughout: the line search's strict `<` (proxgrad.jl:143,186) needs the same arithmetic as the reference; the
// roof is the FP64 pipe (36.6 TFLOP/s measured), not tensor cores — B200's FP64 tensor rate equals its FP64 FMA rate, and
// an error-compensated bf16 split needs >= 28 partial products to carry 53 bits, i.e. less than the FMA pipe delivers
// directly.
//
//   dense_x_kernel      the whole X sweep for a tile: gradient pass over all chunks, then the per-row backtracking line
//                       search.  Rows still searching are compacted to the front of the trial tile after every round and
//                       dealt round-robin over the thread rows, so a round costs ~ceil(active/16)/4 of a full pass.
//   dense_y_pass_kernel one pass of the Y sweep for (row block, chunk): partial G_Y and partial per-feature objectives
//                       (MODE 0), or objectives only for a list of features evaluated at trial blocks (MODE 1).
//   dense_reduce_kernel / dense_y_begin_kernel / dense_y_step_kernel / dense_y_decide_kernel / dense_y_plan_kernel
//                       fixed-order reduction over the row blocks, trial blocks prox(y - (alpha/l) g), accept / reject
//                       per feature (proxgrad.jl:179-200), and the compacted list of features still searching.
// Rows of a tile and features of a chunk keep fixed positions in every reduction, so results do not depend on the grid.
#pragma once
#include "glrm_device.cuh"
#include "glrm_vec.cuh"
#include "glrm_dense_host.h"

namespace glrm {

struct DenseSmem {
  double* As;    // [nbuf][DN_TN][DN_TM]  tiles of A, [feature of the chunk][row]  (bulk-copy destination)
  double* Ys;    // [KT*16][DN_YP]
  double* Xs;    // [k][DN_RP]   rows of X (gradient pass) / trial points of the rows still searching (line search)
  double* Rs;    // [DN_TN][DN_RP]  U, then R, [column][row position]
  double* rowv;  // [6][64] per-row scalars
  double* red;   // [4][DN_TN] per-warp partials
  uint64_t* bar; // [2] mbarriers of the A tile buffers
  int* s_col;    // [2][DN_TN] global Y column of local column jj (-1: unused); two slots: the next chunk's map is built
  int* s_feat;   // [2][DN_TN] feature of chunk position p                      while the current one is still in use
  int* s_foff;   // [2][DN_TN + 2] first local column of chunk position p
  int* s_state;  // [64]
  int* s_perm;   // [64] slot -> row of the tile (rows still searching, in row order)
  int* s_cnt;    // [4]
};
// every region starts on a 16-byte boundary (sizes below are multiples of 16 bytes)
__host__ __device__ inline size_t dense_ys_doubles(int kt) { return ((size_t)kt * 16 * DN_YP + 1) & ~(size_t)1; }
__host__ __device__ inline size_t dense_smem_bytes(int k, int kt, int nbuf) {
  return (size_t)nbuf * DN_TN * DN_TM * sizeof(double) +
         (dense_ys_doubles(kt) + (size_t)k * DN_RP + (size_t)DN_TN * DN_RP + 6 * 64 + 4 * DN_TN) * sizeof(double) +
         2 * sizeof(uint64_t) + (2 * DN_TN + 2 * DN_TN + 2 * (DN_TN + 2) + 64 + 64 + 4) * sizeof(int);
}
// `base` must be the extern __shared__ array itself: plain pointer arithmetic on it keeps the address space known to the
// compiler (LDS / STS instead of generic loads)
__device__ __forceinline__ DenseSmem dense_carve(unsigned char* base, int k, int kt, int nbuf) {
  DenseSmem S;
  double* p = reinterpret_cast<double*>(base);
  S.As = p; p += (size_t)nbuf * DN_TN * DN_TM;
  S.Ys = p; p += dense_ys_doubles(kt);
  S.Xs = p; p += (size_t)k * DN_RP;
  S.Rs = p; p += (size_t)DN_TN * DN_RP;
  S.rowv = p; p += 6 * 64;
  S.red = p; p += 4 * DN_TN;
  S.bar = reinterpret_cast<uint64_t*>(p); p += 2;
  int* q = reinterpret_cast<int*>(p);
  S.s_col = q; q += 2 * DN_TN;
  S.s_feat = q; q += 2 * DN_TN;
  S.s_foff = q; q += 2 * (DN_TN + 2);
  S.s_state = q; q += 64;
  S.s_perm = q; q += 64;
  S.s_cnt = q;
  return S;
}

// ---- TMA-fed tiles of A -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dn_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dn_smem_u32(bar)), "r"(count) : "memory");
}
// Watchdog: a wait that outlasts any plausible copy (2^32 clocks, ~2 s) records where it happened and lets the kernel run
// on (the results are then meaningless; the engine reports the record as an error instead of hanging the device).
static __device__ __noinline__ void dn_give_up(int32_t* diag, int code, int a, int b, int c) {
  if (diag != nullptr && atomicCAS(diag, 0, code) == 0) {
    diag[1] = (int)blockIdx.x; diag[2] = (int)blockIdx.y; diag[3] = (int)threadIdx.x;
    diag[4] = a; diag[5] = b; diag[6] = c;
    __threadfence();
  }
}
__device__ __forceinline__ bool dn_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = dn_smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > (1ll << 32)) return false;
  }
}

// State of the tile buffers; every thread of the CTA carries the same copy (all control flow around it is CTA-uniform).
struct APipe {
  int nbuf;
  bool pend[2];
  uint32_t par[2];
  int64_t held_e0[2];
  int held_c[2];
  __device__ __forceinline__ void init(int nb, int32_t* dg) {
    nbuf = nb; diag = dg;
    pend[0] = pend[1] = false; par[0] = par[1] = 0; nfetch[0] = nfetch[1] = 0;
    held_e0[0] = held_e0[1] = -1; held_c[0] = held_c[1] = -1;
  }
  int32_t* diag;
  int nfetch[2];
  __device__ __forceinline__ void wait(const DenseSmem& S, int b, int site = 0) {
    if (pend[b]) {
      if (!dn_mbar_wait(S.bar + b, par[b]))
        dn_give_up(diag, 1, b | (site << 4) | ((int)par[b] << 8) | (nfetch[b] << 12), held_c[b], (int)(held_e0[b] >> 6));
      par[b] ^= 1u; pend[b] = false;
    }
  }
  // rows [e0, e0 + 64) of the features of chunk c -> buffer b.  Called by all threads of the CTA (it may synchronise them).
  // The caller guarantees (by a __syncthreads since the last element-wise phase that read buffer b) that nobody still
  // reads it.
  __device__ __forceinline__ void fetch(const DenseArgs& P, const DenseSmem& S, int b, int64_t e0, int c) {
    if (held_e0[b] == e0 && held_c[b] == c) return;
    if (pend[b]) {
      // a copy nobody consumed (the speculative "same tile again" of a line search that has ended) is drained first.  The
      // barrier must not be re-armed before EVERY thread has seen that phase complete: a warp that re-arms early lets the
      // new phase complete as well, and a thread still waiting for the old parity would then wait on a phase that only
      // the next fetch can start — forever.
      wait(S, b, 7);
      __syncthreads();
    }
    ++nfetch[b];
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int p0 = P.chunk_ptr[c], nf = P.chunk_ptr[c + 1] - p0;
      const uint32_t bar = dn_smem_u32(S.bar + b);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nf * DN_TM * 8) : "memory");
      __syncwarp();
      double* dst = S.As + (size_t)b * DN_TN * DN_TM;
      for (int p = lane; p < nf; p += 32) {
        const int f = P.feat_list[p0 + p];
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dn_smem_u32(dst + p * DN_TM)), "l"(P.A + (int64_t)f * P.lda + e0), "r"(DN_TM * 8), "r"(bar) : "memory");
      }
    }
    pend[b] = true;
    held_e0[b] = e0; held_c[b] = c;
  }
};

// ---- asynchronous staging of the factor tiles ------------------------------------------------------------------------------
// Both tiles are stored transposed ([factor index][column / row]) so the FMA loops read them conflict-free; an 8-byte
// cp.async (LDGSTS) per element does the transposition on the way in, with all copies of a tile in flight at once.
// src_bytes = 0 zero-fills the destination (rows past the end, factor rows past k).
__device__ __forceinline__ void dn_cp_async8(double* dst, const double* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dn_smem_u32(dst)), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void dn_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// every copy this thread issued has landed, and (barrier) so has everybody else's
__device__ __forceinline__ void dn_async_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
}

// chunk set-up into meta slot `ms`: feature list and local column map of chunk c.  Nobody may still read slot `ms`.
__device__ __forceinline__ void dense_chunk_meta(const DenseArgs& P, const DenseSmem& S, int c, int ms) {
  const int t = threadIdx.x;
  const int p0 = P.chunk_ptr[c], nf = P.chunk_ptr[c + 1] - p0;
  int* s_col = S.s_col + ms * DN_TN;
  int* s_feat = S.s_feat + ms * DN_TN;
  int* s_foff = S.s_foff + ms * (DN_TN + 2);
  for (int jj = t; jj < DN_TN; jj += DN_THREADS) s_col[jj] = -1;
  __syncthreads();
  for (int p = t; p < nf; p += DN_THREADS) {
    const int f = P.feat_list[p0 + p], off = P.feat_off[p0 + p];
    const int64_t y0 = P.ystart[f];
    const int D = (int)(P.ystart[f + 1] - y0);
    s_feat[p] = f;
    s_foff[p] = off;
    if (p == nf - 1) s_foff[nf] = off + D;
    for (int cc = 0; cc < D; ++cc) s_col[off + cc] = (int)(y0 + cc);
  }
  if (nf == 0 && t == 0) s_foff[0] = 0;
  __syncthreads();
}
// the chunk of Y described by meta slot `ms` ([i][jj], rows >= k zero) is put in flight.  The caller guarantees that nobody
// reads Ys any more, and calls dn_async_wait before using it.  A warp per column, lanes over i: 256-byte coalesced reads,
// conflict-free transposed writes (pitch 65).
template <int KT>
__device__ __forceinline__ void dense_chunk_copy(const DenseArgs& P, const DenseSmem& S, int ms) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int* s_col = S.s_col + ms * DN_TN;
  for (int jj = warp; jj < DN_TN; jj += DN_THREADS / 32) {
    const int col = s_col[jj];
    const double* src = P.Ymat + (int64_t)(col >= 0 ? col : 0) * P.stride;
#pragma unroll
    for (int i = lane; i < KT * 16; i += 32) dn_cp_async8(S.Ys + i * DN_YP + jj, src + (i < P.k ? i : 0), col >= 0 && i < P.k);
  }
  dn_async_commit();
}
template <int KT>
__device__ __forceinline__ void dense_chunk_begin(const DenseArgs& P, const DenseSmem& S, int c, int ms) {
  dense_chunk_meta(P, S, c, ms);
  dense_chunk_copy<KT>(P, S, ms);
}

// tile of X -> Xs[i][r] (zero rows past the end): a warp per row, lanes over i (coalesced)
__device__ __forceinline__ void dense_x_begin(const DenseArgs& P, double* Xs, int64_t e0, int nrows) {
  const int k = P.k, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < DN_TM; r += DN_THREADS / 32) {
    const double* src = P.X + (e0 + (r < nrows ? r : 0)) * P.stride;
    for (int i = lane; i < k; i += 32) dn_cp_async8(Xs + i * DN_RP + r, src + i, r < nrows);
  }
  dn_async_commit();
}

// U[r][jj] = sum_i Xt[i][r] * Ys[i][jj]; thread (trow = t/16, tcol = t%16) owns row positions trow*8 .. trow*8+NR-1,
// columns tcol + 16 q.  The result goes to Rs[jj][r] (conflict-free 16-byte stores per column).  NR < 8 is the line
// search with few rows left: slot s sits at position (s % 8) * 8 + s / 8, so the first 8 NR slots are the first NR
// positions of every thread row.
template <int NR>
__device__ __forceinline__ void dense_gemm_u(const double* __restrict__ Xt, const double* __restrict__ Ys, double* __restrict__ Rs, int k) {
  const int t = threadIdx.x, trow = t >> 4, tcol = t & 15;
  double acc[NR][DN_CT];
#pragma unroll
  for (int a = 0; a < NR; ++a)
#pragma unroll
    for (int b = 0; b < DN_CT; ++b) acc[a][b] = 0.0;
  const double* xp = Xt + trow * 8;
  const double* yp = Ys + tcol;
  // few rows per thread = few independent FMA chains: unroll deeper so that the shared-memory loads of several steps are in
  // flight together (one warp per scheduler: nothing else hides their latency)
#pragma unroll (NR <= 2 ? 8 : (NR <= 4 ? 4 : 2))
  for (int i = 0; i < k; ++i) {
    double x[NR];
#pragma unroll
    for (int a = 0; a < NR; a += 2) {
      const double2 v = *reinterpret_cast<const double2*>(xp + i * DN_RP + a);
      x[a] = v.x; x[a + 1] = v.y;
    }
    double y[DN_CT];
#pragma unroll
    for (int b = 0; b < DN_CT; ++b) y[b] = yp[i * DN_YP + 16 * b];
#pragma unroll
    for (int a = 0; a < NR; ++a)
#pragma unroll
      for (int b = 0; b < DN_CT; ++b) acc[a][b] = fma(x[a], y[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < DN_CT; ++b) {
    double* dst = Rs + (tcol + 16 * b) * DN_RP + trow * 8;
#pragma unroll
    for (int a = 0; a < NR; a += 2) *reinterpret_cast<double2*>(dst + a) = make_double2(acc[a][b], acc[a + 1][b]);
  }
}
__device__ __forceinline__ void dense_gemm_u_rows(int nr, const double* Xt, const double* Ys, double* Rs, int k) {
  if (nr <= 2) dense_gemm_u<2>(Xt, Ys, Rs, k);
  else if (nr <= 4) dense_gemm_u<4>(Xt, Ys, Rs, k);
  else if (nr <= 6) dense_gemm_u<6>(Xt, Ys, Rs, k);
  else dense_gemm_u<8>(Xt, Ys, Rs, k);
}

// Element-wise phase on the U tile in shared memory: thread <-> (row position r = t % 64, features p = t/64, t/64 + 2, ...).
// `arow` is the tile row whose entries of A this position holds (== r outside the compacted line search), `valid` whether
// the position carries a real row.  GRAD: U is replaced by dL/dU in place.  Returns this thread's share of the row's loss;
// COLSUM additionally reduces every feature's loss over the 64 rows into red[warp][p] (fixed shuffle tree).
template <int LOSS, bool GRAD, bool COLSUM>
__device__ __forceinline__ double dense_elementwise(const DenseArgs& P, const DenseSmem& S, int ms, const double* __restrict__ At, int nf,
                                                    int arow, bool valid) {
  const int t = threadIdx.x, r = t & 63, half = t >> 6, warp = t >> 5;
  const int* s_feat = S.s_feat + ms * DN_TN;
  const int* s_foff = S.s_foff + ms * (DN_TN + 2);
  double rowsum = 0.0;
  if (!COLSUM && !GRAD && !valid) return 0.0;               // compacted line search: positions past the active rows idle
  for (int p = half; p < nf; p += 2) {
    const int f = s_feat[p], off = s_foff[p], D = s_foff[p + 1] - off;
    const double a = At[p * DN_TM + arow];
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

// slot <-> position of the compacted line search (an involution: swaps the two octal digits)
__device__ __forceinline__ int dense_slot_pos(int s) { return ((s & 7) << 3) | (s >> 3); }

// ---- X sweep ------------------------------------------------------------------------------------------------------------
// rowv slots: 0 obj_old, 1 partial sums, 2 reg of the trial point, 3 alpha, 4 recorded objective
template <int KT, int TG, int TR, int LOSS>
__global__ void __launch_bounds__(DN_THREADS, 1) dense_x_kernel(const DenseArgs P) {
  extern __shared__ __align__(128) unsigned char dn_smem[];
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  const DenseSmem S = dense_carve(dn_smem, P.k, KT, P.nbuf);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k = P.k;
  const int nchunks = *P.nchunks;
  const int64_t ntiles = (P.row1 - P.row0 + DN_TM - 1) / DN_TM;
  double* Gg = P.gscratch + (int64_t)blockIdx.x * DN_TM * P.stride;
  const double l1 = (double)(P.n + 1);                                   // proxgrad.jl:134: length(observed_features[e]) + 1
  double* objold = S.rowv, *part = S.rowv + 64, *regnew = S.rowv + 128, *alpha = S.rowv + 192, *objrec = S.rowv + 256;
  constexpr int NGW = 32 / TG;
  const int lg = lane % TG, gq = lane / TG;
  if (t == 0) { dn_mbar_init(S.bar, 1); dn_mbar_init(S.bar + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  APipe ap;
  ap.init(P.nbuf, P.diag);
  unsigned stepno = 0;
  // buffer of a step: with one or two chunks the tiles of A stay put for the whole tile (no refetch in the line search)
  auto bufof = [&](int c) { return ap.nbuf == 1 ? 0 : (nchunks <= 2 ? c : (int)(stepno & 1u)); };
  auto bufnext = [&](int c) { return ap.nbuf == 1 ? 0 : (nchunks <= 2 ? c : (int)((stepno + 1u) & 1u)); };
  // the chunk of Y in shared memory (or on its way there) and the meta slot that describes it.  With one or two chunks
  // each chunk keeps its own meta slot for the whole kernel (built once); otherwise the slots alternate.
  int y_chunk = -1, y_ms = 0;
  int meta_of[2] = {-1, -1};
  auto want_chunk = [&](int c) {           // Ys and (if it has to be rebuilt) the target meta slot must be free
    if (y_chunk == c) return;
    const int slot = nchunks <= 2 ? c : (y_ms ^ 1);
    if (meta_of[slot] != c) { dense_chunk_meta(P, S, c, slot); meta_of[slot] = c; }
    y_ms = slot;
    dense_chunk_copy<KT>(P, S, slot);
    y_chunk = c;
  };

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t e0 = P.row0 + tile * DN_TM;
    const int nrows = (int)((P.row1 - e0) < DN_TM ? (P.row1 - e0) : DN_TM);
    __syncthreads();                                   // everybody is done with the previous tile's shared memory
    ap.fetch(P, S, bufof(0), e0, 0);
    dense_x_begin(P, S.Xs, e0, nrows);
    want_chunk(0);
    double G[KT][8];
#pragma unroll
    for (int q = 0; q < KT; ++q)
#pragma unroll
      for (int a = 0; a < 8; ++a) G[q][a] = 0.0;
    double rowsum = 0.0;
    // ---- gradient pass (proxgrad.jl:119-135) ----
    for (int c = 0; c < nchunks; ++c) {
      const int b = bufof(c);
      if (y_chunk != c) { __syncthreads(); want_chunk(c); }      // Ys is busy until the previous step's G update is over
      const int nf = P.chunk_ptr[c + 1] - P.chunk_ptr[c];
      ap.fetch(P, S, b, e0, c);                        // normally a no-op: issued one step ahead
      dn_async_wait();                                 // the chunk of Y (and, first step, the tile of X) has landed
      const int ncols = S.s_foff[y_ms * (DN_TN + 2) + nf];
      dense_gemm_u<8>(S.Xs, S.Ys, S.Rs, k);
      __syncthreads();
      const int cn = c + 1 < nchunks ? c + 1 : 0;      // next step: the next chunk, or chunk 0 again (first trial round)
      if (ap.nbuf == 2) ap.fetch(P, S, bufnext(cn), e0, cn);
      ap.wait(S, b, 1);
      rowsum += dense_elementwise<LOSS, true, false>(P, S, y_ms, S.As + (size_t)b * DN_TN * DN_TM, nf, t & 63, (t & 63) < nrows);
      __syncthreads();
      if (ap.nbuf == 1) ap.fetch(P, S, 0, e0, cn);
      dense_gemm_gx<KT>(S.Ys, S.Rs, ncols, G);
      ++stepno;
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
    __syncthreads();                                   // (also: the last G update has finished reading Ys)
    if (nchunks > 1) want_chunk(0);                    // chunk 0 comes back while the line search is set up
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
    // rows that search, in row order
    int na;
    {
      const bool act = t < 64 && S.s_state[t] == 0;
      const unsigned bal = __ballot_sync(FULLMASK, act);
      if (lane == 0) S.s_cnt[warp] = __popc(bal);
      __syncthreads();
      if (act) S.s_perm[(warp ? S.s_cnt[0] : 0) + __popc(bal & ((1u << lane) - 1u))] = t;
      na = S.s_cnt[0] + S.s_cnt[1];
      __syncthreads();
    }
    int ntrials = 0, rounds = 0;
    // ---- line search (proxgrad.jl:136-155): all searching rows of the tile try their step together ----
    // (a step size shrinks by 0.7 per rejected trial, so a search ends within ~log(stepsize / min_stepsize) / log(1 / 0.7)
    // rounds; the cap only trips on a corrupted state, and is reported through the watchdog record)
    while (na > 0) {
      if (++rounds > 4096) { dn_give_up(P.diag, 2, na, (int)tile, rounds); break; }
      // trial points x_new = prox(x - (alpha/l) g) of the active slots -> Xs[.][position of the slot]; a lane group per slot.
      // The row and gradient of the next slot are fetched (L2) while the current one goes through prox / evaluate.
      {
        double2 x0n[TR], gn[TR];
        auto fetch_slot = [&](int base) {
          const int s = base + warp * NGW + gq;
          const int r = S.s_perm[s < na ? s : 0];
#pragma unroll
          for (int rr = 0; rr < TR; ++rr) {
            const int i0 = 2 * (lg + TG * rr);
            x0n[rr] = *reinterpret_cast<const double2*>(P.X + (e0 + r) * P.stride + i0);        // padding past k is zero
            gn[rr] = *reinterpret_cast<const double2*>(Gg + (int64_t)r * P.stride + i0);
          }
        };
        fetch_slot(0);
        for (int base = 0; base < na; base += 4 * NGW) {
          const int s = base + warp * NGW + gq;
          const bool ok = s < na;
          const int r = S.s_perm[ok ? s : 0];
          const int64_t e = e0 + r;
          const int rcode = P.reg_code[P.reg_uniform ? 0 : e];
          const double* rp = P.reg_param + (P.reg_uniform ? 0 : e) * GLRMB200_REG_NPARAM;
          const double stepsize = alpha[r] / l1;                            // :137
          double2 xn[TR];
#pragma unroll
          for (int rr = 0; rr < TR; ++rr) {
            xn[rr].x = fma(-stepsize, gn[rr].x, x0n[rr].x); xn[rr].y = fma(-stepsize, gn[rr].y, x0n[rr].y);   // :140
          }
          if (base + 4 * NGW < na) fetch_slot(base + 4 * NGW);
          reg_prox<TG, TR>(rcode, rp, xn, lg, k, stepsize);                  // :142
          const double rv = reg_eval<TG, TR>(rcode, rp, xn, lg, k);
          if (ok) {
            const int pos = dense_slot_pos(s);
#pragma unroll
            for (int rr = 0; rr < TR; ++rr) {
              const int i0 = 2 * (lg + TG * rr);
              if (i0 < k) S.Xs[i0 * DN_RP + pos] = xn[rr].x;
              if (i0 + 1 < k) S.Xs[(i0 + 1) * DN_RP + pos] = xn[rr].y;
            }
            if (lg == 0) regnew[r] = rv;
          }
        }
      }
      const int myslot = dense_slot_pos(t & 63);                           // the slot whose position this thread serves
      const bool mine = myslot < na;
      const int myrow = S.s_perm[mine ? myslot : 0];
      const int nr = (na + 7) >> 3;                                        // row positions per thread row that are in use
      double trialsum = 0.0;
      for (int c = 0; c < nchunks; ++c) {
        const int b = bufof(c);
        want_chunk(c);                                 // normally on its way already
        const int nf = P.chunk_ptr[c + 1] - P.chunk_ptr[c];
        ap.fetch(P, S, b, e0, c);
        dn_async_wait();                               // chunk of Y landed; trial points written (first chunk)
        dense_gemm_u_rows(nr, S.Xs, S.Ys, S.Rs, k);
        __syncthreads();
        const int cn = c + 1 < nchunks ? c + 1 : 0;
        const int ms = y_ms;                           // the element-wise phase below still reads this chunk's meta slot
        if (nchunks > 1) want_chunk(cn);               // Ys is free: the next chunk travels during the element-wise phase
        if (ap.nbuf == 2) ap.fetch(P, S, bufnext(cn), e0, cn);
        ap.wait(S, b, 2);
        trialsum += dense_elementwise<LOSS, false, false>(P, S, ms, S.As + (size_t)b * DN_TN * DN_TM, nf, myrow, mine);
        __syncthreads();
        if (ap.nbuf == 1) ap.fetch(P, S, 0, e0, cn);
        ++stepno;
      }
      if (t >= 64) part[t - 64] = trialsum;
      __syncthreads();
      if (t < 64 && mine) {
        const double on = (trialsum + part[t]) + regnew[myrow];
        ++ntrials;
        if (on < objold[myrow]) {                                          // :143 (strict; NaN rejects)
          S.s_state[myrow] = 2;                                            // accepted: written back below
          alpha[myrow] *= 1.05;                                            // :145
          objrec[myrow] = on;
        } else {
          alpha[myrow] *= .7;                                              // :149
          if (alpha[myrow] < P.min_stepsize) { alpha[myrow] = P.min_stepsize * 1.1; S.s_state[myrow] = 1; }   // :150-153
        }
      }
      __syncthreads();
      // accepted rows: the trial point becomes the row of X (:144)
      for (int idx = t; idx < na * k; idx += DN_THREADS) {
        const int s = idx / k, i = idx - s * k;
        const int r = S.s_perm[s];
        if (S.s_state[r] == 2) P.X[(e0 + r) * P.stride + i] = S.Xs[i * DN_RP + dense_slot_pos(s)];
      }
      // compact the rows still searching (order kept)
      {
        const int r = (t < 64 && t < na) ? S.s_perm[t] : -1;
        const bool act = r >= 0 && S.s_state[r] == 0;
        const unsigned bal = __ballot_sync(FULLMASK, act);
        __syncthreads();                                                   // every read of the old s_perm is done
        if (lane == 0) S.s_cnt[warp] = __popc(bal);
        if (r >= 0 && S.s_state[r] == 2) S.s_state[r] = 1;
        __syncthreads();
        if (act) S.s_perm[(warp ? S.s_cnt[0] : 0) + __popc(bal & ((1u << lane) - 1u))] = r;
        na = S.s_cnt[0] + S.s_cnt[1];
        __syncthreads();
      }
    }
    if (t < 64 && t < nrows) {
      if (!(P.flags & FLAG_EVAL_ONLY)) P.alpha[e0 + t] = alpha[t];
      if (P.obj_out) P.obj_out[e0 + t] = objrec[t];
    }
    if (ntrials && P.trial_counter) atomicAdd(P.trial_counter, (unsigned long long)ntrials);
  }
  dn_async_wait();                                     // nothing may be in flight when the CTA exits
  ap.wait(S, 0);
  ap.wait(S, 1);
}

// ---- Y sweep: one pass for (row block, chunk) ------------------------------------------------------------------------------
// MODE 0: gradient pass — partial G_Y (k x columns of the chunk) and partial loss sums per feature.
// MODE 1: losses only (trial blocks / objective evaluation).
template <int KT, int LOSS, int MODE>
__global__ void __launch_bounds__(DN_THREADS, 1) dense_y_pass_kernel(const DenseArgs P) {
  extern __shared__ __align__(128) unsigned char dn_smem[];
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  const int c = blockIdx.y;
  if (c >= *P.nchunks) return;
  const DenseSmem S = dense_carve(dn_smem, P.k, KT, P.nbuf);
  const int t = threadIdx.x;
  const int k = P.k;
  const int b = (int)blockIdx.x + P.block0;              // global row block (a rank launches only the blocks it owns)
  const int64_t rb0 = (int64_t)b * P.rows_per_block;
  const int64_t rb1 = (rb0 + P.rows_per_block) < P.row1 ? (rb0 + P.rows_per_block) : P.row1;
  if (t == 0) { dn_mbar_init(S.bar, 1); dn_mbar_init(S.bar + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  APipe ap;
  ap.init(P.nbuf, P.diag);
  if (rb0 < rb1) {
    ap.fetch(P, S, 0, rb0, c);
    dense_x_begin(P, S.Xs, rb0, (int)((rb1 - rb0) < DN_TM ? (rb1 - rb0) : DN_TM));
  }
  dense_chunk_begin<KT>(P, S, c, 0);
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
  int j = 0;
  for (int64_t e0 = rb0; e0 < rb1; e0 += DN_TM, ++j) {
    const int nrows = (int)((rb1 - e0) < DN_TM ? (rb1 - e0) : DN_TM);
    const int buf = ap.nbuf == 2 ? (j & 1) : 0;
    const bool more = e0 + DN_TM < rb1;
    const int nrows_next = more ? (int)((rb1 - e0 - DN_TM) < DN_TM ? (rb1 - e0 - DN_TM) : DN_TM) : 0;
    dn_async_wait();                       // this tile of X (first tile: and the chunk of Y) has landed; also the barrier
                                           // after the previous tile's G_Y update / element-wise phase
    if (ap.nbuf == 2 && more) ap.fetch(P, S, buf ^ 1, e0 + DN_TM, c);
    dense_gemm_u<8>(S.Xs, S.Ys, S.Rs, k);
    __syncthreads();
    if (MODE == 1 && more) dense_x_begin(P, S.Xs, e0 + DN_TM, nrows_next);      // Xs is free: the next tile travels now
    ap.wait(S, buf, 3);
    dense_elementwise<LOSS, MODE == 0, true>(P, S, 0, S.As + (size_t)buf * DN_TN * DN_TM, nf, t & 63, (t & 63) < nrows);
    __syncthreads();
    if (ap.nbuf == 1 && more) ap.fetch(P, S, 0, e0 + DN_TM, c);
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
      if (more) { __syncthreads(); dense_x_begin(P, S.Xs, e0 + DN_TM, nrows_next); }
    }
  }
  dn_async_wait();
  const int ncols = S.s_foff[nf];
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
  ap.wait(S, 0);
  ap.wait(S, 1);
}

// ---- Y sweep: small kernels -----------------------------------------------------------------------------------------------
// Iterations are enqueued ahead of the device (glrm_engine.cu): once the stopping rule has fired, every kernel of the
// iterations still in the queue must leave the model untouched — the small kernels check the same flag as the sweeps.
__device__ __forceinline__ bool dense_stopped(const int* stop) {
  return stop != nullptr && *reinterpret_cast<const volatile int*>(stop) != 0;
}

#ifndef GLRM_DENSE_HELPERS_ONLY   // (non-template kernels: defined once, in dense_inst.cu)
// Two-level fixed-order reduction over the row blocks.  The DN_GROUPS groups are the same whatever the number of GPUs (a
// rank owns whole groups), so the sums — and everything downstream — are bit-identical for 1, 2, 4 and 8 ranks.
// gsum[g][x] = sum over the `bg` blocks of group g (sequential), for the groups g0 <= g < g1
__global__ void dense_reduce_groups_kernel(const double* __restrict__ part, int32_t bg, int32_t g0, int32_t g1, int64_t len,
                                           double* __restrict__ gsum, const int32_t* nactive, const int* stop) {
  if (dense_stopped(stop)) return;
  if (nactive != nullptr && *nactive == 0) return;
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = g0 + (int)blockIdx.y;
  if (x >= len || g >= g1) return;
  double s = 0.0;
  for (int b = g * bg; b < (g + 1) * bg; ++b) s += part[(int64_t)b * len + x];
  gsum[(int64_t)g * len + x] = s;
}
// out[x] = sum over the DN_GROUPS groups (sequential)
__global__ void dense_reduce_total_kernel(const double* __restrict__ gsum, int64_t len, double* __restrict__ out,
                                          const int32_t* nactive, const int* stop) {
  if (dense_stopped(stop)) return;
  if (nactive != nullptr && *nactive == 0) return;
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  double s = 0.0;
#pragma unroll
  for (int g = 0; g < DN_GROUPS; ++g) s += gsum[(int64_t)g * len + x];
  out[x] = s;
}

// plan of the features still searching: compacted list, chunks of <= DN_TN columns made of whole features
__global__ void dense_y_plan_kernel(DenseYState Q) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int nact = 0, nch = 0, used = 0;
  if (!dense_stopped(Q.stop)) {
    const int tn = Q.unit_cols;
    Q.chunk_ptr[0] = 0;
    for (int64_t f = 0; f < Q.n; ++f) {
      if (!Q.active[f]) continue;
      const int64_t y0 = Q.ystart[f];
      const int D = (int)(Q.ystart[f + 1] - y0);
      if (used + D > tn) {
        if (Q.ucol_feat != nullptr) for (int c = used; c < tn; ++c) { Q.ucol_feat[nch * tn + c] = -1; Q.ucol_y[nch * tn + c] = -1; }
        ++nch; Q.chunk_ptr[nch] = nact; used = 0;
      }
      Q.feat_list[nact] = (int32_t)f;
      Q.feat_off[nact] = used;
      if (Q.ucol_feat != nullptr) for (int c = 0; c < D; ++c) { Q.ucol_feat[nch * tn + used + c] = (int32_t)f; Q.ucol_y[nch * tn + used + c] = (int32_t)(y0 + c); }
      used += D;
      ++nact;
    }
    if (nact > 0) {
      if (Q.ucol_feat != nullptr) for (int c = used; c < tn; ++c) { Q.ucol_feat[nch * tn + c] = -1; Q.ucol_y[nch * tn + c] = -1; }
      ++nch; Q.chunk_ptr[nch] = nact;
    }
  }
  *Q.nchunks = nch;
  *Q.nactive = nact;
  // published to the host in a ring of 4 (count, key) slots indexed by the round: the host reads the plan of exactly the
  // round it waits for (never a later one), so that several ranks — whose devices run at different paces — all take the
  // same decision at the same round and keep enqueuing their collectives in lockstep
  volatile int32_t* slot = Q.h_nactive + 2 * (Q.seq & 3);
  slot[0] = nact;                      // the host stops enqueuing line-search rounds once it reads (0, key of this plan)
  __threadfence_system();
  slot[1] = Q.seq;
  __threadfence_system();
}

#endif  // GLRM_DENSE_HELPERS_ONLY

// after the gradient pass: obj_old = loss + ry(y_f), search state (proxgrad.jl:177-179); one lane group per feature column
template <int TG, int TR>
__global__ void __launch_bounds__(128) dense_y_begin_kernel(DenseYState Q) {
  if (dense_stopped(Q.stop)) return;
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
  if (nact == 0 || dense_stopped(Q.stop)) return;
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

#ifndef GLRM_DENSE_HELPERS_ONLY
// accept / reject per feature (proxgrad.jl:186-199)
__global__ void dense_y_decide_kernel(DenseYState Q) {
  if (dense_stopped(Q.stop)) return;
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

#endif  // GLRM_DENSE_HELPERS_ONLY

}  // namespace glrm
