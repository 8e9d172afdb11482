// glrm_dense_mma.cuh — the fully observed path on the FP64 tensor cores (configs 1, 4, 5; proxgrad.jl:117-203 with
// observed_features == fill(1:n, m)).
//
// Same mathematics as glrm_dense.cuh (U = X'Y, R = dL/dU(U, A), G_X = Y R', G_Y = X R), organised like a fused
// attention kernel so that U and R never leave the register file:
//
//   * a CTA owns 128 rows of X (X sweep) or 128 columns of Y (Y sweep) — the OWN tile, resident in shared memory in its
//     natural layout [row][factor index] — and streams the OTHER factor through a ring of 32-row stages filled with bulk
//     copies (cp.async.bulk, SASS UBLKCP) tracked by full / empty mbarriers; the warps take the producer's duty in turn
//     (the fill of stage s + NST - 1 is issued by warp (s + NST - 1) % 8 as it starts stage s), so all 255 registers per
//     thread stay available to the accumulators;
//   * each of the 8 consumer warps owns 16 rows of the own tile.  Per stage it computes its 16 x 32 tile of U with
//     mma.sync.m8n8k4.f64 (SASS DMMA: 256 FMA per instruction, so the issue slots stay free for everything else),
//     applies the loss to the accumulator fragments IN PLACE (A arrives straight from HBM in fragment layout: the 8
//     rows x 2 columns a quad of lanes holds are whole 32-byte sectors, prefetched one stage ahead), and feeds the
//     fragments back as the A operand of the second contraction G_own += R * OTHER — the index permutation between the
//     C and A fragment layouts is absorbed by reading the rows of OTHER in the matching order (conflict-free);
//   * no __syncthreads in the streaming loop: the warps only meet at the pipeline barriers, so one warp's element-wise
//     phase overlaps the other warps' DMMAs;
//   * heterogeneous / vector-valued losses take the same GEMMs; their element-wise phase goes through a per-warp
//     16 x 32 scratch tile (features are packed into 16-column units so that a warp always holds whole features).
//
// Float64 throughout (the line search's strict `<`, proxgrad.jl:143,186): DMMA is IEEE FP64 FMA.  Reduction trees are
// fixed: a row's (feature's) loss is summed over the stage's columns inside the quad / lane in a fixed order, then over
// the stages in order, identically in the gradient pass and in the trial passes (an unchanged trial point is an exact tie
// and is rejected, as in the reference).
#pragma once
#include "glrm_dense.cuh"

namespace glrm {

struct MmSmem {
  double* own;      // [MM_TM][P]          rows of the own factor, natural layout
  double* stage;    // [nst][MM_SC][P]     ring of stages of the other factor
  double* scratch;  // [MM_WARPS][16 * MM_SW]  (heterogeneous losses only)
  double* aw;       // [MM_WARPS][16][32]  prefetched entries of A of the warp's (row, feature) pairs (heterogeneous losses only)
  double* rowv;     // [4][MM_TM]  objold, regnew, alpha, objrec
  double* psum;     // [4][MM_TM]  per-stage loss sums of the trial rows ([0] doubles as the gradient pass's row losses)
  uint64_t* bars;   // [0..3] full, [4..7] empty, [8] own tile
  int* smeta;       // [nst][MI]  per stage: [0..31] feature of local column c; generic: [32..33] features per unit,
                    //            [34..65] feature lists, [66..99] first local column of each feature (+ end)
  int* rstate;      // X sweep: s_state[128], s_perm[128];  Y sweep: feature / Y column of own column c, unit lists
  int* cnt;         // [0..7] compaction counts, [9..10] some trial point of the round needs a pass over Y (by round parity)
};
template <int NT, bool GEN>
__device__ __forceinline__ MmSmem mm_carve(unsigned char* base, int nst) {
  constexpr int P = 8 * NT + 4;
  MmSmem S;
  double* p = reinterpret_cast<double*>(base);
  S.own = p; p += MM_TM * P;
  S.stage = p; p += (size_t)nst * MM_SC * P;
  S.scratch = p; if (GEN) p += MM_WARPS * 16 * MM_SW;
  S.aw = p; if (GEN) p += MM_WARPS * 16 * 32;
  S.rowv = p; p += 4 * MM_TM;
  S.psum = p; p += 4 * MM_TM;
  S.bars = reinterpret_cast<uint64_t*>(p); p += 16;
  int* q = reinterpret_cast<int*>(p);
  S.smeta = q; q += nst * (GEN ? 136 : 32);
  S.rstate = q; q += GEN ? 544 : 256;
  S.cnt = q;
  return S;
}

__device__ __forceinline__ void mm_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mm_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dn_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mm_mbar_expect(uint64_t* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dn_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mm_bulk_copy(double* dst, const double* src, int bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dn_smem_u32(dst)), "l"(src), "r"(bytes), "r"(dn_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mm_wait(uint64_t* bar, uint32_t parity, int32_t* diag, int site) {
  if (!dn_mbar_wait(bar, parity)) dn_give_up(diag, 3, site, (int)parity, 0);
}
// D(8x8) += A(8x4) * B(4x8), FP64.  Lane (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g], D[g][2t], D[g][2t+1].
__device__ __forceinline__ void mm_dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// U[o][s] = sum_i own[o][i] * oth[s][i] for the warp's 8 MT own rows and the stage's 32 rows of the other factor.
// `own` points at own[first row + g][t], `oth` at oth[g][t]; both loads are bank-conflict free (P = 4 or 12 mod 16).
template <int NT, int MT>
__device__ __forceinline__ void mm_gemm1(const double* __restrict__ own, const double* __restrict__ oth, int ks, double (&acc)[2][4][2]) {
  constexpr int P = 8 * NT + 4;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll 2
  for (int kk = 0; kk < ks; ++kk) {
    double a[MT], b[4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) a[mt] = own[mt * 8 * P + 4 * kk];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) b[nt] = oth[nt * 8 * P + 4 * kk];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mm_dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

// G[o][i] += sum_s R[o][s] * oth[s][i] with R in the accumulator fragments of mm_gemm1.  A fragment slot t must carry
// column s_t of an 8-column block; the lane holds columns 2t and 2t+1, so the block is contracted by two DMMAs over the
// column sets {0,2,5,7} and {1,3,4,6} (rows of `oth` 4 doubles apart modulo 16: conflict-free reads).
template <int NT, int MT>
__device__ __forceinline__ void mm_gemm2(const double* __restrict__ oth, const double (&acc)[2][4][2], double (&G)[2][NT][2], int g, int t) {
  constexpr int P = 8 * NT + 4;
  const int hi = t >> 1;
#pragma unroll
  for (int sb = 0; sb < 4; ++sb) {
    const double* r1 = oth + (8 * sb + 2 * t + hi) * P + g;
    const double* r2 = oth + (8 * sb + 2 * t + 1 - hi) * P + g;
    double a1[MT], a2[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      a1[mt] = hi ? acc[mt][sb][1] : acc[mt][sb][0];
      a2[mt] = hi ? acc[mt][sb][0] : acc[mt][sb][1];
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const double b1 = r1[8 * nt], b2 = r2[8 * nt];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mm_dmma(G[mt][nt][0], G[mt][nt][1], a1[mt], b1);
        mm_dmma(G[mt][nt][0], G[mt][nt][1], a2[mt], b2);
      }
    }
  }
}

// ---- element-wise phase, uniform scalar loss: on the fragments ---------------------------------------------------------------
// mown bit mt: the own row 8 mt + g of this lane is real; moth bit 2 nt + e: the other-side index 8 nt + 2 t + e is real.
// GRAD: acc <- dL/dU (0 where not real).  lsum[mt] = the own row's loss over the stage's 32 columns (same in the 4 lanes).
template <int LOSS, bool GRAD, int MT>
__device__ __forceinline__ void mm_elem_uniform(const double* __restrict__ up, double (&acc)[2][4][2], const double (&areg)[2][4][2],
                                                unsigned mown, unsigned moth, double (&lsum)[2]) {
  const double s = up[0], p1 = up[1], p2 = up[2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double p = 0.0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool valid = ((mown >> mt) & 1u) && ((moth >> (2 * nt + e)) & 1u);
        double l, c;
        loss_eval<LOSS, GRAD>(LOSS, s, p1, p2, acc[mt][nt][e], areg[mt][nt][e], l, c);
        if (GRAD) acc[mt][nt][e] = valid ? c : 0.0;
        p += valid ? l : 0.0;
      }
    p += __shfl_xor_sync(FULLMASK, p, 1);
    p += __shfl_xor_sync(FULLMASK, p, 2);
    lsum[mt] = p;
  }
}

// ---- element-wise phase, heterogeneous problem, stage (X sweep) / unit (Y sweep) made of ONE scalar loss type: on the fragments ----
// Features keep their own parameters (scale, threshold ...).  YSIDE = false: the feature belongs to the other-side index
// (column 8 nt + 2 t + e of the stage, colfeat[]); YSIDE = true: to the own row (fown[mt]).  The per-row / per-feature sums
// follow the butterfly order of the scratch path (offsets 16, 8, 4, 2, 1 over the stage index), so a feature's loss is
// the same number whichever path its unit takes (the Y sweep's trial passes regroup the features).
template <int L, bool GRAD, int MT, bool YSIDE>
__device__ __forceinline__ void mm_elem_cols_body(const DenseArgs& P, const int* __restrict__ colfeat, const int (&fown)[2], int t,
                                                  double (&acc)[2][4][2], const double (&aP)[16], unsigned mown, unsigned moth,
                                                  double (&lsum)[2]) {
  double lv[2][4][2];
  double so[2][3];
  if (YSIDE) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const double* lp = P.loss_param + (int64_t)(fown[mt] >= 0 ? fown[mt] : 0) * GLRMB200_LOSS_NPARAM;
      so[mt][0] = lp[0]; so[mt][1] = lp[1]; so[mt][2] = lp[2];
    }
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double s = 1.0, p1 = 0.0, p2 = 0.0;
      if (!YSIDE) {
        const int f = colfeat[8 * nt + 2 * t + e];
        const double* lp = P.loss_param + (int64_t)(f >= 0 ? f : 0) * GLRMB200_LOSS_NPARAM;
        s = lp[0]; p1 = lp[1]; p2 = lp[2];
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const bool valid = ((mown >> mt) & 1u) && ((moth >> (2 * nt + e)) & 1u);
        double l, c;
        if (YSIDE) loss_eval<L, GRAD>(L, so[mt][0], so[mt][1], so[mt][2], acc[mt][nt][e], aP[mt * 8 + nt * 2 + e], l, c);
        else loss_eval<L, GRAD>(L, s, p1, p2, acc[mt][nt][e], aP[mt * 8 + nt * 2 + e], l, c);
        if (GRAD) acc[mt][nt][e] = valid ? c : 0.0;
        lv[mt][nt][e] = valid ? l : 0.0;
      }
    }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      v[e] = (lv[mt][0][e] + lv[mt][2][e]) + (lv[mt][1][e] + lv[mt][3][e]);
      v[e] += __shfl_xor_sync(FULLMASK, v[e], 2);
      v[e] += __shfl_xor_sync(FULLMASK, v[e], 1);
    }
    lsum[mt] = v[0] + v[1];
  }
}
template <bool GRAD, int MT, bool YSIDE>
__device__ __forceinline__ void mm_elem_cols(int code, const DenseArgs& P, const int* __restrict__ colfeat, const int (&fown)[2], int t,
                                             double (&acc)[2][4][2], const double (&aP)[16], unsigned mown, unsigned moth,
                                             double (&lsum)[2]) {
  switch (code) {
#define MM_CASE(L) case L: mm_elem_cols_body<L, GRAD, MT, YSIDE>(P, colfeat, fown, t, acc, aP, mown, moth, lsum); break;
    MM_CASE(GLRMB200_LOSS_QUAD) MM_CASE(GLRMB200_LOSS_L1) MM_CASE(GLRMB200_LOSS_HUBER) MM_CASE(GLRMB200_LOSS_QUANTILE)
    MM_CASE(GLRMB200_LOSS_PERIODIC) MM_CASE(GLRMB200_LOSS_POISSON) MM_CASE(GLRMB200_LOSS_ORDINAL_HINGE)
    MM_CASE(GLRMB200_LOSS_LOGISTIC) MM_CASE(GLRMB200_LOSS_WEIGHTED_HINGE)
#undef MM_CASE
    default: lsum[0] = lsum[1] = NAN; break;
  }
}
// entries of A in fragment layout for rows arow[mt] (X sweep: the stage's columns; -1 = no row / no feature)
template <bool STREAM>
__device__ __forceinline__ void mm_prefetch_a_fx(const DenseArgs& P, const int* __restrict__ colfeat, const int64_t (&arow)[2], int t,
                                                 double (&aP)[16], unsigned& moth) {
  moth = 0;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int f = colfeat[8 * nt + 2 * t + e];
      if (f >= 0) moth |= 1u << (2 * nt + e);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        double v = 0.0;
        if (f >= 0 && arow[mt] >= 0) {
          const double* src = P.A + (int64_t)f * P.lda + arow[mt];
          v = STREAM ? __ldcs(src) : __ldg(src);
        }
        aP[mt * 8 + nt * 2 + e] = v;
      }
    }
}

// ---- element-wise phase, per-feature losses (scalar or vector-valued): through the warp's scratch tile ------------------------
// The entries of A were prefetched into registers a stage ahead (aG); they are parked in the warp's staging area so that
// the feature loops below can stay rolled (one copy of the loss code per kernel).
// X sweep: scratch rows = the warp's 16 own rows, columns = the stage's 32 columns (2 units).  Lane <-> (row = lane % 16,
// features p = lane / 16, lane / 16 + 2, ... of each unit).  Returns the stage loss of row lane % 16.
template <bool GRAD, int MT>
__device__ __forceinline__ double mm_elem_generic_x(const DenseArgs& P, double* __restrict__ Sw, double* __restrict__ Aw,
                                                    const int* __restrict__ meta, double (&acc)[2][4][2], const double (&aG)[16],
                                                    bool rowvalid, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t] = acc[mt][nt][0];
      Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t + 1] = acc[mt][nt][1];
    }
#pragma unroll
  for (int q = 0; q < 16; ++q) Aw[q * 32 + lane] = aG[q];
  __syncwarp();
  const int r = lane & 15, h = lane >> 4;
  double lsum = 0.0;
  // the (row, feature) pairs of this lane, flattened over the two units: j = 0 .. n0 + n1 - 1.  The operands of pair j + 1
  // (ids, loss descriptor, entry of A) are fetched while pair j is evaluated.
  const int nf0 = meta[32], nf1 = meta[33];
  const int n0 = nf0 > h ? (nf0 - h + 1) >> 1 : 0, n1 = nf1 > h ? (nf1 - h + 1) >> 1 : 0;
  struct Ev { int off, D, code; const double* lp; double s, p1, p2, a; };
  auto fetch = [&](int j) {
    const int uu = j >= n0 ? 1 : 0, q = uu ? j - n0 : j, p = h + 2 * q;
    const int f = meta[34 + 16 * uu + p];
    const int o = meta[66 + 17 * uu + p];
    Ev ev;
    ev.off = 16 * uu + o;
    ev.D = meta[66 + 17 * uu + p + 1] - o;
    ev.code = P.loss_code[f];
    ev.lp = P.loss_param + (int64_t)f * GLRMB200_LOSS_NPARAM;
    ev.s = ev.lp[0]; ev.p1 = ev.lp[1]; ev.p2 = ev.lp[2];
    ev.a = Aw[(8 * uu + q) * 32 + lane];
    return ev;
  };
  Ev nx;
  if (n0 + n1 > 0) nx = fetch(0);
#pragma unroll 1
  for (int j = 0; j < n0 + n1; ++j) {
    const Ev ev = nx;
    if (j + 1 < n0 + n1) nx = fetch(j + 1);
    double* up = Sw + r * MM_SW + ev.off;
    double l;
    if (ev.code < GLRMB200_LOSS_MULTINOMIAL) {
      double c;
      loss_eval<0, GRAD>(ev.code, ev.s, ev.p1, ev.p2, up[0], ev.a, l, c);
      if (GRAD) up[0] = rowvalid ? c : 0.0;
    } else {
      double u[VEC_DMAX], gc[VEC_DMAX];
#pragma unroll
      for (int cc = 0; cc < VEC_DMAX; ++cc) { u[cc] = cc < ev.D ? up[cc] : 0.0; gc[cc] = 0.0; }
      l = vec_loss<GRAD>(ev.code, ev.lp, u, ev.D, rowvalid ? ev.a : 1.0, gc);
      if (GRAD) {
#pragma unroll
        for (int cc = 0; cc < VEC_DMAX; ++cc) if (cc < ev.D) up[cc] = rowvalid ? gc[cc] : 0.0;
      }
    }
    lsum += rowvalid ? l : 0.0;
  }
  __syncwarp();
  if (GRAD) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[mt][nt][0] = Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t];
        acc[mt][nt][1] = Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t + 1];
      }
    __syncwarp();
  }
  lsum += __shfl_xor_sync(FULLMASK, lsum, 16);
  return lsum;
}
// A values of the (row lane % 16, feature) pairs mm_elem_generic_x evaluates; grow = global row of that scratch row or -1
template <bool STREAM>
__device__ __forceinline__ void mm_prefetch_a_gx(const DenseArgs& P, const int* __restrict__ meta, int64_t grow, int lane, double (&aG)[16]) {
  const int h = lane >> 4;
#pragma unroll
  for (int uu = 0; uu < 2; ++uu) {
    const int nf = meta[32 + uu];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int p = h + 2 * q;
      double v = 0.0;
      if (p < nf && grow >= 0) {
        const double* src = P.A + (int64_t)meta[34 + 16 * uu + p] * P.lda + grow;
        v = STREAM ? __ldcs(src) : __ldg(src);
      }
      aG[8 * uu + q] = v;
    }
  }
}

// Y sweep: scratch rows = the warp's 16 own columns (one unit), columns = the stage's 32 rows of X; lane <-> row.  `um` points at
// the unit's lists (um[0] = features, um[1..16] = feature ids, um[17..33] = first column of each feature + end).  The loss of
// feature p over the stage's rows is added to featloss in lane p.
template <bool GRAD>
__device__ __forceinline__ void mm_elem_generic_y(const DenseArgs& P, double* __restrict__ Sw, double* __restrict__ Aw,
                                                  const int* __restrict__ um, double (&acc)[2][4][2], const double (&aG)[16],
                                                  bool rowvalid, int lane, double& featloss) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t] = acc[mt][nt][0];
      Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t + 1] = acc[mt][nt][1];
    }
#pragma unroll
  for (int q = 0; q < 16; ++q) Aw[q * 32 + lane] = aG[q];
  __syncwarp();
  const int nf = um[0];
  struct Ev { int off, D, code; const double* lp; double s, p1, p2, a; };
  auto fetch = [&](int p) {
    const int f = um[1 + p];
    Ev ev;
    ev.off = um[17 + p];
    ev.D = um[17 + p + 1] - ev.off;
    ev.code = P.loss_code[f];
    ev.lp = P.loss_param + (int64_t)f * GLRMB200_LOSS_NPARAM;
    ev.s = ev.lp[0]; ev.p1 = ev.lp[1]; ev.p2 = ev.lp[2];
    ev.a = Aw[p * 32 + lane];
    return ev;
  };
  Ev nx;
  if (nf > 0) nx = fetch(0);
#pragma unroll 1
  for (int p = 0; p < nf; ++p) {
    const Ev ev = nx;
    if (p + 1 < nf) nx = fetch(p + 1);
    double* up = Sw + ev.off * MM_SW + lane;
    double l;
    if (ev.code < GLRMB200_LOSS_MULTINOMIAL) {
      double c;
      loss_eval<0, GRAD>(ev.code, ev.s, ev.p1, ev.p2, up[0], ev.a, l, c);
      if (GRAD) up[0] = rowvalid ? c : 0.0;
    } else {
      double u[VEC_DMAX], gc[VEC_DMAX];
#pragma unroll
      for (int cc = 0; cc < VEC_DMAX; ++cc) { u[cc] = cc < ev.D ? up[cc * MM_SW] : 0.0; gc[cc] = 0.0; }
      l = vec_loss<GRAD>(ev.code, ev.lp, u, ev.D, rowvalid ? ev.a : 1.0, gc);
      if (GRAD) {
#pragma unroll
        for (int cc = 0; cc < VEC_DMAX; ++cc) if (cc < ev.D) up[cc * MM_SW] = rowvalid ? gc[cc] : 0.0;
      }
    }
    l = rowvalid ? l : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(FULLMASK, l, o);
    if (lane == p) featloss += l;
  }
  __syncwarp();
  if (GRAD) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[mt][nt][0] = Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t];
        acc[mt][nt][1] = Sw[(8 * mt + g) * MM_SW + 8 * nt + 2 * t + 1];
      }
    __syncwarp();
  }
}

// ---- producer duty (executed by one whole warp) ------------------------------------------------------------------------------------
// one stage of the X sweep: the 32 columns of Y of units 2 st and 2 st + 1 of the plan (rows of `slot`), and the stage's meta
template <int NT, bool GEN>
__device__ __forceinline__ void mm_fill_y_stage(const DenseArgs& P, const MmSmem& S, int slot, int st, int nunits, int lane) {
  constexpr int PT = 8 * NT + 4;
  int* meta = S.smeta + slot * (GEN ? 136 : 32);
  const int u = 2 * st + (lane >> 4), cc = lane & 15;
  int f = -1, ycol = -1;
  if (u < nunits) { f = P.ucol_feat[u * MM_UNIT + cc]; ycol = P.ucol_y[u * MM_UNIT + cc]; }
  meta[lane] = f;
  if (GEN) {
    const int p0 = u < nunits ? P.chunk_ptr[u] : 0, nf = u < nunits ? P.chunk_ptr[u + 1] - p0 : 0;
    if (cc == 0) meta[32 + (lane >> 4)] = nf;
    if (cc < nf) {
      const int fl = P.feat_list[p0 + cc];
      meta[34 + 16 * (lane >> 4) + cc] = fl;
      meta[66 + 17 * (lane >> 4) + cc] = P.feat_off[p0 + cc];
      if (cc == nf - 1) meta[66 + 17 * (lane >> 4) + nf] = P.feat_off[p0 + cc] + (int)(P.ystart[fl + 1] - P.ystart[fl]);
    }
    if (nf == 0 && cc == 0) meta[66 + 17 * (lane >> 4)] = 0;
    // one scalar loss type in the whole stage?  (then the element-wise phase stays on the fragments)
    const int code = f >= 0 ? P.loss_code[f] : -1;
    const bool scalar1 = f < 0 || (code < GLRMB200_LOSS_MULTINOMIAL && P.ystart[f + 1] - P.ystart[f] == 1);
    const unsigned havef = __ballot_sync(FULLMASK, f >= 0);
    const int c0 = __shfl_sync(FULLMASK, code, havef ? __ffs(havef) - 1 : 0);
    const bool same = __all_sync(FULLMASK, scalar1 && (f < 0 || code == c0));
    if (lane == 0) meta[101] = (same && havef) ? c0 : 0;
  }
  double* dst = S.stage + ((size_t)slot * MM_SC + lane) * PT;
  if (ycol < 0) {
#pragma unroll 4
    for (int i = 0; i < 8 * NT; ++i) dst[i] = 0.0;
  }
  const unsigned have = __ballot_sync(FULLMASK, ycol >= 0);
  __syncwarp();
  if (lane == 0) mm_mbar_expect(S.bars + slot, __popc(have) * 64 * NT);
  __syncwarp();
  if (ycol >= 0) mm_bulk_copy(dst, P.Ymat + (int64_t)ycol * P.stride, 64 * NT, S.bars + slot);
}
// one stage of the Y sweep: rows [row0, row0 + 32) of X, zero rows past rb1
template <int NT>
__device__ __forceinline__ void mm_fill_x_stage(const DenseArgs& P, const MmSmem& S, int slot, int64_t row0, int64_t rb1, int lane) {
  constexpr int PT = 8 * NT + 4;
  const int64_t row = row0 + lane;
  double* dst = S.stage + ((size_t)slot * MM_SC + lane) * PT;
  const bool have = row < rb1;
  if (!have) {
#pragma unroll 4
    for (int i = 0; i < 8 * NT; ++i) dst[i] = 0.0;
  }
  const unsigned hm = __ballot_sync(FULLMASK, have);
  __syncwarp();
  if (lane == 0) mm_mbar_expect(S.bars + slot, __popc(hm) * 64 * NT);
  __syncwarp();
  if (have) mm_bulk_copy(dst, P.X + row * P.stride, 64 * NT, S.bars + slot);
}

// ---- X sweep --------------------------------------------------------------------------------------------------------------------
// rowv slots: 0 obj_old, 1 reg of the trial point, 2 alpha, 3 recorded objective
template <int NT, int TG, int TR, int LOSS>
__global__ void __launch_bounds__(MM_THREADS, 1) dense_mma_x_kernel(const DenseArgs P) {
  extern __shared__ __align__(128) unsigned char dn_smem[];
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  constexpr int PT = 8 * NT + 4;
  constexpr bool GEN = LOSS == 0;
  constexpr int MI = GEN ? 136 : 32;
  const int NST = P.nst;
  const MmSmem S = mm_carve<NT, GEN>(dn_smem, NST);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int k = P.k, ks = (k + 3) >> 2;
  const int nunits = *P.nchunks;
  const int nst = (nunits + 1) >> 1;                   // stages per pass over Y
  const bool resident = nst <= NST;                    // all of Y stays in shared memory: loaded once per CTA
  const int64_t ntiles = (P.row1 - P.row0 + MM_TM - 1) / MM_TM;
  uint64_t* full = S.bars, *empty = S.bars + 4, *xfull = S.bars + 8;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { dn_mbar_init(full + s, 1); dn_mbar_init(empty + s, MM_WARPS); }
    dn_mbar_init(xfull, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  double* objold = S.rowv, *regnew = S.rowv + MM_TM, *alpha = S.rowv + 2 * MM_TM, *objrec = S.rowv + 3 * MM_TM;
  double* part = S.psum;
  int* s_state = S.rstate, *s_perm = S.rstate + MM_TM;
  double* Sw = S.scratch + warp * (16 * MM_SW);
  double* Aw = S.aw + warp * (16 * 32);
  constexpr int NGW = 32 / TG;
  const int lg = lane % TG, gq = lane / TG;
  const int ow = warp * 16;
  const double l1 = (double)(P.n + 1);                                     // proxgrad.jl:134: length(observed_features[e]) + 1
  const double up[3] = {P.uparam[0], P.uparam[1], P.uparam[2]};
  uint32_t cons = 0;                                   // stages consumed so far == index of the next fill to consume
  uint32_t xph = 0;
  bool loaded = false;
  // the stage ring (streamed Y).  Fill F (counted over the whole kernel) lives in slot F % NST; it may be issued once fill
  // F - NST has been consumed by all 8 warps (empty barrier).  A pass over Y starts with NST - 1 fills issued by warp 0; the
  // fill of stage st + NST - 1 is then issued by warp (st + NST - 1) % 8 as it starts stage st.
  auto issue = [&](uint32_t base, int st) {
    const uint32_t F = base + (uint32_t)st;
    const int slot = (int)(F % (uint32_t)NST);
    mm_wait(empty + slot, ((F / (uint32_t)NST) & 1u) ^ 1u, P.diag, 10);
    mm_fill_y_stage<NT, GEN>(P, S, slot, st, nunits, lane);
  };
  auto pass_prologue = [&]() {                         // all warps, right after a barrier that ends the previous pass
    if (resident) {
      if (!loaded && warp == 0) for (int st = 0; st < nst; ++st) mm_fill_y_stage<NT, GEN>(P, S, st, st, nunits, lane);
      loaded = true;
    } else if (warp == 0) {
      for (int st = 0; st < NST - 1 && st < nst; ++st) issue(cons, st);
    }
  };
  auto duty = [&](uint32_t base, int st) {
    const int ft = st + NST - 1;
    if (!resident && ft < nst && (ft & (MM_WARPS - 1)) == warp) issue(base, ft);
  };
  auto slot_at = [&](uint32_t base, int st) { return resident ? st : (int)((base + (uint32_t)st) % (uint32_t)NST); };
  auto par_at = [&](uint32_t base, int st) { return resident ? 0u : (((base + (uint32_t)st) / (uint32_t)NST) & 1u); };

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t e0 = P.row0 + tile * MM_TM;
    const int nrows = (int)((P.row1 - e0) < MM_TM ? (P.row1 - e0) : MM_TM);
    double* Gg = P.gscratch + (int64_t)blockIdx.x * MM_TM * P.stride;                 // gradient of the tile's rows
    double* Tg = P.gscratch + (int64_t)(gridDim.x + blockIdx.x) * MM_TM * P.stride;   // last evaluated trial point of each row
    mm_bar_sync(1, MM_THREADS);                        // everybody has left the previous tile
    if (tid == 0) { S.cnt[9] = 0; S.cnt[10] = 0; }
    if (warp == 0) {
      // the own tile: rows of X (zero rows past the end)
      for (int r = lane + (nrows & ~31); r < MM_TM; r += 32) {
        if (r >= nrows) {
          double* dst = S.own + (size_t)r * PT;
          for (int i = 0; i < 8 * NT; ++i) dst[i] = 0.0;
        }
      }
      __syncwarp();
      if (lane == 0) mm_mbar_expect(xfull, nrows * 64 * NT);
      __syncwarp();
      for (int r = lane; r < nrows; r += 32) mm_bulk_copy(S.own + (size_t)r * PT, P.X + (e0 + r) * P.stride, 64 * NT, xfull);
    }
    pass_prologue();
    long long pc0 = 0, pc_form = 0, pc_eval = 0, pc_acc = 0;
    if (P.phase != nullptr && tid == 0) pc0 = clock64();
    mm_wait(xfull, xph, P.diag, 1);
    xph ^= 1u;
    if (P.phase != nullptr && tid == 0) { const long long c = clock64(); atomicAdd(P.phase + 0, (unsigned long long)(c - pc0)); pc0 = c; }
    // ---- gradient pass (proxgrad.jl:119-135) ----
    {
      const uint32_t base = cons;
      double G[2][NT][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) G[mt][nt][0] = G[mt][nt][1] = 0.0;
      int64_t arow[2];
      unsigned mown = 0;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = ow + 8 * mt + g;
        arow[mt] = r < nrows ? e0 + r : -1;
        if (r < nrows) mown |= 1u << mt;
      }
      const int64_t grow = (ow + (lane & 15)) < nrows ? e0 + ow + (lane & 15) : -1;   // generic: row of scratch row lane % 16
      double areg[2][4][2];
      double aG[16];
      unsigned moth = 0;
      double rl[2] = {0.0, 0.0};
      double rlg = 0.0;
      auto prefetch = [&](const int* meta) {
        if constexpr (GEN) {
          if (meta[101]) mm_prefetch_a_fx<true>(P, meta, arow, t, aG, moth);    // one scalar loss type: fragment layout
          else mm_prefetch_a_gx<true>(P, meta, grow, lane, aG);
        } else {
          moth = 0;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int f = meta[8 * nt + 2 * t + e];
              if (f >= 0) moth |= 1u << (2 * nt + e);
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
                areg[mt][nt][e] = (f >= 0 && arow[mt] >= 0) ? __ldcs(P.A + (int64_t)f * P.lda + arow[mt]) : 0.0;
            }
        }
      };
      mm_wait(full + slot_at(base, 0), par_at(base, 0), P.diag, 2);
      prefetch(S.smeta + slot_at(base, 0) * MI);
      for (int st = 0; st < nst; ++st) {
        duty(base, st);
        const int slot = slot_at(base, st);
        const double* oth = S.stage + (size_t)slot * MM_SC * PT;
        double acc[2][4][2];
        mm_gemm1<NT, 2>(S.own + (size_t)(ow + g) * PT + t, oth + g * PT + t, ks, acc);
        if constexpr (GEN) {
          const int* meta = S.smeta + slot * MI;
          if (meta[101]) {
            const int nofeat[2] = {-1, -1};
            double ls[2];
            mm_elem_cols<true, 2, false>(meta[101], P, meta, nofeat, t, acc, aG, mown, moth, ls);
            // stage sums of rows 8 mt + g (quad lanes) -> the lane = row layout of the scratch path
            const double v0 = __shfl_sync(FULLMASK, ls[0], 4 * (lane & 7)), v1 = __shfl_sync(FULLMASK, ls[1], 4 * (lane & 7));
            rlg += (lane & 8) ? v1 : v0;
          } else {
            rlg += mm_elem_generic_x<true, 2>(P, Sw, Aw, meta, acc, aG, grow >= 0, lane);
          }
        } else {
          double ls[2];
          mm_elem_uniform<LOSS, true, 2>(up, acc, areg, mown, moth, ls);
          rl[0] += ls[0]; rl[1] += ls[1];
        }
        if (st + 1 < nst) {
          // the next stage's meta gives the addresses of its entries of A: wait for it now, the loads travel during the
          // second contraction
          const int ns = slot_at(base, st + 1);
          mm_wait(full + ns, par_at(base, st + 1), P.diag, 3);
          prefetch(S.smeta + ns * MI);
        }
        mm_gemm2<NT, 2>(oth, acc, G, g, t);
        if (!resident) { __syncwarp(); if (lane == 0) mm_mbar_arrive(empty + slot); }
        ++cons;
      }
      // gradient -> scratch [row][i] (read back lane-group-wise when trial points are formed); row losses -> part
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          *reinterpret_cast<double2*>(Gg + (int64_t)(ow + 8 * mt + g) * P.stride + 8 * nt + 2 * t) = make_double2(G[mt][nt][0], G[mt][nt][1]);
      if (GEN) { if (lane < 16) part[ow + lane] = rlg; }
      else if (t == 0) { part[ow + g] = rl[0]; part[ow + 8 + g] = rl[1]; }
    }
    mm_bar_sync(1, MM_THREADS);
    if (P.phase != nullptr && tid == 0) { const long long c = clock64(); atomicAdd(P.phase + 1, (unsigned long long)(c - pc0)); pc0 = c; }
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
        x[rr] = i0 < 8 * NT ? *reinterpret_cast<const double2*>(S.own + (size_t)r * PT + i0) : make_double2(0.0, 0.0);
      }
      const double rv = (P.flags & FLAG_NO_REG) ? 0.0 : reg_eval<TG, TR>(rcode, rp, x, lg, k);
      if (lg == 0) {
        const double a0 = r < nrows ? P.alpha[e] : 0.0;
        objold[r] = part[r] + rv;
        objrec[r] = part[r] + rv;
        alpha[r] = a0;
        s_state[r] = (r < nrows && !(P.flags & FLAG_EVAL_ONLY) && a0 > P.min_stepsize) ? 0 : 1;
      }
    }
    mm_bar_sync(1, MM_THREADS);
    // rows that search, in row order
    int na;
    {
      const bool act = tid < MM_TM && s_state[tid] == 0;
      const unsigned bal = __ballot_sync(FULLMASK, act);
      if (lane == 0) S.cnt[warp] = __popc(bal);
      mm_bar_sync(1, MM_THREADS);
      int before = 0;
      for (int w = 0; w < warp; ++w) before += S.cnt[w];
      if (act) s_perm[before + __popc(bal & ((1u << lane) - 1u))] = tid;
      na = 0;
      for (int w = 0; w < MM_WARPS; ++w) na += S.cnt[w];
      mm_bar_sync(1, MM_THREADS);
    }
    int ntrials = 0, rounds = 0;
    if (P.phase != nullptr && tid == 0) { const long long c = clock64(); atomicAdd(P.phase + 2, (unsigned long long)(c - pc0)); pc0 = c; }
    // ---- line search (proxgrad.jl:136-155): all searching rows of the tile try their step together ----
    while (na > 0) {
      if (++rounds > 4096) { dn_give_up(P.diag, 2, na, (int)tile, rounds); break; }
      // trial points x_new = prox(x - (alpha/l) g) of the active slots -> own[slot][.] (the tile's rows of X are not needed
      // any more: x comes from global memory, where it stays untouched until a trial is accepted); a lane group per slot.
      // A trial point that equals the row's previous (rejected) one bit for bit has the same objective and is rejected
      // again — the reference evaluates it and finds an exact tie (proxgrad.jl:143) — so such trials are decided here,
      // without a pass over Y: the step keeps shrinking (:149-153) until the point moves or the step is exhausted.
      // (k-means rows whose assignment is optimal spend all their ~13 trials this way.)
      {
        const bool have_prev = rounds > 1;
        if (tid == 0) S.cnt[9 + ((rounds + 1) & 1)] = 0;
        double2 x0n[TR], gn[TR], pvn[TR];
        auto fetch_slot = [&](int base) {
          const int s = base + warp * NGW + gq;
          const int r = s_perm[s < na ? s : 0];
#pragma unroll
          for (int rr = 0; rr < TR; ++rr) {
            const int i0 = 2 * (lg + TG * rr);
            x0n[rr] = *reinterpret_cast<const double2*>(P.X + (e0 + r) * P.stride + i0);          // padding past k is zero
            gn[rr] = *reinterpret_cast<const double2*>(Gg + (int64_t)r * P.stride + i0);
            pvn[rr] = have_prev ? *reinterpret_cast<const double2*>(Tg + (int64_t)r * P.stride + i0) : make_double2(0.0, 0.0);
          }
        };
        fetch_slot(0);
        for (int base = 0; base < na; base += MM_WARPS * NGW) {
          const int s = base + warp * NGW + gq;
          const bool ok = s < na;
          const int r = s_perm[ok ? s : 0];
          const int64_t e = e0 + r;
          const int rcode = P.reg_code[P.reg_uniform ? 0 : e];
          const double* rp = P.reg_param + (P.reg_uniform ? 0 : e) * GLRMB200_REG_NPARAM;
          double2 x0[TR], gg[TR], pv[TR];
#pragma unroll
          for (int rr = 0; rr < TR; ++rr) { x0[rr] = x0n[rr]; gg[rr] = gn[rr]; pv[rr] = pvn[rr]; }
          if (base + MM_WARPS * NGW < na) fetch_slot(base + MM_WARPS * NGW);
          auto make = [&](double a_, double2 (&out)[TR]) {
            const double stepsize = a_ / l1;                                // :137
#pragma unroll
            for (int rr = 0; rr < TR; ++rr) {
              out[rr].x = fma(-stepsize, gg[rr].x, x0[rr].x); out[rr].y = fma(-stepsize, gg[rr].y, x0[rr].y);   // :140
            }
            reg_prox<TG, TR>(rcode, rp, out, lg, k, stepsize);               // :142
#pragma unroll
            for (int rr = 0; rr < TR; ++rr) {
              const int i0 = 2 * (lg + TG * rr);
              if (i0 >= k) out[rr].x = 0.0;
              if (i0 + 1 >= k) out[rr].y = 0.0;
            }
          };
          double a = alpha[r];
          double2 xn[TR];
          make(a, xn);
          int extra = 0;
          bool stopped = false;
          {
            // repeats: the trial point is the row itself (objective == obj_old, computed by the same reduction tree in the
            // gradient pass) or the row's previous, rejected, trial point
            const unsigned gmask = (TG == 32 ? 0xffffffffu : ((1u << TG) - 1u)) << (gq * TG);
            if (rcode == GLRMB200_REG_UNIT_ONE_SPARSE) {
              // k-means rows (UnitOneSparseConstraint, regularizers.jl:297): the trial point is the one-hot vector at
              // argmax_i (x_i - s g_i), every component linear in the step s.  If it equals the row itself, the row's index
              // wins at s and at 0, hence at every step in between: all remaining trials are exact ties.  The step is
              // shrunk to exhaustion right here (:149-153), one trial counted per shrink.
              bool eq0 = true;
#pragma unroll
              for (int rr = 0; rr < TR; ++rr)
                eq0 = eq0 && __double_as_longlong(xn[rr].x) == __double_as_longlong(x0[rr].x) &&
                      __double_as_longlong(xn[rr].y) == __double_as_longlong(x0[rr].y);
              const unsigned bal0 = __ballot_sync(FULLMASK, eq0);
              if (ok && (bal0 & gmask) == gmask) {
                while (!stopped) {
                  ++extra;
                  a *= .7;
                  if (a < P.min_stepsize) { a = P.min_stepsize * 1.1; stopped = true; }
                  else if (!(a > P.min_stepsize)) stopped = true;            // (the while condition of :136 fails)
                }
              }
            }
            for (;;) {
              bool eq0 = true, eqp = have_prev;
#pragma unroll
              for (int rr = 0; rr < TR; ++rr) {
                eq0 = eq0 && __double_as_longlong(xn[rr].x) == __double_as_longlong(x0[rr].x) &&
                      __double_as_longlong(xn[rr].y) == __double_as_longlong(x0[rr].y);
                eqp = eqp && __double_as_longlong(xn[rr].x) == __double_as_longlong(pv[rr].x) &&
                      __double_as_longlong(xn[rr].y) == __double_as_longlong(pv[rr].y);
              }
              const unsigned bal0 = __ballot_sync(FULLMASK, eq0), balp = __ballot_sync(FULLMASK, eqp);
              const bool same = ok && !stopped && ((bal0 & gmask) == gmask || (balp & gmask) == gmask);
              if (!__any_sync(FULLMASK, same)) break;
              if (same) {
                ++extra;
                a *= .7;                                                     // :149
                if (a < P.min_stepsize) { a = P.min_stepsize * 1.1; stopped = true; }   // :150-153
                else if (!(a > P.min_stepsize)) stopped = true;              // (the while condition of :136 fails)
              }
              double2 xt[TR];
              make(a, xt);                                                   // (warp-wide: the prox shuffles)
              if (same && !stopped) {
#pragma unroll
                for (int rr = 0; rr < TR; ++rr) xn[rr] = xt[rr];
              }
            }
          }
          const double rv = reg_eval<TG, TR>(rcode, rp, xn, lg, k);
          __syncwarp();                                  // every lane of the group has read alpha[r]
          if (ok) {
            if (lg == 0) { alpha[r] = a; ntrials += extra; }
            if (stopped) {
              if (lg == 0) s_state[r] = 1;
            } else {
#pragma unroll
              for (int rr = 0; rr < TR; ++rr) {
                const int i0 = 2 * (lg + TG * rr);
                if (i0 < 8 * NT) *reinterpret_cast<double2*>(S.own + (size_t)s * PT + i0) = xn[rr];
                *reinterpret_cast<double2*>(Tg + (int64_t)r * P.stride + i0) = xn[rr];
              }
              if (lg == 0) { regnew[r] = rv; S.cnt[9 + (rounds & 1)] = 1; }
            }
          }
        }
      }
      mm_bar_sync(1, MM_THREADS);
      if (P.phase != nullptr && tid == 0) { const long long c = clock64(); pc_form += c - pc0; pc0 = c; }
      const bool any_eval = S.cnt[9 + (rounds & 1)] != 0;                  // (false: every trial of this round was a repeat)
      if (any_eval) pass_prologue();
      // losses of the trial points: (16 slots, stage) items.  Y resident: the items are dealt round-robin over the warps and
      // every item leaves its 16 stage sums in psum[stage]; streamed: warp w serves slots 16 w .. 16 w + 15 through all stages.
      if (any_eval) {
        const int nmp = (na + 15) >> 4;
        auto item = [&](int mp, const double* oth, const int* meta, double (&ls)[2], double& lsg) {
          const double* ownp = S.own + (size_t)(16 * mp + g) * PT + t;
          double acc[2][4][2];
          if constexpr (GEN) {
            const int sg = 16 * mp + (lane & 15);
            const int64_t grow = sg < na ? e0 + s_perm[sg] : -1;
            double aG[16];
            if (meta[101]) {
              int64_t arow[2];
              unsigned mown = 0, moth = 0;
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                const int s = 16 * mp + 8 * mt + g;
                arow[mt] = s < na ? e0 + s_perm[s] : -1;
                if (s < na) mown |= 1u << mt;
              }
              mm_prefetch_a_fx<false>(P, meta, arow, t, aG, moth);
              mm_gemm1<NT, 2>(ownp, oth + g * PT + t, ks, acc);
              const int nofeat[2] = {-1, -1};
              double l2[2];
              mm_elem_cols<false, 2, false>(meta[101], P, meta, nofeat, t, acc, aG, mown, moth, l2);
              const double v0 = __shfl_sync(FULLMASK, l2[0], 4 * (lane & 7)), v1 = __shfl_sync(FULLMASK, l2[1], 4 * (lane & 7));
              lsg = (lane & 8) ? v1 : v0;
            } else {
              mm_prefetch_a_gx<false>(P, meta, grow, lane, aG);
              mm_gemm1<NT, 2>(ownp, oth + g * PT + t, ks, acc);
              lsg = mm_elem_generic_x<false, 2>(P, Sw, Aw, meta, acc, aG, grow >= 0, lane);
            }
          } else {
            int64_t arow[2];
            unsigned mown = 0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const int s = 16 * mp + 8 * mt + g;
              arow[mt] = s < na ? e0 + s_perm[s] : -1;
              if (s < na) mown |= 1u << mt;
            }
            const bool two = 16 * mp + 8 < na;
            double areg[2][4][2];
            unsigned moth = 0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int f = meta[8 * nt + 2 * t + e];
                if (f >= 0) moth |= 1u << (2 * nt + e);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                  areg[mt][nt][e] = (f >= 0 && arow[mt] >= 0) ? __ldg(P.A + (int64_t)f * P.lda + arow[mt]) : 0.0;
              }
            ls[0] = ls[1] = 0.0;
            if (two) { mm_gemm1<NT, 2>(ownp, oth + g * PT + t, ks, acc); mm_elem_uniform<LOSS, false, 2>(up, acc, areg, mown, moth, ls); }
            else     { mm_gemm1<NT, 1>(ownp, oth + g * PT + t, ks, acc); mm_elem_uniform<LOSS, false, 1>(up, acc, areg, mown, moth, ls); }
          }
        };
        if (resident) {
          for (int it = warp; it < nmp * nst; it += MM_WARPS) {
            const int mp = it / nst, st = it - mp * nst;
            double ls[2] = {0.0, 0.0}, lsg = 0.0;
            item(mp, S.stage + (size_t)st * MM_SC * PT, S.smeta + st * MI, ls, lsg);
            double* ps = S.psum + st * MM_TM + 16 * mp;
            if (GEN) { if (lane < 16) ps[lane] = lsg; }
            else if (t == 0) { ps[g] = ls[0]; ps[8 + g] = ls[1]; }
          }
        } else {
          const uint32_t base = cons;
          double tl[2] = {0.0, 0.0}, tlg = 0.0;
          for (int st = 0; st < nst; ++st) {
            duty(base, st);
            const int slot = slot_at(base, st);
            mm_wait(full + slot, par_at(base, st), P.diag, 4);
            if (warp < nmp) {
              double ls[2] = {0.0, 0.0}, lsg = 0.0;
              item(warp, S.stage + (size_t)slot * MM_SC * PT, S.smeta + slot * MI, ls, lsg);
              if (GEN) tlg += lsg; else { tl[0] += ls[0]; tl[1] += ls[1]; }
            }
            __syncwarp();
            if (lane == 0) mm_mbar_arrive(empty + slot);
            ++cons;
          }
          if (warp < nmp) {
            double* ps = S.psum + 16 * warp;
            if (GEN) { if (lane < 16) ps[lane] = tlg; }
            else if (t == 0) { ps[g] = tl[0]; ps[8 + g] = tl[1]; }
          }
        }
      }
      mm_bar_sync(1, MM_THREADS);
      if (P.phase != nullptr && tid == 0) { const long long c = clock64(); pc_eval += c - pc0; pc0 = c; }
      if (tid < na && s_state[s_perm[tid]] == 0) {
        const int myrow = s_perm[tid];
        double tot = S.psum[tid];
        if (resident) for (int st = 1; st < nst; ++st) tot += S.psum[st * MM_TM + tid];
        const double on = tot + regnew[myrow];
        ++ntrials;
        if (on < objold[myrow]) {                                            // :143 (strict; NaN rejects)
          s_state[myrow] = 2;                                                // accepted: written back below
          alpha[myrow] *= 1.05;                                              // :145
          objrec[myrow] = on;
        } else {
          alpha[myrow] *= .7;                                                // :149
          if (alpha[myrow] < P.min_stepsize) { alpha[myrow] = P.min_stepsize * 1.1; s_state[myrow] = 1; }   // :150-153
          else if (!(alpha[myrow] > P.min_stepsize)) s_state[myrow] = 1;     // (the while condition of :136 fails)
        }
      }
      mm_bar_sync(1, MM_THREADS);
      // accepted rows: the trial point becomes the row of X (:144)
      for (int idx = tid; idx < na * k; idx += MM_THREADS) {
        const int s = idx / k, i = idx - s * k;
        const int r = s_perm[s];
        if (s_state[r] == 2) P.X[(e0 + r) * P.stride + i] = S.own[(size_t)s * PT + i];
      }
      // compact the rows still searching (order kept)
      {
        const int r = (tid < MM_TM && tid < na) ? s_perm[tid] : -1;
        const bool act = r >= 0 && s_state[r] == 0;
        const unsigned bal = __ballot_sync(FULLMASK, act);
        mm_bar_sync(1, MM_THREADS);                                          // every read of the old s_perm is done
        if (lane == 0) S.cnt[warp] = __popc(bal);
        if (r >= 0 && s_state[r] == 2) s_state[r] = 1;
        mm_bar_sync(1, MM_THREADS);
        int before = 0;
        for (int w = 0; w < warp; ++w) before += S.cnt[w];
        if (act) s_perm[before + __popc(bal & ((1u << lane) - 1u))] = r;
        na = 0;
        for (int w = 0; w < MM_WARPS; ++w) na += S.cnt[w];
        mm_bar_sync(1, MM_THREADS);
      }
      if (P.phase != nullptr && tid == 0) { const long long c = clock64(); pc_acc += c - pc0; pc0 = c; }
    }
    if (P.phase != nullptr && tid == 0) {
      atomicAdd(P.phase + 3, (unsigned long long)pc_form); atomicAdd(P.phase + 4, (unsigned long long)pc_eval);
      atomicAdd(P.phase + 5, (unsigned long long)pc_acc); atomicAdd(P.phase + 6, (unsigned long long)rounds); atomicAdd(P.phase + 7, 1ull);
    }
    if (tid < nrows) {
      if (!(P.flags & FLAG_EVAL_ONLY)) P.alpha[e0 + tid] = alpha[tid];
      if (P.obj_out) P.obj_out[e0 + tid] = objrec[tid];
    }
    if (ntrials && P.trial_counter) atomicAdd(P.trial_counter, (unsigned long long)ntrials);
  }
}

// ---- Y sweep: one pass for (row block, 128 columns of the plan) -------------------------------------------------------------------
// MODE 0: gradient pass — partial G_Y (k x columns) and partial loss sums per feature.  MODE 1: losses only.
template <int NT, int LOSS, int MODE>
__global__ void __launch_bounds__(MM_THREADS, 1) dense_mma_y_kernel(const DenseArgs P) {
  extern __shared__ __align__(128) unsigned char dn_smem[];
  if (P.stop != nullptr && *reinterpret_cast<const volatile int*>(P.stop) != 0) return;
  constexpr int PT = 8 * NT + 4;
  constexpr bool GEN = LOSS == 0;
  const int nunits = *P.nchunks;
  const int u0 = (int)blockIdx.y * MM_WARPS;
  if (u0 >= nunits) return;
  const int NST = P.nst;
  const MmSmem S = mm_carve<NT, GEN>(dn_smem, NST);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int k = P.k, ks = (k + 3) >> 2;
  const int b = (int)blockIdx.x + P.block0;              // global row block (a rank launches only the blocks it owns)
  const int64_t rb0 = (int64_t)b * P.rows_per_block;
  const int64_t rb1 = (rb0 + P.rows_per_block) < P.row1 ? (rb0 + P.rows_per_block) : P.row1;
  const int nstg = rb1 > rb0 ? (int)((rb1 - rb0 + MM_SC - 1) / MM_SC) : 0;
  uint64_t* full = S.bars, *empty = S.bars + 4, *xfull = S.bars + 8;
  int* ownfeat = S.rstate, *owny = S.rstate + MM_TM, *ulist = S.rstate + 2 * MM_TM;    // ulist: [unit][34] (generic)
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { dn_mbar_init(full + s, 1); dn_mbar_init(empty + s, MM_WARPS); }
    dn_mbar_init(xfull, 1);
  }
  for (int c = tid; c < MM_TM; c += MM_THREADS) {
    const int u = u0 + (c >> 4);
    ownfeat[c] = u < nunits ? P.ucol_feat[u * MM_UNIT + (c & 15)] : -1;
    owny[c] = u < nunits ? P.ucol_y[u * MM_UNIT + (c & 15)] : -1;
  }
  if (GEN) {
    for (int x = tid; x < MM_WARPS * 16; x += MM_THREADS) {
      const int uu = x >> 4, p = x & 15, u = u0 + uu;
      int* um = ulist + uu * 34;
      const int p0 = u < nunits ? P.chunk_ptr[u] : 0, nf = u < nunits ? P.chunk_ptr[u + 1] - p0 : 0;
      if (p == 0) { um[0] = nf; if (nf == 0) um[17] = 0; }
      if (p < nf) {
        const int f = P.feat_list[p0 + p];
        um[1 + p] = f;
        um[17 + p] = P.feat_off[p0 + p];
        if (p == nf - 1) um[17 + nf] = P.feat_off[p0 + p] + (int)(P.ystart[f + 1] - P.ystart[f]);
      }
    }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  auto issue = [&](int s) {                              // stage s: rows rb0 + 32 s .. of X (one whole warp)
    const int slot = s % NST;
    mm_wait(empty + slot, (((uint32_t)s / (uint32_t)NST) & 1u) ^ 1u, P.diag, 11);
    mm_fill_x_stage<NT>(P, S, slot, rb0 + (int64_t)s * MM_SC, rb1, lane);
  };
  if (warp == 0) {
    // the CTA's columns of Y (zero rows for unused columns), then the first stages
    for (int c = lane; c < MM_TM; c += 32) {
      if (owny[c] < 0) {
        double* dst = S.own + (size_t)c * PT;
        for (int i = 0; i < 8 * NT; ++i) dst[i] = 0.0;
      }
    }
    int ncopy = 0;
    for (int c = lane; c < MM_TM; c += 32) ncopy += owny[c] >= 0 ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ncopy += __shfl_xor_sync(FULLMASK, ncopy, o);
    __syncwarp();
    if (lane == 0) mm_mbar_expect(xfull, ncopy * 64 * NT);
    __syncwarp();
    for (int c = lane; c < MM_TM; c += 32)
      if (owny[c] >= 0) mm_bulk_copy(S.own + (size_t)c * PT, P.Ymat + (int64_t)owny[c] * P.stride, 64 * NT, xfull);
    for (int s = 0; s < NST - 1 && s < nstg; ++s) issue(s);
  }

  // warp w owns unit u0 + w = columns 16 w .. 16 w + 15 of the CTA
  const int ow = warp * 16;
  const bool active = u0 + warp < nunits;
  const double up[3] = {P.uparam[0], P.uparam[1], P.uparam[2]};
  double* Sw = S.scratch + warp * (16 * MM_SW);
  double* Aw = S.aw + warp * (16 * 32);
  const int* um = ulist + warp * 34;
  int fown[2];
  unsigned mown = 0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    fown[mt] = ownfeat[ow + 8 * mt + g];
    if (fown[mt] >= 0) mown |= 1u << mt;
  }
  double G[2][MODE == 0 ? NT : 1][2];
  if constexpr (MODE == 0) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) G[mt][nt][0] = G[mt][nt][1] = 0.0;
  }
  double cl[2] = {0.0, 0.0};
  double featloss = 0.0;
  double areg[2][4][2];
  double aG[16];
  unsigned moth = 0;
  // heterogeneous problem: a unit made of one scalar loss type keeps its element-wise phase on the fragments
  int ocode = 0;
  if (GEN && active) {
    const int nf = um[0];
    const int code = lane < nf ? P.loss_code[um[1 + lane]] : -1;
    const int c0 = __shfl_sync(FULLMASK, code, 0);
    const bool same = __all_sync(FULLMASK, lane >= nf || code == c0);
    ocode = (same && nf > 0 && um[17 + nf] == nf && c0 < GLRMB200_LOSS_MULTINOMIAL) ? c0 : 0;
  }
  auto prefetch = [&](int s) {
    const int64_t r0 = rb0 + (int64_t)s * MM_SC;
    if constexpr (GEN) {
      if (ocode) {
        moth = 0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t row = r0 + 8 * nt + 2 * t;
          if (row < rb1) moth |= 1u << (2 * nt);
          if (row + 1 < rb1) moth |= 1u << (2 * nt + 1);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            double2 v = make_double2(0.0, 0.0);
            if (fown[mt] >= 0 && row < rb1) v = __ldcs(reinterpret_cast<const double2*>(P.A + (int64_t)fown[mt] * P.lda + row));
            aG[mt * 8 + nt * 2] = v.x; aG[mt * 8 + nt * 2 + 1] = v.y;
          }
        }
      } else {
        const int nf = um[0];
        const int64_t row = r0 + lane;
#pragma unroll
        for (int p = 0; p < 16; ++p) aG[p] = (p < nf && row < rb1) ? __ldcs(P.A + (int64_t)um[1 + p] * P.lda + row) : 0.0;
      }
    } else {
      moth = 0;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int64_t row = r0 + 8 * nt + 2 * t;
        if (row < rb1) moth |= 1u << (2 * nt);
        if (row + 1 < rb1) moth |= 1u << (2 * nt + 1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          double2 v = make_double2(0.0, 0.0);
          if (fown[mt] >= 0 && row < rb1) v = __ldcs(reinterpret_cast<const double2*>(P.A + (int64_t)fown[mt] * P.lda + row));
          areg[mt][nt][0] = v.x; areg[mt][nt][1] = v.y;
        }
      }
    }
  };
  mm_wait(xfull, 0, P.diag, 5);
  if (active && nstg > 0) prefetch(0);
  for (int s = 0; s < nstg; ++s) {
    {
      const int ft = s + NST - 1;
      if (ft < nstg && (ft & (MM_WARPS - 1)) == warp) issue(ft);
    }
    const int slot = s % NST;
    mm_wait(full + slot, ((uint32_t)s / (uint32_t)NST) & 1u, P.diag, 6);
    if (active) {
      const double* oth = S.stage + (size_t)slot * MM_SC * PT;
      double acc[2][4][2];
      mm_gemm1<NT, 2>(S.own + (size_t)(ow + g) * PT + t, oth + g * PT + t, ks, acc);
      if constexpr (GEN) {
        if (ocode) {
          double ls[2];
          mm_elem_cols<MODE == 0, 2, true>(ocode, P, nullptr, fown, t, acc, aG, mown, moth, ls);
          cl[0] += ls[0]; cl[1] += ls[1];
        } else {
          mm_elem_generic_y<MODE == 0>(P, Sw, Aw, um, acc, aG, rb0 + (int64_t)s * MM_SC + lane < rb1, lane, featloss);
        }
      } else {
        double ls[2];
        mm_elem_uniform<LOSS, MODE == 0, 2>(up, acc, areg, mown, moth, ls);
        cl[0] += ls[0]; cl[1] += ls[1];
      }
      if (s + 1 < nstg) prefetch(s + 1);
      if constexpr (MODE == 0) mm_gemm2<NT, 2>(oth, acc, G, g, t);
    }
    __syncwarp();
    if (lane == 0) mm_mbar_arrive(empty + slot);
  }
  if (!active) return;
  // partial loss sums of the block's rows per feature, partial G_Y
  if (GEN && !ocode) {
    if (lane < um[0]) P.objpart[(int64_t)b * P.n + um[1 + lane]] = featloss;
  } else if (t == 0) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) if (fown[mt] >= 0) P.objpart[(int64_t)b * P.n + fown[mt]] = cl[mt];
  }
  if constexpr (MODE == 0) {
    double* gp = P.gpart + (int64_t)b * ((int64_t)P.ystart[P.n] * P.stride);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int col = owny[ow + 8 * mt + g];
      if (col < 0) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        *reinterpret_cast<double2*>(gp + (int64_t)col * P.stride + 8 * nt + 2 * t) = make_double2(G[mt][nt][0], G[mt][nt][1]);
    }
  }
}

}  // namespace glrm
