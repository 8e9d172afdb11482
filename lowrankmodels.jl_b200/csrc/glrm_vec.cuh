// glrm_vec.cuh — units that involve vector-valued losses (losses.jl:354-608: Multinomial, OvA, BvS, Ordistic,
// MultinomialOrdinal).  A column f of A with embedding dimension D_f owns the k x D_f block Y[:, y0_f .. y0_f+D_f)
// (get_yidxs, losses.jl:76-93).  Two kinds of unit need this path:
//   X side, any row of a problem that has such columns: per observed (e,f) it needs the D_f dot products
//     u_c = x_e . Y[:, y0_f+c], the vector gradient, and g += Y_f * grad            (proxgrad.jl:125-131)
//   Y side, a block column f: G_f += x_e * grad', block prox, col_objective over the block   (:168-186)
// The kernel is the straightforward form of the fused update (one warp per unit, group-per-entry, no software
// pipeline): these configurations are correctness-first this round.  D_f <= VEC_DMAX.
#pragma once
#include "glrm_device.cuh"

namespace glrm {

constexpr int VEC_DMAX = 8;

struct VecArgs {
  SweepArgs s;                 // shared fields (lists, schedule, factors, regs, alpha, flags ...)
  const int64_t* ystart;       // [n+1] first Y column of each feature
  int32_t x_side;              // 1: units are rows of X (entries carry their own loss); 0: units are block columns
};

// ---- vector losses: value and gradient (scaled), u[0..D) -------------------------------------------------------
// bin-loss dispatch for OvA / BvS (losses.jl:419,456: any scalar loss; default LogisticLoss(scale))
__device__ __forceinline__ void bin_eval(int bcode, double bs, double bp1, double bp2, double u, bool label, double& l, double& c) {
  loss_eval<0, true>(bcode, bs, bp1, bp2, u, label ? 1.0 : 0.0, l, c);
}

// MultinomialLoss with DD levels (losses.jl:369-398); one exponential per level serves the value and the gradient
// (u_j - u_a - M and u_j - max(u) are the same number up to rounding)
template <int DD, bool WANT_GRAD>
__device__ __forceinline__ double multinomial_fixed(double s, int a, const double (&u)[VEC_DMAX], double (&gc)[VEC_DMAX]) {
  double mx = u[0], ua = 0.0;
#pragma unroll
  for (int j = 0; j < DD; ++j) { mx = jl_maxd(mx, u[j]); ua = (j == a - 1) ? u[j] : ua; }
  const double M = mx - ua;
  double e[DD];
#pragma unroll
  for (int j = 0; j < DD; ++j) e[j] = exp(u[j] - ua - M);
  double sumexp = 0.0;
#pragma unroll
  for (int j = 0; j < DD; ++j) sumexp += e[j];
  if (WANT_GRAD) {
    const double inv = 1.0 / sumexp;
#pragma unroll
    for (int j = 0; j < DD; ++j) gc[j] = s * (e[j] * inv - (j == a - 1 ? 1.0 : 0.0));
  }
  return s * (log(sumexp) + M);
}

template <bool WANT_GRAD>
__device__ __forceinline__ double vec_loss(int code, const double* __restrict__ lp, double (&u)[VEC_DMAX], int D,
                                           double alab, double (&gc)[VEC_DMAX]) {
  const double s = lp[0];
  const int a = (int)alab;             // 1-based level
  double loss = 0.0;
  switch (code) {
    case GLRMB200_LOSS_MULTINOMIAL: {  // losses.jl:369-398.  evaluate: log-sum-exp shifted by max(u); the reference's
      // O(D^2) gradient loop equals softmax(u) - e_a term by term (exp(-M_j)/sumexp_j == exp(u_j-max)/sum exp(u-max)).
      // The level count is fixed at compile time per case: straight-line code, so the D exponentials overlap.
      switch (D) {
#define GLRM_MNL(DD) case DD: loss = multinomial_fixed<DD, WANT_GRAD>(s, a, u, gc); break;
        GLRM_MNL(1) GLRM_MNL(2) GLRM_MNL(3) GLRM_MNL(4) GLRM_MNL(5) GLRM_MNL(6) GLRM_MNL(7) GLRM_MNL(8)
#undef GLRM_MNL
        default: loss = NAN; break;
      }
    } break;
    case GLRMB200_LOSS_OVA:            // :424-438   (scale applied on top of the bin loss's own scale, as in the reference)
    case GLRMB200_LOSS_BVS: {          // :461-475
      const int bcode = (int)lp[3];
      const double bs = lp[4], bp1 = lp[5], bp2 = lp[6];
#pragma unroll
      for (int j = 0; j < VEC_DMAX; ++j) if (j < D) {
        double l, c;
        bin_eval(bcode, bs, bp1, bp2, u[j], code == GLRMB200_LOSS_OVA ? (a == j + 1) : (a > j + 1), l, c);
        loss += l;
        gc[j] = s * c;
      }
      loss *= s;
    } break;
    case GLRMB200_LOSS_ORDISTIC: {     // :499-519;  exp(-M_j)/invlik_j == exp(-u_j^2)/sum_jp exp(-u_jp^2) (shifted by the min)
      double ua = 0.0, mn = INFINITY;
#pragma unroll
      for (int j = 0; j < VEC_DMAX; ++j) if (j < D) { if (j == a - 1) ua = u[j]; mn = jl_mind(mn, u[j] * u[j]); }
      double M = -INFINITY;
#pragma unroll
      for (int j = 0; j < VEC_DMAX; ++j) if (j < D) M = jl_maxd(M, ua * ua - u[j] * u[j]);
      double invlik = 0.0, z = 0.0;
#pragma unroll
      for (int j = 0; j < VEC_DMAX; ++j) if (j < D) { invlik += exp(ua * ua - u[j] * u[j] - M); const double e = exp(mn - u[j] * u[j]); gc[j] = e; z += e; }
      loss = s * (M + log(invlik));
      if (WANT_GRAD) {
#pragma unroll
        for (int j = 0; j < VEC_DMAX; ++j) if (j < D) gc[j] = s * ((j == a - 1 ? 2.0 * u[j] : 0.0) - 2.0 * u[j] * gc[j] / z);
      }
    } break;
    case GLRMB200_LOSS_MULTINOMIAL_ORDINAL: {   // :572-608 (rules enforced on a copy of u)
      const int lmax = (int)lp[2];
      u[0] = jl_mind(-1e-3, u[0]);
#pragma unroll
      for (int j = 1; j < VEC_DMAX; ++j) if (j < D) u[j] = jl_mind(u[j], u[j - 1] - 1e-3);
#pragma unroll
      for (int j = 0; j < VEC_DMAX; ++j) gc[j] = 0.0;
      double ulo = 0.0, uhi = 0.0;     // u[a-2], u[a-1]
#pragma unroll
      for (int j = 0; j < VEC_DMAX; ++j) if (j < D) { if (j == a - 2) ulo = u[j]; if (j == a - 1) uhi = u[j]; }
      if (a == 1) {
        loss = -s * log(exp(0.0) - exp(u[0]));
        if (WANT_GRAD) gc[0] = -s * (-exp(u[0]) / (exp(0.0) - exp(u[0])));
      } else if (a == lmax) {
        loss = -s * ulo;
        if (WANT_GRAD) {
#pragma unroll
          for (int j = 0; j < VEC_DMAX; ++j) if (j == a - 2) gc[j] = -s;
        }
      } else {
        const double den = exp(ulo) - exp(uhi);
        loss = -s * log(den);
        if (WANT_GRAD) {
#pragma unroll
          for (int j = 0; j < VEC_DMAX; ++j) {
            if (j == a - 1) gc[j] = -s * (-exp(uhi) / den);
            if (j == a - 2) gc[j] = -s * (exp(ulo) / den);
          }
        }
      }
    } break;
    default: {                         // scalar loss (D == 1)
      double l, c;
      loss_eval<0, WANT_GRAD>(code, s, lp[1], lp[2], u[0], alab, l, c);
      loss = l;
      gc[0] = c;
    } break;
  }
  return loss;
}

// one pass over the unit's entries.  X side: own = x (R slots), each entry gathers D columns of Y.
// Y side: own = block (D x R slots), each entry gathers one column x_e.
template <int G, int R, int DB, bool GRAD>
__device__ __forceinline__ double vec_pass(const VecArgs& V, int64_t unit, int64_t start, int64_t len, int lane,
                                           const double2 (&own)[DB][R], int Dunit, double2 (&grad)[DB][R]) {
  constexpr int NGW = 32 / G;
  const SweepArgs& A = V.s;
  const int lg = lane % G, gq = lane / G;
  const int stride = A.stride;
  double obj = 0.0;
  if (GRAD) {
#pragma unroll
    for (int c = 0; c < DB; ++c)
#pragma unroll
      for (int r = 0; r < R; ++r) grad[c][r] = make_double2(0.0, 0.0);
  }
  const int64_t nsteps = (len + NGW - 1) / NGW;
  for (int64_t s = 0; s < nsteps; ++s) {
    const int64_t t = s * NGW + gq;
    const bool act = t < len;
    const int64_t q = start + (act ? t : 0);
    const int32_t j = A.idx ? A.idx[q] : (int32_t)(act ? t : 0);
    const double a = A.val[q];
    const int64_t f = V.x_side ? j : unit;                         // the feature whose loss applies
    const int code = A.loss_code[f];
    const double* lp = A.loss_param + f * GLRMB200_LOSS_NPARAM;
    const int64_t y0 = V.ystart[f];
    const int D = act ? (int)(V.ystart[f + 1] - y0) : 0;
    int Dw = D;                                                    // warp-uniform bound for the unrolled loops
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Dw = max(Dw, __shfl_xor_sync(FULLMASK, Dw, o));
    double u[VEC_DMAX], gc[VEC_DMAX];
    double2 xg[R];                                                 // Y side: the gathered x_e
    if (!V.x_side) {
#pragma unroll
      for (int r = 0; r < R; ++r) xg[r] = *reinterpret_cast<const double2*>(A.opp + (int64_t)j * stride + 2 * (lg + G * r));
    }
#pragma unroll
    for (int c = 0; c < VEC_DMAX; ++c) {
      u[c] = 0.0; gc[c] = 0.0;
      if (c < Dw) {
        double d = 0.0;
        if (V.x_side) {
          const double* yp = A.opp + (y0 + (c < D ? c : 0)) * stride + 2 * lg;
#pragma unroll
          for (int r = 0; r < R; ++r) { const double2 y = *reinterpret_cast<const double2*>(yp + 2 * G * r); d = fma(y.x, own[0][r].x, fma(y.y, own[0][r].y, d)); }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) d = fma(xg[r].x, own[c < DB ? c : 0][r].x, fma(xg[r].y, own[c < DB ? c : 0][r].y, d));
        }
        u[c] = group_sum<G>(d);
      }
    }
    if (act) {
      obj += vec_loss<GRAD>(code, lp, u, D, a, gc);
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < VEC_DMAX; ++c) {
          if (c < D) {
            if (V.x_side) {
              const double* yp = A.opp + (y0 + c) * stride + 2 * lg;
#pragma unroll
              for (int r = 0; r < R; ++r) { const double2 y = *reinterpret_cast<const double2*>(yp + 2 * G * r); grad[0][r].x = fma(gc[c], y.x, grad[0][r].x); grad[0][r].y = fma(gc[c], y.y, grad[0][r].y); }
            } else if (c < DB) {
#pragma unroll
              for (int r = 0; r < R; ++r) { grad[c][r].x = fma(gc[c], xg[r].x, grad[c][r].x); grad[c][r].y = fma(gc[c], xg[r].y, grad[c][r].y); }
            }
          }
        }
      }
    }
  }
  obj = cross_group_sum<G>(obj);
  if (GRAD) {
#pragma unroll
    for (int c = 0; c < DB; ++c) {
      if (c < Dunit) {
#pragma unroll
        for (int r = 0; r < R; ++r) { grad[c][r].x = cross_group_sum<G>(grad[c][r].x); grad[c][r].y = cross_group_sum<G>(grad[c][r].y); }
      }
    }
  }
  return obj;
}

// separable regularizers on a k x D block: evaluate / prox column by column (regularizers.jl: Quad, One, NonNeg,
// NonNegOne, Zero and the offset wrappers act element- or row-wise, so the block forms decompose exactly)
constexpr int ORDINAL_FLAGS = GLRMB200_REG_ORDINAL | GLRMB200_REG_MNL_ORDINAL;

// OrdinalReg / MNLOrdinalReg (regularizers.jl:356-407): the first k-1 rows of the block's columns are forced equal —
// mean over the columns, inner prox on that (k-1)-vector, copied back — and MNL additionally makes the last row
// negative and strictly decreasing across the columns.
template <int G, int R, int DB>
__device__ __forceinline__ void block_ordinal_prox(int code, const double* rp, double2 (&v)[DB][R], int D, int lg, int k, double alpha) {
  double2 um[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int c = 0; c < DB; ++c) if (c < D) { sx += v[c][r].x; sy += v[c][r].y; }
    um[r] = make_double2(sx / (double)D, sy / (double)D);
  }
  // inner prox sees rows 0..k-2 (the last-row slot of `um` is scratch and is not copied back)
  reg_prox<G, R>((code & GLRMB200_REG_BASE_MASK) | GLRMB200_REG_LASTENTRY_UNPENALIZED, rp, um, lg, k, alpha);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = 2 * (lg + G * r);
#pragma unroll
    for (int c = 0; c < DB; ++c) if (c < D) {
      if (i0 < k - 1) v[c][r].x = um[r].x;
      if (i0 + 1 < k - 1) v[c][r].y = um[r].y;
    }
    if (code & GLRMB200_REG_MNL_ORDINAL) {                       // :399-402, the lane that owns row k-1 walks the columns
      if (i0 == k - 1) {
        double prev = jl_mind(-1e-3, v[0][r].x);
        v[0][r].x = prev;
#pragma unroll
        for (int c = 1; c < DB; ++c) if (c < D) { prev = jl_mind(v[c][r].x, prev - 1e-3); v[c][r].x = prev; }
      }
      if (i0 + 1 == k - 1) {
        double prev = jl_mind(-1e-3, v[0][r].y);
        v[0][r].y = prev;
#pragma unroll
        for (int c = 1; c < DB; ++c) if (c < D) { prev = jl_mind(v[c][r].y, prev - 1e-3); v[c][r].y = prev; }
      }
    }
  }
}

template <int G, int R, int DB>
__device__ __forceinline__ double block_reg_eval(int code, const double* rp, const double2 (&v)[DB][R], int D, int lg, int k) {
  if (code & ORDINAL_FLAGS)                                      // evaluate(r.r, a[1:end-1, 1])  (:378,405)
    return reg_eval<G, R>((code & GLRMB200_REG_BASE_MASK) | GLRMB200_REG_LASTENTRY_UNPENALIZED, rp, v[0], lg, k);
  double t = 0.0;
#pragma unroll
  for (int c = 0; c < DB; ++c) if (c < D) t += reg_eval<G, R>(code, rp, v[c], lg, k);
  return t;
}

// DB = block capacity: 1 for rows of X, VEC_DMAX for block columns of Y
template <int G, int R, int DB>
__global__ void __launch_bounds__(128) vec_sweep_kernel(const VecArgs V) {
  const SweepArgs& A = V.s;
  const int lane = threadIdx.x & 31, lg = lane % G;
  const int64_t slot = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (slot >= A.n_units || sweep_stopped(A)) return;
  const int64_t unit = A.order[slot];
  const int k = A.k, stride = A.stride;
  int64_t start, len;
  if (A.ptr) { start = A.ptr[unit - A.unit_base]; len = A.ptr[unit - A.unit_base + 1] - start; }
  else { start = (unit - A.unit_base) * A.full_len; len = A.full_len; }
  const int64_t col0 = V.x_side ? unit : V.ystart[unit];
  const int D = V.x_side ? 1 : (int)(V.ystart[unit + 1] - col0);
  double* ownp = A.own + col0 * stride;
  double2 own[DB][R], grad[DB][R];
#pragma unroll
  for (int c = 0; c < DB; ++c)
#pragma unroll
    for (int r = 0; r < R; ++r)
      own[c][r] = c < D ? *reinterpret_cast<const double2*>(ownp + (int64_t)c * stride + 2 * (lg + G * r)) : make_double2(0.0, 0.0);
  const int rcode = A.reg_code[A.reg_uniform ? 0 : unit];
  const double* rp = A.reg_param + (A.reg_uniform ? 0 : unit) * GLRMB200_REG_NPARAM;

  double obj_old = vec_pass<G, R, DB, true>(V, unit, start, len, lane, own, D, grad);
  if (!(A.flags & FLAG_NO_REG)) obj_old += block_reg_eval<G, R, DB>(rcode, rp, own, D, lg, k);
  double alpha = A.alpha[unit];
  double obj_rec = obj_old;
  int ntrials = 0;
  if (!(A.flags & FLAG_EVAL_ONLY)) {
    const double l1 = (double)(len + 1);
    while (alpha > A.min_stepsize) {
      const double stepsize = alpha / l1;
      double2 nw[DB][R], dummy[DB][R];
#pragma unroll
      for (int c = 0; c < DB; ++c) {
#pragma unroll
        for (int r = 0; r < R; ++r) { nw[c][r].x = fma(-stepsize, grad[c][r].x, own[c][r].x); nw[c][r].y = fma(-stepsize, grad[c][r].y, own[c][r].y); }
        if (c < D && !(rcode & ORDINAL_FLAGS)) reg_prox<G, R>(rcode, rp, nw[c], lg, k, stepsize);
      }
      if (rcode & ORDINAL_FLAGS) block_ordinal_prox<G, R, DB>(rcode, rp, nw, D, lg, k, stepsize);
      double obj_new = vec_pass<G, R, DB, false>(V, unit, start, len, lane, nw, D, dummy);
      obj_new += block_reg_eval<G, R, DB>(rcode, rp, nw, D, lg, k);
      ++ntrials;
      if (obj_new < obj_old) {
        if (lane < G) {
#pragma unroll
          for (int c = 0; c < DB; ++c) if (c < D)
#pragma unroll
            for (int r = 0; r < R; ++r) *reinterpret_cast<double2*>(ownp + (int64_t)c * stride + 2 * (lg + G * r)) = nw[c][r];
        }
        alpha *= 1.05;
        obj_rec = obj_new;
        break;
      } else {
        alpha *= .7;
        if (alpha < A.min_stepsize) { alpha = A.min_stepsize * 1.1; break; }
      }
    }
  }
  if (lane == 0) {
    A.alpha[unit] = alpha;
    A.obj_out[unit] = obj_rec;
    if (ntrials && A.trial_counter) atomicAdd(A.trial_counter, (unsigned long long)ntrials);
  }
}

}  // namespace glrm
