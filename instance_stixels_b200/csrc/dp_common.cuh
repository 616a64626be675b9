// Pieces shared by the DP kernel and the backtracking kernel: the per-vB
// transition scalars ("row info") and the per-cell prior selection.  Both
// kernels must produce bit-identical priors, because backtracking re-derives
// the predecessor type the reference stores in index_table
// (StixelsKernels.cu:723-727, 769-773, 828-836).
#pragma once
#include "common.cuh"

namespace isx {

// Dynamic row info Q[vB] (pairwise), p = vB - 1.  C[p] are FINAL row costs.
//  0 gs_k   = pw * min over predecessors for the ground (p < vhor) or sky slot
//  1..5     ground side: P1_hi, P1_mid, P1_lo, t1_hi, t1_lo        (object <- ground)
//           sky side:    P3_yes, P3_no                              (object <- sky)
//  6..10    P2_hi, P2_lo, P2_mid, t2_hi, t2_lo                      (object <- object)
//  11       previous_mean (kept for backtracking/debug)
struct RowInfo {
  float gs_k;
  float a1, a2, a3, a4, a5;
  float p2_hi, p2_lo, p2_mid, t2_hi, t2_lo;
  float pm;
};

struct RowPriors {  // what index_table's predecessor type is decided from
  float g1, g2;     // ground/sky slot: prior from GROUND, from OBJECT
};

// previous_mean / obj_fn: ComputeMean (StixelsKernels.cu:47-60) + clamp (:651-653, 682-684).
__device__ __forceinline__ float segment_mean(float sum_hi, float sum_lo, float valid_hi, float valid_lo, int n,
                                              bool has_invalid) {
  float mean;
  if (has_invalid) {
    const float vd = fsub(valid_hi, valid_lo);
    mean = (vd != 0.0f) ? fmul(fsub(sum_hi, sum_lo), rcp_approx(vd)) : 0.0f;
  } else {
    mean = fmul(fsub(sum_hi, sum_lo), rcp_approx((float)n));
  }
  return clamp_neg(mean);
}

// S = static record of vB (tables.cu), cg/co/cs = C[p][GROUND/OBJECT/SKY].
__device__ __forceinline__ RowInfo make_row_info(const float *S, bool ground_side, float cg, float co, float cs,
                                                 float pm, const float *__restrict__ object_disparity_range,
                                                 const KParams &p, RowPriors *rp) {
  RowInfo q;
  const float pw = p.prior_weight;
  const float inf = inf_f();
  q.pm = pm;
  if (ground_side) {
    // ground <- {ground, object} (:694-711)
    const float g1 = ffma(S[9], pw, cg);
    const float g2 = ffma(S[9], pw, co);
    q.gs_k = fmul(fmin_(g1, g2), pw);
    rp->g1 = g1;
    rp->g2 = g2;
    // object <- ground, three cases of fn vs gf[p] +- eps (:120-144, 789-794)
    q.a1 = ffma(S[3], pw, cg);
    q.a2 = ffma(S[4], pw, cg);
    q.a3 = ffma(S[5], pw, cg);
    q.a4 = S[1];
    q.a5 = S[2];
  } else {
    // sky <- {ground, object} (:735-756)
    const float s1 = ffma(S[6], pw, cg);
    const float s2 = ffma((pm < p.epsilon) ? inf : S[10], pw, co);
    q.gs_k = fmul(fmin_(s1, s2), pw);
    rp->g1 = s1;
    rp->g2 = s2;
    // object <- sky (:173-183, 802-805)
    q.a1 = ffma(S[8], pw, cs);
    q.a2 = ffma(inf, pw, cs);
    q.a3 = q.a4 = q.a5 = 0.0f;
  }
  // object <- object (:146-171, 796-801)
  int ipm = (int)pm;  // F2I.TRUNC
  ipm = ipm < 0 ? 0 : (ipm >= p.max_dis ? p.max_dis - 1 : ipm);
  const float dd = clamp_neg(object_disparity_range[ipm]);
  q.t2_hi = fadd(pm, dd);
  q.t2_lo = fsub(pm, dd);
  const float tr_hi = fadd(S[7], neg_log_div(p.pord, fadd(-dd, fadd(-pm, p.max_disf))));
  const float tr_lo = fadd(S[7], neg_log_div(fadd(-p.pord, 1.0f), q.t2_lo));
  q.p2_hi = ffma(tr_hi, pw, co);
  q.p2_lo = ffma(tr_lo, pw, co);
  q.p2_mid = ffma(inf, pw, co);
  return q;
}

// The three object priors of one cell; fn = clamped segment mean.
__device__ __forceinline__ void object_priors(const RowInfo &q, bool ground_side, float fn, float epsilon,
                                              float &prior1, float &prior2, float &prior3) {
  const float inf = inf_f();
  prior2 = (fn > q.t2_hi) ? q.p2_hi : ((fn < q.t2_lo) ? q.p2_lo : q.p2_mid);
  if (ground_side) {
    prior1 = (fn > q.a4) ? q.a1 : ((fn < q.a5) ? q.a3 : q.a2);
    prior3 = inf;  // C[p][SKY] is +inf below the horizon
  } else {
    prior1 = inf;  // C[p][GROUND] is +inf at/above the horizon
    prior3 = (fn > epsilon) ? q.a1 : q.a2;
  }
}

// Predecessor type rules (:723-727, 769-773, 828-836).
__device__ __forceinline__ int prev_type_gs(float g1, float g2) { return (g1 < g2) ? GROUND : OBJECT; }
__device__ __forceinline__ int prev_type_obj(float prior1, float prior2, float prior3) {
  int t = (prior1 < prior2) ? GROUND : OBJECT;
  if (prior3 < fmin_(prior1, prior2)) t = SKY;
  return t;
}

}  // namespace isx
