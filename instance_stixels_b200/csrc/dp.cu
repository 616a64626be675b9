// The ground/object/sky dynamic program over one stixel column.
// Replaces the first-segment block and the vB loop of StixelsKernel<PAIRWISE>
// (InstanceStixels/src/StixelsKernels.cu:477-839).
//
// Mapping (B200): ONE WARP owns one (frame, column).  The column is walked in
// tiles of 32 rows; lane l owns vT = a + l and keeps the 32 prefix values of
// R[vT+1] in registers for the whole tile.  For every candidate bottom row vB
// the warp-uniform record R[vB] (and, pairwise, the row info Q[vB]) is fetched
// with 128-bit broadcast loads, the cell is evaluated in registers, and a
// running (cost, vB) minimum with the reference's strict-< rule (lowest vB
// wins ties) is updated.  Rows below the tile are final, so the rectangular
// part needs no synchronisation at all; only the 32 steps inside the tile
// form the wavefront, and they exchange the finished row through warp
// shuffles.  No shared-memory barriers, no tensor cores (min-plus recurrence).
//
// Because one of GROUND/SKY is +inf for every row (ground only exists below
// the horizon, sky only at/above it) the cost table keeps two slots per row:
// "gs" (ground if vT < vhor else sky) and "object".
#include "dp_common.cuh"
#include "kernels.h"

namespace isx {

unsigned long long g_launch_count = 0;

namespace {

constexpr int kDpWarps = 4;
constexpr int kDpThreads = kDpWarps * 32;

__device__ __forceinline__ long long i64_from(uint32_t lo, uint32_t hi) {
  return (long long)(((unsigned long long)hi << 32) | lo);
}

template <bool PAIRWISE>
__global__ void __launch_bounds__(kDpThreads)
dp_kernel(const uint4 *__restrict__ records, const float *__restrict__ object_lut, const float *__restrict__ stat,
          float *dyn, const int *__restrict__ vhor_arr, const float *__restrict__ object_disparity_range,
          const float *__restrict__ inverse_height, float4 *__restrict__ dp_out, int ncolumns, KParams p) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gcol = blockIdx.x * kDpWarps + (threadIdx.x >> 5);  // frame * C + column
  if (gcol >= ncolumns) return;
  const int H = p.rows, C = p.realcols;
  const int f = gcol / C;
  const int vhor = vhor_arr[f];
  const bool has_invalid = p.invalid_disparity >= 0.0f;
  const float pw = p.prior_weight, dw = p.disparity_weight, sw = p.segmentation_weight, iw = p.instance_weight;
  const float inf = inf_f();

  const uint4 *rec = records + (size_t)gcol * p.rec_rows * (kRecWords / 4);
  const float *lut = object_lut + (size_t)gcol * p.max_dis * p.lut_stride;
  const float *S = stat + (size_t)f * H * kStatWords;
  float *Q = dyn + (size_t)gcol * H * kDynWords;
  float4 *out = dp_out + (size_t)gcol * H;

  // first-segment priors (:189-199)
  const float first_g_k = fmul(ffma(1.0f, kLn2, p.rows_log), pw);

  for (int a = 0; a < H; a += 32) {
    const int vT = a + lane;
    const bool row_ok = vT < H;
    const int vTc = row_ok ? vT : H - 1;

    // ---- A side: R[vT+1] into registers ----
    uint32_t A[kRecWords];
    {
      const uint4 *ra = rec + (size_t)(vTc + 1) * (kRecWords / 4);
#pragma unroll
      for (int k = 0; k < kRecWords / 4; k++) {
        const uint4 t = __ldg(ra + k);
        A[4 * k] = t.x; A[4 * k + 1] = t.y; A[4 * k + 2] = t.z; A[4 * k + 3] = t.w;
      }
    }
    const long long a_mx = i64_from(A[kRecMx], A[kRecMx + 1]), a_my = i64_from(A[kRecMy], A[kRecMy + 1]);
    const long long a_mx2 = i64_from(A[kRecMx2], A[kRecMx2 + 1]), a_my2 = i64_from(A[kRecMy2], A[kRecMy2 + 1]);
    const float a_disp = __uint_as_float(A[kRecDisp]), a_valid = __uint_as_float(A[kRecValid]);
    const float a_ground = __uint_as_float(A[kRecGround]), a_sky = __uint_as_float(A[kRecSky]);
    const float *lut_a = lut + vTc;
    const float first_o_pvt = (vT <= vhor) ? kLn2 : 0.0f;

    float best_gs = inf, best_o = inf;
    int vb_gs = 0, vb_o = 0;

    const int vb_end = min(a + 31, H - 1);
    for (int vB = 0; vB <= vb_end; vB++) {
      const bool ground_side = (vB == 0) || (vB - 1 < vhor);
      RowInfo q;
      if constexpr (PAIRWISE) {
        if (vB > a) {
          // ---- wavefront step: row pv = vB-1 (lane vB-1-a) is final ----
          const int src = vB - 1 - a;
          const float c_gs = __shfl_sync(full, best_gs, src);
          const float c_o = __shfl_sync(full, best_o, src);
          const int o_vb = __shfl_sync(full, vb_o, src);
          const float hi_d = __shfl_sync(full, a_disp, src), hi_v = __shfl_sync(full, a_valid, src);
          const uint4 lo = __ldg(rec + (size_t)o_vb * (kRecWords / 4) + kRecDisp / 4);
          const float pm = segment_mean(hi_d, __uint_as_float(lo.x), hi_v, __uint_as_float(lo.y), vB - o_vb,
                                        has_invalid);
          RowPriors rp;
          q = make_row_info(S + (size_t)vB * kStatWords, ground_side, ground_side ? c_gs : inf, c_o,
                            ground_side ? inf : c_gs, pm, object_disparity_range, p, &rp);
          if (lane == 0) {
            float4 *qd = reinterpret_cast<float4 *>(Q + (size_t)vB * kDynWords);
            qd[0] = make_float4(q.gs_k, q.a1, q.a2, q.a3);
            qd[1] = make_float4(q.a4, q.a5, q.p2_hi, q.p2_lo);
            qd[2] = make_float4(q.p2_mid, q.t2_hi, q.t2_lo, q.pm);
          }
        } else if (vB > 0) {
          const float4 *qd = reinterpret_cast<const float4 *>(Q + (size_t)vB * kDynWords);
          const float4 q0 = qd[0], q1 = qd[1], q2 = qd[2];
          q.gs_k = q0.x; q.a1 = q0.y; q.a2 = q0.z; q.a3 = q0.w;
          q.a4 = q1.x; q.a5 = q1.y; q.p2_hi = q1.z; q.p2_lo = q1.w;
          q.p2_mid = q2.x; q.t2_hi = q2.y; q.t2_lo = q2.z; q.pm = q2.w;
        }
      }

      // ---- B side: R[vB], warp-uniform ----
      uint32_t Bw[kRecWords];
      {
        const uint4 *rb = rec + (size_t)vB * (kRecWords / 4);
#pragma unroll
        for (int k = 0; k < kRecWords / 4; k++) {
          const uint4 t = __ldg(rb + k);
          Bw[4 * k] = t.x; Bw[4 * k + 1] = t.y; Bw[4 * k + 2] = t.z; Bw[4 * k + 3] = t.w;
        }
      }

      // ---- semantic sums (Cityscapes.h:28-123), exact ints; min commutes with the
      //      monotone int->float conversion and the float add of the offset term ----
      const int s_road = (int)(A[0] - Bw[0]), s_side = (int)(A[1] - Bw[1]);
      const int s_sky = (int)(A[kSkyClass] - Bw[kSkyClass]);
      int s_ni = (int)(A[2] - Bw[2]);
#pragma unroll
      for (int c = 3; c < 10; c++) s_ni = min(s_ni, (int)(A[c] - Bw[c]));
      int s_in = (int)(A[11] - Bw[11]);
#pragma unroll
      for (int c = 12; c < 19; c++) s_in = min(s_in, (int)(A[c] - Bw[c]));
      const int s_off = (int)(A[kRecOff] - Bw[kRecOff]);
      const float nic = fmul((float)s_off, iw);  // ComputeNonInstanceOffsetCost * weight (:618-621)

      // ---- instance variance term (:72-86, 611-616) ----
      const int n = max(vTc + 1 - vB, 1);  // dead lanes (vT < vB) compute a harmless dummy cell
      const float rn = rcp_approx((float)n);
      const float fmx = __ll2float_rn(a_mx - i64_from(Bw[kRecMx], Bw[kRecMx + 1]));
      const float fmy = __ll2float_rn(a_my - i64_from(Bw[kRecMy], Bw[kRecMy + 1]));
      const float fmx2 = __ll2float_rn(a_mx2 - i64_from(Bw[kRecMx2], Bw[kRecMx2 + 1]));
      const float fmy2 = __ll2float_rn(a_my2 - i64_from(Bw[kRecMy2], Bw[kRecMy2 + 1]));
      const float var = ffma(-fmul(fmy, fmy), rn, fadd(fmy2, ffma(-fmul(fmx, fmx), rn, fmx2)));
      const float ic = fmul(var, iw);
      const float seg_o = fmin_(fadd(nic, (float)s_ni), fadd(ic, (float)s_in));
      // In the first-segment block nvcc contracted `min(road, sidewalk) + weight * offsets` into one
      // FFMA (reference SASS of StixelsKernels.cu:502-506); everywhere else it is FMUL + FADD.
      const float seg_gs = vB == 0 ? ffma((float)s_off, iw, (float)min(s_road, s_side))
                                   : fadd(nic, ground_side ? (float)min(s_road, s_side) : (float)s_sky);

      // ---- disparity terms ----
      const float fn = segment_mean(a_disp, __uint_as_float(Bw[kRecDisp]), a_valid, __uint_as_float(Bw[kRecValid]),
                                    n, has_invalid);
      int fni = __float2int_rd(fn);
      fni = fni < 0 ? 0 : (fni >= p.max_dis ? p.max_dis - 1 : fni);
      const float lut_hi = __ldg(lut_a + (size_t)fni * p.lut_stride);
      const float lut_lo = vB > 0 ? __ldg(lut + (size_t)fni * p.lut_stride + (vB - 1)) : 0.0f;
      const float data_o = fsub(lut_hi, lut_lo);
      const float data_gs = ground_side ? fsub(a_ground, __uint_as_float(Bw[kRecGround]))
                                        : fsub(a_sky, __uint_as_float(Bw[kRecSky]));

      // ---- combine (:548-560, 575-584, 700-720, 740-766, 788-824) ----
      float cost_gs, cost_o;
      if constexpr (PAIRWISE) {
        float k_gs, k_o;
        if (vB == 0) {
          k_gs = first_g_k;
          k_o = fmul(fadd(fadd(first_o_pvt, p.rows_log), p.max_dis_log), pw);
        } else {
          float p1, p2, p3;
          object_priors(q, ground_side, fn, p.epsilon, p1, p2, p3);
          k_gs = q.gs_k;
          k_o = fmul(fmin_(p3, fmin_(p1, p2)), pw);
        }
        cost_gs = ffma(seg_gs, sw, ffma(data_gs, dw, k_gs));
        cost_o = ffma(seg_o, sw, ffma(data_o, dw, k_o));
      } else {
        const float ih = __ldg(inverse_height + n);
        cost_gs = ffma(seg_gs, sw, ffma(ih, pw, fmul(data_gs, dw)));
        cost_o = ffma(seg_o, sw, ffma(ih, pw, fmul(data_o, dw)));
      }
      const bool live = row_ok && vT >= vB;
      if (live && cost_gs < best_gs) { best_gs = cost_gs; vb_gs = vB; }
      if (live && cost_o < best_o) { best_o = cost_o; vb_o = vB; }
    }

    if constexpr (PAIRWISE) {
      // last row of the tile: publish Q[a+32] for the next tile
      const int vB = a + 32;
      if (vB < H) {
        const bool ground_side = vB - 1 < vhor;
        const float c_gs = __shfl_sync(full, best_gs, 31);
        const float c_o = __shfl_sync(full, best_o, 31);
        const int o_vb = __shfl_sync(full, vb_o, 31);
        const float hi_d = __shfl_sync(full, a_disp, 31), hi_v = __shfl_sync(full, a_valid, 31);
        const uint4 lo = __ldg(rec + (size_t)o_vb * (kRecWords / 4) + kRecDisp / 4);
        const float pm =
            segment_mean(hi_d, __uint_as_float(lo.x), hi_v, __uint_as_float(lo.y), vB - o_vb, has_invalid);
        RowPriors rp;
        const RowInfo q = make_row_info(S + (size_t)vB * kStatWords, ground_side, ground_side ? c_gs : inf, c_o,
                                        ground_side ? inf : c_gs, pm, object_disparity_range, p, &rp);
        if (lane == 0) {
          float4 *qd = reinterpret_cast<float4 *>(Q + (size_t)vB * kDynWords);
          qd[0] = make_float4(q.gs_k, q.a1, q.a2, q.a3);
          qd[1] = make_float4(q.a4, q.a5, q.p2_hi, q.p2_lo);
          qd[2] = make_float4(q.p2_mid, q.t2_hi, q.t2_lo, q.pm);
        }
      }
    }
    if (row_ok) out[vT] = make_float4(best_gs, best_o, __int_as_float(vb_gs), __int_as_float(vb_o));
    __syncwarp();  // orders this tile's Q stores before the next tile's loads
  }
}

}  // namespace

void launch_dp(const KParams &p, const BatchBuffers &b, int nframes, bool pairwise, cudaStream_t s) {
  const int ncolumns = nframes * p.realcols;
  const int grid = (ncolumns + kDpWarps - 1) / kDpWarps;
  const uint4 *rec = reinterpret_cast<const uint4 *>(b.records);
  if (pairwise)
    dp_kernel<true><<<grid, kDpThreads, 0, s>>>(rec, b.object_lut, b.stat, b.dyn, b.vhor, b.object_disparity_range,
                                                b.inverse_height, b.dp, ncolumns, p);
  else
    dp_kernel<false><<<grid, kDpThreads, 0, s>>>(rec, b.object_lut, b.stat, b.dyn, b.vhor,
                                                 b.object_disparity_range, b.inverse_height, b.dp, ncolumns, p);
  g_launch_count++;
}

}  // namespace isx
