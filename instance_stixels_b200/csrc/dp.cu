// The ground/object/sky dynamic program over one stixel column.
// Replaces the first-segment block and the vB loop of StixelsKernel<PAIRWISE>
// (InstanceStixels/src/StixelsKernels.cu:477-839).
//
// Mapping (B200): ONE CTA (8 warps) owns one (frame, column).  The column's rows
// are cut into tiles of 32 top rows (lane l of a warp owns vT = a + l) and into
// chunks of 32 candidate bottom rows vB.  The CTA walks the vB chunks in
// lockstep: the chunk's 32 prefix records ("B side") are staged once in shared
// memory, and every tile at or above the chunk is one unit of work for one
// warp: it loads its own records R[vT+1] ("A side", coalesced word-major lines)
// into registers, evaluates the 32 x 32 cells of the unit with the B side read
// as 128-bit shared-memory broadcasts, and keeps the running (cost, vB) minimum
// with the reference's strict-< rule (lowest vB wins ties).  Because all warps
// of the CTA are inside the same vB chunk of the same column at the same time,
// the object-LUT gathers of a step hit the same few L1 lines.
//
// Pairwise mode adds the wavefront: the diagonal unit (tile == chunk) is
// processed first by warp 0, which finalises row vB-1 before it evaluates vB,
// exchanges it through warp shuffles and publishes the per-vB transition
// scalars Q[vB] in shared memory for the off-diagonal units of the chunk.
// No tensor cores (min-plus recurrence).
//
// Because one of GROUND/SKY is +inf for every row (ground only exists below
// the horizon, sky only at/above it) the cost table keeps two slots per row:
// "gs" (ground if vT < vhor else sky) and "object".
#include "dp_common.cuh"
#include "kernels.h"

namespace isx {

unsigned long long g_launch_count = 0;

namespace {

constexpr int kDpWarps = 8;
constexpr int kDpThreads = kDpWarps * 32;
constexpr int kChunk = 32;
constexpr int kBStride = 36;  // words per staged B row: >= 32, rows stay 16-byte aligned
constexpr int kStageWords = kRecWords * kChunk;
constexpr int kStagePerThread = (kStageWords + kDpThreads - 1) / kDpThreads;
constexpr int kQsRows = kChunk + 1;

struct DpConsts {
  float pw, dw, sw, iw;
  float dm1f;          // max_dis - 1 as float (LUT row clamp)
  unsigned lut_stride4;  // bytes per fn row of the object LUT
  bool has_invalid;
  float epsilon;
};

__device__ __forceinline__ float f_(uint32_t u) { return __uint_as_float(u); }

// One DP cell per lane: segment (vB .. vT) of this lane's row vT.
//   A      : R[vT+1] in registers,  brow: R[vB] in shared memory (warp-uniform address)
//   nf, n  : segment height as float / int,  rn = MUFU.RCP(nf)
template <bool PAIRWISE, bool FIRST, bool GROUND, bool HAS_INVALID>
__device__ __forceinline__ void dp_cell(const uint32_t (&A)[kRecWords], const uint32_t *__restrict__ brow,
                                        const char *pa, const char *pb,
                                        float nf, float ih, const RowInfo &q, float first_k_gs, float first_k_o,
                                        const DpConsts &c, float &cost_gs, float &cost_o) {
  uint32_t Bw[32];
  {
    const uint4 *b4 = reinterpret_cast<const uint4 *>(brow);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint4 t = b4[k];
      Bw[4 * k] = t.x; Bw[4 * k + 1] = t.y; Bw[4 * k + 2] = t.z; Bw[4 * k + 3] = t.w;
    }
  }
  // ---- semantic sums (Cityscapes.h:28-123), exact ints; min commutes with the
  //      monotone int->float conversion and the float add of the offset term ----
  int s_ni = (int)(A[2] - Bw[2]);
#pragma unroll
  for (int k = 3; k < 10; k++) s_ni = min(s_ni, (int)(A[k] - Bw[k]));
  int s_in = (int)(A[11] - Bw[11]);
#pragma unroll
  for (int k = 12; k < 19; k++) s_in = min(s_in, (int)(A[k] - Bw[k]));
  const int s_off = (int)(A[kRecOff] - Bw[kRecOff]);
  const int s_gs = GROUND ? min((int)(A[0] - Bw[0]), (int)(A[1] - Bw[1])) : (int)(A[kSkyClass] - Bw[kSkyClass]);
  const float f_off = (float)s_off;
  const float nic = fmul(f_off, c.iw);  // ComputeNonInstanceOffsetCost * weight (:618-621)

  // ---- instance variance term (:72-86, 611-616); float prefixes reproduce I2F.S64(diff) ----
  const float rn = rcp_approx(nf);
  const float fmx = fsub(f_(A[kRecMx]), f_(Bw[kRecMx]));
  const float fmy = fsub(f_(A[kRecMy]), f_(Bw[kRecMy]));
  const float fmx2 = fadd(fsub(f_(A[kRecMx2Hi]), f_(Bw[kRecMx2Hi])), fsub(f_(A[kRecMx2Lo]), f_(Bw[kRecMx2Lo])));
  const float fmy2 = fadd(fsub(f_(A[kRecMy2Hi]), f_(Bw[kRecMy2Hi])), fsub(f_(A[kRecMy2Lo]), f_(Bw[kRecMy2Lo])));
  const float var = ffma(-fmul(fmy, fmy), rn, fadd(fmy2, ffma(-fmul(fmx, fmx), rn, fmx2)));
  const float ic = fmul(var, c.iw);
  const float seg_o = fmin_(fadd(nic, (float)s_ni), fadd(ic, (float)s_in));
  // In the first-segment block nvcc contracted `min(road, sidewalk) + weight * offsets` into one
  // FFMA (reference SASS of StixelsKernels.cu:502-506); everywhere else it is FMUL + FADD.
  const float seg_gs = FIRST ? ffma(f_off, c.iw, (float)s_gs) : fadd(nic, (float)s_gs);

  // ---- disparity terms: ComputeMean (:47-60) + clamp (:651-653) ----
  const float sd = fsub(f_(A[kRecDisp]), f_(Bw[kRecDisp]));
  float mean;
  if constexpr (HAS_INVALID) {
    const float vd = fsub(f_(A[kRecValid]), f_(Bw[kRecValid]));
    const float m = fmul(sd, rcp_approx(vd));
    mean = (vd != 0.0f) ? m : 0.0f;
  } else {
    mean = fmul(sd, rn);
  }
  const float fn = fmaxf(mean, 0.0f);  // == clamp_neg for every comparison below (mean is never NaN)
  // floor(fn) as LUT row: add.rz of 2^23 leaves floor(fn) in the mantissa; pa / pb are the byte
  // addresses of LUT[0][vT] / LUT[0][vB-1] minus 0x4B000000 rows, so one 32x32+64 multiply-add
  // (IMAD.WIDE.U32) forms each address.
  const float fbias = __fadd_rz(fminf(fn, c.dm1f), 8388608.0f);
  const unsigned long long roff = (unsigned long long)(unsigned)__float_as_int(fbias) * c.lut_stride4;
  const float lut_hi = __ldg(reinterpret_cast<const float *>(pa + roff));
  const float lut_lo = FIRST ? 0.0f : __ldg(reinterpret_cast<const float *>(pb + roff));
  const float data_o = fsub(lut_hi, lut_lo);
  const float data_gs = GROUND ? fsub(f_(A[kRecGround]), f_(Bw[kRecGround])) : fsub(f_(A[kRecSky]), f_(Bw[kRecSky]));

  // ---- combine (:548-560, 575-584, 700-720, 740-766, 788-824) ----
  if constexpr (PAIRWISE) {
    float k_gs, k_o;
    if constexpr (FIRST) {
      k_gs = first_k_gs;
      k_o = first_k_o;
    } else {
      float p1, p2, p3;
      object_priors(q, GROUND, fn, c.epsilon, p1, p2, p3);
      k_gs = q.gs_k;
      k_o = fmul(fmin_(p3, fmin_(p1, p2)), c.pw);
    }
    cost_gs = ffma(seg_gs, c.sw, ffma(data_gs, c.dw, k_gs));
    cost_o = ffma(seg_o, c.sw, ffma(data_o, c.dw, k_o));
  } else {
    cost_gs = ffma(seg_gs, c.sw, ffma(ih, c.pw, fmul(data_gs, c.dw)));
    cost_o = ffma(seg_o, c.sw, ffma(ih, c.pw, fmul(data_o, c.dw)));
  }
}

__device__ __forceinline__ RowInfo load_row_info(const float *__restrict__ qrow) {
  const float4 *qd = reinterpret_cast<const float4 *>(qrow);
  const float4 q0 = qd[0], q1 = qd[1], q2 = qd[2];
  RowInfo q;
  q.gs_k = q0.x; q.a1 = q0.y; q.a2 = q0.z; q.a3 = q0.w;
  q.a4 = q1.x; q.a5 = q1.y; q.p2_hi = q1.z; q.p2_lo = q1.w;
  q.p2_mid = q2.x; q.t2_hi = q2.y; q.t2_lo = q2.z; q.pm = q2.w;
  return q;
}

__device__ __forceinline__ void store_row_info(float *qrow, const RowInfo &q) {
  float4 *qd = reinterpret_cast<float4 *>(qrow);
  qd[0] = make_float4(q.gs_k, q.a1, q.a2, q.a3);
  qd[1] = make_float4(q.a4, q.a5, q.p2_hi, q.p2_lo);
  qd[2] = make_float4(q.p2_mid, q.t2_hi, q.t2_lo, q.pm);
}

// Steps [k0, k1) of one unit; the whole range lies on one side of the horizon.
//   DIAG: tile == chunk, lane l is live for k <= l only (vT >= vB).
template <bool PAIRWISE, bool GROUND, bool DIAG, bool HAS_INVALID>
__device__ __forceinline__ void dp_steps(const uint32_t (&A)[kRecWords], const uint32_t *__restrict__ bchunk,
                                         const char *lutb, const char *pa, const float *__restrict__ ihs,
                                         const float *__restrict__ qs, int vb0, int k0, int k1, int n0, int lane,
                                         const DpConsts &c, float &best_gs, float &best_o, int &vb_gs, int &vb_o) {
  RowInfo q{};
  float nf = (float)(n0 - k0);       // segment height vT + 1 - vB, kept as a float counter
  const float *ihp = ihs + (n0 - k0);
#pragma unroll 2
  for (int k = k0; k < k1; k++) {
    const int vB = vb0 + k;
    // dead lanes of the diagonal unit (vT < vB) evaluate a harmless dummy cell of height >= 1
    const float nfc = DIAG ? fmaxf(nf, 1.0f) : nf;
    float ih = 0.0f;
    if constexpr (PAIRWISE) q = load_row_info(qs + k * kDynWords);
    else ih = DIAG ? ihs[max(n0 - k, 1)] : *ihp;
    float cost_gs, cost_o;
    dp_cell<PAIRWISE, false, GROUND, HAS_INVALID>(A, bchunk + k * kBStride, pa, lutb + 4 * (vB - 1), nfc, ih, q,
                                                  0.0f, 0.0f, c, cost_gs, cost_o);
    const bool live = !DIAG || lane >= k;
    if (live && cost_gs < best_gs) { best_gs = cost_gs; vb_gs = vB; }
    if (live && cost_o < best_o) { best_o = cost_o; vb_o = vB; }
    nf = fadd(nf, -1.0f);
    ihp--;
  }
}

// previous_mean of the best object segment ending at row pv (:674-685) and the row info of vB = pv + 1.
__device__ __forceinline__ RowInfo finish_row(const uint32_t *__restrict__ rec, int Hp, const float *__restrict__ S,
                                              int vB, int vhor, float c_gs, float c_o, int o_vb, float hi_d,
                                              float hi_v, const float *__restrict__ object_disparity_range,
                                              const KParams &p, bool has_invalid) {
  const bool ground_side = vB - 1 < vhor;
  const float lo_d = f_(__ldg(rec + (size_t)kRecDisp * Hp + o_vb));
  const float lo_v = f_(__ldg(rec + (size_t)kRecValid * Hp + o_vb));
  const float pm = segment_mean(hi_d, lo_d, hi_v, lo_v, vB - o_vb, has_invalid);
  const float inf = inf_f();
  RowPriors rp;
  return make_row_info(S + (size_t)vB * kStatWords, ground_side, ground_side ? c_gs : inf, c_o,
                       ground_side ? inf : c_gs, pm, object_disparity_range, p, &rp);
}

template <bool PAIRWISE, bool HAS_INVALID>
__global__ void __launch_bounds__(kDpThreads, 2)
dp_kernel(const uint32_t *__restrict__ records, const float *__restrict__ object_lut, const float *__restrict__ stat,
          float *__restrict__ pm_out, const int *__restrict__ vhor_arr,
          const float *__restrict__ object_disparity_range, const float *__restrict__ inverse_height,
          float4 *__restrict__ dp_out, KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gcol = blockIdx.x;  // frame * C + column
  const int H = p.rows, C = p.realcols, Hp = p.rec_stride;
  const int f = gcol / C;
  const int vhor = vhor_arr[f];
  const int nchunks = (H + kChunk - 1) / kChunk;
  const float inf = inf_f();

  uint32_t *brec = reinterpret_cast<uint32_t *>(smem_raw);                       // [2][32][kBStride]
  float4 *best = reinterpret_cast<float4 *>(smem_raw + 2 * kChunk * kBStride * 4);  // [32 * nchunks]
  float *extra = reinterpret_cast<float *>(best + kChunk * nchunks);             // unary: ih[H+1]; pairwise: qs[33][12]

  DpConsts c;
  c.pw = p.prior_weight; c.dw = p.disparity_weight; c.sw = p.segmentation_weight; c.iw = p.instance_weight;
  c.dm1f = (float)(p.max_dis - 1);
  c.lut_stride4 = (unsigned)p.lut_stride * 4u;
  c.has_invalid = HAS_INVALID;
  c.epsilon = p.epsilon;

  const uint32_t *rec = records + (size_t)gcol * kRecWords * Hp;
  const float *lut = object_lut + (size_t)gcol * p.max_dis * p.lut_stride;
  // byte address of LUT[0][0] of this column minus the 2^23 bias rows (see dp_cell); opaque to the
  // compiler so that it stays one materialised 64-bit base
  const char *lutb = reinterpret_cast<const char *>(lut) - (unsigned long long)0x4B000000u * c.lut_stride4;
  asm volatile("" : "+l"(lutb));
  const float *S = stat + (size_t)f * H * kStatWords;
  float *pm_col = pm_out + (size_t)gcol * H;

  for (int i = tid; i < kChunk * nchunks; i += kDpThreads) best[i] = make_float4(inf, inf, 0.0f, 0.0f);
  if constexpr (!PAIRWISE)
    for (int i = tid; i <= H; i += kDpThreads) extra[i] = __ldg(inverse_height + i);

  // first-segment priors (:189-199)
  const float first_k_gs = fmul(ffma(1.0f, kLn2, p.rows_log), c.pw);

  // B-side staging: word-major global lines -> row-major shared rows, one chunk ahead in registers
  uint32_t pre[kStagePerThread];
  auto load_stage = [&](int j) {
#pragma unroll
    for (int i = 0; i < kStagePerThread; i++) {
      const int e = tid + i * kDpThreads;
      if (e < kStageWords) pre[i] = __ldg(rec + (size_t)(e >> 5) * Hp + j * kChunk + (e & 31));
    }
  };
  auto store_stage = [&](int buf) {
    uint32_t *dst = brec + buf * kChunk * kBStride;
#pragma unroll
    for (int i = 0; i < kStagePerThread; i++) {
      const int e = tid + i * kDpThreads;
      if (e < kStageWords) dst[(e & 31) * kBStride + (e >> 5)] = pre[i];
    }
  };
  load_stage(0);

  for (int j = 0; j < nchunks; j++) {
    store_stage(j & 1);
    __syncthreads();
    if (j + 1 < nchunks) load_stage(j + 1);
    const uint32_t *bchunk = brec + (j & 1) * kChunk * kBStride;
    const int vb0 = j * kChunk;
    const int nsteps = min(kChunk, H - vb0);
    // steps with vB <= vhor are on the ground side (predecessor row vB-1 below the horizon)
    const int kg = max(0, min(nsteps, vhor + 1 - vb0));

    if constexpr (PAIRWISE) {
      // ---- diagonal unit: warp 0 runs the wavefront and publishes Q[vb0 + 1 ..] ----
      if (warp == 0) {
        float *qs = extra;
        const int vT = vb0 + lane;
        const bool row_ok = vT < H;
        const int vTc = row_ok ? vT : H - 1;
        uint32_t A[kRecWords];
#pragma unroll
        for (int w = 0; w < kRecWords; w++) A[w] = __ldg(rec + (size_t)w * Hp + vTc + 1);
        const float a_disp = f_(A[kRecDisp]), a_valid = f_(A[kRecValid]);
        const char *pa = lutb + 4 * vTc;
        // start from the minima over the earlier chunks (lower vB: they keep winning ties)
        const float4 prev = best[vb0 + lane];
        float best_gs = prev.x, best_o = prev.y;
        int vb_gs = __float_as_int(prev.z), vb_o = __float_as_int(prev.w);
        if (j > 0 && lane < kDynWords) qs[lane] = qs[kChunk * kDynWords + lane];  // Q[vb0] from the previous diagonal
        __syncwarp();
        for (int k = 0; k < nsteps; k++) {
          const int vB = vb0 + k;
          const int n = max(vTc + 1 - vB, 1);
          RowInfo q{};
          if (k > 0) {
            // row vB-1 (lane k-1) is final
            const float c_gs = __shfl_sync(full, best_gs, k - 1), c_o = __shfl_sync(full, best_o, k - 1);
            const int o_vb = __shfl_sync(full, vb_o, k - 1);
            const float hi_d = __shfl_sync(full, a_disp, k - 1), hi_v = __shfl_sync(full, a_valid, k - 1);
            q = finish_row(rec, Hp, S, vB, vhor, c_gs, c_o, o_vb, hi_d, hi_v, object_disparity_range, p,
                           c.has_invalid);
            if (lane == 0) {
              store_row_info(qs + k * kDynWords, q);
              pm_col[vB] = q.pm;
            }
          } else if (vB > 0) {
            q = load_row_info(qs);
          }
          float cost_gs, cost_o;
          const uint32_t *brow = bchunk + k * kBStride;
          if (vB == 0) {
            const float first_k_o = fmul(fadd(fadd((vT <= vhor) ? kLn2 : 0.0f, p.rows_log), p.max_dis_log), c.pw);
            dp_cell<true, true, true, HAS_INVALID>(A, brow, pa, lutb, (float)n, 0.0f, q, first_k_gs, first_k_o, c,
                                                   cost_gs, cost_o);
          } else if (k < kg) {
            dp_cell<true, false, true, HAS_INVALID>(A, brow, pa, lutb + 4 * (vB - 1), (float)n, 0.0f, q, 0.0f,
                                                    0.0f, c, cost_gs, cost_o);
          } else {
            dp_cell<true, false, false, HAS_INVALID>(A, brow, pa, lutb + 4 * (vB - 1), (float)n, 0.0f, q, 0.0f,
                                                     0.0f, c, cost_gs, cost_o);
          }
          const bool live = row_ok && lane >= k;
          if (live && cost_gs < best_gs) { best_gs = cost_gs; vb_gs = vB; }
          if (live && cost_o < best_o) { best_o = cost_o; vb_o = vB; }
        }
        // Q[vb0 + 32] for the next chunk (row vb0 + 31 is final now)
        if (vb0 + kChunk < H) {
          const int vB = vb0 + kChunk;
          const float c_gs = __shfl_sync(full, best_gs, 31), c_o = __shfl_sync(full, best_o, 31);
          const int o_vb = __shfl_sync(full, vb_o, 31);
          const float hi_d = __shfl_sync(full, a_disp, 31), hi_v = __shfl_sync(full, a_valid, 31);
          const RowInfo q = finish_row(rec, Hp, S, vB, vhor, c_gs, c_o, o_vb, hi_d, hi_v, object_disparity_range, p,
                                       c.has_invalid);
          if (lane == 0) {
            store_row_info(qs + kChunk * kDynWords, q);
            pm_col[vB] = q.pm;
          }
        }
        // the diagonal unit is the last one of its tile: the rows are final
        best[vb0 + lane] = make_float4(best_gs, best_o, __int_as_float(vb_gs), __int_as_float(vb_o));
      }
      __syncthreads();
    }

    // ---- units of this chunk: tiles t >= j (pairwise: t > j), round-robin over the warps ----
    for (int t = j + (PAIRWISE ? 1 : 0) + warp; t < nchunks; t += kDpWarps) {
      const int vT = t * kChunk + lane;
      const bool row_ok = vT < H;
      const int vTc = row_ok ? vT : H - 1;
      uint32_t A[kRecWords];
#pragma unroll
      for (int w = 0; w < kRecWords; w++) A[w] = __ldg(rec + (size_t)w * Hp + vTc + 1);
      const char *pa = lutb + 4 * vTc;
      const float4 prev = best[t * kChunk + lane];
      float best_gs = prev.x, best_o = prev.y;
      int vb_gs = __float_as_int(prev.z), vb_o = __float_as_int(prev.w);
      const int n0 = vTc + 1 - vb0;
      int k0 = 0;
      if (j == 0) {
        // first segment, vB = 0 (:481-594)
        float cost_gs, cost_o;
        RowInfo q{};
        const float first_k_o =
            PAIRWISE ? fmul(fadd(fadd((vT <= vhor) ? kLn2 : 0.0f, p.rows_log), p.max_dis_log), c.pw) : 0.0f;
        dp_cell<PAIRWISE, true, true, HAS_INVALID>(A, bchunk, pa, lutb, (float)n0, PAIRWISE ? 0.0f : extra[n0], q,
                                                   first_k_gs, first_k_o, c, cost_gs, cost_o);
        if (cost_gs < best_gs) { best_gs = cost_gs; vb_gs = 0; }
        if (cost_o < best_o) { best_o = cost_o; vb_o = 0; }
        k0 = 1;
      }
      if (!PAIRWISE && t == j) {  // unary only: the diagonal unit needs the vT >= vB predicate
        dp_steps<PAIRWISE, true, true, HAS_INVALID>(A, bchunk, lutb, pa, extra, extra, vb0, k0, max(k0, kg), n0, lane,
                                                    c, best_gs, best_o, vb_gs, vb_o);
        dp_steps<PAIRWISE, false, true, HAS_INVALID>(A, bchunk, lutb, pa, extra, extra, vb0, max(k0, kg), nsteps, n0,
                                                     lane, c, best_gs, best_o, vb_gs, vb_o);
      } else {
        dp_steps<PAIRWISE, true, false, HAS_INVALID>(A, bchunk, lutb, pa, extra, extra, vb0, k0, max(k0, kg), n0, lane,
                                                     c, best_gs, best_o, vb_gs, vb_o);
        dp_steps<PAIRWISE, false, false, HAS_INVALID>(A, bchunk, lutb, pa, extra, extra, vb0, max(k0, kg), nsteps, n0,
                                                      lane, c, best_gs, best_o, vb_gs, vb_o);
      }
      best[t * kChunk + lane] = make_float4(best_gs, best_o, __int_as_float(vb_gs), __int_as_float(vb_o));
    }
  }
  __syncthreads();
  float4 *out = dp_out + (size_t)gcol * H;
  for (int i = tid; i < H; i += kDpThreads) out[i] = best[i];
}

size_t dp_smem_bytes(const KParams &p, bool pairwise) {
  const int nchunks = (p.rows + kChunk - 1) / kChunk;
  size_t b = (size_t)2 * kChunk * kBStride * 4 + (size_t)kChunk * nchunks * sizeof(float4);
  b += pairwise ? (size_t)kQsRows * kDynWords * 4 : (size_t)(p.rows + 1) * 4;
  return (b + 15) & ~(size_t)15;
}

}  // namespace

template <bool PAIRWISE, bool HAS_INVALID>
static void launch_dp_variant(const KParams &p, const BatchBuffers &b, int ncolumns, size_t smem, cudaStream_t s) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(dp_kernel<PAIRWISE, HAS_INVALID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  dp_kernel<PAIRWISE, HAS_INVALID><<<ncolumns, kDpThreads, smem, s>>>(
      b.records, b.object_lut, b.stat, b.pm, b.vhor, b.object_disparity_range, b.inverse_height, b.dp, p);
}

void launch_dp(const KParams &p, const BatchBuffers &b, int nframes, bool pairwise, cudaStream_t s) {
  const int ncolumns = nframes * p.realcols;
  const size_t smem = dp_smem_bytes(p, pairwise);
  const bool has_invalid = p.invalid_disparity >= 0.0f;  // ComputeMean's two modes (StixelsKernels.cu:47-60)
  if (pairwise) {
    if (has_invalid) launch_dp_variant<true, true>(p, b, ncolumns, smem, s);
    else launch_dp_variant<true, false>(p, b, ncolumns, smem, s);
  } else {
    if (has_invalid) launch_dp_variant<false, true>(p, b, ncolumns, smem, s);
    else launch_dp_variant<false, false>(p, b, ncolumns, smem, s);
  }
  g_launch_count++;
}

}  // namespace isx
