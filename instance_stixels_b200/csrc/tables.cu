// Per-column cost tables ("table build").
//
//   frame_tables_kernel  : per frame, the vB-only parts of the pairwise
//                          transition priors (StixelsKernels.cu:40-42, 98-144,
//                          173-187) -- they depend on the road model only.
//   column_tables_kernel : per (frame, column): per-row data costs
//                          (:371-446), the prefix sums the reference computes
//                          with in-kernel Blelloch scans (:452-469,
//                          StixelsKernels.h:73-103) and the object-cost LUT of
//                          ComputeObjectLUT (:236-296, 959-978).
//
// Float prefix sums are order-sensitive; both scan orders of the reference are
// reproduced exactly (see blelloch_prefix_warp / object_lut_rows below).
#include "kernels.h"

namespace isx {

namespace {

// ---------------------------------------------------------------------------
// Static transition record S[vB], vB in [1, H):   (p = vB - 1)
//  0 pc      = ln(H - vB)                       GetPriorCost            (:40-42)
//  1 t1_hi   = max(gf[p],0) + eps               ObjectFromGround cases  (:120-144)
//  2 t1_lo   = max(gf[p],0) - eps
//  3 tr1_hi  4 tr1_mid  5 tr1_lo                its three finite values
//  6 sky_fg  = gf[p] < 1 ? pc : inf             SkyFromGround           (:98-106)
//  7 oo_base = pc + (p < vhor ? -ln .7 : ln 2)  ObjectFromObject prefix (:151-152)
//  8 tr3     = pc + ln(D - eps)                 ObjectFromSky           (:173-183)
//  9 g_prev  = -ln .3 + pc                      GetPriorCostGround      (:185-187)
// 10 sky_fo  = ln 2 + pc                        SkyFromObject           (:88-96)
// ---------------------------------------------------------------------------
__global__ void frame_tables_kernel(const float *__restrict__ ground, const int *__restrict__ vhor_arr,
                                    float *__restrict__ stat, KParams p) {
  const int f = blockIdx.y;
  const int vB = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = p.rows;
  if (vB >= H) return;
  float *S = stat + ((size_t)f * H + vB) * kStatWords;
  if (vB == 0) {
    for (int i = 0; i < kStatWords; i++) S[i] = 0.0f;
    return;
  }
  const float *gf = ground + (size_t)f * 3 * H;
  const int vhor = vhor_arr[f];
  const int pv = vB - 1;
  const float eps = p.epsilon;

  // FFMA(lg2(H-vB), ln2, -0): the -__logf(1.0f) term folds to -0.
  const float pc = ffma(lg2_approx((float)(H - vB)), kLn2, -0.0f);
  const float fnp = clamp_neg(gf[pv]);
  const float t1_hi = fadd(fnp, eps);
  const float t1_lo = fsub(fnp, eps);
  const float og_base = fadd(pc, kNegLog07);
  const float tr1_hi = fadd(neg_log_div(p.pgrav, fsub(fadd(-fnp, p.max_disf), eps)), og_base);
  const float tr1_mid = fadd(neg_log_div(fsub(fadd(-p.pgrav, 1.0f), p.pblg), fadd(eps, eps)), og_base);
  const float tr1_lo = fadd(neg_log_div(p.pblg, t1_lo), og_base);
  const float sky_fg = (gf[pv] < 1.0f) ? pc : inf_f();
  const float oo_base = fadd(pc, (pv < vhor) ? kNegLog07 : kLn2);
  const float tr3 = fadd(pc, ffma(lg2_approx(fsub(p.max_disf, eps)), kLn2, -0.0f));
  const float g_prev = ffma(-kLn2, kLg2_03, pc);
  const float sky_fo = ffma(kLn2, 1.0f, pc);
  S[0] = pc;  S[1] = t1_hi;  S[2] = t1_lo;  S[3] = tr1_hi;  S[4] = tr1_mid;  S[5] = tr1_lo;
  S[6] = sky_fg;  S[7] = oo_base;  S[8] = tr3;  S[9] = g_prev;  S[10] = sky_fo;  S[11] = 0.0f;
}

// ---------------------------------------------------------------------------
// Blelloch-order exclusive prefix sum of e[0..H) (H <= 1024), one warp.
//
// The reference scans in place with the classic work-efficient tree
// (StixelsKernels.h:73-103): block sums are pairwise trees U(.), and ps[i]
// folds, from the most to the least significant set bit of i,
//      prefix = prefix + U(aligned block that bit selects).
// A butterfly (shfl_xor) reproduces U(.) of every aligned block; the value a
// lane receives at level b while its bit b is set is exactly the U(left
// sibling) the down-sweep adds.  Levels 0-4 live inside a 32-row chunk,
// levels 5-9 across the (<= 32) chunk totals.
// ---------------------------------------------------------------------------
__device__ void blelloch_prefix_warp(const float *e, int H, float *ps) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int nchunks = (H + 31) >> 5;
  float total_k = 0.0f;  // lane k: U(chunk k)
  for (int k = 0; k < nchunks; k++) {
    const int v = (k << 5) + lane;
    float x = v < H ? e[v] : 0.0f;
#pragma unroll
    for (int b = 0; b < 5; b++) x = fadd(x, __shfl_xor_sync(full, x, 1 << b));
    if (lane == k) total_k = x;
  }
  float y = total_k, left[5];
#pragma unroll
  for (int b = 0; b < 5; b++) {
    const float o = __shfl_xor_sync(full, y, 1 << b);
    left[b] = o;
    y = fadd(y, o);
  }
  float hi = 0.0f;  // lane k: prefix of chunk k's first row
#pragma unroll
  for (int b = 4; b >= 0; b--)
    if ((lane >> b) & 1) hi = fadd(hi, left[b]);
  const int kmax = H >> 5;  // chunk index of row H itself
  for (int k = 0; k <= kmax; k++) {
    const int v = (k << 5) + lane;
    float x = v < H ? e[v] : 0.0f;
    float lo[5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
      const float o = __shfl_xor_sync(full, x, 1 << b);
      lo[b] = o;
      x = fadd(x, o);
    }
    // k == 32 only for H == 1024, lane 0: the root total.
    float pre = (k < 32) ? __shfl_sync(full, hi, k & 31) : y;
#pragma unroll
    for (int b = 4; b >= 0; b--)
      if ((lane >> b) & 1) pre = fadd(pre, lo[b]);
    if (v <= H) ps[v] = pre;
  }
}

// Exact integer exclusive prefix sums (any order is bit-identical).
template <typename T>
__device__ void exact_prefix_warp(const T *e, int n, T *ps) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  T carry = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    T x = i < n ? e[i] : (T)0;
    T incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T o = __shfl_up_sync(full, incl, d);
      if (lane >= d) incl += o;
    }
    if (i < n) ps[i] = carry + incl - x;
    carry += __shfl_sync(full, incl, 31);
  }
  if (lane == 0) ps[n] = carry;
}

constexpr int kTabThreads = 256;

// Dynamic shared memory carve-up (bytes), Hp = H + 1 rounded up to 4.
// The per-row terms are scanned in place (every scan reads e[v] and writes ps[v] from the same lane), which keeps
// the CTA at 71 KB so that three of them share an SM.
struct TabSmem {
  int Hp, nq;
  size_t off_e[4], off_segps, off_seg, off_i64e, total;
  __host__ __device__ TabSmem(int H, int hs2) {
    Hp = (H + 1 + 3) & ~3;
    nq = H / 8 + 1;  // prefix entries per 1/8-res channel (index v>>3 for v <= H)
    size_t o = 0;
    off_i64e = o; o += (size_t)Hp * 8 * 4;
    for (int i = 0; i < 4; i++) { off_e[i] = o; o += (size_t)Hp * 4; }
    off_seg = o; o += (size_t)21 * (nq + 1) * 4;
    off_segps = o; o += (size_t)21 * (nq + 1) * 4;
    total = (o + 15) & ~(size_t)15;
    (void)hs2;
  }
};

__global__ void __launch_bounds__(kTabThreads, 3)
column_tables_kernel(const float *__restrict__ joined, const int32_t *__restrict__ segmentation,
                     const float *__restrict__ ground, const int *__restrict__ vhor_arr,
                     uint32_t *__restrict__ records_b, int *__restrict__ error_flag, int *__restrict__ col_flags,
                     KParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int H = p.rows, C = p.realcols;
  const int col = blockIdx.x, f = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TabSmem L(H, p.hs2);
  float *e_disp = reinterpret_cast<float *>(smem + L.off_e[0]);
  float *e_valid = reinterpret_cast<float *>(smem + L.off_e[1]);
  float *e_ground = reinterpret_cast<float *>(smem + L.off_e[2]);
  float *e_sky = reinterpret_cast<float *>(smem + L.off_e[3]);
  float *ps_f[4] = {e_disp, e_valid, e_ground, e_sky};                   // scanned in place
  long long *e_i64 = reinterpret_cast<long long *>(smem + L.off_i64e);    // [4][Hp]
  long long *ps_i64 = e_i64;                                              // scanned in place
  int *seg_s = reinterpret_cast<int *>(smem + L.off_seg);                 // [21][nq+1]
  int *seg_ps = reinterpret_cast<int *>(smem + L.off_segps);              // [21][nq+1]
  const int nq = L.nq, segld = nq + 1;

  const float *d_col = joined + ((size_t)f * C + col) * H;
  const int32_t *seg_col = segmentation + ((size_t)f * C + col) * p.n_channels * p.hs2;
  const float *gf = ground + (size_t)f * 3 * H;
  const float *norm_g = gf + H, *inv_g = gf + 2 * H;
  const int vhor = vhor_arr[f];
  const float invalid = p.invalid_disparity;
  const int K = p.n_classes;  // 19; channel K = y offsets, K+1 = x offsets

  // ---- 1/8-resolution channels: the 19 classes, and as a 20th the sum of the two squared offsets (:411-416; the
  //      DP only ever uses the sum, ComputeNonInstanceOffsetCost :62-70) ----
  bool negative_class_value = false;
  for (int i = tid; i < 20 * nq; i += kTabThreads) {
    const int c = i / nq, q = i - c * nq;
    int v = 0;
    if (c < K) {
      if (c < p.n_channels && q < p.hs2) v = seg_col[(size_t)c * p.hs2 + q];
      negative_class_value |= v < 0 || v > (1 << 20);   // the int32 prefix sums over <= 1032 rows cannot wrap
    } else {
      int oy = 0, ox = 0;
      if (q < p.hs2 && K < p.n_channels) oy = seg_col[(size_t)K * p.hs2 + q];
      if (q < p.hs2 && K + 1 < p.n_channels) ox = seg_col[(size_t)(K + 1) * p.hs2 + q];
      oy = oy * oy;
      ox = ox * ox;
      negative_class_value |= oy < 0 || oy > (1 << 19) || ox < 0 || ox > (1 << 19);   // their sum: likewise
      v = oy + ox;
    }
    seg_s[c * segld + q] = v;
  }
  // ---- per-row terms (:371-446) ----
  for (int v = tid; v < H; v += kTabThreads) {
    const float d = d_col[v];
    if (invalid >= 0.0f) {
      const int va = d != invalid;
      e_valid[v] = (float)va;
      e_disp[v] = fmul((float)va, d);
    } else {
      e_valid[v] = 1.0f;  // unused by the reference in this mode (ComputeMean divides by the height)
      e_disp[v] = d;
    }
    // sky_lut: GetDataCostSky (:201-215), zero below the horizon (:424-433)
    float sky = 0.0f;
    if (v >= vhor) {
      sky = p.pnexists_given_sky_log;
      if (d != invalid) {
        const float g = ffma(fmul(d, d), p.inv_sigma2_sky, p.normalization_sky);
        sky = fadd(fmin_(g, p.puniform_sky), p.nopnexists_given_sky_log);
      }
    }
    e_sky[v] = sky;
    // ground_lut: GetDataCostGround (:217-234), +inf at/above the horizon (:437-446)
    float grd = inf_f();
    if (v < vhor) {
      grd = p.pnexists_given_ground_log;
      if (d != invalid) {
        const float diff = fsub(d, gf[v]);
        const float g = ffma(fmul(diff, diff), inv_g[v], norm_g[v]);
        grd = fadd(fmin_(g, p.puniform), p.nopnexists_given_ground_log);
      }
    }
    e_ground[v] = grd;
    // instance means (:400-409): C++ double arithmetic truncated toward zero
    const int q = v / kDownsample;
    const int off_y = (q < p.hs2 && K < p.n_channels) ? seg_col[(size_t)K * p.hs2 + q] : 0;
    const int off_x = (q < p.hs2 && K + 1 < p.n_channels) ? seg_col[(size_t)(K + 1) * p.hs2 + q] : 0;
    const long long mx = (long long)((p.column_step * col + 0.5 * (p.column_step - 1.0)) + off_x + 0.5);
    const long long my = (long long)(v - off_y + 0.5);
    e_i64[0 * L.Hp + v] = mx;
    e_i64[1 * L.Hp + v] = my;
    e_i64[2 * L.Hp + v] = mx * mx;
    e_i64[3 * L.Hp + v] = my * my;
  }
  // The pruning DP kernels bound a segment's class sums from below by those of a shorter one, which needs
  // non-negative values whose int32 prefix sums do not wrap (they are trunc(8 * -log softmax) and squared pixel
  // offsets, wrappers.py:50-61); a column that breaks this is flagged and walked exhaustively.
  const int any_negative = __syncthreads_or(negative_class_value ? 1 : 0);
  if (tid == 0) col_flags[(size_t)f * C + col] = any_negative;

  // ---- prefix sums: warps 0-3 float (Blelloch order), 4-7 int64, all: 1/8-res ints ----
  if (warp < 4) {
    const float *src = warp == 0 ? e_disp : warp == 1 ? e_valid : warp == 2 ? e_ground : e_sky;
    blelloch_prefix_warp(src, H, ps_f[warp]);
  } else {
    exact_prefix_warp<long long>(e_i64 + (warp - 4) * L.Hp, H, ps_i64 + (warp - 4) * L.Hp);
  }
  for (int c = warp; c < 20; c += kTabThreads / 32) exact_prefix_warp<int>(seg_s + c * segld, nq, seg_ps + c * segld);
  __syncthreads();

  // ---- records, v in [0, H]: one 128-byte row per v (common.cuh).  Eight lanes share a row, lane g writes words
  //      4g .. 4g+3 as one 16-byte store, so a warp instruction stores 512 contiguous bytes (four full lines).
  //      Groups 0-4 are the 20 integer words (one formula), 5-7 the float words. ----
  uint4 *recb_col = reinterpret_cast<uint4 *>(records_b + ((size_t)f * C + col) * (size_t)p.rec_stride * kRecBWords);
  bool out_of_range = false;
  const long long lim1 = 1ll << 24, lim2 = 1ll << (24 + kSqSplitBits);
  const long long lomask = (1ll << kSqSplitBits) - 1;
  for (int idx = tid; idx < (H + 1) * (kRecBWords / 4); idx += kTabThreads) {
    const int v = idx >> 3, g = idx & 7;
    uint4 w;
    if (g < 5) {
      // P_c(v) = 8*ps[q] + seg[q]*r  (Cityscapes.h:28-42); word 19 = the same for the summed squared offsets
      const int q = v >> 3, r = v & 7;
      const int *ps = seg_ps + (4 * g) * segld + q, *sg = seg_s + (4 * g) * segld + q;
      w.x = (uint32_t)(ps[0] * kDownsample + sg[0] * r);
      w.y = (uint32_t)(ps[segld] * kDownsample + sg[segld] * r);
      w.z = (uint32_t)(ps[2 * segld] * kDownsample + sg[2 * segld] * r);
      w.w = (uint32_t)(ps[3 * segld] * kDownsample + sg[3 * segld] * r);
    } else if (g == 5) {
      // instance-mean sums as exactly representable floats (see common.cuh): words 20 .. 23
      const long long smx = ps_i64[0 * L.Hp + v], smy = ps_i64[1 * L.Hp + v], smx2 = ps_i64[2 * L.Hp + v];
      out_of_range |= smx <= -lim1 || smx >= lim1 || smy <= -lim1 || smy >= lim1 || smx2 >= lim2;
      w.x = __float_as_uint((float)smx);
      w.y = __float_as_uint((float)smy);
      w.z = __float_as_uint((float)(smx2 & ~lomask));
      w.w = __float_as_uint((float)(smx2 & lomask));
    } else if (g == 6) {
      // words 24 .. 27
      const long long smy2 = ps_i64[3 * L.Hp + v];
      out_of_range |= smy2 >= lim2;
      w.x = __float_as_uint((float)(smy2 & ~lomask));
      w.y = __float_as_uint((float)(smy2 & lomask));
      w.z = __float_as_uint(ps_f[0][v]);
      w.w = __float_as_uint(ps_f[1][v]);
    } else {
      // words 28 .. 31
      w.x = __float_as_uint(ps_f[2][v]);
      w.y = __float_as_uint(ps_f[3][v]);
      w.z = w.w = 0u;
    }
    recb_col[idx] = w;
  }
  if (out_of_range) atomicOr(error_flag + f, kErrOffsetRange);
  (void)lane;
}

// ---------------------------------------------------------------------------
// Object-cost LUT (ComputeObjectLUT, StixelsKernels.cu:236-296, 959-978):
//   LUT[fn][v+1] = sum_{r<=v} obj_cost_lut[fn][(int)d[r]]
// in the reference's summation order (warp_prefix_sum / ComputePrefixSumWarp2): 32-row chunks, the
// running total joins row 0 of the chunk BEFORE a Kogge-Stone scan (lane L: x += x[L-j], j = 1,2,4,8,16).
//
// The reference runs that scan across the lanes of a warp (5 shuffles per chunk and fn row), which made
// the first version of this kernel LSU-bound.  Here one THREAD owns a whole chunk of one fn row: its 32
// values sit in registers, the Kogge-Stone network is 129 FADDs in exactly the reference's association,
// and the chunk carry is a register.  A warp covers 32 consecutive fn rows (lane = fn), reads the
// transposed cost table [dis][fn] (one 128-byte line per row of the chunk) and transposes its 32 x 32
// result tile through shared memory so that every store is a full 128-byte line of one fn row.
// ---------------------------------------------------------------------------
constexpr int kLutWarps = 4;
constexpr int kLutThreads = kLutWarps * 32;

// 32-bit address arithmetic for the two address streams of the kernel (the same trick as the DP's LUT gathers,
// dp.cu): the transposed cost table (64 KB) and one column of the object LUT (< 4 GB, never straddling a multiple of
// 4 GB: lut_column_address) each keep ONE upper address word, so an address costs one IMAD / IADD instead of the
// five instructions of a 64-bit multiply-add (the first version spent 8 instructions per gather, 5 per store).
__device__ __forceinline__ float ldg_lo_hi(unsigned lo, unsigned hi) {
  float r;
  asm("{\n\t.reg .b64 a;\n\tmov.b64 a, {%1, %2};\n\tld.global.nc.f32 %0, [a];\n\t}" : "=f"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void stg_lo_hi(unsigned lo, unsigned hi, float v) {
  asm volatile("{\n\t.reg .b64 a;\n\tmov.b64 a, {%0, %1};\n\tst.global.f32 [a], %2;\n\t}" ::"r"(lo), "r"(hi), "f"(v)
               : "memory");
}

__global__ void __launch_bounds__(kLutThreads)
object_lut_kernel(const float *__restrict__ joined, const float *__restrict__ cost_t,
                  float *__restrict__ object_lut, KParams p) {
  __shared__ __align__(16) uint8_t dis_s[1024 + 32];
  __shared__ float tile[kLutWarps][32][33];
  const int H = p.rows, C = p.realcols, D = p.max_dis;
  const int Dp = (D + 31) & ~31;
  const int col = blockIdx.x, f = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *d_col = joined + ((size_t)f * C + col) * H;
  const int Hc = (H + 31) & ~31;
  for (int v = threadIdx.x; v < Hc; v += kLutThreads) {
    int di = 0;  // rows past H count as disparity 0, like the reference's padded loop (:262-272)
    if (v < H) {
      di = (int)d_col[v];  // (int) d as LUT index (:246-248)
      di = di < 0 ? 0 : (di >= D ? D - 1 : di);
    }
    dis_s[v] = (uint8_t)di;
  }
  __syncthreads();
  const int fn0 = (blockIdx.y * kLutWarps + warp) * 32;
  if (fn0 >= D) return;
  const unsigned long long lut_addr = lut_column_address((unsigned long long)object_lut, (size_t)f * C + col,
                                                         p.lut_cols, (size_t)D * p.lut_stride * 4);
  const unsigned lut_hi = (unsigned)(lut_addr >> 32);
  const unsigned stride4 = (unsigned)p.lut_stride * 4u;
  // column fn of the transposed table (padded to 32 columns, always in range); the table does not straddle 4 GB
  // (isx_initialize checks), so its upper address word is that of its first byte
  const unsigned long long cost_addr = (unsigned long long)(cost_t + fn0 + lane);
  const unsigned cost_lo = (unsigned)cost_addr, cost_hi = (unsigned)(cost_addr >> 32);
  const unsigned row4 = (unsigned)Dp * 4u;   // bytes per disparity row of the transposed table
  float(*tl)[33] = tile[warp];
  const bool full = fn0 + 32 <= D;           // every row of the 32 x 32 tile exists (D is a multiple of 32)
  float carry = 0.0f;
  for (int i = 0; i < Hc; i += 32) {
    float x[32];
    const uint4 *dq = reinterpret_cast<const uint4 *>(dis_s + i);
    const uint4 d0 = dq[0], d1 = dq[1];
    const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int r = 0; r < 32; r++) {
      const unsigned dis = __byte_perm(dw[r >> 2], 0u, 0x4440u | (unsigned)(r & 3));   // byte r & 3, zero-extended
      x[r] = ldg_lo_hi(dis * row4 + cost_lo, cost_hi);
    }
    x[0] = fadd(x[0], carry);
#pragma unroll
    for (int j = 1; j < 32; j <<= 1) {
#pragma unroll
      for (int L = 31; L >= j; L--) x[L] = fadd(x[L], x[L - j]);
    }
    carry = x[31];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; r++) tl[lane][r] = x[r];
    __syncwarp();
    const int v = i + lane;
    if (v < H) {
      unsigned lo = (unsigned)lut_addr + (unsigned)fn0 * stride4 + 4u * (unsigned)v;   // out[v] = LUT[fn][v + 1]
      if (full) {
#pragma unroll
        for (int r = 0; r < 32; r++) {
          stg_lo_hi(lo, lut_hi, tl[r][lane]);
          lo += stride4;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 32; r++) {
          if (fn0 + r < D) stg_lo_hi(lo, lut_hi, tl[r][lane]);
          lo += stride4;
        }
      }
    }
  }
}

}  // namespace

void launch_frame_tables(const KParams &p, const BatchBuffers &b, int nframes, cudaStream_t s) {
  dim3 grid((p.rows + 255) / 256, nframes);
  frame_tables_kernel<<<grid, 256, 0, s>>>(b.ground, b.vhor, b.stat, p);
  g_launch_count++;
}

static size_t tab_smem_bytes(const KParams &p) { return TabSmem(p.rows, p.hs2).total; }

void launch_column_tables(const KParams &p, const BatchBuffers &b, int nframes, cudaStream_t s) {
  dim3 grid(p.realcols, nframes);
  const size_t smem = tab_smem_bytes(p);
  static SmemOptIn optin;
  opt_in_smem(column_tables_kernel, optin);
  column_tables_kernel<<<grid, kTabThreads, smem, s>>>(b.joined, b.segmentation, b.ground, b.vhor, b.records_b,
                                                        b.error_flag, b.col_flags, p);
  dim3 lgrid(p.realcols, (p.max_dis + 32 * kLutWarps - 1) / (32 * kLutWarps), nframes);
  object_lut_kernel<<<lgrid, kLutThreads, 0, s>>>(b.joined, b.obj_cost_lut_t, b.object_lut, p);
  g_launch_count += 2;
}

}  // namespace isx
