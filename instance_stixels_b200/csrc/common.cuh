// Shared declarations of the sm_100a stixel kernels.
//
// Arithmetic contract: the reference is built with `-O3 --use_fast_math`
// (CMakeLists.txt:138-141), i.e. every fp32 op is .ftz, division is
// x * MUFU.RCP(y), __logf(x) is MUFU.LG2(x) * ln2 and nvcc contracts a*b+c
// into FFMA.  Results depend on those exact shapes, so the kernels here spell
// every float operation with an explicit intrinsic (fmul/fadd/ffma/rcp/lg2)
// in the order the reference's sm_100a SASS performs it;
// this translation unit is compiled with -ftz=true -fmad=false so nothing is
// contracted or reordered behind our back.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace isx {

constexpr int kWarp = 32;
constexpr int kMaxSections = 200;      // configuration.h:32
constexpr int kDownsample = 8;         // configuration.h:31
constexpr int kInstanceClasses = 8;    // Stixels.cu:47
constexpr int kFirstInstanceClass = 11;  // StixelsKernels.cu:926
constexpr int kSkyClass = 10;          // Cityscapes.h:117
constexpr int GROUND = 0, OBJECT = 1, SKY = 2;  // types.h:22-24

constexpr float kLn2 = 0.69314718246459960938f;        // 0x3f317218, fast-math __logf scale
constexpr float kNegLog07 = 0.35667496919631958008f;   // -logf(0.7f) folded by nvcc (StixelsKernels.cu:124,151)
constexpr float kLg2_03 = -1.7369655370712280273f;     // lg2(0.3f) folded by nvcc (StixelsKernels.cu:186)

// ---- column tables: prefix sums over rows [0, v) of one stixel column ----
// One record of 32 words (30 used) per row v in [0, H]: records_b[column][v][32], 128-byte rows, rec_stride rows
// per column.  Both access patterns of the DP move whole lines:
//   * "B side": the 32 rows of a vB chunk are one contiguous 4 KB block for cp.async.bulk into shared memory,
//   * "A side": lane l reads the row of vT + 1 = a + l + 1 (its own 128-byte line), once per tile.
// (Round 1 kept a second, word-major copy for the A side: 32 MB per frame of extra writes for a load that the
// pruning kernels issue once per tile.)
// The row count per column is the constant kRecStride (H <= 1024) whatever the image height.
constexpr int kRecStride = 1056;  // 1024 + 1 rounded up to 32
constexpr int kRecWords = 30;
constexpr int kRecSeg = 0;     // 19 words: full-resolution prefix of class c, exact int32
                               //   P_c(v) = 8*ps_c[v/8] + seg_c[v/8]*(v%8)   (Cityscapes.h:28-42)
constexpr int kRecOff = 19;    // int32 prefix of squared offsets x^2+y^2 (StixelsKernels.cu:62-70, 411-416)
// The reference keeps the four instance-mean sums as int64 and converts the DIFFERENCE of two
// prefixes to float (I2F.S64, round to nearest; :72-86, 611-616).  Here the prefixes are stored as
// exactly representable floats so that the same value comes out of plain FADDs:
//   sum(mx), sum(my): |P| < 2^24, float(P) exact, float(P_a) - float(P_b) exact == float(P_a - P_b);
//   sum(mx^2), sum(my^2): P = hi * 4096 + lo, both parts exact floats; (hi_a - hi_b) and
//   (lo_a - lo_b) are exact and their FADD rounds the exact difference once == I2F.S64(P_a - P_b).
// column_tables_kernel checks the ranges (|P| < 2^24, P2 < 2^36) and raises the error flag otherwise.
constexpr int kRecMx = 20;     // float(sum instance_meansx)                  (StixelsKernels.cu:401-403)
constexpr int kRecMy = 21;     // float(sum instance_meansy)                  (:404-405)
constexpr int kRecMx2Hi = 22;  // float((sum meansx^2 >> 12) << 12)           (:406-407)
constexpr int kRecMx2Lo = 23;  // float(sum meansx^2 & 4095)
constexpr int kRecMy2Hi = 24;  // same for meansy^2                           (:408-409)
constexpr int kRecMy2Lo = 25;
constexpr int kRecDisp = 26;   // float Blelloch-order prefix of valid*d      (:385,455)
constexpr int kRecValid = 27;  // float prefix of valid                       (:384,453)
constexpr int kRecGround = 28; // float Blelloch-order prefix of ground_lut   (:437-446,460)
constexpr int kRecSky = 29;    // float Blelloch-order prefix of sky_lut      (:424-433,461)
constexpr int kSqSplitBits = 12;
constexpr int kRecBWords = 32;  // words per stored row (128 bytes)
// bits of the sticky device error flag
constexpr int kErrSectionOverflow = 1;  // a column produced >= 200 stixels (StixelsKernels.cu:950 asserts)
constexpr int kErrOffsetRange = 2;      // instance-offset sums outside the exact-float range above

// ---- per-frame static transition record S[vB] (pairwise only), 12 floats ----
constexpr int kStatWords = 12;
// ---- per-column dynamic row info Q[vB] (pairwise only), 12 floats, lives in shared memory ----
constexpr int kDynWords = 12;

// Parameters shared by all kernels (subset of StixelParameters, types.h:145-184).
struct KParams {
  int rows;          // H
  int cols;          // W (input image)
  int realcols;      // C = (W - width_margin) / column_step  (Stixels.cu:44)
  int column_step;
  int width_margin;
  int max_dis;       // D
  int hs2;           // rows_power2_segmentation
  int n_classes;     // 19
  int n_channels;    // 21
  int median_join;
  int size_filter;
  int min_pts;
  float eps_cluster;
  float invalid_disparity;
  float max_disf;
  float rows_log, max_dis_log;
  float pnexists_given_sky_log, normalization_sky, inv_sigma2_sky, puniform_sky, nopnexists_given_sky_log;
  float pnexists_given_ground_log, puniform, nopnexists_given_ground_log;
  float pord, epsilon, pgrav, pblg;
  float prior_weight, disparity_weight, segmentation_weight, instance_weight;
  // derived strides
  int rec_stride;    // rows per column of the records: always kRecStride
  int lut_stride;    // floats per fn row of the object LUT (>= H, multiple of 32)
  int lut_cols;      // column slots of the object-LUT buffer (chunk * C), see lut_column_address
  // unary branch and bound (dp.cu)
  float obj_cost_min;  // smallest entry of obj_cost_lut
  float obj_cost_absmax;  // largest |entry| of obj_cost_lut (scale of the rounding of the object LUT's prefix sums)
  int prune_unary;     // 0: walk every chunk (debug / A-B runs, ISX_UNARY_PRUNE=0)
  int prune_pairwise;  // same for the pairwise tile-major walk (ISX_PAIRWISE_PRUNE=0)
};

// ---- pinned float ops (all .ftz through -ftz=true) ----
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fadd_rn(a, -b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// MUFU.RCP / MUFU.LG2, what fast-math division and __logf compile to.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// FMNMX: returns the non-NaN operand, like fminf.
__device__ __forceinline__ float fmin_(float a, float b) { return fminf(a, b); }
// NegFastLogDiv(a, b) = -__logf(a) + __logf(b)  (StixelsKernels.cu:35-38):
//   t = FMUL(lg2 a, ln2); FFMA(lg2 b, ln2, -t)
__device__ __forceinline__ float neg_log_div(float a, float b) {
  const float t = fmul(lg2_approx(a), kLn2);
  return ffma(lg2_approx(b), kLn2, -t);
}
// `x < 0 ? 0 : x` as compiled: FSETP.GEU + FSEL (NaN is kept).
__device__ __forceinline__ float clamp_neg(float x) { return (x < 0.0f) ? 0.0f : x; }

__device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }

// ---- placement of the per-column object LUTs ----
// The DP forms LUT addresses with 32-bit arithmetic (one IMAD per gather, dp.cu), so the LUT of a
// column (col_bytes < 4 GB) must not straddle a multiple of 4 GB.  Column g normally sits at
// base + g * col_bytes; a column that would straddle the m-th 4 GB boundary inside the buffer moves
// to a spare slot behind the regular area (two slots per boundary, one of which cannot straddle).
__host__ __device__ inline size_t lut_buffer_bytes(size_t ncols, size_t col_bytes) {
  const size_t boundaries = ((ncols + 2) * col_bytes >> 32) + 2;
  return (ncols + 2 * boundaries) * col_bytes;
}
__host__ __device__ inline unsigned long long lut_column_address(unsigned long long base, size_t gcol, size_t ncols,
                                                                 size_t col_bytes) {
  const unsigned long long a = base + gcol * col_bytes;
  if (((a ^ (a + col_bytes - 1)) >> 32) == 0) return a;
  const unsigned long long m = ((a + col_bytes - 1) >> 32) - (base >> 32) - 1;
  unsigned long long sp = base + (ncols + 2 * m) * col_bytes;
  if (((sp ^ (sp + col_bytes - 1)) >> 32) != 0) sp += col_bytes;
  return sp;
}

}  // namespace isx

