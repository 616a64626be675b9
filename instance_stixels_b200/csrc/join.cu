// Column join: horizontal mean (or median) of `column_step` disparities per
// stixel column, transposed and flipped to bottom-up column-major.
// Replaces JoinColumns (InstanceStixels/src/StixelsKernels.cu:980-1095,
// launch Stixels.cu:509-511).
//
// HBM-bound pass: every input byte is read once with 128-bit coalesced loads,
// the 32x32 result tile goes through shared memory so that the transposed
// store is coalesced too (the reference stores with stride H*4 bytes).
#include "kernels.h"

namespace isx {

namespace {

constexpr int kTile = 32;

// Mean of the valid pixels, summed left to right, times MUFU.RCP(count)
// (reference SASS: FADD.FTZ chain, I2FP, MUFU.RCP, FMUL.FTZ).
template <int STEP>
__device__ __forceinline__ float join_mean(const float *px, int step, float invalid) {
  const int n = STEP > 0 ? STEP : step;
  float sum = 0.0f;
  if (invalid >= 0.0f) {
    int bad = 0;
#pragma unroll
    for (int i = 0; i < n; i++) {
      const float d = px[i];
      if (d != invalid) sum = fadd(sum, d);
      else bad++;
    }
    if (bad == n) return invalid;
    return fmul(rcp_approx((float)(n - bad)), sum);
  }
#pragma unroll
  for (int i = 0; i < n; i++) sum = fadd(sum, px[i]);
  return fmul(rcp_approx((float)n), sum);
}

// Median variant (StixelsKernels.cu:991-1056): order statistics are exact, the
// even case averages the two middle values ((a+b)/2.0f == (a+b)*0.5f).
__device__ float join_median(const float *px, int step, float invalid) {
  float v[16];
  int n = 0;
  for (int i = 0; i < step && i < 16; i++) {
    if (invalid >= 0.0f && px[i] == invalid) continue;
    v[n++] = px[i];
  }
  if (n == 0) return invalid;
  for (int i = 0; i < n / 2 + 1; i++) {
    int m = i;
    for (int j = i + 1; j < n; j++)
      if (v[j] < v[m]) m = j;
    const float t = v[i];
    v[i] = v[m];
    v[m] = t;
  }
  float med = v[n / 2];
  if (n % 2 == 0) med = fmul(fadd(med, v[n / 2 - 1]), 0.5f);
  return med;
}

template <int STEP>
__global__ void __launch_bounds__(256) join_columns_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                           KParams p) {
  __shared__ float tile[kTile][kTile + 1];  // [column][row]
  const int H = p.rows, W = p.cols, C = p.realcols;
  const int step = p.column_step;
  const int col0 = blockIdx.x * kTile, row0 = blockIdx.y * kTile;
  const float *img = in + (size_t)blockIdx.z * H * W;
  float *dst = out + (size_t)blockIdx.z * C * H;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

  for (int r = ty; r < kTile; r += 8) {
    const int row = row0 + r, col = col0 + tx;
    if (row < H && col < C) {
      const float *src = img + (size_t)row * W + p.width_margin + (size_t)col * step;
      float px[16];
      if constexpr (STEP == 8) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        px[0] = a.x; px[1] = a.y; px[2] = a.z; px[3] = a.w;
        px[4] = b.x; px[5] = b.y; px[6] = b.z; px[7] = b.w;
      } else if constexpr (STEP == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
        px[0] = a.x; px[1] = a.y; px[2] = a.z; px[3] = a.w;
      } else {
        for (int i = 0; i < step && i < 16; i++) px[i] = __ldg(src + i);
      }
      tile[tx][r] = p.median_join ? join_median(px, step, p.invalid_disparity)
                                  : join_mean<STEP>(px, step, p.invalid_disparity);
    }
  }
  __syncthreads();
  // out[col][H-1-row]: lanes walk the rows of one column -> contiguous 128 B.
  for (int c = ty; c < kTile; c += 8) {
    const int row = row0 + tx, col = col0 + c;
    if (row < H && col < C) dst[(size_t)col * H + (H - 1 - row)] = tile[c][tx];
  }
}

}  // namespace

void launch_join_columns(const KParams &p, const BatchBuffers &b, int nframes, cudaStream_t s) {
  dim3 grid((p.realcols + kTile - 1) / kTile, (p.rows + kTile - 1) / kTile, nframes);
  const bool vec_ok = (p.width_margin % 4 == 0) && (p.cols % 4 == 0);
  if (p.column_step == 8 && vec_ok)
    join_columns_kernel<8><<<grid, 256, 0, s>>>(b.disparity, b.joined, p);
  else if (p.column_step == 4 && vec_ok)
    join_columns_kernel<4><<<grid, 256, 0, s>>>(b.disparity, b.joined, p);
  else
    join_columns_kernel<0><<<grid, 256, 0, s>>>(b.disparity, b.joined, p);
  g_launch_count++;
}

}  // namespace isx
