// Segmentation ingest: the layout transform the reference bakes into its ONNX
// export, `FlipAndPad` (tools/CNN_training/models/wrappers.py:35-61), as one device pass:
//   CNN output  float [n][21][Hs][Ws]   (19 x -log softmax + 2 offsets, rows top-down, Hs = H/8, Ws = W/8)
//   -> permute(0,3,1,2), rows flipped (index_select 97..0), zero-padded Hs -> Hs2, x *= 8, x.int()
//   -> int32 [n][C][21][Hs2], the tensor Stixels::SetSegmentation / Compute(d_segmentation_local) take
//      (kernel indexing StixelsKernels.cu:393-405, 462-468).
// With column_step < 8 (BASELINE config 4) stixel column c takes CNN column c*step/8.
// HBM-bound transpose: 32 x 32 tiles through shared memory, reads coalesced along Ws, writes along Hs2.
#include "kernels.h"

namespace isx {
namespace {

constexpr int kIngestRows = 128;  // output rows (q) per CTA: four 32 x 32 tiles in flight per thread block

__global__ void __launch_bounds__(256)
flip_and_pad_kernel(const float *__restrict__ cnn, int32_t *__restrict__ seg, int channels, int hs, int ws, int hs2,
                    int realcols, int column_step) {
  __shared__ float tile[kIngestRows][33];
  const int ch = blockIdx.z % channels, f = blockIdx.z / channels;
  const int q0 = blockIdx.x * kIngestRows;   // output row block (flipped rows, 0 = bottom)
  const int c0 = blockIdx.y * 32;            // CNN column block
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float *src = cnn + ((size_t)f * channels + ch) * hs * ws;
  // tile[i][j] = value of flipped row q0 + i, CNN column c0 + j  (zero in the padding q >= hs);
  // all 16 loads of a thread are issued before the barrier
#pragma unroll
  for (int i = ty; i < kIngestRows; i += 8) {
    const int q = q0 + i, c = c0 + tx;
    float v = 0.0f;
    if (q < hs && c < ws) v = __ldg(src + (size_t)(hs - 1 - q) * ws + c);
    tile[i][tx] = v;
  }
  __syncthreads();
  // stixel columns fed by this block of CNN columns: c*step/8 in [c0, c0 + 32)
  const int per = kDownsample / column_step;  // stixel columns per CNN column (1 for width 8, 2 for width 4)
  for (int j = ty; j < 32 * per; j += 8) {
    const int col = c0 * per + j;
    if (col >= realcols) continue;
    int32_t *dst = seg + (((size_t)f * realcols + col) * channels + ch) * hs2;
#pragma unroll
    for (int i = tx; i < kIngestRows; i += 32) {
      const int q = q0 + i;
      // x *= 8; x.int(): fp32 multiply, then truncation toward zero (wrappers.py:59-60)
      if (q < hs2) dst[q] = (int32_t)__fmul_rn(tile[i][j / per], 8.0f);
    }
  }
}

}  // namespace

void launch_flip_and_pad(const KParams &p, const float *cnn, int32_t *seg, int nframes, int hs, int ws,
                         cudaStream_t s) {
  dim3 grid((p.hs2 + kIngestRows - 1) / kIngestRows, (ws + 31) / 32, nframes * p.n_channels);
  flip_and_pad_kernel<<<grid, 256, 0, s>>>(cnn, seg, p.n_channels, hs, ws, p.hs2, p.realcols, p.column_step);
  g_launch_count++;
}

}  // namespace isx
