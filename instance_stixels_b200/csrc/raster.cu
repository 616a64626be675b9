// Stixels -> dense result images on the device.
// Replaces the rectangle drawing of the reference's evaluation tooling
// (tools/visualization/clustering_visualization.py: draw_stixels :397-409 -> label-id and disparity
// result images, draw_instance_masks :118-142 -> instance masks): a stixel (column c, vB..vT) covers
// the inclusive rectangle x in [c*w, c*w + w - 1], y in [H-1-vT, H-1-vB].
//   label image    uint8  [H][W]: Cityscapes label id of the stixel's class (trainId2label[class].id)
//   instance image int32  [H][W]: class*1000 + label for instance stixels with 0 <= label < 1000
//                                 (read_stixel_file :104-111), 0 elsewhere -- one image instead of one
//                                 mask per id; mask(id) == (image == id)
//   disparity image float [H][W]: the stixel's disparity
// One thread per 4 horizontally adjacent pixels: binary search of the row in the column's stixel list
// (sorted top -> bottom), 4/16/16-byte stores, coalesced along x.
#include "kernels.h"

namespace isx {
namespace {

// cityscapesscripts labels.py: trainId 0..18 -> id
__constant__ uint8_t kTrainIdToId[19] = {7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33};

__global__ void __launch_bounds__(256)
instance_table_kernel(const isx_instance *__restrict__ inst, const int *__restrict__ inst_count, int inst_cap,
                      int *__restrict__ table, int realcols) {
  const int f = blockIdx.y;
  const int n = min(inst_count[f], inst_cap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const isx_instance r = inst[(size_t)f * inst_cap + k];
    if (r.label >= 0 && r.label < 1000)
      table[((size_t)f * realcols + r.column) * kMaxSections + r.index] = r.semantic_class * 1000 + r.label;
  }
}

constexpr int kRasterRows = 8;  // image rows per thread: one search, then a walk down the column's stixel list

__global__ void __launch_bounds__(256)
rasterize_kernel(const isx_section *__restrict__ sections, const int *__restrict__ n_sections,
                 const int *__restrict__ table, uint8_t *__restrict__ label_img, int32_t *__restrict__ instance_img,
                 float *__restrict__ disparity_img, int rows, int cols, int realcols, int column_step) {
  const int f = blockIdx.z, y0 = blockIdx.y * kRasterRows;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= cols) return;
  // the (at most two) stixel columns under the four pixels
  const int ca = x0 / column_step, cb = min(x0 + 3, cols - 1) / column_step;
  int split = 4;  // pixels [0, split) belong to column ca, the rest to cb
  if (cb != ca) split = cb * column_step - x0;
  const bool vector_ok = x0 + 3 < cols && (cols & 3) == 0;
  struct Cursor {
    const isx_section *col;
    const int *tab;
    int n, j;
    isx_section s;
  } cur[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int c = k == 0 ? ca : cb;
    Cursor &u = cur[k];
    u.n = 0; u.j = 0; u.col = nullptr; u.tab = nullptr;
    u.s = isx_section{-1, 0, 0, 0.f, 0, 0.f, 0.f, 0.f};
    if (c < realcols && (k == 0 || cb != ca)) {
      u.col = sections + ((size_t)f * realcols + c) * kMaxSections;
      u.tab = table ? table + ((size_t)f * realcols + c) * kMaxSections : nullptr;
      u.n = n_sections[(size_t)f * realcols + c];
      const int v = rows - 1 - y0;
      int lo = 0, hi = u.n - 1;  // vB decreases with the index: smallest index with vB <= v
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u.col[mid].vB <= v) hi = mid; else lo = mid + 1;
      }
      u.j = lo;
      if (u.n > 0) u.s = u.col[lo];
    }
  }
  for (int r = 0; r < kRasterRows; r++) {
    const int y = y0 + r;
    if (y >= rows) break;
    const int v = rows - 1 - y;
    int lab2[2], ins2[2];
    float dis2[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
      Cursor &u = cur[k];
      while (u.col && v < u.s.vB && u.j + 1 < u.n) u.s = u.col[++u.j];   // rows go down, the list goes down
      const bool hit = u.col && u.s.type >= 0 && u.s.vB <= v && v <= u.s.vT;
      const int cls = u.s.semantic_class;
      lab2[k] = (hit && cls >= 0 && cls < 19) ? kTrainIdToId[cls] : 0;
      ins2[k] = (hit && u.tab) ? u.tab[u.j] : 0;
      dis2[k] = hit ? u.s.disparity : 0.0f;
    }
    uint8_t lab[4];
    int ins[4];
    float dis[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int w = k < split ? 0 : 1;
      lab[k] = (uint8_t)lab2[w]; ins[k] = ins2[w]; dis[k] = dis2[w];
    }
    const size_t o = ((size_t)f * rows + y) * cols + x0;
    if (vector_ok) {
      if (label_img) *reinterpret_cast<uchar4 *>(label_img + o) = make_uchar4(lab[0], lab[1], lab[2], lab[3]);
      if (instance_img) *reinterpret_cast<int4 *>(instance_img + o) = make_int4(ins[0], ins[1], ins[2], ins[3]);
      if (disparity_img) *reinterpret_cast<float4 *>(disparity_img + o) = make_float4(dis[0], dis[1], dis[2], dis[3]);
    } else {
      for (int k = 0; k < 4 && x0 + k < cols; k++) {
        if (label_img) label_img[o + k] = lab[k];
        if (instance_img) instance_img[o + k] = ins[k];
        if (disparity_img) disparity_img[o + k] = dis[k];
      }
    }
  }
}

}  // namespace

void launch_rasterize(const KParams &p, const isx_section *sections, const int *n_sections, const isx_instance *inst,
                      const int *inst_count, int inst_cap, int *table, int nframes, uint8_t *label_img,
                      int32_t *instance_img, float *disparity_img, cudaStream_t s) {
  if (instance_img) {
    cudaMemsetAsync(table, 0, sizeof(int) * (size_t)nframes * p.realcols * kMaxSections, s);
    instance_table_kernel<<<dim3(16, nframes), 256, 0, s>>>(inst, inst_count, inst_cap, table, p.realcols);
    g_launch_count++;
  }
  dim3 grid(((p.cols + 3) / 4 + 255) / 256, (p.rows + kRasterRows - 1) / kRasterRows, nframes);
  rasterize_kernel<<<grid, 256, 0, s>>>(sections, n_sections, instance_img ? table : nullptr, label_img, instance_img,
                                        disparity_img, p.rows, p.cols, p.realcols, p.column_step);
  g_launch_count++;
}

}  // namespace isx
