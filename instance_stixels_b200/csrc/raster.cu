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

__global__ void __launch_bounds__(256)
rasterize_kernel(const isx_section *__restrict__ sections, const int *__restrict__ n_sections,
                 const int *__restrict__ table, uint8_t *__restrict__ label_img, int32_t *__restrict__ instance_img,
                 float *__restrict__ disparity_img, int rows, int cols, int realcols, int column_step) {
  const int f = blockIdx.z, y = blockIdx.y;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x0 >= cols) return;
  const int v = rows - 1 - y;
  uint8_t lab[4];
  int ins[4];
  float dis[4];
  int cached_col = -1, cl = 0, ci = 0;
  float cd = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = x0 + k;
    const int c = x / column_step;
    if (x < cols && c < realcols) {
      if (c != cached_col) {
        cached_col = c;
        const isx_section *col = sections + ((size_t)f * realcols + c) * kMaxSections;
        int lo = 0, hi = n_sections[(size_t)f * realcols + c] - 1;  // vB decreases with the index
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (col[mid].vB <= v) hi = mid; else lo = mid + 1;
        }
        const isx_section s = col[lo];
        const bool hit = lo >= 0 && s.type >= 0 && s.vB <= v && v <= s.vT;
        const int cls = s.semantic_class;
        cl = (hit && cls >= 0 && cls < 19) ? kTrainIdToId[cls] : 0;
        ci = (hit && table) ? table[((size_t)f * realcols + c) * kMaxSections + lo] : 0;
        cd = hit ? s.disparity : 0.0f;
      }
      lab[k] = (uint8_t)cl; ins[k] = ci; dis[k] = cd;
    } else {
      lab[k] = 0; ins[k] = 0; dis[k] = 0.0f;
    }
  }
  const size_t o = ((size_t)f * rows + y) * cols + x0;
  if (x0 + 3 < cols && (cols & 3) == 0) {
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

}  // namespace

void launch_rasterize(const KParams &p, const isx_section *sections, const int *n_sections, const isx_instance *inst,
                      const int *inst_count, int inst_cap, int *table, int nframes, uint8_t *label_img,
                      int32_t *instance_img, float *disparity_img, cudaStream_t s) {
  if (instance_img) {
    cudaMemsetAsync(table, 0, sizeof(int) * (size_t)nframes * p.realcols * kMaxSections, s);
    instance_table_kernel<<<dim3(16, nframes), 256, 0, s>>>(inst, inst_count, inst_cap, table, p.realcols);
    g_launch_count++;
  }
  dim3 grid((p.cols / 4 + 256) / 256, p.rows, nframes);
  rasterize_kernel<<<grid, 256, 0, s>>>(sections, n_sections, instance_img ? table : nullptr, label_img, instance_img,
                                        disparity_img, p.rows, p.cols, p.realcols, p.column_step);
  g_launch_count++;
}

}  // namespace isx
