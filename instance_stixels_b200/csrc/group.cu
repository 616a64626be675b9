// Instance grouping on the GPU: size-filtered DBSCAN over the instance-centre
// votes of one semantic class of one frame.
// Replaces Stixels::ClusterInstances -> ML::dbscanFit of the non-vendored cuML
// fork tomsal/cuml@dbscan-sizefilter (InstanceStixels/src/Stixels.cu:639-681).
//
// The fork's source is not available; the semantics implemented here are the
// ones written down in oracle/dbscan_def.h (classic DBSCAN, core points must
// also be size-filter candidates, border points join their lowest-index core
// neighbour, clusters numbered by lowest core index).  One CTA per
// (frame, class): n is tens to a few thousand points, so the O(n^2)
// neighbourhood tests stay on chip (L1/L2) and no per-call allocations, handle
// construction or host round trips remain.
#include "kernels.h"

namespace isx {

namespace {

constexpr int kGroupThreads = 256;

__device__ __forceinline__ bool near_(float2 a, float2 b, float eps2) {
  const float dx = fsub(a.x, b.x), dy = fsub(a.y, b.y);
  return fadd(fmul(dx, dx), fmul(dy, dy)) <= eps2;
}

__global__ void __launch_bounds__(kGroupThreads)
grouping_kernel(const int *__restrict__ cand_count, const float2 *__restrict__ cand_xy,
                const uint8_t *__restrict__ cand_core, int *__restrict__ cand_label, int *__restrict__ scratch,
                KParams p) {
  __shared__ int changed;
  __shared__ int warp_sum[kGroupThreads / 32];
  __shared__ int carry;
  const int k = blockIdx.x, f = blockIdx.y;
  const int n = cand_count[f * kInstanceClasses + k];
  if (n == 0) return;
  const size_t cap = (size_t)p.realcols * kMaxSections;
  const size_t base = ((size_t)f * kInstanceClasses + k) * cap;
  const float2 *xy = cand_xy + base;
  const uint8_t *cand = cand_core + base;
  int *label = cand_label + base;  // final labels
  int *comp = scratch + base;      // component representative (lowest core index), -1 = not core
  const float eps2 = fmul(p.eps_cluster, p.eps_cluster);
  const int tid = threadIdx.x;

  // 1. core points
  for (int i = tid; i < n; i += kGroupThreads) {
    const float2 pi = xy[i];
    int deg = 0;
    for (int j = 0; j < n; j++) deg += near_(pi, xy[j], eps2);
    comp[i] = (cand[i] && deg >= p.min_pts) ? i : -1;
  }
  __syncthreads();
  // 2. connected components of core points: iterate "take the smallest
  //    representative among core neighbours" + pointer jumping to a fixpoint.
  while (true) {
    if (tid == 0) changed = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kGroupThreads) {
      int mine = comp[i];
      if (mine < 0) continue;
      const float2 pi = xy[i];
      int best = mine;
      for (int j = 0; j < n; j++) {
        const int cj = comp[j];
        if (cj >= 0 && cj < best && near_(pi, xy[j], eps2)) best = cj;
      }
      while (comp[best] < best) best = comp[best];  // representatives only ever decrease
      if (best < mine) {
        atomicMin(&comp[i], best);
        atomicMin(&comp[mine], best);
        changed = 1;
      }
    }
    __syncthreads();
    const int again = changed;
    __syncthreads();
    if (!again) break;
  }
  // 3. rank representatives (comp[i] == i) by index -> cluster ids 0..k-1
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < n; start += kGroupThreads) {
    const int i = start + tid;
    const int is_rep = (i < n && comp[i] == i) ? 1 : 0;
    int incl = is_rep;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += o;
    }
    if ((tid & 31) == 31) warp_sum[tid >> 5] = incl;
    __syncthreads();
    int before = carry;
    for (int w = 0; w < (tid >> 5); w++) before += warp_sum[w];
    if (is_rep) label[i] = before + incl - 1;
    __syncthreads();
    if (tid == 0) {
      int t = carry;
      for (int w = 0; w < kGroupThreads / 32; w++) t += warp_sum[w];
      carry = t;
    }
    __syncthreads();
  }
  // 4. core members take their representative's id; border points the id of
  //    their lowest-index core neighbour; the rest is noise (-1).
  for (int i = tid; i < n; i += kGroupThreads) {
    const int c = comp[i];
    if (c == i) continue;  // representative, labelled in step 3
    if (c >= 0) {
      label[i] = label[c];
      continue;
    }
    const float2 pi = xy[i];
    int l = -1;
    for (int j = 0; j < n; j++) {
      if (comp[j] >= 0 && near_(pi, xy[j], eps2)) {
        l = label[comp[j]];
        break;
      }
    }
    label[i] = l;
  }
}

}  // namespace

void launch_grouping(const KParams &p, const BatchBuffers &b, int nframes, cudaStream_t s) {
  dim3 grid(kInstanceClasses, nframes);
  grouping_kernel<<<grid, kGroupThreads, 0, s>>>(b.cand_count, b.cand_xy, b.cand_core, b.cand_label, b.cand_scratch,
                                                 p);
  g_launch_count++;
}

}  // namespace isx
