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
// neighbourhood tests stay on chip and no per-call allocations, handle
// construction or host round trips remain.
//
// Two O(n^2) passes in all: (1) neighbour counts -> core points, (2) one sweep
// over the pairs j < i of core points that unites near pairs in a lock-free
// union-find forest whose links always point to the smaller index, so the root
// of a component is its lowest-index core point.  Points and forest live in
// shared memory when the class has at most kGroupSmemPts candidates (global
// memory otherwise, same code); non-core points are poisoned (x = +inf) in the
// staged copy so that the pair sweep needs no core test.
#include "kernels.h"

namespace isx {

namespace {

// threads per CTA: 256 for batches (one CTA per frame and class fits beside the DP), 1024 when only a few
// frames are in flight and the latency of the largest class is what the caller waits for
constexpr int kGroupThreads = 256;
constexpr int kGroupThreadsLatency = 1024;
constexpr int kGroupSmemPts = 4096;
constexpr size_t kGroupSmemBytes = (size_t)kGroupSmemPts * (sizeof(float2) + sizeof(int));

__device__ __forceinline__ bool near_(float2 a, float2 b, float eps2) {
  const float dx = fsub(a.x, b.x), dy = fsub(a.y, b.y);
  return fadd(fmul(dx, dx), fmul(dy, dy)) <= eps2;
}

// Forest reads bypass L1 (volatile): entries are rewritten by other threads' atomics.
__device__ __forceinline__ int ld_link(const int *comp, int x) { return *reinterpret_cast<const volatile int *>(comp + x); }
__device__ __forceinline__ int find_root(const int *comp, int x) {
  int q = ld_link(comp, x);
  while (q != x) {
    x = q;
    q = ld_link(comp, x);
  }
  return x;
}
// Links the larger root under the smaller one; an entry only ever changes from "root" to a smaller index.
__device__ __forceinline__ int unite(int *comp, int a, int b) {
  while (true) {
    a = find_root(comp, a);
    b = find_root(comp, b);
    if (a == b) return a;
    if (a < b) { const int t = a; a = b; b = t; }
    if (atomicCAS(comp + a, a, b) == a) return b;
  }
}

// P: coordinates (shared copy with non-core points poisoned, or the global array), comp: the forest.
template <bool SMEM>
__device__ __forceinline__ void group_points(float2 *P, const float2 *__restrict__ xy, int *comp,
                                             const uint8_t *__restrict__ cand, int *__restrict__ label, int n,
                                             float eps2, int min_pts, int *warp_sum, int *carry) {
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const float inf = inf_f();
  if (SMEM) {
    for (int i = tid; i < n; i += nthreads) P[i] = xy[i];
    __syncthreads();
  }
  // 1. core points (every point counts as a neighbour, core or not)
  for (int i = tid; i < n; i += nthreads) {
    const float2 pi = P[i];
    int deg = 0;
#pragma unroll 4
    for (int j = 0; j < n; j++) deg += near_(pi, P[j], eps2);
    comp[i] = (cand[i] && deg >= min_pts) ? i : -1;
  }
  __syncthreads();
  if (SMEM) {
    for (int i = tid; i < n; i += nthreads)
      if (comp[i] < 0) P[i].x = inf;
    __syncthreads();
  }
  // 2. connected components of the core points: unite every near pair j < i.  Rows i and n-1-i go to the
  //    same thread so that the triangular sweep is balanced.
  for (int h = tid; 2 * h < n; h += nthreads) {
    for (int side = 0; side < 2; side++) {
      const int i = side == 0 ? h : n - 1 - h;
      if (side == 1 && i == h) break;
      if (ld_link(comp, i) < 0) continue;
      const float2 pi = P[i];
      int ri = i;  // an ancestor of i (its root as far as this thread knows)
      for (int j = 0; j < i; j++) {
        if (!near_(pi, P[j], eps2)) continue;
        const int cj = ld_link(comp, j);
        if (cj == ri) continue;
        if (!SMEM && cj < 0) continue;
        ri = unite(comp, ri, cj);
        if (cj != ri && cj != j) atomicMin(comp + j, ri);  // shortcut for the next visitor (ri is an ancestor of j)
      }
    }
  }
  __syncthreads();
  // every core point -> its root (read-only chase, then a private write: ancestors stay ancestors)
  for (int i = tid; i < n; i += nthreads) {
    if (ld_link(comp, i) < 0) continue;
    const int r = find_root(comp, i);
    if (r != i) atomicMin(comp + i, r);
  }
  __syncthreads();
  // 3. rank the roots (comp[i] == i) by index -> cluster ids 0..k-1
  if (tid == 0) *carry = 0;
  __syncthreads();
  for (int start = 0; start < n; start += nthreads) {
    const int i = start + tid;
    const int is_rep = (i < n && ld_link(comp, i) == i) ? 1 : 0;
    int incl = is_rep;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += o;
    }
    if ((tid & 31) == 31) warp_sum[tid >> 5] = incl;
    __syncthreads();
    int before = *carry;
    for (int w = 0; w < (tid >> 5); w++) before += warp_sum[w];
    if (is_rep) label[i] = before + incl - 1;
    __syncthreads();
    if (tid == 0) {
      int t = *carry;
      for (int w = 0; w < nthreads / 32; w++) t += warp_sum[w];
      *carry = t;
    }
    __syncthreads();
  }
  // 4. core members take their root's id; border points the id of their lowest-index core neighbour;
  //    the rest is noise (-1).
  for (int i = tid; i < n; i += nthreads) {
    const int c = ld_link(comp, i);
    if (c == i) continue;  // root, labelled in step 3
    if (c >= 0) {
      label[i] = *reinterpret_cast<volatile int *>(label + c);
      continue;
    }
    const float2 pi = xy[i];
    int l = -1;
    for (int j = 0; j < n; j++) {
      if (near_(pi, P[j], eps2) && (SMEM || ld_link(comp, j) >= 0)) {
        l = *reinterpret_cast<volatile int *>(label + ld_link(comp, j));
        break;
      }
    }
    label[i] = l;
  }
}

__global__ void __launch_bounds__(kGroupThreadsLatency)
grouping_kernel(const int *__restrict__ cand_count, const float2 *__restrict__ cand_xy,
                const uint8_t *__restrict__ cand_core, int *__restrict__ cand_label, int *__restrict__ scratch,
                KParams p) {
  extern __shared__ __align__(16) unsigned char group_smem[];
  __shared__ int warp_sum[kGroupThreadsLatency / 32];
  __shared__ int carry;
  const int k = blockIdx.x, f = blockIdx.y;
  const int n = cand_count[f * kInstanceClasses + k];
  if (n == 0) return;
  const size_t cap = (size_t)p.realcols * kMaxSections;
  const size_t base = ((size_t)f * kInstanceClasses + k) * cap;
  const float2 *xy = cand_xy + base;
  const uint8_t *cand = cand_core + base;
  int *label = cand_label + base;  // final labels
  const float eps2 = fmul(p.eps_cluster, p.eps_cluster);
  if (n <= kGroupSmemPts) {
    float2 *sxy = reinterpret_cast<float2 *>(group_smem);
    int *scomp = reinterpret_cast<int *>(group_smem + (size_t)kGroupSmemPts * sizeof(float2));
    group_points<true>(sxy, xy, scomp, cand, label, n, eps2, p.min_pts, warp_sum, &carry);
  } else {
    // rare: more candidates than the staging area holds; the forest lives in the scratch array
    group_points<false>(const_cast<float2 *>(xy), xy, scratch + base, cand, label, n, eps2, p.min_pts, warp_sum,
                        &carry);
  }
}

}  // namespace

void launch_grouping(const KParams &p, const BatchBuffers &b, int nframes, cudaStream_t s) {
  static SmemOptIn optin;
  opt_in_smem(grouping_kernel, optin);
  dim3 grid(kInstanceClasses, nframes);
  const int threads = nframes <= 4 ? kGroupThreadsLatency : kGroupThreads;
  grouping_kernel<<<grid, threads, kGroupSmemBytes, s>>>(b.cand_count, b.cand_xy, b.cand_core, b.cand_label,
                                                               b.cand_scratch, p);
  g_launch_count++;
}

// Stand-alone grouping of one point set (the reference's ML::dbscanFit call, Stixels.cu:660-666): host buffers in,
// labels out.  Uses the same kernel as the path (grid 1 x 1).
int dbscan_fit_host(const float *xy, int n, float eps, int min_pts, const uint8_t *core_candidates, int *labels,
                    int threads) {
  if (n == 0) return 0;
  if (threads != kGroupThreads) threads = kGroupThreadsLatency;
  float2 *d_xy = nullptr;
  uint8_t *d_core = nullptr;
  int *d_label = nullptr, *d_scratch = nullptr, *d_count = nullptr;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return r == cudaSuccess; };
  if (ok(cudaMalloc(&d_xy, (size_t)n * sizeof(float2))) && ok(cudaMalloc(&d_core, n)) &&
      ok(cudaMalloc(&d_label, (size_t)n * sizeof(int))) && ok(cudaMalloc(&d_scratch, (size_t)n * sizeof(int))) &&
      ok(cudaMalloc(&d_count, sizeof(int)))) {
    ok(cudaMemcpy(d_xy, xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d_core, core_candidates, n, cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d_count, &n, sizeof(int), cudaMemcpyHostToDevice));
    KParams p{};
    p.realcols = 1;  // candidate capacity per (frame, class) slot is irrelevant for a 1 x 1 grid
    p.eps_cluster = eps;
    p.min_pts = min_pts;
    if (e == cudaSuccess) {
      static SmemOptIn optin;
      opt_in_smem(grouping_kernel, optin);
      grouping_kernel<<<dim3(1, 1), threads, kGroupSmemBytes>>>(d_count, d_xy, d_core, d_label, d_scratch, p);
      g_launch_count++;
      ok(cudaGetLastError());
      ok(cudaDeviceSynchronize());
      ok(cudaMemcpy(labels, d_label, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    }
  }
  cudaFree(d_xy); cudaFree(d_core); cudaFree(d_label); cudaFree(d_scratch); cudaFree(d_count);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace isx
