// Backtracking and stixel emission on the device, plus deterministic
// collection of the instance-grouping candidates.
// Replaces the thread-0 tail of StixelsKernel (InstanceStixels/src/
// StixelsKernels.cu:843-955).
//
// The DP keeps (cost, argmin vB) per row and slot; the predecessor TYPE the
// reference stores in index_table is re-derived here from the final row costs
// with the same float operations (dp_common.cuh), only for the handful of
// cells on the optimal path.  The reference appends candidates with a global
// atomicAdd (:926-941), i.e. in nondeterministic order; here they are
// compacted in (class, column, index-from-top) order.
#include "dp_common.cuh"
#include "kernels.h"

namespace isx {

namespace {

struct Rec {
  uint32_t w[kRecWords];
};

// rec = row-major records of this column (records_b, common.cuh): one 128-byte line per row
__device__ __forceinline__ Rec load_rec(const uint32_t *rec, int v) {
  Rec r;
  const uint4 *row = reinterpret_cast<const uint4 *>(rec + (size_t)v * kRecBWords);
#pragma unroll
  for (int g = 0; g < (kRecWords + 3) / 4; g++) {
    const uint4 t = __ldg(row + g);
    r.w[4 * g] = t.x;
    r.w[4 * g + 1] = t.y;
    if (4 * g + 2 < kRecWords) r.w[4 * g + 2] = t.z;
    if (4 * g + 3 < kRecWords) r.w[4 * g + 3] = t.w;
  }
  return r;
}

__device__ __forceinline__ float rec_f(const Rec &r, int w) { return __uint_as_float(r.w[w]); }
// float(sum over the segment) of the instance means / their squares: the exact-float prefixes
// reproduce the reference's I2F.S64 of the int64 difference (common.cuh).
__device__ __forceinline__ float seg_mean_sum(const Rec &hi, const Rec &lo, int w) {
  return fsub(rec_f(hi, w), rec_f(lo, w));
}
__device__ __forceinline__ float seg_sq_sum(const Rec &hi, const Rec &lo, int w_hi) {
  return fadd(fsub(rec_f(hi, w_hi), rec_f(lo, w_hi)), fsub(rec_f(hi, w_hi + 1), rec_f(lo, w_hi + 1)));
}

__device__ __forceinline__ int seg_sum(const Rec &hi, const Rec &lo, int c) { return (int)(hi.w[c] - lo.w[c]); }

// GetObjectSegmentationClass (Cityscapes.h:85-111): classes 2..18 without sky,
// strict '>' so the lowest class wins ties; costs as floats like the reference.
__device__ int object_class(const Rec &hi, const Rec &lo, float ic, float nic) {
  float best = inf_f();
  int cls = 2;
  for (int c = 2; c < 19; c++) {
    if (c == kSkyClass) continue;
    float cost = fadd(0.0f, c < kSkyClass ? nic : ic);
    cost = fadd(cost, (float)seg_sum(hi, lo, c));
    if (best > cost) {
      best = cost;
      cls = c;
    }
  }
  return cls;
}

// Priors that decide the predecessor type of the segment (vB..vT) of `type`.
template <bool PAIRWISE>
__device__ int predecessor_type(int type, int vB, float fn_clamped, const float4 *dp_col, const float *S,
                                const float *pm_col, int vhor,
                                const float *__restrict__ object_disparity_range, const KParams &p) {
  const int pv = vB - 1;
  const float4 prev = dp_col[pv];
  const bool ground_side = pv < vhor;
  const float inf = inf_f();
  const float cg = ground_side ? prev.x : inf, cs = ground_side ? inf : prev.x, co = prev.y;
  if constexpr (!PAIRWISE) {
    // unary: raw previous costs pick the type (:694-697, 723-727, 782-787, 828-835)
    if (type != OBJECT) return prev_type_gs(cg, co);
    return prev_type_obj(cg, co, cs);
  } else {
    const float pm = pm_col[vB];
    RowPriors rp;
    const RowInfo q = make_row_info(S + (size_t)vB * kStatWords, ground_side, cg, co, cs, pm,
                                    object_disparity_range, p, &rp);
    if (type != OBJECT) return prev_type_gs(rp.g1, rp.g2);
    float p1, p2, p3;
    object_priors(q, ground_side, fn_clamped, p.epsilon, p1, p2, p3);
    // the reference compares the actual three priors; p1/p3 are +inf on the
    // side where C[p][GROUND] / C[p][SKY] are +inf
    return prev_type_obj(p1, p2, p3);
  }
}

// One thread per column, 32 columns per CTA: the per-stixel record loads of a warp touch 32 different lines, so the
// columns are spread over as many SMs as possible instead of sharing the load/store unit of a few.
constexpr int kBacktrackThreads = 32;

template <bool PAIRWISE>
__global__ void __launch_bounds__(kBacktrackThreads)
backtrack_kernel(const uint32_t *__restrict__ records, const float4 *__restrict__ dp, const float *__restrict__ stat,
                 const float *__restrict__ pm, const int *__restrict__ vhor_arr,
                 const float *__restrict__ object_disparity_range, isx_section *__restrict__ sections,
                 int *__restrict__ n_sections, int *error_flag, int ncolumns, KParams p) {
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  if (gcol >= ncolumns) return;
  const int H = p.rows, C = p.realcols;
  const int f = gcol / C;
  const int vhor = vhor_arr[f];
  const bool has_invalid = p.invalid_disparity >= 0.0f;
  const int Hp = p.rec_stride;
  const uint32_t *rec = records + (size_t)gcol * Hp * kRecBWords;
  const float4 *dp_col = dp + (size_t)gcol * H;
  const float *S = stat + (size_t)f * H * kStatWords;
  const float *pm_col = pm + (size_t)gcol * H;
  isx_section *out = sections + (size_t)gcol * kMaxSections;
  const float inf = inf_f();

  int vT = H - 1;
  // top row: OBJECT is the fallback type (:846-861)
  int type = OBJECT;
  {
    const float4 top = dp_col[vT];
    const float last_ground = vT < vhor ? top.x : inf, last_sky = vT < vhor ? inf : top.x, last_object = top.y;
    if (last_ground < last_object) type = GROUND;
    if (last_sky < fmin_(last_ground, last_object)) type = SKY;
  }
  int i = 0;
  while (true) {
    const float4 row = dp_col[vT];
    const int vB = (type == OBJECT) ? __float_as_int(row.w) : __float_as_int(row.z);
    const float cost = (type == OBJECT) ? row.y : row.x;
    const Rec hi = load_rec(rec, vT + 1), lo = load_rec(rec, vB);
    const int n = vT + 1 - vB;

    isx_section sec;
    sec.vT = vT;
    sec.vB = vB;
    sec.type = type;
    // ComputeMean without the clamp (:872-875)
    float mean;
    {
      const float sd = fsub(__uint_as_float(hi.w[kRecDisp]), __uint_as_float(lo.w[kRecDisp]));
      if (has_invalid) {
        const float vd = fsub(__uint_as_float(hi.w[kRecValid]), __uint_as_float(lo.w[kRecValid]));
        mean = (vd != 0.0f) ? fmul(sd, rcp_approx(vd)) : 0.0f;
      } else {
        mean = fmul(sd, rcp_approx((float)n));
      }
    }
    sec.disparity = mean;
    sec.cost = fmin_(cost, 10000.0f);
    const float rn = rcp_approx((float)n);
    const float fmx = seg_mean_sum(hi, lo, kRecMx);
    const float fmy = seg_mean_sum(hi, lo, kRecMy);
    sec.instance_meanx = fmul(fmx, rn);
    sec.instance_meany = fmul(fmy, rn);

    if (type == GROUND) {
      // GetGroundSegmentationClass (Cityscapes.h:52-58)
      sec.semantic_class = ((float)seg_sum(hi, lo, 0) < (float)seg_sum(hi, lo, 1)) ? 0 : 1;
    } else if (type == SKY || sec.disparity < 1.0f) {
      sec.type = SKY;  // far objects become sky (:894-902)
      sec.semantic_class = kSkyClass;
    } else {
      const float fmx2 = seg_sq_sum(hi, lo, kRecMx2Hi);
      const float fmy2 = seg_sq_sum(hi, lo, kRecMy2Hi);
      const float var = ffma(-fmul(fmy, fmy), rn, fadd(fmy2, ffma(-fmul(fmx, fmx), rn, fmx2)));
      const float ic = fmul(var, p.instance_weight);
      const float nic = fmul((float)seg_sum(hi, lo, kRecOff), p.instance_weight);
      sec.semantic_class = object_class(hi, lo, ic, nic);
    }
    out[i] = sec;
    i++;
    if (vB == 0) break;
    if (i >= kMaxSections - 1) {  // reference: assert(i < max_sections) (:950)
      atomicOr(error_flag + f, kErrSectionOverflow);
      break;
    }
    type = predecessor_type<PAIRWISE>(type, vB, clamp_neg(mean), dp_col, S, pm_col, vhor, object_disparity_range, p);
    vT = vB - 1;
  }
  isx_section term;
  term.type = -1;
  term.vB = term.vT = 0;
  term.disparity = term.cost = term.instance_meanx = term.instance_meany = 0.0f;
  term.semantic_class = 0;
  out[i] = term;
  n_sections[gcol] = i;
}

// One CTA per frame: count instance stixels per (column, class), scan over
// columns, then write the candidates in (class, column, index) order.
__global__ void __launch_bounds__(256)
collect_candidates_kernel(const isx_section *__restrict__ sections, const int *__restrict__ n_sections,
                          int *__restrict__ cand_count, int *__restrict__ cand_offset, float2 *__restrict__ cand_xy,
                          int2 *__restrict__ cand_idx, uint8_t *__restrict__ cand_core, KParams p) {
  __shared__ int warp_tot[8][kInstanceClasses];
  __shared__ int running[kInstanceClasses];
  const int f = blockIdx.x, C = p.realcols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const isx_section *fs = sections + (size_t)f * C * kMaxSections;
  const int *ns = n_sections + (size_t)f * C;
  int *off = cand_offset + (size_t)f * (C + 1) * kInstanceClasses;
  const size_t cap = (size_t)C * kMaxSections;
  if (tid < kInstanceClasses) running[tid] = 0;
  __syncthreads();
  for (int base = 0; base < C; base += 256) {
    const int col = base + tid;
    int cnt[kInstanceClasses];
#pragma unroll
    for (int k = 0; k < kInstanceClasses; k++) cnt[k] = 0;
    if (col < C) {
      const int n = ns[col];
      for (int j = 0; j < n; j++) {
        const isx_section s = fs[(size_t)col * kMaxSections + j];
        if (s.type == OBJECT && s.semantic_class >= kFirstInstanceClass) cnt[s.semantic_class - kFirstInstanceClass]++;
      }
    }
    int excl[kInstanceClasses];
#pragma unroll
    for (int k = 0; k < kInstanceClasses; k++) {
      int incl = cnt[k];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      excl[k] = incl - cnt[k];
      if (lane == 31) warp_tot[warp][k] = incl;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kInstanceClasses; k++) {
      int before = running[k];
      for (int w = 0; w < warp; w++) before += warp_tot[w][k];
      excl[k] += before;
    }
    if (col < C) {
#pragma unroll
      for (int k = 0; k < kInstanceClasses; k++) off[(size_t)col * kInstanceClasses + k] = excl[k];
      const int n = ns[col];
      for (int j = 0; j < n; j++) {
        const isx_section s = fs[(size_t)col * kMaxSections + j];
        if (s.type == OBJECT && s.semantic_class >= kFirstInstanceClass) {
          const int k = s.semantic_class - kFirstInstanceClass;
          const size_t dst = ((size_t)f * kInstanceClasses + k) * cap + excl[k]++;
          cand_xy[dst] = make_float2(s.instance_meanx, s.instance_meany);
          cand_idx[dst] = make_int2(col, j);
          cand_core[dst] = (s.vT + 1 - s.vB) >= p.size_filter;  // core candidate = size filter (:940-941)
        }
      }
    }
    __syncthreads();
    if (tid < kInstanceClasses) {
      int t = running[tid];
      for (int w = 0; w < 8; w++) t += warp_tot[w][tid];
      running[tid] = t;
    }
    __syncthreads();
  }
  if (tid < kInstanceClasses) {
    cand_count[f * kInstanceClasses + tid] = running[tid];
    off[(size_t)C * kInstanceClasses + tid] = running[tid];
  }
}

// ---------------------------------------------------------------------------
// Result packing.  A frame yields a few thousand stixels, but the reference's result array is the full-capacity
// [C][200] x 32 B (Stixels.cu:629-633 copies all of it, 1.6 MB per frame, for ~80 KB of payload).  This kernel
// writes what a host caller needs -- the used Sections of every column back to back, the per-column counts, the
// instance records (class, column, index order; GetInstanceStixels, Stixels.cu:744-776) and one descriptor per
// frame -- straight into PINNED HOST memory mapped into the device address space, in 1 KB bursts per warp.  There is
// no device -> host memcpy and no size the host would have to learn first: when the emission stream's event fires
// the packed results are in host memory, and the host expands them into the caller's [C][200] array.
// One CTA per frame.  The frames of a batch take their place in the packed arrays with one atomicAdd per frame on
// the result set's cursors (placement varies from run to run, content does not; the descriptor says where).
// A frame that does not fit (more stixels than the packed arrays were sized for) is flagged and left to the
// padded device arrays, which stay complete in any case.
// ---------------------------------------------------------------------------
constexpr int kPackThreads = 256;

__global__ void __launch_bounds__(kPackThreads)
pack_results_kernel(PackArgs a, KParams p) {
  __shared__ int warp_tot[kPackThreads / 32];
  __shared__ int s_base[2], s_running;
  const int f = blockIdx.x, C = p.realcols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const isx_section *fs = a.sections + (size_t)f * C * kMaxSections;
  const int *ns = a.n_sections + (size_t)f * C;
  int *off = a.col_offset + (size_t)f * (C + 1);
  // ---- exclusive prefix of the stixel counts over the columns ----
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < C; base += kPackThreads) {
    const int col = base + tid;
    const int n = col < C ? ns[col] : 0;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = s_running;
    for (int w = 0; w < warp; w++) before += warp_tot[w];
    if (col < C) {
      off[col] = before + incl - n;
      a.h_counts[(size_t)f * C + col] = n;
    }
    __syncthreads();
    if (tid == 0) {
      int t = s_running;
      for (int w = 0; w < kPackThreads / 32; w++) t += warp_tot[w];
      s_running = t;
    }
    __syncthreads();
  }
  const int n_sec = s_running;
  int n_inst = 0;
  for (int k = 0; k < kInstanceClasses; k++) n_inst += a.cand_count[f * kInstanceClasses + k];
  if (tid == 0) {
    s_base[0] = atomicAdd(a.cursors + 0, n_sec);
    s_base[1] = atomicAdd(a.cursors + 1, n_inst);
  }
  __syncthreads();
  const int sec_base = s_base[0], inst_base = s_base[1];
  const bool sec_fits = sec_base + n_sec <= a.h_sections_cap;
  const bool inst_fits = inst_base + n_inst <= a.h_inst_cap;
  // ---- sections: one warp per column, 32 B per lane (two 16-byte stores; a warp writes up to 1 KB back to back) ----
  if (a.h_padded) {
    // the caller's padded array is device-visible: no packed copy, no expansion on the host
    for (int col = warp; col < C; col += kPackThreads / 32) {
      const int n = ns[col];
      const uint4 *src = reinterpret_cast<const uint4 *>(fs + (size_t)col * kMaxSections);
      uint4 *dst = reinterpret_cast<uint4 *>(a.h_padded + ((size_t)f * C + col) * kMaxSections);
      for (int j = lane; j < 2 * (n + 1); j += 32) dst[j] = src[j];   // n stixels + the type == -1 terminator
    }
  } else if (sec_fits) {
    for (int col = warp; col < C; col += kPackThreads / 32) {
      const int n = ns[col];
      const uint4 *src = reinterpret_cast<const uint4 *>(fs + (size_t)col * kMaxSections);
      uint4 *dst = reinterpret_cast<uint4 *>(a.h_sections + sec_base + off[col]);
      for (int j = lane; j < 2 * n; j += 32) dst[j] = src[j];
    }
  }
  // ---- instance records in (class, column, index) order: device copy (rasteriser, fetch) and host copy ----
  {
    const size_t cap = (size_t)C * kMaxSections;
    int base = 0;
    for (int k = 0; k < kInstanceClasses; k++) {
      const int n = a.cand_count[f * kInstanceClasses + k];
      const size_t src = ((size_t)f * kInstanceClasses + k) * cap;
      for (int i = tid; i < n; i += kPackThreads) {
        const int2 ci = a.cand_idx[src + i];
        const int4 r = make_int4(ci.x, ci.y, a.cand_label[src + i], kFirstInstanceClass + k);  // isx_instance
        const int dst = base + i;
        if (dst < a.inst_cap) *reinterpret_cast<int4 *>(a.inst_out + (size_t)f * a.inst_cap + dst) = r;
        if (inst_fits) *reinterpret_cast<int4 *>(a.h_inst + inst_base + dst) = r;
      }
      base += n;
    }
  }
  if (tid == 0) {
    a.inst_count_out[f] = n_inst;
    isx_packed_frame d;
    d.section_offset = a.h_padded ? -1 : sec_base;
    d.section_count = n_sec;
    d.instance_offset = inst_base;
    d.instance_count = n_inst;
    d.error = a.err[f];
    d.overflow = ((sec_fits || a.h_padded) ? 0 : 1) | (inst_fits ? 0 : 2);
    d.reserved[0] = d.reserved[1] = 0;
    a.h_frames[f] = d;
  }
}

// cost_table / index_table in the reference's layout (parity tests only).
template <bool PAIRWISE>
__global__ void export_tables_kernel(const uint32_t *__restrict__ records, const float4 *__restrict__ dp,
                                     const float *__restrict__ stat, const float *__restrict__ pm,
                                     const int *__restrict__ vhor_arr,
                                     const float *__restrict__ object_disparity_range, int frame,
                                     float *__restrict__ cost_table, int *__restrict__ index_table, KParams p) {
  const int H = p.rows, C = p.realcols;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * H) return;
  const int col = idx / H, vT = idx - col * H;
  const int gcol = frame * C + col;
  const int vhor = vhor_arr[frame];
  const bool has_invalid = p.invalid_disparity >= 0.0f;
  const int Hp = p.rec_stride;
  const uint32_t *rec = records + (size_t)gcol * Hp * kRecBWords;
  const float4 *dp_col = dp + (size_t)gcol * H;
  const float *S = stat + (size_t)frame * H * kStatWords;
  const float *pm_col = pm + (size_t)gcol * H;
  const float4 row = dp_col[vT];
  const float inf = inf_f();
  float *ct = cost_table + ((size_t)col * H + vT) * 3;
  int *it = index_table + ((size_t)col * H + vT) * 3;
  const int gs_type = vT < vhor ? GROUND : SKY;
  ct[GROUND] = gs_type == GROUND ? row.x : inf;
  ct[SKY] = gs_type == SKY ? row.x : inf;
  ct[OBJECT] = row.y;
  it[GROUND] = it[SKY] = it[OBJECT] = -1;
  for (int t = 0; t < 2; t++) {
    const int type = t == 0 ? gs_type : OBJECT;
    const float c = t == 0 ? row.x : row.y;
    const int vB = __float_as_int(t == 0 ? row.z : row.w);
    if (!(c < inf) && !(type == OBJECT && vB == 0)) continue;
    if (vB == 0) {
      it[type] = type == GROUND ? GROUND : OBJECT;  // (:564, 592)
      continue;
    }
    const Rec hi = load_rec(rec, vT + 1), lo = load_rec(rec, vB);
    const float fn = segment_mean(__uint_as_float(hi.w[kRecDisp]), __uint_as_float(lo.w[kRecDisp]),
                                  __uint_as_float(hi.w[kRecValid]), __uint_as_float(lo.w[kRecValid]),
                                  vT + 1 - vB, has_invalid);
    it[type] = vB * 3 + predecessor_type<PAIRWISE>(type, vB, fn, dp_col, S, pm_col, vhor, object_disparity_range, p);
  }
}

}  // namespace

void launch_emit(const KParams &p, const BatchBuffers &b, int nframes, bool pairwise, cudaStream_t s) {
  const int ncolumns = nframes * p.realcols;
  const int grid = (ncolumns + kBacktrackThreads - 1) / kBacktrackThreads;
  const uint32_t *rec = b.records_b;
  if (pairwise)
    backtrack_kernel<true><<<grid, kBacktrackThreads, 0, s>>>(rec, b.dp, b.stat, b.pm, b.vhor, b.object_disparity_range,
                                                b.sections, b.n_sections, b.error_flag, ncolumns, p);
  else
    backtrack_kernel<false><<<grid, kBacktrackThreads, 0, s>>>(rec, b.dp, b.stat, b.pm, b.vhor, b.object_disparity_range,
                                                 b.sections, b.n_sections, b.error_flag, ncolumns, p);
  collect_candidates_kernel<<<nframes, 256, 0, s>>>(b.sections, b.n_sections, b.cand_count, b.cand_offset, b.cand_xy,
                                                     b.cand_idx, b.cand_core, p);
  g_launch_count += 2;
}

void launch_pack(const KParams &p, const PackArgs &a, int nframes, cudaStream_t s) {
  pack_results_kernel<<<nframes, kPackThreads, 0, s>>>(a, p);
  g_launch_count++;
}

void launch_export_tables(const KParams &p, const BatchBuffers &b, int frame, bool pairwise, float *cost_table,
                          int *index_table, cudaStream_t s) {
  const int n = p.realcols * p.rows;
  const uint32_t *rec = b.records_b;
  if (pairwise)
    export_tables_kernel<true><<<(n + 127) / 128, 128, 0, s>>>(rec, b.dp, b.stat, b.pm, b.vhor,
                                                               b.object_disparity_range, frame, cost_table,
                                                               index_table, p);
  else
    export_tables_kernel<false><<<(n + 127) / 128, 128, 0, s>>>(rec, b.dp, b.stat, b.pm, b.vhor,
                                                                b.object_disparity_range, frame, cost_table,
                                                                index_table, p);
  g_launch_count++;
}

}  // namespace isx
