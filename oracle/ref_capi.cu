// TEST INFRASTRUCTURE -- not part of the product.
//
// extern "C" driver around the UNMODIFIED reference `Stixels` class
// (/root/reference/InstanceStixels/src/{Stixels,StixelsKernels}.cu, compiled
// in place by oracle/Makefile into oracle/_ref/libref_stixels.so).  It lets
// tests/ and bench.py (--impl reference) replay the exact call sequence of
// apps/run_cityscapes.cu:335-449 on the GPU box and read the reference's
// intermediates for stage-by-stage parity.
//
// `private` is re-spelled only to READ device pointers of intermediates; the
// class layout is unchanged (all data members already share one access
// section, Stixels.hpp:98-218) and no reference code is altered.
#define private public
#include "Stixels.hpp"
#undef private

#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <vector>

#include "../include/instance_stixels_b200.h"

namespace {
struct RefCtx {
  Stixels stixels;
  StixelConfig cfg;
  StixelsData data;
  std::map<std::pair<int, int>, int> instances;
  std::vector<int32_t> seg_backup;  // Compute destroys d_segmentation in place
};

StixelConfig to_ref(const isx_config &c) {
  StixelConfig r;
  r.rows = c.rows;
  r.cols = c.cols;
  r.max_dis = c.max_dis;
  r.invalid_disparity = c.invalid_disparity;
  r.eps = c.eps;
  r.min_pts = c.min_pts;
  r.size_filter = c.size_filter;
  r.n_semantic_classes = c.n_semantic_classes;
  r.n_offset_channels = c.n_offset_channels;
  r.prior_weight = c.prior_weight;
  r.segmentation_weight = c.segmentation_weight;
  r.instance_weight = c.instance_weight;
  r.disparity_weight = c.disparity_weight;
  r.pairwise = c.pairwise != 0;
  r.column_step = c.column_step;
  r.focal = c.focal;
  r.baseline = c.baseline;
  r.camera_center_x = c.camera_center_x;
  r.camera_center_y = c.camera_center_y;
  r.sigma_disparity_object = c.sigma_disparity_object;
  r.sigma_disparity_ground = c.sigma_disparity_ground;
  r.sigma_sky = c.sigma_sky;
  r.pout = c.pout;
  r.pout_sky = c.pout_sky;
  r.pord = c.pord;
  r.pgrav = c.pgrav;
  r.pblg = c.pblg;
  r.pground_given_nexist = c.pground_given_nexist;
  r.pobject_given_nexist = c.pobject_given_nexist;
  r.psky_given_nexist = c.psky_given_nexist;
  r.pnexist_dis = c.pnexist_dis;
  r.pground = c.pground;
  r.pobject = c.pobject;
  r.psky = c.psky;
  r.width_margin = c.width_margin;
  r.sigma_camera_tilt = c.sigma_camera_tilt;
  r.sigma_camera_height = c.sigma_camera_height;
  r.median_join = c.median_join != 0;
  r.epsilon = c.epsilon;
  r.range_objects_z = c.range_objects_z;
  r.road_vdisparity_threshold = c.road_vdisparity_threshold;
  return r;
}
}  // namespace

extern "C" {

// Writes the reference's own StixelConfig defaults (types.h:30-141) so the
// tests can pin isx_config_init against them.
void ref_config_init(isx_config *c) {
  StixelConfig r;
  c->rows = r.rows;
  c->cols = r.cols;
  c->max_dis = r.max_dis;
  c->invalid_disparity = r.invalid_disparity;
  c->eps = r.eps;
  c->min_pts = r.min_pts;
  c->size_filter = r.size_filter;
  c->n_semantic_classes = r.n_semantic_classes;
  c->n_offset_channels = r.n_offset_channels;
  c->prior_weight = r.prior_weight;
  c->segmentation_weight = r.segmentation_weight;
  c->instance_weight = r.instance_weight;
  c->disparity_weight = r.disparity_weight;
  c->pairwise = r.pairwise;
  c->column_step = r.column_step;
  c->focal = r.focal;
  c->baseline = r.baseline;
  c->camera_center_x = r.camera_center_x;
  c->camera_center_y = r.camera_center_y;
  c->sigma_disparity_object = r.sigma_disparity_object;
  c->sigma_disparity_ground = r.sigma_disparity_ground;
  c->sigma_sky = r.sigma_sky;
  c->pout = r.pout;
  c->pout_sky = r.pout_sky;
  c->pord = r.pord;
  c->pgrav = r.pgrav;
  c->pblg = r.pblg;
  c->pground_given_nexist = r.pground_given_nexist;
  c->pobject_given_nexist = r.pobject_given_nexist;
  c->psky_given_nexist = r.psky_given_nexist;
  c->pnexist_dis = r.pnexist_dis;
  c->pground = r.pground;
  c->pobject = r.pobject;
  c->psky = r.psky;
  c->width_margin = r.width_margin;
  c->sigma_camera_tilt = r.sigma_camera_tilt;
  c->sigma_camera_height = r.sigma_camera_height;
  c->median_join = r.median_join;
  c->epsilon = r.epsilon;
  c->range_objects_z = r.range_objects_z;
  c->road_vdisparity_threshold = r.road_vdisparity_threshold;
}

void *ref_create(void) { return new (std::nothrow) RefCtx(); }

void ref_destroy(void *p) {
  RefCtx *c = static_cast<RefCtx *>(p);
  if (!c) return;
  if (c->stixels.IsInitialized()) c->stixels.Finish();
  delete c;
}

// SetConfig + Initialize (apps/run_cityscapes.cu:335-336).  -1 where the
// reference throws std::invalid_argument.
int ref_configure(void *p, const isx_config *cfg) {
  RefCtx *c = static_cast<RefCtx *>(p);
  try {
    if (c->stixels.IsInitialized()) c->stixels.Finish();
    c->cfg = to_ref(*cfg);
    c->stixels.SetConfig(c->cfg);
    c->stixels.Initialize();
  } catch (const std::invalid_argument &e) {
    std::fprintf(stderr, "ref_configure: %s\n", e.what());
    return -1;
  }
  return 0;
}

int ref_real_cols(void *p) { return static_cast<RefCtx *>(p)->stixels.GetRealCols(); }
int ref_max_sections(void *p) { return static_cast<RefCtx *>(p)->stixels.GetMaxSections(); }
size_t ref_segmentation_elems(void *p) {
  RefCtx *c = static_cast<RefCtx *>(p);
  return (size_t)c->stixels.m_params.rows_power2_segmentation * c->stixels.m_realcols *
         c->stixels.m_segmentation_channels;
}

// One frame through the reference's public API exactly like
// apps/run_cityscapes.cu:346,383-387,406-411,430-431.
int ref_compute_frame(void *p, int pairwise, const float *disparity, size_t n_disp, const int32_t *seg,
                      size_t n_seg, int vhor, float tilt, float height, float alpha, isx_section *sections,
                      isx_frame_meta *meta) {
  RefCtx *c = static_cast<RefCtx *>(p);
  static_assert(sizeof(isx_section) == sizeof(Section), "Section layout");
  std::vector<pixel_t> disp(disparity, disparity + n_disp);
  std::vector<int32_t> segv(seg, seg + n_seg);
  c->stixels.SetDisparityImage(disp);
  c->stixels.SetSegmentation(segv);
  c->stixels.SetRoadParameters(vhor, tilt, height, alpha);
  c->stixels.Compute(pairwise != 0, c->data);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    std::fprintf(stderr, "ref_compute_frame: %s\n", cudaGetErrorString(err));
    return -3;
  }
  c->instances = c->stixels.GetInstanceStixels();
  if (sections)
    std::memcpy(sections, c->data.sections.data(), c->data.sections.size() * sizeof(Section));
  if (meta) {
    meta->rows = c->data.rows;
    meta->cols = c->data.cols;
    meta->realcols = c->data.realcols;
    meta->max_sections = c->data.max_sections;
    meta->max_dis = c->data.max_dis;
    meta->column_step = c->data.column_step;
    meta->semantic_classes = c->data.semantic_classes;
    meta->alpha_ground = c->data.alpha_ground;
    meta->vhor = c->data.vhor;
  }
  return 0;
}

// Timed variant for bench.py --impl reference: same calls, no result copies
// beyond what Compute/GetInstanceStixels do themselves; returns wall seconds
// measured like apps/run_cityscapes.cu:372-416 (sync, steady clock, sync).
double ref_time_frames(void *p, int pairwise, int n_frames, const float *disparity, size_t n_disp,
                       const int32_t *seg, size_t n_seg, int vhor, float tilt, float height, float alpha) {
  RefCtx *c = static_cast<RefCtx *>(p);
  std::vector<std::vector<pixel_t>> disp(n_frames);
  std::vector<std::vector<int32_t>> segv(n_frames);
  for (int f = 0; f < n_frames; f++) {
    disp[f].assign(disparity + (size_t)f * n_disp, disparity + (size_t)(f + 1) * n_disp);
    segv[f].assign(seg + (size_t)f * n_seg, seg + (size_t)(f + 1) * n_seg);
  }
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int f = 0; f < n_frames; f++) {
    c->stixels.SetDisparityImage(disp[f]);
    c->stixels.SetSegmentation(segv[f]);
    c->stixels.SetRoadParameters(vhor, tilt, height, alpha);
    c->stixels.Compute(pairwise != 0, c->data);
    c->instances = c->stixels.GetInstanceStixels();
  }
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (double)ms * 1e-3;
}

extern "C" double ref_dbscan_standin_seconds(int reset);  // dbscan_standin.cu

// The same loop with the host clock around every call of the sequence (they all block: cudaMemcpy / device-wide
// synchronisation inside): out[0] SetDisparityImage + SetSegmentation (H2D), out[1] Compute (road tables, kernels,
// ClusterInstances, D2H of all Sections), out[2] GetInstanceStixels (two full-capacity D2H copies + std::map),
// out[3] the part of out[1] spent in the DBSCAN stand-in (not reference code).  Seconds over all frames.
double ref_time_frames_split(void *p, int pairwise, int n_frames, const float *disparity, size_t n_disp,
                             const int32_t *seg, size_t n_seg, int vhor, float tilt, float height, float alpha,
                             double *out) {
  RefCtx *c = static_cast<RefCtx *>(p);
  std::vector<std::vector<pixel_t>> disp(n_frames);
  std::vector<std::vector<int32_t>> segv(n_frames);
  for (int f = 0; f < n_frames; f++) {
    disp[f].assign(disparity + (size_t)f * n_disp, disparity + (size_t)(f + 1) * n_disp);
    segv[f].assign(seg + (size_t)f * n_seg, seg + (size_t)(f + 1) * n_seg);
  }
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  cudaDeviceSynchronize();
  ref_dbscan_standin_seconds(1);
  double t_set = 0, t_compute = 0, t_get = 0;
  const auto start = clk::now();
  for (int f = 0; f < n_frames; f++) {
    const auto a = clk::now();
    c->stixels.SetDisparityImage(disp[f]);
    c->stixels.SetSegmentation(segv[f]);
    c->stixels.SetRoadParameters(vhor, tilt, height, alpha);
    cudaDeviceSynchronize();   // the disparity copy is the one asynchronous call (pageable source: returns once staged)
    const auto b = clk::now();
    c->stixels.Compute(pairwise != 0, c->data);
    const auto d = clk::now();
    c->instances = c->stixels.GetInstanceStixels();
    const auto e = clk::now();
    t_set += secs(a, b);
    t_compute += secs(b, d);
    t_get += secs(d, e);
  }
  cudaDeviceSynchronize();
  const double total = secs(start, clk::now());
  if (out) {
    out[0] = t_set;
    out[1] = t_compute;
    out[2] = t_get;
    out[3] = ref_dbscan_standin_seconds(1);
  }
  return total;
}

int ref_num_instances(void *p) { return (int)static_cast<RefCtx *>(p)->instances.size(); }

int ref_get_instances(void *p, isx_instance *out, int capacity) {
  RefCtx *c = static_cast<RefCtx *>(p);
  int i = 0;
  const int ms = c->stixels.GetMaxSections();
  for (const auto &kv : c->instances) {
    if (i >= capacity) break;
    out[i].column = kv.first.first;
    out[i].index = kv.first.second;
    out[i].label = kv.second;
    out[i].semantic_class = c->data.sections[(size_t)kv.first.first * ms + kv.first.second].semantic_class;
    i++;
  }
  return i;
}

// ---- intermediates of the last Compute (device -> host) -------------------
size_t ref_tensor_elems(void *p, int tensor) {
  RefCtx *c = static_cast<RefCtx *>(p);
  const Stixels &s = c->stixels;
  const size_t C = s.m_realcols, H = s.m_rows, D = s.m_max_dis;
  switch (tensor) {
    case ISX_T_JOINED_DISPARITY: return C * H;
    case ISX_T_OBJECT_LUT: return C * D * (H + 1);
    default: return 0;
  }
}

int ref_read_tensor(void *p, int tensor, void *host, size_t bytes) {
  RefCtx *c = static_cast<RefCtx *>(p);
  const Stixels &s = c->stixels;
  const size_t C = s.m_realcols, H = s.m_rows, D = s.m_max_dis;
  const size_t need = ref_tensor_elems(p, tensor) * 4;
  if (need == 0 || bytes < need) return -1;
  if (tensor == ISX_T_JOINED_DISPARITY) {
    return cudaMemcpy(host, s.d_disparity, need, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
  }
  if (tensor == ISX_T_OBJECT_LUT) {
    // reference row stride is rows_power2+1 (Stixels.cu:159-160); compact it.
    const size_t stride = (size_t)s.m_params.rows_power2 + 1;
    return cudaMemcpy2D(host, (H + 1) * 4, s.d_object_lut, stride * 4, (H + 1) * 4, C * D,
                        cudaMemcpyDeviceToHost) == cudaSuccess
               ? 0
               : -3;
  }
  return -1;
}

// Host-side per-frame ground tables of the last Compute (Stixels.cu:790-817).
int ref_read_ground_tables(void *p, float *ground_function, float *normalization, float *inv_sigma2) {
  RefCtx *c = static_cast<RefCtx *>(p);
  const Stixels &s = c->stixels;
  std::memcpy(ground_function, s.m_ground_function, sizeof(float) * s.m_rows);
  std::memcpy(normalization, s.m_normalization_ground, sizeof(float) * s.m_rows);
  std::memcpy(inv_sigma2, s.m_inv_sigma2_ground, sizeof(float) * s.m_rows);
  return 0;
}

// Host-side LUTs of Initialize (Stixels.cu:111-129): obj_cost_lut [D][D],
// object_disparity_range [D], and the packed kernel parameters.
int ref_read_init_tables(void *p, float *obj_cost_lut, float *object_disparity_range, float *params38) {
  RefCtx *c = static_cast<RefCtx *>(p);
  const Stixels &s = c->stixels;
  const size_t D = s.m_max_dis;
  if (obj_cost_lut) std::memcpy(obj_cost_lut, s.m_obj_cost_lut, sizeof(float) * D * D);
  if (object_disparity_range)
    std::memcpy(object_disparity_range, s.m_object_disparity_range, sizeof(float) * D);
  if (params38) std::memcpy(params38, &s.m_params, sizeof(StixelParameters));
  return 0;
}

size_t ref_params_bytes(void) { return sizeof(StixelParameters); }

}  // extern "C"
