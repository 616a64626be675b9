// Host part of the model: configuration fan-out and the CPU-side tables.
//
// These tables are inputs of the device path whose VALUES must equal the
// reference's bit for bit (they are produced on the host there too, from the
// same glibc logf/erf/sqrtf: Stixels.cu:76-129, 361-446, 790-887), so each
// expression below keeps the reference's operand order and float/double
// promotion.  Compiled by nvcc like the reference's own host code (same
// <cmath> overload set for the unqualified erf()).
#include "host_model.h"

#include <cmath>
#include <cstdlib>

namespace isx {

namespace {
constexpr float kPiFloat = 3.1416f;       // Stixels.hpp:37
constexpr int kLogLutSize = 1000000;      // configuration.h:30

// Stixels::FastLog (Stixels.cu:786-788): quantised table lookup.
inline float fast_log(const std::vector<float> &lut, float v) {
  long idx = (long)((v)*kLogLutSize + 0.5f);
  if (idx < 0) idx = 0;                       // the reference would read out of bounds here
  if (idx > kLogLutSize) idx = kLogLutSize;
  return lut[(size_t)idx];
}
}  // namespace

std::string validate_config(const isx_config &c) {
  // Same checks, same order and same messages as Stixels::SetConfig (Stixels.cu:294-313).
  if (c.rows == -1 || c.cols == -1) return "Number of rows or columns are not set.";
  if (c.max_dis == -1) return "Maximum disparity value is not set.";
  if (c.eps == -1 || c.min_pts == -1 || c.size_filter == -1) return "Clustering parameters are not set.";
  if (c.prior_weight == -1 || c.segmentation_weight == -1 || c.instance_weight == -1 ||
      c.disparity_weight == -1)
    return "Energy term weights are not set.";
  if (c.column_step == -1) return "Stixel width is not set.";
  if (c.focal == -1 || c.baseline == -1) return "Camera parameters are not set.";
  return std::string();
}

void HostModel::apply(const isx_config &c) {
  rows = (int)c.rows;  // StixelConfig holds floats; SetDisparityParameters takes ints (types.h:33-34)
  cols = (int)c.cols;
  max_dis = c.max_dis;
  invalid_disparity = c.invalid_disparity;
  sigma_disparity_object = c.sigma_disparity_object;
  sigma_disparity_ground = c.sigma_disparity_ground;
  sigma_sky = c.sigma_sky;

  n_classes = c.n_semantic_classes;
  n_channels = c.n_semantic_classes + c.n_offset_channels;

  eps = c.eps;
  min_pts = c.min_pts;
  size_filter = c.size_filter;

  prior_weight = c.prior_weight;
  disparity_weight = c.disparity_weight;
  segmentation_weight = c.segmentation_weight;
  instance_weight = 0.0;
  if (c.segmentation_weight > 1e-5) {
    instance_weight = c.instance_weight / c.segmentation_weight;
    if (c.instance_weight < 1e-8) instance_weight = 0.0;
  }

  pout = c.pout;
  pout_sky = c.pout_sky;
  pnexists_given_ground = (c.pground_given_nexist * c.pnexist_dis) / c.pground;
  pnexists_given_object = (c.pobject_given_nexist * c.pnexist_dis) / c.pobject;
  pnexists_given_sky = (c.psky_given_nexist * c.pnexist_dis) / c.psky;
  pord = c.pord;
  pgrav = c.pgrav;
  pblg = c.pblg;

  column_step = c.column_step;
  median_join = c.median_join != 0;
  epsilon = c.epsilon;
  range_objects_z = c.range_objects_z;
  width_margin = c.width_margin;

  focal = c.focal;
  baseline = c.baseline;
  sigma_camera_tilt = c.sigma_camera_tilt * (kPiFloat) / 180.0f;  // degrees -> radians
  sigma_camera_height = c.sigma_camera_height;
  camera_center_x = c.camera_center_x;
  camera_center_y = c.camera_center_y;
}

void HostModel::derive() {
  realcols = (cols - width_margin) / column_step;
  max_disf = (float)max_dis;
  rows_power2 = (int)powf(2, ceilf(log2f(rows + 1)));
  rows_power2_seg = (int)powf(2, ceilf(log2f(rows / 8 + 1)));

  log_lut.resize((size_t)kLogLutSize + 1);
  for (int i = 0; i < kLogLutSize; i++) {
    const float x = (float)i / ((float)kLogLutSize);
    log_lut[i] = logf(x);
  }
  log_lut[kLogLutSize] = 0.0f;

  max_dis_log = logf(max_disf);
  rows_log = logf((float)rows);
  puniform_sky = max_dis_log - logf(pout_sky);
  puniform = max_dis_log - logf(pout);
  pnexists_given_sky_log = -logf(pnexists_given_sky);
  nopnexists_given_sky_log = -logf(1.0f - pnexists_given_sky);
  pnexists_given_ground_log = -logf(pnexists_given_ground);
  nopnexists_given_ground_log = -logf(1.0f - pnexists_given_ground);
  pnexists_given_object_log = -logf(pnexists_given_object);
  nopnexists_given_object_log = -logf(1.0f - pnexists_given_object);

  // ComputeObjectDisparityRange (Stixels.cu:879-887)
  object_disparity_range.assign(max_dis, 0.0f);
  for (int i = 0; i < max_dis; i++) {
    const float pm = (float)i;
    float range_disp = 0.0f;
    if (pm != 0) {
      const float pmean_plus_z = (baseline * focal / pm) + range_objects_z;
      range_disp = pm - (baseline * focal / pmean_plus_z);
    }
    object_disparity_range[i] = range_disp;
  }

  // PrecomputeSky (Stixels.cu:856-865)
  {
    const float sigma = sigma_sky;
    const float a_range = 0.5f * (erf(max_disf / (sigma * sqrtf(2.0f))) - erf(0.0f));
    normalization_sky =
        fast_log(log_lut, a_range) - logf((1.0f - pout_sky) / (sigma * sqrtf(2.0f * kPiFloat)));
    inv_sigma2_sky = 1.0f / (2.0f * sigma * sigma);
  }

  // PrecomputeObject (Stixels.cu:819-840)
  normalization_object.assign(max_dis, 0.0f);
  inv_sigma2_object.assign(max_dis, 0.0f);
  for (int dis = 0; dis < max_dis; dis++) {
    const float fn = (float)dis;
    const float sigma_object = fn * fn * range_objects_z / (focal * baseline);
    const float sigma = sqrtf(sigma_disparity_object * sigma_disparity_object + sigma_object * sigma_object);
    const float a_range =
        0.5f * (erf((max_disf - fn) / (sigma * sqrtf(2.0f))) - erf((-fn) / (sigma * sqrtf(2.0f))));
    normalization_object[dis] =
        fast_log(log_lut, a_range) - fast_log(log_lut, (1.0f - pout) / (sigma * sqrtf(2.0f * kPiFloat)));
    inv_sigma2_object[dis] = 1.0f / (2.0f * sigma * sigma);
  }

  // GetDataCostObject over all (fn, dis) (Stixels.cu:122-129, 842-854)
  obj_cost_lut.assign((size_t)max_dis * max_dis, 0.0f);
  for (int fn = 0; fn < max_dis; fn++) {
    for (int dis = 0; dis < max_dis; dis++) {
      float data_cost = pnexists_given_object_log;
      if (dis != (int)invalid_disparity) {
        const float model_diff = (float)(dis - fn);
        const float pgaussian = normalization_object[fn] + model_diff * model_diff * inv_sigma2_object[fn];
        const float p_data = fminf(puniform, pgaussian);
        data_cost = p_data + nopnexists_given_object_log;
      }
      obj_cost_lut[(size_t)fn * max_dis + dis] = data_cost;
    }
  }

  // 1./(vT+1-vB): double divide rounded to float (StixelsKernels.cu:485,608)
  inverse_height.assign((size_t)rows + 1, 0.0f);
  for (int n = 1; n <= rows; n++) inverse_height[n] = 1. / n;
}

int HostModel::ground_tables(const isx_road &road, float *out) const {
  const int vhor = rows - road.vhor - 1;
  const float camera_tilt = road.camera_tilt;
  const float camera_height = road.camera_height;
  const float alpha_ground = road.alpha_ground;
  float *ground_function = out;
  float *normalization_ground = out + rows;
  float *inv_sigma2_ground = out + 2 * (size_t)rows;

  const float fb = (focal * baseline) / camera_height;
  for (int v = 0; v < rows; v++) {
    const float fn = alpha_ground * (float)(vhor - v);
    ground_function[v] = fn;

    const float x = camera_tilt + (float)(vhor - v) / focal;
    const float sigma2_road =
        fb * fb *
        (sigma_camera_height * sigma_camera_height * x * x / (camera_height * camera_height) +
         sigma_camera_tilt * sigma_camera_tilt);
    const float sigma = sqrtf(sigma_disparity_ground * sigma_disparity_ground + sigma2_road);
    const float a_range =
        0.5f * (erf((max_disf - fn) / (sigma * sqrtf(2.0f))) - erf((-fn) / (sigma * sqrtf(2.0f))));
    normalization_ground[v] =
        fast_log(log_lut, a_range) - fast_log(log_lut, (1.0f - pout) / (sigma * sqrtf(2.0f * kPiFloat)));
    inv_sigma2_ground[v] = 1.0f / (2.0f * sigma * sigma);
  }
  return vhor;
}

}  // namespace isx
