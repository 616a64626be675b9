// Host-side model state: the configuration the reference keeps in the
// `Stixels` members (Stixels.hpp:98-218) and the tables it derives from it
// in Initialize() / Compute() on the CPU (Stixels.cu:76-129, 790-887).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/instance_stixels_b200.h"

namespace isx {

struct HostModel {
  // ---- raw settings, named after the setters that write them ----
  // SetDisparityParameters (Stixels.cu:425-437)
  int rows = 0, cols = 0, max_dis = 0;
  float invalid_disparity = -1.0f;
  float sigma_disparity_object = 1.0f, sigma_disparity_ground = 2.0f, sigma_sky = 0.1f;
  // SetSegmentationParameters (:402-406)
  int n_classes = 19, n_channels = 21;
  // SetClusteringParameters (:395-400)
  float eps = 0.f;
  int min_pts = 0, size_filter = 0;
  // SetWeightParameters (:408-423) -- instance_weight already divided by segmentation_weight
  float prior_weight = 0.f, disparity_weight = 0.f, segmentation_weight = 0.f, instance_weight = 0.f;
  // SetProbabilities (:361-373)
  float pout = 0.15f, pout_sky = 0.4f;
  float pnexists_given_ground = 0.f, pnexists_given_object = 0.f, pnexists_given_sky = 0.f;
  float pord = 0.2f, pgrav = 0.1f, pblg = 0.04f;
  // SetCameraParameters (:383-393) -- sigma_camera_tilt converted to radians
  float focal = -1.f, baseline = -1.f, sigma_camera_tilt = 0.f, sigma_camera_height = 0.f;
  float camera_center_x = -1.f, camera_center_y = -1.f;
  // SetModelParameters (:439-446)
  int column_step = 0, width_margin = 0;
  bool median_join = false;
  float epsilon = 3.0f, range_objects_z = 10.2f;

  // ---- derived once per Initialize (Stixels.cu:44-133) ----
  int realcols = 0, rows_power2 = 0, rows_power2_seg = 0;
  float max_disf = 0.f, max_dis_log = 0.f, rows_log = 0.f;
  float puniform = 0.f, puniform_sky = 0.f;
  float pnexists_given_sky_log = 0.f, nopnexists_given_sky_log = 0.f;
  float pnexists_given_ground_log = 0.f, nopnexists_given_ground_log = 0.f;
  float pnexists_given_object_log = 0.f, nopnexists_given_object_log = 0.f;
  float normalization_sky = 0.f, inv_sigma2_sky = 0.f;
  std::vector<float> log_lut;                  // 1e6+1 entries (Stixels.cu:79-84)
  std::vector<float> normalization_object, inv_sigma2_object;  // [D]
  std::vector<float> object_disparity_range;   // [D]   (:111-115)
  std::vector<float> obj_cost_lut;             // [D][D] (:122-129)
  std::vector<float> inverse_height;           // [H+1]: (float)(1.0 / n), StixelsKernels.cu:485,608

  void apply(const isx_config &c);             // SetConfig's fan-out (:315-337)
  void derive();                               // Initialize's host part
  // PrecomputeGround + GroundFunction (:790-817, 867-877) for one frame.
  // out = [ground_function | normalization_ground | inv_sigma2_ground], 3*rows floats.
  // Returns the flipped horizon m_vhor = rows - vhor - 1 (:377).
  int ground_tables(const isx_road &road, float *out) const;
};

// Throws nothing; returns an error text (empty = ok) where SetConfig throws
// std::invalid_argument (Stixels.cu:294-313).
std::string validate_config(const isx_config &c);

}  // namespace isx
