// Drop-in replacement header: the plain-data types that cross the library
// boundary.  Member names, types, order and defaults are those of
// InstanceStixels/include/InstanceStixels/types.h:22-205, because the callers
// (apps/run_cityscapes.cu:185-197, apps/stixels_wrapper.cu:23-50,
// apps/stixels_node.cu:94-100) set and read them by name and index
// `sections[col * max_sections + j]`.  StixelParameters (the reference's
// kernel argument block) is an implementation detail and has no counterpart.
#ifndef ISX_DROPIN_TYPES_H_
#define ISX_DROPIN_TYPES_H_

#include <vector>

#include "../instance_stixels_b200.h"

constexpr int GROUND = ISX_GROUND;
constexpr int OBJECT = ISX_OBJECT;
constexpr int SKY = ISX_SKY;

// -1 marks "must be set by the caller"; SetConfig throws std::invalid_argument otherwise.
struct StixelConfig {
    // image
    float rows = -1, cols = -1;
    int max_dis = -1;
    float invalid_disparity = -1.0f;  // >= 0: that value marks holes; < 0: none
    // DBSCAN instance grouping
    float eps = -1;
    int min_pts = -1, size_filter = -1;
    // CNN output
    int n_semantic_classes = -1, n_offset_channels = -1;
    // energy weights
    float prior_weight = -1, segmentation_weight = -1, instance_weight = -1, disparity_weight = -1;
    bool pairwise = false;            // storage only; Compute() takes the flag
    int column_step = -1;             // stixel width
    // camera
    float focal = -1, baseline = -1, camera_center_x = -1, camera_center_y = -1;
    // disparity model
    float sigma_disparity_object = 1.0f, sigma_disparity_ground = 2.0f, sigma_sky = 0.1f;
    // probabilities
    float pout = 0.15f, pout_sky = 0.4f, pord = 0.2f, pgrav = 0.1f, pblg = 0.04f;
    float pground_given_nexist = 0.28, pobject_given_nexist = 0.44, psky_given_nexist = 0.28;
    float pnexist_dis = 0.25f;
    float pground = 1.0f / 3.0f, pobject = 1.0f / 3.0f, psky = 1.0f / 3.0f;
    int width_margin = 0;
    float sigma_camera_tilt = 0.05f, sigma_camera_height = 0.05f;
    bool median_join = false;
    float epsilon = 3.0f, range_objects_z = 10.20f;
    float road_vdisparity_threshold = 0.2f;
};

// One stixel; a column is a run of these ended by type == -1.  Same layout as isx_section.
struct Section {
    int type;
    int vB, vT;
    float disparity;
    int semantic_class;
    float cost;
    float instance_meanx;
    float instance_meany;
};
static_assert(sizeof(Section) == sizeof(isx_section), "Section must match the C ABI record");

struct StixelsData {
    std::vector<Section> sections;
    int rows, cols;
    int realcols, max_sections, max_dis;
    int column_step;
    int semantic_classes;
    float alpha_ground;
    int vhor;
};

#endif  // ISX_DROPIN_TYPES_H_
