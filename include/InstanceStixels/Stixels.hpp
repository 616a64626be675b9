// Drop-in replacement for InstanceStixels/include/InstanceStixels/Stixels.hpp:40-96.
//
// `class Stixels` with the reference's public method names and signatures,
// implemented header-only on top of the C ABI (include/instance_stixels_b200.h,
// libinstance_stixels_b200.so).  apps/run_cityscapes.cu and
// apps/stixels_wrapper.cu compile against it unchanged: same include name,
// same call sequence, same exceptions (std::invalid_argument from SetConfig /
// Get3DVertices), CUDA failures print and exit(1) like CUDA_CHECK_RETURN.
// Host-only code: no CUDA headers are needed to include it.
#ifndef ISX_DROPIN_STIXELS_HPP_
#define ISX_DROPIN_STIXELS_HPP_

#include <stdint.h>

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../instance_stixels_b200.h"
#include "configuration.h"
#include "types.h"
#include "util.h"

constexpr float PIFLOAT = 3.1416f;

class Stixels {
public:
    Stixels() {
        if (isx_create(&h_, device_from_env()) != ISX_OK) die();
    }
    ~Stixels() { isx_destroy(h_); }
    Stixels(const Stixels&) = delete;
    Stixels& operator=(const Stixels&) = delete;

    // Extension: frames per batched call (the reference is always 1).
    void Initialize(int max_batch) { check(isx_initialize(h_, max_batch)); }
    void Initialize() { Initialize(1); }
    void Finish() { check(isx_finish(h_)); }
    bool IsInitialized() { return isx_is_initialized(h_) != 0; }

    float Compute(bool pairwise, StixelsData& stixels, int32_t* d_segmentation_local = nullptr) {
        stixels.sections.resize((size_t)GetRealCols() * GetMaxSections());
        isx_frame_meta m;
        check(isx_compute(h_, pairwise ? 1 : 0, reinterpret_cast<isx_section*>(stixels.sections.data()), &m,
                          d_segmentation_local));
        stixels.rows = m.rows;
        stixels.cols = m.cols;
        stixels.realcols = m.realcols;
        stixels.max_sections = m.max_sections;
        stixels.max_dis = m.max_dis;
        stixels.column_step = m.column_step;
        stixels.semantic_classes = m.semantic_classes;
        stixels.alpha_ground = m.alpha_ground;
        stixels.vhor = m.vhor;
        meta_ = m;
        return -1;  // the reference's timing code is commented out and it returns -1 too
    }
    float ClusterInstances() {
        check(isx_cluster_instances(h_));
        return -1;
    }
    std::map<std::pair<int, int>, int> GetInstanceStixels() {
        int n = 0;
        check(isx_get_instance_stixels(h_, nullptr, 0, &n));
        std::vector<isx_instance> v((size_t)(n > 0 ? n : 1));
        check(isx_get_instance_stixels(h_, v.data(), n, &n));
        std::map<std::pair<int, int>, int> out;
        for (int i = 0; i < n; i++) out[std::make_pair(v[i].column, v[i].index)] = v[i].label;
        return out;
    }
    int GetRealCols() { return isx_real_cols(h_); }
    int GetMaxSections() { return isx_max_sections(h_); }

    void SetConfig(const StixelConfig& c) {
        isx_config k;
        isx_config_init(&k);
        k.rows = c.rows; k.cols = c.cols; k.max_dis = c.max_dis; k.invalid_disparity = c.invalid_disparity;
        k.eps = c.eps; k.min_pts = c.min_pts; k.size_filter = c.size_filter;
        k.n_semantic_classes = c.n_semantic_classes; k.n_offset_channels = c.n_offset_channels;
        k.prior_weight = c.prior_weight; k.segmentation_weight = c.segmentation_weight;
        k.instance_weight = c.instance_weight; k.disparity_weight = c.disparity_weight;
        k.pairwise = c.pairwise; k.column_step = c.column_step;
        k.focal = c.focal; k.baseline = c.baseline;
        k.camera_center_x = c.camera_center_x; k.camera_center_y = c.camera_center_y;
        k.sigma_disparity_object = c.sigma_disparity_object; k.sigma_disparity_ground = c.sigma_disparity_ground;
        k.sigma_sky = c.sigma_sky;
        k.pout = c.pout; k.pout_sky = c.pout_sky; k.pord = c.pord; k.pgrav = c.pgrav; k.pblg = c.pblg;
        k.pground_given_nexist = c.pground_given_nexist; k.pobject_given_nexist = c.pobject_given_nexist;
        k.psky_given_nexist = c.psky_given_nexist; k.pnexist_dis = c.pnexist_dis;
        k.pground = c.pground; k.pobject = c.pobject; k.psky = c.psky;
        k.width_margin = c.width_margin;
        k.sigma_camera_tilt = c.sigma_camera_tilt; k.sigma_camera_height = c.sigma_camera_height;
        k.median_join = c.median_join; k.epsilon = c.epsilon; k.range_objects_z = c.range_objects_z;
        k.road_vdisparity_threshold = c.road_vdisparity_threshold;
        check(isx_set_config(h_, &k));
        cfg_ = c;
    }
    void SetSegmentation(const std::vector<int32_t>& segmentation) {
        check(isx_set_segmentation(h_, segmentation.data(), segmentation.size()));
    }
    void SetSegmentationParameters(const int classes, const int instance_channels) {
        check(isx_set_segmentation_parameters(h_, classes, instance_channels));
    }
    void SetClusteringParameters(const float eps, const int min_pts, const int size_filter) {
        check(isx_set_clustering_parameters(h_, eps, min_pts, size_filter));
    }
    void SetWeightParameters(const float prior_weight, const float disparity_weight,
                             const float segmentation_weight, const float instance_weight) {
        check(isx_set_weight_parameters(h_, prior_weight, disparity_weight, segmentation_weight, instance_weight));
    }
    void SetDisparityImage(const std::vector<pixel_t>& disp_im) {
        check(isx_set_disparity_image(h_, disp_im.data(), disp_im.size()));
    }
    pixel_t* GetInputDisparityImageOnDevice() { return isx_input_disparity_device(h_); }
    void SetProbabilities(float pout, float pout_sky, float pground_given_nexist, float pobject_given_nexist,
                          float psky_given_nexist, float pnexist_dis, float pground, float pobject, float psky,
                          float pord, float pgrav, float pblg) {
        check(isx_set_probabilities(h_, pout, pout_sky, pground_given_nexist, pobject_given_nexist,
                                    psky_given_nexist, pnexist_dis, pground, pobject, psky, pord, pgrav, pblg));
    }
    void SetRoadParameters(int vhor, float camera_tilt, float camera_height, float alpha_ground) {
        check(isx_set_road_parameters(h_, vhor, camera_tilt, camera_height, alpha_ground));
    }
    void SetCameraParameters(float focal, float baseline, float sigma_camera_tilt, float sigma_camera_height,
                             float camera_center_x = -1, float camera_center_y = -1) {
        check(isx_set_camera_parameters(h_, focal, baseline, sigma_camera_tilt, sigma_camera_height,
                                        camera_center_x, camera_center_y));
        cfg_.focal = focal; cfg_.baseline = baseline;
        cfg_.camera_center_x = camera_center_x; cfg_.camera_center_y = camera_center_y;
    }
    void SetDisparityParameters(const int rows, const int cols, const int max_dis, const float invalid_disparity,
                                const float sigma_disparity_object, const float sigma_disparity_ground,
                                float sigma_sky) {
        check(isx_set_disparity_parameters(h_, rows, cols, max_dis, invalid_disparity, sigma_disparity_object,
                                           sigma_disparity_ground, sigma_sky));
        cfg_.rows = rows;
    }
    void SetModelParameters(const int column_step, const bool median_join, float epsilon, float range_objects_z,
                            int width_margin) {
        check(isx_set_model_parameters(h_, column_step, median_join, epsilon, range_objects_z, width_margin));
        cfg_.column_step = column_step;
    }

    // Corner points of every stixel in camera coordinates, 4 x (x, y, z) per stixel, clockwise
    // from the top-left corner (Stixels.cu:683-742).  Pure function of the result.
    std::vector<float> Get3DVertices(const StixelsData& d) {
        if (cfg_.camera_center_x == -1 || cfg_.camera_center_y == -1)
            throw std::invalid_argument("Camera parameters are not set.");
        const float f = cfg_.focal, bf = cfg_.baseline * cfg_.focal;
        const float cx = cfg_.camera_center_x, cy = cfg_.camera_center_y;
        const int step = d.column_step;
        std::vector<float> v;
        for (size_t i = 0; i < (size_t)d.realcols; i++) {
            for (size_t j = 0; j < (size_t)d.max_sections; j++) {
                const Section& s = d.sections[i * d.max_sections + j];
                if (s.type == -1) break;
                const float xl = i * step, xr = xl + step;
                const float yt = d.rows - s.vT - 1, yb = d.rows - s.vB;
                float zt = 0.0, zb = 0.0;  // sky stays at depth 0
                if (s.type == OBJECT) {
                    zt = zb = bf / s.disparity;
                } else if (s.type == GROUND) {
                    zt = bf / (d.alpha_ground * (d.vhor - s.vT));
                    zb = bf / (d.alpha_ground * (d.vhor - s.vB));
                }
                const float corner[4][3] = {{xl, yt, zt}, {xr, yt, zt}, {xr, yb, zb}, {xl, yb, zb}};
                for (const auto& c : corner) {
                    v.push_back(-c[2] / f * (cx - c[0]));
                    v.push_back(-c[2] / f * (cy - c[1]));
                    v.push_back(c[2]);
                }
            }
        }
        return v;
    }

    // The `.stixels` text format read by tools/visualization (Stixels.cu:889-926): one line per
    // column, "type,vB,vT,disparity,class,cost,meanx,meany[,label];" per stixel, then the ground plane.
    static void SaveStixels(Section* stixels, std::map<std::pair<int, int>, int> instance_stixels,
                            const float alpha_ground, const int vhor, const int real_cols,
                            const int max_segments, const char* fname) {
        std::ofstream fp(fname, std::ofstream::out | std::ofstream::trunc);
        if (!fp.is_open()) {
            std::cerr << "Counldn't write file: " << fname << std::endl;
            return;
        }
        for (size_t i = 0; i < (size_t)real_cols; i++) {
            for (size_t j = 0; j < (size_t)max_segments; j++) {
                const Section& s = stixels[i * max_segments + j];
                if (s.type == -1) break;
                fp << s.type << "," << s.vB << "," << s.vT << "," << s.disparity << "," << s.semantic_class << ","
                   << s.cost << "," << s.instance_meanx << "," << s.instance_meany;
                const auto it = instance_stixels.find(std::make_pair((int)i, (int)j));
                if (it != instance_stixels.end()) fp << "," << it->second;
                fp << ";";
            }
            fp << std::endl;
        }
        fp << "groundplane" << alpha_ground << "," << vhor << "\n";
    }

    isx_handle handle() { return h_; }  // for the batched C entry points

private:
    isx_handle h_ = nullptr;
    StixelConfig cfg_;
    isx_frame_meta meta_{};

    static int device_from_env() {
        const char* e = std::getenv("ISX_DEVICE");
        return e ? std::atoi(e) : 0;
    }
    [[noreturn]] void die() {
        std::cerr << "instance_stixels_b200: " << isx_last_error(h_) << std::endl;
        std::exit(1);
    }
    void check(int rc) {
        if (rc == ISX_OK) return;
        if (rc == ISX_ERR_INVALID_ARGUMENT) throw std::invalid_argument(isx_last_error(h_));
        die();
    }
};

#endif  // ISX_DROPIN_STIXELS_HPP_
