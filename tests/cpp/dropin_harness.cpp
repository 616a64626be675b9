// Replays the call sequence of the reference's batch runner (apps/run_cityscapes.cu:185-197,
// 328-343, 346, 383-387, 406-411, 430-449) against the drop-in `class Stixels`, with the file
// loaders replaced by raw binary inputs written by the test:
//   dropin_harness <disparity.f32> <segmentation.i32> <rows> <cols> <pairwise> <vhor> <alpha> <out.stixels>
// With vhor = -1 the road parameters come from the drop-in `class RoadEstimation`, like
// apps/run_cityscapes.cu:332-343, 390-407, and a line "road <vhor> <pitch> <height> <slope>" is printed.
// Prints "sections <n> instances <m>"; exit code 0 on success.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "RoadEstimation.h"   // RoadEstimation
#include "Stixels.hpp"        // Stixels
#include "configuration.h"    // pixel_t

template <typename T>
static std::vector<T> read_all(const char* path, size_t n) {
    std::vector<T> v(n);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
    if (!f) { std::cerr << "short read: " << path << "\n"; std::exit(2); }
    return v;
}

int main(int argc, char* argv[]) {
    if (argc < 9) return 2;
    const int rows = atoi(argv[3]), cols = atoi(argv[4]);
    const bool pairwise = atoi(argv[5]);
    int vhorizon_point = atoi(argv[6]);
    float alpha_ground = atof(argv[7]);
    float camera_tilt = 0.0f, camera_height = 1.18f;

    StixelConfig stixel_config;
    stixel_config.column_step = 8;
    stixel_config.max_dis = 128;
    stixel_config.invalid_disparity = 0.0f;
    stixel_config.n_semantic_classes = 19;
    stixel_config.n_offset_channels = 2;
    stixel_config.prior_weight = pairwise ? 1 : 1e4;
    stixel_config.segmentation_weight = pairwise ? 4.709500548254913f : 11.241965032069425f;
    stixel_config.instance_weight = pairwise ? 0.0031312903639774976f : 0.0017313017435431333f;
    stixel_config.disparity_weight = pairwise ? 0.0001f : 0.0069935800364145494f;
    stixel_config.eps = pairwise ? 18.82232269133926f : 23.89408062110343f;
    stixel_config.min_pts = pairwise ? 3 : 4;
    stixel_config.size_filter = pairwise ? 25 : 42;

    Stixels stixels;
    StixelsData stixels_data;
    bool threw = false;
    try { stixels.SetConfig(stixel_config); } catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { std::cerr << "SetConfig accepted an incomplete config\n"; return 3; }
    stixel_config.rows = rows;
    stixel_config.cols = cols;
    stixel_config.baseline = 0.209313f;
    stixel_config.focal = 2262.52f;
    stixel_config.camera_center_x = 1024.0f;
    stixel_config.camera_center_y = 512.0f;
    stixel_config.pground = stixel_config.pobject = stixel_config.psky = 0.33f;

    if (stixels.IsInitialized()) stixels.Finish();
    stixels.SetConfig(stixel_config);
    stixels.Initialize();

    const int hs2 = 1 << (int)std::ceil(std::log2(rows / 8 + 1));
    const auto disparity_img = read_all<pixel_t>(argv[1], (size_t)rows * cols);
    const auto segmentation = read_all<int32_t>(argv[2], (size_t)(cols / 8) * 21 * hs2);
    stixels.SetDisparityImage(disparity_img);
    stixels.SetSegmentation(segmentation);
    if (vhorizon_point < 0) {
        RoadEstimation road_estimation;
        if (road_estimation.IsInitialized()) road_estimation.Finish();
        road_estimation.Initialize(stixel_config.camera_center_y, stixel_config.baseline, stixel_config.focal,
                                   stixel_config.rows, stixel_config.cols, stixel_config.max_dis,
                                   stixel_config.road_vdisparity_threshold);
        const bool ok = road_estimation.Compute(disparity_img);
        if (!ok) { std::printf("Road estimation failed.\n"); return 5; }
        // the overload the ROS node uses: the image Stixels already holds on the device
        RoadEstimation on_device;
        on_device.Initialize(stixel_config.camera_center_y, stixel_config.baseline, stixel_config.focal,
                             stixel_config.rows, stixel_config.cols, stixel_config.max_dis);
        if (!on_device.Compute(stixels.GetInputDisparityImageOnDevice()) ||
            on_device.GetHorizonPoint() != road_estimation.GetHorizonPoint() ||
            on_device.GetSlope() != road_estimation.GetSlope()) { std::cerr << "device overload differs\n"; return 6; }
        on_device.Finish();
        camera_tilt = road_estimation.GetPitch();
        camera_height = road_estimation.GetCameraHeight();
        vhorizon_point = road_estimation.GetHorizonPoint();
        alpha_ground = road_estimation.GetSlope();
        std::printf("road %d %.9g %.9g %.9g\n", vhorizon_point, camera_tilt, camera_height, alpha_ground);
        road_estimation.Finish();
    }
    stixels.SetRoadParameters(vhorizon_point, camera_tilt, camera_height, alpha_ground);
    stixels.Compute(pairwise, stixels_data);

    Section* stx = stixels_data.sections.data();
    std::map<std::pair<int, int>, int> instance_stixels = stixels.GetInstanceStixels();
    Stixels::SaveStixels(stx, instance_stixels, alpha_ground, stixel_config.rows - 1 - vhorizon_point,
                         stixels.GetRealCols(), stixels.GetMaxSections(), argv[8]);
    size_t n = 0;
    for (int i = 0; i < stixels.GetRealCols(); i++)
        for (int j = 0; j < stixels.GetMaxSections(); j++) {
            if (stx[i * stixels.GetMaxSections() + j].type == -1) break;
            n++;
        }
    const std::vector<float> vertices = stixels.Get3DVertices(stixels_data);
    if (vertices.size() != n * 12) { std::cerr << "Get3DVertices size\n"; return 4; }
    std::printf("sections %zu instances %zu\n", n, instance_stixels.size());
    if (stixels.IsInitialized()) stixels.Finish();
    return 0;
}
