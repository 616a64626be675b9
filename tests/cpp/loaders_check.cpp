// Prints what apps/loaders.h reads from a file, for tests/test_cityscapes_runner.py:
//   loaders_check png <file>     -> "rows cols bit_depth sum_of_pixels first last"
//   loaders_check camera <file>  -> "baseline focal center_y from_file" (9 significant digits)
//   loaders_check npy <file>     -> "ndim d0 d1 ... sum"
#include <cstdio>
#include <cstring>
#include <iostream>

#include "loaders.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    try {
        if (!std::strcmp(argv[1], "png")) {
            const isx_apps::GrayImage im = isx_apps::read_png_gray(argv[2]);
            unsigned long long sum = 0;
            for (uint16_t v : im.pixels) sum += v;
            std::printf("%d %d %d %llu %u %u\n", im.rows, im.cols, im.bit_depth, sum, im.pixels.front(), im.pixels.back());
        } else if (!std::strcmp(argv[1], "camera")) {
            const isx_apps::Camera c = isx_apps::load_camera(argv[2]);
            std::printf("%.9g %.9g %.9g %d\n", c.baseline, c.focal, c.center_y, (int)c.from_file);
        } else if (!std::strcmp(argv[1], "npy")) {
            const isx_apps::NpyInt32 a = isx_apps::load_npy_int32(argv[2]);
            long long sum = 0;
            for (int32_t v : a.data) sum += v;
            std::printf("%zu", a.shape.size());
            for (size_t d : a.shape) std::printf(" %zu", d);
            std::printf(" %lld\n", sum);
        } else {
            return 2;
        }
    } catch (const std::invalid_argument& e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    return 0;
}
