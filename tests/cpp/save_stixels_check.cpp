// Writes a .stixels file with the drop-in Stixels::SaveStixels (include/InstanceStixels/Stixels.hpp) from raw
// inputs prepared by tests/test_stixels_file_format.py:
//   save_stixels_check <sections.bin> <real_cols> <max_segments> <instances.i32: (column, index, label) triples>
//                      <alpha_ground> <vhor> <out.stixels>
#include <cstdlib>
#include <fstream>
#include <map>
#include <vector>

#include "Stixels.hpp"

int main(int argc, char** argv) {
    if (argc < 8) return 2;
    const int real_cols = std::atoi(argv[2]), max_segments = std::atoi(argv[3]);
    std::vector<Section> sections((size_t)real_cols * max_segments);
    std::ifstream f(argv[1], std::ios::binary);
    f.read(reinterpret_cast<char*>(sections.data()), sections.size() * sizeof(Section));
    if (!f) return 3;
    std::map<std::pair<int, int>, int> instances;
    std::ifstream g(argv[4], std::ios::binary);
    int t[3];
    while (g.read(reinterpret_cast<char*>(t), sizeof t)) instances[{t[0], t[1]}] = t[2];
    Stixels::SaveStixels(sections.data(), instances, (float)std::atof(argv[5]), std::atoi(argv[6]), real_cols,
                         max_segments, argv[7]);
    return 0;
}
