// Test helper: h5_check <file> <dataset> <out.bin> [raw]  -- reads a dataset with apps/h5_reader.h and dumps
// "rank dims... " on stdout and the values (int32, or the raw element bytes with `raw`) into <out.bin>.
#include <cstdio>
#include <fstream>
#include <iostream>

#include "h5_reader.h"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    try {
        std::ofstream out(argv[3], std::ios::binary);
        if (argc > 4) {
            const isx_apps::H5Raw r = isx_apps::load_h5_raw(argv[1], argv[2]);
            std::cout << r.shape.size();
            for (size_t d : r.shape) std::cout << " " << d;
            std::cout << " class " << r.type_class << " elem " << r.elem << " signed " << r.is_signed << " be " << r.big_endian << "\n";
            out.write(reinterpret_cast<const char*>(r.bytes.data()), (std::streamsize)r.bytes.size());
        } else {
            const isx_apps::H5Int32 r = isx_apps::load_h5_int32(argv[1], argv[2]);
            std::cout << r.shape.size();
            for (size_t d : r.shape) std::cout << " " << d;
            std::cout << "\n";
            out.write(reinterpret_cast<const char*>(r.data.data()), (std::streamsize)(r.data.size() * 4));
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
