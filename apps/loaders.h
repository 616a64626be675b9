// File loaders of the run_cityscapes harness (apps/run_cityscapes.cpp), without OpenCV / rapidjson / HDF5:
//   * read_png_gray  : what cv::imread(IMREAD_UNCHANGED) gives readDisparityImage for a Cityscapes disparity PNG
//                      (apps/run_cityscapes.cu:109-152): 8- or 16-bit grayscale, non-interlaced; zlib for the inflate.
//   * load_camera    : the three numbers LoadCameraFile takes from the camera JSON (apps/run_cityscapes.cu:51-79):
//                      extrinsic.baseline, intrinsic.fy, intrinsic.v0.
//   * load_npy_int32 : the segmentation tensor.  The reference reads the int32 dataset "nlogprobs" of
//                      <base>_probs.h5 (H5Segmentation.cpp:25-49); HDF5 is not available in this image, so the
//                      harness reads the same array saved as <base>_probs.npy (tools/h5_to_npy.py converts).
#ifndef ISX_APPS_LOADERS_H_
#define ISX_APPS_LOADERS_H_

#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace isx_apps {

inline std::vector<unsigned char> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::invalid_argument("Couldn't read the file " + path);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

struct GrayImage {
    int rows = 0, cols = 0, bit_depth = 0;
    std::vector<uint16_t> pixels;  // row-major, top-left origin; 8-bit images are widened
};

inline uint32_t be32(const unsigned char* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

inline GrayImage read_png_gray(const std::string& path) {
    const std::vector<unsigned char> buf = read_file(path);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) throw std::invalid_argument(path + " is not a PNG file");
    GrayImage im;
    int color_type = -1, interlace = 0;
    std::vector<unsigned char> idat;
    size_t pos = 8;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const std::string type(reinterpret_cast<const char*>(&buf[pos + 4]), 4);
        if (pos + 12 + (size_t)len > buf.size()) throw std::invalid_argument(path + ": truncated PNG chunk");
        const unsigned char* data = &buf[pos + 8];
        if (type == "IHDR") {
            if (len < 13) throw std::invalid_argument(path + ": bad IHDR");
            im.cols = (int)be32(data);
            im.rows = (int)be32(data + 4);
            im.bit_depth = data[8];
            color_type = data[9];
            interlace = data[12];
        } else if (type == "IDAT") {
            idat.insert(idat.end(), data, data + len);
        } else if (type == "IEND") {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (color_type != 0 || (im.bit_depth != 8 && im.bit_depth != 16) || interlace != 0)
        throw std::invalid_argument(path + ": only non-interlaced 8/16-bit grayscale PNGs are supported "
                                           "(the Cityscapes disparity format)");
    const size_t bpp = im.bit_depth / 8, stride = (size_t)im.cols * bpp;
    std::vector<unsigned char> raw((stride + 1) * (size_t)im.rows);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
        throw std::invalid_argument(path + ": zlib could not inflate the image data");
    // undo the scanline filters (PNG specification, section 9)
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    im.pixels.resize((size_t)im.rows * im.cols);
    for (int r = 0; r < im.rows; r++) {
        const unsigned char* line = &raw[(stride + 1) * (size_t)r];
        const int filter = line[0];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int x = line[1 + i];
            switch (filter) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) / 2; break;
                case 4: {
                    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    x += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw std::invalid_argument(path + ": unknown PNG filter type");
            }
            cur[i] = (unsigned char)x;
        }
        for (int col = 0; col < im.cols; col++)
            im.pixels[(size_t)r * im.cols + col] =
                bpp == 1 ? cur[col] : (uint16_t)((cur[2 * col] << 8) | cur[2 * col + 1]);  // big endian samples
        prev.swap(cur);
    }
    return im;
}

// The number that follows "key": inside the object "section" of a (Cityscapes camera) JSON document.
inline bool json_number(const std::string& doc, const std::string& section, const std::string& key, double* out) {
    const size_t s = doc.find("\"" + section + "\"");
    if (s == std::string::npos) return false;
    const size_t open = doc.find('{', s);
    if (open == std::string::npos) return false;
    int depth = 0;
    size_t close = open;
    for (; close < doc.size(); close++) {
        if (doc[close] == '{') depth++;
        if (doc[close] == '}' && --depth == 0) break;
    }
    const size_t k = doc.find("\"" + key + "\"", open);
    if (k == std::string::npos || k > close) return false;
    const size_t colon = doc.find(':', k);
    if (colon == std::string::npos) return false;
    char* end = nullptr;
    const double v = std::strtod(doc.c_str() + colon + 1, &end);
    if (end == doc.c_str() + colon + 1) return false;
    *out = v;
    return true;
}

struct Camera {
    float baseline = 0, focal = 0, center_y = 0;
    bool from_file = false;
};

inline Camera load_camera(const std::string& path) {
    Camera cam;
    std::ifstream f(path);
    if (f) {
        std::stringstream ss;
        ss << f.rdbuf();
        const std::string doc = ss.str();
        double b = 0, fy = 0, v0 = 0;
        if (!json_number(doc, "extrinsic", "baseline", &b) || !json_number(doc, "intrinsic", "fy", &fy) ||
            !json_number(doc, "intrinsic", "v0", &v0))
            throw std::invalid_argument(path + ": extrinsic.baseline / intrinsic.fy / intrinsic.v0 not found");
        cam.baseline = (float)b;
        cam.focal = (float)fy;
        cam.center_y = (float)v0;
        cam.from_file = true;
    } else {
        // the reference's fallback: UEYE parameters (apps/run_cityscapes.cu:66-77)
        const float size_factor = 1000. / 1216.;
        cam.focal = 1495.46f;
        cam.baseline = 0.22087f;
        cam.center_y = 624.896 * size_factor;
    }
    return cam;
}

struct NpyInt32 {
    std::vector<size_t> shape;
    std::vector<int32_t> data;
};

inline NpyInt32 load_npy_int32(const std::string& path) {
    const std::vector<unsigned char> buf = read_file(path);
    if (buf.size() < 10 || std::memcmp(buf.data(), "\x93NUMPY", 6) != 0) throw std::invalid_argument(path + " is not an .npy file");
    const int major = buf[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = buf[8] | (buf[9] << 8); hoff = 10; }
    else { hlen = buf[8] | (buf[9] << 8) | ((size_t)buf[10] << 16) | ((size_t)buf[11] << 24); hoff = 12; }
    if (hoff + hlen > buf.size()) throw std::invalid_argument(path + ": truncated .npy header");
    const std::string header(reinterpret_cast<const char*>(&buf[hoff]), hlen);
    if (header.find("'<i4'") == std::string::npos && header.find("'|i4'") == std::string::npos)
        throw std::invalid_argument(path + ": expected little-endian int32 data ('<i4')");
    if (header.find("'fortran_order': False") == std::string::npos)
        throw std::invalid_argument(path + ": expected C order");
    NpyInt32 out;
    const size_t sp = header.find("'shape':");
    const size_t lp = header.find('(', sp), rp = header.find(')', lp);
    if (sp == std::string::npos || lp == std::string::npos || rp == std::string::npos)
        throw std::invalid_argument(path + ": no shape in the .npy header");
    size_t n = 1;
    const char* p = header.c_str() + lp + 1;
    while (p < header.c_str() + rp) {
        char* end = nullptr;
        const unsigned long long v = std::strtoull(p, &end, 10);
        if (end == p) { p++; continue; }
        out.shape.push_back((size_t)v);
        n *= (size_t)v;
        p = end;
    }
    if (hoff + hlen + n * 4 > buf.size()) throw std::invalid_argument(path + ": fewer values than the shape says");
    out.data.resize(n);
    std::memcpy(out.data.data(), &buf[hoff + hlen], n * 4);
    return out;
}

}  // namespace isx_apps

#endif  // ISX_APPS_LOADERS_H_
