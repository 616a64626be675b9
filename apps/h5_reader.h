// Minimal HDF5 reader for the one thing the run_cityscapes harness needs from libhdf5: an N-dimensional integer
// dataset by name from the root group, converted to int32 -- what H5Segmentation::LoadFile does with "nlogprobs"
// (InstanceStixels/src/H5Segmentation.cpp:25-49: openDataSet, getSimpleExtentDims, read(..., NATIVE_INT)); the files
// are written by h5py `create_dataset('nlogprobs', data=output)` (tools/CNN_training/inference.py:454-455).
// No libhdf5: the file format is walked directly (HDF5 File Format Specification 3.0):
//   superblock        version 0/1 (symbol-table root group) and 2/3 (object-header root group), at offset 0, 512, ...
//   root group        v1 object header -> symbol table message -> B-tree v1 (TREE) -> symbol nodes (SNOD) + local
//                     heap (HEAP); or v2 object header (OHDR) with compact link messages
//   dataset header    v1 or v2 object header incl. continuation blocks: dataspace (v1/v2), datatype (fixed-point,
//                     1/2/4/8 bytes, little or big endian, signed or unsigned), data layout v1-v3 (compact, contiguous,
//                     chunked through a v1 B-tree) and the filter pipeline (none, deflate, shuffle + deflate)
// Not supported (reported by name): dense link storage (fractal heaps), layout version 4, other filters, non-integer
// types.
#ifndef ISX_APPS_H5_READER_H_
#define ISX_APPS_H5_READER_H_

#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace isx_apps {

struct H5Int32 {
    std::vector<size_t> shape;
    std::vector<int32_t> data;
};

// A dataset as stored: element bytes in file order (after undoing chunking and filters).
struct H5Raw {
    std::vector<size_t> shape;
    int type_class = -1;   // 0 fixed-point, 1 floating-point (HDF5 datatype classes)
    int elem = 0;          // bytes per element
    bool is_signed = true, big_endian = false;
    std::vector<unsigned char> bytes;
};

namespace h5detail {

struct File {
    std::vector<unsigned char> buf;
    std::string path;
    size_t base = 0;      // superblock offset (user block)
    int so = 8, sl = 8;   // size of offsets / lengths
    [[noreturn]] void fail(const std::string& what) const { throw std::invalid_argument(path + ": " + what); }
    const unsigned char* at(uint64_t addr, size_t n) const {
        const uint64_t a = addr + base;
        if (a + n > buf.size() || a + n < a) fail("address outside the file");
        return buf.data() + a;
    }
    uint64_t le(uint64_t addr, int n) const {
        const unsigned char* p = at(addr, (size_t)n);
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; i--) v = (v << 8) | p[i];
        return v;
    }
    bool undefined(uint64_t a, int n) const { return n == 8 ? a == ~0ull : a == ((1ull << (8 * n)) - 1); }
};

struct Message {
    int type;
    uint64_t addr;  // of the message data
    size_t size;
};

// All messages of an object header (either version), continuation blocks included.
inline std::vector<Message> object_messages(const File& f, uint64_t addr) {
    std::vector<Message> out;
    std::vector<std::pair<uint64_t, uint64_t>> blocks;  // (address, length) of message blocks still to read
    const bool v2 = std::memcmp(f.at(addr, 4), "OHDR", 4) == 0;
    int flags = 0;
    if (!v2) {
        if (f.le(addr, 1) != 1) f.fail("object header version not 1 or 2");
        const uint64_t header_size = f.le(addr + 8, 4);
        blocks.push_back({addr + 16, header_size});  // 12 bytes of prefix, messages are 8-byte aligned
    } else {
        if (f.le(addr + 4, 1) != 2) f.fail("OHDR version not 2");
        flags = (int)f.le(addr + 5, 1);
        uint64_t p = addr + 6;
        if (flags & 0x20) p += 16;  // four time stamps
        if (flags & 0x10) p += 4;   // max compact / min dense attributes
        const int szbytes = 1 << (flags & 3);
        const uint64_t chunk0 = f.le(p, szbytes);
        p += szbytes;
        blocks.push_back({p, chunk0});
    }
    for (size_t b = 0; b < blocks.size(); b++) {
        uint64_t p = blocks[b].first;
        const uint64_t end = p + blocks[b].second;
        if (v2 && b > 0) {
            if (std::memcmp(f.at(p, 4), "OCHK", 4) != 0) f.fail("bad object header continuation block");
            p += 4;
        }
        const uint64_t tail = v2 ? (b > 0 ? 4 : 0) : 0;  // checksum at the end of a v2 continuation block
        while (p + (v2 ? 4 : 8) <= end - tail) {
            int type;
            size_t size;
            if (!v2) {
                type = (int)f.le(p, 2);
                size = (size_t)f.le(p + 2, 2);
                p += 8;
            } else {
                type = (int)f.le(p, 1);
                size = (size_t)f.le(p + 1, 2);
                p += 4 + ((flags & 0x04) ? 2 : 0);
            }
            if (p + size > end) break;
            if (type == 0x10) {  // continuation: offset, length
                blocks.push_back({f.le(p, f.so), f.le(p + f.so, f.sl)});
            } else if (type != 0 || !v2) {
                out.push_back({type, p, size});
            }
            p += size;
            if (!v2) p = (p + 7) & ~7ull;  // (message data is already padded in v1 files; keep alignment anyway)
        }
    }
    return out;
}

// Object header address of `name` in a symbol-table group (B-tree v1 + local heap).
inline uint64_t find_in_symbol_table(const File& f, uint64_t btree, uint64_t heap, const std::string& name) {
    if (std::memcmp(f.at(heap, 4), "HEAP", 4) != 0) f.fail("bad local heap");
    const uint64_t heap_data = f.le(heap + 8 + 2 * f.sl, f.so);
    std::vector<uint64_t> todo{btree};
    while (!todo.empty()) {
        const uint64_t node = todo.back();
        todo.pop_back();
        if (std::memcmp(f.at(node, 4), "TREE", 4) == 0) {
            if (f.le(node + 4, 1) != 0) f.fail("group B-tree of the wrong type");
            const int used = (int)f.le(node + 6, 2);
            uint64_t p = node + 8 + 2 * f.so;  // past the sibling addresses
            for (int i = 0; i < used; i++) {
                p += f.sl;  // key
                todo.push_back(f.le(p, f.so));
                p += f.so;
            }
        } else if (std::memcmp(f.at(node, 4), "SNOD", 4) == 0) {
            const int n = (int)f.le(node + 6, 2);
            uint64_t p = node + 8;
            for (int i = 0; i < n; i++) {
                const uint64_t name_off = f.le(p, f.so);
                const uint64_t header = f.le(p + f.so, f.so);
                const char* s = reinterpret_cast<const char*>(f.at(heap_data + name_off, 1));
                const size_t room = f.buf.size() - (size_t)(heap_data + name_off + f.base);
                if (strnlen(s, room) == name.size() && std::memcmp(s, name.data(), name.size()) == 0) return header;
                p += 2 * f.so + 4 + 4 + 16;
            }
        } else {
            f.fail("unexpected node in the group B-tree");
        }
    }
    f.fail("no dataset named \"" + name + "\" in the root group");
}

inline std::vector<unsigned char> inflate_all(const File& f, const unsigned char* src, size_t n, size_t expect) {
    std::vector<unsigned char> out(expect);
    uLongf len = (uLongf)expect;
    if (uncompress(out.data(), &len, src, (uLong)n) != Z_OK || len != expect) f.fail("deflate-compressed chunk does not inflate");
    return out;
}

}  // namespace h5detail

inline H5Raw load_h5_raw(const std::string& path, const std::string& dataset) {
    using namespace h5detail;
    File f;
    f.path = path;
    {
        std::ifstream in(path, std::ios::binary);
        if (!in) throw std::invalid_argument("Couldn't read the file " + path);
        f.buf.assign((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    }
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', 0x0d, 0x0a, 0x1a, 0x0a};
    size_t sb = 0;
    bool found = false;
    for (sb = 0; sb + 8 <= f.buf.size(); sb = sb ? sb * 2 : 512)
        if (std::memcmp(f.buf.data() + sb, sig, 8) == 0) { found = true; break; }
    if (!found) f.fail("not an HDF5 file");
    const int version = f.buf[sb + 8];
    uint64_t root_header = 0, root_btree = 0, root_heap = 0;
    bool have_table = false;
    f.base = 0;  // the superblock itself is read at absolute positions; every address in it is relative to the base
                 // address it names (the superblock's own position when the file has a user block)
    if (version <= 1) {
        f.so = f.buf[sb + 13];
        f.sl = f.buf[sb + 14];
        uint64_t p = sb + 24 + (version == 1 ? 4 : 0);
        const uint64_t base_address = f.le(p, f.so);
        p += 4 * (uint64_t)f.so;  // base, free-space, end-of-file, driver-information addresses
        // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
        root_header = f.le(p + f.so, f.so);
        const uint64_t cache_type = f.le(p + 2 * f.so, 4);
        if (cache_type == 1) {
            root_btree = f.le(p + 2 * f.so + 8, f.so);
            root_heap = f.le(p + 3 * f.so + 8, f.so);
            have_table = true;
        }
        f.base = (size_t)(f.undefined(base_address, f.so) ? sb : base_address);
    } else if (version <= 3) {
        f.so = f.buf[sb + 9];
        f.sl = f.buf[sb + 10];
        const uint64_t base_address = f.le(sb + 12, f.so);
        root_header = f.le(sb + 12 + 3 * (uint64_t)f.so, f.so);
        f.base = (size_t)(f.undefined(base_address, f.so) ? sb : base_address);
    } else {
        f.fail("unsupported superblock version");
    }
    if (f.so != 8 && f.so != 4) f.fail("unsupported size of offsets");

    // ---- locate the dataset's object header ----
    uint64_t ds_header = 0;
    if (!have_table) {
        bool located = false;
        for (const Message& m : object_messages(f, root_header)) {
            if (m.type == 0x11) {  // symbol table message: B-tree, local heap
                root_btree = f.le(m.addr, f.so);
                root_heap = f.le(m.addr + f.so, f.so);
                have_table = true;
            } else if (m.type == 0x06) {  // link message
                uint64_t p = m.addr;
                if (f.le(p, 1) != 1) f.fail("link message version");
                const int lf = (int)f.le(p + 1, 1);
                p += 2;
                int link_type = 0;
                if (lf & 0x08) link_type = (int)f.le(p++, 1);
                if (lf & 0x04) p += 8;   // creation order
                if (lf & 0x10) p += 1;   // character set
                const int lsz = 1 << (lf & 3);
                const uint64_t len = f.le(p, lsz);
                p += lsz;
                const std::string name(reinterpret_cast<const char*>(f.at(p, (size_t)len)), (size_t)len);
                p += len;
                if (name == dataset && link_type == 0) {
                    ds_header = f.le(p, f.so);
                    located = true;
                }
            } else if (m.type == 0x02 && !located) {
                // link info: a fractal heap address other than "undefined" means dense link storage
                uint64_t p = m.addr + 2 + ((f.le(m.addr + 1, 1) & 1) ? 8 : 0);
                if (!f.undefined(f.le(p, f.so), f.so)) f.fail("dense link storage (fractal heap) is not supported");
            }
        }
        if (!located && !have_table) f.fail("no dataset named \"" + dataset + "\" in the root group");
        if (located) have_table = false;
    }
    if (!ds_header) ds_header = find_in_symbol_table(f, root_btree, root_heap, dataset);

    // ---- dataspace, datatype, layout, filters ----
    H5Raw out;
    int elem = 0;
    bool is_signed = true, big_endian = false, have_type = false, have_layout = false;
    int layout_class = -1;
    uint64_t data_addr = 0, data_size = 0, chunk_btree = 0;
    std::vector<uint64_t> chunk_dims;
    const unsigned char* compact = nullptr;
    std::vector<int> filters;
    for (const Message& m : object_messages(f, ds_header)) {
        if (m.type == 0x01) {  // dataspace
            const int v = (int)f.le(m.addr, 1), rank = (int)f.le(m.addr + 1, 1);
            uint64_t p = m.addr + (v == 1 ? 8 : 4);
            out.shape.clear();
            for (int i = 0; i < rank; i++) out.shape.push_back((size_t)f.le(p + (uint64_t)i * f.sl, f.sl));
        } else if (m.type == 0x03) {  // datatype
            const int cv = (int)f.le(m.addr, 1);
            out.type_class = cv & 0x0f;
            if (out.type_class > 1) f.fail("dataset \"" + dataset + "\" is neither integer nor floating-point");
            const int bits0 = (int)f.le(m.addr + 1, 1);
            big_endian = bits0 & 1;
            is_signed = (bits0 & 8) != 0;
            elem = (int)f.le(m.addr + 4, 4);
            if (elem != 1 && elem != 2 && elem != 4 && elem != 8) f.fail("unsupported integer size");
            have_type = true;
        } else if (m.type == 0x08) {  // data layout
            const int v = (int)f.le(m.addr, 1);
            if (v == 1 || v == 2) {
                // HDF5 1.6 and older: version, dimensionality, class, 5 reserved, [address], dimensions (4 bytes each;
                // for chunked storage the chunk shape with the element size last), [compact: size + data]
                const int nd = (int)f.le(m.addr + 1, 1);
                layout_class = (int)f.le(m.addr + 2, 1);
                uint64_t p = m.addr + 8;
                if (layout_class != 0) {
                    (layout_class == 2 ? chunk_btree : data_addr) = f.le(p, f.so);
                    p += f.so;
                }
                for (int i = 0; i < nd; i++) chunk_dims.push_back(f.le(p + 4ull * i, 4));
                p += 4ull * nd;
                if (layout_class == 0) {
                    data_size = f.le(p, 4);
                    compact = f.at(p + 4, (size_t)data_size);
                } else if (layout_class == 1) {
                    data_size = ~0ull;   // implied by the dataspace
                    chunk_dims.clear();
                } else if (layout_class != 2) {
                    f.fail("unknown layout class");
                }
                have_layout = true;
                continue;
            }
            if (v != 3) f.fail("data layout message version " + std::to_string(v) + " is not supported");
            layout_class = (int)f.le(m.addr + 1, 1);
            if (layout_class == 0) {
                data_size = f.le(m.addr + 2, 2);
                compact = f.at(m.addr + 4, (size_t)data_size);
            } else if (layout_class == 1) {
                data_addr = f.le(m.addr + 2, f.so);
                data_size = f.le(m.addr + 2 + f.so, f.sl);
            } else if (layout_class == 2) {
                const int nd = (int)f.le(m.addr + 2, 1);
                chunk_btree = f.le(m.addr + 3, f.so);
                for (int i = 0; i < nd; i++) chunk_dims.push_back(f.le(m.addr + 3 + f.so + 4ull * i, 4));
            } else {
                f.fail("unknown layout class");
            }
            have_layout = true;
        } else if (m.type == 0x0b) {  // filter pipeline
            const int v = (int)f.le(m.addr, 1), nf = (int)f.le(m.addr + 1, 1);
            uint64_t p = m.addr + (v == 1 ? 8 : 2);
            for (int i = 0; i < nf; i++) {
                const int id = (int)f.le(p, 2);
                int name_len = 0;
                if (v == 1 || id >= 256) { name_len = (int)f.le(p + 2, 2); p += 2; }
                const int ncd = (int)f.le(p + 4, 2);
                p += 6;
                if (v == 1) name_len = (name_len + 7) & ~7;
                p += name_len + 4ull * ncd;
                if (v == 1 && (ncd & 1)) p += 4;
                filters.push_back(id);
            }
        }
    }
    if (!have_type || !have_layout || out.shape.empty()) f.fail("dataset \"" + dataset + "\" lacks a dataspace, datatype or layout");
    size_t n = 1;
    for (size_t d : out.shape) n *= d;
    std::vector<unsigned char> raw(n * (size_t)elem, 0);
    if (layout_class == 0) {
        if (data_size < raw.size()) f.fail("compact dataset smaller than its dataspace");
        std::memcpy(raw.data(), compact, raw.size());
    } else if (layout_class == 1) {
        if (!f.undefined(data_addr, f.so)) {  // undefined address: never written, all fill value (zero)
            if (data_size < raw.size()) f.fail("contiguous dataset smaller than its dataspace");
            std::memcpy(raw.data(), f.at(data_addr, raw.size()), raw.size());
        }
    } else {
        for (int id : filters)
            if (id != 1 && id != 2) f.fail("filter " + std::to_string(id) + " is not supported (only deflate and shuffle)");
        const size_t rank = out.shape.size();
        if (chunk_dims.size() != rank + 1) f.fail("chunk dimensionality does not match the dataspace");
        size_t chunk_elems = 1;
        for (size_t i = 0; i < rank; i++) chunk_elems *= (size_t)chunk_dims[i];
        const size_t chunk_bytes = chunk_elems * (size_t)elem;
        std::vector<uint64_t> todo;
        if (!f.undefined(chunk_btree, f.so)) todo.push_back(chunk_btree);
        while (!todo.empty()) {
            const uint64_t node = todo.back();
            todo.pop_back();
            if (std::memcmp(f.at(node, 4), "TREE", 4) != 0 || f.le(node + 4, 1) != 1) f.fail("bad chunk B-tree node");
            const int level = (int)f.le(node + 5, 1), used = (int)f.le(node + 6, 2);
            uint64_t p = node + 8 + 2 * (uint64_t)f.so;
            const uint64_t key_size = 8 + 8 * (uint64_t)(rank + 1);
            for (int i = 0; i < used; i++) {
                const uint64_t csize = f.le(p, 4), mask = f.le(p + 4, 4);
                std::vector<uint64_t> off(rank);
                for (size_t d = 0; d < rank; d++) off[d] = f.le(p + 8 + 8 * d, 8);
                const uint64_t child = f.le(p + key_size, f.so);
                p += key_size + f.so;
                if (level > 0) { todo.push_back(child); continue; }
                std::vector<unsigned char> chunk(f.at(child, (size_t)csize), f.at(child, (size_t)csize) + csize);
                // filters are undone in reverse order of the pipeline
                for (int k = (int)filters.size() - 1; k >= 0; k--) {
                    if (mask & (1u << k)) continue;
                    if (filters[k] == 1) {
                        chunk = inflate_all(f, chunk.data(), chunk.size(), chunk_bytes);
                    } else {  // shuffle: byte planes -> elements
                        std::vector<unsigned char> un(chunk.size());
                        const size_t ne = chunk.size() / (size_t)elem;
                        for (size_t e = 0; e < ne; e++)
                            for (int b = 0; b < elem; b++) un[e * elem + b] = chunk[(size_t)b * ne + e];
                        chunk.swap(un);
                    }
                }
                if (chunk.size() < chunk_bytes) f.fail("chunk smaller than its dimensions");
                // scatter the chunk's rows (innermost dimension) into the dataset, clipped at the dataset's edges
                std::vector<size_t> idx(rank, 0);
                const size_t inner = (size_t)chunk_dims[rank - 1];
                const size_t rows_in_chunk = chunk_elems / inner;
                for (size_t r = 0; r < rows_in_chunk; r++) {
                    size_t rem = r;
                    for (size_t d = rank - 1; d-- > 0;) { idx[d] = rem % (size_t)chunk_dims[d]; rem /= (size_t)chunk_dims[d]; }
                    bool inside = true;
                    size_t dst = 0;
                    for (size_t d = 0; d + 1 < rank; d++) {
                        const size_t g = (size_t)off[d] + idx[d];
                        if (g >= out.shape[d]) { inside = false; break; }
                        dst = dst * out.shape[d] + g;
                    }
                    if (!inside || off[rank - 1] >= out.shape[rank - 1]) continue;
                    const size_t take = std::min(inner, out.shape[rank - 1] - (size_t)off[rank - 1]);
                    dst = dst * out.shape[rank - 1] + (size_t)off[rank - 1];
                    std::memcpy(raw.data() + dst * elem, chunk.data() + r * inner * elem, take * (size_t)elem);
                }
            }
        }
    }
    out.elem = elem;
    out.is_signed = is_signed;
    out.big_endian = big_endian;
    out.bytes.swap(raw);
    return out;
}

// The dataset as native int, like H5::DataSet::read(..., H5::PredType::NATIVE_INT) (H5Segmentation.cpp:48).
inline H5Int32 load_h5_int32(const std::string& path, const std::string& dataset) {
    const H5Raw raw = load_h5_raw(path, dataset);
    if (raw.type_class != 0)
        throw std::invalid_argument(path + ": dataset \"" + dataset + "\" is not of an integer type");  // the reference asserts H5T_INTEGER
    H5Int32 out;
    out.shape = raw.shape;
    size_t n = 1;
    for (size_t d : out.shape) n *= d;
    const int elem = raw.elem;
    out.data.resize(n);
    for (size_t i = 0; i < n; i++) {
        const unsigned char* p = raw.bytes.data() + i * (size_t)elem;
        uint64_t v = 0;
        for (int b = 0; b < elem; b++) v |= (uint64_t)p[raw.big_endian ? elem - 1 - b : b] << (8 * b);
        int64_t s = (int64_t)v;
        if (raw.is_signed && elem < 8 && (v >> (8 * elem - 1))) s = (int64_t)(v | (~0ull << (8 * elem)));
        // HDF5's integer conversion clamps values that do not fit the destination
        if (s > INT32_MAX || (!raw.is_signed && elem == 8 && v > (uint64_t)INT32_MAX)) s = INT32_MAX;
        if (s < INT32_MIN) s = INT32_MIN;
        out.data[i] = (int32_t)s;
    }
    return out;
}

}  // namespace isx_apps

#endif  // ISX_APPS_H5_READER_H_
