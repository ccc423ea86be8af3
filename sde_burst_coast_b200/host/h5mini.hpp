// h5mini.hpp - a minimal, dependency-free HDF5 writer for what the reference stores in out_<fileID>.h5
// (/root/reference/src/h5tools.cpp:23-217, src/swarmdyn.cpp:437-462,546-625): a handful of N-dimensional
// float64 datasets at the root ("/part", "/pred", "/swarm", "/swarm_fishNet") or under one group
// "/<fileID>/...". This image has no libhdf5, so the file is produced byte by byte from
// the HDF5 File Format Specification, version 0 structures throughout (the ones every libhdf5 reads):
//   superblock v0 -> root symbol-table entry -> object header v1 (symbol table message)
//   -> B-tree v1 node ("TREE") + local heap ("HEAP") + one symbol-table node ("SNOD")
//   -> per dataset: object header v1 {dataspace v1, datatype (IEEE F64LE), fill value v2, layout v3 contiguous}
// Contiguous, uncompressed layout (the reference writes chunked gzip-9; readers do not care).
// Limits: at most 8 links per group (one symbol-table node, 2 * leaf K), names shorter than 256 bytes.
// Verified here only against the independent reader tests/h5mini_read.py (no HDF5 library is available);
// DESIGN.md section 9 says so.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace h5mini {

struct Dataset {
    std::string name;
    std::vector<uint64_t> dims;
    std::vector<double> data;  // row-major, product(dims) values
};

class Writer {
public:
    // group == "" : datasets at the root; otherwise under "/<group>/"
    explicit Writer(std::string group = "") : group_(std::move(group)) {}
    void add(const std::string &name, const std::vector<uint64_t> &dims, const std::vector<double> &data) {
        Dataset d;
        d.name = name;
        d.dims = dims;
        d.data = data;
        uint64_t n = 1;
        for (uint64_t v : dims) n *= v;
        d.data.resize(n, 0.0);
        sets_.push_back(std::move(d));
    }
    bool write(const std::string &path) {
        if (sets_.empty() || sets_.size() > 8) return false;
        std::sort(sets_.begin(), sets_.end(), [](const Dataset &a, const Dataset &b) { return a.name < b.name; });
        buf_.clear();
        buf_.resize(96, 0);  // superblock, filled last
        // data objects first, groups afterwards (addresses are all known by then)
        std::vector<uint64_t> ds_header(sets_.size());
        for (size_t k = 0; k < sets_.size(); k++) ds_header[k] = write_dataset(sets_[k]);
        std::vector<std::pair<std::string, uint64_t>> links;
        for (size_t k = 0; k < sets_.size(); k++) links.push_back({sets_[k].name, ds_header[k]});
        uint64_t btree = 0, heap = 0;
        uint64_t top_header = write_group(links, btree, heap);
        if (!group_.empty()) {
            std::vector<std::pair<std::string, uint64_t>> root_links = {{group_, top_header}};
            top_header = write_group(root_links, btree, heap);
        }
        write_superblock(top_header, btree, heap);
        FILE *f = std::fopen(path.c_str(), "wb");
        if (!f) return false;
        const bool ok = std::fwrite(buf_.data(), 1, buf_.size(), f) == buf_.size();
        std::fclose(f);
        return ok;
    }

private:
    static constexpr uint64_t UNDEF = ~0ull;
    static constexpr int LEAF_K = 4, INTERNAL_K = 16;
    std::string group_;
    std::vector<Dataset> sets_;
    std::vector<uint8_t> buf_;

    uint64_t here() const { return buf_.size(); }
    void align8() { while (buf_.size() % 8) buf_.push_back(0); }
    void u8(uint8_t v) { buf_.push_back(v); }
    void u16(uint16_t v) { for (int k = 0; k < 2; k++) buf_.push_back((uint8_t)(v >> (8 * k))); }
    void u32(uint32_t v) { for (int k = 0; k < 4; k++) buf_.push_back((uint8_t)(v >> (8 * k))); }
    void u64(uint64_t v) { for (int k = 0; k < 8; k++) buf_.push_back((uint8_t)(v >> (8 * k))); }
    void bytes(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; buf_.insert(buf_.end(), q, q + n); }
    void put64(size_t at, uint64_t v) { for (int k = 0; k < 8; k++) buf_[at + k] = (uint8_t)(v >> (8 * k)); }

    // one header message: type, data (padded to a multiple of 8)
    static void message(std::vector<uint8_t> &out, uint16_t type, const std::vector<uint8_t> &data, uint8_t flags = 0) {
        const size_t padded = (data.size() + 7) / 8 * 8;
        out.push_back((uint8_t)type); out.push_back((uint8_t)(type >> 8));
        out.push_back((uint8_t)padded); out.push_back((uint8_t)(padded >> 8));
        out.push_back(flags); out.push_back(0); out.push_back(0); out.push_back(0);
        out.insert(out.end(), data.begin(), data.end());
        out.insert(out.end(), padded - data.size(), 0);
    }
    // object header v1: 16-byte prefix (incl. 4 bytes of alignment padding) + messages
    uint64_t object_header(const std::vector<uint8_t> &msgs, uint16_t n_msgs) {
        align8();
        const uint64_t at = here();
        u8(1); u8(0); u16(n_msgs); u32(1); u32((uint32_t)msgs.size()); u32(0);
        bytes(msgs.data(), msgs.size());
        return at;
    }

    uint64_t write_dataset(const Dataset &d) {
        align8();
        const uint64_t data_at = here();
        bytes(d.data.data(), d.data.size() * sizeof(double));  // little-endian host assumed (x86-64)
        std::vector<uint8_t> m, t;
        // dataspace v1: version, rank, flags, reserved(1+4), dims
        t = {1, (uint8_t)d.dims.size(), 0, 0, 0, 0, 0, 0};
        for (uint64_t v : d.dims) for (int k = 0; k < 8; k++) t.push_back((uint8_t)(v >> (8 * k)));
        message(m, 0x0001, t);
        // datatype: IEEE 754 binary64, little-endian (class 1 version 1; sign bit 63; exponent 11 bits at 52,
        // mantissa 52 bits at 0, implied leading bit, bias 1023)
        t = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xff, 0x03, 0, 0};
        message(m, 0x0003, t, 1);
        // fill value v2: allocate early, write fill if set, no fill value defined
        t = {2, 1, 2, 0};
        message(m, 0x0005, t, 1);
        // data layout v3, contiguous: address, size
        t = {3, 1};
        for (int k = 0; k < 8; k++) t.push_back((uint8_t)(data_at >> (8 * k)));
        const uint64_t nbytes = d.data.size() * sizeof(double);
        for (int k = 0; k < 8; k++) t.push_back((uint8_t)(nbytes >> (8 * k)));
        message(m, 0x0008, t);
        return object_header(m, 4);
    }

    // old-style group: local heap with the names, one symbol-table node, a one-entry B-tree, object header
    uint64_t write_group(std::vector<std::pair<std::string, uint64_t>> links, uint64_t &btree_at, uint64_t &heap_at) {
        std::sort(links.begin(), links.end());
        // heap data segment: "" at offset 0, then the names, each padded to 8 bytes, then one free block
        std::vector<uint8_t> seg(8, 0);
        std::vector<uint64_t> name_off;
        for (auto &l : links) {
            name_off.push_back(seg.size());
            seg.insert(seg.end(), l.first.begin(), l.first.end());
            seg.push_back(0);
            while (seg.size() % 8) seg.push_back(0);
        }
        const uint64_t free_off = seg.size();
        const uint64_t free_size = 32;
        // free block: offset of the next free block (1 = end of list), size of this block
        for (int k = 0; k < 8; k++) seg.push_back(k == 0 ? 1 : 0);
        for (int k = 0; k < 8; k++) seg.push_back((uint8_t)(free_size >> (8 * k)));
        seg.resize(free_off + free_size, 0);
        align8();
        const uint64_t seg_at = here();
        bytes(seg.data(), seg.size());
        align8();
        heap_at = here();
        bytes("HEAP", 4); u8(0); u8(0); u8(0); u8(0);
        u64(seg.size()); u64(free_off); u64(seg_at);
        // symbol table node
        align8();
        const uint64_t snod_at = here();
        bytes("SNOD", 4); u8(1); u8(0); u16((uint16_t)links.size());
        for (size_t k = 0; k < (size_t)(2 * LEAF_K); k++) {
            if (k < links.size()) { u64(name_off[k]); u64(links[k].second); u32(0); u32(0); u64(0); u64(0); }
            else { for (int q = 0; q < 5; q++) u64(0); }
        }
        // B-tree v1 node, type 0 (group), leaf level, one child
        align8();
        btree_at = here();
        bytes("TREE", 4); u8(0); u8(0); u16(1); u64(UNDEF); u64(UNDEF);
        u64(0);                 // key 0: the empty string
        u64(snod_at);           // child 0
        u64(name_off.back());   // key 1: the largest name in child 0
        for (int k = 0; k < 2 * INTERNAL_K - 1; k++) { u64(0); u64(0); }  // unused (child, key) pairs
        // object header with the symbol table message
        std::vector<uint8_t> m, t;
        for (int k = 0; k < 8; k++) t.push_back((uint8_t)(btree_at >> (8 * k)));
        for (int k = 0; k < 8; k++) t.push_back((uint8_t)(heap_at >> (8 * k)));
        message(m, 0x0011, t);
        return object_header(m, 1);
    }

    void write_superblock(uint64_t root_header, uint64_t btree, uint64_t heap) {
        align8();
        std::vector<uint8_t> sb;
        const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        sb.insert(sb.end(), sig, sig + 8);
        const uint8_t ver[8] = {0, 0, 0, 0, 0, 8, 8, 0};  // versions, size of offsets, size of lengths
        sb.insert(sb.end(), ver, ver + 8);
        sb.push_back(LEAF_K); sb.push_back(0); sb.push_back(INTERNAL_K); sb.push_back(0);
        for (int k = 0; k < 4; k++) sb.push_back(0);  // consistency flags
        auto push64 = [&](uint64_t v) { for (int k = 0; k < 8; k++) sb.push_back((uint8_t)(v >> (8 * k))); };
        push64(0);              // base address
        push64(UNDEF);          // free-space info
        push64(buf_.size());    // end of file
        push64(UNDEF);          // driver info
        push64(0);              // root entry: link name offset
        push64(root_header);    // object header address
        for (int k = 0; k < 4; k++) sb.push_back(k == 0 ? 1 : 0);  // cache type 1: group
        for (int k = 0; k < 4; k++) sb.push_back(0);
        push64(btree);
        push64(heap);
        std::memcpy(buf_.data(), sb.data(), sb.size());
    }
};

}  // namespace h5mini
