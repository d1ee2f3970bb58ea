// Internal declarations shared by the host format layer (dint_host.cpp) and the CUDA
// translation unit (dint_kernels.cu). Not installed; the public surface is include/dint_b200.h.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dint_b200.h"

namespace dint {

constexpr uint32_t kEntry = 16;        // constants::max_entry_size (dint_configuration.hpp:23)
constexpr uint32_t kNumEntries = 65536;  // constants::num_entries   (dint_configuration.hpp:25)
constexpr uint32_t kSelectors = 6;     // constants::num_selectors  (dint_configuration.hpp:20)
constexpr uint32_t kReserved = 7;      // EXCEPTIONS + 5 run codewords (single_dictionary.hpp:22)
constexpr uint32_t kDefaultBlock = 256;  // constants::block_size    (include/util.hpp:35)
constexpr uint32_t kDefaultChunk = 1024;
constexpr uint32_t kUnitInts = 2048;     // a work unit closes once it holds this many integers ...
constexpr uint32_t kUnitSegments = 32;   // ... or this many segments (one per lane of the decoding warp)


// strip tables (lane-sequential decoder): tunables, see dint_strip_limits()
struct strip_limits {
    uint32_t unit_ints;        // integers per strip unit (shared-memory output image per warp)
    uint32_t unit_bytes;       // compressed bytes spanned by a strip unit (shared-memory input buffer per warp)
    uint32_t strip_codewords;  // codewords per strip (steps of the per-thread loop)
    uint32_t strip_ints;       // a strip also closes once it holds this many integers (runs of zeros)
};
const strip_limits& get_strip_limits();

void set_error(const std::string& msg);

inline uint16_t rd16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }

}  // namespace dint

// Device-side view of a dictionary, passed by value to kernels.
struct dint_dict_view {
    const uint32_t* rows;   // [n_rows] x 16 words, 64-byte aligned rows: entry payload, zero padded
    const uint8_t* lens;    // 4-bit log2(entry length) per row, two rows per byte
    uint32_t class_base[8]; // first row of each class (single kinds: one class at 0)
    uint32_t n_rows;
    uint32_t n_classes;
    // lane-sequential decoder: the same 4-bit table, one 16-byte aligned slice per class indexed by the
    // symbol itself; the first lut_rows rows of every class are staged in shared memory
    const uint8_t* lens_cls;
    uint32_t lens_cls_off[8];
    uint32_t class_rows[8];
    uint32_t lut_rows;
    // lane-sequential decoder, tables every CTA stages in shared memory (one contiguous image):
    //   hot0   [n_classes][hot_rows] uint4   words 0..3 of the first hot_rows rows of every class
    //   hot1   [n_classes][hot_rows] uint4   words 4..7
    //   hotlen [n_classes][hot_rows] u8      log2(entry length) of those rows
    //   group  per class, one byte per 32 rows: log2(length) when the 32 rows agree, else 0xFF
    //          (the rows of a DSF dictionary are almost sorted by length beyond the first few hundred)
    const uint8_t* strip_tables;
    uint32_t strip_tables_bytes;  // multiple of 16
    uint32_t hot_rows;            // power of two >= 256 (every u8 symbol is hot)
    uint32_t group_off[8];        // byte offset of the class's group slice inside the image
    uint32_t detail_off[8];       // byte offset of the class's mixed-group detail: group byte 9 + k selects the 16
                                  // bytes of per-row nibbles at detail_off + 16 k; 0xFF = look in lens_cls (global)
};

struct dint_dict {
    int kind = 0;
    int device = -1;
    uint32_t n_rows = 0;
    uint32_t n_classes = 1;
    uint32_t class_base[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t class_size[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> h_rows;  // n_rows * 16
    std::vector<uint8_t> h_lens;   // len - 1
    std::vector<uint32_t> h_sums;  // sum(v_j + 1) over the entry, mod 2^32 (for doc_base of chunks)
    uint32_t* d_rows = nullptr;
    uint8_t* d_lens = nullptr;
    uint8_t* d_lens_cls = nullptr;
    uint32_t* d_sums = nullptr;    // h_sums in HBM (device-side discovery: docID prefix of every strip)
    uint32_t lens_cls_off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t lut_rows = 0;
    uint8_t* d_strip_tables = nullptr;
    uint32_t strip_tables_bytes = 0;
    uint32_t hot_rows = 0;
    uint32_t group_off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t detail_off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // launch shape of the strip decoder for this dictionary (set once by dint_dict_upload)
    int strip_warps = 0, strip_sms = 0, strip_variant = -1;
    size_t strip_smem = 0;
    // device scratch of dint_decode_list_host (grown on demand, reused by every call)
    struct host_scratch {
        std::mutex mu;
        uint8_t* d_in = nullptr;
        size_t in_cap = 0;
        uint32_t* d_out = nullptr;
        size_t out_cap = 0;
        uint64_t* d_cons = nullptr;
        bool attr_set = false;
    };
    mutable host_scratch scratch;
    dint_dict_view view() const {
        dint_dict_view v;
        v.rows = d_rows;
        v.lens = d_lens;
        v.lens_cls = d_lens_cls;
        v.lut_rows = lut_rows;
        v.strip_tables = d_strip_tables;
        v.strip_tables_bytes = strip_tables_bytes;
        v.hot_rows = hot_rows;
        for (int i = 0; i < 8; ++i) v.group_off[i] = group_off[i];
        for (int i = 0; i < 8; ++i) v.detail_off[i] = detail_off[i];
        for (int i = 0; i < 8; ++i) {
            v.lens_cls_off[i] = lens_cls_off[i];
            v.class_rows[i] = class_size[i];
        }
        for (int i = 0; i < 8; ++i) v.class_base[i] = class_base[i];
        v.n_rows = n_rows;
        v.n_classes = n_classes;
        return v;
    }
};

struct dint_vroom_plan {
    uint64_t n_ints = 0;
    std::vector<dint_segment> segments;
    std::vector<uint32_t> unit_first;         // n_units + 1
    std::vector<dint_strip_unit> strip_units; // lane-sequential decoder's work table
    uint64_t cw_stats[16] = {0};              // see dint_vroom_plan_codeword_stats
    std::vector<uint64_t> list_offsets;       // n_lists + 1
    std::vector<uint64_t> list_byte_offsets;  // n_lists + 1
};

struct dint_index_plan {
    int kind = 0;
    uint64_t n_lists = 0, num_docs = 0, n_postings = 0;
    uint64_t lists_off = 0, lists_len = 0, docs_dict_off = 0, freqs_dict_off = 0, end_off = 0;
    std::vector<dint_block> blocks;
    std::vector<uint64_t> list_offsets;  // n_lists + 1 (postings)
    std::vector<uint64_t> endpoints;     // n_lists + 1 (bytes in m_lists)
};

// Strip table + list offsets found on the device (dint_discover.cu); everything lives in HBM.
struct dint_device_plan {
    int device = -1;
    uint64_t n_lists = 0, n_ints = 0, n_units = 0, n_chains = 0, in_bytes = 0;
    dint_strip_unit* d_units = nullptr;
    uint64_t* d_list_offsets = nullptr;       // n_lists + 1 (vroom plans only)
    uint64_t* d_list_byte_offsets = nullptr;  // n_lists + 1 (vroom plans only)
    float discover_ms = 0.f;                  // device timeline of the discovery (CUDA events)
    float phase_ms[3] = {0.f, 0.f, 0.f};      // sync search, count pass(es), emit pass
    uint32_t retries = 0;                     // count passes repeated because a speculative sync point was wrong
    uint64_t n_sync_found = 0;                // sync points the speculative search contributed
    uint64_t n_skimmed = 0;                   // long chains cut into thread-sized pieces by the skim pass
};

// Fixed-size header of a frozen dict_freq_index .bin: everything the host reads of the file when the block table is
// built on the device (dint_index_device_plan_create)
struct dint_index_header {
    uint64_t n_lists = 0, num_docs = 0;
    uint64_t ef_words_off = 0, ef_bits = 0, ef_words = 0;  // byte offset of the bit vector's words, its bits / words
    uint64_t l = 0, hb_off = 0, hb_len = 0, lb_off = 0;    // Elias-Fano geometry (bit offsets inside the vector)
    uint64_t lists_off = 0, lists_len = 0, docs_dict_off = 0, freqs_dict_off = 0, end_off = 0;
};
namespace dint {
int index_header_parse(int kind, const uint8_t* bin, size_t len, dint_index_header* h);  // dint_host.cpp
}

// Block table + list offsets + endpoints of a frozen index, built on the device (dint_discover.cu)
struct dint_device_index {
    int device = -1, kind = 0;
    dint_index_header h;
    uint64_t n_postings = 0, n_blocks = 0;
    uint64_t* d_endpoints = nullptr;     // n_lists + 1 byte offsets in m_lists
    uint64_t* d_list_offsets = nullptr;  // n_lists + 1 postings
    dint_block* d_blocks = nullptr;      // n_blocks
    float build_ms = 0.f;                // device timeline (CUDA events)
};

// implemented in dint_kernels.cu
int dint_dict_upload(dint_dict* d);
void dint_dict_release_device(dint_dict* d);
