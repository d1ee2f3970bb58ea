// TEST INFRASTRUCTURE — not part of the product path.
//
// Thin extern "C" wrapper that instantiates the UNMODIFIED reference decoders
// (headers are included from $(REF) = /root/reference where they lie; nothing
// is copied) so that
//   * tests can pin the C restatement in oracle/dint_oracle.c against the real
//     reference on arbitrary inputs, and
//   * bench.py --impl reference / cpu_baseline(kind="reference") can time the
//     reference's own CPU decode on the GPU box's host cores.
// Built by oracle/Makefile into oracle/_ref/libdint_ref.so (git-ignored).
//
// Reference entry points wrapped:
//   vroom_env/dint_codecs.hpp:315-323  single_opt_dint::decode
//   vroom_env/dint_codecs.hpp:521-619  multi_opt_dint::decode
//   include/dint/dint_codecs.hpp:269-274 opt_dint_single_dict_block::decode
//   include/dint/dint_codecs.hpp:460-510 opt_dint_multi_dict_block::decode
//   include/ds2i/block_codecs.hpp:87-101  TightVariableByte::decode
//   include/ds2i/block_codecs.hpp:130-150 interpolative_block::decode
#include <cassert>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <unordered_map>
#include <vector>

#include <succinct/mapper.hpp>
#include "configuration.hpp"
#include "semiasync_queue.hpp"
#include "block_codecs.hpp"
#include "dint_configuration.hpp"
#include "dictionary_types.hpp"
#include "dint_codecs.hpp"              // include/dint (block-wise codecs)
#include REF_VROOM_DINT_CODECS          // vroom_env/dint_codecs.hpp (whole-list codecs)

using namespace ds2i;

namespace {
enum { KIND_RECT = 0, KIND_PACKED = 1, KIND_MULTI = 2 };

struct ref_dict {
    int kind;
    single_dictionary_rectangular_type rect;
    single_dictionary_packed_type packed;
    multi_dictionary_packed_type multi;
};

template <typename D>
bool load_into(D& d, const char* path) {
    typename D::builder b;
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    b.load(f);
    b.build(d);
    return true;
}
}  // namespace

extern "C" {

int ref_block_size() { return int(constants::block_size); }

void* ref_dict_load(int kind, const char* path) {
    auto* d = new ref_dict;
    d->kind = kind;
    bool ok = false;
    if (kind == KIND_RECT) ok = load_into(d->rect, path);
    if (kind == KIND_PACKED) ok = load_into(d->packed, path);
    if (kind == KIND_MULTI) ok = load_into(d->multi, path);
    if (!ok) { delete d; return nullptr; }
    return d;
}

void ref_dict_free(void* h) { delete static_cast<ref_dict*>(h); }

// whole-list ("vroom") decode of ONE list; out must be zeroed by the caller and
// have n + 256 slots. Returns bytes consumed.
size_t ref_vroom_decode(void* h, const uint8_t* in, uint32_t* out, uint32_t n) {
    auto* d = static_cast<ref_dict*>(h);
    const uint8_t* end = nullptr;
    if (d->kind == KIND_RECT) end = single_opt_dint::decode(d->rect, in, out, 0, n);
    else if (d->kind == KIND_PACKED) end = single_opt_dint::decode(d->packed, in, out, 0, n);
    else end = multi_opt_dint::decode(d->multi, in, out, 0, n);
    return size_t(end - in);
}

// block-wise decode of ONE block (n <= block_size); interpolative when n < block_size.
size_t ref_block_decode(void* h, const uint8_t* in, uint32_t* out, uint32_t sum_of_values, uint32_t n) {
    auto* d = static_cast<ref_dict*>(h);
    const uint8_t* end = nullptr;
    if (d->kind == KIND_RECT) end = opt_dint_single_dict_block::decode(d->rect, in, out, sum_of_values, n);
    else if (d->kind == KIND_PACKED) end = opt_dint_single_dict_block::decode(d->packed, in, out, sum_of_values, n);
    else end = opt_dint_multi_dict_block::decode(d->multi, in, out, sum_of_values, n);
    return size_t(end - in);
}

size_t ref_vbyte_decode(const uint8_t* in, uint32_t* out, size_t n) {
    return size_t(TightVariableByte::decode(in, out, n) - in);
}

// Whole vroom file, reference check_encoded_data semantics (zeroed slack after
// every list, vroom_env/check_encoded_data.cpp:76-110): writes the n decoded
// values of each list back to back into `out` (capacity out_cap ints).
// Returns number of lists, or -1 on overflow.
long ref_vroom_decode_file(void* h, const uint8_t* bin, size_t len, uint32_t* out, size_t out_cap,
                           uint64_t* total_ints) {
    std::vector<uint32_t> buf;
    const uint8_t* p = bin;
    const uint8_t* end = bin + len;
    size_t pos = 0;
    long lists = 0;
    while (p != end) {
        uint32_t hdr[2];
        p = TightVariableByte::decode(p, hdr, 2);  // == vroom_env/codecs.hpp:117-123 header::read
        uint32_t n = hdr[0];
        if (pos + n > out_cap) return -1;
        buf.assign(size_t(n) + 512, 0);
        p += ref_vroom_decode(h, p, buf.data(), n);
        std::memcpy(out + pos, buf.data(), size_t(n) * 4);
        pos += n;
        ++lists;
    }
    if (total_ints) *total_ints = pos;
    return lists;
}

// Timing harness with the semantics of vroom_env/decode.cpp:139-150: one reused
// output buffer that is never re-zeroed, per-list chrono deltas summed. Decodes
// the lists whose headers start in [bin, bin+len) `repeat` times. Returns the
// summed seconds; *ints receives the number of integers decoded in one pass.
//
// scratch: optional caller-owned output buffer of >= constants::max_size (50 M) zeroed ints,
// so that repeated calls do not pay the 200 MB allocation the reference pays once per process
// (vroom_env/decode.cpp:128-129).
double ref_vroom_time_file(void* h, const uint8_t* bin, size_t len, int repeat, uint64_t* ints, uint32_t* scratch) {
    std::vector<uint32_t> own;
    if (!scratch) {
        own.assign(constants::max_size, 0);
        scratch = own.data();
    }
    struct { uint32_t* p; uint32_t* data() { return p; } } decoded{scratch};
    double tot = 0;
    uint64_t cnt = 0;
    for (int r = 0; r < repeat; ++r) {
        const uint8_t* p = bin;
        const uint8_t* end = bin + len;
        cnt = 0;
        while (p != end) {
            uint32_t hdr[2];
            p = TightVariableByte::decode(p, hdr, 2);
            auto t0 = clock_type::now();
            p += ref_vroom_decode(h, p, decoded.data(), hdr[0]);
            auto t1 = clock_type::now();
            tot += std::chrono::duration<double>(t1 - t0).count();
            cnt += hdr[0];
        }
    }
    if (ints) *ints = cnt;
    return tot;
}

}  // extern "C"
