// Host format layer of dint_b200: byte-for-byte consumers of the reference's on-disk
// formats, and the one-time pre-passes that turn a .bin into flat descriptor tables
// for the device decoder. No decoding of integers happens here.
//
// Reference formats followed (paths relative to the reference root):
//   dictionary files      rectangular_dictionary.hpp:72-92, single_dictionary.hpp:72-107,
//                         multi_dictionary.hpp:70-121
//   vroom .bin            vroom_env/encode.cpp:177-190, vroom_env/jobs.hpp:89-95,
//                         vroom_env/codecs.hpp:110-124 (header), :93-107 (vbyte)
//   frozen index .bin     dict_freq_index.hpp:209-214, external/succinct/mapper.hpp:100-156,
//                         external/succinct/bit_vector.hpp:228-231,
//                         include/ds2i/compact_elias_fano.hpp:14-61 + :69-136 (bit layout),
//                         include/dint/dict_posting_list.hpp:10-56 (per-list layout)
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <new>

#include "dint_internal.hpp"

namespace dint {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

static uint32_t env_u32(const char* name, uint32_t dflt, uint32_t lo, uint32_t hi) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    unsigned long x = std::strtoul(v, nullptr, 10);
    return uint32_t(std::min<unsigned long>(std::max<unsigned long>(x, lo), hi));
}

const strip_limits& get_strip_limits() {
    static const strip_limits lim = [] {
        strip_limits l;
        l.unit_ints = env_u32("DINT_B200_UNIT_INTS", 1024, 512, DINT_UNIT_MAX_INTS) & ~3u;
        l.unit_bytes = env_u32("DINT_B200_UNIT_BYTES", 1792, 512, DINT_UNIT_MAX_BYTES - 15);
        l.strip_codewords = env_u32("DINT_B200_STRIP_CODEWORDS", 8, 1, 64);
        l.strip_ints = std::min(512u, l.unit_ints - 256u);
        return l;
    }();
    return lim;
}

static inline uint32_t msb64(uint64_t x) { return 63u - uint32_t(__builtin_clzll(x)); }

// TightVariableByte: 7-bit groups LSB first, terminator has bit 7 SET (block_codecs.hpp:87-101)
static inline const uint8_t* vbyte(const uint8_t* p, const uint8_t* end, uint32_t* v) {
    uint32_t x = 0;
    for (unsigned shift = 0; p < end; shift += 7) {
        uint8_t c = *p++;
        x += uint32_t(c & 127u) << (shift & 31u);
        if (c & 128u) {
            *v = x;
            return p;
        }
    }
    return nullptr;
}

// ---------------------------------------------------------------------------------------
// dictionaries -> unified "expanded rows" representation
//   row r = 16 payload words (zero padded), lens[r] = length - 1, sums[r] = sum(v + 1)
// For packed/multi the (len-1)<<24 | offset indirection is resolved here, once, so the
// device never chases m_offsets -> m_table (two dependent loads per codeword on the CPU).
// ---------------------------------------------------------------------------------------
static void finish_rows(dint_dict* d) {
    d->h_sums.assign(d->n_rows, 0);
    for (uint32_t r = 0; r < d->n_rows; ++r) {
        uint32_t len = uint32_t(d->h_lens[r]) + 1;
        uint32_t s = len;  // runs: all zeros -> sum(v+1) = len
        for (uint32_t j = 0; j < std::min(len, kEntry); ++j) s += d->h_rows[size_t(r) * kEntry + j];
        d->h_sums[r] = s;
    }
}

static int build_rect(dint_dict* d, const uint8_t* table_words, size_t n_table_words) {
    // rows of 17 words: 16 payload + length (rectangular_dictionary.hpp:206-213)
    uint32_t m = uint32_t(n_table_words / (kEntry + 1));
    if (m > kNumEntries) return DINT_E_FORMAT;
    d->n_rows = kNumEntries;
    d->n_classes = 1;
    d->class_size[0] = kNumEntries;
    d->h_rows.assign(size_t(kNumEntries) * kEntry, 0);
    d->h_lens.assign(kNumEntries, 0);
    // builder::init() (rectangular_dictionary.hpp:43-56): reserved rows exist even if the file is short
    for (uint32_t i = 0, size = 256; i < 5; ++i, size /= 2) d->h_lens[2 + i] = uint8_t(size - 1);
    for (uint32_t r = 0; r < m; ++r) {
        const uint8_t* row = table_words + size_t(r) * (kEntry + 1) * 4;
        std::memcpy(&d->h_rows[size_t(r) * kEntry], row, kEntry * 4);
        uint32_t len = rd32(row + kEntry * 4);
        if (len > 256) {
            set_error("rectangular dictionary: entry length > 256");
            return DINT_E_FORMAT;
        }
        d->h_lens[r] = uint8_t(len ? len - 1 : 0);
    }
    finish_rows(d);
    return DINT_OK;
}

static int build_packed(dint_dict* d, const uint8_t* starts, uint32_t n_starts, const uint8_t* offsets,
                        uint32_t n_offsets, const uint8_t* table, uint32_t n_table) {
    d->n_rows = n_offsets;
    if (n_starts) {
        if (n_starts > 8) return DINT_E_FORMAT;
        d->n_classes = n_starts;
        for (uint32_t c = 0; c < n_starts; ++c) {
            d->class_base[c] = rd32(starts + 4 * c);
            if (d->class_base[c] > n_offsets || (c && d->class_base[c] < d->class_base[c - 1])) {
                set_error("multi dictionary: class start offsets are not monotone");
                return DINT_E_FORMAT;
            }
        }
        for (uint32_t c = 0; c < n_starts; ++c) {
            d->class_size[c] = (c + 1 < n_starts ? d->class_base[c + 1] : n_offsets) - d->class_base[c];
            // multi_dictionary.hpp:159-167: a class holds at most num_entries entries + its reserved slots
            if (d->class_size[c] > kNumEntries + kReserved) {
                set_error("multi dictionary: class larger than 65536 + 7 entries");
                return DINT_E_FORMAT;
            }
        }
    } else {
        d->n_classes = 1;
        d->class_size[0] = n_offsets;
    }
    d->h_rows.assign(size_t(n_offsets) * kEntry, 0);
    d->h_lens.assign(n_offsets, 0);
    for (uint32_t r = 0; r < n_offsets; ++r) {
        uint32_t so = rd32(offsets + 4 * size_t(r));
        uint32_t off = so & 0xFFFFFFu, len = (so >> 24) + 1;
        d->h_lens[r] = uint8_t(len - 1);
        uint32_t take = std::min(len, kEntry);
        if (uint64_t(off) + take > n_table) {
            // the reference memcpy's 16 words regardless (table padded by 16, multi_dictionary.hpp:108);
            // only the first `len` matter, and those must exist
            set_error("packed dictionary: entry runs past the table");
            return DINT_E_FORMAT;
        }
        std::memcpy(&d->h_rows[size_t(r) * kEntry], table + 4 * size_t(off), take * 4);
    }
    finish_rows(d);
    return DINT_OK;
}

static int parse_file(dint_dict* d, int kind, const uint8_t* f, size_t len) {
    if (kind == DINT_DICT_RECT) {
        if (len < 4) return DINT_E_FORMAT;
        uint32_t m = rd32(f);
        size_t words = size_t(m) * (kEntry + 1);
        if (m > kNumEntries || len < 4 + words * 4) return DINT_E_FORMAT;
        return build_rect(d, f + 4, words);
    }
    if (kind == DINT_DICT_PACKED) {
        if (len < 12) return DINT_E_FORMAT;
        uint32_t no = rd32(f + 4), nt = rd32(f + 8);
        if (len < 12 + (size_t(no) + nt) * 4) return DINT_E_FORMAT;
        return build_packed(d, nullptr, 0, f + 12, no, f + 12 + size_t(no) * 4, nt);
    }
    if (kind == DINT_DICT_MULTI) {
        if (len < 16) return DINT_E_FORMAT;
        uint32_t ns = rd32(f + 4), no = rd32(f + 8), nt = rd32(f + 12);
        if (len < 16 + (size_t(ns) + no + nt) * 4) return DINT_E_FORMAT;
        return build_packed(d, f + 16, ns, f + 16 + size_t(ns) * 4, no, f + 16 + (size_t(ns) + no) * 4, nt);
    }
    return DINT_E_ARG;
}

// mappable_vector<uint32_t> = u64 size + payload (mapper.hpp:128-146)
static int parse_mapped(dint_dict* d, int kind, const uint8_t* p, size_t len, size_t* consumed) {
    int nvec = kind == DINT_DICT_RECT ? 1 : kind == DINT_DICT_PACKED ? 2 : kind == DINT_DICT_MULTI ? 3 : 0;
    if (!nvec) return DINT_E_ARG;
    const uint8_t* vec[3] = {nullptr, nullptr, nullptr};
    uint64_t cnt[3] = {0, 0, 0};
    size_t off = 0;
    for (int v = 0; v < nvec; ++v) {
        if (off + 8 > len) return DINT_E_FORMAT;
        cnt[v] = rd64(p + off);
        off += 8;
        if (cnt[v] > (len - off) / 4) return DINT_E_FORMAT;
        vec[v] = p + off;
        off += size_t(cnt[v]) * 4;
    }
    if (consumed) *consumed = off;
    if (kind == DINT_DICT_RECT) return build_rect(d, vec[0], size_t(cnt[0]));
    if (kind == DINT_DICT_PACKED) return build_packed(d, nullptr, 0, vec[0], uint32_t(cnt[0]), vec[1], uint32_t(cnt[1]));
    return build_packed(d, vec[0], uint32_t(cnt[0]), vec[1], uint32_t(cnt[1]), vec[2], uint32_t(cnt[2]));
}

static int validate(dint_dict* d) {
    for (uint32_t c = 0; c < d->n_classes; ++c) {
        if (d->class_size[c] < kReserved) {
            set_error("dictionary class smaller than the 7 reserved codewords");
            return DINT_E_FORMAT;
        }
        uint32_t b = d->class_base[c];
        // codeword space (single_dictionary.hpp:39-56): 0,1 exceptions; 2..6 runs of 256..16 zeros
        for (uint32_t i = 0, size = 256; i < 5; ++i, size /= 2) {
            if (uint32_t(d->h_lens[b + 2 + i]) + 1 != size) {
                set_error("dictionary: reserved run codewords have unexpected lengths");
                return DINT_E_FORMAT;
            }
        }
        if (d->h_lens[b] != 0 || d->h_lens[b + 1] != 0) {
            set_error("dictionary: the exception codewords 0 and 1 must have length 1");
            return DINT_E_FORMAT;
        }
        // every other entry is one of constants::target_sizes (dint_configuration.hpp:25): the device tables
        // store log2(length), so anything else would decode differently from the reference
        for (uint32_t i = kReserved; i < d->class_size[c]; ++i) {
            const uint32_t len = uint32_t(d->h_lens[b + i]) + 1;
            if (len > kEntry || (len & (len - 1)) != 0) {
                set_error("dictionary: entry length is not one of 1, 2, 4, 8, 16");
                return DINT_E_FORMAT;
            }
        }
    }
    return DINT_OK;
}

// extent of `nvec` mappable_vector<uint32_t> images (u64 count + payload each, mapper.hpp:128-146): sizes only
static bool mapped_extent(int kind, const uint8_t* p, size_t len, size_t* used) {
    const int nvec = kind == DINT_DICT_RECT ? 1 : kind == DINT_DICT_PACKED ? 2 : 3;
    size_t off = 0;
    for (int v = 0; v < nvec; ++v) {
        if (off + 8 > len) return false;
        const uint64_t cnt = rd64(p + off);
        off += 8;
        if (cnt > (len - off) / 4) return false;
        off += size_t(cnt) * 4;
    }
    *used = off;
    return true;
}

// The header walk of dint_index_plan_create without touching anything that grows with the collection:
// u64 freeze flags | global_parameters 5 x u8 | u64 m_size | u64 m_num_docs | bit_vector { u64 bits, u64 words, words }
// | u64 m_lists.size | m_lists | docs dictionary | freqs dictionary (dict_freq_index.hpp:209-214, mapper.hpp:100-156)
int index_header_parse(int kind, const uint8_t* bin, size_t len, dint_index_header* h) {
    auto bad = [&](const char* m) {
        set_error(std::string("index .bin: ") + m);
        return DINT_E_FORMAT;
    };
    if (kind < 0 || kind > DINT_INDEX_BLOCK_U32) return DINT_E_ARG;
    size_t off = 8;
    if (len < off + 5 + 32) return bad("truncated header");
    const uint32_t log_s0 = bin[off], log_s1 = bin[off + 1];
    off += 5;
    h->n_lists = rd64(bin + off);
    h->num_docs = rd64(bin + off + 8);
    h->ef_bits = rd64(bin + off + 16);
    h->ef_words = rd64(bin + off + 24);
    off += 32;
    if (h->ef_words > (len - off) / 8 || h->ef_bits > h->ef_words * 64) return bad("endpoints bit vector out of range");
    h->ef_words_off = off;
    off += size_t(h->ef_words) * 8;
    if (off + 8 > len) return bad("truncated m_lists");
    h->lists_len = rd64(bin + off);
    off += 8;
    h->lists_off = off;
    if (h->lists_len > len - off) return bad("m_lists out of range");
    off += size_t(h->lists_len);
    h->docs_dict_off = off;
    if (kind < DINT_INDEX_BLOCK_INTERPOLATIVE) {  // (a block_freq_index has nothing behind m_lists)
        size_t used = 0;
        if (!mapped_extent(kind, bin + off, len - off, &used)) return bad("docs dictionary");
        off += used;
        h->freqs_dict_off = off;
        if (!mapped_extent(kind, bin + off, len - off, &used)) return bad("freqs dictionary");
        off += used;
    } else {
        h->freqs_dict_off = off;
    }
    h->end_off = off;
    if (off != len) return bad("trailing bytes");
    const uint64_t n = h->n_lists, universe = h->lists_len;
    if (n > universe || n > h->ef_bits) return bad("list count out of range");
    if (n) {  // compact_elias_fano.hpp:14-61
        h->l = universe > n ? msb64(universe / n) : 0;
        h->hb_len = n + (universe >> h->l) + 2;
        const uint64_t ptr_size = h->hb_len > 1 ? msb64(h->hb_len - 1) + 1 : 0;
        const uint64_t p0 = (h->hb_len - n) >> log_s0, p1 = n >> log_s1;
        h->hb_off = (p0 + p1) * ptr_size;
        h->lb_off = h->hb_off + h->hb_len;
        if (h->lb_off + n * h->l != h->ef_bits) return bad("Elias-Fano geometry mismatch");
    }
    return DINT_OK;
}

}  // namespace dint

using namespace dint;

extern "C" {

const char* dint_last_error(void) { return g_error.c_str(); }
int dint_abi_version(void) { return DINT_B200_ABI_VERSION; }
void dint_strip_limits(uint32_t* unit_max_ints, uint32_t* unit_max_bytes, uint32_t* strip_codewords) {
    const dint::strip_limits& l = dint::get_strip_limits();
    if (unit_max_ints) *unit_max_ints = l.unit_ints;
    if (unit_max_bytes) *unit_max_bytes = l.unit_bytes;
    if (strip_codewords) *strip_codewords = l.strip_codewords;
}

static int dict_finish(dint_dict* d, int rc, int device, dint_dict** out) {
    if (rc == DINT_OK) rc = validate(d);
    if (rc == DINT_OK && device >= 0) {
        d->device = device;
        rc = dint_dict_upload(d);
    }
    if (rc != DINT_OK) {
        if (rc == DINT_E_FORMAT && g_error.empty()) set_error("malformed dictionary");
        dint_dict_release_device(d);  // a failed upload may have left some of the tables allocated
        delete d;
        return rc;
    }
    *out = d;
    return DINT_OK;
}

int dint_dict_create(int kind, const void* file_bytes, size_t len, int device, dint_dict** out) {
    if (!file_bytes || !out) return DINT_E_ARG;
    g_error.clear();
    dint_dict* d = new (std::nothrow) dint_dict;
    if (!d) return DINT_E_NOMEM;
    d->kind = kind;
    try {
        return dict_finish(d, parse_file(d, kind, static_cast<const uint8_t*>(file_bytes), len), device, out);
    } catch (const std::exception&) {  // bad_alloc / length_error from a corrupt size field: never through the C boundary
        dint_dict_release_device(d);
        delete d;
        set_error("dictionary: out of memory");
        return DINT_E_NOMEM;
    }
}

int dint_dict_create_mapped(int kind, const void* bytes, size_t len, int device, dint_dict** out, size_t* consumed) {
    if (!bytes || !out) return DINT_E_ARG;
    g_error.clear();
    dint_dict* d = new (std::nothrow) dint_dict;
    if (!d) return DINT_E_NOMEM;
    d->kind = kind;
    try {
        return dict_finish(d, parse_mapped(d, kind, static_cast<const uint8_t*>(bytes), len, consumed), device, out);
    } catch (const std::exception&) {
        dint_dict_release_device(d);
        delete d;
        set_error("dictionary: out of memory");
        return DINT_E_NOMEM;
    }
}

void dint_dict_destroy(dint_dict* d) {
    if (!d) return;
    dint_dict_release_device(d);
    delete d;
}

int dint_dict_kind(const dint_dict* d) { return d ? d->kind : DINT_E_ARG; }
uint32_t dint_dict_num_rows(const dint_dict* d) { return d ? d->n_rows : 0; }
size_t dint_dict_device_bytes(const dint_dict* d) {
    return (d && d->d_rows) ? size_t(d->n_rows) * (kEntry * 4 + 1) : 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// vroom pre-pass
// ---------------------------------------------------------------------------------------
namespace {

struct walker {
    const dint_dict* d;
    const uint8_t* bin;
    const uint8_t* end;
    // Walks `count` integers of SB-byte codewords (2: u16, 1: u8) of class `cls` starting at p; returns the end
    // pointer or nullptr. acc accumulates sum(v + 1). cut(q, i, acc, symbol, length) is called after every
    // codeword. The body is written without data-dependent branches: in a d-gap stream every other codeword is an
    // exception, which a branch predictor cannot learn.
    template <unsigned SB, typename Strip, typename Each>
    const uint8_t* walk(const uint8_t* p, uint32_t cls, uint64_t count, uint32_t& acc, uint32_t strip_codewords,
                        uint64_t strip_ints, uint64_t* __restrict hist, Strip&& strip, Each&& each) const {
        const uint32_t base = d->class_base[cls], csize = d->class_size[cls];
        const uint8_t* lens = d->h_lens.data() + base;
        const uint32_t* sums = d->h_sums.data() + base;
        const uint8_t* safe = end - std::min<size_t>(size_t(end - bin), SB + 4);  // SB + 4 readable bytes from here back
        uint64_t i = 0, s_i = 0;         // integers walked; integers at the start of the open strip
        uint32_t a = acc, cw = 0, has_run = 0;
        while (i < count) {
            uint32_t w, pay;
            if (p <= safe) {
                w = SB == 2 ? rd16(p) : uint32_t(*p);
                pay = rd32(p + SB);
            } else {  // the last few bytes of the file: bounded reads
                if (p + SB > end) return nullptr;
                w = SB == 2 ? rd16(p) : uint32_t(*p);
                const uint32_t need = w < 2 ? 2u + 2u * w : 0u;
                if (p + SB + need > end) return nullptr;
                pay = 0;
                for (uint32_t k = 0; k < need; ++k) pay |= uint32_t(p[SB + k]) << (8 * k);
            }
            if (w >= csize) return nullptr;
            const bool exc = w < 2;
            const uint32_t value = w ? pay : (pay & 0xFFFFu);       // what an exception codeword carries
            const uint32_t len = exc ? 1u : uint32_t(lens[w]) + 1u;  // rows 0 and 1 are never dictionary entries
            a += exc ? value + 1u : sums[w];
            p += SB + (exc ? 2u + 2u * w : 0u);
            i += len;
            if (i > count) return nullptr;  // entries must tile the list exactly (`i != n` loop)
            const uint32_t run = uint32_t(w - 2u < kReserved - 2u);  // symbols 2..6
            has_run |= run;
            ++cw;
            // 0, 1: exceptions; 2: runs; 3..7: entries of 1, 2, 4, 8, 16 integers; 10: integers produced by runs
            ++hist[exc ? w : run ? 2u : 3u + uint32_t(__builtin_ctz(len))];
            hist[10] += run ? len : 0u;
            each(p, i, a);
            if (i == count || cw >= strip_codewords || i - s_i >= strip_ints) {
                strip(p, i, a, cw, has_run != 0);
                cw = 0;
                has_run = 0;
                s_i = i;
            }
        }
        acc = a;
        return p;
    }
};

// Cuts the codeword streams of a vroom file into strips and packs them into strip units
// (include/dint_b200.h: dint_strip_unit). Fed codeword by codeword by the walker.
struct strip_builder {
    const dint::strip_limits lim;
    const uint8_t* bin;
    std::vector<dint_strip_unit>& units;
    dint_strip_unit cur;
    bool open = false;
    // the strip being filled
    const uint8_t* s_ptr = nullptr;
    uint64_t s_out = 0;
    uint32_t s_cw = 0, s_base = 0, s_sel = 0;
    bool s_list_start = false, s_has_run = false;
    // codeword statistics (dint_vroom_plan_codeword_stats)
    uint64_t* hist = nullptr;
    uint32_t u_max_cw = 0;

    strip_builder(const uint8_t* b, std::vector<dint_strip_unit>& u) : lim(dint::get_strip_limits()), bin(b), units(u) {
        std::memset(&cur, 0, sizeof cur);
    }
    // a codeword stream begins at p (first codeword), producing output from index out_idx on
    void begin_stream(const uint8_t* p, uint64_t out_idx, uint32_t sel, uint32_t acc, bool list_start) {
        s_ptr = p;
        s_out = out_idx;
        s_cw = 0;
        s_base = acc;
        s_sel = sel;
        s_list_start = list_start;
        s_has_run = false;
    }
    // a strip of `codewords` codewords ends at q: out_idx integers are out, acc = sum(v + 1) of the list so far
    void strip(const uint8_t* q, uint64_t out_idx, uint32_t acc, uint32_t codewords, bool has_run) {
        s_cw = codewords;
        s_has_run = has_run;
        emit(q, out_idx);
        begin_stream(q, out_idx, s_sel, acc, false);
    }
    void close_unit() {
        if (!open) return;
        if (hist) {
            hist[11] += u_max_cw;  // trips of the lane-sequential loop = longest strip of the unit
            hist[12] += cur.n_strips;
        }
        u_max_cw = 0;
        for (uint32_t k = cur.n_strips; k < DINT_UNIT_STRIPS; ++k) cur.strip[k] = uint32_t(cur.n_ints) << 13;
        units.push_back(cur);
        open = false;
    }
    void emit(const uint8_t* end_ptr, uint64_t out_end) {
        const uint32_t n = uint32_t(out_end - s_out);
        const uint64_t in0 = uint64_t(s_ptr - bin), in1 = uint64_t(end_ptr - bin);
        if (open && (cur.n_strips >= DINT_UNIT_STRIPS || in1 - cur.in_off > lim.unit_bytes ||
                     cur.n_ints + n > lim.unit_ints || s_out != cur.out_off + cur.n_ints))
            close_unit();
        if (!open) {
            std::memset(&cur, 0, sizeof cur);
            cur.in_off = in0;
            cur.out_off = s_out;
            cur.doc_base = s_base;
            open = true;
        }
        u_max_cw = std::max(u_max_cw, s_cw);
        if (hist) hist[s_sel >= kSelectors ? 8 : 9] += s_cw;
        const uint32_t k = cur.n_strips;
        cur.strip[k] = uint32_t(in0 - cur.in_off) | (uint32_t(cur.n_ints) << 13) | (s_sel << 25);
        if (k && s_list_start) cur.list_starts |= 1u << k;
        if (s_has_run) cur.flags |= DINT_UNIT_HAS_RUNS;
        cur.n_ints = uint16_t(cur.n_ints + n);
        cur.in_len = uint16_t(in1 - cur.in_off);
        cur.n_strips = uint8_t(k + 1);
    }
};

}  // namespace

extern "C" {

static int vroom_plan_create_impl(const dint_dict* d, const void* bin_, size_t len, uint32_t block_size, uint32_t chunk_ints,
                                  dint_vroom_plan** out);
int dint_vroom_plan_create(const dint_dict* d, const void* bin_, size_t len, uint32_t block_size, uint32_t chunk_ints,
                           dint_vroom_plan** out) {
    try {
        return vroom_plan_create_impl(d, bin_, len, block_size, chunk_ints, out);
    } catch (const std::exception&) {
        set_error("vroom .bin: out of memory while building the plan");
        return DINT_E_NOMEM;
    }
}
static int vroom_plan_create_impl(const dint_dict* d, const void* bin_, size_t len, uint32_t block_size, uint32_t chunk_ints,
                                  dint_vroom_plan** out) {
    if (!d || (!bin_ && len) || !out) return DINT_E_ARG;
    g_error.clear();
    if (!block_size) block_size = kDefaultBlock;
    if (!chunk_ints) chunk_ints = kDefaultChunk;
    const uint8_t* bin = static_cast<const uint8_t*>(bin_);
    const uint8_t* end = bin + len;
    std::unique_ptr<dint_vroom_plan> plan(new (std::nothrow) dint_vroom_plan);
    if (!plan) return DINT_E_NOMEM;
    walker wk{d, bin, end};
    strip_builder sb(bin, plan->strip_units);
    sb.hist = plan->cw_stats;
    const uint8_t* p = bin;
    uint64_t out_off = 0;
    plan->list_offsets.push_back(0);
    while (p != end) {
        plan->list_byte_offsets.push_back(uint64_t(p - bin));
        uint32_t n = 0, universe = 0;
        // header::read (vroom_env/codecs.hpp:117-123): n, then universe (unused by DINT decoders)
        p = vbyte(p, end, &n);
        if (p) p = vbyte(p, end, &universe);
        if (!p) {
            set_error("vroom .bin: truncated list header");
            return DINT_E_FORMAT;
        }
        uint32_t acc = 0;
        if (d->kind != DINT_DICT_MULTI) {
            // single_dint::decode (vroom_env/dint_codecs.hpp:37-107): one u16 stream per list
            const uint8_t* seg_p = p;
            uint64_t seg_i = 0;
            uint32_t seg_base = 0;
            sb.begin_stream(p, out_off, 0, 0, true);
            auto strip = [&](const uint8_t* q, uint64_t i, uint32_t a, uint32_t cw, bool has_run) {
                sb.strip(q, out_off + i, a, cw, has_run);
            };
            auto each = [&](const uint8_t* q, uint64_t i, uint32_t a) {  // chunks of a long list for the tile decoder
                if (i - seg_i >= chunk_ints && i < n) {
                    plan->segments.push_back({uint64_t(seg_p - bin), out_off + seg_i, uint32_t(i - seg_i), seg_base,
                                              uint32_t(q - seg_p), a - seg_base});
                    seg_p = q;
                    seg_i = i;
                    seg_base = a;
                }
            };
            p = wk.walk<2>(p, 0, n, acc, sb.lim.strip_codewords, sb.lim.strip_ints, plan->cw_stats, strip, each);
            if (!p) {
                set_error("vroom .bin: codeword stream does not tile the list");
                return DINT_E_FORMAT;
            }
            if (n > seg_i)
                plan->segments.push_back({uint64_t(seg_p - bin), out_off + seg_i, uint32_t(n - seg_i), seg_base,
                                          uint32_t(p - seg_p), acc - seg_base});
        } else {
            // multi_opt_dint::decode (vroom_env/dint_codecs.hpp:521-619): ceil(n / block_size) blocks,
            // each = selector byte + u16 stream (selector < 6) or u8 stream (selector - 6)
            uint64_t done = 0;
            while (done < n) {
                uint32_t size = uint32_t(std::min<uint64_t>(block_size, n - done));
                if (p >= end) {
                    set_error("vroom .bin: truncated multi block");
                    return DINT_E_FORMAT;
                }
                uint32_t sel = *p;
                if (sel >= 2 * kSelectors || (sel % kSelectors) >= d->n_classes) {
                    set_error("vroom .bin: bad selector byte");
                    return DINT_E_FORMAT;
                }
                const uint32_t acc0 = acc;
                sb.begin_stream(p + 1, out_off + done, sel, acc, done == 0);
                auto strip = [&](const uint8_t* qq, uint64_t i, uint32_t a, uint32_t cw, bool has_run) {
                    sb.strip(qq, out_off + done + i, a, cw, has_run);
                };
                auto each = [](const uint8_t*, uint64_t, uint32_t) {};
                const uint8_t* q =
                    sel < kSelectors
                        ? wk.walk<2>(p + 1, sel, size, acc, sb.lim.strip_codewords, sb.lim.strip_ints, plan->cw_stats, strip, each)
                        : wk.walk<1>(p + 1, sel - kSelectors, size, acc, sb.lim.strip_codewords, sb.lim.strip_ints,
                                     plan->cw_stats, strip, each);
                if (!q) {
                    set_error("vroom .bin: codeword stream does not tile the block");
                    return DINT_E_FORMAT;
                }
                plan->segments.push_back({uint64_t(p - bin), out_off + done, size, acc0, uint32_t(q - p), acc - acc0});
                p = q;
                done += size;
            }
        }
        out_off += n;
        plan->list_offsets.push_back(out_off);
    }
    sb.close_unit();
    plan->list_byte_offsets.push_back(len);
    plan->n_ints = out_off;
    if (plan->segments.size() > 0xFFFFFFF0ull) {
        set_error("vroom .bin: more than 2^32 segments");
        return DINT_E_ARG;
    }
    // work units: runs of consecutive segments (contiguous input and output) of ~kUnitInts integers
    plan->unit_first.push_back(0);
    uint64_t unit_ints = 0;
    for (size_t i = 0; i < plan->segments.size(); ++i) {
        unit_ints += plan->segments[i].n;
        if (unit_ints >= kUnitInts || i + 1 - plan->unit_first.back() >= kUnitSegments) {
            plan->unit_first.push_back(uint32_t(i + 1));
            unit_ints = 0;
        }
    }
    if (plan->unit_first.back() != plan->segments.size()) plan->unit_first.push_back(uint32_t(plan->segments.size()));
    *out = plan.release();
    return DINT_OK;
}

void dint_vroom_plan_destroy(dint_vroom_plan* p) { delete p; }
uint64_t dint_vroom_plan_num_lists(const dint_vroom_plan* p) { return p ? p->list_offsets.size() - 1 : 0; }
uint64_t dint_vroom_plan_num_ints(const dint_vroom_plan* p) { return p ? p->n_ints : 0; }
uint64_t dint_vroom_plan_num_segments(const dint_vroom_plan* p) { return p ? p->segments.size() : 0; }
const dint_segment* dint_vroom_plan_segments(const dint_vroom_plan* p) { return p ? p->segments.data() : nullptr; }
void dint_vroom_plan_codeword_stats(const dint_vroom_plan* p, uint64_t out[16]) {
    for (int i = 0; i < 16; ++i) out[i] = p ? p->cw_stats[i] : 0;
}
uint64_t dint_vroom_plan_num_strip_units(const dint_vroom_plan* p) { return p ? p->strip_units.size() : 0; }
const dint_strip_unit* dint_vroom_plan_strip_units(const dint_vroom_plan* p) {
    return p ? p->strip_units.data() : nullptr;
}
uint64_t dint_vroom_plan_num_units(const dint_vroom_plan* p) { return p ? p->unit_first.size() - 1 : 0; }
const uint32_t* dint_vroom_plan_unit_first(const dint_vroom_plan* p) { return p ? p->unit_first.data() : nullptr; }
const uint64_t* dint_vroom_plan_list_offsets(const dint_vroom_plan* p) { return p ? p->list_offsets.data() : nullptr; }
const uint64_t* dint_vroom_plan_list_byte_offsets(const dint_vroom_plan* p) {
    return p ? p->list_byte_offsets.data() : nullptr;
}

// ---------------------------------------------------------------------------------------
// frozen dict_freq_index pre-pass
// ---------------------------------------------------------------------------------------
static inline uint64_t get_bits(const uint8_t* words, uint64_t pos, uint32_t len) {
    if (!len) return 0;
    uint64_t w = pos / 64, s = pos % 64;
    uint64_t lo = rd64(words + w * 8) >> s;
    if (s + len > 64) lo |= rd64(words + (w + 1) * 8) << (64 - s);
    return len == 64 ? lo : lo & ((uint64_t(1) << len) - 1);
}

static int index_plan_create_impl(int kind, const void* bin_, size_t len, uint32_t block_size, dint_index_plan** out);
int dint_index_plan_create(int kind, const void* bin_, size_t len, uint32_t block_size, dint_index_plan** out) {
    try {
        return index_plan_create_impl(kind, bin_, len, block_size, out);
    } catch (const std::exception&) {
        set_error("index .bin: out of memory while building the plan");
        return DINT_E_NOMEM;
    }
}
static int index_plan_create_impl(int kind, const void* bin_, size_t len, uint32_t block_size, dint_index_plan** out) {
    if (!bin_ || !out || kind < 0 || kind > DINT_INDEX_BLOCK_U32) return DINT_E_ARG;
    g_error.clear();
    if (!block_size) block_size = kDefaultBlock;
    const uint8_t* bin = static_cast<const uint8_t*>(bin_);
    std::unique_ptr<dint_index_plan> plan(new (std::nothrow) dint_index_plan);
    if (!plan) return DINT_E_NOMEM;
    plan->kind = kind;
    auto bad = [&](const char* m) {
        set_error(std::string("index .bin: ") + m);
        return DINT_E_FORMAT;
    };
    // u64 freeze flags | global_parameters 5 x u8 | u64 m_size | u64 m_num_docs |
    // bit_vector { u64 bits, mappable_vector<u64> } | mappable_vector<u8> m_lists | dicts
    size_t off = 8;
    if (len < off + 5 + 32) return bad("truncated header");
    const uint32_t log_s0 = bin[off], log_s1 = bin[off + 1];
    off += 5;
    plan->n_lists = rd64(bin + off);
    plan->num_docs = rd64(bin + off + 8);
    const uint64_t ef_bits = rd64(bin + off + 16), ef_words = rd64(bin + off + 24);
    off += 32;
    if (ef_words > (len - off) / 8 || ef_bits > ef_words * 64) return bad("endpoints bit vector out of range");
    const uint8_t* words = bin + off;
    off += size_t(ef_words) * 8;
    if (off + 8 > len) return bad("truncated m_lists");
    plan->lists_len = rd64(bin + off);
    off += 8;
    plan->lists_off = off;
    if (plan->lists_len > len - off) return bad("m_lists out of range");
    off += size_t(plan->lists_len);
    plan->docs_dict_off = off;
    if (kind >= DINT_INDEX_BLOCK_INTERPOLATIVE) {  // block_freq_index: nothing behind m_lists
        plan->freqs_dict_off = plan->end_off = off;
    } else {
        dint_dict tmp;
        size_t used = 0;
        if (parse_mapped(&tmp, kind, bin + off, len - off, &used) != DINT_OK) return bad("docs dictionary");
        off += used;
        plan->freqs_dict_off = off;
        dint_dict tmp2;
        if (parse_mapped(&tmp2, kind, bin + off, len - off, &used) != DINT_OK) return bad("freqs dictionary");
        off += used;
        plan->end_off = off;
    }
    if (off != len) return bad("trailing bytes");

    // Elias-Fano of the n_lists list start offsets, universe = m_lists.size()
    // (dict_freq_index.hpp:96-99; compact_elias_fano.hpp:14-61)
    const uint64_t n = plan->n_lists, universe = plan->lists_len;
    // every list takes at least one byte of m_lists and one bit of the Elias-Fano vector: bound the (untrusted)
    // count before anything is sized by it
    if (n > universe || n > ef_bits) return bad("list count out of range");
    plan->endpoints.assign(n + 1, 0);
    if (n) {
        const uint64_t l = universe > n ? msb64(universe / n) : 0;
        const uint64_t hb_len = n + (universe >> l) + 2;
        const uint64_t ptr_size = hb_len > 1 ? msb64(hb_len - 1) + 1 : 0;
        const uint64_t p0 = (hb_len - n) >> log_s0, p1 = n >> log_s1;
        const uint64_t hb_off = (p0 + p1) * ptr_size, lb_off = hb_off + hb_len;
        if (lb_off + n * l != ef_bits) return bad("Elias-Fano geometry mismatch");
        uint64_t i = 0;
        for (uint64_t pos = 0; pos < hb_len && i < n;) {
            uint32_t take = uint32_t(std::min<uint64_t>(64, hb_len - pos));
            uint64_t w = get_bits(words, hb_off + pos, take);
            while (w && i < n) {
                uint32_t b = uint32_t(__builtin_ctzll(w));
                w &= w - 1;
                uint64_t high = pos + b - i - 1;
                plan->endpoints[i] = (high << l) | get_bits(words, lb_off + i * l, uint32_t(l));
                ++i;
            }
            pos += take;
        }
        if (i != n) return bad("Elias-Fano: fewer endpoints than lists");
    }
    plan->endpoints[n] = universe;

    // per-list layout (dict_posting_list.hpp:10-56): vbyte n | u32 block_max[blocks] |
    // u32 block_endpoint[blocks-1] | blocks {docs bytes, freqs bytes}
    const uint8_t* lists = bin + plan->lists_off;
    const uint8_t* lists_end = lists + plan->lists_len;
    plan->list_offsets.assign(n + 1, 0);
    uint64_t postings = 0;
    for (uint64_t t = 0; t < n; ++t) {
        if (plan->endpoints[t] >= plan->lists_len || (t && plan->endpoints[t] < plan->endpoints[t - 1]))
            return bad("list endpoint out of range");
        const uint8_t* data = lists + plan->endpoints[t];
        uint32_t ln = 0;
        const uint8_t* base = vbyte(data, lists_end, &ln);
        if (!base || !ln) return bad("bad list header");
        const uint64_t blocks = (uint64_t(ln) + block_size - 1) / block_size;
        // offsets are compared as integers: a corrupt n must not move a pointer past the buffer first
        const uint64_t tables = 4 * blocks + 4 * (blocks - 1);
        if (tables > uint64_t(lists_end - base)) return bad("block tables out of range");
        const uint8_t* maxs = base;
        const uint8_t* eps = maxs + 4 * blocks;
        const uint8_t* blocks_data = eps + 4 * (blocks - 1);
        plan->list_offsets[t] = postings;
        for (uint64_t b = 0; b < blocks; ++b) {
            dint_block blk;
            uint32_t endpoint = b ? rd32(eps + 4 * (b - 1)) : 0;
            blk.in_off = uint64_t(blocks_data - lists) + endpoint;
            if (blk.in_off >= plan->lists_len) return bad("block endpoint out of range");
            blk.out_off = postings + b * block_size;
            blk.size = ((b + 1) * block_size <= ln) ? block_size : (ln % block_size);
            blk.cur_base = (b ? rd32(maxs + 4 * (b - 1)) : 0xFFFFFFFFu) + 1;
            blk.block_max = rd32(maxs + 4 * b);
            blk.reserved = 0;
            plan->blocks.push_back(blk);
        }
        postings += ln;
    }
    plan->list_offsets[n] = postings;
    plan->n_postings = postings;
    *out = plan.release();
    return DINT_OK;
}

void dint_index_plan_destroy(dint_index_plan* p) { delete p; }
uint64_t dint_index_plan_num_lists(const dint_index_plan* p) { return p ? p->n_lists : 0; }
uint64_t dint_index_plan_num_docs(const dint_index_plan* p) { return p ? p->num_docs : 0; }
uint64_t dint_index_plan_num_postings(const dint_index_plan* p) { return p ? p->n_postings : 0; }
uint64_t dint_index_plan_num_blocks(const dint_index_plan* p) { return p ? p->blocks.size() : 0; }
const dint_block* dint_index_plan_blocks(const dint_index_plan* p) { return p ? p->blocks.data() : nullptr; }
const uint64_t* dint_index_plan_list_offsets(const dint_index_plan* p) { return p ? p->list_offsets.data() : nullptr; }
const uint64_t* dint_index_plan_endpoints(const dint_index_plan* p) { return p ? p->endpoints.data() : nullptr; }
int dint_index_plan_layout(const dint_index_plan* p, uint64_t* lists_off, uint64_t* lists_len, uint64_t* docs_dict_off,
                           uint64_t* freqs_dict_off, uint64_t* end_off) {
    if (!p) return DINT_E_ARG;
    if (lists_off) *lists_off = p->lists_off;
    if (lists_len) *lists_len = p->lists_len;
    if (docs_dict_off) *docs_dict_off = p->docs_dict_off;
    if (freqs_dict_off) *freqs_dict_off = p->freqs_dict_off;
    if (end_off) *end_off = p->end_off;
    return DINT_OK;
}

}  // extern "C"
