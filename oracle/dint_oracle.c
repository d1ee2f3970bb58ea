/* TEST INFRASTRUCTURE — see dint_oracle.h. Plain-C restatement of the reference's
 * DINT decode path; every function cites the reference lines it follows
 * (paths relative to /root/reference). Scalar, single-threaded, no tricks.
 */
#include "dint_oracle.h"

#include <stdlib.h>
#include <string.h>

#define ENTRY 16u     /* constants::max_entry_size, include/dint/dint_configuration.hpp:23 */
#define NSEL 6u       /* constants::num_selectors,  include/dint/dint_configuration.hpp:20 */
#define EXC 2u        /* EXCEPTIONS,               include/dint/dint_configuration.hpp:6 */

struct oracle_dict {
    int kind;
    uint32_t n_start;
    uint32_t* start_offsets; /* multi only */
    uint32_t n_offsets;
    uint32_t* offsets;       /* packed / multi */
    uint32_t n_table;
    uint32_t* table;         /* padded with ENTRY zero words */
};

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

static uint32_t* dup_words(const uint8_t* p, size_t n, size_t pad) {
    uint32_t* w = (uint32_t*)calloc(n + pad + 1, 4);
    if (w && n) memcpy(w, p, n * 4);
    return w;
}

/* rectangular_dictionary::builder::load   include/dint/rectangular_dictionary.hpp:79-92 (+init :43-56)
 * single_dictionary::builder::load        include/dint/single_dictionary.hpp:88-107
 * multi_dictionary::builder::load         include/dint/multi_dictionary.hpp:93-121 (pads table by 16, :108) */
oracle_dict* oracle_dict_parse(int kind, const uint8_t* f, size_t len) {
    oracle_dict* d = (oracle_dict*)calloc(1, sizeof *d);
    if (!d) return NULL;
    d->kind = kind;
    if (kind == ORACLE_RECT) {
        if (len < 4) goto bad;
        uint32_t m_size = rd32(f);
        size_t words = (size_t)m_size * (ENTRY + 1);
        if (m_size > 65536 || len < 4 + words * 4) goto bad;
        /* init(): full 65536 x 17 table, zero filled, reserved rows' lengths set; file rows overwrite */
        d->n_table = 65536u * (ENTRY + 1);
        d->table = (uint32_t*)calloc((size_t)d->n_table + ENTRY, 4);
        uint32_t pos = ENTRY + 1;
        for (uint32_t i = 0; i < EXC; ++i, pos += ENTRY + 1) d->table[pos - 1] = 1;
        for (uint32_t i = 0, size = 256; i < 5; ++i, pos += ENTRY + 1, size /= 2) d->table[pos - 1] = size;
        memcpy(d->table, f + 4, words * 4);
        return d;
    }
    if (kind == ORACLE_PACKED) {
        if (len < 12) goto bad;
        uint32_t no = rd32(f + 4), nt = rd32(f + 8);
        if (len < 12 + ((size_t)no + nt) * 4) goto bad;
        d->n_offsets = no;
        d->offsets = dup_words(f + 12, no, 0);
        d->n_table = nt;
        d->table = dup_words(f + 12 + (size_t)no * 4, nt, ENTRY);
        return d;
    }
    if (kind == ORACLE_MULTI) {
        if (len < 16) goto bad;
        uint32_t ns = rd32(f + 4), no = rd32(f + 8), nt = rd32(f + 12);
        if (len < 16 + ((size_t)ns + no + nt) * 4) goto bad;
        d->n_start = ns;
        d->start_offsets = dup_words(f + 16, ns, 0);
        d->n_offsets = no;
        d->offsets = dup_words(f + 16 + (size_t)ns * 4, no, 0);
        d->n_table = nt;
        d->table = dup_words(f + 16 + ((size_t)ns + no) * 4, nt, ENTRY);
        return d;
    }
bad:
    oracle_dict_free(d);
    return NULL;
}

/* mappable_vector<T> = u64 size + payload  external/succinct/mapper.hpp:128-146
 * rect  ::map visits m_table                       rectangular_dictionary.hpp:219-222
 * single::map visits m_offsets, m_table            single_dictionary.hpp:245-248
 * multi ::map visits m_start_offsets, m_offsets, m_table  multi_dictionary.hpp:312-316 */
oracle_dict* oracle_dict_parse_mapped(int kind, const uint8_t* p, size_t len, size_t* consumed) {
    oracle_dict* d = (oracle_dict*)calloc(1, sizeof *d);
    size_t off = 0;
    if (!d) return NULL;
    d->kind = kind;
    int nvec = kind == ORACLE_RECT ? 1 : kind == ORACLE_PACKED ? 2 : 3;
    for (int v = 0; v < nvec; ++v) {
        if (off + 8 > len) goto bad;
        uint64_t n = rd64(p + off);
        off += 8;
        if (off + n * 4 > len) goto bad;
        int is_table = (v == nvec - 1);
        uint32_t* w = dup_words(p + off, (size_t)n, is_table ? ENTRY : 0);
        off += (size_t)n * 4;
        if (is_table) { d->table = w; d->n_table = (uint32_t)n; }
        else if (kind == ORACLE_MULTI && v == 0) { d->start_offsets = w; d->n_start = (uint32_t)n; }
        else { d->offsets = w; d->n_offsets = (uint32_t)n; }
    }
    if (consumed) *consumed = off;
    return d;
bad:
    oracle_dict_free(d);
    return NULL;
}

void oracle_dict_free(oracle_dict* d) {
    if (!d) return;
    free(d->start_offsets);
    free(d->offsets);
    free(d->table);
    free(d);
}

/* rectangular_dictionary::copy  rectangular_dictionary.hpp:206-213
 * single_dictionary::copy       single_dictionary.hpp:230-238
 * multi_dictionary::copy        multi_dictionary.hpp:293-304 */
uint32_t oracle_dict_copy(const oracle_dict* d, uint32_t cls, uint32_t i, uint32_t* out) {
    if (d->kind == ORACLE_RECT) {
        const uint32_t* row = d->table + (size_t)i * (ENTRY + 1);
        memcpy(out, row, ENTRY * 4);
        return row[ENTRY];
    }
    uint32_t so = d->kind == ORACLE_MULTI ? d->offsets[d->start_offsets[cls] + i] : d->offsets[i];
    memcpy(out, d->table + (so & 0xFFFFFFu), ENTRY * 4);
    return (so >> 24) + 1;
}

/* TightVariableByte::decode  include/ds2i/block_codecs.hpp:87-101 (== vroom_env/codecs.hpp:93-107):
 * 7-bit groups, LSB first, the byte with bit 7 SET terminates */
size_t oracle_vbyte_decode(const uint8_t* in, uint32_t* out, size_t n) {
    const uint8_t* p = in;
    for (size_t i = 0; i < n; ++i) {
        uint32_t v = 0;
        for (unsigned shift = 0;; shift += 7) {
            uint8_t c = *p++;
            v += (uint32_t)(c & 127u) << shift;
            if (c & 128u) break;
        }
        out[i] = v;
    }
    return (size_t)(p - in);
}

/* bit_reader  include/ds2i/interpolative_coding.hpp:79-153 */
typedef struct { const uint8_t* in; uint32_t avail; uint64_t buf; size_t pos; } bitrd;

static uint32_t br_read(bitrd* r, uint32_t len) { /* :93-107 */
    if (!len) return 0;
    if (r->avail < len) {
        r->buf |= (uint64_t)rd32(r->in) << r->avail;
        r->in += 4;
        r->avail += 32;
    }
    uint32_t val = (uint32_t)(r->buf & (((uint64_t)1 << len) - 1));
    r->buf >>= len;
    r->avail -= len;
    r->pos += len;
    return val;
}

static uint32_t msb32(uint32_t x) { uint32_t b = 0; while (x >>= 1) ++b; return b; }
static uint32_t msb64(uint64_t x) { uint32_t b = 0; while (x >>= 1) ++b; return b; }

static uint32_t br_read_int(bitrd* r, uint32_t u) { /* :109-122, truncated binary */
    uint32_t b = msb32(u);
    uint64_t m = ((uint64_t)1 << (b + 1)) - u;
    uint32_t val = br_read(r, b);
    if (val >= m) val = (uint32_t)(((uint64_t)val << 1) + br_read(r, 1) - m);
    return val;
}

static void br_read_interpolative(bitrd* r, uint32_t* out, size_t n, uint32_t low, uint32_t high) { /* :124-146 */
    size_t h = n / 2;
    uint32_t val = low + br_read_int(r, high - low + 1);
    out[h] = val;
    if (n == 1) return;
    if (h) br_read_interpolative(r, out, h, low, val);
    if (n - h - 1) br_read_interpolative(r, out + h + 1, n - h - 1, val, high);
}

/* interpolative_block::decode  include/ds2i/block_codecs.hpp:130-150 */
size_t oracle_interpolative_decode(const uint8_t* in, uint32_t* out, uint32_t sum_of_values, size_t n) {
    const uint8_t* p = in;
    if (sum_of_values == 0xFFFFFFFFu) p += oracle_vbyte_decode(p, &sum_of_values, 1);
    out[n - 1] = sum_of_values;
    size_t bytes = 0;
    if (n > 1) {
        bitrd r = {p, 0, 0, 0};
        br_read_interpolative(&r, out, n - 1, 0, sum_of_values);
        for (size_t i = n - 1; i > 0; --i) out[i] -= out[i - 1];
        bytes = (r.pos + 7) / 8;
    }
    return (size_t)(p - in) + bytes;
}

/* the u16 codeword loop shared by
 *   single_dint::decode          vroom_env/dint_codecs.hpp:37-107
 *   multi_opt_dint::decode       vroom_env/dint_codecs.hpp:538-571
 *   dint_block::decode           include/dint/dint_codecs.hpp:21-44
 *   opt_dint_multi_dict_block    include/dint/dint_codecs.hpp:469-485 */
static const uint8_t* loop16(const oracle_dict* d, uint32_t cls, const uint8_t* p, uint32_t* out, size_t n) {
    for (size_t i = 0; i != n; p += 2) {
        uint32_t index = rd16(p);
        uint32_t k = 1;
        if (index > EXC - 1) {
            k = oracle_dict_copy(d, cls, index, out);
        } else if (index == 1) { /* 4-byte exception */
            p += 2;
            *out = rd32(p);
            p += 2;
        } else { /* 2-byte exception */
            p += 2;
            *out = rd16(p);
        }
        out += k;
        i += k;
    }
    return p;
}

/* the u8 codeword loop: vroom_env/dint_codecs.hpp:574-609, include/dint/dint_codecs.hpp:487-507 */
static const uint8_t* loop8(const oracle_dict* d, uint32_t cls, const uint8_t* p, uint32_t* out, size_t n) {
    for (size_t i = 0; i != n; p += 1) {
        uint32_t index = *p;
        uint32_t k = 1;
        if (index > EXC - 1) {
            k = oracle_dict_copy(d, cls, index, out);
        } else if (index == 1) {
            p += 1;
            *out = rd32(p);
            p += 3;
        } else {
            p += 1;
            *out = rd16(p);
            p += 1;
        }
        out += k;
        i += k;
    }
    return p;
}

size_t oracle_vroom_decode(const oracle_dict* d, const uint8_t* in, uint32_t* out, size_t n, uint32_t block_size) {
    if (d->kind != ORACLE_MULTI) return (size_t)(loop16(d, 0, in, out, n) - in); /* dint_codecs.hpp:315-323 */
    /* multi_opt_dint::decode  vroom_env/dint_codecs.hpp:521-619 */
    const uint8_t* p = in;
    size_t num_blocks = (n + block_size - 1) / block_size;
    size_t tail = n - (n / block_size * block_size);
    for (size_t b = 0; b != num_blocks; ++b) {
        size_t size = block_size;
        if (b == num_blocks - 1 && tail != 0) size = tail;
        uint8_t sel = *p;
        if (sel < NSEL) p = loop16(d, sel, p + 1, out, size);
        else p = loop8(d, sel - NSEL, p + 1, out, size);
        out += size;
    }
    return (size_t)(p - in);
}

size_t oracle_block_decode(const oracle_dict* d, const uint8_t* in, uint32_t* out, uint32_t sum_of_values, size_t n,
                           uint32_t block_size) {
    /* dint_block::decode :17-19 / opt_dint_multi_dict_block::decode :464-466 */
    if (n < block_size) return oracle_interpolative_decode(in, out, sum_of_values, n);
    if (d->kind != ORACLE_MULTI) return (size_t)(loop16(d, 0, in, out, n) - in);
    uint8_t sel = *in; /* :468-509 */
    if (sel < NSEL) return (size_t)(loop16(d, sel, in + 1, out, n) - in);
    return (size_t)(loop8(d, sel - NSEL, in + 1, out, n) - in);
}

/* vroom_env/check_encoded_data.cpp:76-110 (fresh zeros behind every list) over
 * vroom_env/codecs.hpp:117-123 header::read = vbyte n, vbyte universe */
long oracle_vroom_decode_file(const oracle_dict* d, const uint8_t* bin, size_t len, uint32_t block_size, uint32_t* out,
                              size_t out_cap, uint32_t* list_n, size_t list_cap, uint64_t* total_ints) {
    const uint8_t* p = bin;
    const uint8_t* end = bin + len;
    size_t pos = 0, cap = 0;
    long lists = 0;
    uint32_t* buf = NULL;
    while (p != end) {
        uint32_t hdr[2];
        p += oracle_vbyte_decode(p, hdr, 2);
        size_t n = hdr[0];
        if (pos + n > out_cap || (list_n && (size_t)lists >= list_cap)) { free(buf); return -1; }
        if (n + 512 > cap) { free(buf); cap = 2 * (n + 512); buf = (uint32_t*)malloc(cap * 4); }
        memset(buf, 0, (n + 512) * 4);
        p += oracle_vroom_decode(d, p, buf, n, block_size);
        memcpy(out + pos, buf, n * 4);
        if (list_n) list_n[lists] = (uint32_t)n;
        pos += n;
        ++lists;
    }
    free(buf);
    if (total_ints) *total_ints = pos;
    return lists;
}

/* dict_freq_index::map  include/dint/dict_freq_index.hpp:209-214 through
 * succinct map_visitor  external/succinct/mapper.hpp:100-156:
 *   u64 freeze flags | global_parameters (5 x u8, global_parameters.hpp:14-24) | u64 m_size |
 *   u64 m_num_docs | bit_vector{u64 m_size(bits), mappable_vector<u64> m_bits} (bit_vector.hpp:228-231) |
 *   mappable_vector<u8> m_lists | docs dict | freqs dict */
int oracle_index_parse(int kind, const uint8_t* bin, size_t len, oracle_index_info* info) {
    size_t off = 8 + 5;
    if (len < off + 16 + 16) return -1;
    info->n_lists = rd64(bin + off); off += 8;
    info->num_docs = rd64(bin + off); off += 8;
    info->endpoints_bits = rd64(bin + off); off += 8;
    uint64_t words = rd64(bin + off); off += 8;
    info->endpoints_words_off = off;
    off += (size_t)words * 8;
    if (off + 8 > len) return -1;
    info->lists_len = (size_t)rd64(bin + off); off += 8;
    info->lists_off = off;
    off += info->lists_len;
    if (off > len) return -1;
    info->docs_dict_off = off;
    if (kind >= ORACLE_BLOCK_INTERPOLATIVE) { /* block_freq_index.hpp:138-146: no dictionaries behind m_lists */
        info->freqs_dict_off = info->end_off = off;
        return off == len ? 0 : -2;
    }
    size_t used = 0;
    oracle_dict* d = oracle_dict_parse_mapped(kind, bin + off, len - off, &used);
    if (!d) return -1;
    oracle_dict_free(d);
    off += used;
    info->freqs_dict_off = off;
    d = oracle_dict_parse_mapped(kind, bin + off, len - off, &used);
    if (!d) return -1;
    oracle_dict_free(d);
    off += used;
    info->end_off = off;
    return off == len ? 0 : -2;
}

static uint64_t get_bits(const uint8_t* words, uint64_t pos, uint32_t len) { /* bit_vector::get_bits */
    if (!len) return 0;
    uint64_t w = pos / 64, s = pos % 64;
    uint64_t lo = rd64(words + w * 8) >> s;
    if (s + len > 64) lo |= rd64(words + (w + 1) * 8) << (64 - s);
    return len == 64 ? lo : lo & (((uint64_t)1 << len) - 1);
}

/* compact_elias_fano::offsets  include/ds2i/compact_elias_fano.hpp:14-61 and the bit layout
 * written by ::write :69-136: element i sets high bit (v >> l) + i + 1, low bits at i*l.
 * Called as write(bvb, m_endpoints.begin(), m_lists.size(), m_size, params)
 * (dict_freq_index.hpp:96-99): universe = m_lists.size(), n = number of lists. */
int oracle_index_endpoints(const uint8_t* bin, const oracle_index_info* info, uint64_t* endpoints) {
    const uint8_t* words = bin + info->endpoints_words_off;
    uint64_t universe = info->lists_len, n = info->n_lists;
    uint64_t log_s0 = bin[8], log_s1 = bin[9];
    uint64_t l = universe > n ? msb64(universe / n) : 0;
    uint64_t hb_len = n + (universe >> l) + 2;
    uint64_t ptr_size = hb_len > 1 ? msb64(hb_len - 1) + 1 : 0; /* ceil_log2, util.hpp:67-70 */
    uint64_t p0 = (hb_len - n) >> log_s0, p1 = n >> log_s1;
    uint64_t hb_off = (p0 + p1) * ptr_size;
    uint64_t lb_off = hb_off + hb_len;
    if (lb_off + n * l != info->endpoints_bits) return -1;
    uint64_t i = 0;
    for (uint64_t pos = 0; pos < hb_len && i < n; ++pos) {
        if (get_bits(words, hb_off + pos, 1)) {
            uint64_t high = pos - i - 1;
            endpoints[i] = (high << l) | get_bits(words, lb_off + i * l, (uint32_t)l);
            ++i;
        }
    }
    return i == n ? 0 : -1;
}

/* document_enumerator ctor + decode_docs_block + decode_freqs_block + next()
 * include/dint/dict_posting_list.hpp:90-107, :284-318, :113-124, :164-169 */
long oracle_index_decode(int kind, const uint8_t* bin, size_t len, uint32_t block_size, uint32_t* docids,
                         uint32_t* freqs, size_t cap, uint32_t* list_n, size_t list_cap) {
    oracle_index_info info;
    if (oracle_index_parse(kind, bin, len, &info) != 0) return -1;
    if (list_n && info.n_lists > list_cap) return -1;
    uint64_t* ep = (uint64_t*)malloc((info.n_lists + 1) * 8);
    if (oracle_index_endpoints(bin, &info, ep) != 0) { free(ep); return -1; }
    const int dictless = kind >= ORACLE_BLOCK_INTERPOLATIVE; /* BlockCodec::decode for every block */
    oracle_dict* dd = dictless ? NULL : oracle_dict_parse_mapped(kind, bin + info.docs_dict_off, len - info.docs_dict_off, NULL);
    oracle_dict* fd = dictless ? NULL : oracle_dict_parse_mapped(kind, bin + info.freqs_dict_off, len - info.freqs_dict_off, NULL);
    uint32_t* dbuf = (uint32_t*)malloc((size_t)(block_size + 512) * 4);
    uint32_t* fbuf = (uint32_t*)malloc((size_t)(block_size + 512) * 4);
    size_t pos = 0;
    long rc = 0;
    for (uint64_t t = 0; t < info.n_lists && rc == 0; ++t) {
        const uint8_t* data = bin + info.lists_off + ep[t];
        uint32_t n;
        const uint8_t* base = data + oracle_vbyte_decode(data, &n, 1);
        size_t blocks = ((size_t)n + block_size - 1) / block_size;
        const uint8_t* maxs = base;
        const uint8_t* endpoints = maxs + 4 * blocks;
        const uint8_t* blocks_data = endpoints + 4 * (blocks - 1);
        if (pos + n > cap) { rc = -1; break; }
        if (list_n) list_n[t] = n;
        for (size_t b = 0; b < blocks; ++b) {
            uint32_t endpoint = b ? rd32(endpoints + 4 * (b - 1)) : 0;
            const uint8_t* block_data = blocks_data + endpoint;
            uint32_t size = ((b + 1) * block_size <= n) ? block_size : (n % block_size);
            uint32_t cur_base = (b ? rd32(maxs + 4 * (b - 1)) : 0xFFFFFFFFu) + 1;
            uint32_t bmax = rd32(maxs + 4 * b);
            memset(dbuf, 0, (size_t)(block_size + 512) * 4);
            memset(fbuf, 0, (size_t)(block_size + 512) * 4);
            const uint8_t* fdata;
            if (kind == ORACLE_BLOCK_INTERPOLATIVE) { /* block_posting_list.hpp (decode_docs_block / decode_freqs_block) over block_codecs.hpp:130-150 */
                fdata = block_data + oracle_interpolative_decode(block_data, dbuf, bmax - cur_base - (size - 1), size);
                oracle_interpolative_decode(fdata, fbuf, 0xFFFFFFFFu, size);
            } else if (kind == ORACLE_BLOCK_VBYTE) { /* vbyte_block::decode block_codecs.hpp:370-374 = TightVariableByte::decode :87-101 */
                fdata = block_data + oracle_vbyte_decode(block_data, dbuf, size);
                oracle_vbyte_decode(fdata, fbuf, size);
            } else if (kind == ORACLE_BLOCK_U32) { /* u32_block::decode block_codecs.hpp:390-401: the raw little-endian values */
                for (uint32_t i = 0; i < size; ++i) dbuf[i] = rd32(block_data + 4 * (size_t)i);
                fdata = block_data + 4 * (size_t)size;
                for (uint32_t i = 0; i < size; ++i) fbuf[i] = rd32(fdata + 4 * (size_t)i);
            } else {
                fdata = block_data + oracle_block_decode(dd, block_data, dbuf, bmax - cur_base - (size - 1), size, block_size);
                oracle_block_decode(fd, fdata, fbuf, 0xFFFFFFFFu, size, block_size);
            }
            uint32_t doc = dbuf[0] + cur_base;
            for (uint32_t i = 0; i < size; ++i) {
                if (i) doc += dbuf[i] + 1;
                docids[pos + i] = doc;
                freqs[pos + i] = fbuf[i] + 1;
            }
            pos += size;
        }
    }
    free(dbuf); free(fbuf); free(ep);
    if (dd) oracle_dict_free(dd);
    if (fd) oracle_dict_free(fd);
    return rc ? rc : (long)pos;
}
