/* TEST INFRASTRUCTURE — CPU restatement of the reference's DINT decode path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this; the product library (dint_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference compiled into oracle/_ref/libdint_ref.so (when
 * /root/reference is mounted) and against the committed golden fixtures produced by
 * the reference (tests/golden/manifest.json), which the reference's own checkers
 * (verify_collection, check_encoded_data) accepted.
 */
#ifndef DINT_ORACLE_H
#define DINT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_RECT = 0, ORACLE_PACKED = 1, ORACLE_MULTI = 2,
       /* index files only: block_freq_index<interpolative_block> (include/ds2i/block_freq_index.hpp, block_posting_list.hpp:
        * the container dict_freq_index mirrors) — no dictionaries, every block binary-interpolative coded */
       ORACLE_BLOCK_INTERPOLATIVE = 3,
       ORACLE_BLOCK_VBYTE = 4, /* block_freq_index<vbyte_block>: TightVariableByte blocks (block_codecs.hpp:357-375) */
       ORACLE_BLOCK_U32 = 5 /* block_freq_index<u32_block>: raw u32 blocks (block_codecs.hpp:377-402) */ };

typedef struct oracle_dict oracle_dict;

/* dictionary file -> in-memory dictionary (builder::load of the three dictionary types) */
oracle_dict* oracle_dict_parse(int kind, const uint8_t* file, size_t len);
/* same, from the (m_offsets, m_table) / (m_table) / (m_start_offsets, m_offsets, m_table)
 * mappable_vectors embedded in a frozen index; returns bytes consumed through *consumed */
oracle_dict* oracle_dict_parse_mapped(int kind, const uint8_t* p, size_t len, size_t* consumed);
void oracle_dict_free(oracle_dict* d);
/* Dictionary::copy: writes 16 words to out, returns the entry's length */
uint32_t oracle_dict_copy(const oracle_dict* d, uint32_t cls, uint32_t i, uint32_t* out);

size_t oracle_vbyte_decode(const uint8_t* in, uint32_t* out, size_t n);
size_t oracle_interpolative_decode(const uint8_t* in, uint32_t* out, uint32_t sum_of_values, size_t n);

/* whole-list decode; out must be zeroed and have n + 256 + 16 slots; returns bytes consumed */
size_t oracle_vroom_decode(const oracle_dict* d, const uint8_t* in, uint32_t* out, size_t n, uint32_t block_size);
/* block decode (n <= block_size); out zeroed with 2*block_size slots; returns bytes consumed */
size_t oracle_block_decode(const oracle_dict* d, const uint8_t* in, uint32_t* out, uint32_t sum_of_values, size_t n,
                           uint32_t block_size);

/* whole vroom .bin with check_encoded_data semantics (fresh zeroed buffer per list);
 * list_n (optional) receives each list's length. returns #lists or -1 */
long oracle_vroom_decode_file(const oracle_dict* d, const uint8_t* bin, size_t len, uint32_t block_size, uint32_t* out,
                              size_t out_cap, uint32_t* list_n, size_t list_cap, uint64_t* total_ints);

typedef struct {
    uint64_t n_lists, num_docs;
    uint64_t endpoints_bits;     /* bit length of m_endpoints */
    size_t endpoints_words_off;  /* byte offset of the u64 words of m_endpoints */
    size_t lists_off, lists_len; /* byte offset / length of m_lists */
    size_t docs_dict_off, freqs_dict_off, end_off;
} oracle_index_info;

/* succinct::mapper layout of dict_freq_index (dict_freq_index.hpp:209-214) */
int oracle_index_parse(int kind, const uint8_t* bin, size_t len, oracle_index_info* info);
/* Elias-Fano decode of the n_lists list start offsets (compact_elias_fano.hpp:14-61, 69-136) */
int oracle_index_endpoints(const uint8_t* bin, const oracle_index_info* info, uint64_t* endpoints /* n_lists */);
/* full scan with document_enumerator semantics: docids and freqs of every list, back to back.
 * list_n optional. returns #postings or -1 */
long oracle_index_decode(int kind, const uint8_t* bin, size_t len, uint32_t block_size, uint32_t* docids,
                         uint32_t* freqs, size_t cap, uint32_t* list_n, size_t list_cap);

#ifdef __cplusplus
}
#endif
#endif
