/* dint_b200 — B200-native decoder for DINT-compressed posting lists.
 *
 * C ABI of libdint_b200.so: the drop-in boundary for the reference's DECODE path
 * (jermp/dint). Plain pointers and sizes only; no C++/torch types. All entry points
 * return 0 on success or a negative DINT_E_* code (the reference has no error channel
 * on this path: it asserts, or throws std::runtime_error on file open —
 * vroom_env/decode.cpp:26-28, 98-100).
 *
 * What each group replaces in the reference (paths relative to the reference root):
 *
 *   dint_dict_*            Dictionary::builder::load + builder::build + Dictionary::copy
 *                            include/dint/rectangular_dictionary.hpp:79-92, 206-213
 *                            include/dint/single_dictionary.hpp:88-107, 230-238
 *                            include/dint/multi_dictionary.hpp:93-121, 293-304
 *   dint_vroom_discover /  the `while (begin != end) header::read ...` walk of a vroom .bin, done by the GPU on the
 *   dint_vroom_plan_*        bytes in HBM (discover) or by one host thread (plan): where every list, block and
 *                            strip of <= 8 codewords starts
 *                            vroom_env/decode.cpp:139-150, vroom_env/codecs.hpp:110-124
 *   dint_index_device_plan_* succinct::mapper::map(dict_freq_index) + Elias-Fano endpoints +
 *   / dint_index_plan_*      document_enumerator's per-list block table, on the device / on the host;
 *   dint_index_discover      + the codeword boundaries inside the full blocks (device)
 *                            include/dint/dict_freq_index.hpp:174-181, 209-214
 *                            include/ds2i/compact_elias_fano.hpp:14-61
 *                            include/dint/dict_posting_list.hpp:90-107
 *   dint_decode_segments   single_opt_dint::decode / multi_opt_dint::decode over a batch
 *                            vroom_env/dint_codecs.hpp:37-107, 315-323, 521-619
 *   dint_decode_index      decode_docs_block + decode_freqs_block + next() over a batch,
 *                          i.e. opt_dint_single_dict_block / opt_dint_multi_dict_block::decode
 *                          and interpolative_block::decode for partial blocks
 *                            include/dint/dict_posting_list.hpp:284-318, 113-124, 164-169
 *                            include/dint/dint_codecs.hpp:13-49, 269-274, 460-510
 *                            include/ds2i/block_codecs.hpp:130-150
 *   dint_decode_strips     the same two decoders, lane-sequential: one THREAD walks one strip of
 *                          <= 8 codewords exactly like the reference's scalar loop
 *                            vroom_env/dint_codecs.hpp:37-107 (u16), :574-609 (u8)
 *   dint_decode_list_host  one Decoder::decode(dict, in, out, universe, n) call, host buffers
 *   dint_vroom_ctx_* /     the whole decode_dint() loop of vroom_env/decode.cpp:139-150 on HOST
 *   dint_vroom_decode_host   buffers: H2D of the .bin, batch decode, D2H of the result
 *   dint_decode_index_planned   dint_decode_index through the strip tables of dint_index_discover (fast path)
 *   dint_decode_index_interpolative  the same scan of a block_freq_index<interpolative_block>
 *                            include/ds2i/block_freq_index.hpp, block_posting_list.hpp, block_codecs.hpp:104-150
 *   dint_and_queries       ds2i::and_query<false> for a batch of queries over the decoded docIDs
 *                            include/ds2i/queries.hpp:34-86, src/queries.cpp:15-61
 *   dint_checksum_lists    (new) per-list 64-bit checksums used for large-scale parity
 *
 * Ownership: the caller owns every device buffer; nothing is allocated on the decode
 * path (the *_create / *_discover calls are load-time: they allocate and synchronise). Outputs are PACKED (list after list, no slack) — unlike the reference, which
 * needs a pre-zeroed buffer with 16..256 ints of slack (dint_codecs.hpp:11,
 * dict_posting_list.hpp:104-105). Handles are immutable after creation and may be used
 * from any stream of the device they were created on.
 *
 * There is NO CPU fallback: every decode entry point runs sm_100a kernels or fails.
 */
#ifndef DINT_B200_H
#define DINT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DINT_B200_ABI_VERSION 7 /* 5: device-side discovery (dint_vroom_discover, dint_index_discover, dint_device_plan_*); 6: dint_index_device_plan_create, dint_and_queries, dint_decode_index_interpolative; 7: dint_decode_index_blocks (+ vbyte_block, u32_block) */

/* dictionary kinds == the reference's three dictionary types */
enum {
    DINT_DICT_RECT = 0,   /* single_rect_dint   : rectangular_dictionary   */
    DINT_DICT_PACKED = 1, /* single_packed_dint : single_dictionary<pack>  */
    DINT_DICT_MULTI = 2,  /* multi_packed_dint  : multi_dictionary<pack>   */
    /* index files only (dint_index_plan_create / dint_index_device_plan_create): a frozen
     * block_freq_index<interpolative_block> ("block_interpolative", include/index_types.hpp:41) — the container
     * dict_freq_index mirrors (include/ds2i/block_freq_index.hpp, block_posting_list.hpp), no dictionaries, every
     * block binary-interpolative coded; decoded by dint_decode_index_blocks */
    DINT_INDEX_BLOCK_INTERPOLATIVE = 3,
    /* block_freq_index<vbyte_block> ("block_vbyte", index_types.hpp:45): TightVariableByte blocks
     * (block_codecs.hpp:24-101, 357-375) */
    DINT_INDEX_BLOCK_VBYTE = 4,
    /* block_freq_index<u32_block> ("block_u32", index_types.hpp:44): raw u32 blocks (block_codecs.hpp:377-402) */
    DINT_INDEX_BLOCK_U32 = 5
};

/* output modes of dint_decode_segments */
enum {
    DINT_MODE_RAW = 0,    /* exactly what Decoder::decode writes: d-gap-1 / freq-1        */
    DINT_MODE_DOCIDS = 1, /* fused prefix sum: docid[i] = docid[i-1] + v[i] + 1            */
    DINT_MODE_FREQS = 2   /* fused +1 shift:   freq[i] = v[i] + 1                          */
};

enum {
    DINT_OK = 0,
    DINT_E_ARG = -1,     /* null pointer / bad enum / bad size                              */
    DINT_E_FORMAT = -2,  /* malformed dictionary / .bin                                     */
    DINT_E_CUDA = -3,    /* CUDA runtime error (dint_last_error() has the string)           */
    DINT_E_NOMEM = -4,
    DINT_E_NODEVICE = -5 /* no CUDA device: there is no CPU fallback                        */
};

/* One contiguous codeword stream that starts on a codeword boundary and produces
 * exactly n integers: a whole vroom list, a chunk of a long list, or one 256-int
 * block of a multi_packed vroom list. 32 bytes. */
typedef struct dint_segment {
    uint64_t in_off;   /* byte offset of the first codeword (multi: of the selector byte)  */
    uint64_t out_off;  /* index of the first output integer                                */
    uint32_t n;        /* integers produced                                                */
    uint32_t doc_base; /* DOCIDS mode: docid[first] = doc_base + v[first]                  */
    uint32_t in_len;   /* bytes of the stream (0 = unknown: the decoder may read ahead up
                          to in_bytes)                                                      */
    uint32_t doc_sum;  /* DOCIDS mode: sum of (v + 1) over the segment, mod 2^32, so that
                          doc_base + doc_sum - 1 is its last docid. Needed only when the
                          next segment of the same work unit continues the output
                          (out_off == this out_off + n)                                     */
} dint_segment;

/* One block of a dict_freq_index posting list (docs bytes immediately followed by the
 * freqs bytes). size == block_size -> DINT coded, size < block_size -> interpolative. 32 bytes. */
typedef struct dint_block {
    uint64_t in_off;    /* byte offset of the docs block inside m_lists                    */
    uint64_t out_off;   /* posting index of the block's first posting                      */
    uint32_t size;      /* postings in the block                                           */
    uint32_t cur_base;  /* block_max[b-1] + 1 (0 for b == 0)                               */
    uint32_t block_max; /* last docid of the block                                         */
    uint32_t reserved;
} dint_block;

/* Strip tables: the work decomposition of the lane-sequential decoder (dint_decode_strips).
 * A STRIP is a piece of one codeword stream that starts on a codeword boundary, holds at most
 * a fixed number of codewords (default 8) of ONE dictionary class / symbol width and
 * produces a known number of integers; one thread decodes one strip. A STRIP UNIT is up to 32
 * consecutive strips (one per lane of a warp) whose compressed bytes lie within
 * DINT_UNIT_MAX_BYTES of each other and whose outputs are contiguous in the packed output
 * (at most DINT_UNIT_MAX_INTS integers). 160 bytes, 16-byte aligned. */
#define DINT_UNIT_STRIPS 32
#define DINT_UNIT_HAS_RUNS 1u
#define DINT_UNIT_MAX_INTS 4095u  /* out_rel has 12 bits                                    */
#define DINT_UNIT_MAX_BYTES 8191u /* in_rel has 13 bits                                     */
/* The limits this build of the library actually uses (they size the decoder's shared-memory
 * buffers; defaults 1024 integers / 1792 bytes / 8 codewords per strip; the environment
 * variables DINT_B200_UNIT_INTS, DINT_B200_UNIT_BYTES, DINT_B200_STRIP_CODEWORDS override
 * them at first use, for tuning). Units beyond them are rejected by dint_decode_strips'
 * kernels (their output is left untouched). */
void dint_strip_limits(uint32_t* unit_max_ints, uint32_t* unit_max_bytes, uint32_t* strip_codewords);
typedef struct dint_strip_unit {
    uint64_t in_off;      /* byte offset of the unit's first strip in the .bin                */
    uint64_t out_off;     /* index of the unit's first output integer                         */
    uint32_t doc_base;    /* DOCIDS mode: docid[first] = doc_base + v[first]                  */
    uint32_t list_starts; /* bit k (k >= 1): strip k begins a new list, i.e. the docID prefix
                             sum restarts there (docid = v). Strip 0 is covered by doc_base. */
    uint16_t in_len;      /* bytes from in_off to the end of the last strip                   */
    uint16_t n_ints;      /* integers the unit produces                                       */
    uint8_t n_strips;     /* 1..32                                                            */
    uint8_t flags;        /* DINT_UNIT_HAS_RUNS: some codeword of the unit is a run of zeros, which the
                             decoder does not store: it clears the unit's output image first       */
    uint8_t reserved[2];
    /* strip k: bits 0-12 in_rel (bytes from in_off to its first codeword), bits 13-24 out_rel
     * (integers before it inside the unit), bits 25-28 selector: 0..5 = u16 codewords of class
     * selector, 6..11 = u8 codewords of class selector - 6 (single dictionaries: 0). The strip
     * produces out_rel[k+1] - out_rel[k] integers (the last one n_ints - out_rel). */
    uint32_t strip[DINT_UNIT_STRIPS];
} dint_strip_unit;

typedef struct dint_dict dint_dict;
typedef struct dint_vroom_plan dint_vroom_plan;
typedef struct dint_index_plan dint_index_plan;
typedef struct dint_vroom_ctx dint_vroom_ctx;
typedef struct dint_device_plan dint_device_plan;
typedef struct dint_device_index dint_device_index;

const char* dint_last_error(void);
int dint_abi_version(void);
/* number of kernels launched by this library in the calling process so far */
uint64_t dint_kernel_launches(void);

/* ---- dictionaries --------------------------------------------------------------- */
/* Parse a reference dictionary FILE image (dict.<name>.<type>.DSF-65536-16) and upload it
 * to `device`. device < 0: host-side tables only (enough for the *_plan_* pre-passes). */
int dint_dict_create(int kind, const void* file_bytes, size_t len, int device, dint_dict** out);
/* Same for the mappable_vector images embedded at the end of a frozen index .bin. */
int dint_dict_create_mapped(int kind, const void* bytes, size_t len, int device, dint_dict** out, size_t* consumed);
void dint_dict_destroy(dint_dict* d);
int dint_dict_kind(const dint_dict* d);
uint32_t dint_dict_num_rows(const dint_dict* d);
/* bytes resident in HBM for this dictionary (rows + length table) */
size_t dint_dict_device_bytes(const dint_dict* d);
/* Dictionary::copy through the DEVICE tables (debug/parity): copies the 16-word row of
 * entry i of class cls into out16 and returns its length through *len. */
int dint_dict_copy(const dint_dict* d, uint32_t cls, uint32_t i, uint32_t* out16, uint32_t* len);

/* ---- vroom .bin pre-pass (host, one time per file) ---------------------------------- */
/* Walks the file exactly like vroom_env/decode.cpp does and emits the segment table.
 * chunk_ints: long single-dictionary lists are cut into segments of about this many
 * integers at codeword boundaries (0 = library default). block_size: the reference's
 * constants::block_size the file was encoded with (multi only; 0 = 256). */
int dint_vroom_plan_create(const dint_dict* d, const void* bin, size_t len, uint32_t block_size, uint32_t chunk_ints,
                           dint_vroom_plan** out);
void dint_vroom_plan_destroy(dint_vroom_plan* p);
uint64_t dint_vroom_plan_num_lists(const dint_vroom_plan* p);
uint64_t dint_vroom_plan_num_ints(const dint_vroom_plan* p);
uint64_t dint_vroom_plan_num_segments(const dint_vroom_plan* p);
const dint_segment* dint_vroom_plan_segments(const dint_vroom_plan* p);
/* Work units: unit u = segments [unit_first[u], unit_first[u+1]), consecutive segments with
 * contiguous output holding about 2 K integers; one warp decodes one unit at a time.
 * num_units + 1 entries. */
uint64_t dint_vroom_plan_num_units(const dint_vroom_plan* p);
const uint32_t* dint_vroom_plan_unit_first(const dint_vroom_plan* p);
/* Strip units of the same file (see dint_strip_unit); what dint_decode_strips and the
 * dint_vroom_ctx_* calls consume. */
uint64_t dint_vroom_plan_num_strip_units(const dint_vroom_plan* p);
const dint_strip_unit* dint_vroom_plan_strip_units(const dint_vroom_plan* p);
/* Shape of the file's codeword streams, counted by the pre-pass (diagnostics; no reference counterpart):
 * out[0] 16-bit exceptions, [1] 32-bit exceptions, [2] run codewords, [3..7] dictionary entries of
 * 1, 2, 4, 8, 16 integers, [8] u8 codewords, [9] u16 codewords, [10] integers produced by runs,
 * [11] sum over strip units of the longest strip (codewords), [12] strips. */
void dint_vroom_plan_codeword_stats(const dint_vroom_plan* p, uint64_t out[16]);
/* n_lists + 1 offsets into the packed output; list i is [off[i], off[i+1]) */
const uint64_t* dint_vroom_plan_list_offsets(const dint_vroom_plan* p);
/* n_lists + 1 byte offsets of the list headers; the last one is the file size */
const uint64_t* dint_vroom_plan_list_byte_offsets(const dint_vroom_plan* p);

/* ---- frozen dict_freq_index .bin pre-pass (host, one time per file) -------------------- */
int dint_index_plan_create(int kind, const void* bin, size_t len, uint32_t block_size, dint_index_plan** out);
void dint_index_plan_destroy(dint_index_plan* p);
uint64_t dint_index_plan_num_lists(const dint_index_plan* p);
uint64_t dint_index_plan_num_docs(const dint_index_plan* p);
uint64_t dint_index_plan_num_postings(const dint_index_plan* p);
uint64_t dint_index_plan_num_blocks(const dint_index_plan* p);
const dint_block* dint_index_plan_blocks(const dint_index_plan* p);
const uint64_t* dint_index_plan_list_offsets(const dint_index_plan* p); /* n_lists + 1 */
const uint64_t* dint_index_plan_endpoints(const dint_index_plan* p);    /* n_lists + 1 byte offsets in m_lists */
/* byte ranges inside the .bin */
int dint_index_plan_layout(const dint_index_plan* p, uint64_t* lists_off, uint64_t* lists_len, uint64_t* docs_dict_off,
                           uint64_t* freqs_dict_off, uint64_t* end_off);

/* ---- device-side discovery of the restart points (no host codeword walk) ------------------ */
/* What dint_vroom_plan_create learns by walking every codeword on one host thread is found here by the
 * GPU from the bytes already in HBM (d_in: DEVICE pointer to the .bin image). The file is cut into CHAINS
 * at list starts; one warp walks one chain exactly like vroom_env/decode.cpp:139-150 does (header::read,
 * then codewords until n integers are out: vroom_env/dint_codecs.hpp:45-101, 530-611), 32 symbols per step,
 * without copying dictionary entries; a count pass sizes the tables and an emit pass fills them.
 * Chain starts = byte 0 + sync_offsets[0..n_sync) (HOST array of byte offsets the caller knows to be list
 * starts, e.g. where shard files were concatenated; may be NULL) + list starts a speculative search finds
 * every sync_stride bytes (0 = library default, ~0 = no search): a position is taken when it parses as
 * three consecutive lists whose values sum to the `universe` field of their headers (vroom_env/jobs.hpp:76-85).
 * Every chain must END exactly on the next chain's start; a start that is not landed on is dropped and the
 * count pass repeated, so neither a wrong hint nor a wrong guess can change the result.
 * Synchronises `stream` (this is a load-time call). */
int dint_vroom_discover(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, uint32_t block_size,
                        const uint64_t* sync_offsets, size_t n_sync, uint64_t sync_stride, void* stream,
                        dint_device_plan** out);
/* The same for the full blocks of a frozen dict_freq_index: block starts come from the file's own endpoint
 * tables (d_blocks: DEVICE dint_block table, 16-byte aligned, e.g. dint_index_plan_blocks uploaded), the docs
 * stream of a block is walked first, the freqs stream starts where it ends (dict_posting_list.hpp:311-318).
 * Two strip tables come back: docs (DINT_MODE_DOCIDS, block bases folded in) and freqs. */
int dint_index_discover(const dint_dict* docs_dict, const dint_dict* freqs_dict, const uint8_t* d_lists, size_t lists_bytes,
                        const dint_block* d_blocks, size_t n_blocks, uint32_t block_size, void* stream,
                        dint_device_plan** docs_plan, dint_device_plan** freqs_plan);
/* Device-side twin of dint_index_plan_create for an index file that is already in HBM (d_bin: DEVICE pointer to the
 * whole .bin image on `device`; h_bin: the same bytes on the host, of which ONLY the fixed-size header and the size
 * fields of the embedded vectors are read — nothing that grows with the collection). The GPU expands the
 * Elias-Fano list endpoints (compact_elias_fano.hpp:14-61), parses every list header (dict_posting_list.hpp:90-107)
 * and writes the dint_block table + list offsets; they stay in HBM (accessors below return DEVICE pointers).
 * Synchronises `stream`. */
int dint_index_device_plan_create(int kind, const void* h_bin, size_t len, const uint8_t* d_bin, int device, uint32_t block_size,
                                  void* stream, dint_device_index** out);
void dint_device_index_destroy(dint_device_index* p);
uint64_t dint_device_index_num_lists(const dint_device_index* p);
uint64_t dint_device_index_num_docs(const dint_device_index* p);
uint64_t dint_device_index_num_postings(const dint_device_index* p);
uint64_t dint_device_index_num_blocks(const dint_device_index* p);
double dint_device_index_build_ms(const dint_device_index* p);               /* device timeline, CUDA events */
const dint_block* dint_device_index_blocks(const dint_device_index* p);      /* DEVICE, num_blocks entries */
const uint64_t* dint_device_index_list_offsets(const dint_device_index* p);  /* DEVICE, num_lists + 1 */
const uint64_t* dint_device_index_endpoints(const dint_device_index* p);     /* DEVICE, num_lists + 1 byte offsets in m_lists */
int dint_device_index_layout(const dint_device_index* p, uint64_t* lists_off, uint64_t* lists_len, uint64_t* docs_dict_off,
                             uint64_t* freqs_dict_off, uint64_t* end_off);
void dint_device_plan_destroy(dint_device_plan* p);
uint64_t dint_device_plan_num_lists(const dint_device_plan* p);
uint64_t dint_device_plan_num_ints(const dint_device_plan* p);
uint64_t dint_device_plan_num_units(const dint_device_plan* p);
uint64_t dint_device_plan_num_chains(const dint_device_plan* p);
uint64_t dint_device_plan_num_sync_found(const dint_device_plan* p); /* list starts the speculative search contributed */
uint64_t dint_device_plan_num_skimmed(const dint_device_plan* p);    /* long chains (no list start found inside) cut into pieces at block boundaries */
uint32_t dint_device_plan_retries(const dint_device_plan* p);        /* count passes repeated after a wrong sync point */
double dint_device_plan_discover_ms(const dint_device_plan* p);      /* device timeline of the discovery, CUDA events */
double dint_device_plan_phase_ms(const dint_device_plan* p, int phase); /* 0 sync search, 1 count pass(es), 2 emit pass */
/* DEVICE pointers owned by the plan: num_units strip units (input of dint_decode_strips); num_lists + 1
 * offsets into the packed output / into the file (vroom plans only, NULL for index plans) */
const dint_strip_unit* dint_device_plan_units(const dint_device_plan* p);
const uint64_t* dint_device_plan_list_offsets(const dint_device_plan* p);
const uint64_t* dint_device_plan_list_byte_offsets(const dint_device_plan* p);

/* ---- device decode (stream ordered; all pointers are DEVICE pointers) ------------------ */
/* d_in: the .bin bytes as they are on disk (in_bytes of them; no padding required).
 * d_segs: n_segs descriptors. d_out: packed output, at least max(out_off + n) ints.
 * stream: a cudaStream_t (0 = default stream). block_size: the segments of a multi_packed list already are
 * the blocks the plan cut with this value, so it is only checked for sanity (<= 65536). */
int dint_decode_segments(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_segment* d_segs,
                         size_t n_segs, uint32_t block_size, uint32_t* d_out, int mode, void* stream);
/* Same with an explicit work-unit table (d_unit_first: n_units + 1 DEVICE uint32, see
 * dint_vroom_plan_unit_first). dint_decode_segments groups every 8 consecutive segments. */
int dint_decode_units(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_segment* d_segs, size_t n_segs,
                      const uint32_t* d_unit_first, size_t n_units, uint32_t* d_out, int mode, void* stream);

/* Lane-sequential decode of a strip table (the fast path for whole vroom files): one warp per
 * strip unit, one thread per strip; the unit's bytes are staged in shared memory with 16-byte
 * cp.async, every thread runs the reference's scalar codeword loop over its strip into a
 * shared-memory image of the unit's output, and the warp writes the image with the fused
 * transform of `mode` in 16-byte aligned, fully coalesced stores. d_units: n_units DEVICE
 * dint_strip_unit records. */
int dint_decode_strips(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_strip_unit* d_units,
                       size_t n_units, uint32_t* d_out, int mode, void* stream);

/* d_lists: the m_lists bytes of the index. Writes docids and freqs (enumerator
 * semantics: docid(), freq()) of every block into the packed arrays. Either output
 * may be NULL to skip it (docs blocks are still parsed to locate the freqs bytes). */
int dint_decode_index(const dint_dict* docs_dict, const dint_dict* freqs_dict, const uint8_t* d_lists,
                      size_t lists_bytes, const dint_block* d_blocks, size_t n_blocks, uint32_t block_size,
                      uint32_t* d_docids, uint32_t* d_freqs, void* stream);

/* The same result from the strip tables dint_index_discover built: full blocks go through the lane-sequential
 * decoder (dint_decode_strips: docs with the block bases folded into the docID prefix sum, freqs with +1),
 * partial blocks through the interpolative kernel. d_blocks is still needed for the partial blocks. */
int dint_decode_index_planned(const dint_dict* docs_dict, const dint_dict* freqs_dict, const uint8_t* d_lists,
                              size_t lists_bytes, const dint_device_plan* docs_plan, const dint_device_plan* freqs_plan,
                              const dint_block* d_blocks, size_t n_blocks, uint32_t block_size, uint32_t* d_docids,
                              uint32_t* d_freqs, void* stream);

/* The dictionary-less block codecs behind the same block table ("the other block codecs behind the same
 * batch-decode ABI"): kind DINT_INDEX_BLOCK_INTERPOLATIVE / _VBYTE / _U32 selects interpolative_block, vbyte_block or
 * u32_block ::decode (include/ds2i/block_codecs.hpp:104-150, 357-375, 377-402) for EVERY block of the table — full
 * ones included — and the enumerator's docid() / freq() values are written (block_posting_list.hpp:113-124,
 * 164-169, 195-208). d_blocks from a plan made with the same kind. Any other kind: DINT_E_ARG. */
int dint_decode_index_blocks(int kind, const uint8_t* d_lists, size_t lists_bytes, const dint_block* d_blocks, size_t n_blocks,
                             uint32_t block_size, uint32_t* d_docids, uint32_t* d_freqs, void* stream);

/* dint_decode_index_blocks(DINT_INDEX_BLOCK_INTERPOLATIVE, ...) */
int dint_decode_index_interpolative(const uint8_t* d_lists, size_t lists_bytes, const dint_block* d_blocks, size_t n_blocks,
                                    uint32_t block_size, uint32_t* d_docids, uint32_t* d_freqs, void* stream);

/* sums[i] = sum_{j in list i} (value_j + 1) * (j + 1)  mod 2^64 */
int dint_checksum_lists(const uint32_t* d_values, const uint64_t* d_list_offsets, size_t n_lists, uint64_t* d_sums,
                        void* stream);

/* ---- a consumer on the device: batched AND queries over the decoded docIDs ------------------ */
/* ds2i::and_query<false> (include/ds2i/queries.hpp:34-86; caller: src/queries.cpp:15-61) for a batch of queries over
 * docID lists that are already decoded in HBM (DINT_MODE_DOCIDS output of any decode call above + its list
 * offsets): counts[q] = number of documents that contain every term of query q. Queries in CSR form:
 * d_query_terms[d_query_offsets[q] .. d_query_offsets[q + 1]) are term ids (= list indices; duplicates are dropped
 * like remove_duplicate_terms does; at most 32 terms per query, else counts[q] = UINT64_MAX; a term id >= n_lists
 * gives 0). One CTA per query: each warp looks 32 docIDs of the shortest list up in the other lists at a time. */
int dint_and_queries(const uint32_t* d_docids, const uint64_t* d_list_offsets, size_t n_lists, const uint64_t* d_query_offsets,
                     const uint32_t* d_query_terms, size_t n_queries, uint64_t* d_counts, void* stream);

/* ---- host-buffer convenience: one Decoder::decode(dict, in, out, universe, n) call ------ */
/* in/out are HOST pointers; copies in, decodes on the device, copies out; *consumed is
 * the reference's returned pointer minus `in`. out receives exactly n integers (RAW). */
int dint_decode_list_host(const dint_dict* d, const uint8_t* in, size_t in_avail, uint32_t* out, uint32_t n,
                          uint32_t block_size, size_t* consumed);

/* ---- host-buffer batch decode of a whole vroom .bin --------------------------------------- */
/* A ctx owns the device workspace for files that share `plan`'s layout (device copy of the
 * segment table and list offsets, input / output / checksum buffers): create once, reuse for
 * every call, so the call itself allocates nothing. */
int dint_vroom_ctx_create(const dint_dict* d, const dint_vroom_plan* plan, uint32_t block_size, dint_vroom_ctx** out);
/* The same workspace for a file NO host plan exists for (the cold path): copies h_bin (HOST pointer, len bytes)
 * to the device on `stream`, finds its restart points there (dint_vroom_discover; sync_* as documented
 * there) and sizes the output from what the walk counted. Synchronises `stream`. */
int dint_vroom_ctx_create_discover(const dint_dict* d, const void* h_bin, size_t len, uint32_t block_size,
                                   const uint64_t* sync_offsets, size_t n_sync, uint64_t sync_stride, void* stream,
                                   dint_vroom_ctx** out);
void dint_vroom_ctx_destroy(dint_vroom_ctx* c);
uint64_t dint_vroom_ctx_num_lists(const dint_vroom_ctx* c);
uint64_t dint_vroom_ctx_num_ints(const dint_vroom_ctx* c);
/* the device plan of a ctx made by dint_vroom_ctx_create_discover (NULL otherwise); owned by the ctx */
const dint_device_plan* dint_vroom_ctx_device_plan(const dint_vroom_ctx* c);
/* device pointer of the ctx's packed output (num_ints values, valid after a decode) */
uint32_t* dint_vroom_ctx_device_output(const dint_vroom_ctx* c);
/* h_bin: HOST pointer to the .bin image the plan was made from (pinned memory makes the
 * copies asynchronous). Enqueues on `stream`: H2D of the bytes, the decode kernel in `mode`,
 * then (h_sums != NULL) per-list checksums + D2H of num_lists uint64, and (h_out != NULL) D2H
 * of all num_ints decoded integers. Returns without synchronising the stream. */
int dint_vroom_decode_host(dint_vroom_ctx* c, const void* h_bin, size_t len, int mode, uint32_t* h_out,
                           uint64_t* h_sums, void* stream);
/* Decodes again from the bytes the last dint_vroom_decode_host call left in HBM (no copies):
 * what the reference's decode loop does once the file is mapped. */
int dint_vroom_decode_resident(dint_vroom_ctx* c, int mode, void* stream);
/* Enqueues what dint_vroom_decode_host does after the decode: (h_sums != NULL) per-list checksums + D2H of
 * num_lists uint64, (h_out != NULL) D2H of all decoded integers of the ctx's last decode. */
int dint_vroom_ctx_results(dint_vroom_ctx* c, uint32_t* h_out, uint64_t* h_sums, void* stream);
/* Blocks until everything enqueued on `stream` of the ctx's device has finished. */
int dint_vroom_ctx_sync(dint_vroom_ctx* c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DINT_B200_H */
