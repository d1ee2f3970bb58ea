// dint_b200 — sm_100a kernels and the device half of the C ABI.
//
// Work unit: a "segment" = contiguous codeword stream starting on a codeword boundary
// and producing exactly n integers (include/dint_b200.h). One warp decodes one segment,
// 32 codeword symbols (u16 or u8) per step:
//
//   1. lanes load one symbol each (coalesced; byte-wise when the stream is odd-aligned);
//   2. __ballot_sync masks of the symbols equal to 0 / 1 resolve which symbols are
//      codewords and which are exception payload (a 0 is followed by 1 payload symbol,
//      a 1 by 2; payload may itself be 0/1, so the masks are walked in stream order —
//      a warp-uniform loop with one trip per 0/1 symbol, i.e. ~0.25 % of symbols);
//   3. entry lengths come from the (len-1) byte table staged in shared memory;
//   4. a __shfl_up_sync inclusive scan turns lengths into output offsets; codewords past
//      the segment's n-th integer (bytes of the next list) are dropped;
//   5. the output is produced 32 integers at a time in OUTPUT order: a __reduce_or_sync
//      "head flag" mask + popc tells every lane which codeword covers its position, the
//      value is one 4-byte load from the 64-byte-aligned dictionary row (L1/L2 resident),
//      runs produce zeros without touching memory, exceptions carry their value;
//   6. DOCIDS mode fuses the d-gap prefix sum (second warp scan + running carry), FREQS
//      mode the +1 shift; stores are fully coalesced 128-byte lines.
//
// Reference semantics reproduced (paths relative to the reference root):
//   vroom_env/dint_codecs.hpp:37-107 (u16 loop), :574-609 (u8 loop), :521-536 (selector),
//   include/dint/dint_codecs.hpp:13-49, 460-510, include/ds2i/block_codecs.hpp:87-101,
//   130-150, include/ds2i/interpolative_coding.hpp:79-153,
//   include/dint/dict_posting_list.hpp:284-318 (block bases), :113-124, 164-169.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>

#include "dint_internal.hpp"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarpsPerCta = 16;
constexpr int kThreads = kWarpsPerCta * 32;

enum { KIND_DICT = 0, KIND_RUN = 1, KIND_EXC = 2 };
enum { OUT_RAW = 0, OUT_DOCIDS = 1, OUT_FREQS = 2, OUT_SKIP = 3 };

std::atomic<uint64_t> g_launches{0};

#define CUDA_TRY(expr)                                                                    \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            dint::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));         \
            return DINT_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

__device__ __forceinline__ unsigned lane_id() {
    unsigned l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned lanemask_le() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// streaming loads of the compressed bytes: read once, do not pollute L1
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u16(const uint8_t* p) {
    uint16_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}

// symbol `lane` of the 32-symbol tile that starts at byte `pos`
template <int SYM>
__device__ __forceinline__ uint32_t load_symbol(const uint8_t* __restrict__ in, uint64_t in_bytes, uint64_t pos,
                                                unsigned lane) {
    uint64_t a = pos + uint64_t(SYM) * lane;
    if (SYM == 1) return a < in_bytes ? ldg_u8(in + a) : 0u;
    if (a + 2 > in_bytes) return 0u;
    if (((reinterpret_cast<uintptr_t>(in) + pos) & 1) == 0) return ldg_u16(in + a);  // warp-uniform
    return ldg_u8(in + a) | (ldg_u8(in + a + 1) << 8);
}

// symbol (lane + d) of the concatenation cur | nxt, 1 <= d <= 4
__device__ __forceinline__ uint32_t peek(uint32_t cur, uint32_t nxt, unsigned lane, unsigned d) {
    uint32_t a = __shfl_sync(FULL, cur, (lane + d) & 31);
    uint32_t b = __shfl_sync(FULL, nxt, (lane + d) & 31);
    return (lane + d < 32) ? a : b;
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, unsigned lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, o);
        if (lane >= unsigned(o)) v += t;
    }
    return v;
}

// Decodes one segment with the calling warp. Returns the byte offset one past the last
// consumed symbol (the reference's returned pointer). SYM = bytes per codeword symbol.
//   lut       shared-memory (len-1) table of the segment's class
//   row_base  first dictionary row of the class, class_rows its number of rows
//   out       first output integer of the segment (ignored for OUT_SKIP)
template <int SYM, int MODE>
__device__ __forceinline__ uint64_t decode_stream(const uint32_t* __restrict__ rows, const uint8_t* __restrict__ lut,
                                                  uint32_t row_base, uint32_t class_rows,
                                                  const uint8_t* __restrict__ in, uint64_t in_bytes, uint64_t pos,
                                                  uint32_t n, uint32_t* __restrict__ out, uint32_t doc_base,
                                                  uint2* __restrict__ cwbuf, unsigned lane) {
    constexpr int PZ = SYM == 2 ? 1 : 2;  // payload symbols after a 0 codeword (u16 value)
    constexpr int PO = SYM == 2 ? 2 : 4;  // payload symbols after a 1 codeword (u32 value)
    uint32_t remaining = n;
    uint32_t produced = 0;
    uint32_t acc = doc_base - 1u;  // docid[i] = acc + sum_{j<=i} (v_j + 1)
    uint32_t carry = 0;            // payload symbols spilling in from the previous tile (low bits)
    uint64_t tile_pos = pos;
    uint64_t end_off = pos;
    uint32_t cur = load_symbol<SYM>(in, in_bytes, tile_pos, lane);
    while (remaining) {
        uint32_t nxt = load_symbol<SYM>(in, in_bytes, tile_pos + 32 * SYM, lane);  // prefetch

        // -- which symbols are codewords ---------------------------------------------------
        const unsigned Z = __ballot_sync(FULL, cur == 0u);
        const unsigned O = __ballot_sync(FULL, cur == 1u);
        uint64_t P = carry;
        uint32_t excval = 0;
        if (Z | O) {
            unsigned m = Z | O;
            while (m) {
                int j = __ffs(m) - 1;
                m &= m - 1;
                if (!((P >> j) & 1ull)) {
                    int cnt = ((Z >> j) & 1u) ? PZ : PO;
                    P |= ((1ull << cnt) - 1ull) << (j + 1);
                }
            }
            uint32_t n1 = peek(cur, nxt, lane, 1), n2 = peek(cur, nxt, lane, 2);
            if (SYM == 2) {
                excval = (cur == 0u) ? n1 : (n1 | (n2 << 16));
            } else {
                uint32_t n3 = peek(cur, nxt, lane, 3), n4 = peek(cur, nxt, lane, 4);
                excval = (cur == 0u) ? (n1 | (n2 << 8)) : (n1 | (n2 << 8) | (n3 << 16) | (n4 << 24));
            }
        }
        carry = uint32_t(P >> 32);
        const bool iscw = !((uint32_t(P) >> lane) & 1u);

        // -- lengths and offsets -----------------------------------------------------------
        uint32_t len = 0, kind = KIND_EXC, data = excval;
        if (iscw) {
            if (cur >= 2u) {
                uint32_t idx = min(cur, class_rows - 1u);
                len = uint32_t(lut[idx]) + 1u;
                data = row_base + idx;
                kind = cur < dint::kReserved ? KIND_RUN : KIND_DICT;
            } else {
                len = 1;
            }
        }
        const uint32_t e = warp_inclusive_scan(len, lane);
        const uint32_t s = e - len;
        const bool valid = iscw && (s < remaining);
        const unsigned V = __ballot_sync(FULL, valid);
        const uint32_t T = min(__shfl_sync(FULL, e, 31), remaining);

        if (MODE != OUT_SKIP) {
            if (valid) cwbuf[__popc(V & lanemask_lt())] = make_uint2(data, s | (kind << 16));
            __syncwarp();
            // -- expansion, 32 outputs per round, in output order ---------------------------
            uint32_t kbase = 0;
            for (uint32_t w0 = 0; w0 < T; w0 += 32) {
                const uint32_t rel = s - w0;
                const unsigned heads = __reduce_or_sync(FULL, (valid && rel < 32u) ? (1u << rel) : 0u);
                const uint32_t p = w0 + lane;
                const uint32_t k = kbase + __popc(heads & lanemask_le()) - 1u;
                kbase += __popc(heads);
                uint32_t v = 0;
                const bool act = p < T;
                if (act) {
                    const uint2 info = cwbuf[k];
                    const uint32_t kd = info.y >> 16;
                    if (kd == KIND_DICT) v = __ldg(rows + (size_t(info.x) * dint::kEntry + (p - (info.y & 0xffffu))));
                    else if (kd == KIND_EXC) v = info.x;
                }
                if (MODE == OUT_DOCIDS) {
                    const uint32_t xs = warp_inclusive_scan(act ? v + 1u : 0u, lane);
                    v = acc + xs;
                    acc += __shfl_sync(FULL, xs, 31);
                } else if (MODE == OUT_FREQS) {
                    v += 1u;
                }
                if (act) out[produced + p] = v;
            }
            __syncwarp();
        }
        produced += T;
        remaining -= T;
        if (remaining == 0) {
            // the last codeword of the segment and its payload are the last consumed symbols
            const int kl = 31 - __clz(V);
            const uint32_t lastsym = __shfl_sync(FULL, cur, kl);
            const uint32_t extra = lastsym == 0u ? PZ : (lastsym == 1u ? PO : 0);
            end_off = tile_pos + uint64_t(SYM) * (kl + 1 + extra);
        }
        tile_pos += 32 * SYM;
        cur = nxt;
    }
    return end_off;
}

// single kinds: u16 stream at `pos`. multi: selector byte at `pos`, then u16 (sel < 6) or
// u8 (sel - 6) stream into class sel % 6 (vroom_env/dint_codecs.hpp:536-575).
template <int MODE>
__device__ __forceinline__ uint64_t decode_segment(const dint_dict_view& dv, const uint8_t* __restrict__ lut,
                                                   bool multi, const uint8_t* __restrict__ in, uint64_t in_bytes,
                                                   uint64_t pos, uint32_t n, uint32_t* out, uint32_t doc_base,
                                                   uint2* cwbuf, unsigned lane) {
    const uint32_t* rows = dv.rows;
    if (!multi) return decode_stream<2, MODE>(rows, lut, 0, dv.n_rows, in, in_bytes, pos, n, out, doc_base, cwbuf, lane);
    uint32_t sel = pos < in_bytes ? ldg_u8(in + pos) : 0u;
    const bool wide = sel < dint::kSelectors;
    uint32_t cls = wide ? sel : sel - dint::kSelectors;
    cls = min(cls, dv.n_classes - 1u);
    const uint32_t base = dv.class_base[cls];
    const uint32_t crows = (cls + 1 < dv.n_classes ? dv.class_base[cls + 1] : dv.n_rows) - base;
    if (wide) return decode_stream<2, MODE>(rows, lut + base, base, crows, in, in_bytes, pos + 1, n, out, doc_base, cwbuf, lane);
    return decode_stream<1, MODE>(rows, lut + base, base, min(crows, 256u), in, in_bytes, pos + 1, n, out, doc_base, cwbuf, lane);
}

extern __shared__ __align__(16) uint8_t g_smem[];

// stage the (len-1) table into shared memory; returns the per-warp codeword buffer
__device__ __forceinline__ uint2* stage_lut(const dint_dict_view& dv, uint32_t lut_bytes) {
    const uint4* src = reinterpret_cast<const uint4*>(dv.lens);
    uint4* dst = reinterpret_cast<uint4*>(g_smem);
    for (uint32_t i = threadIdx.x; i < lut_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    return reinterpret_cast<uint2*>(g_smem + lut_bytes) + (threadIdx.x >> 5) * 32;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
decode_segments_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ in,
                       uint64_t in_bytes, const dint_segment* __restrict__ segs, uint64_t n_segs,
                       uint32_t* __restrict__ out) {
    uint2* cwbuf = stage_lut(dv, lut_bytes);
    const unsigned lane = lane_id();
    const uint64_t gwarp = uint64_t(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t nwarps = uint64_t(gridDim.x) * kWarpsPerCta;
    for (uint64_t i = gwarp; i < n_segs; i += nwarps) {
        const dint_segment sg = segs[i];
        decode_segment<MODE>(dv, g_smem, multi != 0, in, in_bytes, sg.in_off, sg.n, out + sg.out_off, sg.doc_base, cwbuf,
                             lane);
    }
}

// Full blocks of a dict_freq_index, pass 1 (docs): decode_docs_block (dict_posting_list.hpp:284-309).
// The block's docs byte length (= where its freqs bytes start) is left in the first slot of
// the block's freqs output for pass 2.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
index_docs_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ lists,
                  uint64_t lists_bytes, const dint_block* __restrict__ blocks, uint64_t n_blocks, uint32_t block_size,
                  uint32_t* __restrict__ docids, uint32_t* __restrict__ freqs) {
    uint2* cwbuf = stage_lut(dv, lut_bytes);
    const unsigned lane = lane_id();
    const uint64_t gwarp = uint64_t(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t nwarps = uint64_t(gridDim.x) * kWarpsPerCta;
    for (uint64_t i = gwarp; i < n_blocks; i += nwarps) {
        const dint_block b = blocks[i];
        if (b.size != block_size) continue;  // partial block: interpolative, index_tails_kernel
        // docid[0] = buf[0] + cur_base (dict_posting_list.hpp:304)
        uint64_t end = decode_segment<MODE>(dv, g_smem, multi != 0, lists, lists_bytes, b.in_off, b.size,
                                            MODE == OUT_SKIP ? nullptr : docids + b.out_off, b.cur_base, cwbuf, lane);
        if (freqs && lane == 0) freqs[b.out_off] = uint32_t(end - b.in_off);
    }
}

// pass 2 (freqs): decode_freqs_block (dict_posting_list.hpp:311-318), freq() = buf[i] + 1 (:168)
__global__ void __launch_bounds__(kThreads)
index_freqs_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ lists,
                   uint64_t lists_bytes, const dint_block* __restrict__ blocks, uint64_t n_blocks, uint32_t block_size,
                   uint32_t* __restrict__ freqs) {
    uint2* cwbuf = stage_lut(dv, lut_bytes);
    const unsigned lane = lane_id();
    const uint64_t gwarp = uint64_t(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
    const uint64_t nwarps = uint64_t(gridDim.x) * kWarpsPerCta;
    for (uint64_t i = gwarp; i < n_blocks; i += nwarps) {
        const dint_block b = blocks[i];
        if (b.size != block_size) continue;
        const uint32_t docs_len = freqs[b.out_off];
        __syncwarp();
        decode_segment<OUT_FREQS>(dv, g_smem, multi != 0, lists, lists_bytes, b.in_off + docs_len, b.size,
                                  freqs + b.out_off, 0, cwbuf, lane);
    }
}

// ---- partial blocks: binary interpolative code, one thread per block ----------------------
// interpolative_block::decode (block_codecs.hpp:130-150) + bit_reader (interpolative_coding.hpp:79-153)
struct bit_reader {
    // Aligned 4-byte loads only touch words that contain at least one byte of the stream,
    // so they never leave the (>= 4-byte granular) allocation that holds it.
    const uint32_t* wp;   // next aligned word to fetch
    const uint8_t* endp;  // one past the last valid byte
    uint32_t shift;       // (start address & 3) * 8
    uint32_t prev;        // previously fetched aligned word
    uint64_t buf;
    uint32_t avail;
    uint64_t pos;  // bits consumed
    __device__ uint32_t aligned() {
        uint32_t v = reinterpret_cast<const uint8_t*>(wp) < endp ? __ldg(wp) : 0u;
        ++wp;
        return v;
    }
    __device__ void init(const uint8_t* base, uint64_t nbytes, uint64_t start) {
        uintptr_t a = reinterpret_cast<uintptr_t>(base) + start;
        wp = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
        endp = base + nbytes;
        shift = uint32_t(a & 3) * 8;
        prev = aligned();
        buf = 0;
        avail = 0;
        pos = 0;
    }
    __device__ uint32_t next32() {  // the u32 at byte start + 4k, as the reference's *m_in++
        if (shift == 0) {
            uint32_t v = prev;
            prev = aligned();
            return v;
        }
        uint32_t nw = aligned();
        uint32_t v = __funnelshift_r(prev, nw, shift);
        prev = nw;
        return v;
    }
    __device__ uint32_t read(uint32_t len) {  // interpolative_coding.hpp:93-107
        if (!len) return 0;
        if (avail < len) {
            buf |= uint64_t(next32()) << avail;
            avail += 32;
        }
        uint32_t val = uint32_t(buf & ((1ull << len) - 1ull));
        buf >>= len;
        avail -= len;
        pos += len;
        return val;
    }
    __device__ uint32_t read_int(uint32_t u) {  // :109-122, truncated binary code
        uint32_t b = 31u - __clz(u);
        uint64_t m = (1ull << (b + 1)) - u;
        uint32_t val = read(b);
        if (val >= m) val = uint32_t((uint64_t(val) << 1) + read(1) - m);
        return val;
    }
};

// pre-order walk of read_interpolative (:124-146) with an explicit stack; emit(idx, val)
template <typename Emit>
__device__ void read_interpolative(bit_reader& br, uint32_t n, uint32_t low0, uint32_t high0, Emit&& emit) {
    uint32_t st_off[10], st_n[10], st_lo[10], st_hi[10];
    int sp = 0;
    st_off[0] = 0; st_n[0] = n; st_lo[0] = low0; st_hi[0] = high0;
    sp = 1;
    while (sp) {
        --sp;
        uint32_t off = st_off[sp], cnt = st_n[sp], lo = st_lo[sp], hi = st_hi[sp];
        for (;;) {
            uint32_t h = cnt / 2;
            uint32_t val = lo + br.read_int(hi - lo + 1u);
            emit(off + h, val);
            if (cnt == 1) break;
            uint32_t rn = cnt - h - 1;
            if (rn) {  // right half later
                st_off[sp] = off + h + 1; st_n[sp] = rn; st_lo[sp] = val; st_hi[sp] = hi;
                ++sp;
            }
            if (!h) break;
            cnt = h;  // left half now
            hi = val;
        }
    }
}

__device__ __forceinline__ uint32_t vbyte_at(const uint8_t* base, uint64_t nbytes, uint64_t& p) {
    uint32_t v = 0;
    for (unsigned shift = 0; p < nbytes; shift += 7) {
        uint32_t c = base[p++];
        v += (c & 127u) << (shift & 31u);
        if (c & 128u) break;
    }
    return v;
}

__global__ void __launch_bounds__(128)
index_tails_kernel(const uint8_t* __restrict__ lists, uint64_t lists_bytes, const dint_block* __restrict__ blocks,
                   uint64_t n_blocks, uint32_t block_size, uint32_t* __restrict__ docids, uint32_t* __restrict__ freqs) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_blocks; i += uint64_t(gridDim.x) * blockDim.x) {
        const dint_block b = blocks[i];
        if (b.size >= block_size) continue;
        const uint32_t n = b.size;
        uint64_t p = b.in_off;
        // docs: prefix sums P_i of the first n-1 values in [0, S], P_{n-1} = S
        // (block_codecs.hpp:139-146); docid_i = cur_base + P_i + i (dict_posting_list.hpp:304-307,122)
        const uint32_t S = b.block_max - b.cur_base - (n - 1u);
        uint32_t* dd = docids ? docids + b.out_off : nullptr;
        if (dd) dd[n - 1] = b.block_max;
        if (n > 1) {
            bit_reader br;
            br.init(lists, lists_bytes, p);
            const uint32_t cb = b.cur_base;
            read_interpolative(br, n - 1, 0, S, [&](uint32_t idx, uint32_t val) {
                if (dd) dd[idx] = cb + val + idx;
            });
            p += (br.pos + 7) / 8;
        }
        if (!freqs) continue;
        // freqs: sum_of_values == -1 -> leading vbyte (block_codecs.hpp:134-136)
        uint32_t* ff = freqs + b.out_off;
        const uint32_t Sf = vbyte_at(lists, lists_bytes, p);
        ff[n - 1] = Sf;
        if (n > 1) {
            bit_reader br;
            br.init(lists, lists_bytes, p);
            read_interpolative(br, n - 1, 0, Sf, [&](uint32_t idx, uint32_t val) { ff[idx] = val; });
            for (uint32_t k = n - 1; k > 0; --k) ff[k] = ff[k] - ff[k - 1] + 1u;  // adjacent differences, freq = v + 1
        }
        ff[0] += 1u;
    }
}

// ---- per-list checksums ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
checksum_kernel(const uint32_t* __restrict__ values, const uint64_t* __restrict__ offs, uint64_t n_lists,
                uint64_t* __restrict__ sums) {
    const unsigned lane = lane_id();
    const uint64_t gwarp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t i = gwarp; i < n_lists; i += nwarps) {
        const uint64_t b = offs[i], e = offs[i + 1];
        uint64_t acc = 0;
        for (uint64_t j = b + lane; j < e; j += 32) acc += (uint64_t(values[j]) + 1ull) * (j - b + 1ull);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) sums[i] = acc;
    }
}

// ---- single-list walk for the host-buffer convenience call ----------------------------------
__global__ void __launch_bounds__(32)
decode_list_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ in, uint64_t in_bytes,
                   uint32_t n, uint32_t block_size, uint32_t* __restrict__ out, uint64_t* __restrict__ consumed) {
    uint2* cwbuf = stage_lut(dv, lut_bytes);
    const unsigned lane = lane_id();
    uint64_t pos = 0;
    if (!multi) {
        pos = decode_segment<OUT_RAW>(dv, g_smem, false, in, in_bytes, 0, n, out, 0, cwbuf, lane);
    } else {
        for (uint32_t done = 0; done < n; done += block_size)  // vroom_env/dint_codecs.hpp:525-534
            pos = decode_segment<OUT_RAW>(dv, g_smem, true, in, in_bytes, pos, min(block_size, n - done), out + done, 0,
                                          cwbuf, lane);
    }
    if (lane == 0) *consumed = pos;
}

// ---- launch helpers -------------------------------------------------------------------------
struct device_info {
    int sms = 0;
    int max_smem_optin = 0;
};

int get_device_info(int device, device_info* di) {
    static std::mutex mu;
    static device_info cache[64];
    static bool have[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return DINT_E_ARG;
    if (!have[device]) {
        CUDA_TRY(cudaDeviceGetAttribute(&cache[device].sms, cudaDevAttrMultiProcessorCount, device));
        CUDA_TRY(cudaDeviceGetAttribute(&cache[device].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        have[device] = true;
    }
    *di = cache[device];
    return DINT_OK;
}

uint32_t lut_bytes_of(const dint_dict* d) { return (d->n_rows + 15u) & ~15u; }
size_t smem_bytes_of(const dint_dict* d, int warps) { return size_t(lut_bytes_of(d)) + size_t(warps) * 32 * sizeof(uint2); }

// persistent grid: as many CTAs as fit (shared memory bound), never more than the work needs
template <typename K>
int plan_grid(K kernel, const dint_dict* d, uint64_t n_units, int* grid, size_t* smem) {
    device_info di;
    int rc = get_device_info(d->device, &di);
    if (rc) return rc;
    *smem = smem_bytes_of(d, kWarpsPerCta);
    if (*smem > size_t(di.max_smem_optin)) {
        dint::set_error("dictionary length table does not fit in shared memory");
        return DINT_E_ARG;
    }
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(*smem)));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, *smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t want = (n_units + kWarpsPerCta - 1) / kWarpsPerCta;
    uint64_t cap = uint64_t(di.sms) * per_sm;
    *grid = int(std::max<uint64_t>(1, std::min(want, cap)));
    return DINT_OK;
}

struct device_guard {
    int prev = -1;
    bool ok = false;
    explicit device_guard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~device_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

// ---- dictionary upload ------------------------------------------------------------------------
int dint_dict_upload(dint_dict* d) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        dint::set_error("no CUDA device: dint_b200 has no CPU fallback");
        return DINT_E_NODEVICE;
    }
    if (d->device >= ndev) return DINT_E_ARG;
    device_guard g(d->device);
    if (!g.ok) {
        dint::set_error("cudaSetDevice failed");
        return DINT_E_CUDA;
    }
    const size_t rows_bytes = size_t(d->n_rows) * dint::kEntry * 4;
    const size_t lens_bytes = lut_bytes_of(d);
    CUDA_TRY(cudaMalloc(&d->d_rows, rows_bytes));
    CUDA_TRY(cudaMalloc(&d->d_lens, lens_bytes));
    CUDA_TRY(cudaMemset(d->d_lens, 0, lens_bytes));
    CUDA_TRY(cudaMemcpy(d->d_rows, d->h_rows.data(), rows_bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d->d_lens, d->h_lens.data(), d->n_rows, cudaMemcpyHostToDevice));
    return DINT_OK;
}

void dint_dict_release_device(dint_dict* d) {
    if (d->d_rows || d->d_lens) {
        device_guard g(d->device);
        cudaFree(d->d_rows);
        cudaFree(d->d_lens);
        d->d_rows = nullptr;
        d->d_lens = nullptr;
    }
}

extern "C" {

uint64_t dint_kernel_launches(void) { return g_launches.load(); }

int dint_dict_copy(const dint_dict* d, uint32_t cls, uint32_t i, uint32_t* out16, uint32_t* len) {
    if (!d || !out16 || !len || cls >= d->n_classes) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident");
        return DINT_E_NODEVICE;
    }
    uint32_t r = d->class_base[cls] + i;
    if (r >= d->n_rows) return DINT_E_ARG;
    device_guard g(d->device);
    uint8_t l = 0;
    CUDA_TRY(cudaMemcpy(out16, d->d_rows + size_t(r) * dint::kEntry, dint::kEntry * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&l, d->d_lens + r, 1, cudaMemcpyDeviceToHost));
    *len = uint32_t(l) + 1;
    return DINT_OK;
}

int dint_decode_segments(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_segment* d_segs,
                         size_t n_segs, uint32_t block_size, uint32_t* d_out, int mode, void* stream) {
    (void)block_size;
    if (!d || mode < DINT_MODE_RAW || mode > DINT_MODE_FREQS) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (n_segs == 0) return DINT_OK;
    if (!d_in || !d_segs || !d_out) return DINT_E_ARG;
    device_guard g(d->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int multi = d->kind == DINT_DICT_MULTI;
    int grid = 0;
    size_t smem = 0;
    int rc;
#define LAUNCH_SEG(M)                                                                                              \
    rc = plan_grid(decode_segments_kernel<M>, d, n_segs, &grid, &smem);                                            \
    if (rc) return rc;                                                                                             \
    decode_segments_kernel<M><<<grid, kThreads, smem, st>>>(d->view(), lut_bytes_of(d), multi, d_in, in_bytes, d_segs, \
                                                           n_segs, d_out);
    if (mode == DINT_MODE_RAW) { LAUNCH_SEG(OUT_RAW) }
    else if (mode == DINT_MODE_DOCIDS) { LAUNCH_SEG(OUT_DOCIDS) }
    else { LAUNCH_SEG(OUT_FREQS) }
#undef LAUNCH_SEG
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

int dint_decode_index(const dint_dict* docs_dict, const dint_dict* freqs_dict, const uint8_t* d_lists,
                      size_t lists_bytes, const dint_block* d_blocks, size_t n_blocks, uint32_t block_size,
                      uint32_t* d_docids, uint32_t* d_freqs, void* stream) {
    if (!docs_dict || !freqs_dict) return DINT_E_ARG;
    if (!docs_dict->d_rows || !freqs_dict->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (docs_dict->device != freqs_dict->device || docs_dict->kind != freqs_dict->kind) return DINT_E_ARG;
    if (n_blocks == 0) return DINT_OK;
    if (!d_lists || !d_blocks || (!d_docids && !d_freqs)) return DINT_E_ARG;
    if (!block_size) block_size = dint::kDefaultBlock;
    device_guard g(docs_dict->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int multi = docs_dict->kind == DINT_DICT_MULTI;
    int grid = 0;
    size_t smem = 0;
    int rc;
    if (d_docids) {
        rc = plan_grid(index_docs_kernel<OUT_DOCIDS>, docs_dict, n_blocks, &grid, &smem);
        if (rc) return rc;
        index_docs_kernel<OUT_DOCIDS><<<grid, kThreads, smem, st>>>(docs_dict->view(), lut_bytes_of(docs_dict), multi,
                                                                   d_lists, lists_bytes, d_blocks, n_blocks, block_size,
                                                                   d_docids, d_freqs);
    } else {
        rc = plan_grid(index_docs_kernel<OUT_SKIP>, docs_dict, n_blocks, &grid, &smem);
        if (rc) return rc;
        index_docs_kernel<OUT_SKIP><<<grid, kThreads, smem, st>>>(docs_dict->view(), lut_bytes_of(docs_dict), multi,
                                                                 d_lists, lists_bytes, d_blocks, n_blocks, block_size,
                                                                 nullptr, d_freqs);
    }
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    if (d_freqs) {
        rc = plan_grid(index_freqs_kernel, freqs_dict, n_blocks, &grid, &smem);
        if (rc) return rc;
        index_freqs_kernel<<<grid, kThreads, smem, st>>>(freqs_dict->view(), lut_bytes_of(freqs_dict), multi, d_lists,
                                                         lists_bytes, d_blocks, n_blocks, block_size, d_freqs);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    {
        device_info di;
        rc = get_device_info(docs_dict->device, &di);
        if (rc) return rc;
        uint64_t want = (n_blocks + 127) / 128;
        int tgrid = int(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(di.sms) * 16)));
        index_tails_kernel<<<tgrid, 128, 0, st>>>(d_lists, lists_bytes, d_blocks, n_blocks, block_size, d_docids, d_freqs);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    return DINT_OK;
}

int dint_checksum_lists(const uint32_t* d_values, const uint64_t* d_list_offsets, size_t n_lists, uint64_t* d_sums,
                        void* stream) {
    if (n_lists == 0) return DINT_OK;
    if (!d_values || !d_list_offsets || !d_sums) return DINT_E_ARG;
    cudaPointerAttributes pa;
    CUDA_TRY(cudaPointerGetAttributes(&pa, d_values));
    int dev = pa.device;
    device_guard g(dev);
    device_info di;
    int rc = get_device_info(dev, &di);
    if (rc) return rc;
    uint64_t want = (n_lists + 7) / 8;
    int grid = int(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(di.sms) * 8)));
    checksum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_values, d_list_offsets, n_lists, d_sums);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

struct dint_vroom_ctx {
    const dint_dict* dict = nullptr;
    uint32_t block_size = 0;
    uint64_t n_lists = 0, n_ints = 0, n_segs = 0, in_bytes = 0;
    uint8_t* d_in = nullptr;
    dint_segment* d_segs = nullptr;
    uint64_t* d_list_offsets = nullptr;
    uint32_t* d_out = nullptr;
    uint64_t* d_sums = nullptr;
};

int dint_vroom_ctx_create(const dint_dict* d, const dint_vroom_plan* plan, uint32_t block_size, dint_vroom_ctx** out) {
    if (!d || !plan || !out) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    device_guard g(d->device);
    dint_vroom_ctx* c = new (std::nothrow) dint_vroom_ctx;
    if (!c) return DINT_E_NOMEM;
    c->dict = d;
    c->block_size = block_size ? block_size : dint::kDefaultBlock;
    c->n_lists = plan->list_offsets.size() - 1;
    c->n_ints = plan->n_ints;
    c->n_segs = plan->segments.size();
    c->in_bytes = plan->list_byte_offsets.back();
    auto fail = [&](cudaError_t e) {
        dint::set_error(std::string("dint_vroom_ctx_create: ") + cudaGetErrorString(e));
        dint_vroom_ctx_destroy(c);
        return DINT_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_in, c->in_bytes + 16)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_segs, (c->n_segs + 1) * sizeof(dint_segment))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_list_offsets, (c->n_lists + 1) * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_out, (c->n_ints + 1) * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_sums, (c->n_lists + 1) * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpy(c->d_segs, plan->segments.data(), c->n_segs * sizeof(dint_segment), cudaMemcpyHostToDevice)) !=
        cudaSuccess)
        return fail(e);
    if ((e = cudaMemcpy(c->d_list_offsets, plan->list_offsets.data(), (c->n_lists + 1) * 8, cudaMemcpyHostToDevice)) !=
        cudaSuccess)
        return fail(e);
    *out = c;
    return DINT_OK;
}

void dint_vroom_ctx_destroy(dint_vroom_ctx* c) {
    if (!c) return;
    device_guard g(c->dict->device);
    cudaFree(c->d_in);
    cudaFree(c->d_segs);
    cudaFree(c->d_list_offsets);
    cudaFree(c->d_out);
    cudaFree(c->d_sums);
    delete c;
}

uint32_t* dint_vroom_ctx_device_output(const dint_vroom_ctx* c) { return c ? c->d_out : nullptr; }

int dint_vroom_decode_host(dint_vroom_ctx* c, const void* h_bin, size_t len, int mode, uint32_t* h_out,
                           uint64_t* h_sums, void* stream) {
    if (!c || (!h_bin && len)) return DINT_E_ARG;
    if (len != c->in_bytes) {
        dint::set_error("dint_vroom_decode_host: .bin size differs from the plan's");
        return DINT_E_ARG;
    }
    device_guard g(c->dict->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (len) CUDA_TRY(cudaMemcpyAsync(c->d_in, h_bin, len, cudaMemcpyHostToDevice, st));
    int rc = dint_decode_segments(c->dict, c->d_in, len, c->d_segs, c->n_segs, c->block_size, c->d_out, mode, stream);
    if (rc) return rc;
    if (h_sums && c->n_lists) {
        rc = dint_checksum_lists(c->d_out, c->d_list_offsets, c->n_lists, c->d_sums, stream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_sums, c->d_sums, c->n_lists * 8, cudaMemcpyDeviceToHost, st));
    }
    if (h_out && c->n_ints) CUDA_TRY(cudaMemcpyAsync(h_out, c->d_out, c->n_ints * 4, cudaMemcpyDeviceToHost, st));
    return DINT_OK;
}

int dint_decode_list_host(const dint_dict* d, const uint8_t* in, size_t in_avail, uint32_t* out, uint32_t n,
                          uint32_t block_size, size_t* consumed) {
    if (!d || !in || (!out && n)) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (!block_size) block_size = dint::kDefaultBlock;
    if (n == 0) {
        if (consumed) *consumed = 0;
        return DINT_OK;
    }
    device_guard g(d->device);
    // a list never needs more than 6 bytes per integer plus one selector byte per block
    size_t need = size_t(n) * 6 + (size_t(n) / block_size + 2) + 16;
    size_t nbytes = std::min(in_avail, need);
    uint8_t* d_in = nullptr;
    uint32_t* d_out = nullptr;
    uint64_t* d_cons = nullptr;
    int rc = DINT_OK;
    cudaError_t e;
    auto fail = [&](const char* what, cudaError_t err) {
        dint::set_error(std::string(what) + ": " + cudaGetErrorString(err));
        rc = DINT_E_CUDA;
    };
    if ((e = cudaMalloc(&d_in, nbytes + 16)) != cudaSuccess) fail("cudaMalloc", e);
    if (!rc && (e = cudaMalloc(&d_out, size_t(n) * 4)) != cudaSuccess) fail("cudaMalloc", e);
    if (!rc && (e = cudaMalloc(&d_cons, 8)) != cudaSuccess) fail("cudaMalloc", e);
    if (!rc && (e = cudaMemcpy(d_in, in, nbytes, cudaMemcpyHostToDevice)) != cudaSuccess) fail("cudaMemcpy H2D", e);
    if (!rc) {
        size_t smem = smem_bytes_of(d, 1);
        e = cudaFuncSetAttribute(decode_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) fail("cudaFuncSetAttribute", e);
        if (!rc) {
            decode_list_kernel<<<1, 32, smem>>>(d->view(), lut_bytes_of(d), d->kind == DINT_DICT_MULTI, d_in, nbytes, n,
                                                block_size, d_out, d_cons);
            ++g_launches;
            if ((e = cudaGetLastError()) != cudaSuccess) fail("decode_list_kernel", e);
        }
    }
    uint64_t cons = 0;
    if (!rc && (e = cudaMemcpy(out, d_out, size_t(n) * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) fail("cudaMemcpy D2H", e);
    if (!rc && (e = cudaMemcpy(&cons, d_cons, 8, cudaMemcpyDeviceToHost)) != cudaSuccess) fail("cudaMemcpy D2H", e);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_cons);
    if (!rc && consumed) *consumed = size_t(cons);
    return rc;
}

}  // extern "C"
