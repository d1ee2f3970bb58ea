// dint_b200 — the byte-aligned block codecs of a frozen block_freq_index (device code, sm_100a).
//
// block_freq_index<vbyte_block> and block_freq_index<u32_block> (include/ds2i/index_types.hpp:44-45) keep the list
// layout of every block index (block_posting_list.hpp:18-47: vbyte n . u32 block_max[blocks] .
// u32 block_endpoint[blocks-1] . blocks) and differ in how a block of up to 256 d-gaps / frequencies-minus-one is
// written: TightVariableByte (block_codecs.hpp:87-101: 7-bit groups, least significant first, the LAST byte of a
// value has bit 7 set) or the raw little-endian u32 values (block_codecs.hpp:377-402). The docs of a block are
// followed at once by its freqs. The block table (dint_block) is the one the DINT index decoders use; one warp
// decodes one block:
//
//   vbyte  32 bytes per trip, one per lane; a ballot of the terminator bits numbers the values that end in the
//          trip, every terminator lane gathers its (<= 5) bytes, a warp scan turns d-gaps into docIDs
//   u32    32 values per trip, funnel loads at any alignment, the same scan
//
// Both write the final docIDs / frequencies (enumerator semantics, block_posting_list.hpp:113-124, 164-169:
// doc_i = cur_base + gap_0 + ... + gap_i + i, freq_i = value_i + 1) with coalesced stores.
#pragma once

#include "dint_strips.cuh"

namespace dintk {

enum { PLAIN_VBYTE = 0, PLAIN_U32 = 1 };

// n TightVariableByte values from byte `pos` on, by one warp. emit(valid, idx, value) is called by ALL lanes once per
// trip (it may hold warp collectives); the lanes whose byte terminates value idx < n pass valid = true. Returns the
// byte position behind the n-th value. Bytes past the end of m_lists read as terminators (the stream always ends).
template <class Emit>
__device__ __forceinline__ uint64_t vbyte_block_stream(const uint8_t* __restrict__ lists, uint64_t lists_len, uint64_t pos,
                                                       uint32_t n, Emit emit) {
    const unsigned lane = lane_id(), lt = lanemask_lt();
    uint32_t done = 0;     // values in front of this trip
    uint32_t pending = 0;  // bytes of an unfinished value in front of this trip
    while (done < n) {
        const uint64_t q = pos + lane;
        const uint32_t c = q < lists_len ? __ldg(lists + q) : 128u;
        const uint32_t term = __ballot_sync(FULL, c & 128u);
        const uint32_t below = term & lt;
        const uint32_t idx = done + uint32_t(__popc(below));
        const bool mine = (c & 128u) && idx < n;
        uint32_t v = 0;
        if (mine) {
            // the value starts behind the previous terminator (or `pending` bytes in front of the trip)
            uint32_t len = below ? lane - (31u - uint32_t(__clz(int(below)))) : lane + 1u + pending;
            len = min(len, 5u);  // longer runs of continuation bytes are malformed; the reference shifts them out
            const uint8_t* s = lists + (q + 1 - len);
            v = c & 127u;
            v <<= 7u * (len - 1u);
            for (uint32_t k = 0; k + 1 < len; ++k) v += (uint32_t(__ldg(s + k)) & 127u) << (7u * k);
        }
        emit(mine, idx, v);
        const uint32_t cnt = uint32_t(__popc(term));
        if (done + cnt >= n) return pos + uint32_t(__fns(term, 0, int(n - done))) + 1u;
        done += cnt;
        pending = term ? uint32_t(__clz(int(term))) : pending + 32u;
        pos += 32;
    }
    return pos;
}

// n raw u32 values from byte `pos` on
template <class Emit>
__device__ __forceinline__ uint64_t u32_block_stream(const uint8_t* __restrict__ lists, uint64_t lists_len, uint64_t pos, uint32_t n,
                                                     Emit emit) {
    const unsigned lane = lane_id();
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t idx = base + lane;
        const uint64_t q = pos + 4ull * idx;
        const bool mine = idx < n && q + 4 <= lists_len;
        emit(idx < n, idx, mine ? ldu32(lists + q) : 0u);
    }
    return pos + 4ull * n;
}

template <int CODEC, class Emit>
__device__ __forceinline__ uint64_t plain_block_stream(const uint8_t* __restrict__ lists, uint64_t lists_len, uint64_t pos,
                                                       uint32_t n, Emit emit) {
    if (CODEC == PLAIN_U32) return u32_block_stream(lists, lists_len, pos, n, emit);
    return vbyte_block_stream(lists, lists_len, pos, n, emit);
}

// one warp per block of the table; d_docids or d_freqs may be null (that stream is skipped — the docs of a vbyte
// block are still walked to find where its freqs start)
template <int CODEC>
__global__ void __launch_bounds__(256) index_plain_blocks_kernel(const uint8_t* __restrict__ lists, uint64_t lists_len,
                                                                 const dint_block* __restrict__ blocks, uint64_t n_blocks,
                                                                 uint32_t* __restrict__ docids, uint32_t* __restrict__ freqs) {
    const uint64_t gwarp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t b = gwarp; b < n_blocks; b += nwarps) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(blocks + b));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(blocks + b) + 1);
        const uint64_t in_off = uint64_t(q0.x) | (uint64_t(q0.y) << 32), out_off = uint64_t(q0.z) | (uint64_t(q0.w) << 32);
        const uint32_t size = q1.x, cur_base = q1.y;
        uint64_t fpos;
        if (docids || CODEC == PLAIN_VBYTE) {
            uint32_t carry = cur_base;  // docID in front of the trip's first value, + its index
            fpos = plain_block_stream<CODEC>(lists, lists_len, in_off, size, [&](bool valid, uint32_t idx, uint32_t v) {
                if (!docids) return;
                const uint32_t incl = warp_inclusive_scan(valid ? v : 0u);
                if (valid) docids[out_off + idx] = carry + incl + idx;
                carry += __shfl_sync(FULL, incl, 31);
            });
        } else {
            fpos = in_off + 4ull * size;
        }
        if (freqs)
            plain_block_stream<CODEC>(lists, lists_len, fpos, size, [&](bool valid, uint32_t idx, uint32_t v) {
                if (valid) freqs[out_off + idx] = v + 1u;
            });
    }
}

}  // namespace dintk
