// dint_b200 — block table of a frozen dict_freq_index, built on the device (device code, sm_100a).
//
// The device-side twin of dint_index_plan_create (dint_host.cpp): the host reads the fixed-size header of the
// file (list count, Elias-Fano geometry, byte ranges) and nothing else; the GPU expands the Elias-Fano encoded
// list endpoints (compact_elias_fano.hpp:14-61, 69-136: pointers0 . pointers1 . high bits . low bits inside one
// bit vector), parses every list header (dict_posting_list.hpp:10-56: vbyte n . u32 block_max[blocks] .
// u32 block_endpoint[blocks-1] . blocks) and writes the dint_block table the decoders consume.
//
//   ef_popc_kernel      ones per 64-bit word of the high-bits vector
//   scan_u32_kernel     exclusive scan (u32 in, u64 out) by one CTA — load-time, the inputs are a few MB
//   ef_emit_kernel      i-th one at bit q of the high bits -> endpoint i = ((q - i - 1) << l) | low bits of i
//   index_lists_kernel  one thread per list: vbyte n at the endpoint, number of blocks, bounds
//   index_blocks_kernel one warp per 32 lists: each lane writes the first blocks of its list, lists with many
//                       blocks are then written by the whole warp
#pragma once

#include "dint_decoder.cuh"

namespace dintk {

// `len` (<= 64) bits at bit position `pos` of a little-endian u64 word array that starts at `words` (any alignment)
__device__ __forceinline__ uint64_t ef_bits(const uint8_t* words, uint64_t pos, uint32_t len) {
    if (!len) return 0;
    const uint64_t w = pos >> 6;
    const uint32_t s = uint32_t(pos & 63);
    uint64_t v = ldu64(words + w * 8) >> s;
    if (s + len > 64u) v |= ldu64(words + (w + 1) * 8) << (64u - s);
    return len == 64u ? v : v & ((uint64_t(1) << len) - 1);
}

struct ef_params {
    const uint8_t* words;   // the bit vector's words inside the file image (device)
    uint64_t n;             // lists
    uint64_t hb_off, hb_len, lb_off;
    uint32_t l;             // low bits per endpoint
    uint64_t n_hwords;      // ceil(hb_len / 64)
    uint32_t* popc;         // [n_hwords]
    const uint64_t* rank;   // [n_hwords] exclusive scan of popc
    uint64_t* endpoints;    // [n + 1]
};

__global__ void __launch_bounds__(256) ef_popc_kernel(const ef_params P) {
    const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < P.n_hwords; j += step) {
        const uint32_t take = uint32_t(min(uint64_t(64), P.hb_len - 64 * j));
        P.popc[j] = uint32_t(__popcll(ef_bits(P.words, P.hb_off + 64 * j, take)));
    }
}

__global__ void __launch_bounds__(256) ef_emit_kernel(const ef_params P) {
    const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < P.n_hwords; j += step) {
        const uint32_t take = uint32_t(min(uint64_t(64), P.hb_len - 64 * j));
        uint64_t w = ef_bits(P.words, P.hb_off + 64 * j, take);
        uint64_t i = P.rank[j];
        while (w && i < P.n) {
            const uint32_t b = uint32_t(__ffsll(static_cast<long long>(w))) - 1u;
            w &= w - 1;
            const uint64_t high = 64 * j + b - i - 1;  // compact_elias_fano.hpp: the i-th one sits at high_i + i + 1
            P.endpoints[i] = (high << P.l) | ef_bits(P.words, P.lb_off + i * P.l, P.l);
            ++i;
        }
    }
}

// exclusive scan of n u32 counts into u64 offsets, total into out[n]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_u32_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t n) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t k = base + threadIdx.x;
        const unsigned long long v = k < n ? in[k] : 0ull;
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = shfl64(x, lane >= unsigned(o) ? int(lane) - o : int(lane));
            if (lane >= unsigned(o)) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        unsigned long long before = s_carry;
        for (unsigned w = 0; w < warp; ++w) before += s_warp[w];
        if (k < n) out[k] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

struct index_lists_params {
    const uint8_t* lists;      // m_lists bytes (device)
    uint64_t lists_len, n;
    uint32_t block_size;
    const uint64_t* endpoints; // [n] (+ [n] = lists_len written by the host)
    uint32_t* list_n;          // [n] postings of every list
    uint32_t* list_blocks;     // [n] blocks of every list
    uint32_t* status;          // != 0: malformed
};

__global__ void __launch_bounds__(256) index_lists_kernel(const index_lists_params P) {
    const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < P.n; t += step) {
        const uint64_t ep = P.endpoints[t];
        bool ok = ep < P.lists_len && (t == 0 || ep >= P.endpoints[t - 1]);
        uint32_t n = 0, k = 0;
        if (ok) {  // TightVariableByte (block_codecs.hpp:87-101): 7-bit groups, the terminator has bit 7 set
            ok = false;
            for (; k < 5u && ep + k < P.lists_len; ++k) {
                const uint32_t c = __ldg(P.lists + ep + k);
                n += (c & 127u) << (7u * k);
                if (c & 128u) {
                    ok = true;
                    break;
                }
            }
        }
        const uint64_t blocks = (uint64_t(n) + P.block_size - 1) / P.block_size;
        // u32 block_max[blocks] . u32 block_endpoint[blocks - 1] must fit behind the header
        if (ok) ok = n != 0u && 4 * blocks + 4 * (blocks - 1) <= P.lists_len - (ep + k + 1);
        if (!ok) {
            atomicMax(P.status, 1u);
            n = 0;
        }
        P.list_n[t] = n;
        P.list_blocks[t] = ok ? uint32_t(blocks) : 0u;
    }
}

struct index_blocks_params {
    const uint8_t* lists;
    uint64_t lists_len, n;
    uint32_t block_size;
    const uint64_t* endpoints;
    const uint32_t* list_n;
    const uint64_t* list_offsets;  // [n + 1] postings in front of every list
    const uint64_t* block_base;    // [n + 1] blocks in front of every list
    dint_block* blocks;
    uint32_t* status;
};

// block b of a list whose tables start at `maxs` (dict_posting_list.hpp:90-107, 284-318)
__device__ __forceinline__ void write_block(const index_blocks_params& P, const uint8_t* maxs, uint64_t blocks, uint64_t data_off,
                                            uint64_t list_off, uint32_t n, uint64_t b, dint_block* dst) {
    const uint8_t* eps = maxs + 4 * blocks;
    const uint32_t endpoint = b ? ldu32(eps + 4 * (b - 1)) : 0u;
    const uint64_t in_off = data_off + endpoint;
    if (in_off >= P.lists_len) atomicMax(P.status, 1u);
    const uint32_t size = (b + 1) * P.block_size <= n ? P.block_size : n % P.block_size;
    const uint32_t cur_base = b ? ldu32(maxs + 4 * (b - 1)) + 1u : 0u;
    const uint32_t block_max = ldu32(maxs + 4 * b);
    const uint64_t out_off = list_off + b * P.block_size;
    uint4* q = reinterpret_cast<uint4*>(dst);
    q[0] = make_uint4(uint32_t(in_off), uint32_t(in_off >> 32), uint32_t(out_off), uint32_t(out_off >> 32));
    q[1] = make_uint4(size, cur_base, block_max, 0u);
}

constexpr uint32_t kLaneBlocks = 4;  // blocks of its list a lane writes by itself

__global__ void __launch_bounds__(256) index_blocks_kernel(const index_blocks_params P) {
    const unsigned lane = lane_id();
    const uint64_t n_groups = (P.n + 31) / 32;
    const uint64_t gwarp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t g = gwarp; g < n_groups; g += nwarps) {
        const uint64_t t = g * 32 + lane;
        uint32_t n = 0;
        uint64_t blocks = 0, bb = 0, lo = 0, data_off = 0;
        const uint8_t* maxs = P.lists;
        if (t < P.n) {
            n = P.list_n[t];
            bb = P.block_base[t];
            blocks = P.block_base[t + 1] - bb;
            lo = P.list_offsets[t];
            uint64_t p = P.endpoints[t];
            while (p < P.lists_len && !(__ldg(P.lists + p) & 128u)) ++p;  // behind the vbyte n
            maxs = P.lists + p + 1;
            data_off = (p + 1) + 4 * blocks + 4 * (blocks ? blocks - 1 : 0);
        }
        for (uint64_t b = 0; b < min(blocks, uint64_t(kLaneBlocks)); ++b) write_block(P, maxs, blocks, data_off, lo, n, b, P.blocks + bb + b);
        // lists with more blocks: the whole warp writes the rest of each, 32 blocks at a time
        uint32_t big = __ballot_sync(FULL, blocks > kLaneBlocks);
        while (big) {
            const int k = __ffs(big) - 1;
            big &= big - 1;
            const uint64_t kblocks = shfl64(blocks, k), kbb = shfl64(bb, k), klo = shfl64(lo, k), kdata = shfl64(data_off, k);
            const uint32_t kn = __shfl_sync(FULL, n, k);
            const uint8_t* kmaxs = reinterpret_cast<const uint8_t*>(shfl64(reinterpret_cast<uint64_t>(maxs), k));
            for (uint64_t b = kLaneBlocks + lane; b < kblocks; b += 32) write_block(P, kmaxs, kblocks, kdata, klo, kn, b, P.blocks + kbb + b);
        }
    }
}

}  // namespace dintk
