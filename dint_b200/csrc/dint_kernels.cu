// dint_b200 — sm_100a kernels and the device half of the C ABI.
//
// The stream decoder itself (tiles, staging ring, fused flush) is in dint_decoder.cuh; this
// file holds the kernels that drive it (work units of a vroom .bin, blocks of a frozen
// dict_freq_index), the interpolative tail-block kernel, the checksum kernel, and the
// extern "C" entry points.
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

#include "dint_cuda_util.hpp"
#include "dint_decoder.cuh"
#include "dint_strips.cuh"
#include "dint_block_codecs.cuh"
#include "dint_internal.hpp"

namespace dintk {
std::atomic<uint64_t> g_launches{0};
}

namespace {

using namespace dintk;

constexpr int kWarpsPerCta = 16;
constexpr int kDefaultFreqsVariant = 7;  // strip decoder variant of frequency streams (7: bulk-copy flush)
constexpr int kThreads = kWarpsPerCta * 32;
constexpr uint32_t kImplicitUnit = 8;  // segments per work unit when the caller passes no unit table

extern __shared__ __align__(16) uint8_t g_smem[];

constexpr uint32_t kClassBytes = 64;  // class_info[8] between the length table and the rings

// Stages the 4-bit log2(length) table and the class table into shared memory and carves the
// calling warp's rings. Returns the class table.
__device__ __forceinline__ const class_info* cta_setup(const dint_dict_view& dv, uint32_t lut_bytes, const uint8_t* in,
                                                       uint64_t in_bytes, uint32_t* out, warp_state& ws) {
    const uint4* src = reinterpret_cast<const uint4*>(dv.lens);
    uint4* dst = reinterpret_cast<uint4*>(g_smem);
    for (uint32_t i = threadIdx.x; i < lut_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    class_info* classes = reinterpret_cast<class_info*>(g_smem + lut_bytes);
    if (threadIdx.x < 8) {
        const uint32_t c = threadIdx.x;
        class_info ci = {0, 0};
        if (c < dv.n_classes) {
            ci.row_base = dv.class_base[c];
            ci.rows = (c + 1 < dv.n_classes ? dv.class_base[c + 1] : dv.n_rows) - ci.row_base;
        }
        classes[c] = ci;
    }
    uint8_t* mine = g_smem + lut_bytes + kClassBytes + (threadIdx.x >> 5) * kWarpSmem;
    ws.stage_s = smem_u32(mine);
    ws.inb_s = ws.stage_s + kStageInts * 4;
    ws.que_s = ws.inb_s + kInBytes;
    ws.lut_s = smem_u32(g_smem);
    for (uint32_t i = lane_id(); i < kStageInts / 4; i += 32) reinterpret_cast<uint4*>(mine)[i] = make_uint4(0, 0, 0, 0);
    ws.out = out;
    ws.flushed = 0;
    ws.staged = 0;
    ws.acc = 0;
    ws.anchor = 0;
    ws.win_end = 0;
    ws.safe_end = 0;
    ws.fast_end = 0;
    ws.lo_rel = 0;
    ws.hi_rel = 0;
    const uint64_t a = reinterpret_cast<uint64_t>(in);
    ws.in_lo = a & ~uint64_t(15);
    ws.in_hi = (a + in_bytes + 15) & ~uint64_t(15);
    ws.vec_ok = (reinterpret_cast<uint64_t>(out) & 15) == 0;
    __syncthreads();
    return classes;
}

// vroom path: one warp per work unit (<= 32 consecutive segments, normally contiguous in the
// file and in the output, which makes the whole unit one flat run).
template <bool MULTI, int MODE, int W>
__global__ void __launch_bounds__(W * 32, W <= 16 ? 2 : 1)
decode_units_kernel(dint_dict_view dv, uint32_t lut_bytes, const uint8_t* __restrict__ in, uint64_t in_bytes,
                    const dint_segment* __restrict__ segs, uint64_t n_segs, const uint32_t* __restrict__ unit_first,
                    uint64_t n_units, uint32_t* __restrict__ out) {
    warp_state ws;
    const class_info* classes = cta_setup(dv, lut_bytes, in, in_bytes, out, ws);
    const unsigned lane = lane_id();
    const uint64_t in_base = reinterpret_cast<uint64_t>(in);
    const uint64_t gwarp = uint64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = uint64_t(gridDim.x) * (blockDim.x >> 5);
    for (uint64_t u = gwarp; u < n_units; u += nwarps) {
        uint64_t s0, s1;
        if (unit_first) {
            s0 = unit_first[u];
            s1 = unit_first[u + 1];
        } else {
            s0 = u * kImplicitUnit;
            s1 = min(n_segs, s0 + kImplicitUnit);
        }
        ws.staged = 0;
        uint32_t tail = 0;  // docid of the last integer decoded so far
        for (uint64_t sb = s0; sb < s1; sb += 32) {
            uint32_t cnt = uint32_t(min(uint64_t(32), s1 - sb));
            dint_segment sg = {0, 0, 0, 0, 0, 0};
            if (lane < cnt) sg = segs[sb + lane];
            while (cnt) {
                // ---- the longest prefix of the batch that forms one flat run -----------------------
                const bool has = lane < cnt;
                const uint64_t out0 = shfl64(sg.out_off, 0), in0 = shfl64(sg.in_off, 0);
                const uint32_t n_incl = warp_inclusive_scan(has ? sg.n : 0u);
                const uint64_t my_end = sg.in_off + sg.in_len;
                const uint64_t prev_end = shfl64(my_end, lane ? int(lane) - 1 : 0);
                bool ok = sg.in_len != 0 && sg.n != 0 && my_end <= in_bytes && my_end - in0 < kMaxRunBytes &&
                          sg.out_off == out0 + (n_incl - sg.n);
                if (lane) ok = ok && sg.in_off >= prev_end && sg.in_off - prev_end <= 12;  // vroom list headers only
                const uint32_t bad = __ballot_sync(FULL, has && !ok);
                const uint32_t r = bad ? uint32_t(__ffs(bad) - 1) : cnt;
                const uint32_t doc_base0 = __shfl_sync(FULL, sg.doc_base, 0);
                if (ws.staged && out0 != ws.flushed + ws.staged) {  // output not contiguous: start over
                    ws.flush_all<MODE>(lane);
                    __syncwarp();
                }
                uint32_t corr0 = 0;
                if (!ws.staged) {
                    ws.flushed = out0;
                    ws.acc = doc_base0 - 1u;  // docid[i] = acc + sum_{j<=i} (v_j + 1)
                } else {
                    corr0 = doc_base0 - 1u - tail;  // re-base the running prefix sum at the list boundary
                }
                if (r == 0) {
                    // segment 0 has no usable byte length (or is empty): walk it alone, like the reference
                    const uint32_t n0 = __shfl_sync(FULL, sg.n, 0);
                    if (corr0) {  // keep it simple: no carried prefix into a segment of unknown length
                        ws.flush_all<MODE>(lane);
                        __syncwarp();
                        ws.flushed = out0;
                        ws.acc = doc_base0 - 1u;
                    }
                    if (in0 < in_bytes) decode_one<MODE>(ws, dv.rows, classes, dv.n_classes, MULTI, in_base + in0, n0, lane);
                    ws.flush_all<MODE>(lane);
                    __syncwarp();
                } else {
                    const bool hasr = lane < r;
                    seg_lane sl;
                    uint32_t hdr = 0, narrow = 0;
                    sl.pk = 0;
                    sl.rows = dv.n_rows;
                    if (MULTI) {
                        const uint32_t sel = hasr ? uint32_t(__ldg(in + sg.in_off)) : 0u;
                        const bool wide = sel < dint::kSelectors;
                        const class_info ci = classes[min(wide ? sel : sel - dint::kSelectors, dv.n_classes - 1u)];
                        narrow = wide ? 0u : 1u;
                        hdr = 1;
                        sl.pk = ci.row_base | (narrow << 31);
                        sl.rows = wide ? ci.rows : min(ci.rows, 256u);
                    }
                    const uint32_t ns = hasr ? (sg.in_len - hdr) >> (narrow ? 0 : 1) : 0u;
                    const uint32_t cs_incl = warp_inclusive_scan(ns);
                    sl.cs = cs_incl - ns;
                    const uint32_t S = __shfl_sync(FULL, cs_incl, 31);
                    const uint32_t N = __shfl_sync(FULL, n_incl, r - 1);
                    const uint64_t run_end = shfl64(my_end, r - 1);
                    const uint32_t a0 = ws.set_anchor(in_base + in0, in_base + run_end + 8);
                    sl.ab = a0 + uint32_t(sg.in_off - in0) + hdr;
                    const uint32_t my_tail = sg.doc_base - 1u + sg.doc_sum;
                    const uint32_t prev_tail = __shfl_up_sync(FULL, my_tail, 1);
                    sl.corr = lane ? sg.doc_base - 1u - prev_tail : corr0;
                    decode_run_queued<MULTI, MODE>(ws, dv.rows, sl, r, S, N, lane);
                    tail = __shfl_sync(FULL, my_tail, r - 1);
                }
                // ---- drop the segments just decoded ---------------------------------------------------
                const uint32_t drop = r ? r : 1;
                cnt -= drop;
                if (cnt) {
                    sg.in_off = shfl64(sg.in_off, (lane + drop) & 31);
                    sg.out_off = shfl64(sg.out_off, (lane + drop) & 31);
                    sg.n = __shfl_sync(FULL, sg.n, (lane + drop) & 31);
                    sg.doc_base = __shfl_sync(FULL, sg.doc_base, (lane + drop) & 31);
                    sg.in_len = __shfl_sync(FULL, sg.in_len, (lane + drop) & 31);
                    sg.doc_sum = __shfl_sync(FULL, sg.doc_sum, (lane + drop) & 31);
                }
            }
        }
        ws.flush_all<MODE>(lane);
        __syncwarp();
    }
}

// Full blocks of a dict_freq_index, pass 1 (docs): decode_docs_block (dict_posting_list.hpp:284-309).
// The block's docs byte length (= where its freqs bytes start) is left in the first slot of
// the block's freqs output for pass 2.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
index_docs_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ lists,
                  uint64_t lists_bytes, const dint_block* __restrict__ blocks, uint64_t n_blocks, uint32_t block_size,
                  uint32_t* __restrict__ docids, uint32_t* __restrict__ freqs) {
    warp_state ws;
    const class_info* classes = cta_setup(dv, lut_bytes, lists, lists_bytes, docids, ws);
    const unsigned lane = lane_id();
    const uint64_t base = reinterpret_cast<uint64_t>(lists);
    const uint64_t gwarp = uint64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = uint64_t(gridDim.x) * (blockDim.x >> 5);
    for (uint64_t i = gwarp; i < n_blocks; i += nwarps) {
        const dint_block b = blocks[i];
        if (b.size != block_size) continue;  // partial block: interpolative, index_tails_kernel
        ws.staged = 0;
        ws.flushed = b.out_off;
        ws.acc = b.cur_base - 1u;  // docid[0] = buf[0] + cur_base (dict_posting_list.hpp:304)
        const uint64_t end = decode_one<MODE>(ws, dv.rows, classes, dv.n_classes, multi != 0, base + b.in_off,
                                              b.size, lane);
        if (MODE != OUT_SKIP) {
            ws.flush_all<MODE>(lane);
            __syncwarp();
        }
        if (freqs && lane == 0) freqs[b.out_off] = uint32_t(end - base - b.in_off);
    }
}

// pass 2 (freqs): decode_freqs_block (dict_posting_list.hpp:311-318), freq() = buf[i] + 1 (:168)
__global__ void __launch_bounds__(kThreads, 2)
index_freqs_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ lists,
                   uint64_t lists_bytes, const dint_block* __restrict__ blocks, uint64_t n_blocks, uint32_t block_size,
                   uint32_t* __restrict__ freqs) {
    warp_state ws;
    const class_info* classes = cta_setup(dv, lut_bytes, lists, lists_bytes, freqs, ws);
    const unsigned lane = lane_id();
    const uint64_t base = reinterpret_cast<uint64_t>(lists);
    const uint64_t gwarp = uint64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = uint64_t(gridDim.x) * (blockDim.x >> 5);
    for (uint64_t i = gwarp; i < n_blocks; i += nwarps) {
        const dint_block b = blocks[i];
        if (b.size != block_size) continue;
        const uint32_t docs_len = freqs[b.out_off];
        __syncwarp();
        ws.staged = 0;
        ws.flushed = b.out_off;
        decode_one<OUT_FREQS>(ws, dv.rows, classes, dv.n_classes, multi != 0, base + b.in_off + docs_len, b.size,
                              lane);
        ws.flush_all<OUT_FREQS>(lane);
        __syncwarp();
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
// `stack`: this thread's column of a level-major shared-memory stack (entry k at stack[k * stride]; depth <= 8 for
// n <= 255), so that the lanes of a warp, which push and pop in lock step, hit consecutive 16-byte slots
template <typename Emit>
__device__ void read_interpolative(bit_reader& br, uint32_t n, uint32_t low0, uint32_t high0, uint4* stack, uint32_t stride,
                                   Emit&& emit) {
    // ONE node per trip, whatever the shape of the tree: lanes that decode tails of similar length then stay in the
    // same instructions (a walk down the left spine inside a pop loop made every lane wait for the longest spine)
    int sp = 0;
    bool have = n != 0u;
    uint32_t off = 0, cnt = n, lo = low0, hi = high0;
    for (;;) {
        if (!have) {
            if (sp == 0) break;
            --sp;
            const uint4 e = stack[sp * stride];
            off = e.x;
            cnt = e.y;
            lo = e.z;
            hi = e.w;
        }
        const uint32_t h = cnt / 2;
        const uint32_t val = lo + br.read_int(hi - lo + 1u);
        emit(off + h, val);
        const uint32_t rn = cnt - h - 1;
        if (rn) {  // right half later
            stack[sp * stride] = make_uint4(off + h + 1, rn, val, hi);
            ++sp;
        }
        have = h != 0u;  // left half now
        cnt = h;
        hi = val;
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

// Writes the n values of a decoded tail (this thread's local array) to dst: scalar stores up to the first 16-byte
// boundary, then 16-byte stores. The interpolative walk produces the values in tree order; storing each one where it
// belongs as it appears meant eight separate 4-byte stores per 32-byte sector, spread over time — L2 then fetches
// every sector it is about to partially overwrite (measured: 2 x the output bytes written + as many read).
// DIFF: the values are prefix sums, store their adjacent differences + 1 (the freqs of a tail).
template <bool DIFF>
__device__ __forceinline__ void write_tail(uint32_t* __restrict__ dst, const uint32_t* P, uint32_t n) {
    uint32_t prev = 0, j = 0;
    auto val = [&](uint32_t k) {
        const uint32_t v = P[k];
        const uint32_t r = DIFF ? v - prev + 1u : v;
        prev = v;
        return r;
    };
    for (; j < n && (reinterpret_cast<uintptr_t>(dst + j) & 15u); ++j) dst[j] = val(j);
    for (; j + 4u <= n; j += 4u) {
        uint4 q;
        q.x = val(j);
        q.y = val(j + 1u);
        q.z = val(j + 2u);
        q.w = val(j + 3u);
        *reinterpret_cast<uint4*>(dst + j) = q;
    }
    for (; j < n; ++j) dst[j] = val(j);
}

// One thread per partial block. A WARP takes kTailChunk consecutive entries of the block table (neighbouring lists:
// their bytes and their outputs are neighbours too, which keeps the accesses of a warp inside a few KB), sorts the
// partial blocks among them by size in shared memory (warp-local counting sort, no CTA barrier), and its lanes take
// them in that order: the 32 lanes then decode tails of (nearly) the same length, whose interpolative trees have the
// same shape, so the warp runs in lock step instead of waiting for its longest tail (tails hold 1..255 postings).
// The freqs prefix sums of a tail stay in the thread's local memory until their adjacent differences are taken
// (lanes in lock step touch the same element index: one 128-byte line per access), so nothing is read back from
// global memory.
constexpr int kTailThreads = 128;
constexpr int kTailChunk = 256;
struct tail_warp_smem {
    unsigned int bins[257];
    uint8_t order[kTailChunk];
    uint16_t sizes[kTailChunk];
};
__global__ void __launch_bounds__(kTailThreads)
index_tails_kernel(const uint8_t* __restrict__ lists, uint64_t lists_bytes, const dint_block* __restrict__ blocks,
                   uint64_t n_blocks, uint32_t block_size, uint32_t chunk, uint32_t* __restrict__ docids,
                   uint32_t* __restrict__ freqs) {
    // block_size: blocks of fewer postings are interpolative coded (dict_freq_index: the partial blocks; a
    // block_freq_index<interpolative_block> is decoded by passing 257: every block of up to 256 postings)
    __shared__ uint4 stacks[kTailThreads * 10];
    __shared__ tail_warp_smem wsm[kTailThreads / 32];
    uint4* stack = stacks + threadIdx.x;  // level-major: entry k at stack[k * kTailThreads]
    tail_warp_smem& W = wsm[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    uint32_t P[256];  // the values of the tail in hand, in list order (local memory)
    // chunk (<= kTailChunk, a multiple of 32): small tables are cut finer so that every SM gets work
    const uint64_t n_chunks = (n_blocks + chunk - 1) / chunk;
    const uint64_t n_warps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t c = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; c < n_chunks; c += n_warps) {
        const uint64_t base = c * chunk;
        const uint32_t cn = uint32_t(min(uint64_t(chunk), n_blocks - base));
        // ---- counting sort of the chunk's partial blocks by size (descending) -------------------------------------
        __syncwarp();
        for (uint32_t t = lane; t < 257u; t += 32u) W.bins[t] = 0;
        __syncwarp();
        for (uint32_t t = lane; t < cn; t += 32u) {
            uint32_t n = blocks[base + t].size;
            if (n >= block_size || n > 256u) n = 0;  // full blocks are DINT coded (strips / tiles)
            W.sizes[t] = uint16_t(n);
            if (n) atomicAdd(&W.bins[256u - n], 1u);
        }
        __syncwarp();
        {
            unsigned int v[8], sum = 0;
            for (int q = 0; q < 8; ++q) {
                v[q] = W.bins[lane * 8 + q];
                sum += v[q];
            }
            const unsigned int incl = warp_inclusive_scan(sum);
            unsigned int acc = incl - sum;
            __syncwarp();
            for (int q = 0; q < 8; ++q) {
                W.bins[lane * 8 + q] = acc;
                acc += v[q];
            }
            if (lane == 31) W.bins[256] = incl;
        }
        __syncwarp();
        const uint32_t n_tails = W.bins[256];
        for (uint32_t t = lane; t < cn; t += 32u) {
            const uint32_t n = W.sizes[t];
            if (n) W.order[atomicAdd(&W.bins[256u - n], 1u)] = uint8_t(t);
        }
        __syncwarp();
        // ---- decode ------------------------------------------------------------------------------------------------
        for (uint32_t k = lane; k < n_tails; k += 32u) {
            const uint64_t i = base + W.order[k];
            const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(blocks + i));
            const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(blocks + i) + 1);
            const uint32_t n = b1.x, cur_base = b1.y, block_max = b1.z;
            uint64_t p = (uint64_t(b0.y) << 32) | b0.x;
            const uint64_t out_off = (uint64_t(b0.w) << 32) | b0.z;
            // docs: prefix sums P_i of the first n-1 values in [0, S], P_{n-1} = S
            // (block_codecs.hpp:139-146); docid_i = cur_base + P_i + i (dict_posting_list.hpp:304-307,122)
            const uint32_t S = block_max - cur_base - (n - 1u);
            uint32_t* dd = docids ? docids + out_off : nullptr;
            if (n > 1) {
                bit_reader br;
                br.init(lists, lists_bytes, p);
                read_interpolative(br, n - 1, 0, S, stack, kTailThreads, [&](uint32_t idx, uint32_t val) {
                    P[idx] = cur_base + val + idx;
                });
                p += (br.pos + 7) / 8;
            }
            if (dd) {
                P[n - 1] = block_max;
                write_tail<false>(dd, P, n);
            }
            if (!freqs) continue;
            // freqs: sum_of_values == -1 -> leading vbyte (block_codecs.hpp:134-136); prefix sums first
            uint32_t* ff = freqs + out_off;
            const uint32_t Sf = vbyte_at(lists, lists_bytes, p);
            if (n == 1) {
                ff[0] = Sf + 1u;  // freq = v + 1 (dict_posting_list.hpp:168)
                continue;
            }
            {
                bit_reader br;
                br.init(lists, lists_bytes, p);
                read_interpolative(br, n - 1, 0, Sf, stack, kTailThreads, [&](uint32_t idx, uint32_t val) { P[idx] = val; });
            }
            // adjacent differences (block_codecs.hpp:147-149), freq = v + 1
            P[n - 1] = Sf;
            write_tail<true>(ff, P, n);
        }
    }
}

// ---- per-list checksums ---------------------------------------------------------------------
// sums[i] = sum_{j in list i} (value_j + 1) * (j + 1). Work is cut by VALUES, not by lists: one warp
// per tile of kSumTile consecutive values, so neither a 100 K-posting list nor a million
// one-posting lists unbalance the grid. Lists that lie inside a tile are summed by one lane
// each (short ones) or by the whole warp; pieces of lists that straddle tiles are added with
// 64-bit atomics (sums[] is zeroed first).
constexpr uint32_t kSumTile = 4096;
constexpr uint32_t kSumShort = 48;

__global__ void __launch_bounds__(256)
checksum_kernel(const uint32_t* __restrict__ values, const uint64_t* __restrict__ offs, uint64_t n_lists,
                unsigned long long* __restrict__ sums) {
    const unsigned lane = lane_id();
    const uint64_t gwarp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const uint64_t total = offs[n_lists];
    const uint64_t n_tiles = (total + kSumTile - 1) / kSumTile;
    for (uint64_t t = gwarp; t < n_tiles; t += nwarps) {
        const uint64_t t0 = t * kSumTile, t1 = min(total, t0 + kSumTile);
        // last list that starts at or before t0 (lists may be empty: take the last such)
        uint64_t lo = 0, hi = n_lists;  // invariant: offs[lo] <= t0 < offs[hi] (offs[n_lists] = total > t0)
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (offs[mid] <= t0) lo = mid;
            else hi = mid;
        }
        uint64_t i = lo;
        while (i < n_lists) {
            const uint64_t li = i + lane;
            uint64_t b = total, e = total;
            if (li < n_lists) {
                b = offs[li];
                e = offs[li + 1];
            }
            const bool act = li < n_lists && b < t1;
            const uint64_t cb = max(b, t0), ce = min(e, t1);
            const bool whole = cb == b && ce == e;
            const bool shortp = act && ce - cb <= kSumShort;
            if (shortp && ce > cb) {
                unsigned long long acc = 0;
                for (uint64_t j = cb; j < ce; ++j) acc += (uint64_t(values[j]) + 1ull) * (j - b + 1ull);
                if (whole) sums[li] = acc;
                else atomicAdd(&sums[li], acc);
            }
            uint32_t m = __ballot_sync(FULL, act && !shortp);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const uint64_t kb = shfl64(b, k), kcb = shfl64(cb, k), kce = shfl64(ce, k);
                unsigned long long acc = 0;
                for (uint64_t j = kcb + lane; j < kce; j += 32) acc += (uint64_t(values[j]) + 1ull) * (j - kb + 1ull);
#pragma unroll
                for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                if (lane == 0) atomicAdd(&sums[i + k], acc);
            }
            // lists that end inside the tile are done; stop at the first one that reaches its end
            const uint32_t done = __ballot_sync(FULL, act && e <= t1);
            const uint32_t nact = __popc(__ballot_sync(FULL, act));
            if (__popc(done) < 32 || nact < 32) break;
            i += 32;
        }
    }
}

// ---- batched AND-intersection over decoded docID lists (ds2i::and_query, include/ds2i/queries.hpp:34-86) --------
// One CTA per query, its warps sharing the candidates. Lane k holds term k's list (start, length); duplicates are
// dropped like remove_duplicate_terms does (:29-32). The reference walks the shortest list and calls next_geq on the
// others in order of increasing length (:48-72); here every warp tests 32 docIDs of the shortest list at once, each
// lane looking its candidate up in the other lists (shortest first, so most candidates die early) by binary search —
// the lists are sorted arrays in HBM once decoded. counts[q] = documents that hold every term (partial counts of the
// warps are added with atomics; the caller zeroes counts); a term id >= n_lists makes it 0.
constexpr uint32_t kMaxQueryTerms = 32;
constexpr uint32_t kAndWarps = 8;  // warps (one CTA) per query: they share the candidates of the shortest list
__global__ void __launch_bounds__(kAndWarps * 32)
and_queries_kernel(const uint32_t* __restrict__ docids, const uint64_t* __restrict__ offs, uint64_t n_lists,
                   const uint64_t* __restrict__ qoff, const uint32_t* __restrict__ qterms, uint64_t n_queries,
                   unsigned long long* __restrict__ counts) {
    const unsigned lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    for (uint64_t q = blockIdx.x; q < n_queries; q += gridDim.x) {  // counts[] is zeroed by the caller
        const uint64_t t0 = qoff[q], t1 = qoff[q + 1];
        const uint32_t nt = uint32_t(min(t1 - t0, uint64_t(kMaxQueryTerms)));
        if (t1 - t0 > kMaxQueryTerms || nt == 0u) {  // more terms than lanes: not supported (flagged), empty query: 0
            if (lane == 0 && warp == 0) counts[q] = nt ? ~0ull : 0ull;
            continue;
        }
        const uint32_t term = lane < nt ? qterms[t0 + lane] : 0xFFFFFFFFu;
        // drop duplicates (keep the first occurrence)
        bool keep = lane < nt;
        for (uint32_t k = 0; k + 1 < nt; ++k) {
            const uint32_t other = __shfl_sync(FULL, term, int(k));
            if (lane > k && other == term) keep = false;
        }
        const bool bad = __ballot_sync(FULL, keep && term >= n_lists) != 0u;
        if (bad) continue;  // counts[q] stays 0
        uint64_t start = 0;
        uint32_t len = 0xFFFFFFFFu;
        if (keep) {
            start = offs[term];
            len = uint32_t(min(offs[term + 1] - start, uint64_t(0xFFFFFFFEu)));
        }
        // order the kept lists by increasing length: rank of every lane's list (ties by lane)
        uint32_t rank = 0;
        for (uint32_t k = 0; k < nt; ++k) {
            const uint32_t ol = __shfl_sync(FULL, len, int(k));
            if (ol < len || (ol == len && k < lane)) ++rank;
        }
        const uint32_t n_kept = __popc(__ballot_sync(FULL, keep));
        // lane r now takes over the list of rank r
        uint64_t s_start = 0;
        uint32_t s_len = 0;
        for (uint32_t k = 0; k < nt; ++k) {
            const uint32_t kr = __shfl_sync(FULL, keep ? rank : 0xFFFFFFFFu, int(k));
            const uint64_t ks = shfl64(start, int(k));
            const uint32_t kl = __shfl_sync(FULL, len, int(k));
            if (kr == lane) {
                s_start = ks;
                s_len = kl;
            }
        }
        const uint64_t a_start = shfl64(s_start, 0);
        const uint32_t a_len = __shfl_sync(FULL, s_len, 0);
        unsigned long long total = 0;
        for (uint32_t c0 = warp * 32u; c0 < a_len; c0 += kAndWarps * 32u) {
            const bool has = c0 + lane < a_len;
            const uint32_t cand = has ? docids[a_start + c0 + lane] : 0u;
            bool alive = has;
            for (uint32_t j = 1; j < n_kept && __any_sync(FULL, alive); ++j) {
                const uint64_t bs = shfl64(s_start, int(j));
                const uint32_t bl = __shfl_sync(FULL, s_len, int(j));
                if (alive) {  // lower bound of cand in list j
                    uint32_t lo = 0, hi = bl;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (docids[bs + mid] < cand) lo = mid + 1;
                        else hi = mid;
                    }
                    alive = lo < bl && docids[bs + lo] == cand;
                }
            }
            total += __popc(__ballot_sync(FULL, alive));
        }
        if (lane == 0 && total) atomicAdd(&counts[q], total);
    }
}

// ---- single-list walk for the host-buffer convenience call ----------------------------------
__global__ void __launch_bounds__(32)
decode_list_kernel(dint_dict_view dv, uint32_t lut_bytes, int multi, const uint8_t* __restrict__ in, uint64_t in_bytes,
                   uint32_t n, uint32_t block_size, uint32_t* __restrict__ out, uint64_t* __restrict__ consumed) {
    warp_state ws;
    const class_info* classes = cta_setup(dv, lut_bytes, in, in_bytes, out, ws);
    const unsigned lane = lane_id();
    const uint64_t base = reinterpret_cast<uint64_t>(in);
    uint64_t addr = base;
    ws.flushed = 0;
    if (!multi) {
        addr = decode_one<OUT_RAW>(ws, dv.rows, classes, dv.n_classes, false, addr, n, lane);
    } else {
        for (uint32_t done = 0; done < n; done += block_size)  // vroom_env/dint_codecs.hpp:525-534
            addr = decode_one<OUT_RAW>(ws, dv.rows, classes, dv.n_classes, true, addr, min(block_size, n - done), lane);
    }
    ws.flush_all<OUT_RAW>(lane);
    if (lane == 0) *consumed = addr - base;
}

// ---- launch helpers -------------------------------------------------------------------------
}  // namespace

int dintk::get_device_info(int device, device_info* di) {
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

namespace {

// 4-bit log2(length) per dictionary row, padded to 16 bytes
uint32_t lut_bytes_of(const dint_dict* d) { return ((d->n_rows + 1u) / 2u + 15u) & ~15u; }
size_t smem_bytes_of(const dint_dict* d, int warps) {
    return size_t(lut_bytes_of(d)) + kClassBytes + size_t(warps) * kWarpSmem;
}

// persistent grid: as many CTAs as fit (shared memory bound), never more than the work needs. A dictionary whose
// 4-bit length table leaves no room for `warps` rings gets a smaller CTA (*warps is lowered; the kernels take
// their warp count from blockDim), down to one warp.
template <typename K>
int plan_grid(K kernel, const dint_dict* d, uint64_t n_units, int* grid, size_t* smem, int* warps) {
    device_info di;
    int rc = get_device_info(d->device, &di);
    if (rc) return rc;
    while (*warps > 1 && smem_bytes_of(d, *warps) > size_t(di.max_smem_optin)) --*warps;
    *smem = smem_bytes_of(d, *warps);
    if (*smem > size_t(di.max_smem_optin)) {
        dint::set_error("dictionary length table does not fit in shared memory");
        return DINT_E_ARG;
    }
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(*smem)));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, *warps * 32, *smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t want = (n_units + *warps - 1) / *warps;
    uint64_t cap = uint64_t(di.sms) * per_sm;
    *grid = int(std::max<uint64_t>(1, std::min(want, cap)));
    return DINT_OK;
}

uint32_t env_u32_k(const char* name, uint32_t dflt) {
    const char* v = std::getenv(name);
    return v && *v ? uint32_t(std::strtoul(v, nullptr, 10)) : dflt;
}

}  // namespace

extern "C" {
static int strip_launch_setup(dint_dict* d);
}

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
    // lengths are powers of two (1..256): 4 bits of log2 per row, two rows per byte
    std::vector<uint8_t> lut(lens_bytes, 0);
    for (uint32_t r = 0; r < d->n_rows; ++r) {
        uint32_t len = uint32_t(d->h_lens[r]) + 1, code = 0;
        while ((1u << code) < len) ++code;
        lut[r >> 1] |= uint8_t(code << ((r & 1u) * 4u));
    }
    CUDA_TRY(cudaMemcpy(d->d_lens, lut.data(), lens_bytes, cudaMemcpyHostToDevice));
    // the same table as one slice per class, indexed by the symbol (lane-sequential decoder). The first
    // lut_rows rows of every class are staged in shared memory: everything for a single dictionary
    // (32 KB), the 8192 most frequent rows of each class of a multi dictionary (<= 24 KB; covers every
    // u8 block, whose symbols are < 256)
    uint32_t max_rows = 0;
    for (uint32_t c = 0; c < d->n_classes; ++c) max_rows = std::max(max_rows, std::min(d->class_size[c], dint::kNumEntries));
    max_rows = std::max((max_rows + 31u) & ~31u, 32u);
    d->lut_rows = (size_t(max_rows) / 2 * d->n_classes <= 40u * 1024u) ? max_rows : 8192u;
    size_t cls_bytes = 0;
    for (uint32_t c = 0; c < d->n_classes; ++c) {
        d->lens_cls_off[c] = uint32_t(cls_bytes);
        const uint32_t rows = std::max((d->class_size[c] + 31u) & ~31u, d->lut_rows);
        cls_bytes += (size_t(rows) / 2 + 15) & ~size_t(15);
    }
    std::vector<uint8_t> lut_cls(cls_bytes + 16, 0);
    for (uint32_t c = 0; c < d->n_classes; ++c)
        for (uint32_t s = 0; s < d->class_size[c]; ++s) {
            uint32_t len = uint32_t(d->h_lens[d->class_base[c] + s]) + 1, code = 0;
            while ((1u << code) < len) ++code;
            lut_cls[d->lens_cls_off[c] + (s >> 1)] |= uint8_t(code << ((s & 1u) * 4u));
        }
    CUDA_TRY(cudaMalloc(&d->d_lens_cls, lut_cls.size()));
    CUDA_TRY(cudaMemcpy(d->d_lens_cls, lut_cls.data(), lut_cls.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&d->d_sums, size_t(d->n_rows) * 4 + 16));
    CUDA_TRY(cudaMemcpy(d->d_sums, d->h_sums.data(), size_t(d->n_rows) * 4, cudaMemcpyHostToDevice));
    // shared-memory image of the lane-sequential decoder (dint_dict_view::strip_tables). A multi dictionary
    // keeps the 256 rows of every class that u8 codewords can address (they are also the most frequent
    // ones, the builder sorts by frequency); a single dictionary keeps its first 2048 rows.
    {
        uint32_t hot = 256;
        if (d->n_classes == 1) {
            const uint32_t cap = env_u32_k("DINT_B200_HOT_ROWS", 2048);
            while (hot < cap && hot < d->class_size[0]) hot *= 2;
        }
        d->hot_rows = hot;
        const size_t nc = d->n_classes;
        const size_t hot0 = 0, hot1 = nc * hot * 16, hotlen = 2 * nc * hot * 16;
        size_t off = hotlen + nc * hot;
        for (uint32_t c = 0; c < d->n_classes; ++c) {
            off = (off + 3) & ~size_t(3);
            d->group_off[c] = uint32_t(off);
            off += (size_t(d->class_size[c]) + 31) / 32;
        }
        // group byte: log2(length) when the 32 rows agree; else 9 + k = the k-th mixed group of the class beyond
        // the hot rows (its per-row nibbles follow in the detail table), 0xFF once a class has more than 246 of them
        std::vector<std::vector<uint8_t>> detail(d->n_classes);
        std::vector<uint8_t> groups(off, 0);
        auto code_of = [&](uint32_t row) {
            uint32_t len = uint32_t(d->h_lens[row]) + 1, code = 0;
            while ((1u << code) < len) ++code;
            return code;
        };
        for (uint32_t c = 0; c < d->n_classes; ++c) {
            const uint32_t rows = d->class_size[c], base = d->class_base[c];
            for (uint32_t g = 0; g * 32 < rows; ++g) {
                const uint32_t first = code_of(base + g * 32), hi = std::min(rows, g * 32 + 32);
                bool same = true;
                for (uint32_t s = g * 32; s < hi; ++s) same = same && code_of(base + s) == first;
                uint8_t v = uint8_t(first);
                if (!same) {
                    const size_t k = detail[c].size() / 16;
                    if (g * 32 + 32 <= hot || k > 245) v = 0xFF;  // hot rows never consult the group table
                    else {
                        v = uint8_t(9 + k);
                        detail[c].resize(detail[c].size() + 16, 0);
                        for (uint32_t s = g * 32; s < hi; ++s)
                            detail[c][k * 16 + ((s & 31u) >> 1)] |= uint8_t(code_of(base + s) << ((s & 1u) * 4u));
                    }
                }
                groups[d->group_off[c] + g] = v;
            }
        }
        for (uint32_t c = 0; c < d->n_classes; ++c) {
            off = (off + 3) & ~size_t(3);
            d->detail_off[c] = uint32_t(off);
            off += detail[c].size();
        }
        const size_t total = (off + 15) & ~size_t(15);
        std::vector<uint8_t> img(total, 0);
        std::memcpy(img.data() + hotlen, groups.data() + hotlen, groups.size() - hotlen);
        for (uint32_t c = 0; c < d->n_classes; ++c) {
            if (!detail[c].empty()) std::memcpy(img.data() + d->detail_off[c], detail[c].data(), detail[c].size());
            const uint32_t rows = d->class_size[c];
            for (uint32_t s = 0; s < std::min(rows, hot); ++s) {
                const uint32_t* row = d->h_rows.data() + size_t(d->class_base[c] + s) * dint::kEntry;
                std::memcpy(img.data() + hot0 + (size_t(c) * hot + s) * 16, row, 16);
                std::memcpy(img.data() + hot1 + (size_t(c) * hot + s) * 16, row + 4, 16);
                img[hotlen + size_t(c) * hot + s] = uint8_t(code_of(d->class_base[c] + s));
            }
        }
        d->strip_tables_bytes = uint32_t(total);
        CUDA_TRY(cudaMalloc(&d->d_strip_tables, total));
        CUDA_TRY(cudaMemcpy(d->d_strip_tables, img.data(), total, cudaMemcpyHostToDevice));
    }
    return strip_launch_setup(d);
}

void dint_dict_release_device(dint_dict* d) {
    if (d->d_rows || d->d_lens || d->d_lens_cls || d->d_strip_tables || d->d_sums || d->scratch.d_in || d->scratch.d_out ||
        d->scratch.d_cons) {
        device_guard g(d->device);
        cudaFree(d->d_rows);
        cudaFree(d->d_lens);
        cudaFree(d->d_lens_cls);
        cudaFree(d->d_strip_tables);
        cudaFree(d->d_sums);
        cudaFree(d->scratch.d_in);
        cudaFree(d->scratch.d_out);
        cudaFree(d->scratch.d_cons);
        d->scratch.d_in = nullptr;
        d->scratch.d_out = nullptr;
        d->scratch.d_cons = nullptr;
        d->scratch.in_cap = d->scratch.out_cap = 0;
        d->d_sums = nullptr;
        d->d_strip_tables = nullptr;
        d->d_rows = nullptr;
        d->d_lens = nullptr;
        d->d_lens_cls = nullptr;
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
    CUDA_TRY(cudaMemcpy(&l, d->d_lens + (r >> 1), 1, cudaMemcpyDeviceToHost));
    *len = 1u << ((l >> ((r & 1u) * 4u)) & 15u);
    return DINT_OK;
}

int dint_decode_units(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_segment* d_segs, size_t n_segs,
                      const uint32_t* d_unit_first, size_t n_units, uint32_t* d_out, int mode, void* stream) {
    if (!d || mode < DINT_MODE_RAW || mode > DINT_MODE_FREQS) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (n_segs == 0) return DINT_OK;
    if (!d_in || !d_segs || !d_out || n_segs > 0xFFFFFFF0ull) return DINT_E_ARG;
    if (!d_unit_first) n_units = (n_segs + kImplicitUnit - 1) / kImplicitUnit;
    if (n_units == 0) return DINT_OK;
    device_guard g(d->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool multi = d->kind == DINT_DICT_MULTI;
    int grid = 0, warps = 0;
    size_t smem = 0;
    int rc;
    // single dictionaries: 32 KB table, two 16-warp CTAs per SM; multi_packed: ~70 KB table shared by
    // one 32-warp CTA per SM
#define LAUNCH_SEG(MU, M, W)                                                                                       \
    warps = W;                                                                                                     \
    rc = plan_grid(decode_units_kernel<MU, M, W>, d, n_units, &grid, &smem, &warps);                               \
    if (rc) return rc;                                                                                             \
    decode_units_kernel<MU, M, W><<<grid, warps * 32, smem, st>>>(d->view(), lut_bytes_of(d), d_in, in_bytes, d_segs, \
                                                                 n_segs, d_unit_first, n_units, d_out);
    if (multi) {
        if (mode == DINT_MODE_RAW) { LAUNCH_SEG(true, OUT_RAW, 32) }
        else if (mode == DINT_MODE_DOCIDS) { LAUNCH_SEG(true, OUT_DOCIDS, 32) }
        else { LAUNCH_SEG(true, OUT_FREQS, 32) }
    } else {
        if (mode == DINT_MODE_RAW) { LAUNCH_SEG(false, OUT_RAW, 16) }
        else if (mode == DINT_MODE_DOCIDS) { LAUNCH_SEG(false, OUT_DOCIDS, 16) }
        else { LAUNCH_SEG(false, OUT_FREQS, 16) }
    }
#undef LAUNCH_SEG
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

// Launch shape of the strip decoder for one dictionary: computed once, when the dictionary is uploaded
// (dint_dict_upload), so that the decode call itself does no getenv / cudaFuncSetAttribute work.
static int strip_launch_setup(dint_dict* d) {
    device_info di;
    int rc = get_device_info(d->device, &di);
    if (rc) return rc;
    const dint::strip_limits& lim = dint::get_strip_limits();
    // one persistent CTA per SM: as many warps as the shared-memory budget holds next to the length table
    // (all of the 227 KB a CTA may have: the 24th warp is worth more than the L1 left for the colder dictionary rows)
    const size_t fixed = size_t(d->strip_tables_bytes) + kStripClassBytes;
    const size_t per_warp = strip_warp_smem(lim.unit_ints, lim.unit_bytes);
    const size_t budget = std::min(size_t(env_u32_k("DINT_B200_STRIP_SMEM_KB", 227)) * 1024, size_t(di.max_smem_optin));
    d->strip_warps = 0;
    if (fixed + per_warp > budget) return DINT_OK;  // reported by dint_decode_strips
    const uint32_t max_warps = std::min(24u, std::max(1u, env_u32_k("DINT_B200_STRIP_WARPS", 24)));
    d->strip_warps = int(std::min<size_t>((budget - fixed) / per_warp, max_warps));
    d->strip_smem = fixed + size_t(d->strip_warps) * per_warp;
    d->strip_sms = di.sms;
    const char* v = std::getenv("DINT_B200_STRIP_VARIANT");
    d->strip_variant = v && *v ? int(std::strtoul(v, nullptr, 10)) & 7 : -1;
    // opt every instance in to the device's maximum once per device
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (!done[d->device]) {
        const int mx = di.max_smem_optin;
#define OPT_IN(M)                                                                                              \
    CUDA_TRY(cudaFuncSetAttribute(decode_strips_kernel<M, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)); \
    CUDA_TRY(cudaFuncSetAttribute(decode_strips_kernel<M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)); \
    CUDA_TRY(cudaFuncSetAttribute(decode_strips_kernel<M, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)); \
    CUDA_TRY(cudaFuncSetAttribute(decode_strips_kernel<M, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        OPT_IN(OUT_RAW)
        OPT_IN(OUT_DOCIDS)
        OPT_IN(OUT_FREQS)
        CUDA_TRY(cudaFuncSetAttribute(decode_strips_kernel<OUT_FREQS, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
#undef OPT_IN
        done[d->device] = true;
    }
    return DINT_OK;
}

int dint_decode_strips(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_strip_unit* d_units,
                       size_t n_units, uint32_t* d_out, int mode, void* stream) {
    if (!d || mode < DINT_MODE_RAW || mode > DINT_MODE_FREQS) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (n_units == 0) return DINT_OK;
    if (!d_in || !d_units || !d_out || (reinterpret_cast<uintptr_t>(d_units) & 15)) return DINT_E_ARG;
    if (d->strip_warps < 1) {
        dint::set_error("strip decoder: shared-memory budget too small for one warp");
        return DINT_E_ARG;
    }
    device_guard g(d->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dint::strip_limits& lim = dint::get_strip_limits();
    strip_params P;
    P.dv = d->view();
    P.in = d_in;
    P.in_bytes = in_bytes;
    P.units = d_units;
    P.n_units = n_units;
    P.out = d_out;
    P.unit_ints = lim.unit_ints;
    P.unit_bytes = lim.unit_bytes;
    P.image_words = strip_image_words(lim.unit_ints);
    P.bits_words = strip_bits_words(lim.unit_ints);
    P.inbuf_bytes = strip_inbuf_bytes(lim.unit_bytes);
    const int warps = d->strip_warps;
    const size_t smem = d->strip_smem;
    const int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((n_units + warps - 1) / warps, uint64_t(d->strip_sms))));
    // kernel variant (see decode_strips_kernel): d-gap streams 0, frequency streams 3; DINT_B200_STRIP_VARIANT overrides
    // (variant 7 = 3 + bulk-copy flush: frequency streams into a 16-byte aligned output)
    int variant = d->strip_variant >= 0 ? d->strip_variant : (mode == DINT_MODE_FREQS ? kDefaultFreqsVariant : 0);
    if (variant > 3 && (mode != DINT_MODE_FREQS || variant != 7 || (reinterpret_cast<uintptr_t>(d_out) & 15))) variant &= 3;
    if (variant == 7) {
        decode_strips_kernel<OUT_FREQS, 7><<<grid, warps * 32, smem, st>>>(P);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
        return DINT_OK;
    }
#define LAUNCH_STRIPS_V(M, V) decode_strips_kernel<M, V><<<grid, warps * 32, smem, st>>>(P);
#define LAUNCH_STRIPS(M)                                                                   \
    if (variant == 0) { LAUNCH_STRIPS_V(M, 0) } else if (variant == 1) { LAUNCH_STRIPS_V(M, 1) } \
    else if (variant == 2) { LAUNCH_STRIPS_V(M, 2) } else { LAUNCH_STRIPS_V(M, 3) }
    if (mode == DINT_MODE_RAW) { LAUNCH_STRIPS(OUT_RAW) }
    else if (mode == DINT_MODE_DOCIDS) { LAUNCH_STRIPS(OUT_DOCIDS) }
    else { LAUNCH_STRIPS(OUT_FREQS) }
#undef LAUNCH_STRIPS_V
#undef LAUNCH_STRIPS
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

int dint_decode_segments(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, const dint_segment* d_segs,
                         size_t n_segs, uint32_t block_size, uint32_t* d_out, int mode, void* stream) {
    // Every segment of a multi_packed list already IS one block (the plan cut it with this block size), so the
    // value only has to be sane: a segment never holds more integers than a block.
    if (block_size > 65536u) {
        dint::set_error("dint_decode_segments: block_size out of range");
        return DINT_E_ARG;
    }
    return dint_decode_units(d, d_in, in_bytes, d_segs, n_segs, nullptr, 0, d_out, mode, stream);
}

static int launch_index_tails(int device, const uint8_t* d_lists, size_t lists_bytes, const dint_block* d_blocks,
                              size_t n_blocks, uint32_t block_size, uint32_t* d_docids, uint32_t* d_freqs, cudaStream_t st) {
    device_info di;
    int rc = get_device_info(device, &di);
    if (rc) return rc;
    if (!n_blocks) return DINT_OK;
    // one warp per chunk of the table; enough chunks to fill the GPU (32 resident warps per SM) before they grow
    const uint64_t warps = uint64_t(di.sms) * 32;
    uint32_t chunk = uint32_t(std::min<uint64_t>(kTailChunk, std::max<uint64_t>(32, (n_blocks / warps) & ~uint64_t(31))));
    const uint64_t n_chunks = (n_blocks + chunk - 1) / chunk;
    const uint64_t want = (n_chunks + kTailThreads / 32 - 1) / (kTailThreads / 32);
    int tgrid = int(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(di.sms) * 8)));
    index_tails_kernel<<<tgrid, kTailThreads, 0, st>>>(d_lists, lists_bytes, d_blocks, n_blocks, block_size, chunk, d_docids,
                                                      d_freqs);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

// the tile decoder's staging ring and the tail kernel's explicit stack are sized for blocks of up to 256 postings
static bool index_block_size_ok(uint32_t block_size) {
    if (block_size >= 1 && block_size <= 256) return true;
    dint::set_error("dint_decode_index: block_size must be in 1..256");
    return false;
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
    if (!d_lists || !d_blocks || (!d_docids && !d_freqs) || (reinterpret_cast<uintptr_t>(d_blocks) & 15)) return DINT_E_ARG;
    if (!block_size) block_size = dint::kDefaultBlock;
    if (!index_block_size_ok(block_size)) return DINT_E_ARG;
    device_guard g(docs_dict->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int multi = docs_dict->kind == DINT_DICT_MULTI;
    int grid = 0, warps = kWarpsPerCta;
    size_t smem = 0;
    int rc;
    if (d_docids) {
        rc = plan_grid(index_docs_kernel<OUT_DOCIDS>, docs_dict, n_blocks, &grid, &smem, &warps);
        if (rc) return rc;
        index_docs_kernel<OUT_DOCIDS><<<grid, warps * 32, smem, st>>>(docs_dict->view(), lut_bytes_of(docs_dict), multi,
                                                                     d_lists, lists_bytes, d_blocks, n_blocks, block_size,
                                                                     d_docids, d_freqs);
    } else {
        rc = plan_grid(index_docs_kernel<OUT_SKIP>, docs_dict, n_blocks, &grid, &smem, &warps);
        if (rc) return rc;
        index_docs_kernel<OUT_SKIP><<<grid, warps * 32, smem, st>>>(docs_dict->view(), lut_bytes_of(docs_dict), multi,
                                                                   d_lists, lists_bytes, d_blocks, n_blocks, block_size,
                                                                   nullptr, d_freqs);
    }
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    if (d_freqs) {
        warps = kWarpsPerCta;
        rc = plan_grid(index_freqs_kernel, freqs_dict, n_blocks, &grid, &smem, &warps);
        if (rc) return rc;
        index_freqs_kernel<<<grid, warps * 32, smem, st>>>(freqs_dict->view(), lut_bytes_of(freqs_dict), multi, d_lists,
                                                          lists_bytes, d_blocks, n_blocks, block_size, d_freqs);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
    }
    return launch_index_tails(docs_dict->device, d_lists, lists_bytes, d_blocks, n_blocks, block_size, d_docids, d_freqs, st);
}

int dint_decode_index_planned(const dint_dict* docs_dict, const dint_dict* freqs_dict, const uint8_t* d_lists,
                              size_t lists_bytes, const dint_device_plan* docs_plan, const dint_device_plan* freqs_plan,
                              const dint_block* d_blocks, size_t n_blocks, uint32_t block_size, uint32_t* d_docids,
                              uint32_t* d_freqs, void* stream) {
    if (!docs_dict || !freqs_dict || !docs_plan || !freqs_plan) return DINT_E_ARG;
    if (n_blocks == 0) return DINT_OK;
    if (!d_lists || !d_blocks || (!d_docids && !d_freqs) || (reinterpret_cast<uintptr_t>(d_blocks) & 15)) return DINT_E_ARG;
    if (!block_size) block_size = dint::kDefaultBlock;
    if (!index_block_size_ok(block_size)) return DINT_E_ARG;
    int rc;
    // full blocks: the lane-sequential decoder over the strip tables dint_index_discover built
    if (d_docids && docs_plan->n_units) {
        rc = dint_decode_strips(docs_dict, d_lists, lists_bytes, docs_plan->d_units, docs_plan->n_units, d_docids,
                                DINT_MODE_DOCIDS, stream);
        if (rc) return rc;
    }
    if (d_freqs && freqs_plan->n_units) {
        rc = dint_decode_strips(freqs_dict, d_lists, lists_bytes, freqs_plan->d_units, freqs_plan->n_units, d_freqs,
                                DINT_MODE_FREQS, stream);
        if (rc) return rc;
    }
    // partial blocks: interpolative
    device_guard g(docs_dict->device);
    return launch_index_tails(docs_dict->device, d_lists, lists_bytes, d_blocks, n_blocks, block_size, d_docids, d_freqs,
                              static_cast<cudaStream_t>(stream));
}

int dint_decode_index_blocks(int kind, const uint8_t* d_lists, size_t lists_bytes, const dint_block* d_blocks, size_t n_blocks,
                             uint32_t block_size, uint32_t* d_docids, uint32_t* d_freqs, void* stream) {
    if (kind != DINT_INDEX_BLOCK_INTERPOLATIVE && kind != DINT_INDEX_BLOCK_VBYTE && kind != DINT_INDEX_BLOCK_U32) {
        dint::set_error("dint_decode_index_blocks: kind is not one of the dictionary-less block codecs");
        return DINT_E_ARG;
    }
    if (n_blocks == 0) return DINT_OK;
    if (!d_lists || !d_blocks || (!d_docids && !d_freqs) || (reinterpret_cast<uintptr_t>(d_blocks) & 15)) return DINT_E_ARG;
    if (!block_size) block_size = dint::kDefaultBlock;
    if (!index_block_size_ok(block_size)) return DINT_E_ARG;
    cudaPointerAttributes pa;
    CUDA_TRY(cudaPointerGetAttributes(&pa, d_lists));
    if (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged) {
        dint::set_error("dint_decode_index_blocks: d_lists is not device memory (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    device_guard g(pa.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // every block is interpolative: the kernel's "partial block" threshold is put above the largest block
    if (kind == DINT_INDEX_BLOCK_INTERPOLATIVE)
        return launch_index_tails(pa.device, d_lists, lists_bytes, d_blocks, n_blocks, block_size + 1u, d_docids, d_freqs, st);
    // vbyte_block / u32_block: one warp per block, 8 warps per CTA
    device_info di;
    int rc = get_device_info(pa.device, &di);
    if (rc) return rc;
    const int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((n_blocks + 7) / 8, uint64_t(di.sms) * 8)));
    if (kind == DINT_INDEX_BLOCK_VBYTE)
        index_plain_blocks_kernel<PLAIN_VBYTE><<<grid, 256, 0, st>>>(d_lists, lists_bytes, d_blocks, n_blocks, d_docids, d_freqs);
    else
        index_plain_blocks_kernel<PLAIN_U32><<<grid, 256, 0, st>>>(d_lists, lists_bytes, d_blocks, n_blocks, d_docids, d_freqs);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

int dint_decode_index_interpolative(const uint8_t* d_lists, size_t lists_bytes, const dint_block* d_blocks, size_t n_blocks,
                                    uint32_t block_size, uint32_t* d_docids, uint32_t* d_freqs, void* stream) {
    return dint_decode_index_blocks(DINT_INDEX_BLOCK_INTERPOLATIVE, d_lists, lists_bytes, d_blocks, n_blocks, block_size, d_docids,
                                    d_freqs, stream);
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
    // persistent grid over value tiles; partial sums of lists that straddle tiles arrive by atomics
    CUDA_TRY(cudaMemsetAsync(d_sums, 0, n_lists * sizeof(uint64_t), static_cast<cudaStream_t>(stream)));
    int grid = di.sms * 8;
    checksum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_values, d_list_offsets, n_lists, reinterpret_cast<unsigned long long*>(d_sums));
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

int dint_and_queries(const uint32_t* d_docids, const uint64_t* d_list_offsets, size_t n_lists, const uint64_t* d_query_offsets,
                     const uint32_t* d_query_terms, size_t n_queries, uint64_t* d_counts, void* stream) {
    if (n_queries == 0) return DINT_OK;
    if (!d_docids || !d_list_offsets || !d_query_offsets || !d_query_terms || !d_counts) return DINT_E_ARG;
    cudaPointerAttributes pa;
    CUDA_TRY(cudaPointerGetAttributes(&pa, d_docids));
    if (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged) {
        dint::set_error("dint_and_queries: d_docids is not device memory (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    device_guard g(pa.device);
    device_info di;
    int rc = get_device_info(pa.device, &di);
    if (rc) return rc;
    // one CTA of kAndWarps warps per query; the warps add their partial counts with atomics
    CUDA_TRY(cudaMemsetAsync(d_counts, 0, n_queries * sizeof(uint64_t), static_cast<cudaStream_t>(stream)));
    const int grid = int(std::max<uint64_t>(1, std::min<uint64_t>(n_queries, uint64_t(di.sms) * 8)));
    and_queries_kernel<<<grid, kAndWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        d_docids, d_list_offsets, n_lists, d_query_offsets, d_query_terms, n_queries,
        reinterpret_cast<unsigned long long*>(d_counts));
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return DINT_OK;
}

struct dint_vroom_ctx {
    const dint_dict* dict = nullptr;
    uint32_t block_size = 0;
    uint64_t n_lists = 0, n_ints = 0, n_units = 0, in_bytes = 0;
    uint8_t* d_in = nullptr;
    const dint_strip_unit* d_units = nullptr;
    const uint64_t* d_list_offsets = nullptr;
    uint32_t* d_out = nullptr;
    uint64_t* d_sums = nullptr;
    // exactly one of the two owns the tables above
    dint_strip_unit* own_units = nullptr;
    uint64_t* own_list_offsets = nullptr;
    dint_device_plan* plan = nullptr;  // dint_vroom_ctx_create_discover
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
    c->n_units = plan->strip_units.size();
    c->in_bytes = plan->list_byte_offsets.back();
    auto fail = [&](cudaError_t e) {
        dint::set_error(std::string("dint_vroom_ctx_create: ") + cudaGetErrorString(e));
        dint_vroom_ctx_destroy(c);
        return DINT_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_in, c->in_bytes + 16)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->own_units, (c->n_units + 1) * sizeof(dint_strip_unit))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->own_list_offsets, (c->n_lists + 1) * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_out, (c->n_ints + 1) * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&c->d_sums, (c->n_lists + 1) * 8)) != cudaSuccess) return fail(e);
    if (c->n_units && (e = cudaMemcpy(c->own_units, plan->strip_units.data(), c->n_units * sizeof(dint_strip_unit),
                                      cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail(e);
    if ((e = cudaMemcpy(c->own_list_offsets, plan->list_offsets.data(), (c->n_lists + 1) * 8, cudaMemcpyHostToDevice)) !=
        cudaSuccess)
        return fail(e);
    c->d_units = c->own_units;
    c->d_list_offsets = c->own_list_offsets;
    *out = c;
    return DINT_OK;
}

int dint_vroom_ctx_create_discover(const dint_dict* d, const void* h_bin, size_t len, uint32_t block_size,
                                   const uint64_t* sync_offsets, size_t n_sync, uint64_t sync_stride, void* stream,
                                   dint_vroom_ctx** out) {
    if (!d || !out || (!h_bin && len)) return DINT_E_ARG;
    if (!d->d_rows) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    device_guard g(d->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dint_vroom_ctx* c = new (std::nothrow) dint_vroom_ctx;
    if (!c) return DINT_E_NOMEM;
    c->dict = d;
    c->block_size = block_size ? block_size : dint::kDefaultBlock;
    c->in_bytes = len;
    auto fail = [&](int rc, cudaError_t e) {
        if (e != cudaSuccess) dint::set_error(std::string("dint_vroom_ctx_create_discover: ") + cudaGetErrorString(e));
        dint_vroom_ctx_destroy(c);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_in, len + 16)) != cudaSuccess) return fail(DINT_E_CUDA, e);
    if (len && (e = cudaMemcpyAsync(c->d_in, h_bin, len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(DINT_E_CUDA, e);
    int rc = dint_vroom_discover(d, c->d_in, len, c->block_size, sync_offsets, n_sync, sync_stride, stream, &c->plan);
    if (rc) return fail(rc, cudaSuccess);
    c->n_lists = c->plan->n_lists;
    c->n_ints = c->plan->n_ints;
    c->n_units = c->plan->n_units;
    c->d_units = c->plan->d_units;
    c->d_list_offsets = c->plan->d_list_offsets;
    if ((e = cudaMalloc(&c->d_out, (c->n_ints + 1) * 4)) != cudaSuccess) return fail(DINT_E_CUDA, e);
    if ((e = cudaMalloc(&c->d_sums, (c->n_lists + 1) * 8)) != cudaSuccess) return fail(DINT_E_CUDA, e);
    *out = c;
    return DINT_OK;
}

void dint_vroom_ctx_destroy(dint_vroom_ctx* c) {
    if (!c) return;
    {
        device_guard g(c->dict->device);
        cudaFree(c->d_in);
        cudaFree(c->own_units);
        cudaFree(c->own_list_offsets);
        cudaFree(c->d_out);
        cudaFree(c->d_sums);
    }
    dint_device_plan_destroy(c->plan);
    delete c;
}

uint32_t* dint_vroom_ctx_device_output(const dint_vroom_ctx* c) { return c ? c->d_out : nullptr; }
uint64_t dint_vroom_ctx_num_lists(const dint_vroom_ctx* c) { return c ? c->n_lists : 0; }
uint64_t dint_vroom_ctx_num_ints(const dint_vroom_ctx* c) { return c ? c->n_ints : 0; }
const dint_device_plan* dint_vroom_ctx_device_plan(const dint_vroom_ctx* c) { return c ? c->plan : nullptr; }

// checksums / decoded integers of the last decode back to the host
static int ctx_results(dint_vroom_ctx* c, uint32_t* h_out, uint64_t* h_sums, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h_sums && c->n_lists) {
        int rc = dint_checksum_lists(c->d_out, c->d_list_offsets, c->n_lists, c->d_sums, stream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_sums, c->d_sums, c->n_lists * 8, cudaMemcpyDeviceToHost, st));
    }
    if (h_out && c->n_ints) CUDA_TRY(cudaMemcpyAsync(h_out, c->d_out, c->n_ints * 4, cudaMemcpyDeviceToHost, st));
    return DINT_OK;
}

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
    int rc = dint_decode_strips(c->dict, c->d_in, len, c->d_units, c->n_units, c->d_out, mode, stream);
    if (rc) return rc;
    return ctx_results(c, h_out, h_sums, stream);
}

int dint_vroom_decode_resident(dint_vroom_ctx* c, int mode, void* stream) {
    if (!c) return DINT_E_ARG;
    return dint_decode_strips(c->dict, c->d_in, c->in_bytes, c->d_units, c->n_units, c->d_out, mode, stream);
}

int dint_vroom_ctx_results(dint_vroom_ctx* c, uint32_t* h_out, uint64_t* h_sums, void* stream) {
    if (!c) return DINT_E_ARG;
    device_guard g(c->dict->device);
    return ctx_results(c, h_out, h_sums, stream);
}

int dint_vroom_ctx_sync(dint_vroom_ctx* c, void* stream) {
    if (!c) return DINT_E_ARG;
    device_guard g(c->dict->device);
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return DINT_OK;
}

// Scratch of dint_decode_list_host, kept with the dictionary and grown on demand: after the first call of a
// given size the call allocates nothing (it is the literal Decoder::decode drop-in, called once per list).
static int list_scratch(const dint_dict* d, size_t in_bytes, size_t out_ints) {
    dint_dict::host_scratch& s = d->scratch;
    if (s.in_cap < in_bytes) {
        cudaFree(s.d_in);
        s.d_in = nullptr;
        s.in_cap = 0;
        const size_t cap = std::max<size_t>(in_bytes * 2, 1u << 16);
        CUDA_TRY(cudaMalloc(&s.d_in, cap));
        s.in_cap = cap;
    }
    if (s.out_cap < out_ints) {
        cudaFree(s.d_out);
        s.d_out = nullptr;
        s.out_cap = 0;
        const size_t cap = std::max<size_t>(out_ints * 2, 1u << 14);
        CUDA_TRY(cudaMalloc(&s.d_out, cap * 4));
        s.out_cap = cap;
    }
    if (!s.d_cons) CUDA_TRY(cudaMalloc(&s.d_cons, 8));
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
    std::lock_guard<std::mutex> lk(d->scratch.mu);
    int rc = list_scratch(d, nbytes + 16, n);
    if (rc) return rc;
    dint_dict::host_scratch& s = d->scratch;
    CUDA_TRY(cudaMemcpy(s.d_in, in, nbytes, cudaMemcpyHostToDevice));
    int warps = 1;
    const size_t smem = smem_bytes_of(d, warps);
    if (!s.attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(decode_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        s.attr_set = true;
    }
    decode_list_kernel<<<1, 32, smem>>>(d->view(), lut_bytes_of(d), d->kind == DINT_DICT_MULTI, s.d_in, nbytes, n, block_size,
                                        s.d_out, s.d_cons);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    uint64_t cons = 0;
    CUDA_TRY(cudaMemcpy(out, s.d_out, size_t(n) * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&cons, s.d_cons, 8, cudaMemcpyDeviceToHost));
    if (consumed) *consumed = size_t(cons);
    return DINT_OK;
}

}  // extern "C"
