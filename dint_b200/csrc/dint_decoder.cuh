// dint_b200 — the warp-level DINT stream decoder (device code, sm_100a).
//
// One warp decodes a RUN of up to 32 consecutive segments (whole short lists, ~1 K-integer
// chunks of long lists, 256-integer blocks of multi_packed lists — include/dint_b200.h) whose
// compressed bytes are (nearly) adjacent and whose outputs are contiguous in the packed output
// array. The codeword symbols of the run form one FLAT symbol space; the warp walks it tile by
// tile, so tiles stay full even when a list holds a single posting:
//
//   input ring   1 KB of shared memory per warp, filled 512 bytes at a time with 16-byte
//                cp.async (LDGSTS) from the compressed stream; the copy for the next chunk
//                is in flight while the current one is parsed. Ring offsets are the low bits
//                of the global byte address, so any stream alignment works.
//   tile         32 consecutive flat symbols (u16, or u8 in multi_packed's narrow blocks), one
//                per lane, always starting on a codeword boundary. While the tile lies inside
//                one segment its parameters are warp-uniform; when a segment starts inside
//                the tile, a __reduce_or_sync mask of the segment starts + popc tells every
//                lane which segment its symbol belongs to (lane k holds segment k's
//                parameters, fetched by shuffle).
//                  * __ballot_sync masks of the symbols equal to 0 / 1 decide which symbols are
//                    codewords and which are exception payload. Candidates that no earlier
//                    candidate can reach are codewords for sure and are resolved bit-parallel;
//                    only real exception codewords hiding behind such payload cost one
//                    warp-uniform loop trip each;
//                  * entry lengths come from a 4-bit log2(length) table in shared memory
//                    (32 KB for a single dictionary);
//                  * a __shfl_up_sync inclusive scan of the lengths gives every codeword its
//                    output offset; codewords that would overflow the staging ring are left
//                    for the next tile.
//   scatter      every codeword lane copies its entry from the L2-resident dictionary (64-byte
//                aligned rows, 16-byte vector loads) / its exception value into the warp's
//                OUTPUT STAGING RING (512 ints of shared memory, indexed by the low bits of
//                the global output index). Runs of zeros cost nothing: the ring is kept
//                zeroed beyond the staged data.
//   flush        whenever a 128-integer aligned group of the output is complete the warp reads it
//                back as one uint4 per lane, applies the fused transform — docID prefix sum
//                (lane-local + one warp scan per 128 integers) or the freq +1 shift — and
//                writes 512 contiguous bytes with 16-byte streaming stores. Groups straddle
//                segment and list boundaries; at a list boundary the d-gap prefix sum is
//                re-based by folding (next doc_base - previous list's last docid) into the
//                first staged value, so the scan itself never has to be segmented.
//
// The kernel is bound by instruction issue, not by HBM (profiles/): the hot loop is written to
// keep the per-tile instruction count down — 32-bit shared-memory addresses, predicated stores,
// shuffle-with-predicate scans, warp-uniform control flow.
//
// Reference semantics reproduced (paths relative to the reference root):
//   vroom_env/dint_codecs.hpp:37-107 (u16 loop), :574-609 (u8 loop), :521-536 (selector),
//   include/dint/dint_codecs.hpp:13-49, 460-510, include/dint/*_dictionary.hpp copy().
#pragma once

#include <cuda_runtime.h>

#include "dint_internal.hpp"

namespace dintk {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kStageInts = 512;  // output staging ring per warp (power of two)
constexpr int kInBytes = 1024;   // input ring per warp (two chunks)
constexpr int kChunk = 512;      // bytes per cp.async round (32 lanes x 16 bytes)
constexpr int kGroup = 128;      // integers per flush round (one uint4 per lane)
constexpr int kQueue = 64;        // codeword queue entries per warp (power of two, 8 bytes each)
constexpr int kWarpSmem = kStageInts * 4 + kInBytes + kQueue * 8;
constexpr uint32_t kMaxRunBytes = 1u << 30;  // input span of one run (32-bit relative offsets)
constexpr uint32_t kUnknownLen = 0xFFFFFFFFu;
constexpr uint32_t kNoSegment = 0xFFFFFFFFu;
constexpr uint32_t kTileLook = 80;  // bytes a single-segment tile may touch: 32 u16 symbols + payload, rounded up

enum { OUT_RAW = 0, OUT_DOCIDS = 1, OUT_FREQS = 2, OUT_SKIP = 3 };

__device__ __forceinline__ unsigned lane_id() {
    unsigned l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ unsigned lanemask_le() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, uint64_t src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(uint32_t* p, uint4 v) { __stcs(reinterpret_cast<uint4*>(p), v); }

// shared-memory accesses through 32-bit addresses (no generic-pointer arithmetic in the hot loop)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// words 1..3 of an entry of length L (1, 2 or >= 4) to consecutive ring slots
__device__ __forceinline__ void sts_tail3(uint32_t a, uint32_t L, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile(
        "{\n\t.reg .pred p1, p2;\n\t"
        "setp.gt.u32 p1, %1, 1;\n\t"
        "setp.gt.u32 p2, %1, 2;\n\t"
        "@p1 st.shared.u32 [%0+4], %2;\n\t"
        "@p2 st.shared.u32 [%0+8], %3;\n\t"
        "@p2 st.shared.u32 [%0+12], %4;\n\t}" ::"r"(a),
        "r"(L), "r"(y), "r"(z), "r"(w)
        : "memory");
}
// the same for an entry of 2^k words (k = 0: nothing, 1: word 1, >= 2: words 1..3)
__device__ __forceinline__ void sts_tail3_log2(uint32_t a, uint32_t k, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile(
        "{\n\t.reg .pred p1, p2;\n\t"
        "setp.gt.u32 p1, %1, 0;\n\t"
        "setp.gt.u32 p2, %1, 1;\n\t"
        "@p1 st.shared.u32 [%0+4], %2;\n\t"
        "@p2 st.shared.u32 [%0+8], %3;\n\t"
        "@p2 st.shared.u32 [%0+12], %4;\n\t}" ::"r"(a),
        "r"(k), "r"(y), "r"(z), "r"(w)
        : "memory");
}
__device__ __forceinline__ void sts_4(uint32_t a, uint4 v) {
    asm volatile(
        "st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+4], %2;\n\tst.shared.u32 [%0+8], %3;\n\tst.shared.u32 [%0+12], %4;" ::"r"(a),
        "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
        : "memory");
}

// little-endian loads at any alignment from the file image (aligned 8-byte loads + funnel)
__device__ __forceinline__ uint64_t ldu64(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), a8 = a & ~uintptr_t(7);
    const uint32_t sh = uint32_t(a & 7) * 8u;
    const uint64_t lo = __ldg(reinterpret_cast<const uint64_t*>(a8));
    if (!sh) return lo;
    const uint64_t hi = __ldg(reinterpret_cast<const uint64_t*>(a8 + 8));
    return (lo >> sh) | (hi << (64u - sh));
}
__device__ __forceinline__ uint32_t ldu32(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), a4 = a & ~uintptr_t(3);
    const uint32_t sh = uint32_t(a & 3) * 8u;
    const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(a4));
    if (!sh) return lo;
    return __funnelshift_r(lo, __ldg(reinterpret_cast<const uint32_t*>(a4 + 4)), sh);
}
// inclusive warp scan: shfl.up hands back a predicate that is false for the lanes without a source
#define DINT_SCAN_STEP(v, o)                                                        \
    asm volatile(                                                                   \
        "{\n\t.reg .pred p;\n\t.reg .u32 t;\n\t"                                    \
        "shfl.sync.up.b32 t|p, %0, " #o ", 0, 0xffffffff;\n\t"                      \
        "@p add.u32 %0, %0, t;\n\t}"                                                \
        : "+r"(v))
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
    DINT_SCAN_STEP(v, 1);
    DINT_SCAN_STEP(v, 2);
    DINT_SCAN_STEP(v, 4);
    DINT_SCAN_STEP(v, 8);
    DINT_SCAN_STEP(v, 16);
    return v;
}
#undef DINT_SCAN_STEP
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(FULL, uint32_t(v), src), hi = __shfl_sync(FULL, uint32_t(v >> 32), src);
    return (uint64_t(hi) << 32) | lo;
}
// tells the compiler a warp-uniform value is uniform (a REDUX result lives in a uniform register),
// so that branches on it need no divergence guards around the warp-level instructions inside
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __reduce_max_sync(FULL, v); }

// Per-warp decoder state. Lives in registers; everything is warp-uniform.
struct warp_state {
    uint32_t stage_s;   // shared address of the output staging ring [kStageInts]
    uint32_t inb_s;     // shared address of the input ring [kInBytes]
    uint32_t lut_s;     // shared address of the 4-bit length table
    uint32_t que_s;     // shared address of the codeword queue [kQueue] x {meta, value}
    uint32_t* out;      // packed output array (global)
    uint64_t flushed;   // output index: everything below it (of this run) is in global memory
    uint32_t staged;    // integers sitting in the ring: [flushed, flushed + staged)
    uint32_t acc;       // DOCIDS: docid of element flushed-1 (doc_base - 1 at a list start)
    bool vec_ok;        // output base is 16-byte aligned
    // input: global byte address = anchor + rel; anchor is a multiple of kInBytes so that the
    // ring offset of an address is rel & (kInBytes - 1)
    uint64_t anchor;
    uint32_t win_end;   // ring holds rel addresses [win_end - kInBytes, win_end); 0 = nothing
    uint32_t safe_end;  // ... of which everything below safe_end has landed
    uint32_t fast_end;  // single-segment tiles starting below this rel address need no housekeeping
    uint32_t lo_rel;    // 16-byte aligned bounds of the bytes this warp may read, rel
    uint32_t hi_rel;
    uint64_t in_lo;     // the same bounds as absolute addresses (whole input buffer)
    uint64_t in_hi;

    // Re-anchors the input side at global address g (start of a run); returns g's rel offset.
    // Bytes at or beyond `end` (absolute) are not fetched.
    __device__ __forceinline__ uint32_t set_anchor(uint64_t g, uint64_t end) {
        cp_async_wait_all();
        __syncwarp();
        anchor = g & ~uint64_t(kInBytes - 1);
        lo_rel = in_lo > anchor ? uint32_t(in_lo - anchor) : 0u;
        const uint64_t hi = end >= in_hi ? in_hi : (end + 15) & ~uint64_t(15);
        hi_rel = uint32_t(min(hi > anchor ? hi - anchor : uint64_t(0), uint64_t(0xFFFFFFF0u)));
        win_end = 0;
        safe_end = 0;
        fast_end = 0;
        return uint32_t(g - anchor);
    }
    __device__ __forceinline__ void load_chunk(uint32_t base, unsigned lane) {
        const uint32_t x = base + 16u * lane;
        const uint32_t so = inb_s + (x & (kInBytes - 1));
        if (x >= lo_rel && x < hi_rel) cp_async16(so, anchor + x);
        else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(so), "r"(0u) : "memory");
    }
    // Input housekeeping: make rel addresses [a, a + look) readable from the ring (look <= kChunk).
    __device__ __forceinline__ void refill(uint32_t a, uint32_t look, unsigned lane) {
        if (win_end == 0 || a + kInBytes < win_end || a >= win_end) {
            cp_async_wait_all();
            __syncwarp();
            const uint32_t b = a & ~uint32_t(kChunk - 1);
            load_chunk(b, lane);
            load_chunk(b + kChunk, lane);
            win_end = b + 2 * kChunk;
            safe_end = b;
        } else if (a + kChunk > win_end) {
            // the warp has moved into the newest chunk: the other slot is dead, refill it
            load_chunk(win_end, lane);
            win_end += kChunk;
        }
        if (a + look > safe_end) {
            cp_async_wait_all();
            __syncwarp();
            safe_end = win_end;
        }
        // a single-segment tile at a' < fast_end has [a', a' + kTileLook) landed and is not yet in
        // the newest chunk
        const uint32_t f = min(safe_end - kTileLook, win_end - uint32_t(kChunk));
        fast_end = safe_end > kTileLook ? f + 1u : 0u;
    }
    __device__ __forceinline__ void ensure_input(uint32_t a, uint32_t look, unsigned lane) {
        if (look > kTileLook || a >= fast_end) refill(a, look, lane);
    }
    __device__ __forceinline__ uint32_t rb(uint32_t a) const { return lds_u8(inb_s + (a & (kInBytes - 1))); }
    // the little-endian u32 at ring address a (any alignment)
    __device__ __forceinline__ uint32_t r32(uint32_t a) const {
        const uint32_t w0 = lds_u32(inb_s + (a & (kInBytes - 4)));
        const uint32_t w1 = lds_u32(inb_s + ((a + 4) & (kInBytes - 4)));
        return __funnelshift_r(w0, w1, (a & 3u) * 8u);
    }

    // ---- output side ---------------------------------------------------------------------
    // Flushes the (possibly partial) 128-aligned group that starts the staged data.
    template <int MODE>
    __device__ __forceinline__ void flush_group(unsigned lane) {
        const uint32_t lo = uint32_t(flushed) & (kGroup - 1);
        const uint32_t hi = min(uint32_t(kGroup), lo + staged);
        const uint64_t gbase = flushed - lo;
        const uint32_t q = 4u * lane;
        const uint32_t sp = stage_s + (((uint32_t(gbase) + q) & (kStageInts - 1)) << 2);
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sp));
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sp), "r"(0u) : "memory");  // keep the ring zeroed
        uint32_t* gp = out + gbase + q;
        if (lo == 0 && hi == kGroup && vec_ok) {
            if (MODE == OUT_DOCIDS) {
                v.x += 1u;
                v.y += v.x + 1u;
                v.z += v.y + 1u;
                v.w += v.z + 1u;
                const uint32_t incl = warp_inclusive_scan(v.w);
                const uint32_t base = acc + (incl - v.w);
                acc += __shfl_sync(FULL, incl, 31);
                v.x += base;
                v.y += base;
                v.z += base;
                v.w += base;
            } else if (MODE == OUT_FREQS) {
                v.x += 1u;
                v.y += 1u;
                v.z += 1u;
                v.w += 1u;
            }
            st_stream4(gp, v);
        } else {
            const bool l0 = q >= lo && q < hi, l1 = q + 1 >= lo && q + 1 < hi, l2 = q + 2 >= lo && q + 2 < hi,
                       l3 = q + 3 >= lo && q + 3 < hi;
            if (MODE == OUT_DOCIDS) {
                v.x = l0 ? v.x + 1u : 0u;
                v.y = v.x + (l1 ? v.y + 1u : 0u);
                v.z = v.y + (l2 ? v.z + 1u : 0u);
                v.w = v.z + (l3 ? v.w + 1u : 0u);
                const uint32_t incl = warp_inclusive_scan(v.w);
                const uint32_t base = acc + (incl - v.w);
                acc += __shfl_sync(FULL, incl, 31);
                v.x += base;
                v.y += base;
                v.z += base;
                v.w += base;
            } else if (MODE == OUT_FREQS) {
                v.x += 1u;
                v.y += 1u;
                v.z += 1u;
                v.w += 1u;
            }
            if (l0) st_stream(gp, v.x);
            if (l1) st_stream(gp + 1, v.y);
            if (l2) st_stream(gp + 2, v.z);
            if (l3) st_stream(gp + 3, v.w);
        }
        flushed += hi - lo;
        staged -= hi - lo;
    }
    template <int MODE>
    __device__ __forceinline__ void flush_full(unsigned lane) {
        while ((uint32_t(flushed) & (kGroup - 1)) + staged >= uint32_t(kGroup)) flush_group<MODE>(lane);
    }
    template <int MODE>
    __device__ __forceinline__ void flush_all(unsigned lane) {
        while (staged) flush_group<MODE>(lane);
    }
};

// What lane k knows about segment k of the run.
struct seg_lane {
    uint32_t cs;    // first flat symbol of the segment
    uint32_t ab;    // rel address of its first symbol
    uint32_t pk;    // row_base | narrow << 31   (narrow: u8 symbols, multi_packed selector >= 6)
    uint32_t rows;  // addressable rows of its class
    uint32_t corr;  // DOCIDS: added to the segment's first integer (re-bases the running prefix sum)
};

// Which of a tile's symbols are exception payload. Z / O: symbols equal to 0 / 1; a 0 codeword is
// followed by PZ payload symbols (a u16 value), a 1 codeword by PO (a u32 value); payload may
// itself look like 0 / 1. `starts`: symbols known to be codewords (bit 0, segment starts).
template <int PZ, int PO>
__device__ __forceinline__ uint32_t payload_mask(uint32_t Z, uint32_t O, uint32_t starts) {
    constexpr uint32_t MZ = (1u << PZ) - 1u, MO = (1u << PO) - 1u;
    const uint32_t E = Z | O;
    uint32_t reach = 0;
#pragma unroll
    for (int d = 1; d <= PZ; ++d) reach |= Z << d;
#pragma unroll
    for (int d = 1; d <= PO; ++d) reach |= O << d;
    const uint32_t A = E & (~reach | starts);  // candidates no other candidate can reach: codewords
    const uint32_t az = A & Z, ao = A & O;
    uint32_t P = 0;
#pragma unroll
    for (int d = 1; d <= PZ; ++d) P |= az << d;
#pragma unroll
    for (int d = 1; d <= PO; ++d) P |= ao << d;
    P &= ~starts;
    uint32_t c = E & ~A & ~P;
    while (c) {  // warp-uniform: a real exception codeword right behind 0/1-valued payload
        const uint32_t bit = c & (0u - c);
        const uint32_t m = (Z & bit) ? MZ : MO;
        P |= (m * bit * 2u) & ~starts;
        c = E & ~A & ~P & ~(bit * 2u - 1u);
    }
    return P;
}

// Decodes a run of `cnt` (<= 32) segments whose parameters sit in lanes 0..cnt-1 of `sl`.
//   S  flat symbols in the run (kUnknownLen: a single segment of unknown byte length; the walk
//      then stops after exactly N integers, like the reference's `i != n` loop)
//   N  integers the run produces (never exceeded, whatever the bytes say)
// Appends to the staging ring. Returns the flat symbol position one past the last consumed symbol.
template <bool MULTI, int MODE>
__device__ __forceinline__ uint32_t decode_run(warp_state& ws, const uint32_t* __restrict__ rows, const seg_lane& sl,
                                               uint32_t cnt_, uint32_t S_, uint32_t N_, unsigned lane) {
    const uint32_t cnt = uniform(cnt_), S = uniform(S_);
    uint32_t F = 0, remaining = uniform(N_);
    const bool has = lane < cnt;
    // parameters of the segment the tile starts in (segment ck), broadcast to every lane
    uint32_t ck = 0;
    uint32_t c_cs = 0, c_ab = uniform(__shfl_sync(FULL, sl.ab, 0));
    uint32_t c_pk = uniform(__shfl_sync(FULL, sl.pk, 0)), c_rows = uniform(__shfl_sync(FULL, sl.rows, 0));
    uint32_t next_start = cnt > 1 ? uniform(__shfl_sync(FULL, sl.cs, 1)) : kNoSegment;  // first symbol of segment ck+1
    while (remaining && F < S) {
        const uint32_t szsh = (MULTI && (c_pk >> 31)) ? 0u : 1u;  // log2(bytes per symbol), uniform per tile
        uint32_t nlive = min(S - F, 32u);                         // symbols of the run in this tile
        uint32_t H = 0, kl = ck;
        uint32_t my_cs = c_cs, my_pk = c_pk, my_rows = c_rows;
        uint32_t a, sym;
        if (next_start - F >= 32u) {
            // ---- the whole tile lies in segment ck -----------------------------------------------------
            const uint32_t a0 = c_ab + ((F - c_cs) << szsh);
            ws.ensure_input(a0, kTileLook, lane);
            a = a0 + (lane << szsh);
            if (szsh == 0) sym = ws.rb(a);
            else if ((a0 & 1u) == 0) sym = lds_u16(ws.inb_s + (a & (kInBytes - 1)));
            else sym = ws.rb(a) | (ws.rb(a + 1) << 8);
        } else {
            // ---- segments start inside the tile: which one does each lane's symbol belong to -----------
            const uint32_t rel = sl.cs - F;
            H = __reduce_or_sync(FULL, (has && rel - 1u < 31u) ? (1u << rel) : 0u);
            kl = ck + __popc(H & lanemask_le());
            my_cs = __shfl_sync(FULL, sl.cs, kl);
            my_pk = __shfl_sync(FULL, sl.pk, kl);
            my_rows = __shfl_sync(FULL, sl.rows, kl);
            const uint32_t my_ab = __shfl_sync(FULL, sl.ab, kl);
            if (MULTI) {
                // one symbol size per tile: stop in front of the first segment of the other size
                const uint32_t other = __ballot_sync(FULL, (my_pk >> 31) != (c_pk >> 31));
                if (other) nlive = min(nlive, uint32_t(__ffs(other) - 1));
            }
            a = my_ab + ((F + lane - my_cs) << szsh);
            // bytes the tile can touch: from lane 0's symbol to lane 31's symbol + the widest payload
            const uint32_t a0 = __shfl_sync(FULL, a, 0), a31 = __shfl_sync(FULL, a, 31);
            ws.ensure_input(a0, max(kTileLook + 1u, min(a31 - a0 + 8u, uint32_t(kChunk))), lane);
            sym = ws.rb(a);
            if (szsh) sym |= ws.rb(a + 1) << 8;
        }

        // ---- which symbols are codewords ----------------------------------------------------
        const bool inlive = lane < nlive;
        const uint32_t Z = __ballot_sync(FULL, inlive && sym == 0u);
        const uint32_t O = __ballot_sync(FULL, inlive && sym == 1u);
        bool iscw = inlive;
        if (Z | O) {
            const uint32_t P = szsh ? payload_mask<1, 2>(Z, O, H | 1u) : payload_mask<2, 4>(Z, O, H | 1u);
            iscw = inlive && !((P >> lane) & 1u);
        }

        // ---- lengths, offsets, how much of the tile is taken ----------------------------------
        // (rows 0 and 1 of every class are the exception codewords: the table holds length 1 for them)
        const uint32_t row = (my_pk & 0x7FFFFFFFu) + min(sym, my_rows - 1u);
        const uint32_t nib = (lds_u8(ws.lut_s + (row >> 1)) >> ((row & 1u) * 4u)) & 15u;
        const uint32_t len = iscw ? (1u << nib) : 0u;
        const uint32_t e = warp_inclusive_scan(len);
        uint32_t limit = remaining;
        if (MODE != OUT_SKIP) limit = min(limit, kStageInts - (uint32_t(ws.flushed) & (kGroup - 1)) - ws.staged);
        const bool valid = iscw && e <= limit;
        const uint32_t T = __reduce_max_sync(FULL, valid ? e : 0u);
        if (T == 0) break;  // malformed stream (a codeword straddles the end of the run)
        // symbols consumed: up to and including the last taken codeword and its payload
        const uint32_t step = sym < 2u ? 1u + ((4u >> (sym == 0u ? 1 : 0)) >> szsh) : 1u;
        const uint32_t adv = __reduce_max_sync(FULL, valid ? lane + step : 0u);

        // ---- scatter into the staging ring ------------------------------------------------------
        if (MODE != OUT_SKIP) {
            const uint32_t idx = (uint32_t(ws.flushed) + ws.staged + e - len) & (kStageInts - 1);
            const uint32_t sa = ws.stage_s + (idx << 2);
            const uint32_t L = (valid && sym >= dint::kReserved) ? len : 0u;  // words to copy from the dictionary
            const uint4* rp = reinterpret_cast<const uint4*>(rows + size_t(row) * dint::kEntry);
            uint4 r0 = make_uint4(0, 0, 0, 0);
            if (L) r0 = __ldg(rp);
            if (valid && sym < 2u) {  // exception: the value follows the codeword (u16 after 0, u32 after 1)
                r0.x = ws.r32(a + (1u << szsh));
                if (sym == 0u) r0.x &= 0xFFFFu;
            }
            if (MODE == OUT_DOCIDS && (H || F == c_cs)) {  // a segment starts in this tile: re-base
                const uint32_t my_corr = __shfl_sync(FULL, sl.corr, kl);
                if (F + lane == my_cs) r0.x += my_corr;
            }
            const bool wraps = idx + 16u > uint32_t(kStageInts);  // the entry may wrap around the ring
            if (valid) sts_u32(sa, r0.x);
            if (!wraps) sts_tail3(sa, L, r0.y, r0.z, r0.w);
            if (__any_sync(FULL, L > 4)) {
                if (L > 4 && !wraps) sts_4(sa + 16, __ldg(rp + 1));
                if (__any_sync(FULL, L > 8)) {
                    if (L > 8 && !wraps) {
                        sts_4(sa + 32, __ldg(rp + 2));
                        sts_4(sa + 48, __ldg(rp + 3));
                    }
                }
            }
            if (wraps && L > 1) {
                const uint32_t* rw = rows + size_t(row) * dint::kEntry;
                for (uint32_t j = 1; j < L; ++j) sts_u32(ws.stage_s + (((idx + j) & (kStageInts - 1)) << 2), __ldg(rw + j));
            }
            ws.staged += T;
            __syncwarp();
            if ((uint32_t(ws.flushed) & (kGroup - 1)) + ws.staged >= uint32_t(kGroup)) {
                ws.flush_full<MODE>(lane);
                __syncwarp();
            }
        }
        remaining -= T;
        F += adv;
        if (F >= next_start) {
            // the next tile starts in another segment
            ck = __popc(__ballot_sync(FULL, has && sl.cs <= F)) - 1u;
            c_cs = uniform(__shfl_sync(FULL, sl.cs, ck));
            c_ab = uniform(__shfl_sync(FULL, sl.ab, ck));
            c_pk = uniform(__shfl_sync(FULL, sl.pk, ck));
            c_rows = uniform(__shfl_sync(FULL, sl.rows, ck));
            next_start = ck + 1 < cnt ? uniform(__shfl_sync(FULL, sl.cs, (ck + 1) & 31)) : kNoSegment;
        }
    }
    return F;
}

// ---- two-phase run decoder: parse tiles into a codeword queue, expand 32 codewords at a time -----
// In sparse posting lists most symbols are exception payload (a 32-bit exception takes 3 u16 or 5 u8
// symbols), so a 32-symbol tile holds only 6..10 codewords and the per-tile work of decode_run
// (length lookup, scan, scatter, flush test) runs with most lanes idle. Here the run is parsed in
// phase A (symbols -> codewords, tile by tile, no output state involved) into a small per-warp
// queue of {dictionary row | kind, value}, and phase B expands FULL batches of 32 queued codewords
// into the staging ring. Requires the run's symbol count S (segment byte lengths), so it serves
// the vroom work units; segments of unknown length go through decode_run.
constexpr uint32_t kMetaExc = 0x80000000u;  // queue meta: exception (value = the integer itself)
constexpr uint32_t kMetaRun = 0x40000000u;  // queue meta: run of zeros (nothing to copy)

template <int MODE>
__device__ __forceinline__ bool expand_batch(warp_state& ws, const uint32_t* __restrict__ rows, uint32_t& qhead,
                                             uint32_t qtail, uint32_t& remaining, unsigned lane) {
    const uint32_t qn = qtail - qhead;
    const bool has = lane < qn;
    const uint32_t qa = ws.que_s + (((qhead + lane) & (kQueue - 1)) << 3);
    uint32_t meta = 0, val = 0;
    if (has) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(meta), "=r"(val) : "r"(qa));
    const uint32_t row = meta & 0x3FFFFFFFu;
    const uint32_t nib = (lds_u8(ws.lut_s + (row >> 1)) >> ((row & 1u) * 4u)) & 15u;
    const uint32_t len = has ? (1u << nib) : 0u;
    const uint32_t e = warp_inclusive_scan(len);
    const uint32_t limit = min(remaining, kStageInts - (uint32_t(ws.flushed) & (kGroup - 1)) - ws.staged);
    const bool valid = has && e <= limit;
    const uint32_t V = __ballot_sync(FULL, valid);
    if (V == 0) return false;  // malformed stream: more integers than the run declares
    const uint32_t T = __shfl_sync(FULL, e, 31 - __clz(V));
    const uint32_t idx = (uint32_t(ws.flushed) + ws.staged + e - len) & (kStageInts - 1);
    const uint32_t sa = ws.stage_s + (idx << 2);
    const uint32_t L = (valid && !(meta & (kMetaExc | kMetaRun))) ? len : 0u;  // words to copy from the dictionary
    const uint4* rp = reinterpret_cast<const uint4*>(rows + size_t(row) * dint::kEntry);
    uint4 r0 = make_uint4(0, 0, 0, 0);
    if (L) r0 = __ldg(rp);
    r0.x += val;  // exception value, or the docID re-basing term of a segment's first codeword, or 0
    const bool wraps = idx + 16u > uint32_t(kStageInts);  // the entry may wrap around the ring
    if (valid) sts_u32(sa, r0.x);
    if (!wraps) sts_tail3(sa, L, r0.y, r0.z, r0.w);
    if (__any_sync(FULL, L > 4)) {
        if (L > 4 && !wraps) sts_4(sa + 16, __ldg(rp + 1));
        if (__any_sync(FULL, L > 8)) {
            if (L > 8 && !wraps) {
                sts_4(sa + 32, __ldg(rp + 2));
                sts_4(sa + 48, __ldg(rp + 3));
            }
        }
    }
    if (wraps && L > 1) {
        const uint32_t* rw = rows + size_t(row) * dint::kEntry;
#pragma unroll 1
        for (uint32_t j = 1; j < L; ++j) sts_u32(ws.stage_s + (((idx + j) & (kStageInts - 1)) << 2), __ldg(rw + j));
    }
    ws.staged += T;
    remaining -= T;
    qhead += __popc(V);
    __syncwarp();
    if ((uint32_t(ws.flushed) & (kGroup - 1)) + ws.staged >= uint32_t(kGroup)) {
        ws.flush_full<MODE>(lane);
        __syncwarp();
    }
    return true;
}

template <bool MULTI, int MODE>
__device__ __forceinline__ void decode_run_queued(warp_state& ws, const uint32_t* __restrict__ rows, const seg_lane& sl,
                                                  uint32_t cnt_, uint32_t S_, uint32_t N_, unsigned lane) {
    const uint32_t cnt = uniform(cnt_), S = uniform(S_);
    uint32_t F = 0, remaining = uniform(N_);
    uint32_t qhead = 0, qtail = 0;
    const bool has = lane < cnt;
    const unsigned le = lanemask_le();
    uint32_t ck = 0;
    uint32_t c_cs = 0, c_ab = uniform(__shfl_sync(FULL, sl.ab, 0));
    uint32_t c_pk = uniform(__shfl_sync(FULL, sl.pk, 0)), c_rows = uniform(__shfl_sync(FULL, sl.rows, 0));
    uint32_t next_start = cnt > 1 ? uniform(__shfl_sync(FULL, sl.cs, 1)) : kNoSegment;
    while (F < S && remaining) {
        // ======== phase A: one tile of symbols -> codewords appended to the queue ========================
        const uint32_t szsh = (MULTI && (c_pk >> 31)) ? 0u : 1u;
        uint32_t nlive = min(S - F, 32u);
        uint32_t H = 0, kl = ck;
        uint32_t my_cs = c_cs, my_pk = c_pk, my_rows = c_rows;
        uint32_t a, sym;
        if (next_start - F >= 32u) {
            const uint32_t a0 = c_ab + ((F - c_cs) << szsh);
            ws.ensure_input(a0, kTileLook, lane);
            a = a0 + (lane << szsh);
            if (szsh == 0) sym = ws.rb(a);
            else if ((a0 & 1u) == 0) sym = lds_u16(ws.inb_s + (a & (kInBytes - 1)));
            else sym = ws.rb(a) | (ws.rb(a + 1) << 8);
        } else {
            const uint32_t rel = sl.cs - F;
            H = __reduce_or_sync(FULL, (has && rel - 1u < 31u) ? (1u << rel) : 0u);
            kl = ck + __popc(H & le);
            my_cs = __shfl_sync(FULL, sl.cs, kl);
            my_pk = __shfl_sync(FULL, sl.pk, kl);
            my_rows = __shfl_sync(FULL, sl.rows, kl);
            const uint32_t my_ab = __shfl_sync(FULL, sl.ab, kl);
            if (MULTI) {
                const uint32_t other = __ballot_sync(FULL, (my_pk >> 31) != (c_pk >> 31));
                if (other) nlive = min(nlive, uint32_t(__ffs(other) - 1));
            }
            a = my_ab + ((F + lane - my_cs) << szsh);
            const uint32_t a0 = __shfl_sync(FULL, a, 0), a31 = __shfl_sync(FULL, a, 31);
            ws.ensure_input(a0, max(kTileLook + 1u, min(a31 - a0 + 8u, uint32_t(kChunk))), lane);
            sym = ws.rb(a);
            if (szsh) sym |= ws.rb(a + 1) << 8;
        }
        const bool inlive = lane < nlive;
        const uint32_t Z = __ballot_sync(FULL, inlive && sym == 0u);
        const uint32_t O = __ballot_sync(FULL, inlive && sym == 1u);
        uint32_t CW = nlive >= 32u ? FULL : ((1u << nlive) - 1u);
        if (Z | O) CW &= ~(szsh ? payload_mask<1, 2>(Z, O, H | 1u) : payload_mask<2, 4>(Z, O, H | 1u));
        const bool iscw = (CW >> lane) & 1u;
        // queue entry of this lane's codeword
        uint32_t meta = (my_pk & 0x3FFFFFFFu) + min(sym, my_rows - 1u);
        uint32_t val = 0, step = 1;
        if (sym < dint::kReserved) meta |= kMetaRun;
        if (iscw && sym < 2u) {  // exception: the value follows the codeword (u16 after 0, u32 after 1)
            val = ws.r32(a + (1u << szsh));
            if (sym == 0u) val &= 0xFFFFu;
            meta |= kMetaExc;
            step = 1u + ((4u >> (sym == 0u ? 1 : 0)) >> szsh);
        }
        if (MODE == OUT_DOCIDS && (H || F == c_cs)) {  // a segment starts in this tile: re-base its first integer
            const uint32_t my_corr = __shfl_sync(FULL, sl.corr, kl);
            if (F + lane == my_cs) val += my_corr;
        }
        if (iscw) {
            const uint32_t slot = (qtail + __popc(CW & (le >> 1))) & (kQueue - 1);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ws.que_s + (slot << 3)), "r"(meta), "r"(val) : "memory");
        }
        qtail += __popc(CW);
        F += __reduce_max_sync(FULL, iscw ? lane + step : 0u);  // symbols consumed, payload included
        if (F >= next_start) {
            ck = __popc(__ballot_sync(FULL, has && sl.cs <= F)) - 1u;
            c_cs = uniform(__shfl_sync(FULL, sl.cs, ck));
            c_ab = uniform(__shfl_sync(FULL, sl.ab, ck));
            c_pk = uniform(__shfl_sync(FULL, sl.pk, ck));
            c_rows = uniform(__shfl_sync(FULL, sl.rows, ck));
            next_start = ck + 1 < cnt ? uniform(__shfl_sync(FULL, sl.cs, (ck + 1) & 31)) : kNoSegment;
        }
        // ======== phase B: expand full batches ============================================================
        __syncwarp();
        while (qtail - qhead >= 32u && remaining) {
            if (!expand_batch<MODE>(ws, rows, qhead, qtail, remaining, lane)) {
                remaining = 0;
                break;
            }
        }
    }
    __syncwarp();
    while (qtail != qhead && remaining) {  // drain
        if (!expand_batch<MODE>(ws, rows, qhead, qtail, remaining, lane)) break;
    }
}

// Per-class parameters of a dictionary, staged in shared memory next to the length table.
struct class_info {
    uint32_t row_base;
    uint32_t rows;
};

// One segment of unknown byte length that starts at global address g and produces n integers
// (a dict_freq_index block, or a whole list for the host convenience call). single kinds: u16
// stream at g. multi: selector byte at g, then a u16 (selector < 6) or u8 (selector - 6)
// stream into class selector % 6 (vroom_env/dint_codecs.hpp:536-575,
// include/dint/dint_codecs.hpp:466-505). Returns the global address one past the last consumed
// byte (the reference's returned pointer).
template <int MODE>
__device__ __forceinline__ uint64_t decode_one(warp_state& ws, const uint32_t* __restrict__ rows,
                                               const class_info* __restrict__ classes, uint32_t n_classes, bool multi,
                                               uint64_t g, uint32_t n, unsigned lane) {
    if (n == 0) return g;
    uint32_t a = ws.set_anchor(g, ~uint64_t(0));
    seg_lane sl;
    sl.cs = 0;
    sl.corr = 0;
    if (!multi) {
        sl.ab = a;
        sl.pk = 0;
        sl.rows = classes[0].rows;
        const uint32_t F = decode_run<false, MODE>(ws, rows, sl, 1, kUnknownLen, n, lane);
        return ws.anchor + a + 2ull * F;
    }
    ws.ensure_input(a, 1, lane);
    const uint32_t sel = ws.rb(a);
    const bool wide = sel < dint::kSelectors;
    const class_info ci = classes[min(wide ? sel : sel - dint::kSelectors, n_classes - 1u)];
    sl.ab = a + 1;
    sl.pk = ci.row_base | (wide ? 0u : 0x80000000u);
    sl.rows = wide ? ci.rows : min(ci.rows, 256u);
    const uint32_t F = decode_run<true, MODE>(ws, rows, sl, 1, kUnknownLen, n, lane);
    return ws.anchor + a + 1ull + (wide ? 2ull : 1ull) * F;
}

}  // namespace dintk
