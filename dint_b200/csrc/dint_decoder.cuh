// dint_b200 — the warp-level DINT stream decoder (device code, sm_100a).
//
// One warp owns a WORK UNIT: a run of consecutive segments (whole short lists and/or
// ~1 K-integer chunks of long lists, include/dint_b200.h) whose outputs are contiguous in the
// packed output array. The warp streams through the unit tile by tile:
//
//   input ring   1 KB of shared memory per warp, filled 512 bytes at a time with 16-byte
//                cp.async (LDGSTS) from the compressed stream; the copy for the next chunk
//                is in flight while the current one is parsed. Ring offsets are the low bits
//                of the global byte address, so any stream alignment works.
//   tile         32 consecutive codeword symbols (u16, or u8 in multi_packed's narrow blocks),
//                one per lane, always starting on a codeword boundary:
//                  * __ballot_sync masks of the symbols equal to 0 / 1 decide which symbols are
//                    codewords and which are exception payload. Candidates with no other
//                    candidate in the preceding 2 (u16) / 4 (u8) symbols are codewords for
//                    sure and are resolved bit-parallel; only the rest (payload that happens
//                    to equal 0/1) is walked in stream order, warp-uniformly;
//                  * entry lengths come from a 4-bit log2(length) table in shared memory
//                    (32 KB for a single dictionary);
//                  * a __shfl_up_sync inclusive scan of the lengths gives every codeword its
//                    output offset; codewords that would overflow the staging ring or the
//                    segment are left for the next tile.
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
constexpr int kWarpSmem = kStageInts * 4 + kInBytes;

enum { OUT_RAW = 0, OUT_DOCIDS = 1, OUT_FREQS = 2, OUT_SKIP = 3 };

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
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, uint64_t src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 ldg_row(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(uint32_t* p, uint4 v) { __stcs(reinterpret_cast<uint4*>(p), v); }

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, unsigned lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, o);
        if (lane >= unsigned(o)) v += t;
    }
    return v;
}

// Per-warp decoder state. Lives in registers; everything is warp-uniform.
struct warp_state {
    uint32_t* stage;    // [kStageInts] shared
    uint8_t* inb;       // [kInBytes] shared
    uint32_t* out;      // packed output array (global)
    uint64_t flushed;   // output index: everything below it (of this unit) is in global memory
    uint32_t staged;    // integers sitting in the ring: [flushed, flushed + staged)
    uint32_t acc;       // DOCIDS: docid of element flushed-1 (doc_base - 1 at a list start)
    uint64_t win_end;   // input ring holds global addresses [win_end - kInBytes, win_end)
    uint64_t safe_end;  // ... of which everything below safe_end has landed
    uint64_t in_lo;     // 16-byte aligned bounds of the bytes this warp may read
    uint64_t in_hi;
    bool vec_ok;        // output base is 16-byte aligned

    __device__ __forceinline__ void load_chunk(uint64_t base, unsigned lane) {
        const uint64_t x = base + 16u * lane;
        const uint32_t so = smem_u32(inb) + (uint32_t(x) & (kInBytes - 1));
        if (x >= in_lo && x < in_hi) cp_async16(so, x);
        else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(so), "r"(0u) : "memory");
    }
    // make [addr, addr + look) readable from the ring
    __device__ __forceinline__ void ensure_input(uint64_t addr, uint32_t look, unsigned lane) {
        if (win_end == 0 || addr + kInBytes < win_end || addr >= win_end) {
            cp_async_wait_all();
            __syncwarp();
            const uint64_t b = addr & ~uint64_t(kChunk - 1);
            load_chunk(b, lane);
            load_chunk(b + kChunk, lane);
            win_end = b + 2 * kChunk;
            safe_end = b;
        } else if (addr + kChunk > win_end) {
            // the warp has moved into the newest chunk: the other slot is dead, refill it
            load_chunk(win_end, lane);
            win_end += kChunk;
        }
        if (addr + look > safe_end) {
            cp_async_wait_all();
            __syncwarp();
            safe_end = win_end;
        }
    }
    __device__ __forceinline__ uint32_t rb(uint32_t a) const { return inb[a & (kInBytes - 1)]; }

    // ---- output side ---------------------------------------------------------------------
    // Flushes the (possibly partial) 128-aligned group that starts the staged data.
    template <int MODE>
    __device__ __forceinline__ void flush_group(unsigned lane) {
        const uint32_t lo = uint32_t(flushed) & (kGroup - 1);
        const uint32_t hi = min(uint32_t(kGroup), lo + staged);
        const uint64_t gbase = flushed - lo;
        const uint32_t q = 4u * lane;
        uint32_t* sp = stage + ((uint32_t(gbase) + q) & (kStageInts - 1));
        uint4 v = *reinterpret_cast<uint4*>(sp);
        *reinterpret_cast<uint4*>(sp) = make_uint4(0, 0, 0, 0);  // keep the ring zeroed behind us
        uint32_t* gp = out + gbase + q;
        if (lo == 0 && hi == kGroup && vec_ok) {
            if (MODE == OUT_DOCIDS) {
                v.x += 1u;
                v.y += v.x + 1u;
                v.z += v.y + 1u;
                v.w += v.z + 1u;
                const uint32_t incl = warp_inclusive_scan(v.w, lane);
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
                const uint32_t incl = warp_inclusive_scan(v.w, lane);
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

// Decodes the codeword stream that starts at global address `addr` and produces exactly n
// integers, appending them to the staging ring. Returns the address one past the last consumed
// symbol (the reference's returned pointer). SYM = bytes per codeword symbol.
//   lut       4-bit log2(length) table, indexed by dictionary row
//   row_base  first row of the stream's class, class_rows its number of rows
//   corr      DOCIDS re-basing term folded into the first integer (0 for a fresh ring)
template <int SYM, int MODE>
__device__ __forceinline__ uint64_t decode_tiles(warp_state& ws, const uint32_t* __restrict__ rows,
                                                 const uint8_t* __restrict__ lut, uint32_t row_base, uint32_t class_rows,
                                                 uint64_t addr, uint32_t n, uint32_t corr, unsigned lane) {
    constexpr int PZ = 2 / SYM;  // payload symbols after a 0 codeword (u16 value)
    constexpr int PO = 4 / SYM;  // payload symbols after a 1 codeword (u32 value)
    constexpr uint32_t LOOK = 32 * SYM + 4;
    uint32_t remaining = n;
    while (remaining) {
        ws.ensure_input(addr, LOOK, lane);
        const uint32_t a = uint32_t(addr) + SYM * lane;
        uint32_t sym;
        if (SYM == 1) sym = ws.rb(a);
        else if ((uint32_t(addr) & 1u) == 0) sym = *reinterpret_cast<const uint16_t*>(ws.inb + (a & (kInBytes - 1)));
        else sym = ws.rb(a) | (ws.rb(a + 1) << 8);

        // ---- which symbols are codewords ----------------------------------------------------
        const uint32_t Z = __ballot_sync(FULL, sym == 0u);
        const uint32_t O = __ballot_sync(FULL, sym == 1u);
        const uint32_t E = Z | O;
        uint32_t P = 0;  // payload symbols of this tile
        if (E) {
            uint32_t prev = 0;
#pragma unroll
            for (int d = 1; d <= PO; ++d) prev |= E << d;
            const uint32_t A = E & ~prev;  // certainly codewords
            const uint32_t az = A & Z, ao = A & O;
#pragma unroll
            for (int d = 1; d <= PZ; ++d) P |= az << d;
#pragma unroll
            for (int d = 1; d <= PO; ++d) P |= ao << d;
            uint32_t U = E & ~A;
            while (U) {  // warp-uniform: payload that happens to look like an exception codeword
                const int j = __ffs(U) - 1;
                U &= U - 1;
                if (!((P >> j) & 1u)) {
                    const uint32_t cnt = ((Z >> j) & 1u) ? PZ : PO;
                    P |= (((1u << cnt) - 1u) << j) << 1;
                }
            }
        }
        const bool iscw = !((P >> lane) & 1u);

        // ---- lengths, offsets, how much of the tile is taken ----------------------------------
        uint32_t len = iscw ? 1u : 0u;
        uint32_t row = 0;
        if (iscw && sym >= 2u) {
            row = row_base + min(sym, class_rows - 1u);
            len = 1u << ((uint32_t(lut[row >> 1]) >> ((row & 1u) * 4u)) & 15u);
        }
        const uint32_t e = warp_inclusive_scan(len, lane);
        const uint32_t s = e - len;
        const uint32_t room = kStageInts - ((uint32_t(ws.flushed) & (kGroup - 1)) + ws.staged);
        const uint32_t limit = MODE == OUT_SKIP ? remaining : min(remaining, room);
        const bool valid = iscw && e <= limit;
        const uint32_t V = __ballot_sync(FULL, valid);
        if (V == 0) break;  // malformed stream (a codeword straddles the end of the segment)
        const int last = 31 - __clz(V);
        const uint32_t T = __shfl_sync(FULL, e, last);
        const uint32_t lastsym = __shfl_sync(FULL, sym, last);
        const uint32_t extra = lastsym == 0u ? PZ : (lastsym == 1u ? PO : 0);

        // ---- scatter into the staging ring ------------------------------------------------------
        if (MODE != OUT_SKIP) {
            const uint32_t base = uint32_t(ws.flushed) + ws.staged + s;
            uint32_t* st = ws.stage;
            const uint32_t L = (valid && sym >= dint::kReserved) ? len : 0u;
            const uint4* rp = reinterpret_cast<const uint4*>(rows + size_t(row) * dint::kEntry);
            uint4 r0 = make_uint4(0, 0, 0, 0);
            if (L) r0 = ldg_row(rp);
            if (valid && sym < 2u) {
                const uint32_t pa = a + SYM;
                r0.x = ws.rb(pa) | (ws.rb(pa + 1) << 8);
                if (sym == 1u) r0.x |= (ws.rb(pa + 2) << 16) | (ws.rb(pa + 3) << 24);
            }
            if (valid) {
                if (s == 0) r0.x += corr;
                st[base & (kStageInts - 1)] = r0.x;
            }
            if (L > 1) st[(base + 1) & (kStageInts - 1)] = r0.y;
            if (L > 2) {
                st[(base + 2) & (kStageInts - 1)] = r0.z;
                st[(base + 3) & (kStageInts - 1)] = r0.w;
            }
            if (__any_sync(FULL, L > 4)) {
                if (L > 4) {
                    const uint4 r1 = ldg_row(rp + 1);
                    st[(base + 4) & (kStageInts - 1)] = r1.x;
                    st[(base + 5) & (kStageInts - 1)] = r1.y;
                    st[(base + 6) & (kStageInts - 1)] = r1.z;
                    st[(base + 7) & (kStageInts - 1)] = r1.w;
                }
                if (__any_sync(FULL, L > 8)) {
                    if (L > 8) {
                        const uint4 r2 = ldg_row(rp + 2);
                        const uint4 r3 = ldg_row(rp + 3);
                        st[(base + 8) & (kStageInts - 1)] = r2.x;
                        st[(base + 9) & (kStageInts - 1)] = r2.y;
                        st[(base + 10) & (kStageInts - 1)] = r2.z;
                        st[(base + 11) & (kStageInts - 1)] = r2.w;
                        st[(base + 12) & (kStageInts - 1)] = r3.x;
                        st[(base + 13) & (kStageInts - 1)] = r3.y;
                        st[(base + 14) & (kStageInts - 1)] = r3.z;
                        st[(base + 15) & (kStageInts - 1)] = r3.w;
                    }
                }
            }
            corr = 0;
            ws.staged += T;
            __syncwarp();
            ws.flush_full<MODE>(lane);
            __syncwarp();
        }
        remaining -= T;
        addr += uint64_t(SYM) * (last + 1 + extra);
    }
    return addr;
}

// One segment. single kinds: u16 stream at `addr`. multi: selector byte at `addr`, then a u16
// (selector < 6) or u8 (selector - 6) stream into class selector % 6
// (vroom_env/dint_codecs.hpp:536-575, include/dint/dint_codecs.hpp:466-505).
template <int MODE>
__device__ __forceinline__ uint64_t decode_one(warp_state& ws, const dint_dict_view& dv, const uint8_t* __restrict__ lut,
                                               bool multi, uint64_t addr, uint32_t n, uint32_t corr, unsigned lane) {
    if (n == 0) return addr;
    if (!multi) return decode_tiles<2, MODE>(ws, dv.rows, lut, 0, dv.n_rows, addr, n, corr, lane);
    ws.ensure_input(addr, 1, lane);
    const uint32_t sel = ws.rb(uint32_t(addr));
    const bool wide = sel < dint::kSelectors;
    const uint32_t cls = min(wide ? sel : sel - dint::kSelectors, dv.n_classes - 1u);
    const uint32_t base = dv.class_base[cls];
    const uint32_t crows = (cls + 1 < dv.n_classes ? dv.class_base[cls + 1] : dv.n_rows) - base;
    if (wide) return decode_tiles<2, MODE>(ws, dv.rows, lut, base, crows, addr + 1, n, corr, lane);
    return decode_tiles<1, MODE>(ws, dv.rows, lut, base, min(crows, 256u), addr + 1, n, corr, lane);
}

}  // namespace dintk
