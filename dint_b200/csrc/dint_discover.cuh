// dint_b200 — device-side discovery of the decode restart points (device code, sm_100a).
//
// The lane-sequential decoder (dint_strips.cuh) needs to know, every few codewords, where the
// reference's two cursors (input byte, output index) stand, plus the running docID. In round 1 a
// serial host walk of every codeword produced that table (dint_vroom_plan_create). Here the GPU
// finds the same restart points itself, from the compressed bytes already in HBM:
//
//   walk_stream   one warp walks ONE codeword stream (a multi_packed block, a whole single-dictionary
//                 list, a full block of a frozen index) 32 symbols per step: __ballot_sync masks of
//                 the symbols equal to 0 / 1 separate codewords from exception payload (3-state
//                 scan, payload_mask), entry lengths come from the shared-memory length tables of
//                 the strip decoder, a __shfl_up_sync inclusive scan gives every codeword its output
//                 offset and a second one its docID prefix sum(v + 1) — no dictionary entry is
//                 copied. Every 8th codeword (and every crossing of a 512-integer boundary, for
//                 streams with run codewords) is a strip start.
//   unit_packer   packs consecutive strips into dint_strip_unit records with the rules of the host
//                 builder (<= 32 strips, unit_ints integers, unit_bytes bytes, contiguous output).
//   vroom_walk_kernel    one warp per CHAIN of a vroom .bin: lists are found one after the other as
//                 in vroom_env/decode.cpp:139-150 (header::read, then the stream ends where the
//                 codewords have produced n integers). Chains start at sync points: list starts the
//                 caller knows (shard / file boundaries) or that find_sync_kernel found. A count pass
//                 sizes the tables, an emit pass fills them, both on the device.
//   find_sync_kernel     speculative search for list starts near evenly spaced targets of a big file:
//                 a candidate byte position is accepted when it parses as `depth` consecutive lists
//                 whose decoded values sum to the header's `universe` field (jobs.hpp:76-85 writes
//                 universe = sum of the values). A wrong guess cannot corrupt anything: the walker of
//                 the previous chain must LAND exactly on every sync point, else the point is dropped
//                 and the walk repeated.
//   index_walk_kernel    full blocks of a frozen dict_freq_index: block starts come from the file's
//                 own endpoints (no walk at all on the host), docs streams first (their end is where
//                 the freqs stream starts, dict_posting_list.hpp:311-318), then freqs streams.
//
// Reference semantics followed: vroom_env/dint_codecs.hpp:37-107, 521-619; vroom_env/codecs.hpp:93-124;
// include/dint/dint_codecs.hpp:13-49, 460-510.
#pragma once

#include "dint_strips.cuh"

namespace dintk {

// per-class constants of the length tables (32 bytes, shared memory)
struct walk_class {
    uint32_t hotlen_s;   // shared address: one byte of log2(length) per hot row
    uint32_t group_s;    // shared address: one byte per 32 rows
    uint32_t detail_s;   // shared address: per-row nibbles of the mixed groups
    uint32_t rows;       // rows of the class (>= 1)
    uint32_t row_base;   // first dictionary row of the class
    uint32_t lut_g_off;  // byte offset of the class's slice in dint_dict_view::lens_cls
    uint32_t pad0, pad1;
};

constexpr uint32_t kWalkClassBytes = 8u * uint32_t(sizeof(walk_class));
constexpr uint32_t kRecBytes = 33u * 16u;  // strip records of one tile
constexpr uint32_t kWalkWarpSmem = uint32_t(kInBytes) + kRecBytes + 32u * 4u;
constexpr uint32_t kIndexGroup = 8;  // blocks of the block table per warp trip

enum { WALK_OK = 0, WALK_MALFORMED = 1, WALK_OVERSHOOT = 2, WALK_TABLE_FULL = 3 };

struct walk_tables {
    uint32_t classes_s;
    uint32_t hot_rows;
    uint32_t n_classes;
    uint32_t hotsum_s;  // shared address of sum(v + 1) of the hot rows, [n_classes][hot_rows] u32 (0: not staged)
    const uint8_t* lens_cls;
    const uint32_t* sums;
};

// the part of the strip decoder's shared-memory image that holds lengths (everything behind hot0 / hot1)
__host__ __device__ inline uint32_t walk_tables_skip(const dint_dict_view& dv) { return 2u * dv.n_classes * dv.hot_rows * 16u; }
__host__ __device__ inline uint32_t walk_tables_bytes(const dint_dict_view& dv) { return dv.strip_tables_bytes - walk_tables_skip(dv); }
__host__ __device__ inline uint32_t walk_hotsum_bytes(const dint_dict_view& dv) { return dv.n_classes * dv.hot_rows * 4u; }

// Stages the length tables + class table at smem; returns the first free byte (16-byte aligned).
template <bool HOTSUMS = false>
__device__ __forceinline__ uint8_t* walk_setup(const dint_dict_view& dv, const uint32_t* sums, uint8_t* smem, walk_tables& wt) {
    const uint32_t skip = walk_tables_skip(dv), bytes = dv.strip_tables_bytes - skip;
    const uint4* src = reinterpret_cast<const uint4*>(dv.strip_tables + skip);
    for (uint32_t i = threadIdx.x; i < bytes / 16u; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = __ldg(src + i);
    walk_class* classes = reinterpret_cast<walk_class*>(smem + bytes);
    if (threadIdx.x < 8) {
        const uint32_t c = min(threadIdx.x, dv.n_classes - 1u);
        walk_class ci;
        ci.hotlen_s = smem_u32(smem) + c * dv.hot_rows;
        ci.group_s = smem_u32(smem) + dv.group_off[c] - skip;
        ci.detail_s = smem_u32(smem) + dv.detail_off[c] - skip;
        ci.rows = max(dv.class_rows[c], 1u);
        ci.row_base = dv.class_base[c];
        ci.lut_g_off = dv.lens_cls_off[c];
        ci.pad0 = ci.pad1 = 0;
        classes[threadIdx.x] = ci;
    }
    wt.classes_s = smem_u32(classes);
    wt.hot_rows = dv.hot_rows;
    wt.n_classes = dv.n_classes;
    wt.lens_cls = dv.lens_cls;
    wt.sums = sums;
    wt.hotsum_s = 0;
    uint8_t* next = smem + bytes + kWalkClassBytes;
    if (HOTSUMS) {
        uint32_t* hs = reinterpret_cast<uint32_t*>(next);
        for (uint32_t i = threadIdx.x; i < dv.n_classes * dv.hot_rows; i += blockDim.x) {
            const uint32_t c = i / dv.hot_rows, r = i % dv.hot_rows;
            hs[i] = r < dv.class_rows[c] ? __ldg(sums + dv.class_base[c] + r) : 0u;
        }
        wt.hotsum_s = smem_u32(hs);
        next += walk_hotsum_bytes(dv);
    }
    return next;
}

__device__ __forceinline__ walk_class load_walk_class(const walk_tables& wt, uint32_t cls) {
    walk_class c;
    const uint32_t a = wt.classes_s + cls * uint32_t(sizeof(walk_class));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.hotlen_s), "=r"(c.group_s), "=r"(c.detail_s), "=r"(c.rows) : "r"(a));
    c.row_base = lds_u32(a + 16u);
    c.lut_g_off = lds_u32(a + 20u);
    c.pad0 = c.pad1 = 0;
    return c;
}

// log2(entry length) of row symc (already clamped to the class) — the strip decoder's lookup
__device__ __forceinline__ uint32_t walk_len_code(const walk_tables& wt, const walk_class& c, uint32_t symc) {
    if (symc < wt.hot_rows) return lds_u8(c.hotlen_s + symc);
    uint32_t code = lds_u8(c.group_s + (symc >> 5));
    if (code > 8u) {
        if (code != 0xFFu) code = lds_u8(c.detail_s + ((code - 9u) << 4) + ((symc & 31u) >> 1));
        else code = __ldg(wt.lens_cls + c.lut_g_off + (symc >> 1));
        code = (code >> ((symc & 1u) << 2)) & 15u;
    }
    return code;
}

// Input side of a walking warp: the cp.async ring of warp_state (dint_decoder.cuh), nothing else.
__device__ __forceinline__ void walk_init_input(warp_state& ws, uint32_t ring_s, const uint8_t* in, uint64_t in_bytes) {
    ws.stage_s = ws.lut_s = ws.que_s = 0;
    ws.inb_s = ring_s;
    ws.out = nullptr;
    ws.flushed = 0;
    ws.staged = 0;
    ws.acc = 0;
    ws.vec_ok = false;
    ws.anchor = 0;
    ws.win_end = ws.safe_end = ws.fast_end = ws.lo_rel = ws.hi_rel = 0;
    const uint64_t a = reinterpret_cast<uint64_t>(in);
    ws.in_lo = a & ~uint64_t(15);
    ws.in_hi = (a + in_bytes + 15) & ~uint64_t(15);
}

// Greedy packing of consecutive strips into strip units (same rules as the host builder, dint_host.cpp).
// Everything is warp-uniform; the strip words of the open unit sit in shared memory.
struct unit_packer {
    dint_strip_unit* units;
    unsigned long long* counter;  // non-null: slots come from this device counter (order is arbitrary)
    uint64_t slot;                // else: next slot of this warp's own range
    uint64_t capacity;            // slots at or beyond it are not written (WALK_TABLE_FULL)
    uint64_t closed;
    uint32_t words_s, max_ints, max_bytes;
    bool open, overflow;
    uint64_t u_in, u_out;
    uint32_t u_base, u_starts, u_ints, u_strips, u_len, u_flags;

    __device__ __forceinline__ void init(dint_strip_unit* u, unsigned long long* cnt, uint64_t first_slot, uint64_t cap,
                                         uint32_t words, uint32_t mi, uint32_t mb) {
        units = u;
        counter = cnt;
        slot = first_slot;
        capacity = cap;
        closed = 0;
        words_s = words;
        max_ints = mi;
        max_bytes = mb;
        open = overflow = false;
        u_in = u_out = 0;
        u_base = u_starts = u_ints = u_strips = u_len = u_flags = 0;
    }
    template <bool EMIT>
    __device__ __forceinline__ void close(unsigned lane) {
        if (!open) return;
        if (EMIT) {
            uint64_t s = slot;
            if (counter) {
                unsigned long long t = 0;
                if (lane == 0) t = atomicAdd(counter, 1ull);
                s = shfl64(t, 0);
            } else {
                ++slot;
            }
            __syncwarp();
            if (s < capacity) {
                const uint32_t w = lane < u_strips ? lds_u32(words_s + 4u * lane) : (u_ints << 13);
                uint32_t* up = reinterpret_cast<uint32_t*>(units + s);
                up[8u + lane] = w;
                if (lane == 0) {
                    reinterpret_cast<uint4*>(up)[0] = make_uint4(uint32_t(u_in), uint32_t(u_in >> 32), uint32_t(u_out), uint32_t(u_out >> 32));
                    reinterpret_cast<uint4*>(up)[1] = make_uint4(u_base, u_starts, u_len | (u_ints << 16), u_strips | (u_flags << 8));
                }
            } else {
                overflow = true;
            }
            __syncwarp();
        }
        ++closed;
        open = false;
    }
    // a strip [in0, in1) of the file producing n integers from output index out0 on; base = sum(v + 1) of its
    // list in front of it (+ the block base of an index block); ls: it begins a list; brk: it must begin a unit
    template <bool EMIT>
    __device__ __forceinline__ void add(unsigned lane, uint64_t in0, uint64_t in1, uint64_t out0, uint32_t n, uint32_t base,
                                        uint32_t sel, bool ls, bool run, bool brk) {
        if (open && (brk || u_strips >= uint32_t(DINT_UNIT_STRIPS) || in1 - u_in > max_bytes || u_ints + n > max_ints ||
                     out0 != u_out + u_ints))
            close<EMIT>(lane);
        if (!open) {
            u_in = in0;
            u_out = out0;
            u_base = base;
            u_starts = u_ints = u_strips = u_len = u_flags = 0;
            open = true;
        }
        if (EMIT && lane == 0) sts_u32(words_s + 4u * u_strips, uint32_t(in0 - u_in) | (u_ints << 13) | (sel << 25));
        if (u_strips && ls) u_starts |= 1u << u_strips;
        if (run) u_flags |= DINT_UNIT_HAS_RUNS;
        u_ints += n;
        u_len = uint32_t(in1 - u_in);
        ++u_strips;
    }
};

// Walks one codeword stream of n integers that starts at ring address a0 (first symbol; the selector byte of a
// multi block is in front of it). sel: 0..5 u16 codewords of class sel, 6..11 u8 codewords of class sel - 6.
// out0: output index of its first integer. acc: sum(v + 1) in front of it, updated. Strips go to pk. Returns false
// for a stream that does not tile n integers. a_end: ring address one past its last byte.
// SKIM: only where the stream ends and what it sums to (no strips): the pass that cuts long chains into pieces.
template <bool EMIT, bool SKIM = false>
__device__ __forceinline__ bool walk_stream(warp_state& ws, const walk_tables& wt, uint64_t in_base, uint32_t a0, uint32_t n,
                                            uint32_t sel, bool ls, bool brk, uint64_t out0, uint32_t& acc, unit_packer& pk,
                                            uint32_t rec_s, uint32_t strip_cw, uint32_t strip_shift, unsigned lane, uint32_t& a_end) {
    a_end = a0;
    if (n == 0) return true;
    const bool narrow = sel >= dint::kSelectors;
    const uint32_t cls = min(narrow ? sel - dint::kSelectors : sel, 7u);
    const walk_class wc = load_walk_class(wt, cls);
    const uint32_t last_row = (narrow ? min(wc.rows, 256u) : wc.rows) - 1u;
    const uint32_t szsh = narrow ? 0u : 1u;
    const unsigned lt = lanemask_lt();
    const uint32_t hotsum_s = wt.hotsum_s + cls * wt.hot_rows * 4u;
    uint32_t F = 0, done = 0, cw = 0, o_prev = 0;
    // the strip whose end is not known yet
    bool p_have = false, p_ls = false, p_run = false, p_brk = false;
    uint32_t p_in = 0, p_out = 0, p_acc = 0;
    auto file_off = [&](uint32_t rel) { return ws.anchor + rel - in_base; };
    while (done < n) {
        const uint32_t at = a0 + (F << szsh);
        ws.ensure_input(at, kTileLook, lane);
        const uint32_t a = at + (lane << szsh);
        uint32_t sym;
        if (narrow) sym = ws.rb(a);
        else if ((at & 1u) == 0u) sym = lds_u16(ws.inb_s + (a & (kInBytes - 1)));
        else sym = ws.rb(a) | (ws.rb(a + 1) << 8);
        const uint32_t Z = __ballot_sync(FULL, sym == 0u), O = __ballot_sync(FULL, sym == 1u);
        bool iscw = true;
        if (Z | O) {
            const uint32_t P = szsh ? payload_mask<1, 2>(Z, O, 1u) : payload_mask<2, 4>(Z, O, 1u);
            iscw = !((P >> lane) & 1u);
        }
        const uint32_t symc = min(sym, last_row);
        const uint32_t code = iscw ? walk_len_code(wt, wc, symc) : 0u;
        const uint32_t len = iscw ? (1u << code) : 0u;
        const uint32_t e = warp_inclusive_scan(len);
        const bool valid = iscw && e <= n - done;
        const uint32_t V = __ballot_sync(FULL, valid);
        if (V == 0u) return false;  // a codeword straddles the end of the stream
        if (__ballot_sync(FULL, valid && sym > last_row)) return false;  // codeword beyond the class
        const int last = 31 - __clz(V);
        const uint32_t T = __shfl_sync(FULL, e, last);
        const uint32_t step = sym < 2u ? 1u + ((sym == 0u ? 2u : 4u) >> szsh) : 1u;
        const uint32_t adv = uint32_t(last) + __shfl_sync(FULL, step, last);
        // sum(v + 1) of every codeword: exception payload + 1, or the row's precomputed sum
        uint32_t sv = 0;
        if (valid) {
            if (sym < 2u) {
                uint32_t pay = ws.r32(a + (1u << szsh));
                if (sym == 0u) pay &= 0xFFFFu;
                sv = pay + 1u;
            } else if (SKIM && symc < wt.hot_rows) {
                sv = lds_u32(hotsum_s + (symc << 2));
            } else {
                sv = __ldg(wt.sums + wc.row_base + symc);
            }
        }
        if (SKIM) {
            acc += __reduce_add_sync(FULL, sv);
            done += T;
            F += adv;
            continue;
        }
        const uint32_t s_incl = warp_inclusive_scan(sv);
        const uint32_t k = cw + __popc(V & lt);
        const uint32_t o = done + e - len;
        const uint32_t prevm = V & lt;
        const uint32_t o_src = __shfl_sync(FULL, o, prevm ? 31 - __clz(prevm) : int(lane));
        const uint32_t op = prevm ? o_src : o_prev;
        const bool is_start = valid && ((k % strip_cw) == 0u || (o >> strip_shift) != (op >> strip_shift));
        const uint32_t S = __ballot_sync(FULL, is_start);
        const bool run_any = __ballot_sync(FULL, valid && sym - 2u < dint::kReserved - 2u) != 0u;
        if (run_any) p_run = true;
        if (S) {
            if (is_start) {
                const uint32_t idx = __popc(S & lt);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rec_s + 16u * idx), "r"(a), "r"(o),
                             "r"(acc + s_incl - sv), "r"(0u)
                             : "memory");
            }
            __syncwarp();
            const uint32_t m = __popc(S);
            for (uint32_t i = 0; i < m; ++i) {
                const uint4 r = lds_v4(rec_s + 16u * i);
                if (p_have) {
                    pk.add<EMIT>(lane, file_off(p_in), file_off(r.x), out0 + p_out, r.y - p_out, p_acc, sel, p_ls, p_run, p_brk);
                    p_ls = p_brk = false;
                } else {
                    p_ls = ls;
                    p_brk = brk;
                }
                p_have = true;
                p_in = r.x;
                p_out = r.y;
                p_acc = r.z;
                p_run = run_any;
            }
            __syncwarp();
        }
        o_prev = __shfl_sync(FULL, o, last);
        acc += __shfl_sync(FULL, s_incl, last);
        done += T;
        F += adv;
        cw += __popc(V);
    }
    a_end = a0 + (F << szsh);
    if (!SKIM && p_have) pk.add<EMIT>(lane, file_off(p_in), file_off(a_end), out0 + p_out, n - p_out, p_acc, sel, p_ls, p_run, p_brk);
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// vroom .bin: one warp per chain
// ---------------------------------------------------------------------------------------------------------
struct vroom_walk_params {
    dint_dict_view dv;
    const uint32_t* sums;
    const uint8_t* in;
    uint64_t in_bytes;
    uint32_t block_size, multi;
    const uint64_t* chain_start;  // n_chains + 1 byte offsets, the last one = in_bytes
    uint64_t n_chains;
    const uint32_t* chain_ids;    // the chains this launch walks (long ones: a warp each, the others: a thread each)
    uint64_t n_ids;
    uint32_t strip_ints;          // thread walker: a strip also closes once it holds this many integers
    uint64_t* counts;             // [n_chains][4] {lists, ints, units, status | end offset << 8}: written by the count pass
    const uint64_t* bases;        // [n_chains][4] {first list, first int, first unit, -}: read by the emit pass
    dint_strip_unit* units;
    uint64_t units_cap;
    uint64_t* list_offsets;       // emit: n_lists (+1 written by the host)
    uint64_t* list_byte_offsets;  // emit: optional
    uint32_t* status;             // emit: != 0 when the pass disagrees with the count pass
    uint32_t unit_ints, unit_bytes, strip_cw, strip_shift;
    // where a chain starts / ends when that is not a list start (pieces of a long chain, cut by vroom_skim_kernel):
    // [c / 2] = {integers of the current list still to come at the chain's start (0: a list header is next),
    //            sum(v + 1) of that list so far, bit 0: the chain ENDS inside a list, integers still to come there}
    unsigned long long* queue;    // thread walker / skim: next entry of chain_ids to hand out (zeroed before the launch)
    const uint4* chain_state;     // null: every chain runs from a list start to a list start
    // skim pass: the cuts of chain c go to cuts[2 * cut_off[c / 2] ...] = {file offset, list_rem | acc << 32}
    const uint64_t* cut_off;      // [n_chains / 2 + 1]
    uint64_t* cuts;
    uint32_t* cut_counts;         // [n_chains / 2]
    uint64_t cut_bytes;           // distance between two cuts
};

template <bool EMIT>
__global__ void __launch_bounds__(512, 2) vroom_walk_kernel(const vroom_walk_params P) {
    extern __shared__ __align__(16) uint8_t smem[];
    walk_tables wt;
    uint8_t* free0 = walk_setup(P.dv, P.sums, smem, wt);
    const unsigned lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint8_t* mine = free0 + size_t(warp) * kWalkWarpSmem;
    const uint32_t ring_s = smem_u32(mine), rec_s = ring_s + uint32_t(kInBytes), words_s = rec_s + kRecBytes;
    warp_state ws;
    walk_init_input(ws, ring_s, P.in, P.in_bytes);
    __syncthreads();
    const uint64_t in_base = reinterpret_cast<uint64_t>(P.in);
    for (uint64_t ci = uint64_t(blockIdx.x) * n_warps + warp; ci < P.n_ids; ci += uint64_t(gridDim.x) * n_warps) {
        const uint64_t c = P.chain_ids[ci];
        const uint64_t start = P.chain_start[c], end = P.chain_start[c + 1];
        uint64_t list_base = 0, int_base = 0, unit_base = 0;
        if (EMIT) {
            list_base = P.bases[4 * c];
            int_base = P.bases[4 * c + 1];
            unit_base = P.bases[4 * c + 2];
        }
        unit_packer pk;
        pk.init(P.units, nullptr, unit_base, P.units_cap, words_s, P.unit_ints, P.unit_bytes);
        uint64_t lists = 0, ints = 0, pos = start;
        uint32_t status = WALK_OK;
        uint32_t p = ws.set_anchor(in_base + start, ~uint64_t(0));
        while (pos < end) {
            if (p > 0x60000000u) p = ws.set_anchor(in_base + pos, ~uint64_t(0));
            // header::read (vroom_env/codecs.hpp:117-123): vbyte n, vbyte universe; the terminator byte has bit 7 set
            ws.ensure_input(p, 16u, lane);
            const uint32_t b = lane < 10u ? ws.rb(p + lane) : 0u;
            const uint32_t term = __ballot_sync(FULL, (b & 128u) != 0u);
            const uint32_t t1 = term ? uint32_t(__ffs(term)) - 1u : 32u;
            const uint32_t term2 = t1 < 31u ? term & ~((2u << t1) - 1u) : 0u;
            const uint32_t t2 = term2 ? uint32_t(__ffs(term2)) - 1u : 32u;
            if (t1 > 4u || t2 > t1 + 5u || pos + t2 + 1u > P.in_bytes) {
                status = WALK_MALFORMED;
                break;
            }
            const uint32_t n = __reduce_add_sync(FULL, lane <= t1 ? (b & 127u) << (7u * lane) : 0u);
            if (uint64_t(n) > 256ull * (P.in_bytes - pos)) {  // a byte encodes at most 256 integers (a run codeword)
                status = WALK_MALFORMED;
                break;
            }
            if (EMIT && lane == 0) {
                P.list_offsets[list_base + lists] = int_base + ints;
                if (P.list_byte_offsets) P.list_byte_offsets[list_base + lists] = pos;
            }
            p += t2 + 1u;
            uint32_t acc = 0, a_end = p;
            bool ok = true;
            if (!P.multi) {
                // single_dint::decode (vroom_env/dint_codecs.hpp:37-107): one u16 stream per list
                ok = walk_stream<EMIT>(ws, wt, in_base, p, n, 0u, true, false, int_base + ints, acc, pk, rec_s, P.strip_cw, P.strip_shift, lane, a_end);
                p = a_end;
            } else {
                // multi_opt_dint::decode (vroom_env/dint_codecs.hpp:521-619): a selector byte in front of every block
                for (uint32_t done = 0; ok && done < n; done += P.block_size) {
                    const uint32_t size = min(P.block_size, n - done);
                    ws.ensure_input(p, 16u, lane);
                    const uint32_t sel = ws.rb(p);
                    if (sel >= 2u * dint::kSelectors || (sel % dint::kSelectors) >= wt.n_classes) {
                        ok = false;
                        break;
                    }
                    ok = walk_stream<EMIT>(ws, wt, in_base, p + 1u, size, sel, done == 0u, false, int_base + ints + done, acc, pk, rec_s,
                                           P.strip_cw, P.strip_shift, lane, a_end);
                    p = a_end;
                }
            }
            pos = ws.anchor + p - in_base;
            if (!ok || pos > P.in_bytes) {
                status = WALK_MALFORMED;
                break;
            }
            ints += n;
            ++lists;
        }
        pk.close<EMIT>(lane);
        if (status == WALK_OK && pos != end) status = WALK_OVERSHOOT;
        if (status == WALK_OK && pk.overflow) status = WALK_TABLE_FULL;
        if (lane == 0) {
            if (!EMIT) {
                P.counts[4 * c] = lists;
                P.counts[4 * c + 1] = ints;
                P.counts[4 * c + 2] = pk.closed;
                P.counts[4 * c + 3] = uint64_t(status) | (pos << 8);
            } else if (status != WALK_OK || lists != P.counts[4 * c] || ints != P.counts[4 * c + 1] || pk.closed != P.counts[4 * c + 2]) {
                atomicMax(P.status, status ? status : uint32_t(WALK_MALFORMED));  // the two passes disagree
            }
        }
    }
    cp_async_wait_all();
}


// ---------------------------------------------------------------------------------------------------------
// vroom .bin: the skim pass — one warp per LONG chain (a stretch without sync points: long lists). It walks the
// chain like vroom_walk_kernel but builds nothing; every cut_bytes it records the next place a thread can start
// from: a list header, or the selector byte of a block of a multi list together with the integers of the list
// still to come and the list's sum(v + 1) so far. The pieces between two cuts are then walked by threads like every
// other chain (a thread needs ~10 instructions per codeword, this tile walk ~4 per codeword and no bookkeeping).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2) vroom_skim_kernel(const vroom_walk_params P) {
    extern __shared__ __align__(16) uint8_t smem[];
    walk_tables wt;
    uint8_t* free0 = walk_setup<true>(P.dv, P.sums, smem, wt);
    const unsigned lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    uint8_t* mine = free0 + size_t(warp) * kWalkWarpSmem;
    const uint32_t ring_s = smem_u32(mine), rec_s = ring_s + uint32_t(kInBytes);
    warp_state ws;
    walk_init_input(ws, ring_s, P.in, P.in_bytes);
    __syncthreads();
    const uint64_t in_base = reinterpret_cast<uint64_t>(P.in);
    unit_packer pk;  // not used by a skimming walk_stream
    pk.init(nullptr, nullptr, 0, 0, 0, 0, 0);
    for (;;) {
        unsigned long long ci = 0;  // the queue hands the chains out longest first
        if (lane == 0) ci = atomicAdd(P.queue, 1ull);
        ci = shfl64(ci, 0);
        if (ci >= P.n_ids) break;
        const uint64_t c = P.chain_ids[ci];
        const uint64_t start = P.chain_start[c], end = P.chain_start[c + 1];
        const uint64_t cap = P.cut_off[(c >> 1) + 1] - P.cut_off[c >> 1];
        uint64_t* dst = P.cuts + 2 * P.cut_off[c >> 1];
        uint64_t n_cuts = 0, last_cut = start, pos = start;
        auto cut = [&](uint64_t at, uint32_t rem, uint32_t acc) {
            if (lane == 0) {
                dst[2 * n_cuts] = at;
                dst[2 * n_cuts + 1] = uint64_t(rem) | (uint64_t(acc) << 32);
            }
            ++n_cuts;
            last_cut = at;
        };
        uint32_t p = ws.set_anchor(in_base + start, ~uint64_t(0));
        while (pos < end) {
            if (p > 0x60000000u) p = ws.set_anchor(in_base + pos, ~uint64_t(0));
            if (pos - last_cut >= P.cut_bytes && n_cuts < cap) cut(pos, 0u, 0u);
            ws.ensure_input(p, 16u, lane);
            const uint32_t b = lane < 10u ? ws.rb(p + lane) : 0u;
            const uint32_t term = __ballot_sync(FULL, (b & 128u) != 0u);
            const uint32_t t1 = term ? uint32_t(__ffs(term)) - 1u : 32u;
            const uint32_t term2 = t1 < 31u ? term & ~((2u << t1) - 1u) : 0u;
            const uint32_t t2 = term2 ? uint32_t(__ffs(term2)) - 1u : 32u;
            if (t1 > 4u || t2 > t1 + 5u || pos + t2 + 1u > P.in_bytes) break;  // malformed: the walkers will say so
            const uint32_t n = __reduce_add_sync(FULL, lane <= t1 ? (b & 127u) << (7u * lane) : 0u);
            if (uint64_t(n) > 256ull * (P.in_bytes - pos)) break;
            p += t2 + 1u;
            uint32_t acc = 0, a_end = p;
            bool ok = true;
            if (!P.multi) {
                ok = walk_stream<false, true>(ws, wt, in_base, p, n, 0u, true, false, 0, acc, pk, rec_s, P.strip_cw, P.strip_shift, lane, a_end);
                p = a_end;
            } else {
                for (uint32_t done = 0; ok && done < n; done += P.block_size) {
                    if (done) {
                        const uint64_t at = ws.anchor + p - in_base;
                        if (at - last_cut >= P.cut_bytes && at < end && n_cuts < cap) cut(at, n - done, acc);
                    }
                    const uint32_t size = min(P.block_size, n - done);
                    ws.ensure_input(p, 16u, lane);
                    const uint32_t sel = ws.rb(p);
                    if (sel >= 2u * dint::kSelectors || (sel % dint::kSelectors) >= wt.n_classes) {
                        ok = false;
                        break;
                    }
                    ok = walk_stream<false, true>(ws, wt, in_base, p + 1u, size, sel, false, false, 0, acc, pk, rec_s, P.strip_cw,
                                                  P.strip_shift, lane, a_end);
                    p = a_end;
                }
            }
            pos = ws.anchor + p - in_base;
            if (!ok || pos > P.in_bytes) break;
        }
        if (lane == 0) P.cut_counts[c >> 1] = uint32_t(n_cuts);
    }
    cp_async_wait_all();
}


// ---------------------------------------------------------------------------------------------------------
// vroom .bin: one THREAD per chain — the reference's scalar walk (vroom_env/decode.cpp:139-150 over
// dint_codecs.hpp:45-101 / 530-611) run by tens of thousands of threads at once, each on its own stretch of the
// file between two sync points. One loop trip handles one token (a codeword, or the switch to the next block /
// list), so the lanes of a warp stay in the same code whatever the lengths of their lists are. A thread that has
// finished its chain takes the next one from a queue inside the same loop (the host hands the chains out longest
// first): no lane waits for the longest chain of its warp.
// ---------------------------------------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256, 3) vroom_thread_walk_kernel(const vroom_walk_params P) {
    extern __shared__ __align__(16) uint8_t smem[];
    walk_tables wt;
    walk_setup<EMIT>(P.dv, P.sums, smem, wt);
    __syncthreads();
    uint64_t c = 0, end = 0;
    uint64_t list_base = 0, int_base = 0, slot = 0;
    const uint64_t in_base = reinterpret_cast<uint64_t>(P.in);
    // The 32 bytes under the cursor live in registers (two aligned 16-byte loads); a codeword is served from them and
    // a new 16-byte piece is fetched only when the cursor leaves the window: every lane reads its own stretch of
    // the file, so loads per codeword are what bounds this kernel (each one is a sector of its own).
    const uint64_t in_hi16 = (in_base + P.in_bytes + 15) & ~uint64_t(15);
    uint4 wlo = make_uint4(0, 0, 0, 0), whi = make_uint4(0, 0, 0, 0);
    uint64_t wbase = ~uint64_t(0) - 64;  // absolute address of wlo (16-byte aligned)
    auto ld16 = [&](uint64_t a) -> uint4 {
        return a < in_hi16 ? __ldg(reinterpret_cast<const uint4*>(a)) : make_uint4(0, 0, 0, 0);
    };
    // the 8 bytes at file offset q (whatever follows the file reads as the bytes of its last 16-byte piece, then
    // zeros; positions are checked against in_bytes by the caller), any alignment
    auto load8 = [&](uint64_t q) -> uint64_t {
        const uint64_t a = in_base + q;
        if (a - wbase >= 24u) {  // also true for a < wbase (wraps)
            if (a - wbase < 40u) {  // the usual step: slide by one piece
                wlo = whi;
                wbase += 16;
                whi = ld16(wbase + 16);
            } else {
                wbase = a & ~uint64_t(15);
                wlo = ld16(wbase);
                whi = ld16(wbase + 16);
            }
        }
        const uint32_t idx = uint32_t(a - wbase);  // 0..23
        const bool up = idx >= 16u;
        const uint32_t a0 = up ? whi.x : wlo.x, a1 = up ? whi.y : wlo.y, a2 = up ? whi.z : wlo.z, a3 = up ? whi.w : wlo.w;
        const uint32_t a4 = up ? 0u : whi.x, a5 = up ? 0u : whi.y;
        const uint32_t j = (idx >> 2) & 3u;
        const uint32_t w0 = j == 0u ? a0 : j == 1u ? a1 : j == 2u ? a2 : a3;
        const uint32_t w1 = j == 0u ? a1 : j == 1u ? a2 : j == 2u ? a3 : a4;
        const uint32_t w2 = j == 0u ? a2 : j == 1u ? a3 : j == 2u ? a4 : a5;
        const uint32_t sh = (idx & 3u) * 8u;
        return (uint64_t(__funnelshift_r(w1, w2, sh)) << 32) | __funnelshift_r(w0, w1, sh);
    };
    // ---- unit packer (unit_packer's rules; the strip words of the open unit live in local memory) ----------
    uint32_t words[DINT_UNIT_STRIPS];
    bool open = false, overflow = false;
    uint64_t u_in = 0, u_out = 0, closed = 0;
    uint32_t u_base = 0, u_starts = 0, u_ints = 0, u_strips = 0, u_len = 0, u_flags = 0;
    auto close_unit = [&]() {
        if (!open) return;
        if (EMIT) {
            if (slot < P.units_cap) {
                uint4* up = reinterpret_cast<uint4*>(P.units + slot);
                up[0] = make_uint4(uint32_t(u_in), uint32_t(u_in >> 32), uint32_t(u_out), uint32_t(u_out >> 32));
                up[1] = make_uint4(u_base, u_starts, u_len | (u_ints << 16), u_strips | (u_flags << 8));
#pragma unroll
                for (uint32_t k = 0; k < DINT_UNIT_STRIPS; k += 4) {
                    uint4 w;
                    w.x = k < u_strips ? words[k] : (u_ints << 13);
                    w.y = k + 1 < u_strips ? words[k + 1] : (u_ints << 13);
                    w.z = k + 2 < u_strips ? words[k + 2] : (u_ints << 13);
                    w.w = k + 3 < u_strips ? words[k + 3] : (u_ints << 13);
                    up[2 + k / 4] = w;
                }
            } else {
                overflow = true;
            }
            ++slot;
        }
        ++closed;
        open = false;
    };
    auto add_strip = [&](uint64_t in0, uint64_t in1, uint64_t out0, uint32_t n, uint32_t base, uint32_t sel, bool ls, bool run) {
        if (open && (u_strips >= uint32_t(DINT_UNIT_STRIPS) || in1 - u_in > P.unit_bytes || u_ints + n > P.unit_ints || out0 != u_out + u_ints))
            close_unit();
        if (!open) {
            u_in = in0;
            u_out = out0;
            u_base = base;
            u_starts = u_ints = u_strips = u_len = u_flags = 0;
            open = true;
        }
        if (EMIT) words[u_strips] = uint32_t(in0 - u_in) | (u_ints << 13) | (sel << 25);
        if (u_strips && ls) u_starts |= 1u << u_strips;
        if (run) u_flags |= DINT_UNIT_HAS_RUNS;
        u_ints += n;
        u_len = uint32_t(in1 - u_in);
        ++u_strips;
    };

    uint64_t pos = 0, lists = 0, ints = 0;
    uint32_t status = WALK_OK;
    uint32_t list_rem = 0, str_rem = 0;  // integers of the list behind the current stream / left in the current stream
    uint32_t acc = 0;                    // sum(v + 1) of the list so far
    bool first_block = false;
    // a piece of a long chain starts and / or ends at a block boundary inside a list (vroom_skim_kernel)
    bool end_mid = false;
    uint32_t end_rem = 0;
    bool walking = false;
    // current stream
    uint32_t sel = 0, sb = 2, last_row = 0, hotlen_s = 0, group_s = 0, detail_s = 0, row_base = 0, lut_g = 0, hotsum_s = 0;
    // open strip
    uint64_t s_pos = 0, s_out = 0;
    uint32_t s_acc = 0, s_ints = 0, cw = 0;
    bool s_ls = false, s_run = false;
    for (;;) {
        if (!walking) {
            // ---- next chain of the queue ------------------------------------------------------------------------
            const unsigned long long k = atomicAdd(P.queue, 1ull);
            if (k >= P.n_ids) break;
            c = P.chain_ids[k];
            pos = P.chain_start[c];
            end = P.chain_start[c + 1];
            if (EMIT) {
                list_base = P.bases[4 * c];
                int_base = P.bases[4 * c + 1];
                slot = P.bases[4 * c + 2];
            }
            lists = ints = closed = 0;
            status = WALK_OK;
            list_rem = str_rem = acc = end_rem = 0;
            first_block = end_mid = open = overflow = false;
            if (P.chain_state) {
                const uint4 cs = __ldg(P.chain_state + (c >> 1));
                list_rem = cs.x;
                acc = cs.y;
                end_mid = (cs.z & 1u) != 0u;
                end_rem = cs.w;
            }
            walking = true;
        }
        bool fin = true;  // left true by every `break` out of the token below: the chain is over (or malformed)
        do {
        if (str_rem == 0u) {
            if (list_rem != 0u && end_mid && pos >= end) break;  // the next piece starts at this block
            if (list_rem == 0u) {
                // ---- between two lists: header::read (vroom_env/codecs.hpp:117-123) ---------------------------------
                if (pos >= end) break;
                const uint64_t h = load8(pos);
                uint32_t n = 0, k = 0, ok = 0;
                for (; k < 5u; ++k) {
                    const uint32_t b = uint32_t(h >> (8u * k)) & 0xFFu;
                    n += (b & 127u) << (7u * k);
                    if (b & 128u) {
                        ok = 1;
                        break;
                    }
                }
                uint32_t hdr = k + 1u;
                if (ok) {  // universe: only its length matters here
                    const uint64_t h2 = load8(pos + hdr);
                    ok = 0;
                    for (k = 0; k < 5u; ++k)
                        if ((h2 >> (8u * k)) & 128u) {
                            ok = 1;
                            break;
                        }
                    hdr += k + 1u;
                }
                // (a byte encodes at most 256 integers — a one-byte run codeword: a larger n cannot be a list of this file)
                if (!ok || pos + hdr > P.in_bytes || uint64_t(n) > 256ull * (P.in_bytes - pos)) {
                    status = WALK_MALFORMED;
                    break;
                }
                if (EMIT) {
                    P.list_offsets[list_base + lists] = int_base + ints;
                    if (P.list_byte_offsets) P.list_byte_offsets[list_base + lists] = pos;
                }
                ++lists;
                pos += hdr;
                list_rem = n;
                acc = 0;
                first_block = true;
                if (n == 0u) {
                    fin = false;
                    break;
                }
            }
            // ---- next stream of the list: a multi block behind its selector byte, or the whole single list -------
            uint32_t size = list_rem;
            sel = 0;
            if (P.multi) {
                size = min(size, P.block_size);
                sel = uint32_t(load8(pos)) & 0xFFu;
                if (pos >= P.in_bytes || sel >= 2u * dint::kSelectors || (sel % dint::kSelectors) >= wt.n_classes) {
                    status = WALK_MALFORMED;
                    break;
                }
                ++pos;
            }
            const bool narrow = sel >= dint::kSelectors;
            const walk_class wc = load_walk_class(wt, narrow ? sel - dint::kSelectors : sel);
            hotlen_s = wc.hotlen_s;
            group_s = wc.group_s;
            detail_s = wc.detail_s;
            row_base = wc.row_base;
            lut_g = wc.lut_g_off;
            hotsum_s = wt.hotsum_s + (narrow ? sel - dint::kSelectors : sel) * wt.hot_rows * 4u;
            last_row = (narrow ? min(wc.rows, 256u) : wc.rows) - 1u;
            sb = narrow ? 1u : 2u;
            list_rem -= size;
            str_rem = size;
            s_pos = pos;
            s_out = int_base + ints;
            s_acc = acc;
            s_ints = 0;
            cw = 0;
            s_ls = first_block;
            s_run = false;
            first_block = false;
        }
        // ---- one codeword ----------------------------------------------------------------------------------------
        const uint64_t bits = load8(pos);
        const uint32_t sym = uint32_t(bits) & (sb == 1u ? 0xFFu : 0xFFFFu);
        uint32_t len = 1, sv = 0, step = sb;
        if (sym < 2u) {
            uint32_t pay = uint32_t(bits >> (8u * sb));
            if (sym == 0u) pay &= 0xFFFFu;
            sv = pay + 1u;
            step += 2u + 2u * sym;
        } else {
            if (sym > last_row) {
                status = WALK_MALFORMED;
                break;
            }
            uint32_t code;
            if (sym < wt.hot_rows) {
                code = lds_u8(hotlen_s + sym);
            } else {
                code = lds_u8(group_s + (sym >> 5));
                if (code > 8u) {
                    if (code != 0xFFu) code = lds_u8(detail_s + ((code - 9u) << 4) + ((sym & 31u) >> 1));
                    else code = __ldg(wt.lens_cls + lut_g + (sym >> 1));
                    code = (code >> ((sym & 1u) << 2)) & 15u;
                }
            }
            len = 1u << code;
            // the docID prefix is only stored by the emit pass; hot rows have their sums in shared memory
            if (EMIT) sv = sym < wt.hot_rows ? lds_u32(hotsum_s + (sym << 2)) : __ldg(wt.sums + row_base + sym);
            s_run = s_run || sym < dint::kReserved;
        }
        pos += step;
        if (len > str_rem || pos > P.in_bytes) {  // entries must tile the stream exactly (`i != n` loop)
            status = WALK_MALFORMED;
            break;
        }
        acc += sv;
        str_rem -= len;
        s_ints += len;
        ints += len;
        ++cw;
        if (str_rem == 0u || cw >= P.strip_cw || s_ints >= P.strip_ints) {
            add_strip(s_pos, pos, s_out, s_ints, s_acc, sel, s_ls, s_run);
            s_pos = pos;
            s_out += s_ints;
            s_acc = acc;
            s_ints = 0;
            cw = 0;
            s_ls = s_run = false;
        }
        fin = false;
        } while (false);
        if (!fin) continue;
        // ---- the chain is over ---------------------------------------------------------------------------------
        close_unit();
        if (status == WALK_OK && (pos != end || list_rem != (end_mid ? end_rem : 0u))) status = WALK_OVERSHOOT;
        if (status == WALK_OK && overflow) status = WALK_TABLE_FULL;
        if (!EMIT) {
            P.counts[4 * c] = lists;
            P.counts[4 * c + 1] = ints;
            P.counts[4 * c + 2] = closed;
            P.counts[4 * c + 3] = uint64_t(status) | (pos << 8);
        } else if (status != WALK_OK || lists != P.counts[4 * c] || ints != P.counts[4 * c + 1] || closed != P.counts[4 * c + 2]) {
            atomicMax(P.status, status ? status : uint32_t(WALK_MALFORMED));
        }
        walking = false;
    }
}

// ---------------------------------------------------------------------------------------------------------
// chain totals -> chain bases: single-pass exclusive scan with decoupled look-back
// ---------------------------------------------------------------------------------------------------------
// After the count pass every chain knows how many lists, integers and strip units it holds; the emit pass needs
// the sums over all chains in front of it. One CTA per tile of 256 chains: local scan, publish the tile's
// aggregate, look back over the preceding tiles' descriptors (aggregate or inclusive prefix, whichever is there)
// instead of waiting for a global barrier. The same kernel checks that every chain ended exactly on the next
// chain's start (if one did not, the host repairs the chain table and this fast path is not taken).
struct chain_scan_params {
    uint64_t n;                       // chains, in file order; chain k occupies slots [8k, 8k + 4) of cnt / bases
    const uint64_t* se;               // [2k] start, [2k + 1] end of chain k
    const uint64_t* cnt;              // count pass: lists, ints, units, status | reached << 8
    uint64_t* bases;                  // out: first list, first int, first unit of chain k
    unsigned long long* tile_agg;     // [tiles][4] aggregate of a tile (3 values)
    unsigned long long* tile_inc;     // [tiles][4] inclusive prefix up to and including a tile
    unsigned int* tile_flag;          // [tiles] 0: nothing yet, 1: aggregate published, 2: inclusive prefix published
    unsigned int* tile_counter;       // hands out tile ids in launch order (forward progress of the look-back)
    unsigned long long* totals;       // out: lists, ints, units, (!= 0: some chain did not land on its successor)
};

__global__ void __launch_bounds__(256) chain_scan_kernel(const chain_scan_params P) {
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[8][3], s_prefix[3];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(P.tile_counter, 1u);
    __syncthreads();
    const uint64_t tile = s_tile, k = tile * 256 + threadIdx.x;
    const bool valid = k < P.n;
    unsigned long long v[3] = {0, 0, 0};
    bool bad = false;
    if (valid) {
        v[0] = P.cnt[8 * k];
        v[1] = P.cnt[8 * k + 1];
        v[2] = P.cnt[8 * k + 2];
        const uint64_t st = P.cnt[8 * k + 3];
        bad = (st & 0xFF) != WALK_OK || (st >> 8) != P.se[2 * k + 1];
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(P.totals + 3, 1ull);
    // CTA-wide exclusive scan of the three counters
    unsigned long long incl[3];
    for (int j = 0; j < 3; ++j) {
        unsigned long long x = v[j];
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = shfl64(x, lane >= unsigned(o) ? int(lane) - o : int(lane));
            if (lane >= unsigned(o)) x += y;
        }
        incl[j] = x;
        if (lane == 31) s_warp[warp][j] = x;
    }
    __syncthreads();
    unsigned long long excl[3], agg[3];
    for (int j = 0; j < 3; ++j) {
        unsigned long long before = 0, total = 0;
        for (unsigned w = 0; w < 8; ++w) {
            const unsigned long long t = s_warp[w][j];
            if (w < warp) before += t;
            total += t;
        }
        excl[j] = before + incl[j] - v[j];
        agg[j] = total;
    }
    // ---- decoupled look-back (thread 0) ----------------------------------------------------------------------
    if (threadIdx.x == 0) {
        volatile unsigned long long* vagg = P.tile_agg;
        volatile unsigned long long* vinc = P.tile_inc;
        volatile unsigned int* vflag = P.tile_flag;
        unsigned long long prefix[3] = {0, 0, 0};
        if (tile == 0) {
            for (int j = 0; j < 3; ++j) vinc[j] = agg[j];
            __threadfence();
            vflag[0] = 2u;
        } else {
            for (int j = 0; j < 3; ++j) vagg[4 * tile + j] = agg[j];
            __threadfence();
            vflag[tile] = 1u;
            for (uint64_t t = tile; t-- > 0;) {
                unsigned int f;
                while ((f = vflag[t]) == 0u) {
                }
                __threadfence();
                if (f == 2u) {
                    for (int j = 0; j < 3; ++j) prefix[j] += vinc[4 * t + j];
                    break;
                }
                for (int j = 0; j < 3; ++j) prefix[j] += vagg[4 * t + j];
            }
            for (int j = 0; j < 3; ++j) vinc[4 * tile + j] = prefix[j] + agg[j];
            __threadfence();
            vflag[tile] = 2u;
        }
        for (int j = 0; j < 3; ++j) s_prefix[j] = prefix[j];
        if ((tile + 1) * 256 >= P.n)
            for (int j = 0; j < 3; ++j) P.totals[j] = prefix[j] + agg[j];
    }
    __syncthreads();
    if (valid)
        for (int j = 0; j < 3; ++j) P.bases[8 * k + j] = s_prefix[j] + excl[j];
}

// ---------------------------------------------------------------------------------------------------------
// speculative sync points of a vroom .bin: one thread per candidate byte position
// ---------------------------------------------------------------------------------------------------------
struct find_sync_params {
    dint_dict_view dv;
    const uint32_t* sums;
    const uint8_t* in;
    uint64_t in_bytes;
    uint32_t block_size, multi;
    uint64_t stride;    // target t = (t + 1) * stride
    uint64_t n_targets;
    uint32_t window;    // bytes searched behind every target (multiple of 32, <= stride)
    uint32_t depth;     // consecutive lists that must check out
    uint32_t max_n;     // longest list accepted as evidence
    uint64_t* found;    // [n_targets] byte offset of a list start, or ~0
};

// thread-sequential check: does a list that starts at byte p decode to values that sum to its `universe`?
// returns the offset one past it, or ~0. (scalar loop of the reference: vroom_env/dint_codecs.hpp:45-101)
__device__ __forceinline__ uint64_t check_list(const find_sync_params& P, const walk_tables& wt, uint64_t p) {
    const uint8_t* in = P.in;
    const uint64_t end = P.in_bytes;
    uint32_t hdr[2];
    for (int h = 0; h < 2; ++h) {
        uint32_t x = 0, k = 0;
        for (;; ++k) {
            if (p >= end || k > 4u) return ~uint64_t(0);
            const uint32_t c = __ldg(in + p++);
            x += (c & 127u) << (7u * k);
            if (c & 128u) break;
        }
        hdr[h] = x;
    }
    const uint32_t n = hdr[0], universe = hdr[1];
    if (n == 0u || n > P.max_n) return ~uint64_t(0);
    uint32_t sum = 0;
    for (uint32_t done = 0; done < n;) {
        uint32_t size = n - done, sel = 0;
        if (P.multi) {
            size = min(size, P.block_size);
            if (p >= end) return ~uint64_t(0);
            sel = __ldg(in + p++);
            if (sel >= 2u * dint::kSelectors || (sel % dint::kSelectors) >= wt.n_classes) return ~uint64_t(0);
        }
        const bool narrow = sel >= dint::kSelectors;
        const walk_class wc = load_walk_class(wt, narrow ? sel - dint::kSelectors : sel);
        const uint32_t last_row = (narrow ? min(wc.rows, 256u) : wc.rows) - 1u;
        const uint32_t sb = narrow ? 1u : 2u;
        for (uint32_t i = 0; i < size;) {
            if (p + sb > end) return ~uint64_t(0);
            uint32_t sym = __ldg(in + p);
            if (!narrow) sym |= uint32_t(__ldg(in + p + 1)) << 8;
            p += sb;
            if (sym > last_row) return ~uint64_t(0);
            if (sym < 2u) {
                const uint32_t nb = 2u + 2u * sym;
                if (p + nb > end) return ~uint64_t(0);
                uint32_t pay = 0;
                for (uint32_t j = 0; j < nb; ++j) pay |= uint32_t(__ldg(in + p + j)) << (8u * j);
                p += nb;
                sum += pay;
                i += 1u;
            } else {
                const uint32_t len = 1u << walk_len_code(wt, wc, sym);
                sum += __ldg(wt.sums + wc.row_base + sym) - len;  // sums[] holds sum(v + 1)
                i += len;
            }
            if (i > size) return ~uint64_t(0);
        }
        done += size;
    }
    return sum == universe ? p : ~uint64_t(0);
}

__global__ void __launch_bounds__(256) find_sync_kernel(const find_sync_params P) {
    extern __shared__ __align__(16) uint8_t smem[];
    walk_tables wt;
    walk_setup(P.dv, P.sums, smem, wt);
    __syncthreads();
    const unsigned lane = lane_id();
    const uint64_t gwarp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t t = gwarp; t < P.n_targets; t += nwarps) {
        const uint64_t t0 = (t + 1) * P.stride;
        uint64_t hit = ~uint64_t(0);
        for (uint32_t w = 0; w < P.window && hit == ~uint64_t(0); w += 32u) {
            uint64_t p = t0 + w + lane;
            const uint64_t cand = p;
            bool ok = p < P.in_bytes;
            for (uint32_t d = 0; ok && d < P.depth; ++d) {
                p = check_list(P, wt, p);
                ok = p != ~uint64_t(0);
                if (ok && p == P.in_bytes) break;  // the chain of lists reached the end of the file
            }
            const uint32_t m = __ballot_sync(FULL, ok);
            if (m) hit = shfl64(cand, __ffs(m) - 1);
        }
        if (lane == 0) P.found[t] = hit;
    }
}

// ---------------------------------------------------------------------------------------------------------
// frozen dict_freq_index: full blocks, docs streams (PASS 0) then freqs streams (PASS 1)
// ---------------------------------------------------------------------------------------------------------
struct index_walk_params {
    dint_dict_view dv;
    const uint32_t* sums;
    const uint8_t* lists;
    uint64_t lists_bytes;
    const dint_block* blocks;
    uint64_t n_blocks;
    uint32_t block_size, multi;
    uint32_t* docs_len;  // per block: bytes of its docs stream (PASS 0 writes, PASS 1 reads)
    dint_strip_unit* units;
    uint64_t units_cap;
    unsigned long long* counter;  // units written
    uint32_t* status;             // != 0: some block was malformed / the table overflowed
    uint32_t unit_ints, unit_bytes, strip_cw, strip_shift;
};

template <int PASS>
__global__ void __launch_bounds__(512, 2) index_walk_kernel(const index_walk_params P) {
    extern __shared__ __align__(16) uint8_t smem[];
    walk_tables wt;
    uint8_t* free0 = walk_setup(P.dv, P.sums, smem, wt);
    const unsigned lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint8_t* mine = free0 + size_t(warp) * kWalkWarpSmem;
    const uint32_t ring_s = smem_u32(mine), rec_s = ring_s + uint32_t(kInBytes), words_s = rec_s + kRecBytes;
    warp_state ws;
    walk_init_input(ws, ring_s, P.lists, P.lists_bytes);
    __syncthreads();
    const uint64_t in_base = reinterpret_cast<uint64_t>(P.lists);
    const uint64_t n_groups = (P.n_blocks + kIndexGroup - 1) / kIndexGroup;
    uint32_t status = WALK_OK;
    for (uint64_t g = uint64_t(blockIdx.x) * n_warps + warp; g < n_groups; g += uint64_t(gridDim.x) * n_warps) {
        unit_packer pk;
        pk.init(P.units, P.counter, 0, P.units_cap, words_s, P.unit_ints, P.unit_bytes);
        bool anchored = false;
        uint32_t prev_max = 0xFFFFFFFFu;
        const uint64_t i1 = min(P.n_blocks, (g + 1) * kIndexGroup);
        for (uint64_t i = g * kIndexGroup; i < i1; ++i) {
            const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(P.blocks + i));
            const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(P.blocks + i) + 1);
            const uint64_t in_off = (uint64_t(b0.y) << 32) | b0.x, out_off = (uint64_t(b0.w) << 32) | b0.z;
            const uint32_t size = b1.x, cur_base = b1.y, block_max = b1.z;
            if (size != P.block_size) continue;  // partial block: interpolative, index_tails_kernel
            const uint64_t off = in_off + (PASS ? P.docs_len[i] : 0u);
            if (off >= P.lists_bytes) {
                status = WALK_MALFORMED;
                continue;
            }
            const uint64_t g_addr = in_base + off;
            uint32_t p;
            if (!anchored || g_addr < ws.anchor || g_addr - ws.anchor > 0x60000000ull) {
                p = ws.set_anchor(g_addr, ~uint64_t(0));
                anchored = true;
            } else {
                p = uint32_t(g_addr - ws.anchor);
            }
            uint32_t sel = 0, a0 = p;
            if (P.multi) {
                ws.ensure_input(p, 16u, lane);
                sel = ws.rb(p);
                if (sel >= 2u * dint::kSelectors || (sel % dint::kSelectors) >= wt.n_classes) {
                    status = WALK_MALFORMED;
                    continue;
                }
                a0 = p + 1u;
            }
            // docid[0] = buf[0] + cur_base (dict_posting_list.hpp:304): the block's base is the running sum
            uint32_t acc = PASS == 0 ? cur_base : 0u, a_end = a0;
            const bool ls = PASS == 0 && cur_base == 0u;
            const bool brk = PASS == 0 && !ls && cur_base != prev_max + 1u;
            if (!walk_stream<true>(ws, wt, in_base, a0, size, sel, ls, brk, out_off, acc, pk, rec_s, P.strip_cw, P.strip_shift, lane, a_end)) {
                status = WALK_MALFORMED;
                pk.close<true>(lane);
                anchored = false;
                continue;
            }
            prev_max = block_max;
            const uint64_t end_off = ws.anchor + a_end - in_base;
            if (end_off > P.lists_bytes) status = WALK_MALFORMED;
            if (PASS == 0 && lane == 0) P.docs_len[i] = uint32_t(end_off - in_off);
        }
        pk.close<true>(lane);
        if (pk.overflow) status = WALK_TABLE_FULL;
    }
    if (status != WALK_OK && lane == 0) atomicMax(P.status, status);
    cp_async_wait_all();
}

}  // namespace dintk
