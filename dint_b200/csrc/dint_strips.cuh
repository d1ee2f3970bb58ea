// dint_b200 — the lane-sequential DINT decoder (device code, sm_100a).
//
// The reference decodes a list with a scalar loop: read a codeword, copy its dictionary entry
// to the output cursor, advance both cursors (vroom_env/dint_codecs.hpp:37-107 for u16
// codewords, :574-609 for u8 codewords). The only loop-carried values are the two cursors, so
// the loop parallelises over *pieces of the stream whose start cursors are known*. The host
// pre-pass (dint_vroom_plan_create) cuts every stream into STRIPS of at most 8 codewords and
// records where each strip starts in the input and in the output (include/dint_b200.h,
// dint_strip_unit). Here one THREAD runs the reference's loop over one strip, 32 strips (one
// STRIP UNIT) per warp, with no warp-level coordination inside the loop:
//
//   input      the unit's compressed bytes (<= 1.75 KB, contiguous in the file) are staged in
//              shared memory with 16-byte cp.async; every thread then reads its own strip with
//              aligned 4-byte shared loads + a funnel shift (any byte alignment), keeping the two
//              words under its cursor in registers from one codeword to the next.
//   length     one byte of log2(length) per row for the hot rows of every dictionary class
//              (the 256 rows u8 codewords can address; 2048 for a single dictionary) in shared
//              memory; colder rows (u16 codewords only) use a per-32-row group table.
//   entry      hot rows: the first 32 bytes of the row from shared memory; colder rows are
//              queued and copied from the L2-resident dictionary 32 codewords at a time.
//   output     an IMAGE of the unit's slice of the packed output array in shared memory
//              (<= 1024 integers by default). Each thread stores its entries at its own cursor. Run
//              codewords store nothing, they only advance the cursor: a unit that holds one
//              (DINT_UNIT_HAS_RUNS, set by the pre-pass) clears its image first.
//   flush      when all strips are done the warp reads the image back one uint4 per lane,
//              applies the fused transform (docID prefix sum with restarts at list starts, or
//              the freq +1 shift) and writes 512 contiguous, 16-byte aligned bytes per
//              instruction with streaming stores.
//
// The kernel is bound by the shared-memory data pipe (LSU wavefronts: every lane reads its own
// strip and writes at its own cursor, so most accesses pay 2-4 wavefronts), see profiles/.
#pragma once

#include "dint_decoder.cuh"

namespace dintk {

constexpr uint32_t kImagePad = 20;  // image words past unit_ints: output misalignment (3) + one entry (16)

// per-class constants, staged in shared memory (32 bytes)
struct strip_class {
    uint32_t hot0_s;     // shared address of the class's slice of hot0 (hot1 lies hot1_delta bytes further)
    uint32_t hotlen_s;   // shared address of the class's slice of hotlen
    uint32_t group_s;    // shared address of the class's group-length slice
    uint32_t rows;       // rows of the class (>= 1)
    uint32_t row_base;   // first dictionary row of the class
    uint32_t lut_g_off;  // byte offset of the class's slice in dint_dict_view::lens_cls
    uint32_t detail_s;   // shared address of the class's mixed-group detail (16 bytes of nibbles per group)
    uint32_t pad;
};

struct strip_params {
    dint_dict_view dv;
    const uint8_t* in;
    uint64_t in_bytes;
    const dint_strip_unit* units;
    uint64_t n_units;
    uint32_t* out;
    uint32_t unit_ints;    // capacity of the output image (integers)
    uint32_t unit_bytes;   // capacity of the input buffer (bytes spanned by a unit)
    uint32_t image_words;  // words of shared memory per image (multiple of 128)
    uint32_t bits_words;   // words of the list-start bitmap (multiple of 4)
    uint32_t inbuf_bytes;  // bytes of the input buffer (multiple of 16)
};

__host__ __device__ inline uint32_t strip_image_words(uint32_t unit_ints) { return (unit_ints + kImagePad + 127u) & ~127u; }
__host__ __device__ inline uint32_t strip_bits_words(uint32_t unit_ints) { return (strip_image_words(unit_ints) / 32u + 3u) & ~3u; }
__host__ __device__ inline uint32_t strip_inbuf_bytes(uint32_t unit_bytes) { return (unit_bytes + 15u + 12u + 15u + 16u) & ~15u; }
constexpr uint32_t kColdQueue = 64;  // entries of 8 bytes; a trip of the decode loop adds at most 32
__host__ __device__ inline uint32_t strip_warp_smem(uint32_t unit_ints, uint32_t unit_bytes) {
    return strip_image_words(unit_ints) * 4u + strip_bits_words(unit_ints) * 4u + kColdQueue * 8u + strip_inbuf_bytes(unit_bytes);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Writes the image [0, a + n_ints) of a unit to out[out_off - a ...] with the transform of MODE.
// a = out_off & 3, so that image quads are aligned quads of the output array.
template <int MODE, bool ZERO>
__device__ __forceinline__ void flush_image(uint32_t image_s, uint32_t bits_s, uint32_t* __restrict__ out, uint64_t out_off,
                                            uint32_t n_ints, uint32_t doc_base, bool unit_heads, bool vec_ok, unsigned lane) {
    const uint32_t a = uint32_t(out_off) & 3u;
    const uint32_t total = a + n_ints;
    uint32_t* gq = out + (out_off - a) + 4u * lane;
    // docid[i] = acc + sum_{j <= i} (v_j + 1). The a image words in front of the unit are read as zeros and
    // scanned like values (each adds 1), so they are discounted here; words past the unit's end cannot
    // change anything in front of them. Only groups that hold a list start need the segmented scan.
    uint32_t acc = doc_base - 1u - a;
    for (uint32_t g0 = 0; g0 < total; g0 += 128u, gq += 128) {
        const uint32_t q = g0 + 4u * lane;
        const uint32_t sp = image_s + (q << 2);
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sp));
        if (ZERO) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sp), "r"(0u) : "memory");
        const uint32_t lo = g0 ? 0u : a;
        if (!ZERO && MODE == OUT_DOCIDS && g0 == 0u && lane == 0u) {  // the words in front of the unit count as zeros
            if (a > 0u) v.x = 0u;
            if (a > 1u) v.y = 0u;
            if (a > 2u) v.z = 0u;
        }
        const uint32_t hi = min(128u, total - g0);
        if (MODE == OUT_DOCIDS) {
            uint32_t f = 0;  // bit i: element i of my quad starts a list
            bool heads = false;
            if (unit_heads) {
                f = (lds_u32(bits_s + ((q >> 5) << 2)) >> (q & 31u)) & 15u;
                heads = __any_sync(FULL, f != 0u);
            }
            if (!heads) {
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
            } else {
                // as in the plain scan, the zeros in front of the unit count 1 each (acc is discounted for them)
                // and whatever lies past its end cannot reach a stored value
                const uint32_t t0 = v.x + 1u, t1 = v.y + 1u, t2 = v.z + 1u, t3 = v.w + 1u;
                // lane-local sums restarting at list starts; r = plain lane total
                const uint32_t p0 = t0;
                const uint32_t p1 = (f & 2u) ? t1 : p0 + t1;
                const uint32_t p2 = (f & 4u) ? t2 : p1 + t2;
                const uint32_t p3 = (f & 8u) ? t3 : p2 + t3;
                const uint32_t r = t0 + t1 + t2 + t3;
                const uint32_t S = warp_inclusive_scan(r);
                // carry into this lane: from the closest earlier lane holding a list start, else from acc
                const uint32_t prevH = __ballot_sync(FULL, f != 0u) & lanemask_lt();
                const uint32_t hl = prevH ? 31u - uint32_t(__clz(prevH)) : 0u;
                const uint32_t from_head = __shfl_sync(FULL, p3 - S, hl);
                const uint32_t c = (prevH ? from_head - 1u : acc) + (S - r);
                v.x = (f & 1u) ? p0 - 1u : c + p0;
                v.y = (f & 3u) ? p1 - 1u : c + p1;
                v.z = (f & 7u) ? p2 - 1u : c + p2;
                v.w = (f & 15u) ? p3 - 1u : c + p3;
                acc = __shfl_sync(FULL, v.w, 31);
            }
        } else if (MODE == OUT_FREQS) {
            v.x += 1u;
            v.y += 1u;
            v.z += 1u;
            v.w += 1u;
        }
        if (lo == 0u && hi == 128u && vec_ok) {
            st_stream4(gq, v);
        } else {
            const uint32_t e = 4u * lane;
            const bool l0 = e >= lo && e < hi, l3 = e + 3 >= lo && e + 3 < hi;
            if (l0 && l3 && vec_ok) {
                st_stream4(gq, v);
            } else {
                if (l0) st_stream(gq, v.x);
                if (e + 1 >= lo && e + 1 < hi) st_stream(gq + 1, v.y);
                if (e + 2 >= lo && e + 2 < hi) st_stream(gq + 2, v.z);
                if (l3) st_stream(gq + 3, v.w);
            }
        }
    }
}

// FREQS, VARIANT bit 2: the image already holds the final values (every store of the decode loop adds the +1), so
// the flush is a copy: the aligned middle of the unit's slice goes out with ONE bulk shared -> global copy issued by
// lane 0 (cp.async.bulk: the copy engine reads the image, no LDS.128 / STG.128 per lane), the partial quads at its
// two ends with scalar stores. The caller waits for the copy to have READ the image before the next unit writes it.
__device__ __forceinline__ void flush_image_bulk(uint32_t image_s, uint32_t* __restrict__ out, uint64_t out_off, uint32_t n_ints,
                                                 unsigned lane) {
    const uint32_t a = uint32_t(out_off) & 3u;
    const uint32_t total = a + n_ints;
    uint32_t* g0 = out + (out_off - a);                 // image word w <-> g0[w]
    const uint32_t first = a ? min(4u, total) : 0u;     // words [a, first): head quad, partial
    const uint32_t last = max(first, total & ~3u);      // words [last, total): tail quad, partial
    if (lane < 4u) {
        if (lane >= a && lane < first) st_stream(g0 + lane, lds_u32(image_s + (lane << 2)));
        const uint32_t w = last + lane;
        if (w < total && w >= first) st_stream(g0 + w, lds_u32(image_s + (w << 2)));
    }
    // generic-proxy writes of every lane -> visible to the async proxy, then one lane issues the copy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0u && last > first) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g0 + first), "r"(image_s + (first << 2)),
                     "r"((last - first) << 2)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
__device__ __forceinline__ void bulk_flush_wait_read(unsigned lane) {
    if (lane == 0u) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
}

// predicated shared loads: the destination keeps its value when the predicate is false (no branch)
__device__ __forceinline__ void lds_u8_if(uint32_t& v, uint32_t a, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.u8 %0, [%1];\n\t}" : "+r"(v) : "r"(a), "r"(uint32_t(p)));
}
__device__ __forceinline__ void lds_v4_if(uint4& v, uint32_t a, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
                 : "r"(a), "r"(uint32_t(p)));
}
// a register quad whose contents do not matter yet (no instruction; avoids zeroing it every trip)
__device__ __forceinline__ uint4 undefined_v4() {
    uint4 v;
    asm volatile("" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// header words of a strip unit (two uint4 + this lane's strip word), see dint_strip_unit
struct unit_header {
    uint4 h0, h1;
    uint32_t word;
    __device__ __forceinline__ void load(const dint_strip_unit* units, uint64_t u, unsigned lane) {
        const uint4* hp = reinterpret_cast<const uint4*>(units + u);
        h0 = __ldg(hp);
        h1 = __ldg(hp + 1);
        word = __ldg(reinterpret_cast<const uint32_t*>(hp + 2) + lane);
    }
    __device__ __forceinline__ uint64_t in_off() const { return (uint64_t(h0.y) << 32) | h0.x; }
    __device__ __forceinline__ uint64_t out_off() const { return (uint64_t(h0.w) << 32) | h0.z; }
    __device__ __forceinline__ uint32_t in_len() const { return h1.z & 0xFFFFu; }
    __device__ __forceinline__ uint32_t n_ints() const { return h1.z >> 16; }
    __device__ __forceinline__ uint32_t n_strips() const { return h1.w & 0xFFu; }
    __device__ __forceinline__ bool has_runs() const { return (h1.w >> 8) & DINT_UNIT_HAS_RUNS; }
    // a unit this build can hold (warp-uniform); anything else is skipped, its output left untouched
    __device__ __forceinline__ bool ok(const strip_params& P) const {
        return !(n_ints() > P.unit_ints || in_len() > P.unit_bytes || n_strips() > 32u || in_off() + in_len() > P.in_bytes ||
                 n_ints() == 0u);
    }
};

constexpr uint32_t kStripClassBytes = 8u * uint32_t(sizeof(strip_class));

// VARIANT bit 0: keep the input words under the cursor in registers instead of reloading them every trip;
// bit 1: clear the image only for units with run codewords (else the flush always re-zeroes it). Both pay off
// on streams of short codewords without runs or exceptions (freqs: -9 %) and cost 3 % each on d-gap streams
// (half of the codewords are exceptions, half of the integers come from runs), see profiles/README.md.
template <int MODE, int VARIANT>
__global__ void __launch_bounds__(768, 1) decode_strips_kernel(const strip_params P) {
    constexpr bool REUSE = (VARIANT & 1) != 0, RUNFLAG = (VARIANT & 2) != 0;
    // bit 2 (FREQS with RUNFLAG only): +1 folded into the stores, bulk-copy flush (needs a 16-byte aligned output)
    constexpr bool BULK = (VARIANT & 4) != 0 && MODE == OUT_FREQS && RUNFLAG;
    constexpr uint32_t kFill = BULK ? 1u : 0u;  // what a run codeword's integers read as in the image
    extern __shared__ __align__(16) uint8_t smem[];
    const unsigned lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const dint_dict_view& dv = P.dv;

    // ---- CTA setup: dictionary image (hot rows, lengths) + class table ---------------------------
    for (uint32_t i = threadIdx.x; i < dv.strip_tables_bytes / 16u; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(dv.strip_tables) + i);
    const uint32_t hot_rows = dv.hot_rows;
    const uint32_t hot1_delta = dv.n_classes * hot_rows * 16u;
    strip_class* classes = reinterpret_cast<strip_class*>(smem + dv.strip_tables_bytes);
    if (threadIdx.x < 8) {
        const uint32_t c = min(threadIdx.x, dv.n_classes - 1u);
        strip_class ci;
        ci.hot0_s = smem_u32(smem) + c * hot_rows * 16u;
        ci.hotlen_s = smem_u32(smem) + 2u * hot1_delta + c * hot_rows;
        ci.group_s = smem_u32(smem) + dv.group_off[c];
        ci.rows = max(dv.class_rows[c], 1u);
        ci.row_base = dv.class_base[c];
        ci.lut_g_off = dv.lens_cls_off[c];
        ci.detail_s = smem_u32(smem) + dv.detail_off[c];
        ci.pad = 0;
        classes[threadIdx.x] = ci;
    }
    const uint32_t classes_s = smem_u32(classes);
    uint8_t* mine = smem + dv.strip_tables_bytes + kStripClassBytes +
                    size_t(warp) * (P.image_words * 4u + P.bits_words * 4u + kColdQueue * 8u + P.inbuf_bytes);
    const uint32_t image_s = smem_u32(mine);
    const uint32_t bits_s = image_s + P.image_words * 4u;
    const uint32_t queue_s = bits_s + P.bits_words * 4u;
    const uint32_t inbuf_s = queue_s + kColdQueue * 8u;
    for (uint32_t i = lane; i < (P.image_words + P.bits_words) / 4u; i += 32) reinterpret_cast<uint4*>(mine)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const uint64_t in_base = reinterpret_cast<uint64_t>(P.in);
    const uint64_t in_hi = (in_base + P.in_bytes + 15) & ~uint64_t(15);
    const bool vec_ok = (reinterpret_cast<uint64_t>(P.out) & 15) == 0;

    // copies the bytes of a unit into this warp's input buffer (16-byte cp.async; the caller waits). Aligned
    // 16-byte pieces that hold at least one byte of the file are read whole; the ones past its end are zeroed.
    auto stage = [&](uint64_t in_off, uint32_t in_len) {
        const uint64_t g = in_base + in_off, g0 = g & ~uint64_t(15);
        const uint32_t n_chunks = ((uint32_t(g) & 15u) + in_len + 12u + 15u) >> 4;  // a thread reads up to 12 bytes past its symbol
        const bool at_end = g0 + 16ull * n_chunks > in_hi;                         // only the last units of a file
        const uint32_t n_file = at_end ? uint32_t((in_hi - g0) >> 4) : n_chunks;
        const uint64_t ga = g0 + 16u * lane;
        const uint32_t sa = inbuf_s + 16u * lane;
        if (P.inbuf_bytes <= 4u * 512u) {  // default limits: four straight-line rounds
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                const uint32_t i = lane + 32u * j;
                if (i < n_file) cp_async16(sa + 512u * j, ga + 512u * j);
                else if (i < n_chunks) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sa + 512u * j), "r"(0u) : "memory");
            }
            return;
        }
        for (uint32_t i = lane, k = 0; i < n_chunks; i += 32, k += 512) {
            if (i < n_file) cp_async16(sa + k, ga + k);
            else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sa + k), "r"(0u) : "memory");
        }
    };

    const uint64_t stride = uint64_t(gridDim.x) * n_warps;
    uint64_t u = uint64_t(blockIdx.x) * n_warps + warp;
    unit_header H;
    H.h0 = H.h1 = make_uint4(0, 0, 0, 0);
    H.word = 0;
    bool valid = false;  // H passed ok() (checked once, when its bytes were requested)
    if (u < P.n_units) {
        H.load(P.units, u, lane);
        valid = H.ok(P);
        if (valid) stage(H.in_off(), H.in_len());
    }
    while (u < P.n_units) {
        // ---- the next unit's header is fetched while this one is decoded ---------------------------------
        const uint64_t un = u + stride;
        const bool more = un < P.n_units;
        unit_header N = H;
        if (more) N.load(P.units, un, lane);

        const uint64_t out_off = H.out_off();
        const uint32_t n_ints = H.n_ints();
        const uint32_t doc_base = H.h1.x, starts = H.h1.y;
        uint32_t start_word_s = 0;
        if (valid) {
            // ---- my strip ------------------------------------------------------------------------------------
            const uint32_t in_len = H.in_len(), n_strips = H.n_strips();
            const uint32_t skew = uint32_t(in_base + H.in_off()) & 15u;
            const bool has = lane < n_strips;
            const uint32_t word = H.word;
            const uint32_t in_rel = min(word & 0x1FFFu, in_len), out_rel = min((word >> 13) & 0xFFFu, n_ints), sel = (word >> 25) & 15u;
            uint32_t nxt = __shfl_down_sync(FULL, out_rel, 1);
            if (lane + 1u >= n_strips) nxt = n_ints;
            const uint32_t n = (has && nxt > out_rel) ? nxt - out_rel : 0u;
            const uint32_t a = uint32_t(out_off) & 3u;
            uint32_t o = image_s + ((a + out_rel) << 2);
            const uint32_t o_end = o + (n << 2);
            uint32_t p = inbuf_s + skew + in_rel;
            const uint32_t p_lim = inbuf_s + skew + in_len;
            const bool narrow = sel >= dint::kSelectors;
            const uint32_t cls = min(narrow ? sel - dint::kSelectors : sel, 7u);
            uint32_t hot0_s, hotlen_s, group_s, c_rows;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(hot0_s), "=r"(hotlen_s), "=r"(group_s), "=r"(c_rows)
                         : "r"(classes_s + cls * uint32_t(sizeof(strip_class))));
            const uint32_t c_row_base = lds_u32(classes_s + cls * uint32_t(sizeof(strip_class)) + 16u);
            const uint32_t c_lut_g = lds_u32(classes_s + cls * uint32_t(sizeof(strip_class)) + 20u);
            const uint32_t detail_s = lds_u32(classes_s + cls * uint32_t(sizeof(strip_class)) + 24u);
            const uint32_t symbits = narrow ? 8u : 16u, symmask = narrow ? 0xFFu : 0xFFFFu, symbytes = narrow ? 1u : 2u;
            const uint32_t last_row = (narrow ? min(c_rows, 256u) : c_rows) - 1u;
            const uint8_t* lutg = dv.lens_cls + c_lut_g;
            if (MODE == OUT_DOCIDS && has && lane && ((starts >> lane) & 1u)) {
                const uint32_t pos = a + out_rel;
                start_word_s = bits_s + ((pos >> 5) << 2);
                asm volatile("atom.shared.or.b32 _, [%0], %1;" ::"r"(start_word_s), "r"(1u << (pos & 31u)) : "memory");
            }
            if (BULK) bulk_flush_wait_read(lane);  // the previous unit's copy has read the image
            if (RUNFLAG && H.has_runs())  // run codewords store nothing: they rely on a cleared image
                for (uint32_t q = 4u * lane; q < a + n_ints; q += 128u)
                    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(image_s + (q << 2)), "r"(kFill) : "memory");
            cp_async_wait_all();
            __syncwarp();

            // ---- the reference's scalar loop, one strip per thread ------------------------------------------------
            // Lanes whose strip is finished keep stepping through whatever follows (no stores, addresses
            // clamped), so the cursor updates need no guard.
            // Entries beyond the hot rows (u16 codewords only) come from the L2-resident dictionary. Their
            // length is known at once (group table), so the cursors move on; the copy itself is queued
            // {image address | log2(length) << 24, dictionary row} and done for up to 64 codewords at a
            // time, so that one round trip to L2 serves all of them.
            uint32_t qn = 0;
            auto plus = [](uint4 v) {  // BULK: the image holds final freqs (v + 1)
                if (BULK) {
                    v.x += 1u;
                    v.y += 1u;
                    v.z += 1u;
                    v.w += 1u;
                }
                return v;
            };
            auto drain = [&]() {
                __syncwarp();
                for (uint32_t k = lane; k < qn; k += 32u) {
                    uint32_t ex, ey;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex), "=r"(ey) : "r"(queue_s + 8u * k));
                    const uint4* rp = reinterpret_cast<const uint4*>(dv.rows + size_t(ey) * dint::kEntry);
                    const uint4 r = plus(__ldg(rp));
                    const uint32_t oo = ex & 0xFFFFFFu, cc = ex >> 24;
                    sts_u32(oo, r.x);
                    sts_tail3_log2(oo, cc, r.y, r.z, r.w);
                    if (cc > 2u) {
                        sts_4(oo + 16u, plus(__ldg(rp + 1)));
                        if (cc > 3u) {
                            sts_4(oo + 32u, plus(__ldg(rp + 2)));
                            sts_4(oo + 48u, plus(__ldg(rp + 3)));
                        }
                    }
                }
                qn = 0;
                __syncwarp();
            };
            // length of a row beyond the hot rows + its place in the copy queue
            auto cold_row = [&](uint32_t sym, uint32_t oo, uint32_t cold_mask) -> uint32_t {
                const uint32_t symc = min(sym, last_row);
                uint32_t code = lds_u8(group_s + (symc >> 5));
                if (code > 8u) {  // the 32 rows of the group differ in length
                    if (code != 0xFFu) code = lds_u8(detail_s + ((code - 9u) << 4) + ((symc & 31u) >> 1));
                    else code = __ldg(lutg + (symc >> 1));
                    code = (code >> ((symc & 1u) << 2)) & 15u;
                }
                const uint32_t slot = queue_s + 8u * (qn + __popc(cold_mask & lanemask_lt()));
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot), "r"(oo | (code << 24)), "r"(c_row_base + symc) : "memory");
                return code;
            };
            // stores of one codeword: exception payload, or the entry's words (runs store nothing: the image is zeroed)
            auto put = [&](uint32_t oo, uint32_t sym, bool on, bool dichot, uint32_t code, uint4 r, uint32_t pay) {
                const uint32_t words = dichot ? code : 0u;
                if (sym < 2u) r.x = pay;
                r = plus(r);
                if (on && (sym < 2u || dichot)) sts_u32(oo, r.x);
                sts_tail3_log2(oo, words, r.y, r.z, r.w);
                if (words > 2u) {
                    sts_4(oo + 16u, plus(lds_v4(hot0_s + hot1_delta + (sym << 4))));
                    if (words > 3u) {
                        const uint4* rp = reinterpret_cast<const uint4*>(dv.rows + size_t(c_row_base + sym) * dint::kEntry);
                        sts_4(oo + 32u, plus(__ldg(rp + 2)));
                        sts_4(oo + 48u, plus(__ldg(rp + 3)));
                    }
                }
            };

            // One codeword per trip. The only loop-carried values are the two cursors; the input cursor depends
            // on nothing but the symbol, so the words under the NEXT symbol are requested before the table
            // lookups and the stores of this one (shared-memory stores would otherwise fence them).
            bool act = o < o_end;
            uint32_t w0 = lds_u32(p & ~3u), w1 = lds_u32((p & ~3u) + 4);
            while (__any_sync(FULL, act)) {
                const uint32_t a4 = p & ~3u, sh = p << 3;  // the funnel shift takes its count mod 32
                const uint32_t lo = __funnelshift_r(w0, w1, sh);
                const uint32_t sym = lo & symmask;
                uint32_t w2 = w1;
                if (sym == 1u) w2 = lds_u32(a4 + 8);
                const uint32_t pn = min(p + symbytes + (sym < 2u ? 2u + 2u * sym : 0u), p_lim);
                const uint32_t a4n = pn & ~3u;  // a4, a4 + 4 or a4 + 8: most trips reuse a word already in a register
                uint32_t n0 = w0, n1 = w1;
                if (!REUSE) {
                    n0 = lds_u32(a4n);
                    n1 = lds_u32(a4n + 4);
                } else if (a4n != a4) {
                    n0 = w1;
                    if (a4n != a4 + 4u) n0 = lds_u32(a4n);
                    n1 = lds_u32(a4n + 4);
                }
                const bool hot = sym < hot_rows;
                uint32_t code = 0;  // log2(entry length); rows 0 and 1 hold length 1, rows 2..6 the run lengths
                uint4 r = undefined_v4();  // only the words a predicate below selects are ever stored
                const bool dic = act && sym >= dint::kReserved;
                lds_u8_if(code, hotlen_s + sym, hot);
                lds_v4_if(r, hot0_s + (sym << 4), dic && hot);
                // what follows the codeword: 32 bits after a 1, 16 bits after a 0
                uint32_t pay = __funnelshift_r(lo, __funnelshift_r(w1, w2, sh), symbits);
                if (sym != 1u) pay &= 0xFFFFu;
                const bool cold = dic && !hot;
                const uint32_t cm = __ballot_sync(FULL, cold);
                if (cm) {  // warp-uniform
                    if (cold) code = cold_row(sym, o, cm);
                    qn += __popc(cm);
                    if (qn > kColdQueue - 32u) drain();  // the queue only grows here
                }
                put(o, sym, act, dic && hot, code, r, pay);
                o += 4u << code;
                p = pn;
                w0 = n0;
                w1 = n1;
                act = o < o_end;
            }
            if (qn) drain();
            __syncwarp();
        }

        // ---- the next unit's bytes travel while this one is written ---------------------------------------
        const bool n_valid = more && N.ok(P);
        if (n_valid) stage(N.in_off(), N.in_len());

        // ---- write the unit ------------------------------------------------------------------------------------
        if (valid && BULK) {
            flush_image_bulk(image_s, P.out, out_off, n_ints, lane);
        } else if (valid) {
            flush_image<MODE, !RUNFLAG>(image_s, bits_s, P.out, out_off, n_ints, doc_base, (starts & ~1u) != 0u, vec_ok, lane);
            if (MODE == OUT_DOCIDS && start_word_s) sts_u32(start_word_s, 0u);
            __syncwarp();
        }
        u = un;
        H = N;
        valid = n_valid;
    }
    cp_async_wait_all();
    if (BULK) {  // the last copy must be complete before the CTA (and its shared memory) goes away
        if (lane == 0u) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }
}

}  // namespace dintk
