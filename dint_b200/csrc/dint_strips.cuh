// dint_b200 — the lane-sequential DINT decoder (device code, sm_100a).
//
// The reference decodes a list with a scalar loop: read a codeword, copy its dictionary entry
// to the output cursor, advance both cursors (vroom_env/dint_codecs.hpp:37-107 for u16
// codewords, :574-609 for u8 codewords). The only loop-carried values are the two cursors, so
// the loop parallelises over *pieces of the stream whose start cursors are known*. The host
// pre-pass (dint_vroom_plan_create) cuts every stream into STRIPS of at most 16 codewords and
// records where each strip starts in the input and in the output (include/dint_b200.h,
// dint_strip_unit). Here one THREAD runs the reference's loop over one strip, 32 strips (one
// STRIP UNIT) per warp, with no warp-level coordination inside the loop:
//
//   input      the unit's compressed bytes (<= 3.5 KB, contiguous in the file) are staged in
//              shared memory with 16-byte cp.async; every thread then reads its own strip with
//              aligned 4-byte shared loads + a funnel shift (any byte alignment).
//   length     4-bit log2(length) table, one slice per dictionary class indexed by the symbol;
//              the first lut_rows rows of every class live in shared memory, the rest (rare
//              symbols of multi_packed's u16 blocks) are read through L1 from global memory.
//   entry      16-byte vector loads from the expanded dictionary rows (L1/L2 resident).
//   output     an IMAGE of the unit's slice of the packed output array in shared memory
//              (<= 2048 integers), pre-zeroed, so runs of zeros only advance the cursor. Each
//              thread stores its entries at its own cursor.
//   flush      when all strips are done the warp reads the image back one uint4 per lane,
//              applies the fused transform (docID prefix sum with restarts at list starts, or
//              the freq +1 shift), writes 512 contiguous, 16-byte aligned bytes per
//              instruction with streaming stores, and re-zeroes the image.
//
// Work per codeword is ~50 warp instructions for 32 codewords; the warp-wide work (staging,
// flush) is amortised over a whole unit (~512 codewords).
#pragma once

#include "dint_decoder.cuh"

namespace dintk {

constexpr uint32_t kImagePad = 20;  // image words past unit_ints: output misalignment (3) + one entry (16)

// per-class constants, staged in shared memory (32 bytes)
struct strip_class {
    uint32_t row_base;   // first dictionary row of the class
    uint32_t rows;       // rows of the class
    uint32_t lut_s;      // shared address of the class's length-table slice
    uint32_t lut_g_off;  // byte offset of the class's slice in dint_dict_view::lens_cls
    uint32_t pad[4];
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
__host__ __device__ inline uint32_t strip_warp_smem(uint32_t unit_ints, uint32_t unit_bytes) {
    return strip_image_words(unit_ints) * 4u + strip_bits_words(unit_ints) * 4u + strip_inbuf_bytes(unit_bytes);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Writes the image [0, a + n_ints) of a unit to out[out_off - a ...] with the transform of MODE and
// re-zeroes it. a = out_off & 3, so that image quads are aligned quads of the output array.
template <int MODE>
__device__ __forceinline__ void flush_image(uint32_t image_s, uint32_t bits_s, uint32_t* __restrict__ out, uint64_t out_off,
                                            uint32_t n_ints, uint32_t doc_base, bool vec_ok, unsigned lane) {
    const uint32_t a = uint32_t(out_off) & 3u;
    const uint32_t total = a + n_ints;
    uint32_t* gq = out + (out_off - a) + 4u * lane;
    uint32_t acc = doc_base - 1u;  // docid[i] = acc + sum_{j <= i} (v_j + 1)
    for (uint32_t g0 = 0; g0 < total; g0 += 128u, gq += 128) {
        const uint32_t q = g0 + 4u * lane;
        const uint32_t sp = image_s + (q << 2);
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sp));
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sp), "r"(0u) : "memory");
        const uint32_t lo = g0 ? 0u : a;
        const uint32_t hi = min(128u, total - g0);
        uint32_t f = 0;  // bit i: element i of my quad starts a list
        bool heads = false;
        if (MODE == OUT_DOCIDS) {
            f = (lds_u32(bits_s + ((q >> 5) << 2)) >> (q & 31u)) & 15u;
            heads = __any_sync(FULL, f != 0u);
        }
        if (lo == 0u && hi == 128u && !heads && vec_ok) {
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
            st_stream4(gq, v);
        } else {
            const uint32_t e = 4u * lane;
            const bool l0 = e >= lo && e < hi, l1 = e + 1 >= lo && e + 1 < hi, l2 = e + 2 >= lo && e + 2 < hi,
                       l3 = e + 3 >= lo && e + 3 < hi;
            if (MODE == OUT_DOCIDS) {
                const uint32_t t0 = l0 ? v.x + 1u : 0u, t1 = l1 ? v.y + 1u : 0u, t2 = l2 ? v.z + 1u : 0u,
                               t3 = l3 ? v.w + 1u : 0u;
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
            } else if (MODE == OUT_FREQS) {
                v.x += 1u;
                v.y += 1u;
                v.z += 1u;
                v.w += 1u;
            }
            if (l0 && l3 && vec_ok) {
                st_stream4(gq, v);
            } else {
                if (l0) st_stream(gq, v.x);
                if (l1) st_stream(gq + 1, v.y);
                if (l2) st_stream(gq + 2, v.z);
                if (l3) st_stream(gq + 3, v.w);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(768, 1) decode_strips_kernel(const strip_params P) {
    extern __shared__ __align__(16) uint8_t smem[];
    const unsigned lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const dint_dict_view& dv = P.dv;

    // ---- CTA setup: length-table slices + class table -------------------------------------------
    const uint32_t slice_bytes = dv.lut_rows / 2u;  // multiple of 16
    {
        const uint32_t per = slice_bytes / 16u;
        for (uint32_t i = threadIdx.x; i < per * dv.n_classes; i += blockDim.x) {
            const uint32_t c = i / per, k = i - c * per;
            reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(dv.lens_cls + dv.lens_cls_off[c]) + k);
        }
    }
    const uint32_t lut_total = slice_bytes * dv.n_classes;
    strip_class* classes = reinterpret_cast<strip_class*>(smem + lut_total);
    if (threadIdx.x < 8) {
        const uint32_t c = threadIdx.x;
        strip_class ci = {0, 1, smem_u32(smem), 0, {0, 0, 0, 0}};
        if (c < dv.n_classes) {
            ci.row_base = dv.class_base[c];
            ci.rows = max(dv.class_rows[c], 1u);
            ci.lut_s = smem_u32(smem) + c * slice_bytes;
            ci.lut_g_off = dv.lens_cls_off[c];
        }
        classes[c] = ci;
    }
    const uint32_t classes_s = smem_u32(classes);
    uint8_t* mine = smem + lut_total + 8 * sizeof(strip_class) + size_t(warp) * (P.image_words * 4u + P.bits_words * 4u + P.inbuf_bytes);
    const uint32_t image_s = smem_u32(mine);
    const uint32_t bits_s = image_s + P.image_words * 4u;
    const uint32_t inbuf_s = bits_s + P.bits_words * 4u;
    for (uint32_t i = lane; i < (P.image_words + P.bits_words) / 4u; i += 32) reinterpret_cast<uint4*>(mine)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const uint64_t in_base = reinterpret_cast<uint64_t>(P.in);
    const uint64_t in_lo = in_base & ~uint64_t(15), in_hi = (in_base + P.in_bytes + 15) & ~uint64_t(15);
    const bool vec_ok = (reinterpret_cast<uint64_t>(P.out) & 15) == 0;
    const uint32_t lut_rows = dv.lut_rows;

    for (uint64_t u = uint64_t(blockIdx.x) * n_warps + warp; u < P.n_units; u += uint64_t(gridDim.x) * n_warps) {
        // ---- unit header (uniform loads) + my strip --------------------------------------------------
        const uint4* hp = reinterpret_cast<const uint4*>(P.units + u);
        const uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(hp + 2) + lane);
        const uint64_t in_off = (uint64_t(h0.y) << 32) | h0.x, out_off = (uint64_t(h0.w) << 32) | h0.z;
        const uint32_t doc_base = h1.x, starts = h1.y, in_len = h1.z & 0xFFFFu, n_ints = h1.z >> 16, n_strips = h1.w & 0xFFu;
        if (n_ints > P.unit_ints || in_len > P.unit_bytes || n_strips > 32u || in_off + in_len > P.in_bytes || n_ints == 0u)
            continue;  // not a unit this build can hold (warp-uniform)

        // ---- stage the unit's bytes ----------------------------------------------------------------------
        const uint64_t g = in_base + in_off, g0 = g & ~uint64_t(15);
        const uint32_t skew = uint32_t(g) & 15u;
        const uint32_t n_chunks = (skew + in_len + 12u + 15u) >> 4;  // a thread reads up to 12 bytes past its symbol
        for (uint32_t i = lane; i < n_chunks; i += 32) {
            const uint64_t ga = g0 + 16ull * i;
            const uint32_t sa = inbuf_s + 16u * i;
            if (ga >= in_lo && ga < in_hi) cp_async16(sa, ga);
            else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sa), "r"(0u) : "memory");
        }

        // ---- my strip ----------------------------------------------------------------------------------------
        const bool has = lane < n_strips;
        const uint32_t in_rel = min(word & 0x1FFFu, in_len), out_rel = min((word >> 13) & 0xFFFu, n_ints), sel = (word >> 25) & 15u;
        uint32_t nxt = __shfl_down_sync(FULL, out_rel, 1);
        if (lane + 1u >= n_strips) nxt = n_ints;
        const uint32_t n = (has && nxt > out_rel) ? nxt - out_rel : 0u;
        const uint32_t a = uint32_t(out_off) & 3u;
        uint32_t o = image_s + ((a + out_rel) << 2);
        const uint32_t o_end = o + (n << 2);
        uint32_t p = inbuf_s + skew + in_rel;
        const bool narrow = sel >= dint::kSelectors;
        const uint32_t cls = min(narrow ? sel - dint::kSelectors : sel, dv.n_classes - 1u);
        uint32_t c_row_base, c_rows, c_lut_s, c_lut_g;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(c_row_base), "=r"(c_rows), "=r"(c_lut_s), "=r"(c_lut_g)
                     : "r"(classes_s + cls * uint32_t(sizeof(strip_class))));
        const uint32_t symbits = narrow ? 8u : 16u, symmask = narrow ? 0xFFu : 0xFFFFu, symbytes = narrow ? 1u : 2u;
        const uint32_t last_row = (narrow ? min(c_rows, 256u) : c_rows) - 1u;
        const uint32_t* rowp = dv.rows + size_t(c_row_base) * dint::kEntry;
        const uint8_t* lutg = dv.lens_cls + c_lut_g;
        uint32_t start_word_s = 0;
        if (MODE == OUT_DOCIDS && has && lane && ((starts >> lane) & 1u)) {
            const uint32_t pos = a + out_rel;
            start_word_s = bits_s + ((pos >> 5) << 2);
            asm volatile("atom.shared.or.b32 _, [%0], %1;" ::"r"(start_word_s), "r"(1u << (pos & 31u)) : "memory");
        }
        cp_async_wait_all();
        __syncwarp();

        // ---- the reference's scalar loop, one strip per thread ------------------------------------------------
        bool act = o < o_end;
        while (__any_sync(FULL, act)) {
            const uint32_t a4 = p & ~3u, sh = (p & 3u) << 3;
            const uint32_t w0 = lds_u32(a4), w1 = lds_u32(a4 + 4);
            const uint32_t lo = __funnelshift_r(w0, w1, sh);
            const uint32_t sym = lo & symmask;
            uint32_t pay = lo >> symbits;  // what follows the codeword
            const bool exc = sym < 2u;     // 0: u16 value follows, 1: u32 value follows
            if (sym == 1u) pay |= __funnelshift_r(w1, lds_u32(a4 + 8), sh) << (32u - symbits);
            else pay &= 0xFFFFu;
            const uint32_t symc = min(sym, last_row);
            uint32_t lb;
            if (symc < lut_rows) lb = lds_u8(c_lut_s + (symc >> 1));
            else lb = __ldg(lutg + (symc >> 1));
            const uint32_t nib = (lb >> ((symc & 1u) << 2)) & 15u;  // rows 0 and 1 hold length 1, rows 2..6 the run lengths
            const uint4* rp = reinterpret_cast<const uint4*>(rowp + size_t(symc) * dint::kEntry);
            const bool dic = act && sym >= dint::kReserved;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (dic) r = __ldg(rp);
            if (exc) r.x = pay;
            const uint32_t L = dic ? (1u << nib) : 0u;  // dictionary words to copy (runs copy nothing: the image is zeroed)
            if (act && (exc || dic)) sts_u32(o, r.x);
            sts_tail3(o, L, r.y, r.z, r.w);
            if (L > 4u) {
                for (uint32_t k = 1; k < (L >> 2); ++k) sts_4(o + 16u * k, __ldg(rp + k));
            }
            if (act) {
                o += 4u << nib;
                p += symbytes + (exc ? (sym ? 4u : 2u) : 0u);
            }
            act = o < o_end;
        }
        __syncwarp();

        // ---- write the unit ------------------------------------------------------------------------------------
        flush_image<MODE>(image_s, bits_s, P.out, out_off, n_ints, doc_base, vec_ok, lane);
        if (MODE == OUT_DOCIDS && start_word_s) sts_u32(start_word_s, 0u);
        __syncwarp();
    }
}

}  // namespace dintk
