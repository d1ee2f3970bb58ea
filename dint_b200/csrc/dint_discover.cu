// dint_b200 — host side of the device-side discovery (dint_discover.cuh): picks the chains, runs the
// count and emit passes, owns the resulting device-resident plan. Part of the C ABI (include/dint_b200.h).
//
// Replaces, for files that are already in HBM, the serial host walk of dint_vroom_plan_create: the only
// host work left is O(chains) bookkeeping (a prefix sum over a few thousand chain totals).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <memory>
#include <vector>

#include "dint_cuda_util.hpp"
#include "dint_discover.cuh"
#include "dint_index_blocks.cuh"

using namespace dintk;

namespace {

struct dev_buf {  // cold-path scratch, freed on scope exit
    void* p = nullptr;
    ~dev_buf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
};

uint32_t strip_shift_of(const dint::strip_limits& lim) {
    // a strip also ends where the output crosses a multiple of 2^shift integers, so that a strip of run codewords
    // (256 integers each) stays below unit_ints: (2^shift - 1) + 256 <= unit_ints
    uint32_t s = 9;
    while (s > 4 && (1u << s) + 255u > lim.unit_ints) --s;
    return s;
}

size_t walk_smem(const dint_dict* d, int warps) {
    dint_dict_view v = d->view();
    return size_t(walk_tables_bytes(v)) + kWalkClassBytes + size_t(warps) * kWalkWarpSmem;
}

template <typename K>
int walk_grid(K kernel, const dint_dict* d, int warps, uint64_t want_warps, int* grid, size_t* smem, size_t extra_smem = 0) {
    device_info di;
    int rc = get_device_info(d->device, &di);
    if (rc) return rc;
    *smem = walk_smem(d, warps) + extra_smem;
    if (*smem > size_t(di.max_smem_optin)) {
        dint::set_error("discovery: length tables do not fit in shared memory");
        return DINT_E_ARG;
    }
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(*smem)));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, *smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t want = (want_warps + warps - 1) / warps;
    *grid = int(std::max<uint64_t>(1, std::min<uint64_t>(want, uint64_t(di.sms) * per_sm)));
    return DINT_OK;
}

constexpr int kWalkWarps = 16;

// DINT_B200_TRACE=1: host timestamps of the discovery's steps on stderr (diagnostics)
struct tracer {
    bool on;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    tracer() : on(std::getenv("DINT_B200_TRACE") != nullptr), t0(now()) {}
    void operator()(const char* what) {
        if (on) std::fprintf(stderr, "[dint trace] %8.2f ms  %s\n", now() - t0, what);
    }
};

}  // namespace

extern "C" {

void dint_device_plan_destroy(dint_device_plan* p) {
    if (!p) return;
    {
        device_guard g(p->device);
        cudaFree(p->d_units);
        cudaFree(p->d_list_offsets);
        cudaFree(p->d_list_byte_offsets);
    }
    delete p;
}
uint64_t dint_device_plan_num_lists(const dint_device_plan* p) { return p ? p->n_lists : 0; }
uint64_t dint_device_plan_num_ints(const dint_device_plan* p) { return p ? p->n_ints : 0; }
uint64_t dint_device_plan_num_units(const dint_device_plan* p) { return p ? p->n_units : 0; }
uint64_t dint_device_plan_num_chains(const dint_device_plan* p) { return p ? p->n_chains : 0; }
uint64_t dint_device_plan_num_sync_found(const dint_device_plan* p) { return p ? p->n_sync_found : 0; }
uint64_t dint_device_plan_num_skimmed(const dint_device_plan* p) { return p ? p->n_skimmed : 0; }
uint32_t dint_device_plan_retries(const dint_device_plan* p) { return p ? p->retries : 0; }
double dint_device_plan_discover_ms(const dint_device_plan* p) { return p ? double(p->discover_ms) : 0.0; }
double dint_device_plan_phase_ms(const dint_device_plan* p, int phase) {
    return p && phase >= 0 && phase < 3 ? double(p->phase_ms[phase]) : 0.0;
}
const dint_strip_unit* dint_device_plan_units(const dint_device_plan* p) { return p ? p->d_units : nullptr; }
const uint64_t* dint_device_plan_list_offsets(const dint_device_plan* p) { return p ? p->d_list_offsets : nullptr; }
const uint64_t* dint_device_plan_list_byte_offsets(const dint_device_plan* p) { return p ? p->d_list_byte_offsets : nullptr; }

int dint_vroom_discover(const dint_dict* d, const uint8_t* d_in, size_t in_bytes, uint32_t block_size,
                        const uint64_t* sync_offsets, size_t n_sync, uint64_t sync_stride, void* stream,
                        dint_device_plan** out) {
    if (!d || !out || (!d_in && in_bytes) || (!sync_offsets && n_sync)) return DINT_E_ARG;
    if (!d->d_rows || !d->d_sums) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (!block_size) block_size = dint::kDefaultBlock;
    device_guard g(d->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    device_info di;
    int rc = get_device_info(d->device, &di);
    if (rc) return rc;
    std::unique_ptr<dint_device_plan, void (*)(dint_device_plan*)> plan(new (std::nothrow) dint_device_plan, dint_device_plan_destroy);
    if (!plan) return DINT_E_NOMEM;
    plan->device = d->device;
    plan->in_bytes = in_bytes;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs = nullptr, evc = nullptr;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventCreate(&evs));
    CUDA_TRY(cudaEventCreate(&evc));
    struct ev_guard {
        cudaEvent_t a, b, c, d;
        ~ev_guard() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
            cudaEventDestroy(c);
            cudaEventDestroy(d);
        }
    } evg{ev0, ev1, evs, evc};
    CUDA_TRY(cudaEventRecord(ev0, st));
    const dint::strip_limits& lim = dint::get_strip_limits();
    const bool multi = d->kind == DINT_DICT_MULTI;
    auto env_bytes = [](const char* name, uint64_t dflt) -> uint64_t {  // tuning knobs (diagnostics)
        const char* v = std::getenv(name);
        return v && *v ? std::strtoull(v, nullptr, 10) : dflt;
    };

    // ---- chain starts: byte 0, the caller's list starts, and what the speculative search finds ---------------
    std::vector<uint64_t> sync;
    uint64_t used_stride = 0;
    try {
        sync.push_back(0);
        for (size_t i = 0; i < n_sync; ++i)
            if (sync_offsets[i] > 0 && sync_offsets[i] < in_bytes) sync.push_back(sync_offsets[i]);
        if (sync_stride != ~uint64_t(0) && in_bytes > 0) {
            // default: about one chain per resident walking THREAD (1024 per SM), never closer than 4 KB
            uint64_t stride = sync_stride ? sync_stride : std::max<uint64_t>(4096, in_bytes / (uint64_t(di.sms) * 1024));
            stride = std::max<uint64_t>(stride, 1024);
            used_stride = stride;
            const uint64_t n_targets = (in_bytes - 1) / stride;
            if (n_targets) {
                dev_buf found;
                CUDA_TRY(cudaMalloc(&found.p, n_targets * 8));
                find_sync_params F;
                F.dv = d->view();
                F.sums = d->d_sums;
                F.in = d_in;
                F.in_bytes = in_bytes;
                F.block_size = block_size;
                F.multi = multi;
                F.stride = stride;
                F.n_targets = n_targets;
                F.window = uint32_t(std::min<uint64_t>(stride, 2048) & ~uint64_t(31));
                F.depth = uint32_t(env_bytes("DINT_B200_SYNC_DEPTH", 3));
                F.max_n = 256;  // evidence = three consecutive short lists (cheap to check; a Zipfian collection is full of them)
                F.found = found.as<uint64_t>();
                const size_t smem = walk_smem(d, 0);
                CUDA_TRY(cudaFuncSetAttribute(find_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                const int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((n_targets + 7) / 8, uint64_t(di.sms) * 4)));
                find_sync_kernel<<<grid, 256, smem, st>>>(F);
                ++g_launches;
                CUDA_TRY(cudaGetLastError());
                std::vector<uint64_t> h(n_targets);
                CUDA_TRY(cudaMemcpyAsync(h.data(), found.p, n_targets * 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                for (uint64_t v : h)
                    if (v != ~uint64_t(0) && v > 0 && v < in_bytes) {
                        sync.push_back(v);
                        ++plan->n_sync_found;
                    }
            }
        }
        std::sort(sync.begin(), sync.end());
        sync.erase(std::unique(sync.begin(), sync.end()), sync.end());
    } catch (const std::bad_alloc&) {
        return DINT_E_NOMEM;
    }

    CUDA_TRY(cudaEventRecord(evs, st));
    vroom_walk_params P;
    P.dv = d->view();
    P.sums = d->d_sums;
    P.in = d_in;
    P.in_bytes = in_bytes;
    P.block_size = block_size;
    P.multi = multi;
    P.unit_ints = lim.unit_ints;
    P.unit_bytes = lim.unit_bytes;
    P.strip_cw = lim.strip_codewords;
    P.strip_shift = strip_shift_of(lim);
    P.strip_ints = lim.strip_ints;
    P.units = nullptr;
    P.units_cap = 0;
    P.list_offsets = P.list_byte_offsets = nullptr;
    P.bases = nullptr;
    P.chain_state = nullptr;
    P.queue = nullptr;
    P.cut_off = nullptr;
    P.cuts = nullptr;
    P.cut_counts = nullptr;
    P.cut_bytes = 0;
    dev_buf d_status;  // u32 status | u32 - | u64 chain queue of the thread walker / the skim pass
    CUDA_TRY(cudaMalloc(&d_status.p, 16));
    P.status = d_status.as<uint32_t>();
    P.queue = reinterpret_cast<unsigned long long*>(d_status.as<uint32_t>() + 2);
    // chains sorted longest first in O(n): buckets of 512 bytes
    auto longest_first = [](std::vector<uint32_t>& ids, size_t lo, size_t hi, const std::vector<uint64_t>& se) {
        constexpr uint64_t kBins = 512;
        auto bin = [&](uint32_t slot) { return kBins - 1 - std::min<uint64_t>((se[slot + 1] - se[slot]) >> 9, kBins - 1); };
        std::vector<uint32_t> first(kBins + 1, 0), tmp(ids.begin() + lo, ids.begin() + hi);
        for (uint32_t v : tmp) ++first[bin(v) + 1];
        for (uint64_t b = 0; b < kBins; ++b) first[b + 1] += first[b];
        for (uint32_t v : tmp) ids[lo + first[bin(v)]++] = v;
    };

    // One chain = one candidate list start + where the walk from it has to stop. Chains up to kLongChain bytes are
    // walked by one thread each, longer ones (stretches without a sync point: very long lists) and the few chains of a
    // repair round by one warp each. A chain keeps its walker: the two passes must pack its units identically.
    // (measured: a thread walks ~2-3 KB/ms, a warp ~7 KB/ms, and the longest chain of either kind bounds the pass)
    // Long chains are first skimmed by a warp each (vroom_skim_kernel) and cut into pieces of about kCutBytes that
    // start at list headers or at block boundaries inside a list; threads walk the pieces. A long chain the skim
    // cannot cut (a long list of a single-stream codec, a malformed stretch) keeps its warp.
    // (measured: a thread walks ~2-3 KB/ms, a warp ~7 KB/ms, and the longest chain of either kind bounds the pass)
    // (gov2 sample x 256, 6.4 GB: long 48 KB / cut 16 KB: 130 ms; 63 / 21 KB: 158 ms; 96 / 32 KB: 182 ms; no skim: 234 ms)
    // (gov2 sample x 1, 25 MB, stride 4 KB: long 8 KB / cut 4 KB: 19 ms; 16 / 8 KB: 29 ms; 24 / 8 KB: 35 ms — a small
    // file is bound by its longest thread chain; the 1.5 GB freqs file x 256, stride 10 KB: 24 / 8 KB: 118 ms,
    // 11.5 / 4 KB: 144 ms — a large one by the bytes the skim pass has to walk)
    const uint64_t min_long = std::min<uint64_t>(std::max<uint64_t>(uint64_t(in_bytes) >> 16, 8 * 1024), 24 * 1024);
    const uint64_t kLongChain = env_bytes("DINT_B200_LONG_CHAIN", std::max<uint64_t>(min_long, std::min<uint64_t>(used_stride + used_stride / 8, 96 * 1024)));
    uint64_t kCutBytes = env_bytes("DINT_B200_CUT_BYTES", std::max<uint64_t>(std::max<uint64_t>(min_long / 3, 4 * 1024), std::min<uint64_t>(3 * used_stride / 8, 32 * 1024)));
    if (kCutBytes) kCutBytes = std::max<uint64_t>(kCutBytes, 64);  // (the cut table is sized by chain length / kCutBytes)
    struct chain {
        uint64_t start, end;             // walk from start to the first list start >= end
        uint64_t lists, ints, units, reached;
        uint32_t status;
        bool by_warp, counted;
        // a piece of a skimmed chain: integers of the current list still to come at start (0: a list header),
        // the list's sum(v + 1) so far, whether the piece ends inside a list and the integers to come there
        uint32_t rem0 = 0, acc0 = 0, end_rem = 0;
        bool end_mid = false;
    };
    // ---- one pass over chains[ids]: device arrays are indexed by the position k in ids (slot 2k / 8k) ----------------
    struct pass_bufs {
        dev_buf d_se, d_cnt, d_order, d_bases, d_state;
        uint64_t n = 0, n_short = 0, n_long = 0;
    };
    auto prepare = [&](const std::vector<chain>& chains, const std::vector<uint32_t>& ids, pass_bufs& pb) -> int {
        const uint64_t n = pb.n = ids.size();
        pb.n_short = pb.n_long = 0;
        if (!n) return DINT_OK;
        std::vector<uint64_t> se(2 * n + 2);
        std::vector<uint32_t> order(n);
        std::vector<uint32_t> state(4 * n);
        // walk_params wants chain c = [chain_start[c], chain_start[c + 1]): give every chain its own pair
        for (uint64_t k = 0; k < n; ++k) {
            const chain& c = chains[ids[k]];
            se[2 * k] = c.start;
            se[2 * k + 1] = c.end;
            state[4 * k] = c.rem0;
            state[4 * k + 1] = c.acc0;
            state[4 * k + 2] = c.end_mid ? 1u : 0u;
            state[4 * k + 3] = c.end_rem;
        }
        for (uint64_t k = 0; k < n; ++k)
            if (!chains[ids[k]].by_warp) order[pb.n_short++] = uint32_t(2 * k);
        for (uint64_t k = 0; k < n; ++k)
            if (chains[ids[k]].by_warp) order[pb.n_short + pb.n_long++] = uint32_t(2 * k);
        longest_first(order, 0, pb.n_short, se);
        CUDA_TRY(cudaMalloc(&pb.d_se.p, se.size() * 8));
        CUDA_TRY(cudaMalloc(&pb.d_cnt.p, 8 * n * 8));
        CUDA_TRY(cudaMalloc(&pb.d_bases.p, 8 * n * 8));
        CUDA_TRY(cudaMalloc(&pb.d_order.p, n * 4));
        CUDA_TRY(cudaMalloc(&pb.d_state.p, n * 16));
        CUDA_TRY(cudaMemcpyAsync(pb.d_state.p, state.data(), n * 16, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(pb.d_se.p, se.data(), se.size() * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(pb.d_order.p, order.data(), n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));  // se / order go out of scope
        return DINT_OK;
    };
    auto launch = [&](pass_bufs& pb, bool emit) -> int {
        if (!pb.n) return DINT_OK;
        P.chain_start = pb.d_se.as<uint64_t>();
        P.n_chains = 2 * pb.n;
        P.counts = pb.d_cnt.as<uint64_t>();
        P.bases = pb.d_bases.as<uint64_t>();
        P.chain_state = pb.d_state.as<uint4>();
        CUDA_TRY(cudaMemsetAsync(d_status.p, 0, 16, st));
        if (pb.n_short) {
            P.chain_ids = pb.d_order.as<uint32_t>();
            P.n_ids = pb.n_short;
            const size_t smem = walk_smem(d, 0) + (emit ? walk_hotsum_bytes(d->view()) : 0);
            // persistent CTAs (the threads take chains from the queue): 3 per SM fit the emit pass's registers
            const uint64_t ctas = emit ? 3 : env_bytes("DINT_B200_COUNT_CTAS", 3);
            const int grid = int(std::min<uint64_t>((pb.n_short + 255) / 256, uint64_t(di.sms) * ctas));
            if (emit) {
                CUDA_TRY(cudaFuncSetAttribute(vroom_thread_walk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                vroom_thread_walk_kernel<true><<<grid, 256, smem, st>>>(P);
            } else {
                CUDA_TRY(cudaFuncSetAttribute(vroom_thread_walk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                vroom_thread_walk_kernel<false><<<grid, 256, smem, st>>>(P);
            }
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
        }
        if (pb.n_long) {
            P.chain_ids = pb.d_order.as<uint32_t>() + pb.n_short;
            P.n_ids = pb.n_long;
            int grid = 0, rc2;
            size_t smem = 0;
            if (emit) {
                rc2 = walk_grid(vroom_walk_kernel<true>, d, kWalkWarps, pb.n_long, &grid, &smem);
                if (rc2) return rc2;
                vroom_walk_kernel<true><<<grid, kWalkWarps * 32, smem, st>>>(P);
            } else {
                rc2 = walk_grid(vroom_walk_kernel<false>, d, kWalkWarps, pb.n_long, &grid, &smem);
                if (rc2) return rc2;
                vroom_walk_kernel<false><<<grid, kWalkWarps * 32, smem, st>>>(P);
            }
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
        }
        return DINT_OK;
    };
    auto fetch = [&](pass_bufs& pb, std::vector<chain>& chains, const std::vector<uint32_t>& ids) -> int {
        const uint64_t n = pb.n;
        if (!n) return DINT_OK;
        std::vector<uint64_t> h(8 * n);
        CUDA_TRY(cudaMemcpyAsync(h.data(), pb.d_cnt.p, 8 * n * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (uint64_t k = 0; k < n; ++k) {
            chain& c = chains[ids[k]];
            c.lists = h[8 * k];
            c.ints = h[8 * k + 1];
            c.units = h[8 * k + 2];
            c.status = uint32_t(h[8 * k + 3] & 0xFF);
            c.reached = h[8 * k + 3] >> 8;
            c.counted = true;
        }
        return DINT_OK;
    };
    tracer trace0;
    tracer* tr_ptr = &trace0;
    // count pass of a repair round / emit pass with host-computed bases (the slow path)
    auto run_pass = [&](std::vector<chain>& chains, const std::vector<uint32_t>& ids, bool emit, const std::vector<uint64_t>* bases) -> int {
        pass_bufs pb;
        int rc2 = prepare(chains, ids, pb);
        if (rc2 || !pb.n) return rc2;
        const uint64_t n = pb.n;
        if (emit) tr_ptr->operator()("  emit: prepared");
        if (emit) {  // counts of the count pass (checked by the emit pass) + bases, both at slot 8k
            std::vector<uint64_t> wide(8 * n, 0), b(8 * n, 0);
            for (uint64_t k = 0; k < n; ++k) {
                const chain& c = chains[ids[k]];
                wide[8 * k] = c.lists;
                wide[8 * k + 1] = c.ints;
                wide[8 * k + 2] = c.units;
                for (int j = 0; j < 3; ++j) b[8 * k + j] = (*bases)[4 * ids[k] + j];
            }
            CUDA_TRY(cudaMemcpyAsync(pb.d_cnt.p, wide.data(), 8 * n * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(pb.d_bases.p, b.data(), 8 * n * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (emit) tr_ptr->operator()("  emit: uploaded");
        rc2 = launch(pb, emit);
        if (rc2) return rc2;
        if (!emit) return fetch(pb, chains, ids);
        tr_ptr->operator()("  emit: launched");
        CUDA_TRY(cudaStreamSynchronize(st));
        tr_ptr->operator()("  emit: synced");
        return DINT_OK;
    };
    auto finish = [&](uint64_t lists, uint64_t ints) -> int {
        const uint64_t tail[2] = {ints, uint64_t(in_bytes)};
        CUDA_TRY(cudaMemcpyAsync(plan->d_list_offsets + lists, &tail[0], 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(plan->d_list_byte_offsets + lists, &tail[1], 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(ev1, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaEventElapsedTime(&plan->discover_ms, ev0, ev1));
        CUDA_TRY(cudaEventElapsedTime(&plan->phase_ms[0], ev0, evs));
        CUDA_TRY(cudaEventElapsedTime(&plan->phase_ms[1], evs, evc));
        CUDA_TRY(cudaEventElapsedTime(&plan->phase_ms[2], evc, ev1));
        return DINT_OK;
    };
    auto allocate_outputs = [&](uint64_t lists, uint64_t ints, uint64_t units, uint64_t n_chains) -> int {
        plan->n_lists = lists;
        plan->n_ints = ints;
        plan->n_units = units;
        plan->n_chains = n_chains;
        CUDA_TRY(cudaMalloc(&plan->d_units, (units + 1) * sizeof(dint_strip_unit)));
        CUDA_TRY(cudaMalloc(&plan->d_list_offsets, (lists + 1) * 8));
        CUDA_TRY(cudaMalloc(&plan->d_list_byte_offsets, (lists + 1) * 8));
        P.units = plan->d_units;
        P.units_cap = units;
        P.list_offsets = plan->d_list_offsets;
        P.list_byte_offsets = plan->d_list_byte_offsets;
        return DINT_OK;
    };

    tracer tr;
    try {
        // ---- count pass + repair rounds ------------------------------------------------------------------------------
        // Every candidate start gets a chain that runs to the next candidate. A chain that starts at a true list start
        // only ever stops at true list starts, so the file is then re-linked from byte 0: chain -> where it stopped ->
        // the chain that starts there. A stop that no chain starts at (its target was a wrong guess of the speculative
        // search, or a wrong hint) becomes a new chain, walked in the next round; chains nobody lands on are dropped.
        if (in_bytes == 0) sync.clear();
        std::vector<chain> chains;
        chains.reserve(sync.size() + 64);
        for (size_t i = 0; i < sync.size(); ++i) {
            const uint64_t e = i + 1 < sync.size() ? sync[i + 1] : uint64_t(in_bytes);
            chains.push_back(chain{sync[i], e, 0, 0, 0, 0, 0, e - sync[i] > kLongChain, false});
        }
        // ---- skim pass: long chains -> pieces -------------------------------------------------------------------------
        std::vector<uint32_t> long_ids;
        for (size_t i = 0; i < chains.size(); ++i)
            if (chains[i].by_warp) long_ids.push_back(uint32_t(i));
        if (!long_ids.empty() && kCutBytes) {
            const uint64_t nl = long_ids.size();
            std::vector<uint64_t> se(2 * nl + 2), cut_off(nl + 1, 0);
            std::vector<uint32_t> order(nl);
            for (uint64_t j = 0; j < nl; ++j) {
                const chain& c = chains[long_ids[j]];
                se[2 * j] = c.start;
                se[2 * j + 1] = c.end;
                order[j] = uint32_t(2 * j);
                cut_off[j + 1] = cut_off[j] + (c.end - c.start) / kCutBytes + 1;
            }
            longest_first(order, 0, nl, se);
            dev_buf d_se, d_order, d_off, d_cuts, d_cnt;
            CUDA_TRY(cudaMalloc(&d_se.p, se.size() * 8));
            CUDA_TRY(cudaMalloc(&d_order.p, nl * 4));
            CUDA_TRY(cudaMalloc(&d_off.p, (nl + 1) * 8));
            CUDA_TRY(cudaMalloc(&d_cuts.p, cut_off[nl] * 16));
            CUDA_TRY(cudaMalloc(&d_cnt.p, nl * 4));
            CUDA_TRY(cudaMemcpyAsync(d_se.p, se.data(), se.size() * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_order.p, order.data(), nl * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_off.p, cut_off.data(), (nl + 1) * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemsetAsync(d_cnt.p, 0, nl * 4, st));
            CUDA_TRY(cudaMemsetAsync(d_status.p, 0, 16, st));
            P.chain_start = d_se.as<uint64_t>();
            P.n_chains = 2 * nl;
            P.chain_ids = d_order.as<uint32_t>();
            P.n_ids = nl;
            P.cut_off = d_off.as<uint64_t>();
            P.cuts = d_cuts.as<uint64_t>();
            P.cut_counts = d_cnt.as<uint32_t>();
            P.cut_bytes = kCutBytes;
            int grid = 0;
            size_t smem = 0;
            rc = walk_grid(vroom_skim_kernel, d, kWalkWarps, nl, &grid, &smem, walk_hotsum_bytes(d->view()));
            if (rc) return rc;
            vroom_skim_kernel<<<grid, kWalkWarps * 32, smem, st>>>(P);
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
            std::vector<uint32_t> n_cuts(nl);
            std::vector<uint64_t> cuts(2 * cut_off[nl]);
            CUDA_TRY(cudaMemcpyAsync(n_cuts.data(), d_cnt.p, nl * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(cuts.data(), d_cuts.p, cuts.size() * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            P.cut_off = nullptr;
            P.cuts = nullptr;
            P.cut_counts = nullptr;
            std::vector<chain> pieces;
            pieces.reserve(chains.size() + cut_off[nl]);
            size_t j = 0;
            for (size_t i = 0; i < chains.size(); ++i) {
                if (j >= nl || long_ids[j] != i) {
                    pieces.push_back(chains[i]);
                    continue;
                }
                const chain whole = chains[i];
                const uint64_t* cj = cuts.data() + 2 * cut_off[j];
                const uint32_t m = std::min<uint64_t>(n_cuts[j], cut_off[j + 1] - cut_off[j]);
                ++j;
                chain cur = whole;  // [start, first cut), then [cut, next cut) ..., the last piece keeps the chain's end
                for (uint32_t k = 0; k < m; ++k) {
                    const uint64_t at = cj[2 * k];
                    const uint32_t rem = uint32_t(cj[2 * k + 1]), acc = uint32_t(cj[2 * k + 1] >> 32);
                    if (at <= cur.start || at >= whole.end) break;  // (cuts are strictly increasing inside the chain)
                    cur.end = at;
                    cur.end_mid = rem != 0;
                    cur.end_rem = rem;
                    // (a warp starts and stops at list headers only: a long list of a single-stream codec keeps its warp)
                    cur.by_warp = cur.rem0 == 0 && !cur.end_mid && cur.end - cur.start > kLongChain;
                    pieces.push_back(cur);
                    cur = whole;
                    cur.start = at;
                    cur.rem0 = rem;
                    cur.acc0 = acc;
                }
                cur.by_warp = cur.rem0 == 0 && cur.end - cur.start > kLongChain;
                pieces.push_back(cur);
            }
            chains.swap(pieces);
            plan->n_skimmed = nl;
            tr("skimmed");
        }
        std::vector<uint32_t> todo(chains.size());
        for (size_t i = 0; i < chains.size(); ++i) todo[i] = uint32_t(i);
        std::vector<uint32_t> linked;  // the chains of the file, in order
        // ---- fast path: everything stays on the device ------------------------------------------------------------------
        // count pass -> chain_scan_kernel (every chain landed on its successor? + decoupled look-back scan of the chain
        // totals into chain bases) -> the host reads back four numbers, sizes the tables -> emit pass
        bool first_counted = false;
        if (!chains.empty()) {
            pass_bufs pb;
            rc = prepare(chains, todo, pb);
            if (rc) return rc;
            tr("chains uploaded");
            rc = launch(pb, false);
            if (rc) return rc;
            const uint64_t n = pb.n, tiles = (n + 255) / 256;
            dev_buf d_scan;  // tile_agg [tiles][4] | tile_inc [tiles][4] | totals [4] | flags [tiles] | counter
            const size_t scan_bytes = (8 * tiles + 4) * 8 + (tiles + 4) * 4;
            CUDA_TRY(cudaMalloc(&d_scan.p, scan_bytes));
            CUDA_TRY(cudaMemsetAsync(d_scan.p, 0, scan_bytes, st));
            chain_scan_params S;
            S.n = n;
            S.se = pb.d_se.as<uint64_t>();
            S.cnt = pb.d_cnt.as<uint64_t>();
            S.bases = pb.d_bases.as<uint64_t>();
            S.tile_agg = d_scan.as<unsigned long long>();
            S.tile_inc = S.tile_agg + 4 * tiles;
            S.totals = S.tile_inc + 4 * tiles;
            S.tile_flag = reinterpret_cast<unsigned int*>(S.totals + 4);
            S.tile_counter = S.tile_flag + tiles;
            chain_scan_kernel<<<int(tiles), 256, 0, st>>>(S);
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
            uint64_t totals[4] = {0, 0, 0, 0};
            CUDA_TRY(cudaMemcpyAsync(totals, S.totals, 32, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            tr("counted");
            if (totals[3] == 0) {
                CUDA_TRY(cudaEventRecord(evc, st));
                rc = allocate_outputs(totals[0], totals[1], totals[2], n);
                if (rc) return rc;
                rc = launch(pb, true);
                if (rc) return rc;
                uint32_t h_status = 0;
                CUDA_TRY(cudaMemcpyAsync(&h_status, d_status.p, 4, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                if (h_status != WALK_OK) {
                    dint::set_error("vroom .bin: the emit pass of the device walk disagrees with the count pass");
                    return DINT_E_FORMAT;
                }
                rc = finish(totals[0], totals[1]);
                if (rc) return rc;
                tr("emitted");
                *out = plan.release();
                return DINT_OK;
            }
            // some chain did not land on its successor (or is malformed): repair on the host, from these counts
            tr("fast path refused");
            rc = fetch(pb, chains, todo);
            if (rc) return rc;
            first_counted = true;
            tr("counts fetched");
        }
        for (uint32_t round = 0;; ++round) {
            if (!(round == 0 && first_counted)) {
                rc = run_pass(chains, todo, false, nullptr);
                if (rc) return rc;
            }
            todo.clear();
            // index of the chain that starts at a given byte: candidates are sorted by start, repairs are appended
            std::vector<std::pair<uint64_t, uint32_t>> by_start(chains.size());
            for (size_t i = 0; i < chains.size(); ++i) by_start[i] = {chains[i].start, uint32_t(i)};
            {  // the candidates are in file order, only the repairs appended behind them are not
                const auto mid = std::is_sorted_until(by_start.begin(), by_start.end());
                std::sort(mid, by_start.end());
                std::inplace_merge(by_start.begin(), mid, by_start.end());
            }
            // the chain that continues at byte `off` behind chain `prev` (-1: none): a chain that starts at a list header
            // if prev stopped at a list boundary; the next piece of the same skimmed chain if prev stopped inside a list
            auto find = [&](uint64_t off, int64_t prev) -> int64_t {
                const bool mid = prev >= 0 && chains[size_t(prev)].end_mid;
                for (auto it = std::lower_bound(by_start.begin(), by_start.end(), std::make_pair(off, uint32_t(0)));
                     it != by_start.end() && it->first == off; ++it) {
                    const chain& c = chains[it->second];
                    if (!mid && c.rem0 == 0) return int64_t(it->second);
                    if (mid && c.rem0 != 0 && uint32_t(prev) + 1 == it->second) return int64_t(it->second);
                }
                return -1;
            };
            auto next_start = [&](uint64_t off) -> uint64_t {  // first candidate list start > off
                auto it = std::upper_bound(by_start.begin(), by_start.end(), std::make_pair(off, ~uint32_t(0)));
                while (it != by_start.end() && chains[it->second].rem0 != 0) ++it;
                return it != by_start.end() ? it->first : uint64_t(in_bytes);
            };
            // every overshoot points at a list start nobody covers yet (if the chain itself is a true one): repair all
            // of them in one round, not only the first one met from byte 0
            for (size_t i = 0; i < chains.size(); ++i) {
                const chain& c = chains[i];
                // (a piece that was to stop inside a list and did not is not a true one: nothing to repair behind it)
                if (c.counted && c.status == WALK_OVERSHOOT && !c.end_mid && c.reached < in_bytes && find(c.reached, -1) < 0) {
                    const uint64_t e = next_start(c.reached);
                    chains.push_back(chain{c.reached, e, 0, 0, 0, 0, 0, true, false});
                    by_start.insert(std::upper_bound(by_start.begin(), by_start.end(), std::make_pair(c.reached, uint32_t(0))),
                                    {c.reached, uint32_t(chains.size() - 1)});
                    todo.push_back(uint32_t(chains.size() - 1));
                }
            }
            if (!todo.empty()) {
                if (round >= 64) {
                    dint::set_error("vroom .bin: the device walk did not converge");
                    return DINT_E_FORMAT;
                }
                ++plan->retries;
                continue;
            }
            // link from byte 0
            linked.clear();
            uint64_t off = 0;
            int64_t prev = -1;
            bool complete = true;
            while (off < in_bytes) {
                int64_t i = -1;
                if (prev >= 0 && size_t(prev) + 1 < chains.size() && chains[size_t(prev) + 1].start == off) {
                    // the usual case: the next candidate in file order (same acceptance rule as find)
                    const bool mid = chains[size_t(prev)].end_mid;
                    if (mid == (chains[size_t(prev) + 1].rem0 != 0)) i = prev + 1;
                }
                if (i < 0) i = find(off, prev);
                if (i < 0 && prev >= 0 && chains[size_t(prev)].end_mid) {
                    dint::set_error("vroom .bin: malformed list (device walk), chain at byte " + std::to_string(off));
                    return DINT_E_FORMAT;
                }
                if (i < 0) {  // cannot happen after the repairs above; be safe
                    chains.push_back(chain{off, next_start(off), 0, 0, 0, 0, 0, true, false});
                    todo.push_back(uint32_t(chains.size() - 1));
                    complete = false;
                    break;
                }
                const chain& c = chains[size_t(i)];
                if (c.status == WALK_MALFORMED || c.status == WALK_TABLE_FULL || c.reached <= off) {
                    dint::set_error("vroom .bin: malformed list (device walk), chain at byte " + std::to_string(off));
                    return DINT_E_FORMAT;
                }
                linked.push_back(uint32_t(i));
                off = c.reached;
                prev = i;
            }
            if (complete) break;
            if (round >= 64) {
                dint::set_error("vroom .bin: the device walk did not converge");
                return DINT_E_FORMAT;
            }
            ++plan->retries;
        }
        tr("linked");
        CUDA_TRY(cudaEventRecord(evc, st));
        // ---- bases of every linked chain (host: one pass over the chain totals) ---------------------------------------
        std::vector<uint64_t> bases(chains.size() * 4 + 4, 0);
        uint64_t lists = 0, ints = 0, units = 0;
        for (uint32_t i : linked) {
            chain& c = chains[i];
            c.end = c.reached;  // the emit pass stops exactly where the count pass did
            bases[4 * size_t(i)] = lists;
            bases[4 * size_t(i) + 1] = ints;
            bases[4 * size_t(i) + 2] = units;
            lists += c.lists;
            ints += c.ints;
            units += c.units;
        }
        tr("bases");
        rc = allocate_outputs(lists, ints, units, linked.size());
        if (rc) return rc;
        tr("allocated");
        if (!linked.empty()) {
            rc = run_pass(chains, linked, true, &bases);
            if (rc) return rc;
            tr("emit pass done");
            uint32_t h_status = 0;
            CUDA_TRY(cudaMemcpy(&h_status, d_status.p, 4, cudaMemcpyDeviceToHost));
            if (h_status != WALK_OK) {
                dint::set_error("vroom .bin: the emit pass of the device walk disagrees with the count pass");
                return DINT_E_FORMAT;
            }
        }
        rc = finish(lists, ints);
        if (rc) return rc;
    } catch (const std::bad_alloc&) {
        return DINT_E_NOMEM;
    }
    *out = plan.release();
    return DINT_OK;
}

int dint_index_discover(const dint_dict* docs_dict, const dint_dict* freqs_dict, const uint8_t* d_lists, size_t lists_bytes,
                        const dint_block* d_blocks, size_t n_blocks, uint32_t block_size, void* stream,
                        dint_device_plan** docs_plan, dint_device_plan** freqs_plan) {
    if (!docs_dict || !freqs_dict || !docs_plan || !freqs_plan) return DINT_E_ARG;
    if (!docs_dict->d_rows || !freqs_dict->d_rows || !docs_dict->d_sums || !freqs_dict->d_sums) {
        dint::set_error("dictionary is not device resident (no CPU fallback)");
        return DINT_E_NODEVICE;
    }
    if (docs_dict->device != freqs_dict->device || docs_dict->kind != freqs_dict->kind) return DINT_E_ARG;
    if (n_blocks && (!d_lists || !d_blocks || (reinterpret_cast<uintptr_t>(d_blocks) & 15))) return DINT_E_ARG;
    if (!block_size) block_size = dint::kDefaultBlock;
    device_guard g(docs_dict->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dint::strip_limits& lim = dint::get_strip_limits();
    if (block_size > lim.unit_ints) {
        dint::set_error("index discovery: block_size exceeds the strip decoder's unit size");
        return DINT_E_ARG;
    }
    using plan_ptr = std::unique_ptr<dint_device_plan, void (*)(dint_device_plan*)>;
    plan_ptr plans[2] = {plan_ptr(new (std::nothrow) dint_device_plan, dint_device_plan_destroy),
                         plan_ptr(new (std::nothrow) dint_device_plan, dint_device_plan_destroy)};
    if (!plans[0] || !plans[1]) return DINT_E_NOMEM;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    struct ev_guard {
        cudaEvent_t* e;
        ~ev_guard() {
            for (int i = 0; i < 3; ++i) cudaEventDestroy(e[i]);
        }
    } evg{ev};
    dev_buf d_docs_len, d_ctr;
    CUDA_TRY(cudaMalloc(&d_docs_len.p, (n_blocks + 1) * 4));
    CUDA_TRY(cudaMalloc(&d_ctr.p, 64));
    CUDA_TRY(cudaMemsetAsync(d_ctr.p, 0, 64, st));
    CUDA_TRY(cudaMemsetAsync(d_docs_len.p, 0, (n_blocks + 1) * 4, st));
    const uint64_t n_groups = (n_blocks + kIndexGroup - 1) / kIndexGroup;
    CUDA_TRY(cudaEventRecord(ev[0], st));
    for (int pass = 0; pass < 2; ++pass) {
        const dint_dict* d = pass ? freqs_dict : docs_dict;
        dint_device_plan* plan = plans[pass].get();
        plan->device = d->device;
        plan->in_bytes = lists_bytes;
        plan->n_chains = n_groups;
        CUDA_TRY(cudaMalloc(&plan->d_units, (n_blocks + 1) * sizeof(dint_strip_unit)));
        index_walk_params P;
        P.dv = d->view();
        P.sums = d->d_sums;
        P.lists = d_lists;
        P.lists_bytes = lists_bytes;
        P.blocks = d_blocks;
        P.n_blocks = n_blocks;
        P.block_size = block_size;
        P.multi = d->kind == DINT_DICT_MULTI;
        P.docs_len = d_docs_len.as<uint32_t>();
        P.units = plan->d_units;
        P.units_cap = n_blocks;
        P.counter = d_ctr.as<unsigned long long>() + pass;
        P.status = d_ctr.as<uint32_t>() + 8 + pass;
        P.unit_ints = lim.unit_ints;
        P.unit_bytes = lim.unit_bytes;
        P.strip_cw = lim.strip_codewords;
        P.strip_shift = strip_shift_of(lim);
        if (n_groups) {
            int grid = 0;
            size_t smem = 0;
            int rc;
            if (pass == 0) {
                rc = walk_grid(index_walk_kernel<0>, d, kWalkWarps, n_groups, &grid, &smem);
                if (rc) return rc;
                index_walk_kernel<0><<<grid, kWalkWarps * 32, smem, st>>>(P);
            } else {
                rc = walk_grid(index_walk_kernel<1>, d, kWalkWarps, n_groups, &grid, &smem);
                if (rc) return rc;
                index_walk_kernel<1><<<grid, kWalkWarps * 32, smem, st>>>(P);
            }
            ++g_launches;
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaEventRecord(ev[1 + pass], st));
    }
    uint64_t h[8] = {0};
    CUDA_TRY(cudaMemcpyAsync(h, d_ctr.p, 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(h) + 8;
    for (int pass = 0; pass < 2; ++pass) {
        if (hs[pass] != WALK_OK) {
            dint::set_error(std::string("index .bin: malformed ") + (pass ? "freqs" : "docs") + " block (device walk)");
            return DINT_E_FORMAT;
        }
        dint_device_plan* plan = plans[pass].get();
        plan->n_units = h[pass];
        CUDA_TRY(cudaEventElapsedTime(&plan->discover_ms, ev[pass], ev[1 + pass]));
        // the table was sized for one unit per block: give the slack back
        if (plan->n_units + 1 < (n_blocks + 1) / 4 * 3) {
            dint_strip_unit* tight = nullptr;
            CUDA_TRY(cudaMalloc(&tight, (plan->n_units + 1) * sizeof(dint_strip_unit)));
            cudaError_t e = cudaMemcpyAsync(tight, plan->d_units, plan->n_units * sizeof(dint_strip_unit), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                cudaFree(tight);
                dint::set_error(std::string("index discovery: ") + cudaGetErrorString(e));
                return DINT_E_CUDA;
            }
            cudaFree(plan->d_units);
            plan->d_units = tight;
        }
    }
    *docs_plan = plans[0].release();
    *freqs_plan = plans[1].release();
    return DINT_OK;
}


// ---- block table of a frozen index, built on the device -------------------------------------------------------
void dint_device_index_destroy(dint_device_index* p) {
    if (!p) return;
    {
        device_guard g(p->device);
        cudaFree(p->d_endpoints);
        cudaFree(p->d_list_offsets);
        cudaFree(p->d_blocks);
    }
    delete p;
}
uint64_t dint_device_index_num_lists(const dint_device_index* p) { return p ? p->h.n_lists : 0; }
uint64_t dint_device_index_num_docs(const dint_device_index* p) { return p ? p->h.num_docs : 0; }
uint64_t dint_device_index_num_postings(const dint_device_index* p) { return p ? p->n_postings : 0; }
uint64_t dint_device_index_num_blocks(const dint_device_index* p) { return p ? p->n_blocks : 0; }
double dint_device_index_build_ms(const dint_device_index* p) { return p ? double(p->build_ms) : 0.0; }
const dint_block* dint_device_index_blocks(const dint_device_index* p) { return p ? p->d_blocks : nullptr; }
const uint64_t* dint_device_index_list_offsets(const dint_device_index* p) { return p ? p->d_list_offsets : nullptr; }
const uint64_t* dint_device_index_endpoints(const dint_device_index* p) { return p ? p->d_endpoints : nullptr; }
int dint_device_index_layout(const dint_device_index* p, uint64_t* lists_off, uint64_t* lists_len, uint64_t* docs_dict_off,
                             uint64_t* freqs_dict_off, uint64_t* end_off) {
    if (!p) return DINT_E_ARG;
    if (lists_off) *lists_off = p->h.lists_off;
    if (lists_len) *lists_len = p->h.lists_len;
    if (docs_dict_off) *docs_dict_off = p->h.docs_dict_off;
    if (freqs_dict_off) *freqs_dict_off = p->h.freqs_dict_off;
    if (end_off) *end_off = p->h.end_off;
    return DINT_OK;
}

int dint_index_device_plan_create(int kind, const void* h_bin, size_t len, const uint8_t* d_bin, int device, uint32_t block_size,
                                  void* stream, dint_device_index** out) {
    if (!h_bin || !d_bin || !out) return DINT_E_ARG;
    if (!block_size) block_size = dint::kDefaultBlock;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        dint::set_error("no CUDA device: dint_b200 has no CPU fallback");
        return DINT_E_NODEVICE;
    }
    std::unique_ptr<dint_device_index, void (*)(dint_device_index*)> ix(new (std::nothrow) dint_device_index, dint_device_index_destroy);
    if (!ix) return DINT_E_NOMEM;
    ix->device = device;
    ix->kind = kind;
    int rc = dint::index_header_parse(kind, static_cast<const uint8_t*>(h_bin), len, &ix->h);  // O(1) host work
    if (rc) return rc;
    const dint_index_header& H = ix->h;
    device_guard g(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    device_info di;
    rc = get_device_info(device, &di);
    if (rc) return rc;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    struct ev_guard {
        cudaEvent_t a, b;
        ~ev_guard() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } evg{ev0, ev1};
    CUDA_TRY(cudaEventRecord(ev0, st));
    const uint64_t n = H.n_lists;
    CUDA_TRY(cudaMalloc(&ix->d_endpoints, (n + 1) * 8));
    CUDA_TRY(cudaMalloc(&ix->d_list_offsets, (n + 1) * 8));
    dev_buf d_status;
    CUDA_TRY(cudaMalloc(&d_status.p, 16));
    CUDA_TRY(cudaMemsetAsync(d_status.p, 0, 16, st));
    const uint64_t lists_len = H.lists_len;
    CUDA_TRY(cudaMemcpyAsync(ix->d_endpoints + n, &lists_len, 8, cudaMemcpyHostToDevice, st));
    if (n) {
        // ---- Elias-Fano endpoints --------------------------------------------------------------------------------
        ef_params E;
        E.words = d_bin + H.ef_words_off;
        E.n = n;
        E.hb_off = H.hb_off;
        E.hb_len = H.hb_len;
        E.lb_off = H.lb_off;
        E.l = uint32_t(H.l);
        E.n_hwords = (H.hb_len + 63) / 64;
        dev_buf d_popc, d_rank;
        CUDA_TRY(cudaMalloc(&d_popc.p, E.n_hwords * 4));
        CUDA_TRY(cudaMalloc(&d_rank.p, (E.n_hwords + 1) * 8));
        E.popc = d_popc.as<uint32_t>();
        E.rank = d_rank.as<uint64_t>();
        E.endpoints = ix->d_endpoints;
        const int grid = int(std::max<uint64_t>(1, std::min<uint64_t>((E.n_hwords + 255) / 256, uint64_t(di.sms) * 8)));
        ef_popc_kernel<<<grid, 256, 0, st>>>(E);
        scan_u32_kernel<<<1, 1024, 0, st>>>(E.popc, d_rank.as<uint64_t>(), E.n_hwords);
        ef_emit_kernel<<<grid, 256, 0, st>>>(E);
        g_launches += 3;
        CUDA_TRY(cudaGetLastError());
        uint64_t ones = 0;
        CUDA_TRY(cudaMemcpyAsync(&ones, d_rank.as<uint64_t>() + E.n_hwords, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (ones < n) {
            dint::set_error("index .bin: Elias-Fano: fewer endpoints than lists");
            return DINT_E_FORMAT;
        }
        // ---- list headers -> postings and blocks per list -> their prefix sums ---------------------------------------
        dev_buf d_n, d_nb, d_bbase;
        CUDA_TRY(cudaMalloc(&d_n.p, n * 4));
        CUDA_TRY(cudaMalloc(&d_nb.p, n * 4));
        CUDA_TRY(cudaMalloc(&d_bbase.p, (n + 1) * 8));
        index_lists_params L;
        L.lists = d_bin + H.lists_off;
        L.lists_len = H.lists_len;
        L.n = n;
        L.block_size = block_size;
        L.endpoints = ix->d_endpoints;
        L.list_n = d_n.as<uint32_t>();
        L.list_blocks = d_nb.as<uint32_t>();
        L.status = d_status.as<uint32_t>();
        const int lgrid = int(std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, uint64_t(di.sms) * 8)));
        index_lists_kernel<<<lgrid, 256, 0, st>>>(L);
        scan_u32_kernel<<<1, 1024, 0, st>>>(L.list_n, ix->d_list_offsets, n);
        scan_u32_kernel<<<1, 1024, 0, st>>>(L.list_blocks, d_bbase.as<uint64_t>(), n);
        g_launches += 3;
        CUDA_TRY(cudaGetLastError());
        uint64_t totals[2] = {0, 0};
        uint32_t h_status = 0;
        CUDA_TRY(cudaMemcpyAsync(&totals[0], ix->d_list_offsets + n, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&totals[1], d_bbase.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&h_status, d_status.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_status) {
            dint::set_error("index .bin: bad list header or endpoint (device walk)");
            return DINT_E_FORMAT;
        }
        ix->n_postings = totals[0];
        ix->n_blocks = totals[1];
        // ---- the block table ----------------------------------------------------------------------------------------
        CUDA_TRY(cudaMalloc(&ix->d_blocks, (ix->n_blocks + 1) * sizeof(dint_block)));
        index_blocks_params B;
        B.lists = L.lists;
        B.lists_len = H.lists_len;
        B.n = n;
        B.block_size = block_size;
        B.endpoints = ix->d_endpoints;
        B.list_n = L.list_n;
        B.list_offsets = ix->d_list_offsets;
        B.block_base = d_bbase.as<uint64_t>();
        B.blocks = ix->d_blocks;
        B.status = d_status.as<uint32_t>();
        const uint64_t groups = (n + 31) / 32;
        const int bgrid = int(std::max<uint64_t>(1, std::min<uint64_t>((groups + 7) / 8, uint64_t(di.sms) * 8)));
        index_blocks_kernel<<<bgrid, 256, 0, st>>>(B);
        ++g_launches;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&h_status, d_status.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_status) {
            dint::set_error("index .bin: block endpoint out of range (device walk)");
            return DINT_E_FORMAT;
        }
    } else {
        const uint64_t zero = 0;
        CUDA_TRY(cudaMemcpyAsync(ix->d_list_offsets, &zero, 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMalloc(&ix->d_blocks, sizeof(dint_block)));
    }
    CUDA_TRY(cudaEventRecord(ev1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventElapsedTime(&ix->build_ms, ev0, ev1));
    *out = ix.release();
    return DINT_OK;
}

}  // extern "C"
