#!/usr/bin/env python3
"""bench.py — decoded integers per second of the DINT decode path on B200.

One "step" = one pass of the hot path over the whole synthetic collection shard of this
rank: the docs stream decoded with the fused d-gap prefix sum (docIDs) and the freqs stream
with the fused +1 shift, i.e. 2 x postings integers, through the C ABI of libdint_b200.so.

Workload (config.workload): the Gov2-shaped synthetic collection of SURVEY.md §8d config 4
(25 M docs, ~6 G postings, Zipfian list lengths, geometric+bursty gaps). The reference's own
CPU encoder produced a 1/R stratified sample of it (oracle/make_synthetic.py, checked by the
reference's check_encoded_data); the sample is replicated R times into distinct HBM regions
so that inputs (~9 GB) and outputs (~48 GB) dwarf the 126 MB L2. If bench_data/ is absent the
committed test_collection fixtures are replicated instead and config.workload says so.

Multi-GPU: one process per GPU (torchrun), every rank decodes its own shard of the same
shape (term-range sharding, dictionary replicated, no traffic on the hot path -> "weak"
scaling); NCCL only for the final {count, checksum} all-reduce and the max-over-ranks time.

`--impl reference` times the reference's own CPU decoder (oracle/_ref/libdint_ref.so, built
from /root/reference; else the C port in oracle/) on all host cores, same metric and config.
"""
import argparse
import ctypes as C
import hashlib
import json
import lzma
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DICT_NAMES = {"single_rect_dint": "rectangular", "single_packed_dint": "single_packed",
              "multi_packed_dint": "multi_packed"}
KINDS = {"single_rect_dint": 0, "single_packed_dint": 1, "multi_packed_dint": 2}
FULL_POSTINGS = {"gov2": 6_000_000_000, "clueweb09": 15_000_000_000}


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def _read(path):
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    with lzma.open(path + ".xz") as f:
        return f.read()


def _materialize(path):
    """make sure an uncompressed copy exists on disk (the reference loader wants a path)"""
    if not os.path.exists(path):
        with lzma.open(path + ".xz") as fi, open(path + f".tmp{os.getpid()}", "wb") as fo:
            fo.write(fi.read())
        os.replace(path + f".tmp{os.getpid()}", path)
    return path


def load_workload(name, type_, postings_target):
    """returns dict(parts={docs,freqs -> (bin bytes, dict bytes, dict path)}, expected digests, replicas ...)"""
    d = os.path.join(ROOT, "bench_data", name)
    w = {"type": type_, "block_size": 256}
    if os.path.exists(os.path.join(d, "meta.json")) and \
            (os.path.exists(os.path.join(d, f"vroom.{type_}.docs.bin")) or
             os.path.exists(os.path.join(d, f"vroom.{type_}.docs.bin.xz"))):
        meta = json.load(open(os.path.join(d, "meta.json")))
        w["name"] = name
        w["desc"] = (f"synthetic {meta['shape']}-shaped collection ({meta['num_docs'] / 1e6:.0f}M docs, "
                     f"{meta['full_scale_postings'] / 1e9:.0f}G postings law): reference-encoded 1/{meta['replicas']} "
                     f"stratified sample x {meta['replicas']} replicas in HBM")
        w["replicas"] = meta["replicas"]
        w["n_postings"] = meta["n_postings"]
        w["sha"] = {"docs": meta["docids_sha256"], "freqs": meta["freqs_sha256"]}
        w["parts"] = {}
        for part in ("docs", "freqs"):
            dict_file = os.path.join(d, f"dict.{name}.{part}.{DICT_NAMES[type_]}.DSF-65536-16")
            w["parts"][part] = (_read(os.path.join(d, f"vroom.{type_}.{part}.bin")), _read(dict_file),
                                _materialize(dict_file))
    else:
        # fallback: the committed test_collection fixtures (tests/golden), replicated to the same size
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import golden_data
        G = golden_data.Golden()
        w["name"] = "test_collection"
        w["n_postings"] = G.n_postings
        w["replicas"] = max(1, round(postings_target / G.n_postings))
        w["desc"] = (f"FALLBACK (bench_data/{name} absent): reference-encoded test_collection "
                     f"(3.3M postings) x {w['replicas']} replicas in HBM")
        w["sha"] = {"docs": G.sha["docids_sha256"], "freqs": G.sha["freqs_sha256"]}
        w["parts"] = {}
        for part in ("docs", "freqs"):
            w["parts"][part] = (G.bytes(G.vroom_name(type_, part)), G.bytes(G.dict_name(type_, part)),
                                G.path(G.dict_name(type_, part)))
    return w


# ------------------------------------------------------------------------------------------
# reference CPU decoder (the checker, timed as a baseline only)
# ------------------------------------------------------------------------------------------
class CpuDecoder:
    def __init__(self):
        ref = os.path.join(ROOT, "oracle", "_ref", "libdint_ref.so")
        self.kind = "reference" if os.path.exists(ref) else "port"
        if self.kind == "reference":
            L = self.lib = C.CDLL(ref)
            L.ref_dict_load.restype = C.c_void_p
            L.ref_dict_load.argtypes = [C.c_int, C.c_char_p]
            L.ref_vroom_time_file.restype = C.c_double
            L.ref_vroom_time_file.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
        else:
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, stdout=subprocess.DEVNULL)
            L = self.lib = C.CDLL(os.path.join(ROOT, "oracle", "libdint_oracle.so"))
            L.oracle_dict_parse.restype = C.c_void_p
            L.oracle_dict_parse.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
            L.oracle_vroom_decode_file.restype = C.c_long
            L.oracle_vroom_decode_file.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p,
                                                   C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]

    def prepare(self, w):
        self.streams = []
        for part in ("docs", "freqs"):
            bin_, dict_bytes, dict_path = w["parts"][part]
            a = np.frombuffer(bin_, dtype=np.uint8)
            if self.kind == "reference":
                h = self.lib.ref_dict_load(KINDS[w["type"]], dict_path.encode())
            else:
                db = np.frombuffer(dict_bytes, dtype=np.uint8)
                h = self.lib.oracle_dict_parse(KINDS[w["type"]], db.ctypes.data, len(db))
            assert h, "cannot load dictionary for the CPU baseline"
            self.streams.append((h, a))
        self.ints_per_pass = 2 * w["n_postings"]
        self.scratch = {}

    def _scratch(self):
        # one lazily-mapped 50 M-int output buffer per thread, as the reference allocates once
        tid = threading.get_ident()
        if tid not in self.scratch:
            self.scratch[tid] = np.zeros(50_000_000 + 1024, dtype=np.uint32)
        return self.scratch[tid]

    def one_pass(self, repeat=1):
        """decode docs + freqs of the sample `repeat` times on the calling thread.
        returns (wall seconds, seconds as summed by the reference's per-list chrono)"""
        t0 = time.perf_counter()
        chrono = 0.0
        for h, a in self.streams:
            if self.kind == "reference":
                tot = C.c_uint64(0)
                chrono += self.lib.ref_vroom_time_file(h, a.ctypes.data, len(a), repeat, C.byref(tot),
                                                       self._scratch().ctypes.data)
            else:
                out = np.zeros(self.ints_per_pass // 2 + 16, dtype=np.uint32)
                tot = C.c_uint64(0)
                for _ in range(repeat):
                    self.lib.oracle_vroom_decode_file(h, a.ctypes.data, len(a), 256, out.ctypes.data, len(out) - 16,
                                                      None, 0, C.byref(tot))
        wall = time.perf_counter() - t0
        return wall, (chrono if self.kind == "reference" else wall)

    def run_all_cores(self, threads, repeat):
        from concurrent.futures import ThreadPoolExecutor
        if getattr(self, "pool_size", 0) != threads:
            self.pool = ThreadPoolExecutor(max_workers=threads)
            self.pool_size = threads
        t0 = time.perf_counter()
        futs = [self.pool.submit(self.one_pass, repeat) for _ in range(threads)]
        for f in futs:
            f.result()
        wall = time.perf_counter() - t0
        return threads * repeat * self.ints_per_pass / wall, wall


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(w, budget_s=12.0):
    dec = CpuDecoder()
    dec.prepare(w)
    wall1, chrono1 = dec.one_pass(1)  # single thread, as shipped (sum of per-list chrono deltas)
    single = dec.ints_per_pass / chrono1
    threads = host_threads()
    repeat = max(1, int(budget_s / max(wall1 * 1.5, 1e-3)))
    repeat = min(repeat, 50)
    allv, wall = dec.run_all_cores(threads, repeat)
    return {"value": allv / 1e9, "unit": "G ints/s", "cores": threads, "kind": dec.kind,
            "single_thread_value": single / 1e9,
            "sample": (f"{w['name']} sample (docs+freqs, {dec.ints_per_pass} ints), {repeat} passes on each of "
                       f"{threads} threads (all cores, wall clock); single thread = 1 pass as shipped "
                       f"(sum of per-list chrono deltas, vroom_env/decode.cpp:142-147)")}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 50 ms from before the warm-up on; only the samples whose timestamp falls inside
    the timed region [mark_begin, mark_end] are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        import datetime
        sm, mx, reasons, sm_all, mx_all = [], [], set(), [], []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            sm_all.append(clk)
            mx_all.append(cmax)
            if self.t0 is not None and self.t1 is not None and not (self.t0 <= ts <= self.t1 + 0.02):
                continue
            sm.append(clk)
            mx.append(cmax)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # clock skew between nvidia-smi's local time and ours: fall back to the last samples
            sm, mx = sm_all[-3:], mx_all[-3:]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ReplicatedStream:
    """R back-to-back copies of one vroom .bin in HBM + the matching segment table."""

    def __init__(self, torch, capi, dev, type_, bin_bytes, dict_bytes, replicas, block_size, engine="strips"):
        self.capi = capi
        self.engine = engine
        self.dict = capi.Dictionary(KINDS[type_], dict_bytes, device=dev.index)
        t0 = time.perf_counter()
        self.plan = capi.VroomPlan(self.dict, bin_bytes, block_size=block_size)
        self.plan_s = time.perf_counter() - t0
        self.R = replicas
        self.in_bytes = len(bin_bytes)
        self.n_ints = int(self.plan.n_ints)
        self.n_lists = int(self.plan.n_lists)
        self.n_segs = int(self.plan.n_segments)
        base = torch.frombuffer(bytearray(bin_bytes), dtype=torch.uint8).to(dev)
        self.d_in = base.repeat(replicas)
        segs = torch.from_numpy(self.plan.segments.view(np.int64).reshape(-1, 4).copy()).to(dev)
        shift = torch.zeros(replicas, 1, 4, dtype=torch.int64, device=dev)
        r = torch.arange(replicas, dtype=torch.int64, device=dev)
        shift[:, 0, 0] = r * self.in_bytes
        shift[:, 0, 1] = r * self.n_ints
        self.d_segs = (segs[None, :, :] + shift).reshape(-1, 4).contiguous()
        self.n_units = int(self.plan.n_units)
        uf = torch.from_numpy(self.plan.unit_first.astype(np.int64)).to(dev)
        self.d_unit_first = torch.cat([(uf[:-1][None, :] + (r * self.n_segs)[:, None]).reshape(-1),
                                       torch.tensor([replicas * self.n_segs], dtype=torch.int64, device=dev)]
                                      ).to(torch.int32)
        # strip units (lane-sequential decoder): 160-byte records, in_off / out_off are the first two u64
        self.n_strip_units = int(self.plan.n_strip_units)
        su = torch.from_numpy(self.plan.strip_units.view(np.int64).reshape(-1, 20).copy()).to(dev)
        sshift = torch.zeros(replicas, 1, 20, dtype=torch.int64, device=dev)
        sshift[:, 0, 0] = r * self.in_bytes
        sshift[:, 0, 1] = r * self.n_ints
        self.d_strip_units = (su[None, :, :] + sshift).reshape(-1, 20).contiguous()
        lo = torch.from_numpy(self.plan.list_offsets.astype(np.int64)).to(dev)
        self.d_list_offsets = torch.cat([(lo[:-1][None, :] + (r * self.n_ints)[:, None]).reshape(-1),
                                         torch.tensor([replicas * self.n_ints], dtype=torch.int64, device=dev)])
        self.out = torch.empty(replicas * self.n_ints, dtype=torch.int32, device=dev)
        self.algo_bytes = replicas * (self.in_bytes + 4 * self.n_ints)

    def launch(self, mode, stream):
        if self.engine == "strips":
            self.capi.decode_strips(self.dict, self.d_in.data_ptr(), self.R * self.in_bytes, self.d_strip_units.data_ptr(),
                                    self.R * self.n_strip_units, self.out.data_ptr(), mode, stream=stream)
            return
        self.capi.decode_units(self.dict, self.d_in.data_ptr(), self.R * self.in_bytes, self.d_segs.data_ptr(),
                               self.R * self.n_segs, self.d_unit_first.data_ptr(), self.R * self.n_units,
                               self.out.data_ptr(), mode, stream=stream)


def bind_to_gpu_numa_node(torch, local):
    """Multi-rank runs: keep this rank's host threads (and therefore its pinned staging buffers, first touch)
    on the CPUs next to its GPU, so that the end-to-end copies do not cross sockets. Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from dint_b200 import capi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        bind_to_gpu_numa_node(torch, local)
        dist.init_process_group("nccl", device_id=dev)

    w = load_workload(args.workload, args.type, FULL_POSTINGS.get(args.workload.split("_")[0], 6_000_000_000))
    R = args.replicas or w["replicas"]
    if R != w["replicas"]:
        w["desc"] += f" [--replicas {R} override]"
    streams = {}
    for part in ("docs", "freqs"):
        bin_, dict_bytes, _ = w["parts"][part]
        streams[part] = ReplicatedStream(torch, capi, dev, args.type, bin_, dict_bytes, R, w["block_size"], args.engine)
    sd, sf = streams["docs"], streams["freqs"]
    ints_per_step = R * (sd.n_ints + sf.n_ints)
    st = torch.cuda.current_stream(dev)

    def step(events=None):
        if events is not None:
            events[0].record(st)
        sd.launch(capi.MODE_DOCIDS, st.cuda_stream)
        if events is not None:
            events[1].record(st)
        sf.launch(capi.MODE_FREQS, st.cuda_stream)
        if events is not None:
            events[2].record(st)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- warm-up + parity (untimed) ---------------------------------------------------------
    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize(dev)
    parity = {}
    for part, s in streams.items():
        first = s.out[:s.n_ints].cpu().numpy().view(np.uint32)
        last = s.out[(R - 1) * s.n_ints:].cpu().numpy().view(np.uint32)
        ok = hashlib.sha256(first.tobytes()).hexdigest() == w["sha"][part] and np.array_equal(first, last)
        sums = torch.empty(R * s.n_lists, dtype=torch.int64, device=dev)
        capi.checksum_lists(s.out.data_ptr(), s.d_list_offsets.data_ptr(), R * s.n_lists, sums.data_ptr(),
                            stream=st.cuda_stream)
        sums = sums.view(R, s.n_lists)
        ok = ok and bool((sums == sums[0:1]).all().item())
        parity[part] = ok
        s.checksum = int(sums.sum().item())
        del sums
    if not all(parity.values()):
        raise SystemExit(f"bench.py: PARITY FAILURE on rank {rank}: {parity} — refusing to report a number")

    # ---- timed region -----------------------------------------------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = capi.kernel_launches()
    sampler.mark_begin()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(st)
    for i in range(K):
        step(ev[i])
    t_end.record(st)
    torch.cuda.synchronize(dev)
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = capi.kernel_launches() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    docs_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    freqs_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    tot = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    red = torch.tensor([ints_per_step, (sd.checksum + sf.checksum) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dist.all_reduce(red, op=dist.ReduceOp.SUM)  # the only collective of the job: {count, checksum}
    elapsed_ms = float(tot.item())
    total_ints_per_step = int(red[0].item())

    # ---- end-to-end through the host-buffer C-ABI call ----------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, capi, dev, w, args, R, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = sd.algo_bytes / (docs_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        key = f"{w['name']}:{args.type}:docs"
        if key in t:
            traffic = t[key]["dram_bytes_per_int"] * R * sd.n_ints
    value = total_ints_per_step * K / (elapsed_ms * 1e-3) / 1e9
    line = {
        "metric": "decoded ints/sec (billions)", "value": value, "unit": "G ints/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": w["desc"], "codec": args.type, "dictionary": "DSF-65536-16",
                   "ints_per_step_per_gpu": ints_per_step, "postings_per_gpu": R * sd.n_ints,
                   "streams": "docs (fused docID prefix sum) + freqs (fused +1)",
                   "compressed_bytes_per_gpu": R * (sd.in_bytes + sf.in_bytes),
                   "sharding": f"term-range, {world} rank(s), dictionary replicated, no hot-path collective",
                   "l2": "inputs+outputs per step >> 126 MB L2 (no flush needed)",
                   "parity": "sha256 of replica 0 == reference collection; all replicas' per-list checksums equal"},
        "roofline": {"bound": "hbm", "kernel": ("decode_strips_kernel<DOCIDS>" if args.engine == "strips" else
                                                  "decode_units_kernel<DOCIDS>") + " (docs stream)", "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_launch": sd.algo_bytes, "avg_launch_ms": docs_ms, "traffic": traffic,
                     "freqs_kernel": {"achieved": sf.algo_bytes / (freqs_ms * 1e-3) / 1e9, "avg_launch_ms": freqs_ms,
                                      "algorithmic_bytes_per_launch": sf.algo_bytes},
                     "bytes_per_int": {"docs": sd.algo_bytes / (R * sd.n_ints), "freqs": sf.algo_bytes / (R * sf.n_ints)}},
        "clocks": clocks, "gpu_launches": int(launches),
        "checksum": int(red[1].item()),
        # one-time host pre-pass per .bin (dint_vroom_plan_create: header walk + strip table), NOT in any timed region;
        # a vroom file has no list index, so one sequential walk is inherent (vroom_env/decode.cpp:139-150 is that walk)
        "prepass": {"what": "dint_vroom_plan_create, docs + freqs sample files, 1 host thread",
                    "ms": (sd.plan_s + sf.plan_s) * 1e3,
                    "G_ints_per_s": (sd.n_ints + sf.n_ints) / (sd.plan_s + sf.plan_s) / 1e9,
                    "strip_table_bytes_per_int": 160.0 * (sd.n_strip_units + sf.n_strip_units) / (sd.n_ints + sf.n_ints),
                    "codewords": {"docs": sd.plan.codeword_stats, "freqs": sf.plan.codeword_stats}},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(w)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(torch, capi, dev, w, args, R, world):
    """same metric through dint_vroom_decode_host: HOST buffers in (pinned .bin images), per-list
    checksums out; every step copies every replica's bytes over PCIe again. The call is stream
    ordered, so the caller overlaps the copies of one file with the decode of the previous one by
    alternating between two contexts on two streams (what a host driver with >1 file would do)."""
    import torch.distributed as dist
    st = torch.cuda.current_stream(dev)
    lanes = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    decs, ints = [], 0
    h2d = d2h = 0
    # one host "file" per call = K back-to-back copies of the sample's .bin (a real collection is one
    # multi-GB file; K keeps a call's copies large enough for PCIe to stream)
    K = max(1, min(args.e2e_batch, R))
    while R % K:
        K -= 1
    for part, mode in (("docs", capi.MODE_DOCIDS), ("freqs", capi.MODE_FREQS)):
        bin_, dict_bytes, _ = w["parts"][part]
        big = bin_ * K
        d = capi.Dictionary(KINDS[args.type], dict_bytes, device=dev.index)
        plan = capi.VroomPlan(d, big, block_size=w["block_size"])
        h_bin = torch.frombuffer(bytearray(big), dtype=torch.uint8).pin_memory()
        per_lane = []
        for _ in lanes:
            dec = capi.VroomHostDecoder(plan, block_size=w["block_size"])
            h_sums = torch.empty(int(plan.n_lists), dtype=torch.int64).pin_memory()
            per_lane.append((dec, h_sums))
        decs.append((per_lane, h_bin, mode, d, plan))
        ints += int(plan.n_ints) // K
        h2d += len(bin_)
        d2h += int(plan.n_lists) // K * 8
    steps = max(1, min(args.steps, args.e2e_steps))

    def one_step():
        for r in range(R // K):
            ln = r & 1
            for per_lane, h_bin, mode, _, _ in decs:
                dec, h_sums = per_lane[ln]
                dec.decode(h_bin.data_ptr(), h_bin.numel(), mode=mode, h_sums_ptr=h_sums.data_ptr(),
                           stream=lanes[ln].cuda_stream)

    def fork():
        for s in lanes:
            s.wait_stream(st)

    def join():
        for s in lanes:
            st.wait_stream(s)
    fork()
    one_step()
    join()
    torch.cuda.synchronize(dev)
    # the checksums that came back over PCIe must be the ones of the resident run
    for per_lane, _, _, _, plan in decs:
        assert bool((per_lane[0][1] == per_lane[1][1]).all()), "e2e: lanes disagree"
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(st)
    fork()
    for _ in range(steps):
        one_step()
    join()
    t1.record(st)
    torch.cuda.synchronize(dev)
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    val = world * R * ints * steps / (float(ms.item()) * 1e-3) / 1e9
    e2e_sum = sum(int(per_lane[0][1].sum().item()) & 0xFFFFFFFFFFFF for per_lane, _, _, _, _ in decs)
    return {"value": val, "unit": "G ints/s", "h2d_bytes_per_step": R * h2d, "d2h_bytes_per_step": R * d2h,
            "steps": steps, "call": "dint_vroom_decode_host (pinned .bin in, per-list checksums out), "
            f"{2 * (R // K)} calls per step ({K} sample files per call) alternating between 2 contexts/streams", "ms_per_step": float(ms.item()) / steps,
            "checksum48": e2e_sum}


# ------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    w = load_workload(args.workload, args.type, FULL_POSTINGS.get(args.workload.split("_")[0], 6_000_000_000))
    dec = CpuDecoder()
    dec.prepare(w)
    threads = host_threads()
    wall1, _ = dec.one_pass(1)
    # bounded sample: each step = every thread decodes the sample `repeat` times (~3 s per step)
    repeat = max(1, min(20, int(3.0 / max(wall1 * 1.3, 1e-3))))
    for _ in range(args.warmup):
        dec.run_all_cores(threads, 1)
    t0 = time.perf_counter()
    ints = 0
    for _ in range(args.steps):
        v, wall = dec.run_all_cores(threads, repeat)
        ints += threads * repeat * dec.ints_per_pass
    elapsed = time.perf_counter() - t0
    value = ints / elapsed / 1e9
    sample = (f"{w['name']} sample (docs+freqs, {dec.ints_per_pass} ints) decoded {repeat}x per step by each of "
              f"{threads} host threads; output is raw d-gaps/freq-1 as the reference's decode writes them")
    line = {"impl": "reference", "metric": "decoded ints/sec (billions)", "value": value, "unit": "G ints/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["desc"], "codec": args.type, "dictionary": "DSF-65536-16"},
            "cpu_baseline": {"value": value, "unit": "G ints/s", "cores": threads, "kind": dec.kind, "sample": sample},
            "e2e": {"value": value, "unit": "G ints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dint_b200", choices=["dint_b200", "reference"])
    ap.add_argument("--workload", default="gov2_r256")
    ap.add_argument("--type", default=None, choices=sorted(KINDS))
    ap.add_argument("--engine", default="strips", choices=("strips", "units"),
                    help="strips = lane-sequential strip decoder (default); units = warp-per-unit tile decoder")
    ap.add_argument("--replicas", type=int, default=0, help="override the number of HBM replicas (debug)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-batch", type=int, default=8, help="sample .bin copies per host-buffer call")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.type is None:
        d = os.path.join(ROOT, "bench_data", args.workload)
        has_multi = any(os.path.exists(os.path.join(d, f"vroom.multi_packed_dint.docs.bin{s}")) for s in ("", ".xz"))
        args.type = "multi_packed_dint" if has_multi or not os.path.exists(os.path.join(d, "meta.json")) \
            else "single_packed_dint"
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
