#!/usr/bin/env python3
"""probe: device-side discovery time vs file size (K sample replicas per file)"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from dint_b200 import capi

w = bench.load_workload("gov2_r256", "multi_packed_dint")
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
out = []
for part in ("docs", "freqs"):
    bin_, dict_bytes, _, _ = w["parts"][part]
    d = capi.Dictionary(2, dict_bytes, device=0)
    base = torch.frombuffer(bytearray(bin_), dtype=torch.uint8).to(dev)
    for K in (1, 8, 32, 256):
        d_in = base.repeat(K)
        torch.cuda.synchronize()
        for rep in range(2):
            t0 = time.perf_counter()
            dp = capi.DevicePlan.discover_vroom(d, d_in.data_ptr(), d_in.numel(), 256, None, 0, stream=st)
            wall = (time.perf_counter() - t0) * 1e3
        out.append({"part": part, "K": K, "MB": d_in.numel() / 1e6, "wall_ms": wall, "device_ms": dp.discover_ms, "phases": dp.phase_ms,
                    "chains": dp.n_chains, "skimmed": dp.n_skimmed, "retries": dp.retries, "G_ints_s": dp.n_ints / dp.discover_ms / 1e6})
        print(json.dumps(out[-1]), flush=True)
        del dp, d_in
