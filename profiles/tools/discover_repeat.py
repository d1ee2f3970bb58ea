#!/usr/bin/env python3
"""probe: repeat the device-side discovery of the replicated freqs / docs file, print phases (variance hunt)"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from dint_b200 import capi
w = bench.load_workload("gov2_r256", "multi_packed_dint")
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for part in ("freqs", "docs"):
    bin_, dict_bytes, _, _ = w["parts"][part]
    d = capi.Dictionary(2, dict_bytes, device=0)
    d_in = torch.frombuffer(bytearray(bin_), dtype=torch.uint8).to(dev).repeat(K)
    big = torch.empty(int(40e9), dtype=torch.uint8, device=dev) if len(sys.argv) > 2 else None  # memory pressure like the bench
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dp = capi.DevicePlan.discover_vroom(d, d_in.data_ptr(), d_in.numel(), 256, None, 0, stream=st)
        wall = (time.perf_counter() - t0) * 1e3
        print(part, rep, round(wall, 1), [round(x, 1) for x in dp.phase_ms], dp.retries, flush=True)
        del dp
    del d_in, big
