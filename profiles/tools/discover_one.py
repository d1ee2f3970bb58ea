#!/usr/bin/env python3
"""one device-side discovery of K replicas of the gov2 sample's docs file (for ncu)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from dint_b200 import capi
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
part = sys.argv[2] if len(sys.argv) > 2 else "docs"
w = bench.load_workload("gov2_r256", "multi_packed_dint")
bin_, dict_bytes, _, _ = w["parts"][part]
d = capi.Dictionary(2, dict_bytes, device=0)
d_in = torch.frombuffer(bytearray(bin_), dtype=torch.uint8).cuda().repeat(K)
st = torch.cuda.current_stream().cuda_stream
dp = capi.DevicePlan.discover_vroom(d, d_in.data_ptr(), d_in.numel(), 256, None, 0, stream=st)
print(dp.n_chains, dp.phase_ms)
