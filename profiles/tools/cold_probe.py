#!/usr/bin/env python3
"""probe: where a cold first-touch call spends its wall time (create_discover / decode+results / destroy)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from dint_b200 import capi
w = bench.load_workload("gov2_r256", "multi_packed_dint")
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for part, mode in (("docs", 1), ("freqs", 2)):
    bin_, dict_bytes, _, _ = w["parts"][part]
    d = capi.Dictionary(2, dict_bytes, device=0)
    h_bin = torch.frombuffer(bytearray(bin_ * K), dtype=torch.uint8).pin_memory()
    h_sums = torch.empty(K * w["n_lists"], dtype=torch.int64).pin_memory()
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec = capi.VroomHostDecoder.discover(d, h_bin.data_ptr(), h_bin.numel(), block_size=256, stream=st)
        t1 = time.perf_counter()
        dec.decode_resident(mode=mode, stream=st)
        dec.results(h_sums_ptr=h_sums.data_ptr(), stream=st)
        dec.sync(st)
        t2 = time.perf_counter()
        ph = dec.device_plan.phase_ms
        dms = dec.device_plan.discover_ms
        dec.close()
        t3 = time.perf_counter()
        print(part, rep, "create", round((t1 - t0) * 1e3, 1), "(device discover", round(dms, 1), [round(x, 1) for x in ph], ") decode+results",
              round((t2 - t1) * 1e3, 1), "destroy", round((t3 - t2) * 1e3, 1), flush=True)
