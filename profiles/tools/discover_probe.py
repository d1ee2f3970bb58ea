#!/usr/bin/env python3
"""probe: device-side discovery over the replicated bench file (no host plan) + decode with the device tables"""
import hashlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from dint_b200 import capi, driver

def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    name = sys.argv[2] if len(sys.argv) > 2 else "gov2_r256"
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    w = bench.load_workload(name, "multi_packed_dint", 0)
    dev = torch.device("cuda", 0)
    out = {}
    for part, mode in (("docs", capi.MODE_DOCIDS), ("freqs", capi.MODE_FREQS)):
        bin_, dict_bytes, _ = w["parts"][part]
        d = capi.Dictionary(2, dict_bytes, device=0)
        base = torch.frombuffer(bytearray(bin_), dtype=torch.uint8).to(dev)
        d_in = base.repeat(R)
        torch.cuda.synchronize()
        st = torch.cuda.current_stream().cuda_stream
        res = []
        for rep in range(2):
            t0 = time.perf_counter()
            dp = capi.DevicePlan.discover_vroom(d, d_in.data_ptr(), d_in.numel(), 256, None, stride, stream=st)
            wall = time.perf_counter() - t0
            res.append({"wall_ms": wall * 1e3, "device_ms": dp.discover_ms, "phase_ms": dp.phase_ms, "chains": dp.n_chains, "sync_found": dp.n_sync_found,
                        "retries": dp.retries, "units": dp.n_units, "lists": dp.n_lists, "ints": dp.n_ints})
        n = dp.n_ints
        o = torch.empty(n, dtype=torch.int32, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(2):
            capi.decode_strips(d, d_in.data_ptr(), d_in.numel(), dp.units_ptr, dp.n_units, o.data_ptr(), mode, stream=st)
        ev[0].record()
        for _ in range(5):
            capi.decode_strips(d, d_in.data_ptr(), d_in.numel(), dp.units_ptr, dp.n_units, o.data_ptr(), mode, stream=st)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        per = n // R
        first = o[:per].cpu().numpy().view(np.uint32)
        last = o[(R - 1) * per:].cpu().numpy().view(np.uint32)
        ok = hashlib.sha256(first.tobytes()).hexdigest() == w["sha"][part] and np.array_equal(first, last)
        out[part] = {"discover": res, "decode_ms": ms, "G_ints_s": n / ms / 1e6, "parity": bool(ok),
                     "table_bytes_per_int": 160.0 * dp.n_units / n,
                     "discover_G_ints_s": n / res[-1]["device_ms"] / 1e6}
        del o, d_in, dp
    print(json.dumps(out, indent=1))

main()
