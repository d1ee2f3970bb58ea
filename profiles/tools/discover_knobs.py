#!/usr/bin/env python3
"""device-side discovery of K replicas of the gov2 sample under different chain-cutting knobs
(DINT_B200_LONG_CHAIN / DINT_B200_CUT_BYTES, bytes; CUT_BYTES=0: no skim pass): phase times + table sizes"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from dint_b200 import capi
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
configs = [tuple(c.split(":")) for c in sys.argv[2:]] or [("", "")]
w = bench.load_workload("gov2_r256", "multi_packed_dint")
st = torch.cuda.current_stream().cuda_stream
out = []
for part in ("docs", "freqs"):
    bin_, dict_bytes, _, _ = w["parts"][part]
    d = capi.Dictionary(2, dict_bytes, device=0)
    d_in = torch.frombuffer(bytearray(bin_), dtype=torch.uint8).cuda().repeat(K)
    for rep in range(2):
        for lc, cb in configs:
            for k, v in (("DINT_B200_LONG_CHAIN", lc), ("DINT_B200_CUT_BYTES", cb)):
                if v:
                    os.environ[k] = v
                else:
                    os.environ.pop(k, None)
            dp = capi.DevicePlan.discover_vroom(d, d_in.data_ptr(), d_in.numel(), 256, None, 0, stream=st)
            row = {"part": part, "long": lc, "cut": cb, "rep": rep, "ms": round(dp.discover_ms, 1), "phases": [round(x, 1) for x in dp.phase_ms],
                   "chains": dp.n_chains, "skimmed": dp.n_skimmed, "retries": dp.retries, "lists": dp.n_lists, "ints": dp.n_ints, "units": dp.n_units}
            print(json.dumps(row), flush=True)
            out.append(row)
            del dp
    del d_in
