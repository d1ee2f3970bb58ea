#!/usr/bin/env python3
"""Shape of the strip-unit table of the bench workload: strips per unit, integers and bytes per unit/strip."""
import os, sys, lzma
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dint_b200 import capi
KINDS = {"single_packed_dint": 1, "multi_packed_dint": 2}
DN = {"single_packed_dint": "single_packed", "multi_packed_dint": "multi_packed"}
def rd(p):
    return open(p, "rb").read() if os.path.exists(p) else lzma.open(p + ".xz").read()
d = os.path.join(ROOT, "bench_data", "gov2_r256")
for type_ in sys.argv[1:] or ["multi_packed_dint", "single_packed_dint"]:
    for part in ("docs", "freqs"):
        dic = capi.Dictionary(KINDS[type_], rd(f"{d}/dict.gov2_r256.{part}.{DN[type_]}.DSF-65536-16"))
        b = rd(f"{d}/vroom.{type_}.{part}.bin")
        plan = capi.VroomPlan(dic, b)
        su = plan.strip_units
        raw = su.view(np.uint8).reshape(len(su), 160)
        in_len = raw[:, 24:26].copy().view(np.uint16).ravel().astype(np.int64)
        n_ints = raw[:, 26:28].copy().view(np.uint16).ravel().astype(np.int64)
        n_strips = raw[:, 28].astype(np.int64)
        words = raw[:, 32:].copy().view(np.uint32).reshape(len(su), 32)
        out_rel = ((words >> 13) & 0xFFF).astype(np.int64)
        in_rel = (words & 0x1FFF).astype(np.int64)
        sel = (words >> 25) & 15
        # ints per strip
        nxt = np.concatenate([out_rel[:, 1:], n_ints[:, None]], axis=1)
        k = np.arange(32)[None, :]
        last = (k + 1 >= n_strips[:, None])
        nxt = np.where(last, n_ints[:, None], nxt)
        sint = np.where(k < n_strips[:, None], nxt - out_rel, 0)
        nin = np.concatenate([in_rel[:, 1:], in_len[:, None]], axis=1)
        nin = np.where(last, in_len[:, None], nin)
        sbytes = np.where(k < n_strips[:, None], nin - in_rel, 0)   # includes list headers between strips
        print(f"{type_} {part}: {len(b)} B, {int(plan.n_ints)} ints, {int(plan.n_lists)} lists, units {len(su)}, "
              f"strips/unit {n_strips.mean():.1f}, ints/unit {n_ints.mean():.0f}, bytes/unit {in_len.mean():.0f}, "
              f"ints/strip {sint.sum()/n_strips.sum():.1f}, max ints/strip per unit (mean) {sint.max(1).mean():.0f}, "
              f"bytes/strip {sbytes.sum()/n_strips.sum():.1f} max-bytes/strip per unit mean {sbytes.max(1).mean():.0f}")
        print("   units with n_strips<32: %.1f%%; narrow strips %.1f%%" % (100*(n_strips<32).mean(), 100*((sel>=6)&(k<n_strips[:,None])).sum()/n_strips.sum()))
        cs = plan.codeword_stats
        ncw = cs["u8_codewords"] + cs["u16_codewords"]
        print("   codewords %d (%.2f ints/codeword): " % (ncw, int(plan.n_ints) / ncw) +
              ", ".join(f"{k} {100 * v / ncw:.1f}%" for k, v in cs.items() if k in ("exc16", "exc32", "runs", "len1", "len2", "len4", "len8", "len16", "u8_codewords")) +
              f"; run ints {100 * cs['run_ints'] / int(plan.n_ints):.1f}% of ints; loop trips {cs['unit_trips']} for {ncw} codewords -> lane utilisation {100 * ncw / (32 * cs['unit_trips']):.1f}%")
