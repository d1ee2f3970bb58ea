#!/usr/bin/env python3
"""Shared-memory wavefronts per SASS instruction from `ncu --page source --csv` (which instruction pays for bank conflicts).
usage: ncu_wavefronts.py sass.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, s in enumerate(secs):
    e = secs[n + 1] if n + 1 < len(secs) else len(rows)
    hdr = rows[s + 1]
    body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
    iw, ii, isrc, ie = (hdr.index(k) for k in ('L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal', 'Source', 'Instructions Executed'))
    tot = sum(float(r[iw] or 0) for r in body)
    ideal = sum(float(r[ii] or 0) for r in body)
    inst = sum(int(r[ie] or 0) for r in body)
    print(f"== {rows[s][1][:70]} (launch {n}): {inst} warp instructions, {tot:.0f} shared wavefronts ({ideal:.0f} ideal)")
    L = [(float(r[iw] or 0), float(r[ii] or 0), int(r[ie] or 0), k, r[isrc]) for k, r in enumerate(body) if float(r[iw] or 0) > 0]
    for w, i, c, k, src in sorted(L, reverse=True)[:top]:
        print(f"  {100 * w / tot:5.1f}%  {w / max(c, 1):5.2f} wavefronts/inst (ideal {i / max(c, 1):4.2f})  x{c:<9d} #{k:<4d} {src[:64]}")
