#!/usr/bin/env python3
"""Join an ncu SASS source page with nvdisasm line info: instructions executed and stall
samples per CUDA source line.

  cuobjdump -xelf all dint_b200/libdint_b200.so ; nvdisasm -g -c dint_kernels.sm_100a.cubin > dis.txt
  ncu -i prof.ncu-rep --page source --csv > sass.csv
  python profiles/tools/ncu_lines.py sass.csv dis.txt <kernel substring> [launch index]
"""
import csv
import re
import sys
from collections import defaultdict


def parse_dis(path, want):
    funcs, cur, line = {}, None, None
    for t in open(path, errors="replace"):
        m = re.match(r"^(_Z\w+):\s*$", t)
        if m:
            cur = m.group(1)
            funcs[cur] = {}
            line = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', t)
        if m:
            if "inlined at" not in t:
                line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", t)
        if m and cur:
            funcs[cur][int(m.group(1), 16)] = (line, m.group(2).strip())
    return funcs


def main():
    sass, dis, want = sys.argv[1], sys.argv[2], sys.argv[3]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(sass)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    norm = lambda t: re.sub(r"\((int|bool)\)", "", t)
    secs = [i for i in secs if want in norm(rows[i][1])]
    s = secs[launch]
    nxt = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > s]
    e = nxt[0] if nxt else len(rows)
    hdr = rows[s + 1]
    ia, ii, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = rows[s + 2:e]
    a0 = int(body[0][ia], 16)
    funcs = parse_dis(dis, want)
    name = re.sub(r"[^A-Za-z0-9]", "", rows[s][1].split("(")[0])
    cand = [f for f in funcs if all(tok in f for tok in re.findall(r"[a-z_]{4,}", want.split("<")[0]))]
    # pick the function with the same instruction count
    f = [c for c in cand if len(funcs[c]) == len(body)]
    if not f:
        print("no disassembled function with", len(body), "instructions among", [(c[-40:], len(funcs[c])) for c in cand])
        return
    fn = funcs[f[0]]
    per = defaultdict(lambda: [0, 0])
    tot = totsm = 0
    for r in body:
        off = int(r[ia], 16) - a0
        ln, _ = fn.get(off, (None, ""))
        n, smp = int(r[ii] or 0), int(r[ism] or 0)
        per[ln][0] += n
        per[ln][1] += smp
        tot += n
        totsm += smp
    print(f"{rows[s][1][:80]}\n total warp-instructions {tot}, samples {totsm}")
    src = {}
    for k in per:
        if k and k[0] not in src:
            for base in ("dint_b200/csrc/", "include/"):
                try:
                    src[k[0]] = open(base + k[0]).read().split("\n")
                except OSError:
                    pass
    for k, (n, smp) in sorted(per.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 60]:
        text = src.get(k[0], [""] * 100000)[k[1] - 1].strip()[:90] if k else ""
        print(f"{100 * n / tot:5.1f}% inst {100 * smp / max(totsm, 1):5.1f}% smp  {k[0] + ':' + str(k[1]) if k else '?':24} {text}")


if __name__ == "__main__":
    main()
