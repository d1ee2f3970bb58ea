#!/usr/bin/env python3
"""TEST/BENCH INFRASTRUCTURE — synthetic Gov2-/ClueWeb-shaped collections, encoded by the
UNMODIFIED reference encoder (runs only where /root/reference is mounted).

Full-scale law: num_docs documents, `terms` posting lists with
Zipfian lengths n_t = min(num_docs, floor(C / t)), C solved so that sum(n_t) = postings;
d-gaps geometric with the list's density plus bursts of consecutive docIDs; freqs
1 + Geometric(0.6). Encoding 6-15 G postings with the reference's optimal-parse encoder
takes many hours, and the result would not fit a repo snapshot, so what is produced here is
a 1/R stratified SAMPLE of that collection which bench.py replicates R times in HBM:

  * lists with n_t >= R * 4096 contribute a contiguous window of n_t / R postings with the
    list's own density (long lists are cut into ~1 K-integer segments by the decoder anyway);
  * shorter lists are taken whole, every R-th term.

So every (list length class, density) stratum keeps its share of the postings. Dictionaries
are built by the reference's own builder (DSF-65536-16) from a 1/8 subsample of the sample,
then every list is encoded by the reference's vroom_env/encode and checked by
vroom_env/check_encoded_data; the frozen index is built and verified by the reference too.

Outputs: bench_data/<name>/{vroom.<type>.<part>.bin, dict.*, index.<type>.bin, meta.json}
(+ xz copies when --xz). PRNG: numpy PCG64 seeded (seed, term id).
"""
import argparse
import json
import lzma
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFBIN = os.path.join(HERE, "_ref")
TYPES = {"single_rect_dint": "rectangular", "single_packed_dint": "single_packed", "multi_packed_dint": "multi_packed"}

SHAPES = {
    # name: (num_docs, terms, postings, seed)
    "gov2": (25_000_000, 35_000_000, 6_000_000_000, 42),
    "clueweb09": (50_000_000, 90_000_000, 15_000_000_000, 43),
}


def solve_C(num_docs, terms, postings):
    t = np.arange(1, terms + 1, dtype=np.float64)
    lo, hi = 1.0, float(postings)
    for _ in range(100):
        C = (lo + hi) / 2
        s = np.minimum(num_docs, np.floor(C / t)).clip(min=1).sum()
        if s < postings:
            lo = C
        else:
            hi = C
    return hi


def gen_list(rng, n, density, num_docs):
    """n strictly increasing docids with the given density, bursty; returns (docs, freqs)"""
    p = min(max(density, 1e-9), 1.0)
    burst = rng.random(n) < 0.3 / 8  # a burst starts here...
    blen = rng.geometric(1 / 32, size=n)
    in_burst = np.zeros(n, dtype=bool)
    idx = np.nonzero(burst)[0]
    for i, l in zip(idx, blen[idx]):
        in_burst[i:i + l] = True
    frac_burst = in_burst.mean() if n else 0.0
    # keep the mean gap ~ 1/p: non-burst gaps are stretched accordingly
    mean_gap = 1.0 / p
    rest = max(mean_gap - frac_burst, 1.0) / max(1.0 - frac_burst, 1e-6)
    q = min(1.0, 1.0 / rest)
    gaps = rng.geometric(q, size=n).astype(np.int64)
    gaps[in_burst] = 1
    docs = np.cumsum(gaps) - 1
    # stay inside the universe: shrink the (gap - 1) parts, which keeps strict monotonicity
    limit = num_docs - 1
    if n and docs[-1] > limit:
        scale = max(limit - (n - 1), 0) / max(int(docs[-1]) - (n - 1), 1)
        gaps = 1 + np.floor((gaps - 1) * scale).astype(np.int64)
        docs = np.cumsum(gaps) - 1
    assert n == 0 or (docs[-1] <= limit and (n == 1 or np.diff(docs).min() >= 1))
    freqs = rng.geometric(0.6, size=n).astype(np.int64)
    return docs.astype(np.uint32), freqs.astype(np.uint32)


def write_collection(path, lists, num_docs):
    with open(path + ".docs", "wb") as fd, open(path + ".freqs", "wb") as ff, open(path + ".sizes", "wb") as fs:
        np.array([1, num_docs], dtype="<u4").tofile(fd)
        sizes = np.ones(num_docs, dtype="<u4")
        np.array([num_docs], dtype="<u4").tofile(fs)
        sizes.tofile(fs)
        for d, f in lists:
            np.array([len(d)], dtype="<u4").tofile(fd)
            d.astype("<u4").tofile(fd)
            np.array([len(f)], dtype="<u4").tofile(ff)
            f.astype("<u4").tofile(ff)


def sh(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="gov2", choices=sorted(SHAPES))
    ap.add_argument("--replicas", type=int, default=256, help="R: bench.py replicates the sample R times")
    ap.add_argument("--types", default="single_packed_dint,multi_packed_dint")
    ap.add_argument("--index", action="store_true", help="also build frozen dict_freq_index files")
    ap.add_argument("--out", default=os.path.join(ROOT, "bench_data"))
    ap.add_argument("--xz", action="store_true")
    a = ap.parse_args()
    num_docs, terms, postings, seed = SHAPES[a.shape]
    R = a.replicas
    name = f"{a.shape}_r{R}"
    out = os.path.join(a.out, name)
    work = os.path.join(out, "_work")
    os.makedirs(work, exist_ok=True)

    C = solve_C(num_docs, terms, postings)
    lists = []
    n_long = 0
    t = 1
    # long lists: one window of n_t / R postings each
    while True:
        n_t = int(min(num_docs, C // t))
        if n_t < R * 4096:
            break
        rng = np.random.default_rng([seed, t])
        lists.append(gen_list(rng, n_t // R, n_t / num_docs, num_docs))
        n_long += 1
        t += 1
    first_short = t
    # short lists: whole, every R-th term
    for t in range(first_short + R // 2, terms + 1, R):
        n_t = int(max(1, min(num_docs, C // t)))
        rng = np.random.default_rng([seed, t])
        lists.append(gen_list(rng, n_t, n_t / num_docs, num_docs))
    # interleave long and short lists so term order looks like a real (lexicographic) index
    order = np.random.default_rng(seed).permutation(len(lists))
    lists = [lists[i] for i in order]
    total = sum(len(d) for d, _ in lists)
    print(f"{name}: C={C:.0f} long windows={n_long} lists={len(lists)} postings={total} "
          f"(x{R} = {total * R / 1e9:.2f} G of {postings / 1e9:.1f} G)", flush=True)
    base = os.path.join(work, name)
    write_collection(base, lists, num_docs)
    # dictionary-building subsample (every 8th list)
    sub = os.path.join(work, name + "_dictsample")
    write_collection(sub, lists[::8], num_docs)

    meta = {"shape": a.shape, "replicas": R, "num_docs": num_docs, "full_scale_terms": terms,
            "full_scale_postings": postings, "C": C, "seed": seed, "n_lists": len(lists), "n_postings": total,
            "long_windows": n_long, "block_size": 256, "files": {}}
    for type_ in a.types.split(","):
        dname = TYPES[type_]
        sh([os.path.join(REFBIN, "ref_index"), "dict", type_, sub], cwd=work)
        for part in ("docs", "freqs"):
            src = os.path.join(work, f"dict.{name}_dictsample.{part}.{dname}.DSF-65536-16")
            dict_file = f"dict.{name}.{part}.{dname}.DSF-65536-16"
            shutil.copy(src, os.path.join(work, dict_file))  # the index builder looks it up by this name in CWD
            shutil.copy(src, os.path.join(out, dict_file))
            vb = f"vroom.{type_}.{part}.bin"
            sh([os.path.join(REFBIN, "vroom_encode"), type_, f"{base}.{part}", "--dict", dict_file, "--out",
                os.path.join(out, vb)], cwd=work)
            sh([os.path.join(REFBIN, "vroom_check_encoded_data"), type_, f"{base}.{part}", os.path.join(out, vb),
                "--dict", dict_file], cwd=work)
            for f in (dict_file, vb):
                meta["files"][f] = {"bytes": os.path.getsize(os.path.join(out, f))}
        if a.index:
            idx = f"index.{type_}.bin"
            sh([os.path.join(REFBIN, "ref_index"), "build", type_, base, os.path.join(out, idx)], cwd=work)
            meta["files"][idx] = {"bytes": os.path.getsize(os.path.join(out, idx))}
    # golden digests of the raw collection (what the decoders must reproduce)
    import hashlib
    docs_flat = np.concatenate([d for d, _ in lists]).astype("<u4")
    freqs_flat = np.concatenate([f for _, f in lists]).astype("<u4")
    meta["docids_sha256"] = hashlib.sha256(docs_flat.tobytes()).hexdigest()
    meta["freqs_sha256"] = hashlib.sha256(freqs_flat.tobytes()).hexdigest()
    with open(os.path.join(out, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    if a.xz:
        for f in meta["files"]:
            with open(os.path.join(out, f), "rb") as fi, lzma.open(os.path.join(out, f + ".xz"), "wb", preset=6) as fo:
                fo.write(fi.read())
    shutil.rmtree(work)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    sys.exit(main())
