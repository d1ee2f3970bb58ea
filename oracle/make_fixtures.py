#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — generates the golden fixtures under tests/golden/.

Runs ONLY where /root/reference is mounted (this container). Everything here is
produced by the UNMODIFIED reference code compiled by oracle/Makefile:

  * dictionaries  : dict_freq_index::builder::build_model  (include/dint/dict_freq_index.hpp:53-68,139-161)
  * index .bin    : builder.build + succinct::mapper::freeze, checked by the
                    reference's verify_collection (include/ds2i/verify_collection.hpp:7-52)
  * vroom .bin    : vroom_env/encode.cpp:133-191, checked by the reference's
                    vroom_env/check_encoded_data.cpp:20-119
  * golden output : sha256 of what the reference decoders produce (libdint_ref.so /
                    ref_index dump), recorded in tests/golden/manifest.json

Usage:  python oracle/make_fixtures.py [--bs 256] [--seed-dir /tmp/probe/run]
`--seed-dir` may hold dictionary files from an earlier run of the same reference
builder (multi_packed takes ~10 min to build); they are copied instead of rebuilt.
"""
import argparse
import ctypes
import hashlib
import json
import lzma
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DINT_REFERENCE", "/root/reference")
COLL = os.path.join(REF, "test/test_data/test_collection")
TYPES = {"single_rect_dint": ("rectangular", 0), "single_packed_dint": ("single_packed", 1),
         "multi_packed_dint": ("multi_packed", 2)}


def sh(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def read_collection(path, skip_first):
    a = np.fromfile(path, dtype="<u4")
    lists, i = [], 0
    while i < len(a):
        n = int(a[i])
        lists.append(a[i + 1:i + 1 + n])
        i += 1 + n
    return lists[1:] if skip_first else lists


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=256)
    ap.add_argument("--seed-dir", default="/tmp/probe/run")
    ap.add_argument("--no-commit-fixtures", action="store_true")
    ap.add_argument("--types", default=",".join(TYPES), help="comma-separated subset of the codec types")
    ap.add_argument("--index-only", action="store_true",
                    help="frozen dict_freq_index files only (the vroom streams of the single dictionaries do not depend on block_size)")
    args = ap.parse_args()

    sh(["make", "-C", HERE, "ref", f"BS={args.bs}", "-j8"])
    out = os.path.join(HERE, "_ref") if args.bs == 256 else os.path.join(HERE, "_ref", f"bs{args.bs}")
    data = os.path.join(out, "data", "test_collection")
    os.makedirs(data, exist_ok=True)

    # dictionaries do not depend on block_size for the single kinds; multi statistics do
    # (statistics_collectors.hpp:93-104), so seeds are only reused at the shipped block size.
    types = {t: TYPES[t] for t in args.types.split(",")}
    for t, (dname, _) in types.items():
        for part in ("docs", "freqs"):
            f = f"dict.test_collection.{part}.{dname}.DSF-65536-16"
            src = os.path.join(args.seed_dir, f)
            if os.path.exists(src) and not os.path.exists(os.path.join(data, f)) and \
                    (args.bs == 256 or dname != "multi_packed"):
                shutil.copy(src, os.path.join(data, f))

    manifest = {"block_size": args.bs, "collection": "test/test_data/test_collection", "files": {}, "golden": {}}
    docs = read_collection(COLL + ".docs", True)
    freqs = read_collection(COLL + ".freqs", False)
    assert len(docs) == len(freqs)
    docs_flat = np.concatenate(docs).astype("<u4")
    freqs_flat = np.concatenate(freqs).astype("<u4")
    gaps = []
    for d in docs:
        g = np.empty_like(d)
        g[0] = d[0]
        g[1:] = d[1:] - d[:-1] - 1
        gaps.append(g)
    gaps_flat = np.concatenate(gaps).astype("<u4")
    manifest["n_lists"] = len(docs)
    manifest["n_postings"] = int(len(docs_flat))
    manifest["golden"]["docids_sha256"] = sha(docs_flat.tobytes())
    manifest["golden"]["freqs_sha256"] = sha(freqs_flat.tobytes())
    manifest["golden"]["vroom_docs_gaps_sha256"] = sha(gaps_flat.tobytes())
    manifest["golden"]["vroom_freqs_minus1_sha256"] = sha((freqs_flat - 1).astype("<u4").tobytes())

    lib = ctypes.CDLL(os.path.join(out, "libdint_ref.so"))
    lib.ref_dict_load.restype = ctypes.c_void_p
    lib.ref_dict_load.argtypes = [ctypes.c_int, ctypes.c_char_p]
    lib.ref_vroom_decode_file.restype = ctypes.c_long
    lib.ref_vroom_decode_file.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
    assert lib.ref_block_size() == args.bs

    for t, (dname, kind) in types.items():
        idx = f"index.{t}.bin"
        sh([os.path.join(out, "ref_index"), "build", t, COLL, idx], cwd=data)  # also builds missing dicts
        sh([os.path.join(out, "ref_index"), "dump", t, idx, "dump.docs.u32", "dump.freqs.u32"], cwd=data)
        dd = np.fromfile(os.path.join(data, "dump.docs.u32"), dtype="<u4")
        df = np.fromfile(os.path.join(data, "dump.freqs.u32"), dtype="<u4")
        assert sha(dd.tobytes()) == manifest["golden"]["docids_sha256"], "reference enumerator != collection"
        assert sha(df.tobytes()) == manifest["golden"]["freqs_sha256"]
        os.remove(os.path.join(data, "dump.docs.u32"))
        os.remove(os.path.join(data, "dump.freqs.u32"))
        for part, want in (() if args.index_only else (("docs", gaps_flat), ("freqs", (freqs_flat - 1).astype("<u4")))):
            dict_file = f"dict.test_collection.{part}.{dname}.DSF-65536-16"
            vb = f"vroom.{t}.{part}.bin"
            sh([os.path.join(out, "vroom_encode"), t, f"{COLL}.{part}", "--dict", dict_file, "--out", vb], cwd=data)
            sh([os.path.join(out, "vroom_check_encoded_data"), t, f"{COLL}.{part}", vb, "--dict", dict_file],
               cwd=data)
            # golden output through the reference decoder itself
            h = lib.ref_dict_load(kind, os.path.join(data, dict_file).encode())
            assert h
            raw = np.fromfile(os.path.join(data, vb), dtype=np.uint8)
            outbuf = np.zeros(len(want) + 16, dtype="<u4")
            tot = ctypes.c_uint64(0)
            nl = lib.ref_vroom_decode_file(h, raw.ctypes.data, len(raw), outbuf.ctypes.data, len(outbuf),
                                           ctypes.byref(tot))
            assert nl == len(docs) and tot.value == len(want), (nl, tot.value)
            assert np.array_equal(outbuf[:len(want)], want), f"reference decoder output != collection for {vb}"
            for f in (dict_file, vb):
                b = open(os.path.join(data, f), "rb").read()
                manifest["files"][f] = {"bytes": len(b), "sha256": sha(b)}
        b = open(os.path.join(data, idx), "rb").read()
        manifest["files"][idx] = {"bytes": len(b), "sha256": sha(b)}

    # rect and packed vroom streams are byte-identical (SURVEY §8c cross-format pin)
    for part in (() if args.index_only or not {"single_rect_dint", "single_packed_dint"} <= set(types) else ("docs", "freqs")):
        a = manifest["files"][f"vroom.single_rect_dint.{part}.bin"]["sha256"]
        b = manifest["files"][f"vroom.single_packed_dint.{part}.bin"]["sha256"]
        assert a == b, "rect/packed vroom streams differ"

    if not args.no_commit_fixtures:
        gold = os.path.join(ROOT, "tests", "golden") if args.bs == 256 else \
            os.path.join(ROOT, "tests", "golden", f"bs{args.bs}")
        os.makedirs(gold, exist_ok=True)
        for f in manifest["files"]:
            if f.startswith("vroom.single_rect_dint"):
                manifest["files"][f]["same_as"] = f.replace("single_rect_dint", "single_packed_dint")
                continue
            with open(os.path.join(data, f), "rb") as fi, lzma.open(os.path.join(gold, f + ".xz"), "wb", preset=9) as fo:
                fo.write(fi.read())
        with open(os.path.join(gold, "manifest.json"), "w") as fo:
            json.dump(manifest, fo, indent=1, sort_keys=True)
    print(json.dumps(manifest["golden"], indent=1))


if __name__ == "__main__":
    sys.exit(main())
