#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (runs only where /root/reference is mounted). Builds the frozen
block_freq_index<interpolative_block | vbyte_block | u32_block> of test/test_data/test_collection with the UNMODIFIED
reference (oracle/_ref/ref_index build <type>: its own builder + verify_collection), checks that the reference
enumerator's dump of each is the collection, and commits them as tests/golden/index.<type>.bin.xz (+ manifest
entries). SURVEY §8f rank 4: the other block codecs behind the same batch-decode ABI."""
import hashlib, json, lzma, os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DINT_REFERENCE", "/root/reference")
COLL = os.path.join(REF, "test/test_data/test_collection")
GOLD = os.path.join(ROOT, "tests", "golden")
TYPES = sys.argv[1:] or ["block_interpolative", "block_vbyte", "block_u32"]
exe = os.path.join(HERE, "_ref", "ref_index")
subprocess.run(["make", "-C", HERE, "ref", "-j8"], check=True, stdout=subprocess.DEVNULL)
manifest = json.load(open(os.path.join(GOLD, "manifest.json")))
sha = lambda b: hashlib.sha256(b).hexdigest()
for t in TYPES:
    name = f"index.{t}.bin"
    with tempfile.TemporaryDirectory() as tmp:
        idx = os.path.join(tmp, name)
        subprocess.run([exe, "build", t, COLL, idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([exe, "dump", t, idx, os.path.join(tmp, "d.u32"), os.path.join(tmp, "f.u32")], check=True, stderr=subprocess.DEVNULL)
        assert sha(np.fromfile(os.path.join(tmp, "d.u32"), dtype="<u4").tobytes()) == manifest["golden"]["docids_sha256"]
        assert sha(np.fromfile(os.path.join(tmp, "f.u32"), dtype="<u4").tobytes()) == manifest["golden"]["freqs_sha256"]
        b = open(idx, "rb").read()
        manifest["files"][name] = {"bytes": len(b), "sha256": sha(b)}
        with lzma.open(os.path.join(GOLD, name + ".xz"), "wb", preset=6) as fo:
            fo.write(b)
    print(name, manifest["files"][name])
with open(os.path.join(GOLD, "manifest.json"), "w") as fo:
    json.dump(manifest, fo, indent=1, sort_keys=True)
