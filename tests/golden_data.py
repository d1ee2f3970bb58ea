"""Access to the committed golden fixtures (tests/golden/*.xz + manifest.json).

The fixtures were produced by the UNMODIFIED reference (oracle/make_fixtures.py) from
test/test_data/test_collection; manifest.json pins their sha256 and the sha256 of the
outputs the reference decoders produce from them.
"""
import hashlib
import json
import lzma
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CACHE = os.environ.get("DINT_GOLDEN_CACHE", os.path.join(GOLD, "_cache"))

TYPES = ("single_rect_dint", "single_packed_dint", "multi_packed_dint")
DICT_NAMES = {"single_rect_dint": "rectangular", "single_packed_dint": "single_packed",
              "multi_packed_dint": "multi_packed"}


def sha(b):
    return hashlib.sha256(b).hexdigest()


class Golden:
    def __init__(self, sub=""):
        self.dir = os.path.join(GOLD, sub)
        self.cache = os.path.join(CACHE, sub)
        with open(os.path.join(self.dir, "manifest.json")) as f:
            self.manifest = json.load(f)
        self.block_size = self.manifest["block_size"]
        self.n_lists = self.manifest["n_lists"]
        self.n_postings = self.manifest["n_postings"]
        self.sha = self.manifest["golden"]

    def path(self, name):
        """decompress (once) and return the path of a raw fixture file"""
        meta = self.manifest["files"][name]
        name = meta.get("same_as", name)
        os.makedirs(self.cache, exist_ok=True)
        p = os.path.join(self.cache, name)
        if not os.path.exists(p) or os.path.getsize(p) != meta["bytes"]:
            with lzma.open(os.path.join(self.dir, name + ".xz")) as fi:
                b = fi.read()
            assert sha(b) == meta["sha256"], f"fixture {name} corrupted"
            tmp = p + f".tmp{os.getpid()}"
            with open(tmp, "wb") as fo:
                fo.write(b)
            os.replace(tmp, p)
        return p

    def bytes(self, name):
        with open(self.path(name), "rb") as f:
            return f.read()

    def dict_name(self, type_, part):
        return f"dict.test_collection.{part}.{DICT_NAMES[type_]}.DSF-65536-16"

    def vroom_name(self, type_, part):
        return f"vroom.{type_}.{part}.bin"

    def index_name(self, type_):
        return f"index.{type_}.bin"


def checksum_lists(values, list_n):
    """per-list 64-bit checksum  sum((v+1)*(pos+1)) mod 2^64  (SURVEY §8d parity row)"""
    values = values.astype(np.uint64)
    out = np.zeros(len(list_n), dtype=np.uint64)
    off = 0
    with np.errstate(over="ignore"):
        for i, n in enumerate(list_n):
            n = int(n)
            v = values[off:off + n]
            out[i] = np.sum((v + np.uint64(1)) * np.arange(1, n + 1, dtype=np.uint64), dtype=np.uint64)
            off += n
    return out
