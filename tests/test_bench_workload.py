"""The bench workload itself (bench_data/gov2_r256: the Gov2-shaped synthetic sample encoded by the
reference's own encoder) through the CUDA path: full-output sha256 against the
collection the reference encoded (meta.json), the oracle on a prefix, and the per-list checksum kernel
against numpy. bench.py repeats the sha256 + checksum check at full scale before it prints a number."""
import hashlib
import json
import lzma
import os

import numpy as np
import pytest
import torch

import golden_data as g
from dint_b200 import capi, driver
from oracle_lib import KINDS

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "bench_data", "gov2_r256")


def _read(name):
    p = os.path.join(DATA, name)
    if os.path.exists(p):
        with open(p, "rb") as f:
            return f.read()
    with lzma.open(p + ".xz") as f:
        return f.read()


@pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "meta.json")), reason="bench_data/gov2_r256 not present")
@pytest.mark.parametrize("type_", ("multi_packed_dint", "single_packed_dint"))
def test_bench_sample_decodes_to_the_reference_collection(oracle, type_):
    meta = json.load(open(os.path.join(DATA, "meta.json")))
    dn = g.DICT_NAMES[type_]
    for part, mode, key in (("docs", capi.MODE_DOCIDS, "docids_sha256"), ("freqs", capi.MODE_FREQS, "freqs_sha256")):
        dict_bytes = _read(f"dict.gov2_r256.{part}.{dn}.DSF-65536-16")
        bin_ = _read(f"vroom.{type_}.{part}.bin")
        vf = driver.VroomFile(type_, bin_, dict_bytes)
        assert vf.n_ints == meta["n_postings"] and vf.n_lists == meta["n_lists"]
        out = vf.decode(mode=mode)
        host = out.cpu().numpy().view(np.uint32)
        assert hashlib.sha256(host.tobytes()).hexdigest() == meta[key], (type_, part)
        # per-list checksums: kernel vs numpy on a sample of lists (long, short, first, last)
        sums = vf.checksums(out).cpu().numpy().view(np.uint64)
        lo = vf.plan.list_offsets.astype(np.int64)
        lens = np.diff(lo)
        rng = np.random.default_rng(5)
        pick = np.unique(np.concatenate([[0, vf.n_lists - 1, int(lens.argmax())], rng.integers(0, vf.n_lists, 200)]))
        with np.errstate(over="ignore"):
            for i in pick:
                v = host[lo[i]:lo[i + 1]].astype(np.uint64)
                want = np.sum((v + np.uint64(1)) * np.arange(1, len(v) + 1, dtype=np.uint64), dtype=np.uint64)
                assert sums[i] == want, i
        # raw values of a prefix of the file against the oracle (element by element), both decoders agree on all of it
        raw = vf.decode(mode=capi.MODE_RAW).cpu().numpy().view(np.uint32)
        n_l = 3000
        cut = int(vf.plan.list_byte_offsets[n_l])
        ho = oracle.dict_parse(KINDS[type_], dict_bytes)
        want, ln = oracle.vroom_decode_file(ho, bin_[:cut], int(lo[n_l]), n_l)
        oracle.dict_free(ho)
        assert len(ln) == n_l and np.array_equal(raw[:lo[n_l]], want)
        tiles = driver.VroomFile(type_, bin_, dict_bytes, engine="units").decode(mode=capi.MODE_RAW)
        assert torch.equal(tiles, vf.decode(mode=capi.MODE_RAW))
        del vf, out, tiles
        torch.cuda.empty_cache()
