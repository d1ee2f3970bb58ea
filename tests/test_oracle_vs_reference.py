"""Pins the C restatement against the UNMODIFIED reference compiled into oracle/_ref/libdint_ref.so
(built by `make -C oracle ref` where /root/reference is mounted; the .so travels with the repo
snapshot). Skipped with a clear reason when the reference build is not present."""
import numpy as np
import pytest

import golden_data as g
import stream_builder as sb
from oracle_lib import KINDS, Reference, reference_so

pytestmark = pytest.mark.skipif(reference_so() is None, reason="oracle/_ref/libdint_ref.so not built")


@pytest.fixture(scope="module")
def ref():
    return Reference()


def _lists(plan_bytes, oracle, h, golden, type_, part):
    from dint_b200 import capi
    d = capi.Dictionary(capi.KIND_OF_TYPE[type_], golden.bytes(golden.dict_name(type_, part)))
    return capi.VroomPlan(d, plan_bytes)


@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("part", ("docs", "freqs"))
def test_vroom_file_equal(oracle, ref, golden, type_, part):
    bin_ = golden.bytes(golden.vroom_name(type_, part))
    ho = oracle.dict_parse(KINDS[type_], golden.bytes(golden.dict_name(type_, part)))
    hr = ref.dict_load(KINDS[type_], golden.path(golden.dict_name(type_, part)))
    a, _ = oracle.vroom_decode_file(ho, bin_, golden.n_postings, golden.n_lists)
    b, nl = ref.vroom_decode_file(hr, bin_, golden.n_postings)
    assert nl == golden.n_lists and np.array_equal(a, b)


@pytest.mark.parametrize("type_", g.TYPES)
def test_vroom_per_list_bytes_consumed(oracle, ref, golden, type_):
    """the returned pointer (bytes consumed) must agree list by list"""
    plan = _lists(golden.bytes(golden.vroom_name(type_, "docs")), oracle, None, golden, type_, "docs")
    raw = np.frombuffer(golden.bytes(golden.vroom_name(type_, "docs")), dtype=np.uint8)
    ho = oracle.dict_parse(KINDS[type_], golden.bytes(golden.dict_name(type_, "docs")))
    hr = ref.dict_load(KINDS[type_], golden.path(golden.dict_name(type_, "docs")))
    rng = np.random.default_rng(7)
    lens = np.diff(plan.list_offsets.astype(np.int64))
    pick = set(rng.integers(0, plan.n_lists, 300).tolist()) | set(np.argsort(-lens)[:20].tolist())
    for i in sorted(pick):
        n = int(lens[i])
        start = int(plan.list_byte_offsets[i])
        hdr = np.zeros(2, dtype=np.uint32)
        used_h = oracle.lib.oracle_vbyte_decode(raw[start:].ctypes.data, hdr.ctypes.data, 2)
        assert hdr[0] == n
        body = raw[start + used_h:]
        va, ua = oracle.vroom_decode_list(ho, body, n)
        vb, ub = ref.vroom_decode_list(hr, body, n)
        assert ua == ub == int(plan.list_byte_offsets[i + 1]) - start - used_h
        assert np.array_equal(va, vb)


@pytest.mark.parametrize("sum_known", (True, False))
def test_interpolative_blocks_random(oracle, ref, sum_known):
    """partial blocks (n < 256): interpolative_block::decode, both sum conventions
    (docs: sum given; freqs: sum_of_values == -1 -> leading vbyte), sizes of test_block_codecs.cpp:42-48"""
    rng = np.random.default_rng(11)
    hr = None
    for n in [1, 2, 3, 16, 17, 31, 100, 254, 255]:
        for mag in (1, 2, 7, 100, 10_000, 1_000_000):
            vals = rng.integers(0, mag, n).astype(np.uint32)
            data, s = sb.encode_interpolative_block(vals, sum_known)
            buf = np.frombuffer(data + b"\0" * 16, dtype=np.uint8)
            sov = s if sum_known else 0xFFFFFFFF
            a, ua = oracle.block_decode(None, buf, sov, n)
            if hr is None:
                import golden_data
                G = golden_data.Golden()
                hr = ref.dict_load(1, G.path(G.dict_name("single_packed_dint", "docs")))
            b, ub = ref.block_decode(hr, buf, sov, n)
            assert np.array_equal(a, vals) and np.array_equal(b, vals), (n, mag)
            assert ua == ub == len(data)


@pytest.mark.parametrize("type_", g.TYPES)
def test_synthetic_streams_with_exceptions(oracle, ref, golden, type_):
    """u32 exceptions, payload words equal to 0/1, runs: absent from test_collection"""
    kind = KINDS[type_]
    di = sb.DictImage(kind, golden.bytes(golden.dict_name(type_, "docs")))
    ho = oracle.dict_parse(kind, golden.bytes(golden.dict_name(type_, "docs")))
    hr = ref.dict_load(kind, golden.path(golden.dict_name(type_, "docs")))
    rng = np.random.default_rng(3)
    cases = [
        [0, 1, 0, 1, 65536, 1, 0, 70000, 65537, 0, 0, 1],
        [1 << 31, 0xFFFFFFFF, 65536 + 1, 65536, 1, 0] * 9,
        [0] * 600 + [5] + [0] * 17,
        rng.integers(0, 3, 777).tolist(),
        rng.integers(0, 1 << 20, 300).tolist(),
        [0] * 256, [0] * 255, [7], [0],
    ]
    sels = [0, 6, 3, 9, 1, 7]
    for vals in cases:
        for kw in ({}, {"use_dict": False}, {"use_runs": False}):
            data = sb.encode_vroom_list(di, vals, selectors=sels, **kw)
            hdr = np.zeros(2, dtype=np.uint32)
            raw = np.frombuffer(data + b"\0" * 64, dtype=np.uint8)
            used_h = oracle.lib.oracle_vbyte_decode(raw.ctypes.data, hdr.ctypes.data, 2)
            body = raw[used_h:]
            va, ua = oracle.vroom_decode_list(ho, body, len(vals))
            vb, ub = ref.vroom_decode_list(hr, body, len(vals))
            assert np.array_equal(va, np.asarray(vals, dtype=np.uint32)), (type_, kw)
            assert np.array_equal(va, vb) and ua == ub == len(data) - used_h
