"""Pins the C restatement (oracle/dint_oracle.c) to the golden fixtures the reference produced.

tests/golden/manifest.json holds the sha256 of what the UNMODIFIED reference decoders output
for each fixture (oracle/make_fixtures.py); both reference checkers accepted the fixtures.
"""
import numpy as np
import pytest

import golden_data as g
from oracle_lib import KINDS

PARTS = (("docs", "vroom_docs_gaps_sha256"), ("freqs", "vroom_freqs_minus1_sha256"))


def test_fixture_integrity(golden):
    for name, meta in golden.manifest["files"].items():
        b = golden.bytes(name)
        assert len(b) == meta["bytes"] and g.sha(b) == meta["sha256"], name


def test_rect_and_packed_streams_identical(golden):
    # SURVEY §8c cross-format pin
    for part in ("docs", "freqs"):
        assert golden.manifest["files"][f"vroom.single_rect_dint.{part}.bin"]["sha256"] == \
            golden.manifest["files"][f"vroom.single_packed_dint.{part}.bin"]["sha256"]


@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("part,key", PARTS)
def test_oracle_vroom_matches_reference_digest(oracle, golden, type_, part, key):
    h = oracle.dict_parse(KINDS[type_], golden.bytes(golden.dict_name(type_, part)))
    out, ln = oracle.vroom_decode_file(h, golden.bytes(golden.vroom_name(type_, part)), golden.n_postings,
                                       golden.n_lists, golden.block_size)
    oracle.dict_free(h)
    assert len(ln) == golden.n_lists and len(out) == golden.n_postings
    assert g.sha(out.astype("<u4").tobytes()) == golden.sha[key]


@pytest.mark.parametrize("type_", g.TYPES)
def test_oracle_index_matches_reference_digest(oracle, golden, type_):
    d, f, ln = oracle.index_decode(KINDS[type_], golden.bytes(golden.index_name(type_)), golden.n_postings,
                                   golden.block_size)
    assert len(ln) == golden.n_lists and int(ln.sum()) == golden.n_postings
    assert g.sha(d.astype("<u4").tobytes()) == golden.sha["docids_sha256"]
    assert g.sha(f.astype("<u4").tobytes()) == golden.sha["freqs_sha256"]


@pytest.mark.parametrize("type_", ("single_rect_dint", "single_packed_dint"))
def test_oracle_index_block128_matches_reference_digest(oracle, golden128, type_):
    """config 2 as BASELINE.json words it: 128-integer blocks (the reference ships 256)"""
    assert golden128.block_size == 128
    d, f, ln = oracle.index_decode(KINDS[type_], golden128.bytes(golden128.index_name(type_)), golden128.n_postings, 128)
    assert len(ln) == golden128.n_lists and int(ln.sum()) == golden128.n_postings
    assert g.sha(d.astype("<u4").tobytes()) == golden128.sha["docids_sha256"]
    assert g.sha(f.astype("<u4").tobytes()) == golden128.sha["freqs_sha256"]


def test_reserved_codewords(oracle, golden):
    """codeword space of SURVEY §9: 0,1 exceptions (len 1); 2..6 runs of 256..16 zeros"""
    for type_ in g.TYPES:
        h = oracle.dict_parse(KINDS[type_], golden.bytes(golden.dict_name(type_, "docs")))
        for cls in range(6 if type_ == "multi_packed_dint" else 1):
            for i, want in zip(range(2, 7), (256, 128, 64, 32, 16)):
                n, row = oracle.dict_copy(h, cls, i)
                assert n == want and not row.any()
        oracle.dict_free(h)


def test_rect_and_packed_dictionaries_entrywise_identical(oracle, golden):
    for part in ("docs", "freqs"):
        a = oracle.dict_parse(0, golden.bytes(golden.dict_name("single_rect_dint", part)))
        b = oracle.dict_parse(1, golden.bytes(golden.dict_name("single_packed_dint", part)))
        rng = np.random.default_rng(1)
        for i in list(range(0, 64)) + rng.integers(7, 65536, 2000).tolist():
            na, ra = oracle.dict_copy(a, 0, i)
            nb, rb = oracle.dict_copy(b, 0, i)
            if i >= 2:
                assert na == nb, i
            assert np.array_equal(ra[:min(na, 16)], rb[:min(na, 16)])
        oracle.dict_free(a)
        oracle.dict_free(b)


@pytest.mark.parametrize("kind,type_", ((3, "block_interpolative"), (4, "block_vbyte"), (5, "block_u32")))
def test_block_codec_index_decodes_to_the_collection(golden, oracle, kind, type_):
    """block_freq_index<interpolative_block | vbyte_block | u32_block> of the test collection, built and verified by
    the reference (oracle/make_block_codec_fixture.py): the C restatement must reproduce the collection from each"""
    import hashlib
    data = golden.bytes(f"index.{type_}.bin")
    d, f, ln = oracle.index_decode(kind, data, golden.n_postings + 16)
    assert len(d) == golden.n_postings and len(ln) == golden.n_lists
    assert hashlib.sha256(d.tobytes()).hexdigest() == golden.sha["docids_sha256"]
    assert hashlib.sha256(f.tobytes()).hexdigest() == golden.sha["freqs_sha256"]
