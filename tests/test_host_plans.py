"""Host format layer (dint_b200/csrc/dint_host.cpp) against the oracle: no GPU needed."""
import numpy as np
import pytest

import golden_data as g
import stream_builder as sb
from dint_b200 import capi
from oracle_lib import KINDS


@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("part", ("docs", "freqs"))
def test_vroom_plan_matches_oracle_walk(oracle, golden, type_, part):
    kind = KINDS[type_]
    bin_ = golden.bytes(golden.vroom_name(type_, part))
    d = capi.Dictionary(kind, golden.bytes(golden.dict_name(type_, part)))
    plan = capi.VroomPlan(d, bin_, chunk_ints=300)
    ho = oracle.dict_parse(kind, golden.bytes(golden.dict_name(type_, part)))
    vals, ln = oracle.vroom_decode_file(ho, bin_, golden.n_postings, golden.n_lists)
    assert plan.n_lists == golden.n_lists and plan.n_ints == golden.n_postings
    assert np.array_equal(np.diff(plan.list_offsets.astype(np.int64)), ln.astype(np.int64))
    assert int(plan.list_byte_offsets[-1]) == len(bin_)
    segs = plan.segments
    # segments tile the output exactly, in order
    assert int(segs["out_off"][0]) == 0
    assert np.array_equal(segs["out_off"][1:], (segs["out_off"] + segs["n"])[:-1])
    assert int(segs["out_off"][-1] + segs["n"][-1]) == golden.n_postings
    # doc_base of a segment = sum(v + 1) over the list's earlier postings (0 at a list start)
    starts = plan.list_offsets[:-1]
    first = np.isin(segs["out_off"], starts)
    if type_ != "multi_packed_dint":
        assert first.sum() == golden.n_lists
    csum = np.concatenate([np.zeros(1, dtype=np.uint64), np.cumsum(vals.astype(np.uint64) + np.uint64(1))])
    list_of = np.searchsorted(plan.list_offsets, segs["out_off"], side="right") - 1
    want = (csum[segs["out_off"].astype(np.int64)] - csum[plan.list_offsets[list_of].astype(np.int64)]) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(segs["doc_base"].astype(np.uint64), want)
    # doc_sum = sum(v + 1) over the segment; in_len = its bytes (segments of one list are back to back)
    ends = (segs["out_off"] + segs["n"]).astype(np.int64)
    assert np.array_equal(segs["doc_sum"].astype(np.uint64),
                          (csum[ends] - csum[segs["out_off"].astype(np.int64)]) & np.uint64(0xFFFFFFFF))
    same_list = list_of[1:] == list_of[:-1]
    assert np.array_equal((segs["in_off"] + segs["in_len"])[:-1][same_list], segs["in_off"][1:][same_list])
    assert int(segs["in_off"][-1] + segs["in_len"][-1]) == len(bin_)
    # work units tile the segment table, in order, none empty
    uf = plan.unit_first.astype(np.int64)
    assert uf[0] == 0 and uf[-1] == len(segs) and np.all(np.diff(uf) > 0) and len(uf) == plan.n_units + 1
    # every segment starts on a codeword boundary: decoding it alone reproduces its slice
    raw = np.frombuffer(bin_ + b"\0" * 64, dtype=np.uint8)
    rng = np.random.default_rng(5)
    for i in rng.integers(0, len(segs), 200):
        s = segs[i]
        a, _ = oracle.vroom_decode_list(ho, raw[int(s["in_off"]):], int(s["n"])) if type_ != "multi_packed_dint" \
            else oracle.block_decode(ho, raw[int(s["in_off"]):], 0, int(s["n"]), block_size=int(s["n"]))
        assert np.array_equal(a, vals[int(s["out_off"]):int(s["out_off"]) + int(s["n"])])


@pytest.mark.parametrize("type_", g.TYPES)
def test_index_plan_matches_oracle(oracle, golden, type_):
    kind = KINDS[type_]
    bin_ = golden.bytes(golden.index_name(type_))
    plan = capi.IndexPlan(kind, bin_)
    ep, info = oracle.index_endpoints(kind, bin_)
    assert plan.n_lists == info.n_lists == golden.n_lists
    assert plan.num_docs == info.num_docs == 10000
    assert (plan.lists_off, plan.lists_len) == (info.lists_off, info.lists_len)
    assert (plan.docs_dict_off, plan.freqs_dict_off, plan.end_off) == (info.docs_dict_off, info.freqs_dict_off, len(bin_))
    assert np.array_equal(plan.endpoints[:-1], ep) and int(plan.endpoints[-1]) == info.lists_len
    assert plan.n_postings == golden.n_postings
    d, f, ln = oracle.index_decode(kind, bin_, golden.n_postings)
    assert np.array_equal(np.diff(plan.list_offsets.astype(np.int64)), ln.astype(np.int64))
    b = plan.blocks
    assert int(b["size"].astype(np.int64).sum()) == golden.n_postings
    # block_max is the docid of the block's last posting; cur_base = previous block_max + 1
    last = (b["out_off"] + b["size"] - 1).astype(np.int64)
    assert np.array_equal(d[last], b["block_max"])
    firsts = np.isin(b["out_off"], plan.list_offsets[:-1])
    assert np.all(b["cur_base"][firsts] == 0)
    assert np.array_equal(b["cur_base"][~firsts], b["block_max"][np.nonzero(~firsts)[0] - 1] + 1)


def test_embedded_dictionaries_equal_dictionary_files(oracle, golden):
    """the dictionaries frozen inside the index are the ones in the dict files"""
    for type_ in g.TYPES:
        kind = KINDS[type_]
        bin_ = golden.bytes(golden.index_name(type_))
        plan = capi.IndexPlan(kind, bin_)
        a = oracle.lib.oracle_dict_parse_mapped(kind, np.frombuffer(bin_, dtype=np.uint8)[plan.docs_dict_off:].ctypes.data,
                                                len(bin_) - plan.docs_dict_off, None)
        b = oracle.dict_parse(kind, golden.bytes(golden.dict_name(type_, "docs")))
        for cls in range(6 if kind == 2 else 1):
            for i in (0, 1, 2, 6) + ((7, 8, 100, 130) if cls < 5 else ()):  # class 5 holds only the reserved 7
                na, ra = oracle.dict_copy(a, cls, i)
                nb, rb = oracle.dict_copy(b, cls, i)
                assert na == nb and np.array_equal(ra[:min(na, 16)], rb[:min(nb, 16)])


def test_malformed_inputs_are_rejected(golden):
    kind = 1
    dict_bytes = golden.bytes(golden.dict_name("single_packed_dint", "docs"))
    with pytest.raises(capi.DintError) as e:
        capi.Dictionary(kind, dict_bytes[:1000])
    assert e.value.code == -2
    with pytest.raises(capi.DintError):
        capi.Dictionary(7, dict_bytes)
    d = capi.Dictionary(kind, dict_bytes)
    bin_ = golden.bytes(golden.vroom_name("single_packed_dint", "docs"))
    with pytest.raises(capi.DintError) as e:
        capi.VroomPlan(d, bin_[:len(bin_) - 3])  # truncated in the middle of the last list
    assert e.value.code == -2
    idx = golden.bytes(golden.index_name("single_packed_dint"))
    with pytest.raises(capi.DintError):
        capi.IndexPlan(kind, idx[:len(idx) - 8])
    with pytest.raises(capi.DintError):
        capi.IndexPlan(0, idx)  # wrong dictionary kind for this file


def _multi_dict_parts(dict_bytes):
    """header words + the three arrays of a multi_packed dictionary file (multi_dictionary.hpp:70-91)"""
    w = np.frombuffer(dict_bytes, dtype="<u4").copy()
    ns, no, nt = int(w[1]), int(w[2]), int(w[3])
    return w, 4, 4 + ns, 4 + ns + no, nt


def test_malformed_dictionaries_are_rejected(golden):
    """what a dictionary must satisfy for the device tables to mean what the reference's copy() does:
    monotone class starts, classes of at most 65536 + 7 entries, exception rows of length 1, entry lengths
    in {1, 2, 4, 8, 16} (the device stores log2(length)) — anything else is DINT_E_FORMAT, not a wrong decode"""
    good = golden.bytes(golden.dict_name("multi_packed_dint", "docs"))
    capi.Dictionary(2, good)  # the fixture itself is fine
    w, so, oo, to, nt = _multi_dict_parts(good)
    bad = w.copy()
    bad[so + 1], bad[so + 2] = w[so + 2], w[so + 1]  # class starts swapped
    with pytest.raises(capi.DintError) as e:
        capi.Dictionary(2, bad.tobytes())
    assert e.value.code == -2
    first_entry = int(w[so]) + 7  # first non-reserved row of class 0
    bad = w.copy()
    bad[oo + first_entry] = (2 << 24) | (int(w[oo + first_entry]) & 0xFFFFFF)  # length 3
    with pytest.raises(capi.DintError) as e:
        capi.Dictionary(2, bad.tobytes())
    assert e.value.code == -2
    bad = w.copy()
    bad[oo + int(w[so])] = (1 << 24)  # the 16-bit exception row claims length 2
    with pytest.raises(capi.DintError) as e:
        capi.Dictionary(2, bad.tobytes())
    assert e.value.code == -2
    bad = w.copy()
    bad[oo + int(w[so]) + 2] = 0x0F000000  # run codeword 2 must be 256 zeros
    with pytest.raises(capi.DintError):
        capi.Dictionary(2, bad.tobytes())
    # single_packed: the same length rule
    sp = np.frombuffer(golden.bytes(golden.dict_name("single_packed_dint", "docs")), dtype="<u4").copy()
    sp[3 + 100] = (4 << 24) | (int(sp[3 + 100]) & 0xFFFFFF)  # length 5
    with pytest.raises(capi.DintError) as e:
        capi.Dictionary(1, sp.tobytes())
    assert e.value.code == -2


def test_corrupt_index_header_is_an_error_not_a_crash(golden):
    """an absurd list count in the header must come back as an error code (no allocation sized by it, no exception
    through the C boundary)"""
    idx = bytearray(golden.bytes(golden.index_name("single_packed_dint")))
    for n in (2 ** 62, 2 ** 40, len(idx) * 9):
        bad = bytearray(idx)
        bad[13:21] = int(n).to_bytes(8, "little")  # m_size, behind the 8-byte flags and 5 global parameters
        with pytest.raises(capi.DintError) as e:
            capi.IndexPlan(1, bytes(bad))
        assert e.value.code in (-2, -4)


def test_empty_vroom_file(golden):
    d = capi.Dictionary(1, golden.bytes(golden.dict_name("single_packed_dint", "docs")))
    plan = capi.VroomPlan(d, b"")
    assert plan.n_lists == 0 and plan.n_ints == 0 and plan.n_segments == 0


def test_chunking_respects_exception_payload(golden):
    """chunk boundaries may never fall between an exception codeword and its payload"""
    di = sb.DictImage(1, golden.bytes(golden.dict_name("single_packed_dint", "docs")))
    d = capi.Dictionary(1, golden.bytes(golden.dict_name("single_packed_dint", "docs")))
    vals = [70000, 1, 0, 65536] * 50
    data = sb.encode_vroom_list(di, vals, use_dict=False, use_runs=False)
    for chunk in (1, 2, 3, 7):
        plan = capi.VroomPlan(d, data, chunk_ints=chunk)
        assert plan.n_ints == len(vals)
        # with only exceptions every integer costs 4 or 6 bytes: offsets must follow that grid
        pos = len(sb.vbyte(len(vals))) + len(sb.vbyte(sum(vals) & 0xFFFFFFFF))
        expect = {}
        for j, v in enumerate(vals):
            expect[j] = pos
            pos += 4 if v < 65536 else 6
        for s in plan.segments:
            assert expect[int(s["out_off"])] == int(s["in_off"])


@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("part", ("docs", "freqs"))
def test_strip_units_tile_the_file(oracle, golden, type_, part):
    """strip table of the lane-sequential decoder: units tile the output, every strip starts on a
    codeword boundary of the right class and decodes (with the oracle) to its slice of the output"""
    kind = KINDS[type_]
    bin_ = golden.bytes(golden.vroom_name(type_, part))
    d = capi.Dictionary(kind, golden.bytes(golden.dict_name(type_, part)))
    plan = capi.VroomPlan(d, bin_)
    lim = capi.strip_limits()
    ho = oracle.dict_parse(kind, golden.bytes(golden.dict_name(type_, part)))
    vals, _ = oracle.vroom_decode_file(ho, bin_, golden.n_postings, golden.n_lists)
    U = plan.strip_units
    assert len(U) == plan.n_strip_units > 0
    n_ints = U["n_ints"].astype(np.int64)
    assert int(U["out_off"][0]) == 0 and np.array_equal(U["out_off"][1:].astype(np.int64), (U["out_off"].astype(np.int64) + n_ints)[:-1])
    assert int(U["out_off"][-1]) + int(n_ints[-1]) == golden.n_postings
    assert n_ints.min() >= 1 and n_ints.max() <= lim["unit_ints"] and int(U["in_len"].max()) <= lim["unit_bytes"]
    assert U["n_strips"].min() >= 1 and U["n_strips"].max() <= 32
    assert np.all(U["in_off"].astype(np.int64) + U["in_len"] <= len(bin_))
    in_rel, out_rel, sel = U["strip"] & 0x1FFF, (U["strip"] >> 13) & 0xFFF, (U["strip"] >> 25) & 15
    k = np.arange(32)[None, :]
    used = k < U["n_strips"][:, None]
    assert np.all(out_rel[:, 0] == 0) and np.all(in_rel[:, 0] == 0)
    assert np.all(out_rel[~used] == np.broadcast_to(U["n_ints"][:, None], out_rel.shape)[~used])
    assert np.all(np.diff(np.where(used, out_rel.astype(np.int64), 1 << 20), axis=1)[used[:, 1:]] > 0)  # no empty strips
    assert np.all(sel[used] < (12 if kind == 2 else 1))
    # docID bookkeeping: doc_base = sum(v + 1) over the list's postings before the unit; list_starts marks list heads
    csum = np.concatenate([np.zeros(1, dtype=np.uint64), np.cumsum(vals.astype(np.uint64) + np.uint64(1))])
    lo = plan.list_offsets.astype(np.int64)
    list_of = np.searchsorted(lo, U["out_off"].astype(np.int64), side="right") - 1
    want = (csum[U["out_off"].astype(np.int64)] - csum[lo[list_of]]) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(U["doc_base"].astype(np.uint64), want)
    pos = U["out_off"].astype(np.int64)[:, None] + out_rel
    is_head = np.isin(pos, lo[:-1]) & used
    is_head[:, 0] = False
    bits = ((U["list_starts"][:, None] >> k) & 1).astype(bool)
    assert np.array_equal(bits, is_head)
    heads_covered = np.isin(lo[:-1], pos[used])
    assert heads_covered.all()  # every list begins a strip
    # decode sampled units strip by strip
    raw = np.frombuffer(bin_ + b"\0" * 64, dtype=np.uint8)
    rng = np.random.default_rng(11)
    flagged = 0
    for ui in np.unique(np.concatenate([rng.integers(0, len(U), 60), [0, len(U) - 1]])):
        u = U[ui]
        ends = list(out_rel[ui][1:u["n_strips"]]) + [int(u["n_ints"])]
        has_run = False
        for j in range(int(u["n_strips"])):
            n = int(ends[j]) - int(out_rel[ui][j])
            start = int(u["in_off"]) + int(in_rel[ui][j])
            # walk the strip's codewords (reference loop, lengths from the oracle's dictionary): <= strip_codewords of
            # them, and DINT_UNIT_HAS_RUNS must be set exactly when one of them is a run
            narrow = int(sel[ui][j]) >= 6
            cls = int(sel[ui][j]) - 6 if narrow else int(sel[ui][j])
            q, produced, cws = start, 0, 0
            while produced < n:
                sym = int(raw[q]) if narrow else int(raw[q]) | (int(raw[q + 1]) << 8)
                q += 1 if narrow else 2
                if sym < 2:
                    q += 2 + 2 * sym
                    produced += 1
                else:
                    produced += oracle.dict_copy(ho, cls, sym)[0]
                    has_run = has_run or sym < 7
                cws += 1
            assert produced == n and 1 <= cws <= lim["strip_codewords"]
            if kind != 2:
                a, used_bytes = oracle.vroom_decode_list(ho, raw[start:], n)
            else:
                fake = np.concatenate([np.array([sel[ui][j]], dtype=np.uint8), raw[start:start + 8 * n + 64]])
                a, used_bytes = oracle.block_decode(ho, fake, 0, n, block_size=n)
                used_bytes -= 1
            o = int(u["out_off"]) + int(out_rel[ui][j])
            assert np.array_equal(a, vals[o:o + n])
            assert start + used_bytes <= int(u["in_off"]) + int(u["in_len"]) and q == start + used_bytes
        assert bool(u["flags"] & 1) == has_run
        flagged += has_run
    cs = plan.codeword_stats
    assert cs["u8_codewords"] + cs["u16_codewords"] == sum(cs[k] for k in ("exc16", "exc32", "runs", "len1", "len2", "len4", "len8", "len16"))
    assert cs["exc16"] + cs["exc32"] + cs["len1"] + 2 * cs["len2"] + 4 * cs["len4"] + 8 * cs["len8"] + 16 * cs["len16"] + \
        cs["run_ints"] == golden.n_postings
    assert cs["strips"] == int(U["n_strips"].sum()) and (cs["u8_codewords"] == 0 or kind == 2)
    assert (cs["runs"] > 0) == bool((U["flags"] & 1).any()) and (flagged == 0 or cs["runs"] > 0)


BLOCK_CODECS = ("block_interpolative", "block_vbyte", "block_u32")  # index_types.hpp:41, 45, 44


@pytest.mark.parametrize("type_", BLOCK_CODECS)
def test_block_codec_index_plan(golden, oracle, type_):
    """the host pre-pass of a dictionary-less block_freq_index: same container as dict_freq_index
    (block_posting_list.hpp mirrors dict_posting_list.hpp), nothing behind m_lists"""
    kind = capi.KIND_OF_TYPE[type_]
    data = golden.bytes(f"index.{type_}.bin")
    plan = capi.IndexPlan(kind, data)
    assert plan.n_lists == golden.n_lists and plan.n_postings == golden.n_postings
    assert plan.docs_dict_off == plan.freqs_dict_off == plan.end_off == len(data)
    info = oracle.index_parse(kind, data)
    assert (plan.lists_off, plan.lists_len) == (info.lists_off, info.lists_len)
    _, _, ln = oracle.index_decode(kind, data, golden.n_postings + 16)
    assert np.array_equal(np.diff(plan.list_offsets.astype(np.int64)), ln.astype(np.int64))
    with pytest.raises(capi.DintError):  # a DINT index is not a block_freq_index: bytes behind m_lists
        capi.IndexPlan(kind, golden.bytes(golden.index_name("single_packed_dint")))
