"""Parity tests proper: the CUDA path, called through the C ABI, against
  (1) the sha256 of the reference decoders' own output (tests/golden/manifest.json), and
  (2) the oracle (oracle/dint_oracle.c) on the same inputs, element by element.
Bit-exact: this is integer work, there is no tolerance."""
import os

import numpy as np
import pytest
import torch

import golden_data as g
import stream_builder as sb
from dint_b200 import capi, driver
from oracle_lib import KINDS

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.fixture(scope="module")
def vroom_files(golden):
    cache = {}

    def get(type_, part, chunk_ints=0, engine="strips"):
        key = (type_, part, chunk_ints, engine)
        if key not in cache:
            cache[key] = driver.VroomFile(type_, golden.bytes(golden.vroom_name(type_, part)),
                                          golden.bytes(golden.dict_name(type_, part)), chunk_ints=chunk_ints, engine=engine)
        return cache[key]
    return get


ENGINES = ("strips", "units")  # lane-sequential strip decoder (default) / warp-per-unit tile decoder


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("part,key", (("docs", "vroom_docs_gaps_sha256"), ("freqs", "vroom_freqs_minus1_sha256")))
def test_vroom_raw_equals_reference_output(golden, oracle, vroom_files, type_, part, key, engine):
    """config 1: vroom decode of the whole test collection, full-output comparison"""
    vf = vroom_files(type_, part, engine=engine)
    before = capi.kernel_launches()
    out = _np(vf.decode(mode=capi.MODE_RAW))
    assert capi.kernel_launches() == before + 1
    assert len(out) == golden.n_postings
    assert g.sha(out.tobytes()) == golden.sha[key]
    ho = oracle.dict_parse(KINDS[type_], golden.bytes(golden.dict_name(type_, part)))
    want, _ = oracle.vroom_decode_file(ho, golden.bytes(golden.vroom_name(type_, part)), golden.n_postings, golden.n_lists)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("engine,chunk", (("strips", 0), ("units", 0), ("units", 1), ("units", 33), ("units", 100)))
@pytest.mark.parametrize("type_", g.TYPES)
def test_vroom_fused_docids_and_freqs(golden, vroom_files, type_, engine, chunk):
    """fused prefix sum / +1 shift must reproduce the .docs / .freqs collection itself"""
    d = _np(vroom_files(type_, "docs", chunk, engine).decode(mode=capi.MODE_DOCIDS))
    assert g.sha(d.tobytes()) == golden.sha["docids_sha256"]
    f = _np(vroom_files(type_, "freqs", chunk, engine).decode(mode=capi.MODE_FREQS))
    assert g.sha(f.tobytes()) == golden.sha["freqs_sha256"]
    r = _np(vroom_files(type_, "docs", chunk, engine).decode(mode=capi.MODE_RAW))
    assert g.sha(r.tobytes()) == golden.sha["vroom_docs_gaps_sha256"]


@pytest.mark.parametrize("type_", g.TYPES)
def test_index_scan_equals_reference_enumerator(golden, oracle, type_):
    """config 2: every docid()/freq() of the frozen dict_freq_index (verify_collection semantics)"""
    idx = driver.IndexFile(type_, golden.bytes(golden.index_name(type_)))
    d, f = idx.scan()
    d, f = _np(d), _np(f)
    assert g.sha(d.tobytes()) == golden.sha["docids_sha256"]
    assert g.sha(f.tobytes()) == golden.sha["freqs_sha256"]
    wd, wf, _ = oracle.index_decode(KINDS[type_], golden.bytes(golden.index_name(type_)), golden.n_postings)
    assert np.array_equal(d, wd) and np.array_equal(f, wf)
    # either stream alone
    d2, _ = idx.scan(want_freqs=False)
    _, f2 = idx.scan(want_docs=False)
    assert np.array_equal(_np(d2), wd) and np.array_equal(_np(f2), wf)


@pytest.mark.parametrize("type_", ("single_rect_dint", "single_packed_dint"))
def test_index_scan_block128(golden128, oracle, type_):
    """config 2 with 128-integer blocks (BASELINE.json configs[1]); the index was built by the reference
    compiled with constants::block_size = 128"""
    data = golden128.bytes(golden128.index_name(type_))
    idx = driver.IndexFile(type_, data, block_size=128)
    d, f = idx.scan()
    d, f = _np(d), _np(f)
    assert g.sha(d.tobytes()) == golden128.sha["docids_sha256"]
    assert g.sha(f.tobytes()) == golden128.sha["freqs_sha256"]
    wd, wf, _ = oracle.index_decode(KINDS[type_], data, golden128.n_postings, 128)
    assert np.array_equal(d, wd) and np.array_equal(f, wf)


def test_query_log_ragged_batch(golden, oracle):
    """config 3: ragged batch = only the posting lists touched by the reference's query log
    (tests/golden/queries.json <- test/test_data/queries: 500 AND queries, 1088 distinct terms,
    731 581 postings), decoded in one launch"""
    import json
    type_ = "single_packed_dint"
    vf = driver.VroomFile(type_, golden.bytes(golden.vroom_name(type_, "docs")), golden.bytes(golden.dict_name(type_, "docs")))
    with open(os.path.join(g.GOLD, "queries.json")) as fq:
        queries = json.load(fq)["queries"]
    terms = np.array(sorted({t for q in queries for t in q}), dtype=np.int64)
    assert len(queries) == 500 and len(terms) == 1088 and terms.max() < vf.n_lists
    lo = vf.plan.list_offsets.astype(np.int64)
    assert int((lo[terms + 1] - lo[terms]).sum()) == 731581
    segs = vf.plan.segments
    list_of = np.searchsorted(vf.plan.list_offsets, segs["out_off"], side="right") - 1
    pick = segs[np.isin(list_of, terms)].copy()
    lens = pick["n"].astype(np.int64)
    new_off = np.concatenate([[0], np.cumsum(lens)])
    src_off = pick["out_off"].astype(np.int64)
    pick["out_off"] = new_off[:-1]
    d_segs = torch.from_numpy(pick.view(np.uint8).reshape(-1)).cuda()
    out = torch.zeros(int(new_off[-1]), dtype=torch.int32, device="cuda")
    capi.decode_segments(vf.dict, vf.d_in.data_ptr(), vf.plan.in_bytes, d_segs.data_ptr(), len(pick), out.data_ptr(),
                         capi.MODE_DOCIDS, stream=torch.cuda.current_stream().cuda_stream)
    full = _np(vf.decode(mode=capi.MODE_DOCIDS))
    want = np.concatenate([full[s:s + n] for s, n in zip(src_off, lens)])
    assert np.array_equal(_np(out), want)


def _decode_lists(type_, dict_bytes, lists_vals, block_size=256, chunk_ints=0, mode=capi.MODE_RAW, engine="strips", **kw):
    di = sb.DictImage(capi.KIND_OF_TYPE[type_], dict_bytes)
    data = b"".join(sb.encode_vroom_list(di, v, block_size=block_size, **kw) for v in lists_vals)
    vf = driver.VroomFile(type_, data, dict_bytes, chunk_ints=chunk_ints, block_size=block_size, engine=engine)
    return _np(vf.decode(mode=mode)), data, vf


EDGE_LISTS = [
    [0, 1, 0, 1, 65536, 1, 0, 70000, 65537, 0, 0, 1],
    [1 << 31, 0xFFFFFFFF, 65537, 65536, 1, 0] * 9,
    [0] * 600 + [5] + [0] * 17,
    [0] * 256, [0] * 255, [0] * 16, [7], [0], [1], [65535], [65536],
    [3] * 31, [3] * 32, [3] * 33, [70000] * 31 + [1], [70000] * 32, [1, 0] * 40,
]


@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("kw", ({}, {"use_dict": False}, {"use_runs": False, "use_dict": False}))
def test_edge_case_streams(golden, oracle, type_, kw):
    """32-bit exceptions, payload equal to 0/1, exceptions straddling 32-symbol tiles, runs,
    single-integer lists, 8-bit blocks (multi): none of these occur in test_collection"""
    rng = np.random.default_rng(9)
    lists = EDGE_LISTS + [rng.integers(0, 3, 777).tolist(), rng.integers(0, 1 << 20, 300).tolist(),
                          rng.integers(0, 70000, 1000).tolist()]
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    want = np.concatenate([np.asarray(v, dtype=np.uint32) for v in lists])
    for sels in ([0, 6, 3, 9, 1, 7], [8], [2]):
        for engine, chunk in (("strips", 0), ("units", 0), ("units", 1), ("units", 5)):
            out, data, vf = _decode_lists(type_, dict_bytes, lists, chunk_ints=chunk, engine=engine, selectors=sels, **kw)
            assert np.array_equal(out, want), (sels, engine, chunk)
            ho = oracle.dict_parse(KINDS[type_], dict_bytes)
            o, _ = oracle.vroom_decode_file(ho, data, len(want), len(lists))
            assert np.array_equal(o, want)
            # fused modes against numpy
            docs = _np(vf.decode(mode=capi.MODE_DOCIDS))
            exp = np.concatenate([(np.cumsum(np.asarray(v, dtype=np.uint64) + 1) - 1).astype(np.uint32) for v in lists])
            assert np.array_equal(docs, exp)
            assert np.array_equal(_np(vf.decode(mode=capi.MODE_FREQS)), want + np.uint32(1))


def _random_lists(di, rng, n_lists, long_list):
    """lists assembled from real dictionary entries of every length (hot rows, cold rows, all classes), bursts
    of zeros (run codewords), small values and >= 65 536 values (32-bit exceptions)"""
    pools = []
    for ents in di.classes:
        by_len = {}
        for i, t in enumerate(ents[:65536]):
            if t is not None:
                by_len.setdefault(len(t), []).append(t)
        if by_len:  # the last class of a multi dictionary may hold nothing but the reserved rows
            pools.append(by_len)
    lists = []
    for li in range(n_lists):
        n = long_list if li == 0 else int(rng.choice([1, 2, 3, 7, 15, 16, 17, 31, 100, 255, 256, 257, 700, 3000]))
        vals = []
        bursts = (li // 15) % 2 == 0  # stretches of lists without zero bursts: units without run codewords
        while len(vals) < n:
            kind = rng.integers(0, 10)
            if kind in (5, 6) and not bursts:
                kind = 7
            if kind < 5:
                by_len = pools[int(rng.integers(0, len(pools)))]
                ln = int(rng.choice(list(by_len)))
                ents = by_len[ln]
                vals.extend(ents[int(rng.integers(0, len(ents)))])
            elif kind < 6:
                vals.extend([0] * int(rng.choice([1, 5, 16, 17, 40, 128, 300])))
            elif kind < 9:
                vals.extend(rng.integers(0, 70000, int(rng.integers(1, 6))).tolist())
            else:
                vals.extend(rng.integers(65536, 1 << 32, int(rng.integers(1, 4))).tolist())
        lists.append(vals[:n])
    return lists


@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("seed", (1, 2))
def test_random_streams_property(golden, oracle, type_, seed):
    """randomised lists (one of 60 000 integers) through both decoders and all modes == oracle == numpy"""
    rng = np.random.default_rng(seed)
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs" if seed == 1 else "freqs"))
    di = sb.DictImage(capi.KIND_OF_TYPE[type_], dict_bytes)
    lists = _random_lists(di, rng, 120, 60000)
    want = np.concatenate([np.asarray(v, dtype=np.uint32) for v in lists])
    sels = rng.integers(0, 12, 64).tolist()
    ho = oracle.dict_parse(KINDS[type_], dict_bytes)
    for engine in ENGINES:
        out, data, vf = _decode_lists(type_, dict_bytes, lists, engine=engine, selectors=sels)
        o, _ = oracle.vroom_decode_file(ho, data, len(want), len(lists))
        assert np.array_equal(o, want)  # the stream is valid and says what we meant
        assert np.array_equal(out, want), engine
        exp = np.concatenate([(np.cumsum(np.asarray(v, dtype=np.uint64) + 1) - 1).astype(np.uint32) for v in lists])
        assert np.array_equal(_np(vf.decode(mode=capi.MODE_DOCIDS)), exp), engine
        assert np.array_equal(_np(vf.decode(mode=capi.MODE_FREQS)), want + np.uint32(1)), engine
        if engine == "strips":  # the run flag is per unit: some units of this stream have runs, some do not
            fl = vf.plan.strip_units["flags"] & 1
            assert fl.any() and not fl.all()
    oracle.dict_free(ho)


def _synthetic_packed_dictionary(n_classes, rows_per_class, seed):
    """a single_packed (n_classes = 1) or multi_packed dictionary FILE whose entry lengths change from row to row, so
    that every group of 32 rows is mixed: more mixed groups than the decoder's per-class detail table holds (246),
    which forces the fall-back to the global 4-bit length table (no reference-built dictionary does that)"""
    rng = np.random.default_rng(seed)
    offsets, starts, table = [], [], [0] * 16
    for c in range(n_classes):
        starts.append(len(offsets))
        offsets += [0, 0, 0xFF000000, 0x7F000000, 0x3F000000, 0x1F000000, 0x0F000000]  # single_dictionary.hpp:43-55
        for i in range(7, rows_per_class):
            ln = (1, 2, 4, 8, 16)[int(rng.integers(0, 5))]
            offsets.append(((ln - 1) << 24) | len(table))
            # distinct tuples: the first value identifies (class, row); keep them < 65536 so that the greedy test
            # encoder never confuses an entry with an exception payload
            table += [(i * 7 + c) % 65521] + rng.integers(0, 50, ln - 1).tolist()
    table += [0] * 16
    if n_classes == 1:
        hdr = [rows_per_class, len(offsets), len(table)]
    else:
        hdr = [rows_per_class, n_classes, len(offsets), len(table)] + starts
    return np.array(hdr + offsets + table, dtype="<u4").tobytes()


@pytest.mark.parametrize("type_,n_classes", (("single_packed_dint", 1), ("multi_packed_dint", 6)))
def test_dictionary_with_mixed_lengths_in_every_group(oracle, type_, n_classes):
    """cold rows whose lengths come from the per-class detail table AND from the global length table"""
    dict_bytes = _synthetic_packed_dictionary(n_classes, 20000, 3)
    kind = capi.KIND_OF_TYPE[type_]
    di = sb.DictImage(kind, dict_bytes)
    rng = np.random.default_rng(4)
    lists = []
    for li in range(60):
        vals = []
        n = int(rng.choice([5, 40, 300, 2000]))
        cls = int(rng.integers(0, n_classes))
        while len(vals) < n:
            row = int(rng.integers(7, 20000))  # mostly beyond the hot rows, all over the group table
            vals.extend(di.classes[cls][row])
            if rng.integers(0, 8) == 0:
                vals.extend([0] * int(rng.choice([3, 16, 40])))
            if rng.integers(0, 8) == 0:
                vals.append(int(rng.integers(65536, 1 << 32)))
        lists.append(vals[:n])
    want = np.concatenate([np.asarray(v, dtype=np.uint32) for v in lists])
    ho = oracle.dict_parse(KINDS[type_], dict_bytes)
    sels = [0, 1, 2, 3, 4, 5, 7, 9] if n_classes > 1 else None
    for engine in ENGINES:
        out, data, vf = _decode_lists(type_, dict_bytes, lists, engine=engine, selectors=sels)
        o, _ = oracle.vroom_decode_file(ho, data, len(want), len(lists))
        assert np.array_equal(o, want)
        assert np.array_equal(out, want), engine
        exp = np.concatenate([(np.cumsum(np.asarray(v, dtype=np.uint64) + 1) - 1).astype(np.uint32) for v in lists])
        assert np.array_equal(_np(vf.decode(mode=capi.MODE_DOCIDS)), exp), engine
    cs = vf.plan.codeword_stats
    assert cs["u16_codewords"] > 1000 and min(cs[k] for k in ("len1", "len2", "len4", "len8", "len16")) > 50
    oracle.dict_free(ho)


@pytest.mark.parametrize("type_", ("single_packed_dint", "multi_packed_dint"))
@pytest.mark.parametrize("shift", (1, 2, 3))
def test_strips_unaligned_output_and_input(golden, type_, shift):
    """output array / .bin that do not start on a 16-byte boundary: the strip decoder falls back to
    scalar stores and stages the input from the enclosing aligned blocks"""
    bin_ = golden.bytes(golden.vroom_name(type_, "docs"))
    vf = driver.VroomFile(type_, bin_, golden.bytes(golden.dict_name(type_, "docs")))
    want = _np(vf.decode(mode=capi.MODE_DOCIDS))
    assert g.sha(want.tobytes()) == golden.sha["docids_sha256"]
    out = torch.full((vf.n_ints + 8,), -1, dtype=torch.int32, device="cuda")
    d_in = torch.zeros(len(bin_) + 64, dtype=torch.uint8, device="cuda")
    d_in[shift * 5:shift * 5 + len(bin_)] = vf.d_in
    capi.decode_strips(vf.dict, d_in.data_ptr() + shift * 5, len(bin_), vf.d_strip_units.data_ptr(),
                       int(vf.plan.n_strip_units), out.data_ptr() + 4 * shift, capi.MODE_DOCIDS,
                       stream=torch.cuda.current_stream().cuda_stream)
    got = _np(out)
    assert np.array_equal(got[shift:shift + vf.n_ints], want)
    assert np.all(got[:shift] == 0xFFFFFFFF) and np.all(got[shift + vf.n_ints:] == 0xFFFFFFFF)


@pytest.mark.parametrize("type_", ("single_packed_dint", "multi_packed_dint"))
@pytest.mark.parametrize("shift", (0, 1, 2, 3))
def test_strips_freqs_bulk_flush_and_its_fallback(golden, type_, shift):
    """frequency streams leave the shared-memory image with one bulk copy per unit (cp.async.bulk) when the output is
    16-byte aligned, and with the per-lane flush otherwise: both must give the collection's freqs, and nothing may be
    written outside the output slice (partial quads at both ends of every unit)"""
    bin_ = golden.bytes(golden.vroom_name(type_, "freqs"))
    vf = driver.VroomFile(type_, bin_, golden.bytes(golden.dict_name(type_, "freqs")))
    out = torch.full((vf.n_ints + 8,), -1, dtype=torch.int32, device="cuda")
    capi.decode_strips(vf.dict, vf.d_in.data_ptr(), len(bin_), vf.d_strip_units.data_ptr(), int(vf.plan.n_strip_units),
                       out.data_ptr() + 4 * shift, capi.MODE_FREQS, stream=torch.cuda.current_stream().cuda_stream)
    got = _np(out)
    assert g.sha(got[shift:shift + vf.n_ints].tobytes()) == golden.sha["freqs_sha256"]
    assert np.all(got[:shift] == 0xFFFFFFFF) and np.all(got[shift + vf.n_ints:] == 0xFFFFFFFF)


def test_strips_rejects_units_beyond_the_limits(golden, vroom_files):
    """a unit the decoder cannot hold (corrupt table: too many integers / bytes / strips, or input past the end)
    is skipped: its slice of the output stays untouched, every other unit decodes, nothing faults"""
    vf = vroom_files("single_packed_dint", "docs")
    good = _np(vf.decode(mode=capi.MODE_DOCIDS))
    U = vf.plan.strip_units.copy()
    lim = capi.strip_limits()
    bad = {3: ("n_ints", lim["unit_ints"] + 4), 10: ("in_len", lim["unit_bytes"] + 1), 17: ("n_strips", 33),
           25: ("in_off", vf.plan.in_bytes), 40: ("n_ints", 0)}
    for i, (field, value) in bad.items():
        U[field][i] = value
    d_units = torch.from_numpy(U.view(np.uint8).reshape(-1)).cuda()
    out = torch.full((vf.n_ints,), -559038737, dtype=torch.int32, device="cuda")  # 0xDEADBEEF
    capi.decode_strips(vf.dict, vf.d_in.data_ptr(), vf.plan.in_bytes, d_units.data_ptr(), len(U), out.data_ptr(),
                       capi.MODE_DOCIDS, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = _np(out)
    orig = vf.plan.strip_units
    skipped = np.zeros(vf.n_ints, dtype=bool)
    for i in bad:
        o, n = int(orig["out_off"][i]), int(orig["n_ints"][i])
        skipped[o:o + n] = True
    assert np.all(got[skipped] == 0xDEADBEEF)
    assert np.array_equal(got[~skipped], good[~skipped])


def test_empty_batch_and_empty_file(golden):
    dict_bytes = golden.bytes(golden.dict_name("single_packed_dint", "docs"))
    vf = driver.VroomFile("single_packed_dint", b"", dict_bytes)
    assert vf.n_ints == 0 and vf.decode().numel() == 0


@pytest.mark.parametrize("type_", g.TYPES)
def test_dict_copy_device_tables(golden, oracle, type_):
    """Dictionary::copy through the device-resident tables"""
    kind = KINDS[type_]
    data = golden.bytes(golden.dict_name(type_, "freqs"))
    d = capi.Dictionary(kind, data, device=0)
    ho = oracle.dict_parse(kind, data)
    rng = np.random.default_rng(2)
    for cls in range(5 if kind == 2 else 1):
        hi = 120 if kind == 2 else 65536
        for i in [2, 3, 6, 7, 8] + rng.integers(7, hi, 40).tolist():
            n, row = d.copy(cls, i)
            wn, wrow = oracle.dict_copy(ho, cls, i)
            assert n == wn and np.array_equal(row[:min(n, 16)], wrow[:min(n, 16)])


@pytest.mark.parametrize("type_", g.TYPES)
def test_decode_list_host_is_decoder_decode(golden, oracle, type_):
    """the per-list drop-in: values AND the advanced input pointer"""
    bin_ = golden.bytes(golden.vroom_name(type_, "docs"))
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    d = capi.Dictionary(KINDS[type_], dict_bytes, device=0)
    plan = capi.VroomPlan(d, bin_)
    ho = oracle.dict_parse(KINDS[type_], dict_bytes)
    raw = np.frombuffer(bin_, dtype=np.uint8)
    lens = np.diff(plan.list_offsets.astype(np.int64))
    rng = np.random.default_rng(4)
    for i in sorted(set(rng.integers(0, plan.n_lists, 12).tolist()) | set(np.argsort(-lens)[:3].tolist())):
        n = int(lens[i])
        start = int(plan.list_byte_offsets[i])
        hdr = np.zeros(2, dtype=np.uint32)
        start += oracle.lib.oracle_vbyte_decode(raw[start:].ctypes.data, hdr.ctypes.data, 2)
        vals, used = capi.decode_list_host(d, bin_[start:], n)
        want, wused = oracle.vroom_decode_list(ho, np.frombuffer(bin_[start:] + b"\0" * 64, dtype=np.uint8), n)
        assert used == wused == int(plan.list_byte_offsets[i + 1]) - start
        assert np.array_equal(vals, want)


def test_checksums_and_replicated_collection(golden, vroom_files):
    """size-independent property for large inputs: decoding K back-to-back copies of the file
    yields K identical per-list checksum vectors (checksum of checksums)"""
    vf = vroom_files("single_packed_dint", "docs")
    out = vf.decode()
    sums = vf.checksums(out).cpu().numpy().view(np.uint64)
    want = g.checksum_lists(_np(out), np.diff(vf.plan.list_offsets.astype(np.int64)))
    assert np.array_equal(sums, want)
    K = 3
    bin_ = golden.bytes(golden.vroom_name("single_packed_dint", "docs"))
    big = driver.VroomFile("single_packed_dint", bin_ * K, golden.bytes(golden.dict_name("single_packed_dint", "docs")))
    bo = big.decode()
    bs = big.checksums(bo).cpu().numpy().view(np.uint64).reshape(K, -1)
    assert all(np.array_equal(bs[k], want) for k in range(K))
    assert np.array_equal(_np(bo).reshape(K, -1)[K - 1], _np(out))


@pytest.mark.parametrize("sum_known", (True, False))
def test_interpolative_tail_blocks_random(golden, sum_known):
    """partial blocks through the index path: hand-built m_lists with one tail block per list"""
    rng = np.random.default_rng(13)
    kind = 1
    idxbin = golden.bytes(golden.index_name("single_packed_dint"))
    plan0 = capi.IndexPlan(kind, idxbin)
    dd = capi.Dictionary(kind, idxbin[plan0.docs_dict_off:plan0.freqs_dict_off], device=0, mapped=True)
    fd = capi.Dictionary(kind, idxbin[plan0.freqs_dict_off:plan0.end_off], device=0, mapped=True)
    lists = bytearray()
    blocks, want_d, want_f = [], [], []
    pos = 0
    for n in [1, 2, 3, 16, 17, 31, 100, 254, 255] * 3:
        for mag in (1, 2, 100, 100000):
            gaps = rng.integers(0, mag, n).astype(np.uint32)
            fm1 = rng.integers(0, mag, n).astype(np.uint32)
            base = int(rng.integers(0, 1000))
            docs = base + np.cumsum(gaps.astype(np.uint64) + 1) - 1
            db, s = sb.encode_interpolative_block(gaps, True)
            fb, _ = sb.encode_interpolative_block(fm1, False)
            blocks.append((len(lists), pos, n, base, int(docs[-1]), 0))
            lists += db + fb
            want_d.append(docs.astype(np.uint32))
            want_f.append(fm1 + 1)
            pos += n
    barr = np.array(blocks, dtype=capi.BLOCK_DTYPE)
    d_lists = torch.frombuffer(bytearray(lists), dtype=torch.uint8).cuda()
    d_blocks = torch.from_numpy(barr.view(np.uint8).reshape(-1)).cuda()
    d_docs = torch.zeros(pos, dtype=torch.int32, device="cuda")
    d_freqs = torch.zeros(pos, dtype=torch.int32, device="cuda")
    capi.decode_index(dd, fd, d_lists.data_ptr(), len(lists), d_blocks.data_ptr(), len(barr), 256, d_docs.data_ptr(),
                      d_freqs.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(_np(d_docs), np.concatenate(want_d))
    assert np.array_equal(_np(d_freqs), np.concatenate(want_f))


@pytest.mark.parametrize("type_", g.TYPES)
def test_cpp_host_driver(golden, tmp_path, type_):
    """dint_b200_decode (C++ over the C ABI, mirror of vroom_env/decode.cpp): stats line + per-list
    checksums of the fused docID decode must match the collection itself"""
    import json
    import os
    import subprocess
    from dint_b200 import _build
    _build.build()
    binp = golden.path(golden.vroom_name(type_, "docs"))
    dictp = golden.path(golden.dict_name(type_, "docs"))
    sums_file = str(tmp_path / "sums.bin")
    out = subprocess.run([_build.DRIVER, type_, binp, "--dict", dictp, "--mode", "docids", "--repeat", "3",
                          "--sums", sums_file], check=True, capture_output=True, text=True).stdout
    st = json.loads(out.strip().splitlines()[-1])
    assert int(st["num_sequences"]) == golden.n_lists and int(st["num_integers"]) == golden.n_postings
    assert st["type"] == type_ and int(st["ints_x_sec"]) > 0 and st["kernel_launches"] >= 4
    assert st["plan"] == "device discovery" and st["discover_device_ms"] > 0  # no host walk of the codewords
    out2 = subprocess.run([_build.DRIVER, type_, binp, "--dict", dictp, "--mode", "docids", "--repeat", "1", "--host-plan"],
                          check=True, capture_output=True, text=True).stdout
    st2 = json.loads(out2.strip().splitlines()[-1])
    assert st2["plan"] == "host walk" and st2["checksum"] == st["checksum"] and st2["num_sequences"] == st["num_sequences"]
    vf = driver.VroomFile(type_, golden.bytes(golden.vroom_name(type_, "docs")), golden.bytes(golden.dict_name(type_, "docs")))
    docs = _np(vf.decode(mode=capi.MODE_DOCIDS))
    assert g.sha(docs.tobytes()) == golden.sha["docids_sha256"]
    want = g.checksum_lists(docs, np.diff(vf.plan.list_offsets.astype(np.int64)))
    got = np.fromfile(sums_file, dtype=np.uint64)
    assert np.array_equal(got, want)
    assert int(st["checksum"]) == int(want.sum(dtype=np.uint64))
    # unknown type: logged, exit code 0, like the reference (decode.cpp:259)
    r = subprocess.run([_build.DRIVER, "nope", binp, "--dict", dictp], capture_output=True, text=True)
    assert r.returncode == 0 and "unknown type" in r.stderr


def test_sharded_vroom_file_single_process(golden):
    """term-range shards decoded one after the other on one GPU add up to the whole file"""
    type_ = "multi_packed_dint"
    bin_ = golden.bytes(golden.vroom_name(type_, "docs"))
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    whole = driver.VroomFile(type_, bin_, dict_bytes)
    full = _np(whole.decode(mode=capi.MODE_DOCIDS))
    parts, count, check = [], 0, 0
    for r in range(3):
        sh = driver.ShardedVroomFile(type_, bin_, dict_bytes, r, 3)
        out = sh.decode(mode=capi.MODE_DOCIDS)
        parts.append(_np(out))
        c, k = sh.totals(out)
        count += c
        check = (check + k) & 0x7FFFFFFFFFFFFFFF
    assert np.array_equal(np.concatenate(parts), full) and count == golden.n_postings
    want = int(whole.checksums(whole.decode(mode=capi.MODE_DOCIDS)).cpu().numpy().view(np.uint64).sum(dtype=np.uint64))
    assert check == want & 0x7FFFFFFFFFFFFFFF
