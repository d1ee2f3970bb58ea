"""Device-side discovery of the decode restart points (dint_vroom_discover / dint_index_discover):
files decoded with NO host plan must give what the reference decoders give, and the tables the GPU finds
must agree with the serial host walk (dint_vroom_plan_create) on everything that is comparable
(lists, integers, list offsets in the output and in the file)."""
import ctypes as C

import numpy as np
import pytest
import torch

import golden_data as g
import stream_builder as sb
from dint_b200 import capi, driver
from oracle_lib import KINDS
from test_gpu_parity import EDGE_LISTS, _random_lists

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy().view(np.uint32)


# sync_stride variants: library default / every 16 KB / no speculative search at all (one chain)
STRIDES = (0, 16384, capi.NO_SYNC_SEARCH)


@pytest.mark.parametrize("stride", STRIDES)
@pytest.mark.parametrize("type_", g.TYPES)
@pytest.mark.parametrize("part,raw_key,fused_key,mode", (("docs", "vroom_docs_gaps_sha256", "docids_sha256", capi.MODE_DOCIDS),
                                                         ("freqs", "vroom_freqs_minus1_sha256", "freqs_sha256", capi.MODE_FREQS)))
def test_vroom_without_host_plan(golden, type_, part, raw_key, fused_key, mode, stride):
    """config 1 decoded from a file no host code has walked: sha256 of the reference decoder's output"""
    bin_ = golden.bytes(golden.vroom_name(type_, part))
    dict_bytes = golden.bytes(golden.dict_name(type_, part))
    vf = driver.VroomFile(type_, bin_, dict_bytes, plan="device", sync_stride=stride)
    assert vf.plan is None and vf.n_lists == golden.n_lists and vf.n_ints == golden.n_postings
    assert g.sha(_np(vf.decode(mode=capi.MODE_RAW)).tobytes()) == golden.sha[raw_key]
    assert g.sha(_np(vf.decode(mode=mode)).tobytes()) == golden.sha[fused_key]
    # list offsets (output and file) == the host walk's
    host = capi.VroomPlan(capi.Dictionary(KINDS[type_], dict_bytes), bin_)
    assert np.array_equal(vf.d_list_offsets.cpu().numpy().astype(np.uint64), host.list_offsets)
    assert np.array_equal(vf.d_list_byte_offsets.cpu().numpy().astype(np.uint64), host.list_byte_offsets)
    if stride == capi.NO_SYNC_SEARCH:  # one chain: a warp skims it and cuts it into pieces for the thread walker
        assert vf.dplan.n_sync_found == 0 and vf.dplan.n_skimmed == 1 and vf.dplan.n_chains > 1
    elif stride:
        assert vf.dplan.n_chains > 8  # the speculative search found list starts to cut the file at
    # a speculative sync point may be wrong (it then costs a repair round of a few chains), never the result
    assert vf.dplan.retries <= 3


@pytest.mark.parametrize("type_", g.TYPES)
def test_sync_hints_right_and_wrong(golden, type_):
    """caller-provided list starts are used; offsets that are NOT list starts are dropped (the previous chain does
    not land on them) and the result is unchanged"""
    bin_ = golden.bytes(golden.vroom_name(type_, "docs"))
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    host = capi.VroomPlan(capi.Dictionary(KINDS[type_], dict_bytes), bin_)
    starts = host.list_byte_offsets
    good = starts[[1000, 5000, 50000, 100000]]
    vf = driver.VroomFile(type_, bin_, dict_bytes, plan="device", sync_offsets=good, sync_stride=capi.NO_SYNC_SEARCH)
    assert vf.dplan.n_sync_found == 0 and vf.dplan.n_chains >= 5 and vf.dplan.retries == 0
    assert g.sha(_np(vf.decode(mode=capi.MODE_DOCIDS)).tobytes()) == golden.sha["docids_sha256"]
    bad = np.array([starts[1000], starts[5000] + 1, starts[50000] + 3, starts[100000], len(bin_) - 2], dtype=np.uint64)
    vf = driver.VroomFile(type_, bin_, dict_bytes, plan="device", sync_offsets=bad, sync_stride=capi.NO_SYNC_SEARCH)
    assert vf.dplan.retries >= 1
    assert vf.n_lists == golden.n_lists and vf.n_ints == golden.n_postings
    assert g.sha(_np(vf.decode(mode=capi.MODE_DOCIDS)).tobytes()) == golden.sha["docids_sha256"]
    assert np.array_equal(vf.d_list_offsets.cpu().numpy().astype(np.uint64), host.list_offsets)


@pytest.mark.parametrize("type_", g.TYPES)
def test_discover_edge_streams(golden, oracle, type_):
    """32-bit exceptions, payload equal to 0/1, exceptions straddling tiles, runs, 1-integer lists, u8 blocks,
    a 60 000-integer list: device-found tables decode to what was encoded, in every mode"""
    rng = np.random.default_rng(11)
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    di = sb.DictImage(capi.KIND_OF_TYPE[type_], dict_bytes)
    lists = EDGE_LISTS + _random_lists(di, rng, 150, 60000)
    want = np.concatenate([np.asarray(v, dtype=np.uint32) for v in lists])
    exp_docs = np.concatenate([(np.cumsum(np.asarray(v, dtype=np.uint64) + 1) - 1).astype(np.uint32) for v in lists])
    ho = oracle.dict_parse(KINDS[type_], dict_bytes)
    for sels in ([0, 6, 3, 9, 1, 7], [8], [2], rng.integers(0, 12, 64).tolist()):
        data = b"".join(sb.encode_vroom_list(di, v, block_size=256, selectors=sels) for v in lists)
        o, _ = oracle.vroom_decode_file(ho, data, len(want), len(lists))
        assert np.array_equal(o, want)
        for stride in (4096, capi.NO_SYNC_SEARCH):
            vf = driver.VroomFile(type_, data, dict_bytes, plan="device", sync_stride=stride)
            assert vf.n_lists == len(lists) and vf.n_ints == len(want)
            assert np.array_equal(_np(vf.decode(mode=capi.MODE_RAW)), want), (sels[:6], stride)
            assert np.array_equal(_np(vf.decode(mode=capi.MODE_DOCIDS)), exp_docs), (sels[:6], stride)
            assert np.array_equal(_np(vf.decode(mode=capi.MODE_FREQS)), want + np.uint32(1)), (sels[:6], stride)
    oracle.dict_free(ho)


@pytest.mark.parametrize("knobs", ((2048, 1024), (4096, 300), (1024, 0), (1 << 30, 1024)))
@pytest.mark.parametrize("type_", g.TYPES)
def test_skim_pass_cuts_anywhere(golden, type_, knobs, monkeypatch):
    """the skim pass (long chains -> pieces that start at block boundaries inside a list) with the thresholds turned
    down so that nearly every chain is cut every few blocks / with the skim switched off (warps walk the long chains) /
    with nothing long enough to skim: the tables decode to the reference's output and the list table is the host walk's"""
    long_chain, cut = knobs
    monkeypatch.setenv("DINT_B200_LONG_CHAIN", str(long_chain))
    monkeypatch.setenv("DINT_B200_CUT_BYTES", str(cut))
    for part, fused_key, mode in (("docs", "docids_sha256", capi.MODE_DOCIDS), ("freqs", "freqs_sha256", capi.MODE_FREQS)):
        bin_ = golden.bytes(golden.vroom_name(type_, part))
        dict_bytes = golden.bytes(golden.dict_name(type_, part))
        host = capi.VroomPlan(capi.Dictionary(KINDS[type_], dict_bytes), bin_)
        for stride in (0, capi.NO_SYNC_SEARCH):
            vf = driver.VroomFile(type_, bin_, dict_bytes, plan="device", sync_stride=stride)
            assert vf.n_lists == golden.n_lists and vf.n_ints == golden.n_postings
            assert g.sha(_np(vf.decode(mode=mode)).tobytes()) == golden.sha[fused_key], (part, stride)
            assert np.array_equal(vf.d_list_offsets.cpu().numpy().astype(np.uint64), host.list_offsets)
            assert np.array_equal(vf.d_list_byte_offsets.cpu().numpy().astype(np.uint64), host.list_byte_offsets)
            if cut == 0 or long_chain == 1 << 30:
                assert vf.dplan.n_skimmed == 0
            else:
                assert vf.dplan.n_skimmed >= 1
                if type_ == "multi_packed_dint" and stride == capi.NO_SYNC_SEARCH:
                    assert vf.dplan.n_chains > len(bin_) // (4 * max(cut, 1024))  # cut inside the lists as well


def test_skim_pass_malformed_long_list(golden):
    """a long multi_packed list with a broken selector byte in the middle: the skim stops cutting there, the walk of
    the last piece reports the file as malformed"""
    type_ = "multi_packed_dint"
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    di = sb.DictImage(capi.KIND_OF_TYPE[type_], dict_bytes)
    rng = np.random.default_rng(5)
    lists = _random_lists(di, rng, 3, 200000)
    data = bytearray(b"".join(sb.encode_vroom_list(di, v, block_size=256, selectors=[0, 6, 3]) for v in lists))
    vf = driver.VroomFile(type_, bytes(data), dict_bytes, plan="device", sync_stride=capi.NO_SYNC_SEARCH)
    assert vf.n_ints == sum(len(v) for v in lists) and vf.dplan.n_skimmed == 1
    host = capi.VroomPlan(capi.Dictionary(KINDS[type_], dict_bytes), bytes(data))
    # the selector byte of a block in the middle of the longest list
    k = int(np.argmax([len(v) for v in lists]))
    segs = host.segments
    lo, hi = int(host.list_offsets[k]), int(host.list_offsets[k + 1])
    mine = segs[(segs["out_off"] >= lo) & (segs["out_off"] < hi)]
    assert len(mine) > 8
    data[int(mine[len(mine) // 2]["in_off"])] = 0xFF
    with pytest.raises(capi.DintError) as e:
        driver.VroomFile(type_, bytes(data), dict_bytes, plan="device", sync_stride=capi.NO_SYNC_SEARCH)
    assert e.value.code == -2


def test_discover_rejects_malformed_files(golden):
    type_ = "single_packed_dint"
    bin_ = golden.bytes(golden.vroom_name(type_, "docs"))
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    with pytest.raises(capi.DintError) as e:
        driver.VroomFile(type_, bin_[:len(bin_) - 5], dict_bytes, plan="device", sync_stride=capi.NO_SYNC_SEARCH)
    assert e.value.code == -2
    # a header that announces more integers than the rest of the file can hold (a byte encodes at most 256): refused at
    # the header, not after walking to the end of the file
    host = capi.VroomPlan(capi.Dictionary(KINDS[type_], dict_bytes), bin_)
    broken = bytearray(bin_)
    at = int(host.list_byte_offsets[len(host.list_byte_offsets) // 2])
    broken[at:at + 5] = bytes([0x7F, 0x7F, 0x7F, 0x7F, 0x8F])
    for stride in (0, capi.NO_SYNC_SEARCH):
        with pytest.raises(capi.DintError) as e:
            driver.VroomFile(type_, bytes(broken), dict_bytes, plan="device", sync_stride=stride)
        assert e.value.code == -2
    # an empty file is fine: nothing to decode
    vf = driver.VroomFile(type_, b"", dict_bytes, plan="device")
    assert vf.n_lists == 0 and vf.n_ints == 0


@pytest.mark.parametrize("type_", g.TYPES)
def test_cold_host_call_equals_planned_call(golden, type_):
    """dint_vroom_ctx_create_discover (H2D + device-side discovery, no host plan) then decode + checksums:
    the per-list checksums are those of the host-planned path and of the collection itself"""
    bin_ = golden.bytes(golden.vroom_name(type_, "docs"))
    dict_bytes = golden.bytes(golden.dict_name(type_, "docs"))
    d = capi.Dictionary(KINDS[type_], dict_bytes, device=0)
    h_bin = torch.frombuffer(bytearray(bin_), dtype=torch.uint8).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    dec = capi.VroomHostDecoder.discover(d, h_bin.data_ptr(), h_bin.numel(), stream=st)
    assert dec.n_lists == golden.n_lists and dec.n_ints == golden.n_postings
    dec.decode_resident(mode=capi.MODE_DOCIDS, stream=st)
    sums = torch.empty(dec.n_lists, dtype=torch.int64).pin_memory()
    out = torch.empty(dec.n_ints, dtype=torch.int32).pin_memory()
    dec.results(h_out_ptr=out.data_ptr(), h_sums_ptr=sums.data_ptr(), stream=st)
    dec.sync(st)
    assert g.sha(out.numpy().view(np.uint32).tobytes()) == golden.sha["docids_sha256"]
    plan = capi.VroomPlan(d, bin_)
    ref = capi.VroomHostDecoder(plan)
    sums2 = torch.empty(dec.n_lists, dtype=torch.int64).pin_memory()
    ref.decode(h_bin.data_ptr(), h_bin.numel(), mode=capi.MODE_DOCIDS, h_sums_ptr=sums2.data_ptr(), stream=st)
    ref.sync(st)
    assert torch.equal(sums, sums2)
    lens = np.diff(plan.list_offsets.astype(np.int64))
    want = g.checksum_lists(out.numpy().view(np.uint32), lens)
    assert np.array_equal(sums.numpy().view(np.uint64), want)


@pytest.mark.parametrize("type_", g.TYPES)
def test_index_scan_with_discovered_strips(golden, oracle, type_):
    """config 2 through the lane-sequential decoder: the full blocks' strip tables come from dint_index_discover
    (block starts from the file's own endpoints, codeword boundaries found on the device)"""
    idx = driver.IndexFile(type_, golden.bytes(golden.index_name(type_)))
    d, f = idx.scan(engine="strips")
    d, f = _np(d), _np(f)
    assert g.sha(d.tobytes()) == golden.sha["docids_sha256"]
    assert g.sha(f.tobytes()) == golden.sha["freqs_sha256"]
    assert idx.docs_plan.n_units > 0 and idx.freqs_plan.n_units > 0
    d2, _ = idx.scan(engine="strips", want_freqs=False)
    _, f2 = idx.scan(engine="strips", want_docs=False)
    assert np.array_equal(_np(d2), d) and np.array_equal(_np(f2), f)


@pytest.mark.parametrize("type_", ("single_rect_dint", "single_packed_dint"))
def test_index_discovered_strips_block128(golden128, type_):
    data = golden128.bytes(golden128.index_name(type_))
    idx = driver.IndexFile(type_, data, block_size=128)
    d, f = idx.scan(engine="strips")
    assert g.sha(_np(d).tobytes()) == golden128.sha["docids_sha256"]
    assert g.sha(_np(f).tobytes()) == golden128.sha["freqs_sha256"]


def test_index_discover_subset_of_blocks(golden):
    """a ragged block table (every third list): units must not carry a docID prefix across a gap"""
    type_ = "single_packed_dint"
    idx = driver.IndexFile(type_, golden.bytes(golden.index_name(type_)))
    full_d, full_f = idx.scan()
    full_d, full_f = _np(full_d), _np(full_f)
    p = idx.plan
    blocks = p.blocks
    list_of = np.searchsorted(p.list_offsets, blocks["out_off"], side="right") - 1
    long_lists = np.unique(list_of[blocks["size"] == 256])
    pick = blocks[np.isin(list_of, long_lists[::3])].copy()
    # drop every 5th block as well, and pack the outputs
    pick = pick[np.arange(len(pick)) % 5 != 2].copy()
    src = pick["out_off"].astype(np.int64)
    n = pick["size"].astype(np.int64)
    new_off = np.concatenate([[0], np.cumsum(n)])
    pick["out_off"] = new_off[:-1]
    d_blocks = torch.from_numpy(pick.view(np.uint8).reshape(-1)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    dp, fp = capi.DevicePlan.discover_index(idx.docs_dict, idx.freqs_dict, idx.d_bin.data_ptr() + p.lists_off, p.lists_len,
                                            d_blocks.data_ptr(), len(pick), 256, stream=st)
    dd = torch.zeros(int(new_off[-1]), dtype=torch.int32, device="cuda")
    ff = torch.zeros_like(dd)
    capi.decode_index_planned(idx.docs_dict, idx.freqs_dict, idx.d_bin.data_ptr() + p.lists_off, p.lists_len, dp, fp,
                              d_blocks.data_ptr(), len(pick), 256, dd.data_ptr(), ff.data_ptr(), stream=st)
    want_d = np.concatenate([full_d[s:s + k] for s, k in zip(src, n)])
    want_f = np.concatenate([full_f[s:s + k] for s, k in zip(src, n)])
    assert np.array_equal(_np(dd), want_d) and np.array_equal(_np(ff), want_f)


def test_full_size_multi_dictionary_every_engine(oracle):
    """a multi_packed dictionary at its maximum size (6 classes x 65536 + 7 entries: a 192 KB 4-bit length table)
    must decode through every engine: the tile decoder shrinks its CTA until table + rings fit in shared memory
    (it used to refuse), the strip decoder and the device-side discovery use the group tables"""
    from test_gpu_parity import _synthetic_packed_dictionary, _decode_lists
    type_ = "multi_packed_dint"
    dict_bytes = _synthetic_packed_dictionary(6, 65543, 7)
    di = sb.DictImage(capi.KIND_OF_TYPE[type_], dict_bytes)
    rng = np.random.default_rng(8)
    lists = []
    for li in range(40):
        vals = []
        n = int(rng.choice([3, 60, 300, 1500]))
        cls = int(rng.integers(0, 6))
        ents = [t for t in di.classes[cls][7:65536] if t is not None]
        while len(vals) < n:
            if rng.random() < 0.8:
                vals.extend(ents[int(rng.integers(0, len(ents)))])
            else:
                vals.extend(rng.integers(0, 1 << 20, 2).tolist())
        lists.append(vals[:n])
    want = np.concatenate([np.asarray(v, dtype=np.uint32) for v in lists])
    sels = [0, 1, 2, 3, 4, 5, 7, 9]
    ho = oracle.dict_parse(KINDS[type_], dict_bytes)
    for engine in ("strips", "units"):
        out, data, vf = _decode_lists(type_, dict_bytes, lists, engine=engine, selectors=sels)
        o, _ = oracle.vroom_decode_file(ho, data, len(want), len(lists))
        assert np.array_equal(o, want)
        assert np.array_equal(out, want), engine
    vd = driver.VroomFile(type_, data, dict_bytes, plan="device", sync_stride=2048)
    assert np.array_equal(_np(vd.decode(mode=capi.MODE_RAW)), want)
    # the literal Decoder::decode drop-in (one list, host buffers) with the same dictionary: its scratch is reused
    d = capi.Dictionary(2, dict_bytes, device=0)
    host_plan = capi.VroomPlan(d, data)
    lbo = host_plan.list_byte_offsets.astype(np.int64)
    for li in (0, 5, 39):
        hdr = len(sb.vbyte(len(lists[li]))) + len(sb.vbyte(sum(lists[li]) & 0xFFFFFFFF))
        vals, used = capi.decode_list_host(d, data[lbo[li] + hdr:], len(lists[li]))
        assert np.array_equal(vals, np.asarray(lists[li], dtype=np.uint32)) and used == lbo[li + 1] - lbo[li] - hdr
    oracle.dict_free(ho)


def test_index_block_size_out_of_range_is_rejected(golden):
    idx = driver.IndexFile("single_packed_dint", golden.bytes(golden.index_name("single_packed_dint")))
    p = idx.plan
    out = torch.zeros(idx.n_postings, dtype=torch.int32, device="cuda")
    with pytest.raises(capi.DintError) as e:
        capi.decode_index(idx.docs_dict, idx.freqs_dict, idx.d_bin.data_ptr() + p.lists_off, p.lists_len,
                          idx.d_blocks.data_ptr(), int(p.n_blocks), 2048, out.data_ptr(), None)
    assert e.value.code == -1


@pytest.mark.parametrize("type_", g.TYPES)
def test_index_block_table_built_on_the_device(golden, type_):
    """dint_index_device_plan_create: Elias-Fano endpoints, list offsets and the block table found by the GPU are
    exactly the host pre-pass's, and the scan through them gives the reference enumerator's output"""
    data = golden.bytes(golden.index_name(type_))
    host = capi.IndexPlan(capi.KIND_OF_TYPE[type_], data)
    idx = driver.IndexFile(type_, data, plan="device")
    p = idx.plan
    assert (p.n_lists, p.n_postings, p.n_blocks, p.num_docs) == (host.n_lists, host.n_postings, host.n_blocks, host.num_docs)
    assert (p.lists_off, p.lists_len, p.docs_dict_off, p.freqs_dict_off, p.end_off) == \
        (host.lists_off, host.lists_len, host.docs_dict_off, host.freqs_dict_off, host.end_off)
    ep = driver.device_view(p.endpoints_ptr, p.n_lists + 1, "<i8", idx.dev).cpu().numpy().astype(np.uint64)
    assert np.array_equal(ep, host.endpoints)
    assert np.array_equal(idx.d_list_offsets.cpu().numpy().astype(np.uint64), host.list_offsets)
    blocks = idx.d_blocks.cpu().numpy().view(capi.BLOCK_DTYPE)
    assert np.array_equal(blocks, host.blocks)
    for engine in ("strips", "tiles"):
        d, f = idx.scan(engine=engine)
        assert g.sha(_np(d).tobytes()) == golden.sha["docids_sha256"], engine
        assert g.sha(_np(f).tobytes()) == golden.sha["freqs_sha256"], engine


def test_index_device_plan_block128_and_corrupt_files(golden, golden128):
    data = golden128.bytes(golden128.index_name("single_packed_dint"))
    idx = driver.IndexFile("single_packed_dint", data, block_size=128, plan="device")
    d, f = idx.scan(engine="strips")
    assert g.sha(_np(d).tobytes()) == golden128.sha["docids_sha256"] and g.sha(_np(f).tobytes()) == golden128.sha["freqs_sha256"]
    good = golden.bytes(golden.index_name("single_packed_dint"))
    d_bin = torch.frombuffer(bytearray(good), dtype=torch.uint8).cuda()
    for bad in (good[:len(good) - 8], good[:40]):
        with pytest.raises(capi.DintError):
            capi.DeviceIndexPlan(1, bad, d_bin.data_ptr())
    # a corrupt list header inside m_lists is found by the device walk
    host = capi.IndexPlan(1, good)
    broken = bytearray(good)
    pos = host.lists_off + int(host.endpoints[5])
    broken[pos:pos + 5] = b"\x00\x00\x00\x00\x00"  # vbyte without a terminator
    d_broken = torch.frombuffer(bytearray(broken), dtype=torch.uint8).cuda()
    with pytest.raises(capi.DintError) as e:
        capi.DeviceIndexPlan(1, bytes(broken), d_broken.data_ptr())
    assert e.value.code == -2


@pytest.mark.parametrize("plan", ("host", "device"))
@pytest.mark.parametrize("type_", ("block_interpolative", "block_vbyte", "block_u32"))
def test_block_codec_index_scan(golden, oracle, type_, plan):
    """frozen block_freq_index<interpolative_block | vbyte_block | u32_block> files (reference-built)
    decoded through the same block-table ABI (dint_decode_index_blocks): == collection == oracle"""
    data = golden.bytes(f"index.{type_}.bin")
    idx = driver.IndexFile(type_, data, plan=plan)
    assert idx.docs_dict is None and idx.n_postings == golden.n_postings
    d, f = idx.scan()
    d, f = _np(d), _np(f)
    assert g.sha(d.tobytes()) == golden.sha["docids_sha256"]
    assert g.sha(f.tobytes()) == golden.sha["freqs_sha256"]
    wd, wf, _ = oracle.index_decode(capi.KIND_OF_TYPE[type_], data, golden.n_postings + 16)
    assert np.array_equal(d, wd) and np.array_equal(f, wf)
    d2, _ = idx.scan(want_freqs=False)
    _, f2 = idx.scan(want_docs=False)
    assert np.array_equal(_np(d2), wd) and np.array_equal(_np(f2), wf)
