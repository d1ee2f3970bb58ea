"""the consumer of the decode path on the device — batched AND queries (ds2i::and_query,
include/ds2i/queries.hpp:34-86) over the docID lists decoded into HBM. The expected counts come from the
REFERENCE's own and_query run on the frozen indexes of the test collection (tests/golden/queries.json:
and_counts, produced by oracle/make_query_fixture.py through oracle/_ref/ref_index)."""
import json
import os

import numpy as np
import pytest
import torch

import golden_data as g
from dint_b200 import capi, driver

pytestmark = pytest.mark.gpu


def _fixture():
    with open(os.path.join(g.GOLD, "queries.json")) as f:
        d = json.load(f)
    return d["queries"], d["and_counts"]


def _csr(queries):
    off = np.concatenate([[0], np.cumsum([len(q) for q in queries])]).astype(np.int64)
    terms = np.array([t for q in queries for t in q], dtype=np.int64)
    return torch.from_numpy(off).cuda(), torch.from_numpy(terms).to(torch.int32).cuda()


def _run(docids, list_offsets, n_lists, queries):
    qoff, qterms = _csr(queries)
    counts = torch.full((len(queries),), -7, dtype=torch.int64, device="cuda")
    capi.and_queries(docids.data_ptr(), list_offsets.data_ptr(), n_lists, qoff.data_ptr(), qterms.data_ptr(), len(queries),
                     counts.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    return counts.cpu().numpy()


@pytest.mark.parametrize("type_", g.TYPES)
def test_and_queries_equal_reference_counts(golden, type_):
    """all 500 queries of the reference's log, docIDs from the index scan (device-discovered strips)"""
    queries, want = _fixture()
    idx = driver.IndexFile(type_, golden.bytes(golden.index_name(type_)))
    d, _ = idx.scan(engine="strips", want_freqs=False)
    got = _run(d, idx.d_list_offsets, idx.n_lists, queries)
    assert got.tolist() == want
    assert sum(want) == 19940 and sum(1 for c in want if c) == 303  # the fixture itself (reference run)


def test_and_queries_on_vroom_docids_and_numpy(golden):
    """the same from a vroom file decoded with no host plan; cross-checked with numpy intersections"""
    queries, want = _fixture()
    type_ = "multi_packed_dint"
    vf = driver.VroomFile(type_, golden.bytes(golden.vroom_name(type_, "docs")), golden.bytes(golden.dict_name(type_, "docs")),
                          plan="device")
    d = vf.decode(mode=capi.MODE_DOCIDS)
    got = _run(d, vf.d_list_offsets, vf.n_lists, queries)
    assert got.tolist() == want
    h = d.cpu().numpy().view(np.uint32)
    lo = vf.d_list_offsets.cpu().numpy()
    for qi in (0, 2, 17, 123, 499):
        sets = [h[lo[t]:lo[t + 1]] for t in set(queries[qi])]
        inter = sets[0]
        for s_ in sets[1:]:
            inter = np.intersect1d(inter, s_)
        assert len(inter) == want[qi]


def test_and_queries_edge_cases(golden):
    type_ = "single_packed_dint"
    vf = driver.VroomFile(type_, golden.bytes(golden.vroom_name(type_, "docs")), golden.bytes(golden.dict_name(type_, "docs")))
    d = vf.decode(mode=capi.MODE_DOCIDS)
    lo = vf.plan.list_offsets.astype(np.int64)
    lens = np.diff(lo)
    big = int(np.argmax(lens))
    qs = [[big], [big, big, big], [big, vf.n_lists + 5], [], [0, 1, 2, 3] * 9, [big, int(np.argmin(lens))]]
    got = _run(d, vf.d_list_offsets, vf.n_lists, qs)
    assert got[0] == lens[big] and got[1] == lens[big]       # one term / duplicates of it: the whole list
    assert got[2] == 0 and got[3] == 0                       # unknown term id, empty query
    assert got[4] == -1                                      # 36 terms: more than the kernel takes (UINT64_MAX)
    h = d.cpu().numpy().view(np.uint32)
    a, b = h[lo[big]:lo[big + 1]], h[lo[qs[5][1]]:lo[qs[5][1] + 1]]
    assert got[5] == len(np.intersect1d(a, b))


def test_and_queries_on_touched_lists_only(golden):
    """config 3 end to end: decode only the 1 088 lists the queries touch (ragged batch), answer the queries there"""
    queries, want = _fixture()
    type_ = "single_packed_dint"
    vf = driver.VroomFile(type_, golden.bytes(golden.vroom_name(type_, "docs")), golden.bytes(golden.dict_name(type_, "docs")))
    terms = np.array(sorted({t for q in queries for t in q}), dtype=np.int64)
    remap = {int(t): i for i, t in enumerate(terms)}
    segs = vf.plan.segments
    list_of = np.searchsorted(vf.plan.list_offsets, segs["out_off"], side="right") - 1
    pick = segs[np.isin(list_of, terms)].copy()
    lens = pick["n"].astype(np.int64)
    new_off = np.concatenate([[0], np.cumsum(lens)])
    pick["out_off"] = new_off[:-1]
    d_segs = torch.from_numpy(pick.view(np.uint8).reshape(-1)).cuda()
    out = torch.zeros(int(new_off[-1]), dtype=torch.int32, device="cuda")
    capi.decode_segments(vf.dict, vf.d_in.data_ptr(), vf.in_bytes, d_segs.data_ptr(), len(pick), out.data_ptr(),
                         capi.MODE_DOCIDS, stream=torch.cuda.current_stream().cuda_stream)
    full_lo = vf.plan.list_offsets.astype(np.int64)
    sub_lo = np.concatenate([[0], np.cumsum(full_lo[terms + 1] - full_lo[terms])])
    d_lo = torch.from_numpy(sub_lo).cuda()
    got = _run(out, d_lo, len(terms), [[remap[t] for t in q] for q in queries])
    assert got.tolist() == want
