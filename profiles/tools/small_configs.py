#!/usr/bin/env python3
"""Latency of the small configurations of SURVEY §8(d) (BASELINE.json configs[0..2]) on one GPU, through the
C ABI: config 1 = vroom decode of test_collection, config 2 = full scan of the frozen dict_freq_index,
config 3 = ragged batches of the lists touched by the reference's 500-query log. Prints one JSON object.
These are latency-bound (13 MB of output): time = back-to-back launches on one stream, CUDA events."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_data as g
from dint_b200 import capi, driver

G = g.Golden()
dev = torch.device("cuda", 0)

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps  # us

res = {"config1_vroom_test_collection": {}, "config2_index_scan": {}, "config3_query_log": {}}
for type_ in ("single_packed_dint", "multi_packed_dint"):
    for part, mode in (("docs", capi.MODE_DOCIDS), ("freqs", capi.MODE_FREQS)):
        vf = driver.VroomFile(type_, G.bytes(G.vroom_name(type_, part)), G.bytes(G.dict_name(type_, part)))
        out = torch.empty(vf.n_ints, dtype=torch.int32, device=dev)
        us = timed(lambda: vf.decode(out, mode=mode), 500)
        # the same file with NO host plan: restart points found on the device (load time), then the same decoder
        vd = driver.VroomFile(type_, G.bytes(G.vroom_name(type_, part)), G.bytes(G.dict_name(type_, part)), plan="device")
        us_d = timed(lambda: vd.decode(out, mode=mode), 500)
        res["config1_vroom_test_collection"][f"{type_}:{part}"] = {
            "us_per_pass": us, "lists": vf.n_lists, "ints": vf.n_ints, "in_bytes": vf.plan.in_bytes,
            "G_ints_per_s": vf.n_ints / us / 1e3,
            "device_plan": {"us_per_pass": us_d, "discover_ms": vd.dplan.discover_ms, "phases_ms": vd.dplan.phase_ms,
                            "chains": vd.dplan.n_chains, "repair_rounds": vd.dplan.retries}}
for type_ in g.TYPES:
    idx = driver.IndexFile(type_, G.bytes(G.index_name(type_)))
    d = torch.empty(idx.n_postings, dtype=torch.int32, device=dev); f = torch.empty_like(d)
    us = timed(lambda: idx.scan(d, f), 200)
    dp, fp = idx.discover()
    us_s = timed(lambda: idx.scan(d, f, engine="strips"), 200)
    res["config2_index_scan"][type_] = {"us_per_pass_tiles": us, "us_per_pass_strips": us_s, "postings": idx.n_postings,
                                        "block_size": idx.block_size, "lists_bytes": int(idx.plan.lists_len),
                                        "G_ints_per_s_strips": 2 * idx.n_postings / us_s / 1e3,
                                        "discover_ms": dp.discover_ms + fp.discover_ms}
# the other block codecs behind the same block table: block_freq_index<interpolative_block | vbyte_block | u32_block>
REF_CPU = {"block_interpolative": "113 ms per scan (58.7 M ints/s)", "block_vbyte": "36 ms per scan (183 M ints/s)",
           "block_u32": "20 ms per scan (327 M ints/s)"}
for t in ("block_interpolative", "block_vbyte", "block_u32"):
    bi = driver.IndexFile(t, G.bytes(f"index.{t}.bin"), plan="device")
    d = torch.empty(bi.n_postings, dtype=torch.int32, device=dev); f = torch.empty_like(d)
    us_bi = timed(lambda: bi.scan(d, f), 200)
    res[f"{t}_index_scan"] = {"us_per_pass": us_bi, "postings": bi.n_postings, "G_ints_per_s": 2 * bi.n_postings / us_bi / 1e3,
                              "lists_bytes": int(bi.plan.lists_len), "block_table_device_ms": bi.plan.build_ms,
                              "reference_cpu": f"ref_index scan {t}: {REF_CPU[t]}, 1 thread, build container"}
# config 3
type_ = "single_packed_dint"
vf = driver.VroomFile(type_, G.bytes(G.vroom_name(type_, "docs")), G.bytes(G.dict_name(type_, "docs")))
queries = json.load(open(os.path.join(g.GOLD, "queries.json")))["queries"]
segs = vf.plan.segments
list_of = np.searchsorted(vf.plan.list_offsets, segs["out_off"], side="right") - 1
st = torch.cuda.current_stream().cuda_stream

def batch(terms):
    pick = segs[np.isin(list_of, np.asarray(terms))].copy()
    lens = pick["n"].astype(np.int64)
    pick["out_off"] = np.concatenate([[0], np.cumsum(lens)])[:-1]
    return torch.from_numpy(pick.view(np.uint8).reshape(-1)).to(dev), len(pick), int(lens.sum())

all_terms = sorted({t for q in queries for t in q})
d_segs, n_segs, n_post = batch(all_terms)
out = torch.empty(n_post, dtype=torch.int32, device=dev)
us_all = timed(lambda: capi.decode_segments(vf.dict, vf.d_in.data_ptr(), vf.plan.in_bytes, d_segs.data_ptr(), n_segs,
                                            out.data_ptr(), capi.MODE_DOCIDS, stream=st), 500)
per_q = [batch(q) for q in queries]
outq = torch.empty(max(p[2] for p in per_q), dtype=torch.int32, device=dev)
def all_queries():
    for ds, ns, _ in per_q:
        capi.decode_segments(vf.dict, vf.d_in.data_ptr(), vf.plan.in_bytes, ds.data_ptr(), ns, outq.data_ptr(),
                             capi.MODE_DOCIDS, stream=st)
us_q = timed(all_queries, 20) / len(per_q)
res["config3_query_log"] = {"queries": len(queries), "distinct_terms": len(all_terms), "postings_distinct": n_post,
                            "postings_with_repeats": sum(p[2] for p in per_q),
                            "one_batch_all_terms_us": us_all, "one_batch_per_query_us_avg": us_q,
                            "note": "dint_decode_segments (tile decoder) on a segment table filtered to the touched lists; docIDs fused"}
# config 3 answered: batched AND queries (ds2i::and_query) on the device over the decoded lists
full = vf.decode(mode=capi.MODE_DOCIDS)
qoff = torch.from_numpy(np.concatenate([[0], np.cumsum([len(q) for q in queries])]).astype(np.int64)).to(dev)
qterms = torch.tensor([t for q in queries for t in q], dtype=torch.int32, device=dev)
counts = torch.empty(len(queries), dtype=torch.int64, device=dev)
us_and = timed(lambda: capi.and_queries(full.data_ptr(), vf.d_list_offsets.data_ptr(), vf.n_lists, qoff.data_ptr(), qterms.data_ptr(),
                                        len(queries), counts.data_ptr(), stream=st), 500)
want = json.load(open(os.path.join(g.GOLD, "queries.json")))["and_counts"]
assert counts.cpu().tolist() == want
# one query at a time (latency of a single AND query on resident docIDs)
def one_by_one():
    for i in range(0, len(queries), 25):
        capi.and_queries(full.data_ptr(), vf.d_list_offsets.data_ptr(), vf.n_lists, qoff[i:].data_ptr(), qterms.data_ptr(), 1,
                         counts.data_ptr(), stream=st)
us_one = timed(one_by_one, 50) / 20
res["config3_and_queries"] = {"queries": len(queries), "results_total": int(sum(want)), "batch_of_500_us": us_and,
                              "us_per_query_in_batch": us_and / len(queries), "single_query_launch_us": us_one,
                              "decode_touched_lists_plus_and_us": us_all + us_and,
                              "reference_cpu": "and_query<false> on the frozen index: 5.6 ms for the 500 queries (11 us each), 1 thread, build container",
                              "note": "dint_and_queries on docIDs resident in HBM; counts equal the reference's for all 500 queries"}
print(json.dumps(res, indent=1))
