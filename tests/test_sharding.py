"""Multi-GPU host logic on CPU: term-range sharding balanced by postings + the job's only
collective (a SUM all-reduce of {count, checksum}) over gloo with world_size 2 and 3.
Each rank decodes its shard with the ORACLE here (no GPU in this suite); on the GPU box the same
shards go through libdint_b200.so (bench.py / driver.ShardedVroomFile)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_data as g
from dint_b200 import capi, shard
from oracle_lib import KINDS, Oracle


def test_split_by_postings_properties():
    rng = np.random.default_rng(3)
    for n, world in ((1, 1), (1, 4), (10, 3), (1000, 8), (113306, 8)):
        lengths = np.maximum(1, (50000 / rng.zipf(1.3, n)).astype(np.int64))
        b = shard.split_by_postings(lengths, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        per = np.array([lengths[b[r]:b[r + 1]].sum() for r in range(world)])
        assert per.sum() == lengths.sum()
        if n >= 100 * world:
            assert per.max() <= lengths.sum() / world + lengths.max()  # off by at most one list


def _worker(rank, world, port, type_, part, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = g.Golden()
    O = Oracle()
    bin_ = G.bytes(G.vroom_name(type_, part))
    dict_bytes = G.bytes(G.dict_name(type_, part))
    d = capi.Dictionary(KINDS[type_], dict_bytes)  # host tables only
    plan = capi.VroomPlan(d, bin_)
    lengths = np.diff(plan.list_offsets.astype(np.int64))
    mine, first, last = shard.shard_vroom(bin_, lengths, plan.list_byte_offsets, rank, world)
    # the slice is a valid vroom file of its own: plan it, decode it (oracle stands in for the GPU)
    sub = capi.VroomPlan(d, mine)
    assert sub.n_lists == last - first and sub.n_ints == int(lengths[first:last].sum())
    ho = O.dict_parse(KINDS[type_], dict_bytes)
    vals, ln = O.vroom_decode_file(ho, mine, sub.n_ints, sub.n_lists)
    assert np.array_equal(ln.astype(np.int64), lengths[first:last])
    sums = g.checksum_lists(vals, ln.astype(np.int64))
    count, checksum = shard.reduce_totals(sub.n_ints, int(sums.sum(dtype=np.uint64)))
    if rank == 0:
        out.put((count, checksum))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", (2, 3))
@pytest.mark.parametrize("type_,part", (("single_packed_dint", "docs"), ("multi_packed_dint", "freqs")))
def test_sharded_decode_reduces_to_whole_file(golden, oracle, world, type_, part):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world * 7 + (0 if part == "docs" else 1)
    procs = [ctx.Process(target=_worker, args=(r, world, port, type_, part, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    count, checksum = out.get(timeout=10)
    # whole-file truth from the oracle
    bin_ = golden.bytes(golden.vroom_name(type_, part))
    ho = oracle.dict_parse(KINDS[type_], golden.bytes(golden.dict_name(type_, part)))
    vals, ln = oracle.vroom_decode_file(ho, bin_, golden.n_postings, golden.n_lists)
    want = int(g.checksum_lists(vals, ln.astype(np.int64)).sum(dtype=np.uint64)) & 0x7FFFFFFFFFFFFFFF
    assert count == golden.n_postings and checksum == want


def test_virtual_cuts_of_a_replicated_collection():
    """bench.py cuts R back-to-back copies of the sample into term ranges of equal postings without materialising
    R x lists entries: the cuts must be those of shard.split_by_postings on the materialised lengths"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(5)
    lengths = np.maximum(1, (3000 / rng.zipf(1.2, 500)).astype(np.int64))
    lo = np.concatenate([[0], np.cumsum(lengths)])
    n_ints, n_lists = int(lo[-1]), len(lengths)
    for R, world in ((8, 1), (8, 2), (8, 8), (7, 3), (5, 4), (640, 8)):
        cuts = bench.virtual_cuts(lo, n_ints, R, world)
        assert cuts[0] == (0, 0) and cuts[-1] == (R, 0) and len(cuts) == world + 1
        flat = [q * n_lists + j for q, j in cuts]
        assert all(b >= a for a, b in zip(flat, flat[1:]))
        per = [(q1 * n_ints + lo[j1]) - (q0 * n_ints + lo[j0]) for (q0, j0), (q1, j1) in zip(cuts, cuts[1:])]
        assert sum(per) == R * n_ints
        assert max(per) - R * n_ints / world <= lengths.max()  # off by at most one list
        if R <= 8:
            b = shard.split_by_postings(np.tile(lengths, R), world)
            assert [int(x) for x in b] == flat
        if R % world == 0:  # whole replicas per rank: exactly equal shards
            assert len(set(per)) == 1 and all(j == 0 for _, j in cuts)
