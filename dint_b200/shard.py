"""Term-range sharding of a vroom-encoded collection across the GPUs of one box.

Posting lists are independent units, so the multi-GPU path has no exchange step:
rank r takes a contiguous range of term ids holding ~1/world of the POSTINGS (not of the lists —
list lengths are Zipfian), slices the .bin at list boundaries, plans and decodes its slice with its
own replica of the dictionary, and only the final {decoded count, checksum} pair is all-reduced
(NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def split_by_postings(list_lengths, world):
    """boundaries[r] .. boundaries[r+1] = the term-id range of rank r (world + 1 entries).
    Contiguous, covers every list exactly once, balanced by postings."""
    lengths = np.asarray(list_lengths, dtype=np.int64)
    n = len(lengths)
    if world < 1:
        raise ValueError("world must be >= 1")
    csum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(csum[-1])
    targets = (np.arange(1, world, dtype=np.float64) * total / world)
    cuts = np.searchsorted(csum, targets, side="left")
    # snap each cut to the nearer list boundary
    cuts = np.clip(cuts, 0, n)
    for i, (c, t) in enumerate(zip(cuts, targets)):
        if c > 0 and abs(csum[c - 1] - t) <= abs(csum[min(c, n)] - t):
            cuts[i] = c - 1
    b = np.concatenate([[0], np.maximum.accumulate(cuts), [n]]).astype(np.int64)
    return b


def shard_byte_range(list_byte_offsets, boundaries, rank):
    """byte range [lo, hi) of rank's lists inside the .bin (list headers included)"""
    lbo = np.asarray(list_byte_offsets, dtype=np.int64)
    return int(lbo[boundaries[rank]]), int(lbo[boundaries[rank + 1]])


def shard_vroom(bin_bytes, list_lengths, list_byte_offsets, rank, world):
    """returns (slice of the .bin for this rank, first list id, one-past-last list id)"""
    b = split_by_postings(list_lengths, world)
    lo, hi = shard_byte_range(list_byte_offsets, b, rank)
    return bin_bytes[lo:hi], int(b[rank]), int(b[rank + 1])


def reduce_totals(count, checksum, device=None):
    """SUM all-reduce of {decoded integers, checksum mod 2^63} over the default process group
    (the job's only collective). Returns python ints; a no-op without torch.distributed."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(count), int(checksum) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        hi = t >> 32
        lo = t & 0xFFFFFFFF
        parts = torch.stack([hi, lo])  # sum halves separately: no int64 overflow for <= 2^31 ranks
        dist.all_reduce(parts, op=dist.ReduceOp.SUM)
        tot = [(int(parts[0][i]) << 32) + int(parts[1][i]) for i in range(2)]
        return tot[0], tot[1] & 0x7FFFFFFFFFFFFFFF
    return int(t[0]), int(t[1])
