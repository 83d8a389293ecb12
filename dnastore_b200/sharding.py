"""Multi-GPU sharding of a read batch.

Reads are independent (the reference's decodeFastSeqs loop carries no state between reads,
reference src/viterbi.cpp:312-318), so a batch is cut into `world` CONTIGUOUS ranges of the read
index space, balanced by DP work (sum of len+1), one range per rank / GPU.  There is no data-path
collective: every rank decodes its range with its own Decoder and only the decoded strings and
log-likelihoods are gathered on the host, in input order.
"""
import numpy as np


def shard_bounds(read_len, world):
    """Contiguous ranges [lo, hi) per rank with ~equal sum(len+1). Returns an int64 array of world+1 cuts."""
    read_len = np.asarray(read_len, dtype=np.int64)
    n = len(read_len)
    if world <= 1 or n == 0:
        return np.array([0] + [n] * max(world, 1), dtype=np.int64)
    work = np.cumsum(read_len + 1)
    total = work[-1]
    cuts = [0]
    for r in range(1, world):
        # first index whose cumulative work reaches r/world of the total
        cuts.append(int(np.searchsorted(work, total * r / world, side="left")))
    cuts.append(n)
    return np.maximum.accumulate(np.array(cuts, dtype=np.int64))


def my_range(read_len, rank, world):
    cuts = shard_bounds(read_len, world)
    return int(cuts[rank]), int(cuts[rank + 1])


def gather_in_order(local_results, rank, world, group=None):
    """Host-side ordered gather of per-rank result lists (decoded strings, log-likelihoods ...).

    Uses torch.distributed.all_gather_object (works on gloo and nccl groups); because the shards are
    contiguous ranges in rank order, concatenating in rank order restores input order."""
    if world <= 1:
        return list(local_results)
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, list(local_results), group=group)
    out = []
    for p in parts:
        out.extend(p)
    return out


def decode_sharded(decode_fn, reads, rank, world, group=None):
    """Decode `reads` (identical list on every rank) cooperatively: rank r decodes its contiguous
    shard with decode_fn(list_of_reads) -> list_of_results, then everything is gathered in input order."""
    lo, hi = my_range([len(r) for r in reads], rank, world)
    local = decode_fn(reads[lo:hi]) if hi > lo else []
    return gather_in_order(local, rank, world, group)
