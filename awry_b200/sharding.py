"""Range partitioning of a query batch over GPUs / ranks, and the host-side gather.

The search path shards by independent queries (fm_index.rs:455-487 maps over them): the index is
replicated, the batch is cut into contiguous ranges balanced by query BYTES, every rank searches
its range, results are concatenated in range order.  No collective sits on the data path; the
only communication is the final gather of results to the caller's rank.

`split_by_bytes` is the same rule the C++ library applies across replicas inside one process
(`for_each_replica_range`, awry_b200/csrc/batch.cu).
"""
from typing import Callable, List, Optional, Tuple

import numpy as np


def split_by_bytes(qoff: np.ndarray, n_parts: int) -> List[Tuple[int, int]]:
    """Contiguous query ranges [(lo, hi), ...] covering [0, nq) with ~equal total query bytes."""
    qoff = np.asarray(qoff, dtype=np.uint64)
    nq = len(qoff) - 1
    if n_parts <= 1 or nq < 2 * n_parts:
        return [(0, nq)] + [(nq, nq)] * (max(n_parts, 1) - 1)
    total = int(qoff[nq] - qoff[0])
    cuts = [0]
    for i in range(1, n_parts):
        target = int(qoff[0]) + total * i // n_parts
        c = int(np.searchsorted(qoff[:nq], np.uint64(target), side="left"))
        cuts.append(max(c, cuts[-1]))
    cuts.append(nq)
    return [(cuts[i], cuts[i + 1]) for i in range(n_parts)]


def rank_slice(qbytes: np.ndarray, qoff: np.ndarray, rank: int, world: int):
    """The (qbytes, qoff) pair of this rank's range, rebased to start at offset 0, plus the range."""
    lo, hi = split_by_bytes(qoff, world)[rank]
    b0, b1 = int(qoff[lo]), int(qoff[hi])
    return qbytes[b0:b1], (qoff[lo:hi + 1] - qoff[lo]).astype(np.uint64), (lo, hi)


def sharded_count(qbytes: np.ndarray, qoff: np.ndarray, count_fn: Callable[[np.ndarray, np.ndarray], np.ndarray],
                  group=None, dst: int = 0) -> Optional[np.ndarray]:
    """parallel_count across the ranks of a torch.distributed group: every rank runs `count_fn`
    (e.g. FmIndex.count_packed) on its range; rank `dst` returns the counts in input order."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sb, so, (lo, hi) = rank_slice(qbytes, qoff, rank, world)
    local = np.ascontiguousarray(count_fn(sb, so), dtype=np.uint64) if hi > lo else np.zeros(0, np.uint64)
    ranges = split_by_bytes(qoff, world)
    longest = max(h - l for l, h in ranges)
    buf = torch.zeros(max(longest, 1), dtype=torch.int64)
    buf[: hi - lo] = torch.from_numpy(local.view(np.int64))
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.zeros(len(qoff) - 1, dtype=np.uint64)
    for (l, h), t in zip(ranges, gathered):
        out[l:h] = t[: h - l].numpy().view(np.uint64)
    return out


def sharded_locate(qbytes: np.ndarray, qoff: np.ndarray,
                   locate_fn: Callable[[np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray]],
                   group=None, dst: int = 0):
    """parallel_locate across ranks: returns (hit_off[nq+1], hits[n,2]) on rank `dst`."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sb, so, (lo, hi) = rank_slice(qbytes, qoff, rank, world)
    if hi > lo:
        off, hits = locate_fn(sb, so)
    else:
        off, hits = np.zeros(1, np.uint64), np.zeros((0, 2), np.uint64)
    payload = (lo, hi, np.asarray(off, dtype=np.uint64), np.asarray(hits, dtype=np.uint64))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    nq = len(qoff) - 1
    hit_off = np.zeros(nq + 1, dtype=np.uint64)
    parts, base = [], 0
    for l, h, off, hits in sorted(gathered, key=lambda p: p[0]):
        if h > l:
            hit_off[l:h + 1] = off + np.uint64(base)
        parts.append(hits)
        base += len(hits)
    hit_off[nq] = base
    return hit_off, (np.concatenate(parts) if parts else np.zeros((0, 2), np.uint64))
