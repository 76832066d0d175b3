"""Read sharding for the multi-GPU EM: contiguous read ranges with about H / G hits each.

Same greedy rule as the reference uses to split reads over its threads (/root/reference/EM.cpp:135-157):
worker i keeps taking reads while at least one read is left for every remaining worker and (unless it is the
last worker) its hit count is still below floor(nHits / nWorkers).  Row order is preserved, so `.ofg` / posterior
order is independent of the number of GPUs.  Host-side index arithmetic only."""
from __future__ import annotations

import numpy as np


def shard_reads(row_ptr, n_ranks: int):
    """-> list of (first_read, last_read_exclusive) per rank"""
    row_ptr = np.asarray(row_ptr, dtype=np.uint64)
    n = len(row_ptr) - 1
    n_hits = int(row_ptr[-1])
    n_ranks = max(1, min(n_ranks, n)) if n > 0 else 1
    thr = n_hits // n_ranks
    out, cur = [], 0
    for i in range(n_ranks):
        left_workers = n_ranks - i - 1
        if i == n_ranks - 1:
            end = n
        else:
            # smallest end with hits(cur, end) >= thr, capped so that every later worker still gets a read
            target = int(row_ptr[cur]) + thr
            end = int(np.searchsorted(row_ptr, target, side="left"))
            end = max(end, cur)          # thr == 0: take nothing unless forced
            end = min(end, n - left_workers)
            if end < cur:
                end = cur
        out.append((cur, end))
        cur = end
    return out


def slice_csr(row_ptr, first: int, last: int):
    """row_ptr of the shard (rebased to 0) and the hit range it covers"""
    row_ptr = np.asarray(row_ptr, dtype=np.uint64)
    h0, h1 = int(row_ptr[first]), int(row_ptr[last])
    return (row_ptr[first:last + 1] - np.uint64(h0)).astype(np.uint64), h0, h1
