"""Multi-GPU layer of the hot path: one process per GPU, torch.distributed for the
plumbing.

Two modes:

* replicas  -- the index fits one GPU (BASELINE configs[1..3]): every rank holds the whole graph and the
               queries are split; no data-path collective (`split_queries`).
* shards    -- contiguous id-range shards (configs[3], [4]): rank g holds ids [g*N/G, (g+1)*N/G) with its
               own independent single-layer graph (entry = its first node); every query is searched on
               every shard, the per-shard top-k (dist f32, label u64, count i32) are exchanged with ONE
               all-gather and merged per query by `pgemb_merge_topk_device` in the reference's
               (dist,label) pair order (hnswalg.cpp:236-247).

The exchange (`exchange_topk`) is backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the local search
and the merge are injected so that the CPU tests can drive the same host logic with the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous id ranges: shard g = [g*n//world, (g+1)*n//world)."""
    return [(g * n // world, (g + 1) * n // world) for g in range(world)]


def split_queries(nq: int, world: int, rank: int) -> tuple[int, int]:
    """Replica mode: the slice of a query batch this rank serves."""
    lo, hi = shard_bounds(nq, world)[rank]
    return lo, hi


def exchange_topk(dists, labels, counts, group=None):
    """All-gather the per-shard results.  Inputs: dists [nq,k] f32, labels [nq,k] i64 (u64 bit pattern),
    counts [nq] i32 on this rank.  Returns (D [world,nq,k], L [world,nq,k], N [world,nq]) -- the layout
    pgemb_merge_topk_device expects ([shard][query][k])."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    outs = []
    for t in (dists, labels, counts):
        t = t.contiguous()
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t, group=group)
        outs.append(torch.stack(bufs, 0))
    return tuple(outs)


class ShardedSearch:
    """Search every shard, exchange, merge.  `local_search(queries, ef) -> (dists, labels, counts)` and
    `merge(D, L, N, k) -> (dists, labels, counts)` are injected (GPU: HnswIndex + merge kernel)."""

    def __init__(self, local_search: Callable, merge: Callable, group=None):
        self.local_search, self.merge, self.group = local_search, merge, group

    def search(self, queries, ef: int):
        d, l, n = self.local_search(queries, ef)
        D, L, N = exchange_topk(d, l, n, self.group)
        return self.merge(D, L, N, ef)


def gpu_local_search(index):
    """local_search for a pg_embedding_b200.HnswIndex whose labels are global (e.g. global node ids)."""
    import torch
    from . import _lib

    lib = _lib.load()

    def run(queries, ef):
        nq = queries.shape[0]
        d = torch.empty((nq, ef), dtype=torch.float32, device=queries.device)
        l = torch.empty((nq, ef), dtype=torch.int64, device=queries.device)
        n = torch.empty((nq,), dtype=torch.int32, device=queries.device)
        _lib.check(lib.pgemb_search_batch_device(index.dev, nq, queries.data_ptr(), ef, l.data_ptr(), d.data_ptr(), None,
                                                  n.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
        return d, l, n

    return run


def gpu_merge():
    import torch
    from . import _lib

    lib = _lib.load()

    def run(D, L, N, k):
        world, nq, _ = D.shape
        od = torch.empty((nq, k), dtype=torch.float32, device=D.device)
        ol = torch.empty((nq, k), dtype=torch.int64, device=D.device)
        on = torch.empty((nq,), dtype=torch.int32, device=D.device)
        _lib.check(lib.pgemb_merge_topk_device(nq, world, k, D.data_ptr(), L.data_ptr(), N.data_ptr(), od.data_ptr(), ol.data_ptr(),
                                                on.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return od, ol, on

    return run
