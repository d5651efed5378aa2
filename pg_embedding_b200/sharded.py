"""Multi-GPU layer of the hot path: one process per GPU, torch.distributed for the
plumbing.

Two modes:

* replicas  -- the index fits one GPU (BASELINE configs[1..3]): every rank holds the whole graph and the
               queries are split; no data-path collective (`split_queries`).
* shards    -- contiguous id-range shards (configs[3], [4]): rank g holds ids [g*N/G, (g+1)*N/G) with its
               own independent single-layer graph (entry = its first node); every query is searched on
               every shard and the per-shard top-k lists are merged per query in the reference's
               (dist,label) pair order (hnswalg.cpp:236-247).  Two exchanges:

               - `ShardedSearch`: the lists of a rank are ONE packed buffer
                 [labels u64 nq*k | dists f32 nq*k | counts i32 nq]; ONE all-gather (NCCL on GPUs, gloo in the
                 CPU tests) delivers all of them and `pgemb_merge_topk_packed_device` merges: 1 collective + 1 kernel;
               - `PeerExchange` (GPU only): no collective at all -- every rank's search writes into a buffer its peers
                 have mapped (CUDA IPC), a 4-byte flag copy per peer publishes the step, and ONE kernel per rank waits
                 for the flags, reads the peers' lists over NVLink and merges (`pgemb_exchange_*`, csrc/capi.cu).

The local search and the merge of `ShardedSearch` are injected so that the CPU tests can drive the same host logic with
the oracle.
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


def packed_bytes(nq: int, k: int) -> int:
    """Size of one rank's packed result buffer, padded to 8 bytes so that every shard's block starts aligned."""
    return (nq * k * 12 + nq * 4 + 7) & ~7


def pack_topk(dists, labels, counts):
    """(dists [nq,k] f32, labels [nq,k] i64, counts [nq] i32) -> one uint8 tensor [labels | dists | counts | pad]."""
    import torch
    nq, k = dists.shape
    out = torch.zeros(packed_bytes(nq, k), dtype=torch.uint8, device=dists.device)
    out[: nq * k * 8] = labels.contiguous().view(torch.uint8).reshape(-1)
    out[nq * k * 8: nq * k * 12] = dists.contiguous().view(torch.uint8).reshape(-1)
    out[nq * k * 12: nq * k * 12 + nq * 4] = counts.contiguous().view(torch.uint8).reshape(-1)
    return out


def unpack_topk(buf, world: int, nq: int, k: int):
    """[world * packed_bytes] uint8 -> (D [world,nq,k] f32, L [world,nq,k] i64, N [world,nq] i32)."""
    import torch
    b = buf.view(world, -1)
    L = b[:, : nq * k * 8].contiguous().view(torch.int64).view(world, nq, k)
    D = b[:, nq * k * 8: nq * k * 12].contiguous().view(torch.float32).view(world, nq, k)
    N = b[:, nq * k * 12: nq * k * 12 + nq * 4].contiguous().view(torch.int32).view(world, nq)
    return D, L, N


def exchange_packed(packed, group=None):
    """THE collective of the sharded path: one all-gather of the packed per-shard results.  Returns [world * bytes] uint8."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty(world * packed.numel(), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out


class ShardedSearch:
    """Search every shard, ONE all-gather, merge.  `local_search(queries, ef) -> (dists, labels, counts)` or -> a packed
    uint8 tensor; `merge_packed(buf, world, nq, k) -> (dists, labels, counts)` (GPU: HnswIndex + merge kernel)."""

    def __init__(self, local_search: Callable, merge_packed: Callable, group=None):
        self.local_search, self.merge_packed, self.group = local_search, merge_packed, group
        self.collectives = 0

    def search(self, queries, ef: int):
        import torch.distributed as dist
        res = self.local_search(queries, ef)
        packed = res if not isinstance(res, tuple) else pack_topk(*res)
        buf = exchange_packed(packed, self.group)
        self.collectives += 1
        return self.merge_packed(buf, dist.get_world_size(self.group), queries.shape[0], ef)


def gpu_local_search_packed(index):
    """local_search for a pg_embedding_b200.HnswIndex whose labels are global (e.g. global node ids): the traversal writes its
    three outputs straight into one packed buffer (no packing pass)."""
    import torch
    from . import _lib

    lib = _lib.load()

    def run(queries, ef):
        nq = queries.shape[0]
        buf = torch.zeros(packed_bytes(nq, ef), dtype=torch.uint8, device=queries.device)
        base = buf.data_ptr()
        _lib.check(lib.pgemb_search_batch_device(index.dev, nq, queries.data_ptr(), ef, base, base + nq * ef * 8, None,
                                                  base + nq * ef * 12, None, torch.cuda.current_stream().cuda_stream))
        return buf

    return run


def gpu_local_scan_packed(index):
    """local_search for the brute-force scan (BASELINE configs[4]): `run(queries, k)` scans this rank's rows (pgemb_scan_topk_device,
    the tensor-core filter + exact re-scoring) straight into one packed buffer."""
    import torch
    from . import _lib

    lib = _lib.load()

    def run(queries, k):
        nq = queries.shape[0]
        buf = torch.zeros(packed_bytes(nq, k), dtype=torch.uint8, device=queries.device)
        base = buf.data_ptr()
        _lib.check(lib.pgemb_scan_topk_device(index.dev, nq, queries.data_ptr(), k, base, base + nq * k * 8, base + nq * k * 12,
                                               torch.cuda.current_stream().cuda_stream))
        return buf

    return run


def gpu_merge_packed():
    import torch
    from . import _lib

    lib = _lib.load()

    def run(buf, world, nq, k):
        od = torch.empty((nq, k), dtype=torch.float32, device=buf.device)
        ol = torch.empty((nq, k), dtype=torch.int64, device=buf.device)
        on = torch.empty((nq,), dtype=torch.int32, device=buf.device)
        _lib.check(lib.pgemb_merge_topk_packed_device(nq, world, k, buf.data_ptr(), packed_bytes(nq, k), od.data_ptr(), ol.data_ptr(),
                                                       on.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return od, ol, on

    return run


class PeerExchange:
    """Sharded search with NO collective in the data path (pgemb_exchange_*): set-up exchanges the CUDA-IPC handles once
    through torch.distributed; a step is the local traversal (+ a 4-byte flag copy per peer) and one wait+merge kernel."""

    def __init__(self, index, max_nq: int, k: int, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib

        self.lib, self.index, self.k = _lib.load(), index, int(k)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ex = C.c_void_p()
        _lib.check(self.lib.pgemb_exchange_create(index.device, self.rank, self.world, int(max_nq), self.k, C.byref(self.ex)))
        mine = (C.c_char * 64)()
        _lib.check(self.lib.pgemb_exchange_handle(self.ex, mine))
        t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).cuda()
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, t, group=group)                     # set-up only, not in the data path
        raw = bytes(allh.cpu().numpy().tobytes())
        _lib.check(self.lib.pgemb_exchange_attach(self.ex, raw, 0))
        dist.barrier(group)

    def search(self, queries, ef: int):
        import torch
        from . import _lib
        assert ef == self.k
        nq = queries.shape[0]
        st = torch.cuda.current_stream().cuda_stream
        od = torch.empty((nq, ef), dtype=torch.float32, device=queries.device)
        ol = torch.empty((nq, ef), dtype=torch.int64, device=queries.device)
        on = torch.empty((nq,), dtype=torch.int32, device=queries.device)
        _lib.check(self.lib.pgemb_sharded_search_device(self.index.dev, self.ex, nq, queries.data_ptr(), ef, st))
        _lib.check(self.lib.pgemb_sharded_merge_device(self.ex, nq, ol.data_ptr(), od.data_ptr(), on.data_ptr(), st))
        return od, ol, on

    def scan(self, queries, k: int):
        """The same exchange with the brute-force scan as the local step."""
        import torch
        from . import _lib
        assert k == self.k
        nq = queries.shape[0]
        st = torch.cuda.current_stream().cuda_stream
        od = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
        ol = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
        on = torch.empty((nq,), dtype=torch.int32, device=queries.device)
        _lib.check(self.lib.pgemb_sharded_scan_device(self.index.dev, self.ex, nq, queries.data_ptr(), k, st))
        _lib.check(self.lib.pgemb_sharded_merge_device(self.ex, nq, ol.data_ptr(), od.data_ptr(), on.data_ptr(), st))
        return od, ol, on

    def merge_ms(self) -> float:
        return float(self.lib.pgemb_exchange_last_merge_ms(self.ex))

    def error(self) -> int:
        return int(self.lib.pgemb_exchange_error(self.ex))

    def close(self):
        if self.ex:
            self.lib.pgemb_exchange_destroy(self.ex)
            self.ex = C.c_void_p()
