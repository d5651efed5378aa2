"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in pg_embedding_b200/sharded.py:
id-range partition, query split, the packed per-rank result buffer and the ONE all-gather that exchanges it.
The local search and the merge are injected: here they are the ORACLE (reference per shard, SURVEY.md
section 8(e) "the oracle for sharded configs") and a numpy merge, so no compute runs in the product."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def merge_topk_numpy(D, L, N, k):
    world, nq, _ = D.shape
    od = np.full((nq, k), np.inf, dtype=np.float32)
    ol = np.full((nq, k), -1, dtype=np.int64)
    on = np.zeros(nq, dtype=np.int32)
    for q in range(nq):
        pairs = sorted((float(D[s, q, i]), int(np.uint64(L[s, q, i]))) for s in range(world) for i in range(int(N[s, q])))[:k]
        on[q] = len(pairs)
        for i, (d, l) in enumerate(pairs):
            od[q, i], ol[q, i] = d, l
    return od, ol, on


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    from pg_embedding_b200 import sharded

    rng = np.random.default_rng(3)
    n, dims, m, efc, ef, nq = 1201, 12, 6, 24, 16, 40
    x = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal((nq, dims)).astype(np.float32)
    bounds = sharded.shard_bounds(n, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    lo, hi = bounds[rank]
    shard = oracle.FlatIndex("port", dims, m, efc, ef, "l2", capacity=hi - lo)
    shard.build(x[lo:hi], np.arange(lo, hi, dtype=np.uint64))      # labels = global ids

    def local_search(queries, k):
        r = shard.search_many(queries.numpy(), k)
        lab = r["labels"].astype(np.uint64).view(np.int64)
        d = np.full((queries.shape[0], k), np.inf, np.float32)
        for i in range(queries.shape[0]):
            c = int(r["n"][i])
            if c:
                d[i, :c] = oracle.dist_many("port", "l2", queries[i].numpy(), x[r["labels"][i, :c].astype(np.int64)])
        return torch.from_numpy(d), torch.from_numpy(lab.copy()), torch.from_numpy(r["n"].copy())

    def merge(buf, world_, nq_, k):
        D, L, N = sharded.unpack_topk(buf, world_, nq_, k)
        return merge_topk_numpy(D.numpy(), L.numpy(), N.numpy(), k)

    s = sharded.ShardedSearch(local_search, merge)
    od, ol, on = s.search(torch.from_numpy(q), ef)
    assert s.collectives == 1                                     # one all-gather per step, nothing else
    # the sharded brute-force scan (BASELINE configs[4]'s host logic): every rank scans ITS rows for the whole batch -- here
    # the oracle's distances, ordered by (dist,label) -- and the same exchange + merge yields the top-k over the whole table
    kscan = 7
    lab_all = (np.arange(n, dtype=np.uint64) * np.uint64(3)) + np.uint64(1)

    def local_scan(queries, k):
        qn = queries.numpy()
        d = np.full((qn.shape[0], k), np.inf, np.float32); lab = np.full((qn.shape[0], k), -1, np.int64); cnt = np.zeros(qn.shape[0], np.int32)
        for i in range(qn.shape[0]):
            dd = oracle.dist_many("port", "l2", qn[i], x[lo:hi])
            order = sorted((float(dd[j]), int(lab_all[lo + j])) for j in range(hi - lo))[:k]
            cnt[i] = len(order)
            d[i, :len(order)] = [o[0] for o in order]; lab[i, :len(order)] = [o[1] for o in order]
        return torch.from_numpy(d), torch.from_numpy(lab), torch.from_numpy(cnt)

    s2 = sharded.ShardedSearch(local_scan, merge)
    sd, sl, sn = s2.search(torch.from_numpy(q[:9]), kscan)
    assert s2.collectives == 1
    # replica mode: query split covers the batch exactly once
    a, b = sharded.split_queries(nq, world, rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), {"od": od, "ol": ol, "on": on, "split": (a, b), "sd": sd, "sl": sl, "sn": sn}, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_world2(tmp_path, oracle_mod):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npy", allow_pickle=True).item() for r in range(world)]
    # every rank ends with the same merged answer
    assert (res[0]["ol"] == res[1]["ol"]).all() and (res[0]["on"] == res[1]["on"]).all()
    assert res[0]["split"] == (0, 20) and res[1]["split"] == (20, 40)
    # the sharded scan: identical on both ranks and equal to a brute force over the whole table in (dist,label) order
    assert (res[0]["sl"] == res[1]["sl"]).all() and (res[0]["sd"] == res[1]["sd"]).all() and (res[0]["sn"] == res[1]["sn"]).all()
    rng0 = np.random.default_rng(3)
    x0 = rng0.standard_normal((1201, 12)).astype(np.float32); q0 = rng0.standard_normal((40, 12)).astype(np.float32)
    lab0 = (np.arange(1201, dtype=np.uint64) * np.uint64(3)) + np.uint64(1)
    for i in range(9):
        dd = oracle_mod.dist_many("port", "l2", q0[i], x0)
        want = sorted((float(dd[j]), int(lab0[j])) for j in range(1201))[:7]
        assert res[0]["sn"][i] == 7 and res[0]["sl"][i].tolist() == [w[1] for w in want]
        assert np.array([w[0] for w in want], np.float32).tobytes() == np.asarray(res[0]["sd"][i], np.float32).tobytes()
    # and it equals the oracle run per shard + merged on one process
    from pg_embedding_b200 import sharded
    rng = np.random.default_rng(3)
    n, dims, m, efc, ef, nq = 1201, 12, 6, 24, 16, 40
    x = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal((nq, dims)).astype(np.float32)
    D, L, N = [], [], []
    for lo, hi in sharded.shard_bounds(n, world):
        sh = oracle_mod.FlatIndex("port", dims, m, efc, ef, "l2", capacity=hi - lo)
        sh.build(x[lo:hi], np.arange(lo, hi, dtype=np.uint64))
        r = sh.search_many(q, ef)
        d = np.full((nq, ef), np.inf, np.float32)
        for i in range(nq):
            c = int(r["n"][i])
            d[i, :c] = oracle_mod.dist_many("port", "l2", q[i], x[r["labels"][i, :c].astype(np.int64)])
        D.append(d); L.append(r["labels"].view(np.int64)); N.append(r["n"])
    od, ol, on = merge_topk_numpy(np.stack(D), np.stack(L), np.stack(N), ef)
    assert (ol == res[0]["ol"]).all() and (on == res[0]["on"]).all()
    # merged lists are sorted by (dist,label) and hold distinct global ids
    for i in range(nq):
        c = on[i]
        assert list(od[i, :c]) == sorted(od[i, :c]) and len(set(ol[i, :c].tolist())) == c
