"""2-GPU test of the sharded path: id-range shards, one NCCL all-gather of per-shard
top-k, merge kernel; checked against the oracle run per shard + the same merge on CPU.  Skipped on 1 GPU."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT)
    import pg_embedding_b200 as pg
    from pg_embedding_b200 import sharded
    rng = np.random.default_rng(11)
    n, dims, m, efc, ef, nq = 6000, 64, 8, 48, 32, 256
    x = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal((nq, dims)).astype(np.float32)
    lo, hi = sharded.shard_bounds(n, world)[rank]
    idx = pg.HnswIndex(dims, m, efc, ef, "l2", capacity=hi - lo, device=rank)
    idx.append(x[lo:hi], np.arange(lo, hi, dtype=np.uint64))       # labels = global ids
    idx.build_appended(0, hi - lo, 1)                              # exact sequential build of the shard
    s = sharded.ShardedSearch(sharded.gpu_local_search(idx), sharded.gpu_merge())
    od, ol, on = s.search(torch.from_numpy(q).cuda(), ef)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), {"ol": ol.cpu().numpy(), "on": on.cpu().numpy(), "od": od.cpu().numpy()}, allow_pickle=True)
    dist.barrier(); dist.destroy_process_group()


def test_sharded_two_gpus(tmp_path, oracle_mod):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pg_embedding_b200 import sharded
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npy", allow_pickle=True).item() for r in range(world)]
    assert (res[0]["ol"] == res[1]["ol"]).all()
    rng = np.random.default_rng(11)
    n, dims, m, efc, ef, nq = 6000, 64, 8, 48, 32, 256
    x = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal((nq, dims)).astype(np.float32)
    pairs = [[] for _ in range(nq)]
    for lo, hi in sharded.shard_bounds(n, world):
        sh = oracle_mod.FlatIndex("port", dims, m, efc, ef, "l2", capacity=hi - lo)
        sh.build(x[lo:hi], np.arange(lo, hi, dtype=np.uint64))
        r = sh.search_many(q, ef)
        for i in range(nq):
            c = int(r["n"][i]); labs = r["labels"][i, :c].astype(np.int64)
            dd = oracle_mod.dist_many("port", "l2", q[i], x[labs])
            pairs[i] += list(zip(dd.tolist(), labs.tolist()))
    for i in range(nq):
        want = sorted(pairs[i])[:ef]
        assert res[0]["on"][i] == len(want)
        assert res[0]["ol"][i, :len(want)].tolist() == [w[1] for w in want]
        assert np.array([w[0] for w in want], np.float32).tobytes() == res[0]["od"][i, :len(want)].tobytes()
