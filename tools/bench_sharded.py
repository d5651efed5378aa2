#!/usr/bin/env python
"""Sharded-index measurement (BASELINE configs[3] shape: dims 1536, L2, index sharded by id range over the GPUs of one box,
one NCCL all-gather of per-shard top-k + merge kernel).  Launch: torchrun --nproc-per-node G tools/bench_sharded.py --n <total N>
Every rank searches ALL queries on its shard; value = queries/s of the whole job (max over ranks)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=1536); ap.add_argument("--rows", dest="n", type=int, default=4_000_000)
ap.add_argument("--metric", default="l2"); ap.add_argument("--m", type=int, default=32)
ap.add_argument("--efc", type=int, default=200); ap.add_argument("--efs", type=int, default=64)
ap.add_argument("--batch", type=int, default=16384); ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.load()
lo, hi = sharded.shard_bounds(a.n, world)[rank]
g = torch.Generator(device="cuda"); g.manual_seed(99)
centres = torch.randn((max(4, int(round(a.n ** 0.5))), a.dims), generator=g, device="cuda")
sigma = 0.3 * (2.0 * a.dims) ** 0.5 / a.dims ** 0.5
def gen(k, seed):
    g.manual_seed(seed)
    out = torch.empty((k, a.dims), device="cuda")
    for s in range(0, k, 1 << 16):
        e = min(k, s + (1 << 16))
        c = torch.randint(0, centres.shape[0], (e - s,), generator=g, device="cuda")
        out[s:e] = centres[c] + sigma * torch.randn((e - s, a.dims), generator=g, device="cuda")
    return out
X = gen(hi - lo, 1234 + rank)                     # this rank's id range
Q = gen(a.batch * (a.steps + 2), 5678)            # same queries on every rank
idx = pg.HnswIndex(a.dims, a.m, a.efc, a.efs, a.metric, capacity=hi - lo, device=local)
labels = torch.arange(lo, hi, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.pgemb_index_append_device(idx.dev, hi - lo, X.data_ptr(), labels.data_ptr(), None, st)); torch.cuda.synchronize()
build_s = idx.build_appended(0, hi - lo, 4096)
if world > 1:
    srch = sharded.ShardedSearch(sharded.gpu_local_search(idx), sharded.gpu_merge())
    run = lambda q: srch.search(q, a.efs)
else:
    loc = sharded.gpu_local_search(idx)
    run = lambda q: loc(q, a.efs)
B = a.batch
for s in range(2): run(Q[s * B:(s + 1) * B])
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(2, 2 + a.steps): out = run(Q[s * B:(s + 1) * B])
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"mode": f"sharded x{world}" if world > 1 else "single", "shape": vars(a), "shard_rows": hi - lo, "build_s": round(build_s, 1),
                      "qps": round(B * a.steps / (ms.item() * 1e-3), 1), "ms_per_step": round(ms.item() / a.steps, 2),
                      "exchange_bytes_per_rank_per_step": B * a.efs * 12 + B * 4}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
