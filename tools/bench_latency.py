#!/usr/bin/env python
"""Single-query latency of the reference-shaped hnsw_search (what a Postgres backend would call, one query per call)
on the BASELINE 1M x 768 cosine graph, vs the reference CPU code single-threaded on the same graph."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
n = int(os.environ.get("PGEMB_BENCH_N", 1_000_000))
lib = _lib.load()
X, Q = bench.make_data(torch, n, 512)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=n)
_lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
idx.build_appended(0, n, 4096)
q = Q.cpu().numpy()
res = {"n": n}
for mode, tag in (("1", "latency_mode"), ("0", "throughput_mode")):
    os.environ["PGEMB_COOP"] = mode     # read by the library at every launch
    for i in range(20): idx.search(q[i])
    ts = []
    for i in range(20, 320):
        t0 = time.perf_counter(); idx.search(q[i]); ts.append(time.perf_counter() - t0)
    out = idx.search_batch(q[20:21], 64)
    res[tag] = {"hnsw_search_ms_median": round(1e3 * float(np.median(ts)), 3), "hnsw_search_ms_p95": round(1e3 * float(np.percentile(ts, 95)), 3),
                "kernel_ms_one_query": round(out["kernel_ms"], 3)}
    for nb in (16, 148, 512):
        idx.search_batch(q[:nb], 64)
        o2 = idx.search_batch(q[:nb], 64)
        res[tag][f"kernel_ms_batch{nb}"] = round(o2["kernel_ms"], 3)
os.environ.pop("PGEMB_COOP")
if "--cpu" in sys.argv:
    which, kind = bench.pick_checker()
    orc = bench.host_graph(idx, n, which)
    orc.search_many(q[:64], 64, nthreads=1, want_labels=False)
    r = orc.search_many(q[64:320], 64, nthreads=1, want_labels=False)
    res["cpu_reference_ms_per_query_1thread"] = round(1e3 * r["seconds"] / 256, 3)
    res["cpu_kind"] = kind
print(json.dumps(res))
