#!/usr/bin/env python
"""A/B of the latency prototypes in ONE process (the 1M x 768 cosine graph is built once): single-query hnsw_search latency
and one-query kernel time for a list of flag combinations.  The flags are read by the library at every call, so they can be
switched through os.environ between measurements.  Needs the prototype library:

    PGEMB_LIB_VARIANT=proto python tools/bench_latency_ab.py
"""
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
assert b"+proto" in lib.pgemb_version(), "set PGEMB_LIB_VARIANT=proto"
X, Q = bench.make_data(torch, n, 512)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=n)
_lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
idx.build_appended(0, n, 4096)
q = Q.cpu().numpy()
ALL = ["PGEMB_VISITED_PAIRS", "PGEMB_SMEM_VISITED", "PGEMB_FAST_SMALL", "PGEMB_GATHER_LDGSTS"]
COMBOS = [
    {},
    {"PGEMB_VISITED_PAIRS": "1"},
    {"PGEMB_SMEM_VISITED": "4096"},
    {"PGEMB_GATHER_LDGSTS": "1"},
    {"PGEMB_FAST_SMALL": "1"},
    {"PGEMB_VISITED_PAIRS": "1", "PGEMB_SMEM_VISITED": "4096"},
    {"PGEMB_VISITED_PAIRS": "1", "PGEMB_SMEM_VISITED": "4096", "PGEMB_GATHER_LDGSTS": "1"},
    {"PGEMB_VISITED_PAIRS": "1", "PGEMB_SMEM_VISITED": "4096", "PGEMB_GATHER_LDGSTS": "1", "PGEMB_FAST_SMALL": "1"},
]
ref = None
for combo in COMBOS:
    for k in ALL:
        os.environ.pop(k, None)
    os.environ.update(combo)
    for i in range(20): idx.search(q[i])
    ts, labs = [], []
    for i in range(20, 320):
        t0 = time.perf_counter(); r = idx.search(q[i]); ts.append(time.perf_counter() - t0); labs.append(r.tolist())
    if ref is None: ref = labs
    out = idx.search_batch(q[20:21], 64)
    b16 = idx.search_batch(q[:16], 64); b16 = idx.search_batch(q[:16], 64)
    print(json.dumps({"flags": combo, "hnsw_search_ms_median": round(1e3 * float(np.median(ts)), 3), "hnsw_search_ms_p95": round(1e3 * float(np.percentile(ts, 95)), 3),
                      "kernel_ms_one_query": round(out["kernel_ms"], 3), "kernel_ms_batch16": round(b16["kernel_ms"], 3), "same_results_as_default": labs == ref}), flush=True)
