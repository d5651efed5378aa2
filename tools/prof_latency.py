#!/usr/bin/env python
"""Profile target for the latency-mode kernel: build the BASELINE graph, warm up, then bracket a few single-query
hnsw_search calls with cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
n = int(os.environ.get("PGEMB_BENCH_N", 1_000_000))
lib = _lib.load()
X, Q = bench.make_data(torch, n, 64)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=n)
_lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
idx.build_appended(0, n, 4096)
q = Q.cpu().numpy()
for i in range(10): idx.search(q[i])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(10, 12): idx.search(q[i])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
