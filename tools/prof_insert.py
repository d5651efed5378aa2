#!/usr/bin/env python
"""Profile target for the insert path (K4): a bulk-built graph of PGEMB_BENCH_N rows of the BASELINE shape, then a few exact
hnsw_bind_point calls inside cudaProfilerStart/Stop: raw-mode traversal (latency mode) + select_kernel + backlink_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
n = int(os.environ.get("PGEMB_BENCH_N", 200_000))
lib = _lib.load()
X, Q = bench.make_data(torch, n + 16, 8)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=n + 16)
_lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
idx.build_appended(0, n, 4096)
x = X[n:].cpu().numpy()
idx.insert_many(x[:4])
torch.cuda.profiler.start()
idx.insert_many(x[4:6])
torch.cuda.profiler.stop()
print("done", len(idx))
