#!/usr/bin/env python
"""Profile target for the latency-mode kernel: build the BASELINE graph, warm up, then bracket a few single-query
launches with cudaProfilerStart/Stop (run under `ncu --profile-from-start off`).  The launches go through the
DEVICE-pointer entry point: a replayed kernel must not depend on a concurrent copy stream (profiles/README.md)."""
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
ef = bench.EFS
d_lab = torch.empty((1, ef), dtype=torch.int64, device="cuda")
d_n = torch.empty((1,), dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
def one(i):
    _lib.check(lib.pgemb_search_batch_device(idx.dev, 1, Q[i:i + 1].data_ptr(), ef, d_lab.data_ptr(), None, None, d_n.data_ptr(), None, stream))
    torch.cuda.synchronize()
for i in range(10): one(i)
torch.cuda.profiler.start()
for i in range(10, 12): one(i)
torch.cuda.profiler.stop()
print("done")
