#!/usr/bin/env python
"""Profile target for the exact-scan kernels: append PGEMB_BENCH_N rows of the BASELINE shape, warm up, then bracket one
pgemb_scan_topk call with cudaProfilerStart/Stop (run under `ncu --profile-from-start off`; the flags PGEMB_SCAN_TILED /
PGEMB_SCAN_TC and PGEMB_LIB_VARIANT=proto select the variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
n = int(os.environ.get("PGEMB_BENCH_N", 200_000))
nq = int(os.environ.get("PGEMB_PROF_SCAN", 64))
k = int(os.environ.get("PGEMB_PROF_SCAN_K", 10))
lib = _lib.load()
X, Q = bench.make_data(torch, n, nq)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=n)
_lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
q = Q.cpu().numpy()
idx.scan_topk(q[:2], k)
torch.cuda.profiler.start()
out = idx.scan_topk(q, k)
torch.cuda.profiler.stop()
print("done", out["n"][:4].tolist())
