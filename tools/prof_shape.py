#!/usr/bin/env python
"""Profile target for the throughput-mode traversal at any shape: build the graph, warm up, bracket ONE batch launch with
cudaProfilerStart/Stop (run under `ncu --profile-from-start off`).  Device-pointer entry point.
usage: prof_shape.py --dims 1536 --n 500000 --metric l2 --m 32 [--batch 32768]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=768); ap.add_argument("--n", type=int, default=1000000); ap.add_argument("--metric", default="cosine")
ap.add_argument("--m", type=int, default=32); ap.add_argument("--batch", type=int, default=32768); ap.add_argument("--efs", type=int, default=64)
a = ap.parse_args()
lib = _lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(99)
centres = torch.randn((max(4, int(round(a.n ** 0.5))), a.dims), generator=g, device="cuda")
gen = bench.gen_points if a.metric == "cosine" else bench.gen_points_raw
X, Q = gen(torch, a.n, 1234, centres), gen(torch, a.batch * 3, 5678, centres)
idx = pg.HnswIndex(a.dims, a.m, 200, a.efs, a.metric, capacity=a.n)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.pgemb_index_append_device(idx.dev, a.n, X.data_ptr(), None, None, st)); torch.cuda.synchronize()
idx.build_appended(0, a.n, 4096)
B = a.batch
d_lab = torch.empty((B, a.efs), dtype=torch.int64, device="cuda"); d_n = torch.empty((B,), dtype=torch.int32, device="cuda")
def step(s):
    _lib.check(lib.pgemb_search_batch_device(idx.dev, B, Q[s * B:(s + 1) * B].data_ptr(), a.efs, d_lab.data_ptr(), None, None, d_n.data_ptr(), None, st))
    torch.cuda.synchronize()
step(0); step(1)
torch.cuda.profiler.start()
step(2)
torch.cuda.profiler.stop()
print("done", float(lib.pgemb_last_kernel_ms(idx.dev)))
