#!/usr/bin/env python
"""hnsw_bind_point at index size N: a base graph of N rows (GPU bulk build, exported to the CPU checker), then `--inserts`
exact sequential inserts timed on the reference (one CPU core -- inserts are serial by design, embedding.c:624-629) and on
the GPU (pgemb_insert_batch == n x hnsw_bind_point); the link lists of the whole graph must be identical afterwards."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
from oracle import oracle
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200000); ap.add_argument("--inserts", type=int, default=200)
a = ap.parse_args()
lib = _lib.load()
X, _ = bench.make_data(torch, a.n + a.inserts, 8)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=a.n + a.inserts)
_lib.check(lib.pgemb_index_append_device(idx.dev, a.n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
idx.build_appended(0, a.n, 4096)
which = "ref" if oracle.available("ref") else "port"
orc = oracle.FlatIndex(which, bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=a.n + a.inserts)
for s in range(0, a.n, 1 << 16):
    orc.load_records(idx.export_records(s, min(1 << 16, a.n - s)))
new = X[a.n:].cpu().numpy()
t0 = time.perf_counter()
for i in range(a.inserts):
    orc.add(new[i], a.n + i)
t_cpu = time.perf_counter() - t0
idx.insert_many(new[:8])            # warm-up
t0 = time.perf_counter(); idx.insert_many(new[8:]); t_gpu = time.perf_counter() - t0
same = bool(idx.links().tobytes() == orc.links().tobytes())
print(json.dumps({"n": a.n, "inserts": a.inserts, "cpu_reference_ms_per_insert": round(1e3 * t_cpu / a.inserts, 3),
                  "gpu_exact_ms_per_insert": round(1e3 * t_gpu / (a.inserts - 8), 3), "links_identical": same, "checker": which}))
