#!/usr/bin/env python
"""Insert-path timing for DESIGN.md: exact sequential binds on the GPU (pgemb_insert_batch == n x hnsw_bind_point)
vs the reference on one CPU core (inserts are serial by design, embedding.c:624-629), same data, link lists compared."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pg_embedding_b200 as pg
from oracle import oracle
ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=768); ap.add_argument("--n", type=int, default=4000)
ap.add_argument("--metric", default="cosine"); ap.add_argument("--m", type=int, default=32); ap.add_argument("--efc", type=int, default=200)
a = ap.parse_args()
rng = np.random.default_rng(1)
c = rng.standard_normal((max(4, int(a.n ** 0.5)), a.dims)).astype(np.float32)
x = c[rng.integers(0, len(c), a.n)] + 0.42 * rng.standard_normal((a.n, a.dims)).astype(np.float32)
x = np.ascontiguousarray(x / np.linalg.norm(x, axis=1, keepdims=True), dtype=np.float32)
which = "ref" if oracle.available("ref") else "port"
orc = oracle.FlatIndex(which, a.dims, a.m, a.efc, 64, a.metric, capacity=a.n)
t_cpu = orc.build(x)
idx = pg.HnswIndex(a.dims, a.m, a.efc, 64, a.metric, capacity=a.n)
idx.insert_many(x[:64])            # warm-up (workspace allocation, module load)
t0 = time.perf_counter(); idx.insert_many(x[64:]); t_gpu = time.perf_counter() - t0
same = bool(idx.links().tobytes() == orc.links().tobytes())
print(json.dumps({"shape": vars(a), "cpu_reference_ms_per_insert": round(1e3 * t_cpu / a.n, 3), "gpu_exact_ms_per_insert": round(1e3 * t_gpu / (a.n - 64), 3),
                  "links_identical": same, "checker": which}))
