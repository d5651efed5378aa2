#!/usr/bin/env python
"""Build-path timing: pgemb_build_exact (bit-identical to sequential inserts) vs pgemb_build_bulk, same data as bench.py."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000); ap.add_argument("--bmax", type=int, default=256)
ap.add_argument("--mode", default="exact"); ap.add_argument("--bulk-first", type=int, default=0)
a = ap.parse_args()
lib = _lib.load()
X, Q = bench.make_data(torch, a.n, 1024)
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=a.n)
_lib.check(lib.pgemb_index_append_device(idx.dev, a.n, X.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
if a.mode == "exact":
    if a.bulk_first:
        idx.build_appended(0, a.bulk_first, 4096)      # steady-state probe: bulk-build a prefix, then exact for the rest
    secs, st = idx.build_exact(a.bulk_first, a.n - a.bulk_first, a.bmax)
else:
    secs, st = idx.build_appended(0, a.n, 4096), {}
out = idx.search_batch(Q.cpu().numpy(), 64)
truth = torch.topk(Q @ X.T, 10, dim=1).indices.cpu().numpy()
rec = float(np.mean([len(set(truth[i].tolist()) & set(out["labels"][i, :10].tolist())) / 10 for i in range(1024)]))
print(json.dumps({"n": a.n, "mode": a.mode, "bmax": a.bmax, "build_s": round(secs, 2), "us_per_insert": round(1e6 * secs / (a.n - a.bulk_first), 1), "stats": st,
                  "avg_accepted_per_batch": round((a.n - a.bulk_first) / st["batches"], 1) if st else None, "recall_at_10": round(rec, 4)}))
