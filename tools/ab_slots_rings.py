#!/usr/bin/env python
"""Sweep of slots x rings per SM (PGEMB_WARPS / PGEMB_RINGS) of the throughput-mode traversal on the headline index, built once:
is the configuration make_search_config chooses still the best one?"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
lib = _lib.load()
n, steps = 1_000_000, 10
B, ef, W = 32768, bench.EFS, 3
X, Q = bench.make_data(torch, n, B * (steps + W))
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, ef, bench.METRIC, capacity=n)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, st)); torch.cuda.synchronize()
idx.build_appended(0, n, 4096)
d_lab = torch.empty((B, ef), dtype=torch.int64, device="cuda"); d_n = torch.empty((B,), dtype=torch.int32, device="cuda")
def step(s):
    _lib.check(lib.pgemb_search_batch_device(idx.dev, B, Q[s * B:(s + 1) * B].data_ptr(), ef, d_lab.data_ptr(), None, None, d_n.data_ptr(), None, st))
for cfg in (sys.argv[1:] or ["0:0", "12:6", "11:6", "10:6", "9:6", "8:7", "8:6", "13:5", "14:5", "16:5", "16:4", "0:0"]):
    w, r = cfg.split(":")
    for k, v in (("PGEMB_WARPS", w), ("PGEMB_RINGS", r)):
        if v != "0": os.environ[k] = v
        else: os.environ.pop(k, None)
    try:
        for s in range(W): step(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(W, W + steps): step(s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"slots": int(w), "rings": int(r), "qps": round(B / (ms * 1e-3), 0), "ms": round(ms, 3)}), flush=True)
    except Exception as e:
        print(json.dumps({"slots": int(w), "rings": int(r), "error": repr(e)[:120]}), flush=True)
