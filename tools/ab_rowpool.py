#!/usr/bin/env python
"""A/B of the throughput-mode gather on the headline index (built once): ring pool vs the per-row pool (PGEMB_ROW_POOL) at several
slot counts.  Same queries for every configuration; labels must be identical to the first configuration's.
usage: ab_rowpool.py [--n 1000000] [--configs "0:0:0,1:8:0,1:10:0,1:12:0"]   (ROW_POOL:WARPS:RINGS, 0 = default)"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--configs", default="0:0:0,1:8:0,1:10:0,1:12:0,1:9:0,1:11:0")
ap.add_argument("--dims", type=int, default=0); ap.add_argument("--metric", default=""); ap.add_argument("--m", type=int, default=0)
a = ap.parse_args()
lib = _lib.load()
dims, metric, m = a.dims or bench.DIMS, a.metric or bench.METRIC, a.m or bench.M
B, ef, W = 32768, bench.EFS, 3
g = torch.Generator(device="cuda"); g.manual_seed(99)
centres = torch.randn((max(4, int(round(a.n ** 0.5))), dims), generator=g, device="cuda")
gen = bench.gen_points if metric == "cosine" else bench.gen_points_raw
X, Q = gen(torch, a.n, 1234, centres), gen(torch, B * (a.steps + W), 5678, centres)
idx = pg.HnswIndex(dims, m, bench.EFC, ef, metric, capacity=a.n)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.pgemb_index_append_device(idx.dev, a.n, X.data_ptr(), None, None, st)); torch.cuda.synchronize()
idx.build_appended(0, a.n, 4096)
d_lab = torch.empty((B, ef), dtype=torch.int64, device="cuda"); d_n = torch.empty((B,), dtype=torch.int32, device="cuda")
d_st = torch.empty((B, 4), dtype=torch.int32, device="cuda")
peak, _ = bench.measured_peak_gbs()
def step(s, stats=False):
    _lib.check(lib.pgemb_search_batch_device(idx.dev, B, Q[s * B:(s + 1) * B].data_ptr(), ef, d_lab.data_ptr(), None, None, d_n.data_ptr(),
                                              d_st.data_ptr() if stats else None, st))
ref = None
for cfg in a.configs.split(","):
    rp, w, r = cfg.split(":")
    os.environ["PGEMB_ROW_POOL"] = rp
    for k, v in (("PGEMB_WARPS", w), ("PGEMB_RINGS", r)):
        if v != "0": os.environ[k] = v
        else: os.environ.pop(k, None)
    try:
        for s in range(W): step(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(W, W + a.steps): step(s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        step(W, True); torch.cuda.synchronize()
        stt = d_st.cpu().numpy().astype(np.int64); nres = d_n.cpu().numpy().astype(np.int64)
        byt = int((stt[:, 0] * dims * 4 + stt[:, 2] * 4 + nres * 8).sum())
        lab = d_lab.cpu().numpy().copy()
        if ref is None: ref = lab
        print(json.dumps({"row_pool": int(rp), "warps": int(w), "rings": int(r), "qps": round(B / (ms * 1e-3), 0), "ms": round(ms, 3),
                          "frac": round(byt / (ms * 1e-3) / 1e9 / peak, 4), "same_labels_as_first": bool((lab == ref).all()), "err": int(lib.pgemb_index_poll_error(idx.dev, st))}), flush=True)
    except Exception as e:
        print(json.dumps({"row_pool": int(rp), "warps": int(w), "rings": int(r), "error": repr(e)[:200]}), flush=True)
