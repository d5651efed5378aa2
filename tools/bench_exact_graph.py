#!/usr/bin/env python
"""The headline workload on the REFERENCE-EXACT graph: pgemb_build_exact builds the 1M x 768 cosine graph bit-identical to one
million sequential hnsw_add_point calls (speculative batches, tests/test_gpu_parity.py::test_exact_parallel_build_equals_sequential),
which the CPU reference would need hours for; then the same search measurement as bench.py (device-resident batches), recall@10
against exact brute force, and label parity of a query sample against the compiled reference searching the same graph.
bench.py itself uses the 9-second bulk build; this script reports what changes on the exact graph."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--bmax", type=int, default=1024)
ap.add_argument("--parity-queries", type=int, default=2048)
a = ap.parse_args()
lib = _lib.load()
B, ef, W = 32768, bench.EFS, 3
X, Q = bench.make_data(torch, a.n, B * (a.steps + W))
idx = pg.HnswIndex(bench.DIMS, bench.M, bench.EFC, bench.EFS, bench.METRIC, capacity=a.n)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.pgemb_index_append_device(idx.dev, a.n, X.data_ptr(), None, None, st)); torch.cuda.synchronize()
t0 = time.time()
secs, stats = idx.build_exact(0, a.n, a.bmax)
wall = time.time() - t0
d_lab = torch.empty((B, ef), dtype=torch.int64, device="cuda"); d_n = torch.empty((B,), dtype=torch.int32, device="cuda")
d_st = torch.empty((a.steps, B, 4), dtype=torch.int32, device="cuda"); d_nn = torch.empty((a.steps, B), dtype=torch.int32, device="cuda")
def step(s, kt=None):
    _lib.check(lib.pgemb_search_batch_device(idx.dev, B, Q[s * B:(s + 1) * B].data_ptr(), ef, d_lab.data_ptr(), None, None,
                                              (d_nn[kt] if kt is not None else d_n).data_ptr(), d_st[kt].data_ptr() if kt is not None else None, st))
for s in range(W): step(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(W, W + a.steps): step(s, s - W)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
stt = d_st.cpu().numpy().astype(np.int64).reshape(-1, 4); nres = d_nn.cpu().numpy().astype(np.int64).reshape(-1)
alg = int((stt[:, 0] * bench.DIMS * 4 + stt[:, 2] * 4 + nres * 8).sum()) // a.steps
peak, _ = bench.measured_peak_gbs()
ns = 1000
truth = torch.topk(Q[W * B:W * B + ns] @ X.T, 10, dim=1).indices.cpu().numpy()
step(W); torch.cuda.synchronize()
got = d_lab[:ns, :10].cpu().numpy()
recall = float(np.mean([len(set(truth[i].tolist()) & set(got[i].tolist())) / 10.0 for i in range(ns)]))
par = None
if a.parity_queries:
    orc = bench.host_graph(idx, a.n, bench.pick_checker()[0])
    qh = Q[W * B:W * B + a.parity_queries].cpu().numpy()
    ref = orc.search_many(qh, ef, nthreads=os.cpu_count() or 1)
    par = bool((ref["labels"] == d_lab.cpu().numpy()[:a.parity_queries].view(np.uint64)).all())
print(json.dumps({"graph": "reference-exact (pgemb_build_exact)", "n": a.n, "build_s": round(secs, 1), "build_wall_s": round(wall, 1), "us_per_insert": round(1e6 * secs / a.n, 1),
                  "build_stats": stats, "qps": round(B / (ms * 1e-3), 1), "ms_per_step": round(ms, 3), "frac_of_hbm_peak": round(alg / (ms * 1e-3) / 1e9 / peak, 4),
                  "dist_evals_per_query": float(stt[:, 0].mean()), "expansions_per_query": float(stt[:, 1].mean()), "recall_at_10": round(recall, 4),
                  "labels_identical_to_cpu_reference_on_sample": par, "parity_queries": a.parity_queries}))
