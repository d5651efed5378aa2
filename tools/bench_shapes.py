#!/usr/bin/env python
"""Secondary measurements for DESIGN.md: search QPS / HBM roofline fraction at other BASELINE shapes
(not the bench.py headline).  Same synthetic generator and bulk build as bench.py.
usage: bench_shapes.py --dims 128 --n 100000 --metric l2 --m 16 [--efc 200 --efs 64 --batch 32768 --steps 10]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pg_embedding_b200 as pg
from pg_embedding_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=128); ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--metric", default="l2"); ap.add_argument("--m", type=int, default=16)
ap.add_argument("--efc", type=int, default=200); ap.add_argument("--efs", type=int, default=64)
ap.add_argument("--batch", type=int, default=32768); ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--build-batch", type=int, default=4096); ap.add_argument("--dist", default="clustered", choices=["clustered", "iid"])
ap.add_argument("--scan-queries", type=int, default=0, help="also time the exact scan (pgemb_scan_topk) on this many queries and use it as ground truth")
a = ap.parse_args()
lib = _lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(99)
centres = torch.randn((max(4, int(round(a.n ** 0.5))), a.dims), generator=g, device="cuda")
sigma = 0.3 * (2.0 * a.dims) ** 0.5 / a.dims ** 0.5
def gen(k, seed):
    g.manual_seed(seed)
    out = torch.empty((k, a.dims), device="cuda")
    for s in range(0, k, 1 << 16):
        e = min(k, s + (1 << 16))
        c = torch.randint(0, centres.shape[0], (e - s,), generator=g, device="cuda")
        x = torch.randn((e - s, a.dims), generator=g, device="cuda") if a.dist == "iid" else centres[c] + sigma * torch.randn((e - s, a.dims), generator=g, device="cuda")
        out[s:e] = x / x.norm(dim=1, keepdim=True) if a.metric == "cosine" else x
    return out
X, Q = gen(a.n, 1234), gen(a.batch * (a.steps + 3), 5678)
idx = pg.HnswIndex(a.dims, a.m, a.efc, a.efs, a.metric, capacity=a.n)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.pgemb_index_append_device(idx.dev, a.n, X.data_ptr(), None, None, st)); torch.cuda.synchronize()
build_s = idx.build_appended(0, a.n, a.build_batch)
B, ef = a.batch, a.efs
d_lab = torch.empty((B, ef), dtype=torch.int64, device="cuda"); d_n = torch.empty((B,), dtype=torch.int32, device="cuda")
d_st = torch.empty((B, 4), dtype=torch.int32, device="cuda")
def step(s, stats=False):
    _lib.check(lib.pgemb_search_batch_device(idx.dev, B, Q[s * B:(s + 1) * B].data_ptr(), ef, d_lab.data_ptr(), None, None, d_n.data_ptr(),
                                              d_st.data_ptr() if stats else None, st))
for s in range(3): step(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(3, 3 + a.steps): step(s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
step(3, True); torch.cuda.synchronize()
kms = float(lib.pgemb_last_kernel_ms(idx.dev))
stt = d_st.cpu().numpy().astype(np.int64); nres = d_n.cpu().numpy().astype(np.int64)
byt = int((stt[:, 0] * a.dims * 4 + stt[:, 2] * 4 + nres * 8).sum())
ns = min(1000, B); qs = Q[3 * B:3 * B + ns]
if a.metric == "cosine": truth = torch.topk(qs @ X.T, 10, dim=1).indices
else:
    p = 1 if a.metric == "manhattan" else 2
    truth = torch.cat([torch.topk(torch.cdist(qs[i:i + 64], X, p=p), 10, dim=1, largest=False).indices for i in range(0, ns, 64)])
got = d_lab[:ns, :10].cpu().numpy(); truth = truth.cpu().numpy()
rec = float(np.mean([len(set(truth[i].tolist()) & set(got[i].tolist())) / 10 for i in range(ns)]))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
scan = None
if a.scan_queries:
    import time
    qh = qs[:a.scan_queries].cpu().numpy()
    idx.scan_topk(qh[:1], 10)  # warm: staging buffers, (prototype) library loading
    t0 = time.perf_counter(); sc = idx.scan_topk(qh, 10); t1 = time.perf_counter() - t0
    agree = float(np.mean([len(set(sc["labels"][i].tolist()) & set(truth[i].tolist())) / 10 for i in range(len(qh))]))
    scan = {"queries": len(qh), "seconds": round(t1, 3), "pairs_per_s": round(len(qh) * a.n / t1, 0), "top10_overlap_with_torch_truth": round(agree, 4)}
print(json.dumps({"scan_topk": scan, "shape": vars(a), "build_s": round(build_s, 2), "qps": round(B / (ms * 1e-3), 1), "ms_per_step": round(ms, 3),
                  "achieved_gbs": round(byt / (kms * 1e-3) / 1e9, 1), "frac": round(byt / (kms * 1e-3) / 1e9 / peak, 4),
                  "dist_evals_per_query": float(stt[:, 0].mean()), "expansions_per_query": float(stt[:, 1].mean()), "recall_at_10": round(rec, 4)}))
