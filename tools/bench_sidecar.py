"""Throughput / latency of the reference-shaped hnsw_search (one query per call) when P backend PROCESSES call it at the
same time through pgemb_sidecar (the forked-backend deployment) -- on a B200:

    python tools/bench_sidecar.py [--rows 1000000 --dims 768 --m 32 --metric cosine --backends 1,16,64,128 --seconds 5]

Prints one JSON line per backend count: aggregate queries/s, per-call latency percentiles, and how the sidecar batched
the calls (launches, mean and largest batch).  The graph is built once in the sidecar (bulk build); parity of the results
is checked by tests/test_sidecar.py, not here.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def backend_main(a):
    """One backend: raw ctypes loop around hnsw_search (no numpy in the timed loop)."""
    from pg_embedding_b200 import sidecar
    sidecar.connect(a.shm)
    idx = sidecar.RemoteIndex(1, a.dims, a.m, a.efc, a.efs, a.metric, capacity=1)
    rng = np.random.default_rng(1000 + a.backend_id)
    q = np.load(a.queries)
    q = np.ascontiguousarray(q[rng.permutation(q.shape[0])], dtype=np.float32)
    lib = sidecar.client()
    free = C.CDLL(None).free
    free.argtypes = [C.c_void_p]
    n, res = C.c_size_t(), C.POINTER(C.c_uint64)()
    meta = C.byref(idx.h.meta)
    ptrs = [q[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(q.shape[0])]
    open(a.out + ".ready", "w").close()
    while not os.path.exists(a.go):
        time.sleep(0.001)
    lat = []
    t_end = time.perf_counter() + a.seconds
    i = 0
    while True:
        t0 = time.perf_counter()
        if t0 >= t_end:
            break
        if not lib.hnsw_search(meta, ptrs[i % len(ptrs)], C.byref(n), C.byref(res)):
            raise RuntimeError(lib.pgemb_client_last_error().decode())
        free(res)
        lat.append(time.perf_counter() - t0)
        i += 1
    np.save(a.out, np.array(lat, np.float64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dims", type=int, default=768)
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--efs", type=int, default=64)
    ap.add_argument("--metric", default="cosine")
    ap.add_argument("--backends", default="1,16,64,128")
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--linger-us", type=int, default=None, help="sidecar's --linger-us (default: the sidecar's own default)")
    ap.add_argument("--lib", default=None, help="C-ABI library the sidecar loads (default: the product library)")
    ap.add_argument("--numpy-data", action="store_true", help="iid numpy data instead of bench.py's generator (no torch / CUDA in this process: emulated runs)")
    # internal: backend mode
    ap.add_argument("--backend-id", type=int, default=-1)
    ap.add_argument("--shm"), ap.add_argument("--queries"), ap.add_argument("--out"), ap.add_argument("--go")
    a = ap.parse_args()
    if a.backend_id >= 0:
        return backend_main(a)

    from pg_embedding_b200 import build, sidecar
    build.build()
    X = None
    if a.dims == 768 and not a.numpy_data:
        import bench  # the BASELINE data generator (clustered mixture, seeds 1234/5678)
        import torch
        X, Q = bench.make_data(torch, a.rows, 8192)
    if X is None:
        rng = np.random.default_rng(1234)
        X = rng.standard_normal((a.rows, a.dims)).astype(np.float32)
        Q = rng.standard_normal((8192, a.dims)).astype(np.float32)
    X = X.cpu().numpy() if hasattr(X, "cpu") else X
    Q = Q.cpu().numpy() if hasattr(Q, "cpu") else Q
    shm = f"/pgemb_bench_{os.getpid()}"
    srv = sidecar.SidecarProcess(shm, lib=a.lib, slots=512, bulk_mb=64, linger_us=a.linger_us)
    srv.wait_ready(120)
    tmp = f"/tmp/pgemb_bench_{os.getpid()}"
    os.makedirs(tmp, exist_ok=True)
    try:
        idx = sidecar.RemoteIndex(1, a.dims, a.m, a.efc, a.efs, a.metric, capacity=a.rows)
        rs = idx.record_bytes
        t0 = time.time()
        step = 16384  # 16384 x 3340 B = 55 MB per request: within the 64 MB bulk area
        for lo in range(0, a.rows, step):
            hi = min(a.rows, lo + step)
            rec = np.zeros((hi - lo, rs), np.uint8)
            rec[:, (2 * a.m + 1) * 4:(2 * a.m + 1) * 4 + a.dims * 4] = np.ascontiguousarray(X[lo:hi]).view(np.uint8)
            rec[:, rs - 8:] = np.arange(lo, hi, dtype=np.uint64).view(np.uint8).reshape(hi - lo, 8)
            idx.append_records(rec)
        t_ship = time.time() - t0
        t_build = idx.build(0, a.rows, batch_max=4096, exact=False)
        print(f"# shipped {a.rows} records in {t_ship:.1f}s, bulk build {t_build:.1f}s", file=sys.stderr)
        qf = os.path.join(tmp, "q.npy")
        np.save(qf, Q)
        for P in [int(x) for x in a.backends.split(",")]:
            go = os.path.join(tmp, f"go{P}")
            s0 = sidecar.stats()
            procs = []
            for b in range(P):
                out = os.path.join(tmp, f"lat_{P}_{b}.npy")
                cmd = [sys.executable, os.path.abspath(__file__), "--backend-id", str(b), "--shm", shm, "--queries", qf, "--out", out, "--go", go,
                       "--dims", str(a.dims), "--m", str(a.m), "--efc", str(a.efc), "--efs", str(a.efs), "--metric", a.metric, "--seconds", str(a.seconds)]
                procs.append((subprocess.Popen(cmd), out))
            while not all(os.path.exists(o + ".ready") or p.poll() is not None for p, o in procs):
                time.sleep(0.01)
            open(go, "w").close()
            for p, _ in procs:
                assert p.wait() == 0
            lat = np.concatenate([np.load(o) for _, o in procs])
            s1 = sidecar.stats()
            nb, ns = s1["batches"] - s0["batches"], s1["searches"] - s0["searches"]
            print(json.dumps({"backends": P, "queries_per_s": round(lat.size / a.seconds, 1), "calls": int(lat.size),
                              "latency_ms": {k: round(float(np.percentile(lat, v)) * 1e3, 3) for k, v in (("p50", 50), ("p90", 90), ("p99", 99))},
                              "launches": nb, "mean_batch": round(ns / max(nb, 1), 2), "max_batch_so_far": s1["max_batch"],
                              "workload": f"dims={a.dims} N={a.rows} {a.metric} m={a.m} efS={a.efs}, one query per hnsw_search call per backend process"}))
    finally:
        sidecar.client().pgemb_client_disconnect()
        srv.stop()


if __name__ == "__main__":
    main()
