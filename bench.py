#!/usr/bin/env python
"""bench.py -- QPS of the HNSW candidate-scoring path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
    dims=768, N=1M synthetic fp32 vectors (clustered mixture, L2-normalised), cosine `<=>`,
    hnsw(m=32, efconstruction=200, efsearch=64); a *step* = one batch of `--batch` k-NN queries
    (k = efsearch = 64) through the search path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Our arm     : the CUDA path.  `value` = queries/s with the query batch already resident in HBM
              (pgemb_search_batch_device on torch's stream, CUDA-event timed, max over ranks);
              `e2e`   = the same through the host-pointer C-ABI call pgemb_search_batch with pinned HOST
              buffers (H2D of the queries and D2H of labels+counts inside the timed region).
Reference arm (`--impl reference`): the reference's own CPU implementation (oracle/_ref = unmodified
              hnswalg.cpp + distfunc.c on a flat-memory host; falls back to the oracle port if the
              prebuilt .so is absent) on all host threads, timed on a bounded sample of the same queries
              against the SAME graph.

The graph: a sequential reference-exact build of 1M x 768 takes hours on any hardware (SURVEY.md
section 6), so both arms search the graph produced by the GPU bulk builder (pgemb_build_bulk: the
reference's search + heuristics applied in batches).  The build is setup, not
timed.  Search parity on that graph is checked in-run: the CPU reference and the GPU must return
identical labels for the sampled queries (`parity` in the JSON line).

N>1 (`torchrun`): the index (3.3 GB) fits one GPU, so ranks hold replicas and split the queries
(SURVEY.md section 8(e)): no data-path collective, "scaling": "weak" (per-GPU batch fixed).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS, M, EFC, EFS = 768, 32, 200, 64
JSON_OUT = sys.stdout
METRIC = "cosine"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", dest="n", type=int, default=int(os.environ.get("PGEMB_BENCH_N", 1_000_000)),
                    help="index size (default 1M = the BASELINE config; smaller values are for development only)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PGEMB_BENCH_BATCH", 32768)), help="queries per step per GPU")
    ap.add_argument("--build-batch", type=int, default=int(os.environ.get("PGEMB_BENCH_BUILD_BATCH", 4096)))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic data: mixture of ~sqrt(N) Gaussian centres, noise norm = 0.3 x the
# typical inter-centre distance, L2-normalised for cosine; fixed seeds 1234 (base) / 5678 (queries).
# ---------------------------------------------------------------------------------------------------
def gen_points(torch, n, seed, centres, chunk=1 << 16):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    dims = centres.shape[1]
    spacing = float((2.0 * dims) ** 0.5)          # E|c_i - c_j| for N(0,I) centres
    sigma = 0.3 * spacing / float(dims ** 0.5)    # per-coordinate noise
    out = torch.empty((n, dims), dtype=torch.float32, device="cuda")
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        a = torch.randint(0, centres.shape[0], (e - s,), generator=g, device="cuda")
        x = centres[a] + sigma * torch.randn((e - s, dims), generator=g, device="cuda")
        out[s:e] = x / x.norm(dim=1, keepdim=True)
    return out


def make_data(torch, n, nq):
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    centres = torch.randn((max(4, int(round(n ** 0.5))), DIMS), generator=g, device="cuda")
    return gen_points(torch, n, 1234, centres), gen_points(torch, nq, 5678, centres)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic_bytes(batch):
    """DRAM read+write bytes of one traversal launch from the committed ncu --set full capture (profiles/), valid for
    the default 32768-query launch of this workload only; None otherwise."""
    p = os.path.join(ROOT, "profiles", "r1_search_kernel_cosine768_metrics.csv")
    if batch != 32768 or not os.path.isfile(p):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = 0.0
    try:
        for line in open(p):
            f = line.strip().split(",")
            if len(f) >= 4 and f[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[-1]) * scale.get(f[-2], 1.0)
    except Exception:
        return None
    return int(tot) if tot > 0 else None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner) write to fd 1 too, so fd 1 is
    pointed at stderr for the whole run and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    args = parse()
    global JSON_OUT
    JSON_OUT = protect_stdout()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0  # the reference arm is a single-process CPU run
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    import pg_embedding_b200 as pg
    from pg_embedding_b200 import _lib
    lib = _lib.load()
    if pg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")

    n, B, K, W = args.n, args.batch, args.steps, max(args.warmup, 3 if args.impl == "ours" else 0)
    workload = f"dims={DIMS} N={n} cosine m={M} efC={EFC} efS={EFS} (BASELINE configs[2])" + ("" if n == 1_000_000 else " [REDUCED N: development run]")
    t0 = time.time()
    nq_total = B * (K + W)
    X, Q = make_data(torch, n, nq_total if args.impl == "ours" else max(B, 4096))
    torch.cuda.synchronize()
    log(f"[rank {rank}] data generated in {time.time() - t0:.1f}s")

    # ---- the device index + bulk build (setup, untimed) --------------------------------------------
    idx = pg.HnswIndex(DIMS, M, EFC, EFS, METRIC, capacity=n, device=local)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.pgemb_index_append_device(idx.dev, n, X.data_ptr(), None, None, stream))
    torch.cuda.synchronize()
    t0 = time.time()
    build_s = idx.build_appended(0, n, args.build_batch)
    log(f"[rank {rank}] bulk build of {n} nodes: {build_s:.1f}s device ({time.time() - t0:.1f}s wall)")

    ef = EFS
    if args.impl == "reference":
        return reference_arm(args, torch, pg, idx, X, Q, n, K, W)

    # ---- device-resident outputs ---------------------------------------------------------------------
    d_lab = torch.empty((B, ef), dtype=torch.int64, device="cuda")
    d_n = torch.empty((B,), dtype=torch.int32, device="cuda")
    d_stats = torch.empty((B, 4), dtype=torch.int32, device="cuda")

    def step_device(s, want_stats=False):
        q = Q[s * B:(s + 1) * B]
        _lib.check(lib.pgemb_search_batch_device(idx.dev, B, q.data_ptr(), ef, d_lab.data_ptr(), None, None, d_n.data_ptr(),
                                                  d_stats.data_ptr() if want_stats else None, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        step_device(s)
    barrier()
    log(f"[rank {rank}] warm-up done, timing {K} steps of {B} queries")
    if os.environ.get("PGEMB_PROFILE"):   # ncu --profile-from-start off: capture exactly the timed region
        torch.cuda.profiler.start()
    launches0 = int(lib.pgemb_launch_count())
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(W, W + K):
        step_device(s)
    ev1.record()
    torch.cuda.synchronize()
    if os.environ.get("PGEMB_PROFILE"):
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.pgemb_launch_count()) - launches0
    clocks = sampler.stop()
    tms = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * B * K / (ms_max * 1e-3)
    log(f"[rank {rank}] timed region: {ms_max:.1f} ms (max over ranks)")

    # ---- roofline of the dominant kernel (the traversal = gather+score): one more step with counters ----
    step_device(W, want_stats=True)
    torch.cuda.synchronize()
    kms = float(lib.pgemb_last_kernel_ms(idx.dev))
    st = d_stats.cpu().numpy().astype(np.int64)
    nres = d_n.cpu().numpy().astype(np.int64)
    alg_bytes = int((st[:, 0] * DIMS * 4 + st[:, 2] * 4 + nres * 8).sum())
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "search_kernel<cosine> (K3: TMA row gather + exact distance + queue update)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "peak_source": peak_src, "traffic": ncu_traffic_bytes(B) if n == 1_000_000 else None,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(kms, 3),
                "per_query": {"dist_evals": float(st[:, 0].mean()), "expansions": float(st[:, 1].mean()),
                              "bytes": float(alg_bytes / B)}}

    # ---- e2e through the host-pointer C ABI with pinned host buffers ------------------------------------
    NB = min(K + W, 4)                                   # distinct pinned query batches, cycled
    hq = torch.empty((NB, B, DIMS), dtype=torch.float32).pin_memory()
    hq.copy_(Q[: NB * B].view(NB, B, DIMS).cpu())
    hl = torch.empty((B, ef), dtype=torch.int64).pin_memory()
    hn = torch.empty((B,), dtype=torch.int32).pin_memory()
    fp, u64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)

    def step_host(s):
        # the call a host application makes: host pointers in, host pointers out
        _lib.check(lib.pgemb_search_batch(idx.dev, B, C.cast(hq[s % NB].data_ptr(), fp), ef, C.cast(hl.data_ptr(), u64p), None, None,
                                          C.cast(hn.data_ptr(), i32p), None))

    for s in range(2):
        step_host(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        step_host(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    log(f"[rank {rank}] e2e region: {float(te.item()):.3f} s")
    e2e = {"value": round(world * B * K / float(te.item()), 1), "unit": "queries/s",
           "h2d_bytes_per_step": B * DIMS * 4, "d2h_bytes_per_step": B * ef * 8 + B * 4}

    # ---- recall@10 vs exact brute force (rank 0; reported, not tuned) -----------------------------------
    out = None
    if rank == 0:
        ns = min(1000, B)
        qs = Q[W * B: W * B + ns]
        truth = torch.topk(qs @ X.T, 10, dim=1).indices.cpu().numpy()
        step_device(W)
        torch.cuda.synchronize()
        got = d_lab[:ns, :10].cpu().numpy()
        recall = float(np.mean([len(set(truth[i].tolist()) & set(got[i].tolist())) / 10.0 for i in range(ns)]))
        cpu_baseline, parity = None, None
        if not args.no_cpu and world == 1:
            cpu_baseline, parity = cpu_leg(args, idx, Q[W * B:(W + 1) * B], d_lab.cpu().numpy(), d_n.cpu().numpy(), n)
        out = {
            "metric": "QPS @ recall@10, dims=768 N=1M efSearch=64", "value": round(value, 1), "unit": "queries/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_max / K, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "queries_per_step_per_gpu": B, "k": ef, "parallelism": f"replicas x{world}, queries split",
                       "l2": "inputs larger than L2 (3.3 GB index vs 126 MB L2); distinct queries every step",
                       "graph": f"GPU bulk build (batch<={args.build_batch}), {build_s:.1f}s, shared by both arms",
                       "distribution": "mixture of sqrt(N) Gaussians, noise 0.3x inter-centre spacing, L2-normalised; seeds 1234/5678"},
            "recall_at_10": round(recall, 4),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity": parity,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        JSON_OUT.write(json.dumps(out) + "\n")
        JSON_OUT.flush()
    return 0


def host_graph(idx, n, which):
    """Copy the GPU index (reference record layout) into the CPU checker's flat host."""
    from oracle import oracle
    orc = oracle.FlatIndex(which, DIMS, M, EFC, EFS, METRIC, capacity=n)
    chunk = 1 << 16
    for s in range(0, n, chunk):
        orc.load_records(idx.export_records(s, min(chunk, n - s)))
    return orc


def pick_checker():
    from oracle import oracle
    if oracle.available("ref"):
        return "ref", "reference"
    oracle.build("port")
    return "port", "port"


def cpu_leg(args, idx, q_dev, gpu_labels, gpu_n, n):
    which, kind = pick_checker()
    orc = host_graph(idx, n, which)
    cores = os.cpu_count() or 1
    q = q_dev.cpu().numpy()
    cal = orc.search_many(q[:cores * 2], EFS, nthreads=cores, want_labels=False)
    qps_est = max(1.0, cores * 2 / max(cal["seconds"], 1e-6))
    reps = 3                                                          # warm cache, >= 3 repetitions, median
    ns = int(min(q.shape[0], max(cores * 4, qps_est * args.cpu_seconds / reps)))
    # untimed pass over the WHOLE step: warms the 3.3 GB graph (page faults, caches) and is the parity check
    full = orc.search_many(q, EFS, nthreads=cores)
    same = bool((full["labels"] == gpu_labels[:q.shape[0]].view(np.uint64)).all() and (full["n"] == gpu_n[:q.shape[0]]).all())
    secs = sorted(orc.search_many(q[:ns], EFS, nthreads=cores, want_labels=False)["seconds"] for _ in range(reps))
    med = secs[len(secs) // 2]
    base = {"value": round(ns / med, 1), "unit": "queries/s", "cores": cores, "kind": kind,
            "sample": f"{ns} of the step's queries, one reader thread per host core, median of {reps} warm passes "
                      f"({secs[0]:.1f}/{med:.1f}/{secs[-1]:.1f}s), same graph"}
    par = {"queries": int(q.shape[0]), "labels_identical_to_cpu_reference": same}
    orc.close()
    return base, par


def reference_arm(args, torch, pg, idx, X, Q, n, K, W):
    which, kind = pick_checker()
    orc = host_graph(idx, n, which)
    cores = os.cpu_count() or 1
    q = Q.cpu().numpy()
    cal = orc.search_many(q[:cores * 2], EFS, nthreads=cores, want_labels=False)
    qps_est = max(1.0, cores * 2 / max(cal["seconds"], 1e-6))
    per_step = int(min(q.shape[0], max(cores * 2, qps_est * max(2.0, 60.0 / max(1, K + W)))))
    for s in range(W):
        orc.search_many(q[:per_step], EFS, nthreads=cores, want_labels=False)
    t = 0.0
    for s in range(K):
        off = (s * per_step) % max(1, q.shape[0] - per_step + 1)
        t += orc.search_many(q[off:off + per_step], EFS, nthreads=cores, want_labels=False)["seconds"]
    v = round(per_step * K / t, 1)
    sample = f"{per_step} queries per step, one reader thread per host core ({cores}), same GPU-built graph"
    out = {"impl": "reference", "metric": "QPS @ recall@10, dims=768 N=1M efSearch=64", "value": v, "unit": "queries/s",
           "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": round(1e3 * t / K, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"dims={DIMS} N={n} cosine m={M} efC={EFC} efS={EFS} (BASELINE configs[2])", "k": EFS,
                      "queries_per_step": per_step, "graph": "GPU bulk build, shared by both arms"},
           "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    JSON_OUT.write(json.dumps(out) + "\n")
    JSON_OUT.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
