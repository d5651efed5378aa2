#!/bin/bash
# First GPU run of round 2: validate everything that round 1 could not re-run, then A/B the
# opt-in prototypes.  One gpurun call, ~6 min of box time.  Output: gpurun_out/r2_first.log (+ bench JSONs).
mkdir -p gpurun_out
L=gpurun_out/r2_first.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "pytest -m gpu (product library: the measured kernels)"
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
say "pytest -m gpu of the prototypes (libpgemb_b200_proto.so)"
PGEMB_LIB_VARIANT=proto timeout 900 python -m pytest tests/test_gpu_prototypes.py -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
say "smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $L
say "bench default"
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "exit $?" | tee -a $L
say "bench PGEMB_VISITED_PAIRS=1 (no cpu leg)"
PGEMB_LIB_VARIANT=proto PGEMB_VISITED_PAIRS=1 timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench_vpairs.json 2> gpurun_out/r2_bench_vpairs.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
for f in ("r2_bench_default", "r2_bench_vpairs"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
say "latency (both kernel modes), product library"
timeout 600 python tools/bench_latency.py 2>&1 | tail -1 | tee -a $L
say "latency prototypes, A/B in one process on the prototype library (graph built once)"
PGEMB_LIB_VARIANT=proto timeout 900 python tools/bench_latency_ab.py 2> gpurun_out/r2_latency_ab.err | tee -a $L
say "1536-d L2 (configs[3] row shape): 4 lanes/row vs 8 lanes/row"
timeout 600 python tools/bench_shapes.py --dims 1536 --n 500000 --metric l2 --m 32 2>&1 | tail -1 | cut -c1-500 | tee -a $L
PGEMB_LIB_VARIANT=proto PGEMB_L2_TPR8=1 timeout 600 python tools/bench_shapes.py --dims 1536 --n 500000 --metric l2 --m 32 2>&1 | tail -1 | cut -c1-500 | tee -a $L
say "configs[1] shape (dims 128, N 100K, L2, m 16: lives in L2, issue-bound): bulk-copy gather vs LDGSTS gather"
timeout 600 python tools/bench_shapes.py --dims 128 --n 100000 --metric l2 --m 16 2>&1 | tail -1 | cut -c1-500 | tee -a $L
PGEMB_LIB_VARIANT=proto PGEMB_GATHER_LDGSTS=1 timeout 600 python tools/bench_shapes.py --dims 128 --n 100000 --metric l2 --m 16 2>&1 | tail -1 | cut -c1-500 | tee -a $L
say "north-star shape with the LDGSTS gather (rows then allocate in L2 without the evict-first policy)"
PGEMB_LIB_VARIANT=proto PGEMB_GATHER_LDGSTS=1 timeout 600 python bench.py --no-cpu > gpurun_out/r2_bench_ldgsts.json 2> gpurun_out/r2_bench_ldgsts.err; echo "exit $?" | tee -a $L; cut -c1-300 gpurun_out/r2_bench_ldgsts.json | tee -a $L
say "exact scan: per-pair kernel vs tiled (64 queries x 1M rows)"
timeout 600 python tools/bench_shapes.py --dims 768 --n 1000000 --metric cosine --m 32 --steps 2 --scan-queries 64 2>&1 | tail -1 | cut -c1-400 | tee -a $L
PGEMB_LIB_VARIANT=proto PGEMB_SCAN_TILED=1 timeout 600 python tools/bench_shapes.py --dims 768 --n 1000000 --metric cosine --m 32 --steps 2 --scan-queries 64 2>&1 | tail -1 | cut -c1-400 | tee -a $L
say "exact scan through the tensor-core filter (PGEMB_SCAN_TC=1), 64 and 1024 queries x 1M rows"
PGEMB_LIB_VARIANT=proto PGEMB_SCAN_TC=1 timeout 600 python tools/bench_shapes.py --dims 768 --n 1000000 --metric cosine --m 32 --steps 2 --scan-queries 1024 2>&1 | tail -1 | cut -c1-400 | tee -a $L
say "sidecar: one-query-per-call hnsw_search from 1..128 backend processes (tools/bench_sidecar.py)"
timeout 900 python tools/bench_sidecar.py --backends 1,64 --seconds 3 2> gpurun_out/r2_sidecar.err | tee -a $L
say "exact parallel build, steady state at N~1M (20K exact inserts after a bulk-built prefix): default vs batch clamp"
PGEMB_LIB_VARIANT=proto timeout 600 python tools/bench_build.py --n 1000000 --bulk-first 980000 --bmax 1024 2>&1 | tail -1 | cut -c1-300 | tee -a $L
PGEMB_LIB_VARIANT=proto PGEMB_EXACT_CLAMP_SMS=1 timeout 600 python tools/bench_build.py --n 1000000 --bulk-first 980000 --bmax 1024 2>&1 | tail -1 | cut -c1-300 | tee -a $L
say "sidecar over the prototype library: single-stream small batches + shared-memory visited set"
PGEMB_FAST_SMALL=1 PGEMB_SMEM_VISITED=4096 timeout 900 python tools/bench_sidecar.py --lib pg_embedding_b200/libpgemb_b200_proto.so --backends 1 --seconds 3 2>> gpurun_out/r2_sidecar.err | tee -a $L
