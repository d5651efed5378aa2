#!/bin/bash
# Fourth GPU call of round 2: the re-scoring kernel with a CTA per query, x16 chunk growth, the staged insert heuristic, the
# pinned landing area of the small-batch path -- tests, the bench line, host-side time line of a scan, insert path at N.
mkdir -p gpurun_out
L=gpurun_out/r2_fourth.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -8 | tee -a $L
say "smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $L
say "bench default"
timeout 900 python bench.py > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/r4_bench.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print("scan_topk", json.dumps(d.get("scan_topk")))
    print("configs1", d["configs1"]["value"], d["configs1"]["parity"], d["configs1"]["cpu_baseline"])
except Exception as e:
    print("bench FAILED", e)
PY
say "host-side time line of pgemb_scan_topk (1024 x 1M; second call = warm)"
PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing\|done" | tee -a $L
say "same, 1 query and 64 queries"
PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L
PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=64 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L
say "latency (hnsw_search, one query per call)"
timeout 600 python tools/bench_latency.py 2>&1 | tail -1 | tee -a $L
say "insert path: N = 4K built from scratch (both sides), then inserts at N = 200K"
timeout 600 python tools/bench_insert.py 2>&1 | tail -1 | tee -a $L
PGEMB_SELECT_STAGED=0 timeout 600 python tools/bench_insert.py 2>&1 | tail -1 | tee -a $L
timeout 600 python tools/bench_insert_at.py --n 200000 --inserts 200 2>&1 | tail -1 | tee -a $L
say "sidecar: 1 and 64 backends"
timeout 600 python tools/bench_sidecar.py --backends 1,64 --seconds 3 2> gpurun_out/r4_sidecar.err | tee -a $L
