#!/bin/bash
# Fifth GPU call of round 2: sliced re-scoring with threshold refresh, the scan path releasing the traversal's L2 persistence.
mkdir -p gpurun_out
L=gpurun_out/r2_fifth.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "K6 tests"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -4 | tee -a $L
say "time line of one scan (1024 x 1M), chunk growth 16 / 8 / 32"
for g in 16 8 32; do PGEMB_SCAN_TC_GROWTH=$g PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done
say "bench default (scan leg after the traversal legs: L2 persistence released by the scan)"
timeout 900 python bench.py --no-cpu > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "exit $?" | tee -a $L
PGEMB_SCAN_L2RESET=0 timeout 900 python bench.py --no-cpu > gpurun_out/r5_bench_noreset.json 2>> gpurun_out/r5_bench.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
for f in ("r5_bench", "r5_bench_noreset"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        sc = d.get("scan_topk") or {}
        print(f, "value", d["value"], "frac", d["roofline"]["frac"], "scan s", sc.get("seconds"), "x_of_bound", sc.get("x_of_that_bound"), "rescored", sc.get("rescored_fraction"), sc.get("parity"), "configs1", (d.get("configs1") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
