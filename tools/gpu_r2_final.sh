#!/bin/bash
# Final validation of the round's code: full -m gpu suite, smoke, the bench line (both arms), then the headline workload on the
# reference-exact graph (tools/bench_exact_graph.py).
mkdir -p gpurun_out
L=gpurun_out/r2_final.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -8 | tee -a $L
say "smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $L
say "bench reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/rf_bench_ref.json 2> gpurun_out/rf_bench_ref.err; echo "exit $?" | tee -a $L
cut -c1-400 gpurun_out/rf_bench_ref.json | tee -a $L
say "bench default"
timeout 900 python bench.py > gpurun_out/rf_bench.json 2> gpurun_out/rf_bench.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/rf_bench.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d["gpu_launches"], "clocks", d["clocks"])
    sc = d.get("scan_topk") or {}
    print("scan_topk s", sc.get("seconds"), "x_of_bound", sc.get("x_of_that_bound"), "tensor", sc.get("tensor"), "rescored", sc.get("rescored_fraction"), sc.get("parity"))
    print("configs1", d["configs1"]["value"], d["configs1"]["parity"], d["configs1"]["cpu_baseline"])
except Exception as e:
    print("bench FAILED", e)
PY
say "headline workload on the reference-exact graph (1M exact inserts on the GPU)"
timeout 900 python tools/bench_exact_graph.py 2> gpurun_out/rf_exact.err | tail -1 | tee -a $L
