#!/bin/bash
# Second GPU call of round 2: everything written since the first one -- K6 (tcgen05 scan filter), the promoted variants, f2 at the
# C ABI, bench.py's new legs.  The K6 tests run FIRST and alone under a short timeout: a tensor-core pipeline bug must not cost
# more than a minute of box time (its waits are bounded and trap, see scan_umma_kernel.cuh).
mkdir -p gpurun_out
L=gpurun_out/r2_second.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "K6 raw products (descriptors / swizzle / TMEM read-back)"
timeout 300 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q -x --timeout=240 -p no:cacheprovider -k "product" 2>&1 | tail -15 | tee -a $L
say "K6 scan == exact path"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider -k "not product" -s 2>&1 | grep -v "^$" | tail -40 | tee -a $L
say "pytest -m gpu (everything else)"
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --deselect tests/test_gpu_scan_umma.py 2>&1 | tail -15 | tee -a $L
say "smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $L
say "bench default (headline + scan_topk + configs1 legs)"
timeout 900 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/r2_bench2.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print("scan_topk", json.dumps(d.get("scan_topk")))
    print("configs1", json.dumps(d.get("configs1")))
except Exception as e:
    print("bench FAILED", e)
PY
tail -5 gpurun_out/r2_bench2.err | tee -a $L
say "latency (hnsw_search, one query per call)"
timeout 600 python tools/bench_latency.py 2>&1 | tail -1 | tee -a $L
say "1536-d L2 shape (8 lanes per row now default)"
timeout 600 python tools/bench_shapes.py --dims 1536 --n 500000 --metric l2 --m 32 2>&1 | tail -1 | cut -c1-500 | tee -a $L
say "gather ceiling probe (random 3 KB row gathers, no dependency chain): UBLKCP vs UTMALDG gather4"
for cfg in "0 8 1" "0 4 2" "0 6 1" "0 9 1" "2 8 1" "2 4 2" "1 8 1"; do timeout 60 tools/probe/gather_peak $cfg 2>&1 | tail -1 | tee -a $L; done
