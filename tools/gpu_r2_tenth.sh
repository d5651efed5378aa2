#!/bin/bash
# Round 2, call 16: (1) the repair scheme of the exact parallel build (PGEMB_EXACT_REPAIR) A/B at N ~ 1M, (2) device-I/O scan tests,
# (3) bench.py with the configs[4]-shaped leg on one shard (12.5M x 768).
mkdir -p gpurun_out
L=gpurun_out/r2_tenth.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "pytest scan + exact build"
timeout 900 python -m pytest tests/test_gpu_scan_umma.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -p no:cacheprovider -k "scan or exact_parallel" 2>&1 | tail -4 | tee -a $L
say "exact build steady state at N ~ 1M (20000 inserts after a 1M bulk prefix)"
for cfg in "0 1024" "1 1024" "1 148" "1 32" "1 8"; do
  set -- $cfg
  echo "REPAIR=$1 bmax=$2" | tee -a $L
  PGEMB_EXACT_REPAIR=$1 timeout 600 python tools/bench_build.py --n 1020000 --bulk-first 1000000 --bmax $2 2>&1 | tail -1 | tee -a $L
done
say "bench default (with the configs4 leg)"
timeout 1200 python bench.py > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; echo "exit $?" | tee -a $L
tail -5 gpurun_out/r10_bench.err | tee -a $L
python - <<'PY' | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/r10_bench.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "parity", d.get("parity"))
    print("configs4", json.dumps(d.get("configs4")))
except Exception as e:
    print("bench FAILED", e)
PY
