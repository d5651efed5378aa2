#!/bin/bash
# Seventh GPU call: re-scoring with the running top-k in shared memory, filter epilogue with one slot reservation per chunk.
mkdir -p gpurun_out
L=gpurun_out/r2_seventh.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "K6 tests"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -4 | tee -a $L
say "time line, k = 64 and k = 10"
for k in 64 10; do PGEMB_PROF_SCAN_K=$k PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done
say "launch list, k = 64"
PGEMB_PROF_SCAN_K=64 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r7_scan_launches.csv python tools/prof_scan.py > gpurun_out/r7_list.log 2>&1
python - <<'PY' | tee -a $L
import csv
try:
    rows = [r for r in csv.reader(open("gpurun_out/r7_scan_launches.csv")) if len(r) > 5 and r[0].isdigit()]
    for r in rows: print(r[4][:50], r[-1], r[-2])
except Exception as e:
    print("launch list FAILED", e)
PY
