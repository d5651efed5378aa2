#!/bin/bash
# coalesced staged re-scoring: tests, growth sweep, launch list
mkdir -p gpurun_out
L=gpurun_out/r2_ninth.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "K6 tests"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -4 | tee -a $L
for k in 64 10; do for g in 2 4 8; do echo "k $k growth $g" | tee -a $L; PGEMB_SCAN_TC_GROWTH=$g PGEMB_PROF_SCAN_K=$k PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done; done
say "launch list, k = 64, default growth"
PGEMB_PROF_SCAN_K=64 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r9_scan_launches.csv python tools/prof_scan.py > gpurun_out/r9_list.log 2>&1
python - <<'PY' | tee -a $L
import csv
try:
    rows = [r for r in csv.reader(open("gpurun_out/r9_scan_launches.csv")) if len(r) > 5 and r[0].isdigit()]
    f = sum(float(r[-1]) for r in rows if "filter" in r[4]); s = sum(float(r[-1]) for r in rows if "rescore" in r[4])
    print("launches", len(rows), "filter total ns", f, "rescore total ns", s)
    for r in rows[-6:]: print(r[4][:50], r[-1], r[-2])
except Exception as e:
    print("launch list FAILED", e)
PY
say "nq 1 and 64, k 64"
for nq in 64 1; do PGEMB_PROF_SCAN_K=64 PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=$nq PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done
