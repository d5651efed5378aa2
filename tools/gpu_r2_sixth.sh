#!/bin/bash
# Sixth GPU call: where does a k = 64 scan spend its time (launch list + full capture of the last re-scoring launch)
mkdir -p gpurun_out
L=gpurun_out/r2_sixth.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "time line, k = 64 and k = 10"
for k in 64 10; do PGEMB_PROF_SCAN_K=$k PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done
say "launch list, k = 64"
PGEMB_PROF_SCAN_K=64 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r6_scan_launches.csv python tools/prof_scan.py > gpurun_out/r6_list.log 2>&1
python - <<'PY' | tee -a $L
import csv
try:
    rows = [r for r in csv.reader(open("gpurun_out/r6_scan_launches.csv")) if len(r) > 5 and r[0].isdigit()]
    for r in rows: print(r[4][:50], r[-1], r[-2])
except Exception as e:
    print("launch list FAILED", e)
PY
say "ncu --set full of the last re-scoring launch (k = 64)"
PGEMB_PROF_SCAN_K=64 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -k regex:scan_rescore -s 4 -c 1 -o gpurun_out/r6_rescore python tools/prof_scan.py > gpurun_out/r6_prof.log 2>&1
python tools/ncu_summary.py gpurun_out/r6_rescore.ncu-rep gpurun_out/r6_rescore 2>&1 | tail -1 | tee -a $L
ncu -i gpurun_out/r6_rescore.ncu-rep --page source --csv 2>/dev/null > gpurun_out/r6_rescore_source.csv
rm -f gpurun_out/r6_rescore.ncu-rep
