#!/bin/bash
# chunk growth sweep for the K6 scan (k = 64 and 10)
mkdir -p gpurun_out
L=gpurun_out/r2_eighth.log; : > $L
for k in 64 10; do for g in 2 3 4 8; do echo "k $k growth $g" | tee -a $L; PGEMB_SCAN_TC_GROWTH=$g PGEMB_PROF_SCAN_K=$k PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done; done
echo "nq 64 / nq 1, k 64, growth 4 and 8" | tee -a $L
for nq in 64 1; do for g in 4 8; do PGEMB_SCAN_TC_GROWTH=$g PGEMB_PROF_SCAN_K=64 PGEMB_SCAN_TIMING=1 PGEMB_PROF_SCAN=$nq PGEMB_BENCH_N=1000000 timeout 300 python tools/prof_scan.py 2>&1 | grep -i "timing" | tail -1 | tee -a $L; done; done
