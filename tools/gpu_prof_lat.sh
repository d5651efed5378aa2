#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/r1_latency python tools/prof_latency.py > gpurun_out/prof_lat.log 2>&1
tail -3 gpurun_out/prof_lat.log; ls -la gpurun_out/*.ncu-rep
