#!/bin/bash
# ncu --set full of the traversal with the per-row pool (why is it slower than the rings?)
mkdir -p gpurun_out
L=gpurun_out/r2_prof_rowpool.log; : > $L
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
PGEMB_ROW_POOL=1 PGEMB_WARPS=12 timeout 600 $NCU -k regex:search_kernel -c 1 -o gpurun_out/rp_rowpool python tools/prof_shape.py > gpurun_out/rp_prof.log 2>&1
tail -2 gpurun_out/rp_prof.log | tee -a $L
python tools/ncu_summary.py gpurun_out/rp_rowpool.ncu-rep gpurun_out/rp_rowpool 2>&1 | tail -1 | tee -a $L
ls -la gpurun_out/*.ncu-rep | tee -a $L
