#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "kat or empty or (search_identical and d16m8n2000 and cosine) or (bind_links and d8m4n500 and l2) or (exact_parallel and d3m3)" -p no:cacheprovider > gpurun_out/san_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/san_$tool.log | tail -4
done
