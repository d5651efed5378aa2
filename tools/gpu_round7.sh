#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for v in "PGEMB_X=0" "PGEMB_WARPS=10 PGEMB_RINGS=6" "PGEMB_WARPS=14 PGEMB_RINGS=5" "PGEMB_WARPS=16 PGEMB_RINGS=5" "PGEMB_WARPS=12 PGEMB_RINGS=5" "PGEMB_WARPS=8 PGEMB_RINGS=7"; do
  env $v timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  echo "== $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_var.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['recall_at_10'])" 2>&1 | tail -1)"
done
PGEMB_PROFILE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -o gpurun_out/prof_search_r1c python bench.py --no-cpu --steps 1 > gpurun_out/ncu_full.log 2>&1
