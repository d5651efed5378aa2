#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
PGEMB_COPY=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or bind" --timeout=600 -p no:cacheprovider 2>&1 | tail -2
for v in "PGEMB_STAGE_KB=12 PGEMB_STAGES=2" "PGEMB_STAGE_KB=12 PGEMB_STAGES=1" "PGEMB_STAGE_KB=12 PGEMB_STAGES=1 PGEMB_COPY=1" "PGEMB_STAGE_KB=12 PGEMB_STAGES=2 PGEMB_COPY=1" "PGEMB_STAGE_KB=24 PGEMB_STAGES=1" "PGEMB_STAGE_KB=12 PGEMB_STAGES=1 PGEMB_VISITED_HASH=0" "PGEMB_STAGE_KB=16 PGEMB_STAGES=1" "PGEMB_STAGE_KB=8 PGEMB_STAGES=1" "PGEMB_STAGE_KB=24 PGEMB_STAGES=1 PGEMB_COPY=1" "PGEMB_STAGE_KB=48 PGEMB_STAGES=1 PGEMB_COPY=1" "PGEMB_STAGE_KB=12 PGEMB_STAGES=1 PGEMB_L2_PERSIST=0"; do
  env $v timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  echo "== $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_var.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['recall_at_10'])" 2>&1 | tail -1)"
done
