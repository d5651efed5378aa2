#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench exit $?"; tail -3 gpurun_out/bench_1m.err; cat gpurun_out/bench_1m.json
for v in "PGEMB_STAGES=1" "PGEMB_TPR=2" "PGEMB_TPR=1" "PGEMB_SLOTS_PER_SM=2"; do
  env $v timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  echo "== $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_var.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'])")"
done
PGEMB_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --no-cpu --steps 3 > gpurun_out/ncu_launch.log 2>&1
PGEMB_PROFILE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -o gpurun_out/prof_search_r1 python bench.py --no-cpu --steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
