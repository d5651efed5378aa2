#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench exit $?"; tail -2 gpurun_out/bench_1m.err; cat gpurun_out/bench_1m.json
PGEMB_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --no-cpu --steps 5 > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/launches_r1.csv | tail -6
