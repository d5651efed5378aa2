#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench exit $?"; tail -2 gpurun_out/bench_1m.err; cat gpurun_out/bench_1m.json
PGEMB_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --no-cpu --steps 5 > gpurun_out/ncu_launch.log 2>&1
PGEMB_PROFILE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -o gpurun_out/prof_search_r1_final python bench.py --no-cpu --steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -4
