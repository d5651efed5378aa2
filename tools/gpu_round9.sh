#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python bench.py --no-cpu --steps 20 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err; echo "bench exit $?"; tail -2 gpurun_out/bench_var.err
python -c "import json;d=json.load(open('gpurun_out/bench_var.json'));print(d['value'], d['roofline']['frac'], d['e2e'], d['recall_at_10'])"
