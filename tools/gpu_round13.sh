#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -3
for v in "PGEMB_X=0" "PGEMB_WARPS=12 PGEMB_RINGS=6" "PGEMB_WARPS=10 PGEMB_RINGS=6"; do
  env $v timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  echo "== $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_var.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'])" 2>&1 | tail -1)"
done
