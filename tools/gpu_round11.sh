#!/bin/bash
mkdir -p gpurun_out
python pg_embedding_b200/build.py > /dev/null
for v in "PGEMB_VH_PER_EF=64" "PGEMB_VH_PER_EF=128" "PGEMB_VH_PER_EF=256"; do
  env $v timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
  echo "== $v: $(python -c "import json;d=json.load(open('gpurun_out/bench_var.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'])" 2>&1 | tail -1)"
done
timeout 600 python tools/bench_shapes.py --dims 128 --n 100000 --metric l2 --m 16 2>&1 | tail -1
timeout 600 python tools/bench_shapes.py --dims 1536 --n 1000000 --metric l2 --m 32 2>&1 | tail -1
timeout 600 python tools/bench_shapes.py --dims 768 --n 1000000 --metric manhattan --m 32 2>&1 | tail -1
