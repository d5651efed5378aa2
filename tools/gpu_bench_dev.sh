#!/bin/bash
mkdir -p gpurun_out
for n in 50000 200000; do
  timeout 900 python bench.py --rows $n --steps 5 --warmup 3 --batch 8192 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "n=$n exit $?"; tail -3 gpurun_out/bench_n$n.err; cat gpurun_out/bench_n$n.json
done
