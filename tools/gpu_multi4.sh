#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "g4 exit $?"
tail -12 gpurun_out/bench_g4.err; echo ---; wc -l gpurun_out/bench_g4.json; cat gpurun_out/bench_g4.json | cut -c1-300; free -g | head -2
