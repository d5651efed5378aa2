#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "g2 exit $?"; tail -4 gpurun_out/bench_g2.err; cat gpurun_out/bench_g2.json
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
