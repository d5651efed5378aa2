#!/bin/bash
# 8-GPU call: the world-8 sharded test and bench.py --gpus 8 (what the driver's scaling run does at N = 8)
mkdir -p gpurun_out
L=gpurun_out/r2_multi8.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "pytest tests/test_gpu_sharded.py world 8"
timeout 500 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=400 -p no:cacheprovider -k "8" 2>&1 | tail -6 | tee -a $L
say "bench.py --gpus 8 (torchrun)"
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 \
    > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "exit $?" | tee -a $L
python - <<PY | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n8.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"])
    print("sharded", json.dumps(d.get("sharded")))
    print("configs4", json.dumps(d.get("configs4")))
except Exception as e:
    print("bench FAILED", e)
PY
tail -12 gpurun_out/r2_bench_n8.err | cut -c1-300 | tee -a $L
