#!/bin/bash
# Multi-GPU call of round 2 (gpurun --gpus N): the sharded path's tests (both exchanges) and bench.py under torchrun with the
# sharded leg.  usage: bash tools/gpu_r2_multi.sh <N>
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r2_multi$N.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "nvidia-smi topo"; nvidia-smi topo -m 2>&1 | head -12 | tee -a $L
if [ -z "$SKIP_K6" ]; then
say "K6 tests (one GPU of the box)"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -4 | tee -a $L
fi
say "pytest tests/test_gpu_sharded.py (world 2..$N)"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=800 -p no:cacheprovider 2>&1 | tail -12 | tee -a $L
say "bench.py --gpus $N (torchrun): replicas headline + sharded leg"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "exit $?" | tee -a $L
python - <<PY | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n$N.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"])
    print("sharded", json.dumps(d.get("sharded")))
    print("configs4", json.dumps(d.get("configs4")))
except Exception as e:
    print("bench FAILED", e)
PY
tail -8 gpurun_out/r2_bench_n$N.err | cut -c1-300 | tee -a $L
