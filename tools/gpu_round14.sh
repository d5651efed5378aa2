#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exact_parallel" --timeout=900 -p no:cacheprovider 2>&1 | tail -8
timeout 600 python tools/bench_build.py --n 100000 --bmax 256 2>&1 | tail -1
timeout 600 python tools/bench_build.py --n 100000 --bmax 1024 2>&1 | tail -1
timeout 600 python tools/bench_build.py --n 100000 --mode bulk 2>&1 | tail -1
