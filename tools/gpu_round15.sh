#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exact_parallel" --timeout=900 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/bench_build.py --n 100000 --bmax 64 2>&1 | tail -1
timeout 600 python tools/bench_build.py --n 100000 --bmax 256 2>&1 | tail -1
