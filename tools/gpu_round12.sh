#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/bench_shapes.py --dims 1536 --n 1000000 --metric l2 --m 32 2>&1 | tail -1
timeout 600 python tools/bench_shapes.py --dims 128 --n 100000 --metric l2 --m 16 2>&1 | tail -1
timeout 600 python tools/bench_insert.py 2>&1 | tail -1
timeout 600 python tools/bench_insert.py --dims 128 --metric l2 --m 16 --n 10000 2>&1 | tail -1
