#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python tools/bench_latency.py --cpu 2>&1 | tail -1
timeout 300 python tools/bench_insert.py 2>&1 | tail -1
