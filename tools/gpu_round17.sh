#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/bench_latency.py --cpu 2>&1 | tail -1
timeout 600 python tools/bench_insert.py 2>&1 | tail -1
timeout 600 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
timeout 900 python tools/bench_build.py --n 1000000 --bmax 256 --bulk-first 990000 2>&1 | tail -1
