#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python tools/bench_shapes.py --dims 768 --n 1000000 --metric cosine --m 32 --dist iid --scan-queries 64 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_1m.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'], d['parity'], d['recall_at_10'])"
