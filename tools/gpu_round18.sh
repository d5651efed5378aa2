#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "kat or empty" -p no:cacheprovider > gpurun_out/san_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck.log | tail -3
grep -E "Race reported" gpurun_out/san_racecheck.log | sed -E 's/\+0x[0-9a-f]+//; s/=========//' | sort | uniq -c | sort -rn | head -6
