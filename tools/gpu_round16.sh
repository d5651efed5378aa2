#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_golden_fixtures.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_build.csv python tools/bench_build.py --n 300000 --mode bulk > gpurun_out/ncu_build.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_build.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) > mv:
        name = r[kn].split('(')[0][:60]
        agg[name][0] += 1; agg[name][1] += float(r[mv]) / 1e6
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{t:9.1f} ms  {c:6d} launches  {k}")
PY
