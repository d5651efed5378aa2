#!/bin/bash
# Third GPU call of round 2: re-check what changed (K6 re-scoring with one lane per candidate, debug products, the optional
# second ring), A/B the second ring on the headline, then the ncu evidence: a launch list of one pgemb_scan_topk call and
# `--set full` captures of the kernels round 1 had no profile of.
mkdir -p gpurun_out
L=gpurun_out/r2_third.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "K6 tests (products + scans)"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
timeout 600 python bench.py --no-cpu --no-legs --steps 20 > gpurun_out/r3_bench_a.json 2>> gpurun_out/r3_bench.err; echo "exit $?" | tee -a $L
timeout 600 python bench.py --no-cpu --steps 20 > gpurun_out/r3_bench_c.json 2>> gpurun_out/r3_bench.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
for f in ("a", "b", "c"):
    try:
        d = json.load(open(f"gpurun_out/r3_bench_{f}.json"))
        print(f, "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "scan", (d.get("scan_topk") or {}).get("seconds"), (d.get("scan_topk") or {}).get("parity"))
    except Exception as e:
        print(f, "FAILED", e)
PY
say "ncu launch list of one pgemb_scan_topk call (1024 queries x 1M rows): who takes the time"
PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_scan_launches.csv python tools/prof_scan.py > gpurun_out/r3_prof_scan_list.log 2>&1
python - <<'PY' | tee -a $L
import csv
try:
    rows = [r for r in csv.reader(open("gpurun_out/r3_scan_launches.csv")) if len(r) > 5 and r[0].isdigit()]
    for r in rows: print(r[4][:60], r[-1], r[-2])
except Exception as e:
    print("launch list FAILED", e)
PY
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
say "ncu --set full: K6 filter (the largest chunk = 5th filter launch) and its re-scoring"
PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 $NCU -k regex:scan_filter_umma -s 4 -c 1 -o gpurun_out/r3_scan_filter python tools/prof_scan.py > gpurun_out/r3_prof_a.log 2>&1
PGEMB_PROF_SCAN=1024 PGEMB_BENCH_N=1000000 timeout 400 $NCU -k regex:scan_rescore -s 4 -c 1 -o gpurun_out/r3_scan_rescore python tools/prof_scan.py > gpurun_out/r3_prof_b.log 2>&1
say "ncu --set full: latency-mode traversal (one query), 1536-d L2 throughput traversal, manhattan traversal, exact tiled scan, insert path"
timeout 300 $NCU -k regex:search_kernel -c 1 -o gpurun_out/r3_latency python tools/prof_latency.py > gpurun_out/r3_prof_c.log 2>&1
timeout 400 $NCU -k regex:search_kernel -c 1 -o gpurun_out/r3_l2_1536 python tools/prof_shape.py --dims 1536 --n 500000 --metric l2 > gpurun_out/r3_prof_d.log 2>&1
PGEMB_SCAN_TC=0 PGEMB_PROF_SCAN=64 PGEMB_BENCH_N=200000 timeout 300 $NCU -k regex:scan_tile_kernel -c 1 -o gpurun_out/r3_scan_tile python tools/prof_scan.py > gpurun_out/r3_prof_e.log 2>&1
PGEMB_BENCH_N=200000 timeout 300 $NCU -k "regex:select_kernel|backlink_kernel" -c 2 -o gpurun_out/r3_insert python tools/prof_insert.py > gpurun_out/r3_prof_f.log 2>&1
for r in r3_scan_filter r3_scan_rescore r3_latency r3_l2_1536 r3_scan_tile r3_insert; do
  [ -f gpurun_out/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r 2>&1 | tail -1 | tee -a $L
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tee -a $L
# gpurun brings back at most 64 MiB: the summaries above are what profiles/ keeps; two reports travel for the source page
rm -f gpurun_out/r3_l2_1536.ncu-rep gpurun_out/r3_scan_tile.ncu-rep gpurun_out/r3_insert.ncu-rep gpurun_out/r3_scan_rescore.ncu-rep
du -sh gpurun_out | tee -a $L
