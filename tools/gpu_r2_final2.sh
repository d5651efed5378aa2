#!/bin/bash
# Final validation of round 2's last code: full -m gpu suite, smoke, bench (both arms), the ncu launch list of the bench command and
# one ncu --set full capture of the headline kernel (what roofline.traffic is read from).
mkdir -p gpurun_out
L=gpurun_out/r2_final2.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6 | tee -a $L
say "smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee -a $L
say "bench reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/rf2_bench_ref.json 2> gpurun_out/rf2_bench_ref.err; echo "exit $?" | tee -a $L
cut -c1-500 gpurun_out/rf2_bench_ref.json | tee -a $L
say "bench default"
timeout 900 python bench.py > gpurun_out/rf2_bench.json 2> gpurun_out/rf2_bench.err; echo "exit $?" | tee -a $L
python - <<'PY' | tee -a $L
import json
try:
    d = json.load(open("gpurun_out/rf2_bench.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d["gpu_launches"], "clocks", d["clocks"])
    for leg in ("scan_topk", "configs1", "configs4"):
        v = d.get(leg) or {}
        print(leg, {k: v.get(k) for k in ("value", "seconds", "ms_per_step", "parity", "tensor", "error") if k in v})
except Exception as e:
    print("bench FAILED", e)
PY
say "ncu launch list of the bench command (headline only)"
PGEMB_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-legs --no-cpu > gpurun_out/rf2_list.log 2>&1; echo "exit $?" | tee -a $L   # PGEMB_PROFILE: cudaProfilerStart/Stop bracket exactly the timed region
grep -c search_kernel gpurun_out/r2_launches.csv | tee -a $L
say "ncu --set full of the headline traversal launch"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -k regex:search_kernel -c 1 -o gpurun_out/r2_search_kernel_cosine768 python tools/prof_shape.py > gpurun_out/rf2_prof.log 2>&1; echo "exit $?" | tee -a $L
python tools/ncu_summary.py gpurun_out/r2_search_kernel_cosine768.ncu-rep gpurun_out/r2_search_kernel_cosine768 2>&1 | tail -1 | tee -a $L
rm -f gpurun_out/r2_search_kernel_cosine768.ncu-rep
du -sh gpurun_out | tee -a $L
