#!/bin/bash
# Second GPU call of round 2: ncu captures (one GPU, --set full, one launch each) of what round 1 has no profile of.
# Each capture is capped; summaries go to gpurun_out/ (copy the ones to keep into profiles/).  ~8 min of box time.
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -c 1 -f"
say() { echo "== $*" | tee -a gpurun_out/r2_prof.log; }
say "latency-mode traversal kernel, one query (device-pointer entry point)"
timeout 300 $NCU -k regex:search_kernel -o gpurun_out/r2_latency python tools/prof_latency.py > gpurun_out/r2_prof_lat.log 2>&1
say "same with the prototype library: shared-memory visited set + paired test-and-set"
PGEMB_LIB_VARIANT=proto PGEMB_SMEM_VISITED=4096 PGEMB_VISITED_PAIRS=1 timeout 300 $NCU -k regex:search_kernel -o gpurun_out/r2_latency_proto python tools/prof_latency.py > gpurun_out/r2_prof_lat_proto.log 2>&1
say "exact scan: tiled kernel and the tensor-core filter's select kernel (64 queries x 200K rows)"
PGEMB_LIB_VARIANT=proto PGEMB_SCAN_TILED=1 PGEMB_PROF_SCAN=64 PGEMB_BENCH_N=200000 timeout 300 $NCU -k regex:scan_tile_kernel -o gpurun_out/r2_scan_tile python tools/prof_scan.py > gpurun_out/r2_prof_scan_tile.log 2>&1
PGEMB_LIB_VARIANT=proto PGEMB_SCAN_TC=1 PGEMB_PROF_SCAN=64 PGEMB_BENCH_N=200000 timeout 300 $NCU -k regex:scan_select_tc -o gpurun_out/r2_scan_tc python tools/prof_scan.py > gpurun_out/r2_prof_scan_tc.log 2>&1
for r in r2_latency r2_latency_proto r2_scan_tile r2_scan_tc; do
  [ -f gpurun_out/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r 2>&1 | tail -1 | tee -a gpurun_out/r2_prof.log
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tee -a gpurun_out/r2_prof.log
