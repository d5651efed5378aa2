#!/bin/bash
# ncu --set full of the traversal on BASELINE configs[1] (dims 128, N 100K, L2, m 16): what bounds the L2-resident case
mkdir -p gpurun_out
L=gpurun_out/r2_prof_c1.log; : > $L
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
timeout 500 $NCU -k regex:search_kernel -c 1 -o gpurun_out/r2_configs1 python tools/prof_shape.py --dims 128 --n 100000 --metric l2 --m 16 > gpurun_out/rc1_prof.log 2>&1
tail -1 gpurun_out/rc1_prof.log | tee -a $L
python tools/ncu_summary.py gpurun_out/r2_configs1.ncu-rep gpurun_out/r2_configs1 2>&1 | tail -1 | tee -a $L
ncu -i gpurun_out/r2_configs1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k in ('lts__t_bytes.sum.per_second','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__m_xbar2l1tex_read_bytes.sum.per_second','lts__t_sector_hit_rate.pct','dram__bytes_read.sum.per_second','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.avg.per_cycle_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed'):
    if k in h: print(k, r[h.index(k)], rows[1][h.index(k)])
" | tee -a $L
