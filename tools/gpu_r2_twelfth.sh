#!/bin/bash
# calls 19/20: row pool variants (residue columns; then free allocation + grouping table), A/B again
mkdir -p gpurun_out
L=gpurun_out/r2_thirteenth.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "A/B headline (cosine 768, 1M)"
timeout 600 python tools/ab_rowpool.py --configs "0:0:0,1:8:0,1:10:0,1:12:0,1:14:0" 2>&1 | grep -v Warning | tee -a $L
say "A/B dims 1536 L2 500K"
timeout 600 python tools/ab_rowpool.py --n 500000 --dims 1536 --metric l2 --m 32 --configs "0:0:0,1:8:0,1:10:0" 2>&1 | grep -v Warning | tee -a $L
say "row pool variant tests"
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q --timeout=500 -p no:cacheprovider -k row_pool 2>&1 | tail -3 | tee -a $L
