#!/bin/bash
# call 18: the scan race fix re-validated (K6 tests), then the per-row pool gather of K3 A/B on the headline index and on the 1536-d L2 shape
mkdir -p gpurun_out
L=gpurun_out/r2_eleventh.log; : > $L
say() { echo "== $*" | tee -a $L; }
say "K6 + sharded-scan tests after the rescore fix"
timeout 600 python -m pytest tests/test_gpu_scan_umma.py -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -3 | tee -a $L
say "row pool parity tests (PGEMB_ROW_POOL=1 through the whole search parity file)"
PGEMB_ROW_POOL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=800 -p no:cacheprovider -x 2>&1 | tail -4 | tee -a $L
say "A/B headline (cosine 768, 1M)"
timeout 600 python tools/ab_rowpool.py 2>&1 | grep -v Warning | tee -a $L
say "A/B dims 1536 L2 500K"
timeout 600 python tools/ab_rowpool.py --n 500000 --dims 1536 --metric l2 --m 32 --configs "0:0:0,1:8:0,1:10:0,1:6:0" 2>&1 | grep -v Warning | tee -a $L
