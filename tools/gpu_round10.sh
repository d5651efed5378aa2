#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider --durations=8 2>&1 | tail -16
