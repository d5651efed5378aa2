#!/bin/bash
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_sharded.py --rows 2000000 2>&1 | tail -2
