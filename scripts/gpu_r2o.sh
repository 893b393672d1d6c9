#!/bin/bash
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/nccl_latency.py 2>&1 | grep -i "can_access\|pipelined" | cut -c1-220
