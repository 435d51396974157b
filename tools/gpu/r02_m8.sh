#!/bin/bash
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread" ; nvidia-smi topo -m 2>/dev/null | head -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_probe_multi.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -20
