#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 tests/multi_gpu_check.py > gpurun_out/mgcheck_$N.log 2>&1; echo "rc $?"; grep -n "\[multi\]" gpurun_out/mgcheck_$N.log | tail -30 | cut -c1-200; grep -n "Error\|error\|assert\|Traceback" gpurun_out/mgcheck_$N.log | head -20 | cut -c1-250
echo "elapsed ${SECONDS}s"
