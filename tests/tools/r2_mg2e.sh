#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "det_space" > gpurun_out/t_mg2e.log 2>&1; tail -12 gpurun_out/t_mg2e.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_deterministic.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t_mg2e1.log 2>&1; tail -15 gpurun_out/t_mg2e1.log | cut -c1-300
echo "elapsed ${SECONDS}s"
