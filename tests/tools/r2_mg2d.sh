#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "det_space or frisys_mol_driver" > gpurun_out/t_mg2d.log 2>&1; tail -12 gpurun_out/t_mg2d.log
timeout 600 python -m pytest tests/test_gpu_drivers.py -x -q -m gpu -k "det_space" > gpurun_out/t_mg2d1.log 2>&1; tail -3 gpurun_out/t_mg2d1.log
echo "elapsed ${SECONDS}s"
