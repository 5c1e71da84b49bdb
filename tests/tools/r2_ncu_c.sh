#!/bin/bash
# round 2 evidence, part 2: launch list of the default bench command (every launch of a 5-iteration run) + the 3000-parent H.v parity test
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "3000_parents" > gpurun_out/t_hv3000.log 2>&1; tail -3 gpurun_out/t_hv3000.log
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r02c.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_h2o.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r02_l.log 2>&1; tail -2 gpurun_out/ncu_r02_l.log | cut -c1-200
echo "elapsed ${SECONDS}s"
