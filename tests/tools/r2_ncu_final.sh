#!/bin/bash
# round 2 final evidence: launch list of the default bench command + full captures of the slowest stage kernel and the fused vector kernel
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_r02c.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02c_launches_h2o.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r02c_l.log 2>&1; tail -2 gpurun_out/ncu_r02c_l.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hbpp_stage2_kernel -s 19 -c 1 -f -o gpurun_out/prof_r02c_stage4_h2o python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r02c_s.log 2>&1; tail -2 gpurun_out/ncu_r02c_s.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vec_phase_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02c_vecphase_h2o python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_r02c_v.log 2>&1; tail -2 gpurun_out/ncu_r02c_v.log | cut -c1-200
echo "elapsed ${SECONDS}s"
