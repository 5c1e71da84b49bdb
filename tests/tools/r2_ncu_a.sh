#!/bin/bash
mkdir -p gpurun_out
python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/plain_c5.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hbpp_stage2_kernel -s 18 -c 1 -f -o gpurun_out/prof_r2a_stage3_c5 \
    python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/ncu_r2a_c5.log 2>&1; tail -3 gpurun_out/ncu_r2a_c5.log
python bench.py --steps 2 --warmup 6 > gpurun_out/plain_ne.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hbpp_stage2_kernel -s 43 -c 1 -f -o gpurun_out/prof_r2a_stage3_ne \
    python bench.py --steps 2 --warmup 6 > gpurun_out/ncu_r2a_ne.log 2>&1; tail -3 gpurun_out/ncu_r2a_ne.log
echo "elapsed ${SECONDS}s"
