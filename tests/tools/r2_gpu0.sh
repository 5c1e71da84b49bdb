#!/bin/bash
# round 2, first GPU call: the whole tier with -rxX, the two round-1 XFAILs with their assertion text, bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -rxX > gpurun_out/r2_t0.log 2>&1; tail -25 gpurun_out/r2_t0.log
timeout 300 python -m pytest tests/test_zz_gpu_fullsize.py tests/test_zz_gpu_hostapi.py -m gpu --runxfail -q -x --tb=long > gpurun_out/r2_xfail_a.log 2>&1; tail -60 gpurun_out/r2_xfail_a.log
timeout 200 python -m pytest tests/test_zz_gpu_hostapi.py -m gpu --runxfail -q --tb=long > gpurun_out/r2_xfail_b.log 2>&1; tail -60 gpurun_out/r2_xfail_b.log
timeout 120 python bench.py > gpurun_out/r2_bench_ne0.log 2> gpurun_out/r2_bench_ne0.err; cut -c1-600 gpurun_out/r2_bench_ne0.log
timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_0.log 2>/dev/null; cut -c1-600 gpurun_out/r2_bench_c5_0.log
echo "elapsed ${SECONDS}s"
