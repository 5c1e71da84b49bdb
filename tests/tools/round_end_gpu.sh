#!/bin/bash
# One GPU call at the end of a round: the GPU tier, the bench line, the pivotal-sampler measurement and two ncu captures.
# Every step has its own timeout; the last capture only runs if enough of the call's budget is left (FRIES_GPU_BUDGET s).
BUDGET=${FRIES_GPU_BUDGET:-700}
mkdir -p gpurun_out
# -rxX: the pending cases (tests/test_zz_gpu_*.py, non-strict xfail, child processes) report XPASS / XFAIL with their reason
timeout 400 python -m pytest tests -x -q -m gpu -rxX > gpurun_out/t_final.log 2>&1; tail -15 gpurun_out/t_final.log
# A/B of the one-CTA-per-SM stage kernels
FRIES_STAGE_CTAS=1 timeout 90 python bench.py > gpurun_out/bench_ctas1.log 2> gpurun_out/bench_ctas1.err; cut -c1-300 gpurun_out/bench_ctas1.log
# the same A/B at 1.25e7 samples per GPU, where the stage kernels are issue bound
timeout 120 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_ctas2.log 2>/dev/null; cut -c1-200 gpurun_out/bench_c5_ctas2.log
FRIES_STAGE_CTAS=1 timeout 120 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_ctas1.log 2>/dev/null; cut -c1-200 gpurun_out/bench_c5_ctas1.log
timeout 90 python bench.py > gpurun_out/bench_final2.log 2> gpurun_out/bench_final2.err; cut -c1-300 gpurun_out/bench_final2.log
timeout 45 python tests/tools/bench_piv.py > gpurun_out/bench_piv.log 2>&1; tail -1 gpurun_out/bench_piv.log
timeout 70 ncu --set full --clock-control none --import-source on -k regex:piv_samp_kernel -c 1 -f -o gpurun_out/prof_r1d_piv \
    python tests/tools/bench_piv.py --n 8000000 --n_samp 1600000 --reps 1 > gpurun_out/ncu_piv.log 2>&1; tail -2 gpurun_out/ncu_piv.log
echo "elapsed ${SECONDS}s"
if [ $SECONDS -lt $((BUDGET - 100)) ]; then
    timeout $((BUDGET - SECONDS)) ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats \
        --section MemoryWorkloadAnalysis --section InstructionStats --section Occupancy --section LaunchStats \
        --clock-control none -k regex:hbpp_stage_kernel --launch-skip 19 --launch-count 1 -f -o gpurun_out/prof_r1d_stage4_c5 \
        python bench.py --config c5 --steps 1 --warmup 3 > gpurun_out/ncu_c5.log 2>&1; tail -2 gpurun_out/ncu_c5.log
fi
echo "elapsed ${SECONDS}s"
