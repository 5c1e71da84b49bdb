#!/bin/bash
# vec_phase changes (round-robin P1 tiles, threshold instead of keep flags, helper CTA for the dot products, streaming solve)
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 900 python -m pytest tests/test_gpu_vecphase.py tests/test_gpu_fullsize.py tests/test_gpu_drivers.py tests/test_gpu_energy_parity.py -x -q -m gpu > gpurun_out/t_g6.log 2>&1; tail -5 gpurun_out/t_g6.log
FRIES_CTA_MARKS=1 python bench.py --steps 20 --warmup 5 > gpurun_out/b_g6_h2o.log 2> gpurun_out/b_g6_h2o.err; cut -c1-150 gpurun_out/b_g6_h2o.log; grep "cta marks stage [45]" gpurun_out/b_g6_h2o.err | cut -c1-200
python - <<'P'
import json
d=json.loads(open("gpurun_out/b_g6_h2o.log").read().strip().splitlines()[-1])
print(d["roofline"]["kernels_ms"], d["roofline"]["stage_bracket"], d["roofline"]["find_preserve_bracket"])
P
python bench.py --config ne --steps 20 --warmup 5 > gpurun_out/b_g6_ne.log 2>&1; cut -c1-150 gpurun_out/b_g6_ne.log
echo "elapsed ${SECONDS}s"
