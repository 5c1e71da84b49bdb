#!/bin/bash
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_hbpp_exact_limit.py tests/test_gpu_energy_parity.py -x -q -m gpu -k "hbpp or energy" 2>&1 | tail -3
timeout 200 python bench.py > gpurun_out/r2_b12_h2o.log 2>gpurun_out/r2_b12_h2o.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b12_h2o.log').read().strip().splitlines()[-1]); print('h2o', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels_ms'])
PY
timeout 200 python bench.py --config ne > gpurun_out/r2_b12_ne.log 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b12_ne.log').read().strip().splitlines()[-1]); print('ne', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels_ms'])
PY
timeout 400 python bench.py --config n2full --steps 5 --warmup 3 > gpurun_out/r2_b12_n2full.log 2>gpurun_out/r2_b12_n2full.err; tail -c 1800 gpurun_out/r2_b12_n2full.log; tail -3 gpurun_out/r2_b12_n2full.err
echo "elapsed ${SECONDS}s"
