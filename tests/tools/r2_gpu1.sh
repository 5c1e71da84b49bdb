#!/bin/bash
# engine v2: parity subset, stage-by-stage debug, bench Ne + C5 (each under its own timeout: a barrier mismatch in a cooperative kernel hangs)
mkdir -p gpurun_out
timeout 60 python tests/tools/debug_hbpp_stages.py --n_det 20000 --n_samp 30000 --giant 0 2>&1 | cut -c1-400 | tail -12
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_bracket.py tests/test_hbpp_exact_limit.py tests/test_gpu_unbiased.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; tail -15 gpurun_out/r2_t1.log
timeout 90 python tests/tools/debug_hbpp_stages.py 2>&1 | cut -c1-300 | tail -12
timeout 120 python bench.py > gpurun_out/r2_bench_ne1.log 2> gpurun_out/r2_bench_ne1.err; python - <<'PY'
import json
for f in ['gpurun_out/r2_bench_ne1.log']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, d['ms_per_step'], r['kernels_ms']); print(r.get('stage_phase_us',{}).get('us')); print(r.get('stage_bracket'))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.log','.err')).read()[-1500:])
PY
FRIES_ENGINE=1 timeout 120 python bench.py > gpurun_out/r2_bench_ne1_old.log 2>/dev/null; cut -c1-120 gpurun_out/r2_bench_ne1_old.log
timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_1.log 2>gpurun_out/r2_bench_c5_1.err; python - <<'PY'
import json
for f in ['gpurun_out/r2_bench_c5_1.log']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, d['ms_per_step'], r['kernels_ms']); print(r.get('stage_phase_us',{}).get('us')); print(r.get('stage_bracket'))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.log','.err')).read()[-1500:])
PY
echo "elapsed ${SECONDS}s"
