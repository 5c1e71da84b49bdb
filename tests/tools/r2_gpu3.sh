#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    km=r['kernels_ms']; st=[km['hbpp_stage%d'%i] for i in range(5)]
    print(f.split('/')[-1], 'ms/step', d['ms_per_step'], 'stages', st, 'sum', round(sum(st),4), {k:v for k,v in km.items() if not k.startswith('hbpp_stage')})
    print('   phases', r.get('stage_phase_us',{}).get('us'))
except Exception as e: print(f,'ERR',e)
PY
}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_bracket.py tests/test_hbpp_exact_limit.py tests/test_gpu_unbiased.py tests/test_gpu_drivers.py -x -q -m gpu > gpurun_out/r2_t3.log 2>&1; tail -4 gpurun_out/r2_t3.log
timeout 100 python bench.py > gpurun_out/r2_b8_ne.log 2>/dev/null; show gpurun_out/r2_b8_ne.log
timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_b8_c5.log 2>/dev/null; show gpurun_out/r2_b8_c5.log
timeout 150 python bench.py --config h2o > gpurun_out/r2_b8_h2o.log 2>/dev/null; show gpurun_out/r2_b8_h2o.log
echo "elapsed ${SECONDS}s"
