#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    km=r['kernels_ms']; st=[km['hbpp_stage%d'%i] for i in range(5)]
    print(f.split('/')[-1], 'ms/step', d['ms_per_step'], 'stages', st, 'sum', round(sum(st),4), 'fin', km['hbpp_finalize'])
    print('   phases', r.get('stage_phase_us',{}).get('us'))
except Exception as e: print(f,'ERR',e)
PY
}
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "hbpp or hb_rows" > gpurun_out/r2_t2.log 2>&1; tail -3 gpurun_out/r2_t2.log
timeout 100 python bench.py > gpurun_out/r2_b2_ne_c1.log 2>/dev/null; show gpurun_out/r2_b2_ne_c1.log
FRIES_STAGE2_CTAS=2 timeout 100 python bench.py > gpurun_out/r2_b2_ne_c2.log 2>/dev/null; show gpurun_out/r2_b2_ne_c2.log
timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_b2_c5_c1.log 2>/dev/null; show gpurun_out/r2_b2_c5_c1.log
FRIES_STAGE2_CTAS=2 timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_b2_c5_c2.log 2>/dev/null; show gpurun_out/r2_b2_c5_c2.log
echo "elapsed ${SECONDS}s"
