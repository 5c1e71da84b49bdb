#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    km=r['kernels_ms']; st=[km['hbpp_stage%d'%i] for i in range(5)]
    print(f.split('/')[-1], 'ms/step', d['ms_per_step'], 'stages', st, 'sum', round(sum(st),4), {k:v for k,v in km.items() if not k.startswith('hbpp_stage')})
except Exception as e: print(f,'ERR',e)
PY
}
export FRIES_BENCH_CPU_ITERS=2
timeout 100 python -m pytest tests/test_gpu_vecphase.py -x -q -m gpu 2>&1 | tail -2
for v in 1 2; do for s in 1 2; do
FRIES_VECPHASE_CTAS=$v FRIES_STAGE2_CTAS=$s timeout 150 python bench.py --steps 10 > gpurun_out/r2_b10_h2o_v${v}s${s}.log 2>/dev/null; show gpurun_out/r2_b10_h2o_v${v}s${s}.log
done; done
FRIES_VECPHASE_CTAS=2 timeout 100 python bench.py --config ne --steps 10 > gpurun_out/r2_b10_ne_v2.log 2>/dev/null; show gpurun_out/r2_b10_ne_v2.log
FRIES_VECPHASE_CTAS=1 timeout 100 python bench.py --config ne --steps 10 > gpurun_out/r2_b10_ne_v1.log 2>/dev/null; show gpurun_out/r2_b10_ne_v1.log
timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_b10_c5.log 2>/dev/null; show gpurun_out/r2_b10_c5.log
echo "elapsed ${SECONDS}s"
