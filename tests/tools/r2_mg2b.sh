#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_vecphase.py -x -q -m gpu 2>&1 | tail -2
export FRIES_BENCH_CPU_ITERS=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_h2o_g2.log 2> gpurun_out/r2_bench_h2o_g2.err; grep -n "Error" -B2 gpurun_out/r2_bench_h2o_g2.err | tail -8; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_h2o_g2.log').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['route'].get('nvlink_GBps_per_gpu'), d['kernels_ms_rank0'], d['config']['stored_dets'], d['rounds'])
    print(d['stage_phase_us']['us'])
except Exception as e: print('ERR', e)
PY
timeout 150 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/r2_b11_c5.log 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b11_c5.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernels_ms'])
PY
echo "elapsed ${SECONDS}s"
