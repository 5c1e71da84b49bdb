#!/bin/bash
# 2 GPUs: the multi-GPU test tier (torchrun check + both C++ drivers under fries_launch) and the default bench line
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/t_mg2c.log 2>&1; tail -6 gpurun_out/t_mg2c.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_h2o_g2.log 2> gpurun_out/r2c_bench_h2o_g2.err; grep -n "Error" -B2 gpurun_out/r2c_bench_h2o_g2.err | tail -8; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_h2o_g2.log').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['route'].get('nvlink_GBps_per_gpu'), d['kernels_ms_rank0'], d['config']['stored_dets'], d['rounds'])
except Exception as e: print('ERR', e)
PY
echo "elapsed ${SECONDS}s"
