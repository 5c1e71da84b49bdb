#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py > gpurun_out/r2_mg2.log 2>&1; grep "^\[multi\]\|Error\|error\|assert" gpurun_out/r2_mg2.log | tail -15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_h2o_g2.log 2> gpurun_out/r2_bench_h2o_g2.err; tail -c 1500 gpurun_out/r2_bench_h2o_g2.err | tail -5; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_h2o_g2.log').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['route'].get('nvlink_GBps_per_gpu'), d['kernels_ms_rank0'], d['config']['stored_dets'])
except Exception as e: print('ERR', e)
PY
timeout 200 python bench.py > gpurun_out/r2_bench_h2o_g1.log 2>gpurun_out/r2_bench_h2o_g1.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_h2o_g1.log').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_table'])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/r2_bench_h2o_g1.err
echo "elapsed ${SECONDS}s"
