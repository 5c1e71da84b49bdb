#!/bin/bash
# 8 GPUs: the torchrun multi-GPU check on 8 ranks, then the default bench line at N = 4 and N = 8
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "torchrun" > gpurun_out/t_mg8.log 2>&1; tail -6 gpurun_out/t_mg8.log | cut -c1-300
for n in 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2c_bench_h2o_g$n.log 2> gpurun_out/r2c_bench_h2o_g$n.err; grep -n "Error" -B2 gpurun_out/r2c_bench_h2o_g$n.err | tail -8
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_h2o_g$n.log').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['route'].get('nvlink_GBps_per_gpu'), d['kernels_ms_rank0'], d['config']['stored_dets'], d['rounds'])
except Exception as e: print('ERR', e)
PY
done
echo "elapsed ${SECONDS}s"
