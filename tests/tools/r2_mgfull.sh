#!/bin/bash
# bench.py --config n2full on N GPUs (BASELINE configs[3])
mkdir -p gpurun_out
N=${1:-2}
export FRIES_BENCH_CPU_ITERS=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2957$N bench.py --config n2full --gpus $N --steps 5 --warmup 3 > gpurun_out/r2c_bench_n2full_g$N.log 2> gpurun_out/r2c_bench_n2full_g$N.err; echo "rc $?"; grep -n "Error\|error" -B2 gpurun_out/r2c_bench_n2full_g$N.err | tail -12 | cut -c1-250
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_n2full_g$N.log').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['stored_dets'], d['config']['spawned_per_iteration'], d['roofline']['kernels_ms_per_iteration_rank0'], d['route']['nvlink_GBps_per_gpu'], d['energy_est'], d['comm_error_epoch'])
except Exception as e: print('ERR', e)
PY
echo "elapsed ${SECONDS}s"
