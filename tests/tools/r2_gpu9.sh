#!/bin/bash
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for c in h2o ne; do
python bench.py --config $c --steps 40 --warmup 10 > gpurun_out/b_g9_$c.log 2> gpurun_out/b_g9_$c.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g9_$c.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernels_ms"], r["bracket_hits"], r["stage_bracket"], r["find_preserve_bracket"])
P
done
timeout 600 python -m pytest tests/test_gpu_vecphase.py tests/test_gpu_parity.py tests/test_gpu_hbpp_piv.py -x -q -m gpu 2>&1 | tail -3
echo "elapsed ${SECONDS}s"
