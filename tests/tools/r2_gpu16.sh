#!/bin/bash
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "synthetic_size" 2>&1 | tail -4
python bench.py --config c5 --steps 8 --warmup 10 > gpurun_out/b_fin2_c5.log 2> gpurun_out/b_fin2_c5.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_fin2_c5.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("c5", d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernels_ms"], r["bracket_hits"], r["kernel"], r["frac"], r.get("iter_frac"))
P
echo "elapsed ${SECONDS}s"
