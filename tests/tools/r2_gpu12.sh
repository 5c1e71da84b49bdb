#!/bin/bash
# stage kernels: global-list solve restored for overflowing staging areas (out of line); CTA-0 reduce out of line
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for c in c5 h2o ne; do
st=40; [ $c = c5 ] && st=8
python bench.py --config $c --steps $st --warmup 10 > gpurun_out/b_g12_$c.log 2> gpurun_out/b_g12_$c.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g12_$c.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernels_ms"], r["bracket_hits"], r["stage_bracket"], r["frac"], r.get("iter_frac"))
P
done
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_vecphase.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -3
echo "elapsed ${SECONDS}s"
