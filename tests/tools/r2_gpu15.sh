#!/bin/bash
# final bench lines of round 2 (one box) + the fused vector kernel at the 1.25e7 size + a GPU test subset
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for c in h2o ne c5 n2full; do
st=40; wu=10; [ $c = c5 ] && st=8; [ $c = n2full ] && st=5 && wu=3
python bench.py --config $c --steps $st --warmup $wu > gpurun_out/b_fin_$c.log 2> gpurun_out/b_fin_$c.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_fin_$c.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernels_ms") or r.get("kernels_ms_per_iteration"), r.get("bracket_hits"), r["frac"], r.get("iter_frac"))
P
done
FRIES_FUSED_VEC=1 python bench.py --config c5 --steps 8 --warmup 10 > gpurun_out/b_fin_c5_fused.log 2> gpurun_out/b_fin_c5_fused.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_fin_c5_fused.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("c5 fused", d["value"], d["ms_per_step"], r["kernels_ms"], r["bracket_hits"])
P
timeout 900 python -m pytest tests/test_gpu_bracket.py tests/test_gpu_vecphase.py tests/test_gpu_fullsize.py tests/test_gpu_drivers.py -x -q -m gpu 2>&1 | tail -3
echo "elapsed ${SECONDS}s"
