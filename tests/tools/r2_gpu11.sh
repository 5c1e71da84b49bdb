#!/bin/bash
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for dcy in 0 0.7; do
for c in c5 ne; do
st=40; [ $c = c5 ] && st=8
FRIES_PRED_DECAY=$dcy python bench.py --config $c --steps $st --warmup 10 > gpurun_out/b_g11_${c}_$dcy.log 2> gpurun_out/b_g11_${c}_$dcy.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g11_${c}_$dcy.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c decay $dcy", d["value"], d["ms_per_step"], r["kernels_ms"], r["bracket_hits"], r["stage_bracket"], r["find_preserve_bracket"])
P
done
done
echo "elapsed ${SECONDS}s"
