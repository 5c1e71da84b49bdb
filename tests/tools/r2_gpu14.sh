#!/bin/bash
# same-box A/B of the bracket's safety factor
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for f in 6 4 3; do
for c in h2o ne c5; do
st=40; [ $c = c5 ] && st=8
FRIES_PRED_FACTOR=$f python bench.py --config $c --steps $st --warmup 10 > gpurun_out/b_g14_${c}_$f.log 2> gpurun_out/b_g14_${c}_$f.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g14_${c}_$f.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c factor $f", d["value"], d["ms_per_step"], [r["kernels_ms"].get(k) for k in ("hbpp_stage0","hbpp_stage1","hbpp_stage2","hbpp_stage3","hbpp_stage4","vec_phase","find_preserve")], r["bracket_hits"], r["stage_bracket"]["candidates"], r["find_preserve_bracket"]["candidates"])
P
done
done
echo "elapsed ${SECONDS}s"
