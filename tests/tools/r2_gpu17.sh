#!/bin/bash
# same-box A/B: 8 (cur) vs 16 (v2) candidates per thread in the CTA-0 solve
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for rep in 1 2; do
for v in cur v2; do
for c in h2o ne; do
lib=""; [ $v = v2 ] && lib=$PWD/fries_b200/libfries_b200_v2.so
FRIES_B200_LIB=$lib python bench.py --config $c --steps 40 --warmup 10 > gpurun_out/b_g17_${c}_$v.log 2> gpurun_out/b_g17_${c}_$v.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g17_${c}_$v.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c $v", d["value"], d["ms_per_step"], [r["kernels_ms"][k] for k in ("hbpp_stage0","hbpp_stage1","hbpp_stage2","hbpp_stage3","hbpp_stage4","vec_phase")], r["stage_bracket"]["candidates"], r["find_preserve_bracket"]["candidates"], r["stage_phase_us"]["candidate_rounds_us"])
P
done
done
done
echo "elapsed ${SECONDS}s"
