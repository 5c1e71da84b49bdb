#!/bin/bash
# same-box A/B: stage kernels with (v0) and without (v1) the out-of-line global-list solve
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for rep in 1 2; do
for v in v0 v1; do
for c in h2o ne; do
lib=""; [ $v = v1 ] && lib=$PWD/fries_b200/libfries_b200_v1.so
FRIES_B200_LIB=$lib python bench.py --config $c --steps 40 --warmup 10 > gpurun_out/b_g13_${c}_$v.log 2> gpurun_out/b_g13_${c}_$v.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g13_${c}_$v.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c $v", d["value"], d["ms_per_step"], [r["kernels_ms"][k] for k in ("hbpp_stage0","hbpp_stage1","hbpp_stage2","hbpp_stage3","hbpp_stage4","vec_phase")], d["clocks"]["sm_mhz"])
P
done
done
done
echo "elapsed ${SECONDS}s"
