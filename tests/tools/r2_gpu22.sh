#!/bin/bash
# same-box A/B: row chunks of 4 (cur) vs 2 (c2) vs 6 (c6)
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for rep in 1 2; do
for v in cur c2 c6; do
for c in ne h2o; do
lib=""; [ $v != cur ] && lib=$PWD/fries_b200/libfries_b200_$v.so
FRIES_B200_LIB=$lib python bench.py --config $c --steps 40 --warmup 10 > gpurun_out/b_g22.log 2> gpurun_out/b_g22.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g22.log").read().strip().splitlines()[-1])
r=d["roofline"]; k=r["kernels_ms"]
print("$c $v", d["value"], d["ms_per_step"], [k[x] for x in ("hbpp_stage0","hbpp_stage1","hbpp_stage2","hbpp_stage3","hbpp_stage4","hbpp_finalize","vec_phase")])
P
done
done
done
echo "elapsed ${SECONDS}s"
