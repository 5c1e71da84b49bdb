#!/bin/bash
# same-box A/B: per-stage CTA choice (new) vs the previous build (old) vs row chunks of 8 (c8)
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for rep in 1 2; do
for v in new old c8; do
for c in ne h2o; do
lib=""; [ $v = old ] && lib=$PWD/fries_b200/libfries_b200_old.so; [ $v = c8 ] && lib=$PWD/fries_b200/libfries_b200_c8.so
FRIES_B200_LIB=$lib python bench.py --config $c --steps 40 --warmup 10 > gpurun_out/b_g21.log 2> gpurun_out/b_g21.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g21.log").read().strip().splitlines()[-1])
r=d["roofline"]; k=r["kernels_ms"]
print("$c $v", d["value"], d["ms_per_step"], [k[x] for x in ("hbpp_stage0","hbpp_stage1","hbpp_stage2","hbpp_stage3","hbpp_stage4","vec_phase")])
P
done
done
done
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_golden.py tests/test_gpu_bracket.py -x -q -m gpu 2>&1 | tail -3
echo "elapsed ${SECONDS}s"
