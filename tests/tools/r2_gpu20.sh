#!/bin/bash
# same-box A/B of the CTAs-per-SM choices with the final build
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
run() {  # label, config
python bench.py --config $2 --steps 40 --warmup 10 > gpurun_out/b_g20.log 2> gpurun_out/b_g20.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g20.log").read().strip().splitlines()[-1])
r=d["roofline"]; k=r["kernels_ms"]
print("$1 $2", d["value"], d["ms_per_step"], [k[x] for x in ("hbpp_stage0","hbpp_stage1","hbpp_stage2","hbpp_stage3","hbpp_stage4","hbpp_finalize","vec_phase")])
P
}
run default h2o; run default ne
FRIES_STAGE2_CTAS=1 run stage1cta h2o
FRIES_STAGE2_CTAS=2 run stage2cta ne
FRIES_VECPHASE_CTAS=1 run vp1cta h2o
FRIES_VECPHASE_CTAS=1 run vp1cta ne
FRIES_VECPHASE_CTAS=2 run vp2cta ne
run default h2o; run default ne
echo "elapsed ${SECONDS}s"
