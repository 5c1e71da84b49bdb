#!/bin/bash
# single-rank build of the stage kernels (no cross-rank code, staged solve for long lists, row chunks of 4): full GPU tier + bench
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/t_g10.log 2>&1; tail -5 gpurun_out/t_g10.log | cut -c1-300
for c in h2o ne c5; do
st=40; [ $c = c5 ] && st=8
python bench.py --config $c --steps $st --warmup 10 > gpurun_out/b_g10_$c.log 2> gpurun_out/b_g10_$c.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g10_$c.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernels_ms"], r["bracket_hits"], r["frac"], r.get("iter_frac"))
P
done
echo "elapsed ${SECONDS}s"
