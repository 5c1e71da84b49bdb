#!/bin/bash
# merge in one pass: parity tests, then same-box A/B against the two-pass kernels
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_drivers.py tests/test_gpu_deterministic.py tests/test_gpu_hostapi.py tests/test_gpu_fullsize.py tests/test_reftests_b200.py -x -q -m gpu 2>&1 | tail -3
for v in fused twopass; do
for c in h2o c5 n2full; do
st=40; wu=10; [ $c = c5 ] && st=8; [ $c = n2full ] && st=5 && wu=3
tp=""; [ $v = twopass ] && tp=1
FRIES_MERGE_TWO_PASS=$tp python bench.py --config $c --steps $st --warmup $wu > gpurun_out/b_g18_${c}_$v.log 2> gpurun_out/b_g18_${c}_$v.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g18_${c}_$v.log").read().strip().splitlines()[-1])
r=d["roofline"]; k=r.get("kernels_ms") or r.get("kernels_ms_per_iteration")
print("$c $v", d["value"], d["ms_per_step"], k.get("merge_insert"), k.get("merge_accum"), d.get("energy_est"))
P
done
done
echo "elapsed ${SECONDS}s"
