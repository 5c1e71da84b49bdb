#!/bin/bash
# one-pass merge only for an index beyond the L2 and few insertions: final check
mkdir -p gpurun_out
export FRIES_BENCH_CPU_ITERS=2
for c in h2o c5 n2full ne; do
st=40; wu=10; [ $c = c5 ] && st=8; [ $c = n2full ] && st=5 && wu=3
python bench.py --config $c --steps $st --warmup $wu > gpurun_out/b_g19_$c.log 2> gpurun_out/b_g19_$c.err
python - <<P
import json
d=json.loads(open("gpurun_out/b_g19_$c.log").read().strip().splitlines()[-1])
r=d["roofline"]; k=r.get("kernels_ms") or r.get("kernels_ms_per_iteration")
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], k, r["kernel"], r["frac"], r.get("iter_frac"), r.get("traffic"))
P
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_drivers.py -x -q -m gpu 2>&1 | tail -3
echo "elapsed ${SECONDS}s"
