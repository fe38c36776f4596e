#!/bin/bash
OUT=gpurun_out
for dbg in 0 1 2; do
BAOFIT_FUSED_DEBUG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2c_$dbg.json 2> $OUT/bench_r2c_$dbg.err; echo "bench dbg=$dbg exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2c_$dbg.json"))
print("dbg $dbg ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"])
PY
done
