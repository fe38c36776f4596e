#!/bin/bash
OUT=gpurun_out
for cfg in "1 20" "1 24" "2 20" "2 24" "3 50" "3 100" "3 200"; do
set -- $cfg
BAOFIT_FUSED_PACE_MODE=$1 BAOFIT_FUSED_PACE=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2o_$1_$2.json 2> $OUT/bench_r2o_$1_$2.err; echo "bench mode=$1 pace=$2 exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2o_$1_$2.json"))
print("mode $1 pace $2 ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"], "checksum", d["chi2_checksum"])
PY
done
