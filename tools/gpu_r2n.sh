#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest "tests/test_oracle_fixtures.py" tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_r2n.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_r2n.log
for cfg in "0 0" "2 0" "0 20" "0 24" "0 28" "0 32"; do
set -- $cfg
BAOFIT_FUSED_DEBUG=$1 BAOFIT_FUSED_PACE=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2n_$1_$2.json 2> $OUT/bench_r2n_$1_$2.err; echo "bench dbg=$1 pace=$2 exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2n_$1_$2.json"))
print("dbg $1 pace $2 ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"], "checksum", d["chi2_checksum"], "e2e", d["e2e"]["value"])
PY
done
