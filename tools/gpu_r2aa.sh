#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_oracle_fixtures.py tests/test_gpu_properties.py -m gpu -q -x > $OUT/pytest_r2aa.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_r2aa.log
for dense in 0 1; do
if [ $dense = 1 ]; then export BAOFIT_FUSED_DENSE=1; fi
timeout 300 python bench.py --steps 50 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2aa_$dense.json 2> $OUT/bench_r2aa_$dense.err; echo "bench dense=$dense exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2aa_$dense.json"))
print("dense $dense ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"], "checksum", d["chi2_checksum"], "e2e", d["e2e"]["value"])
PY
done
