#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest "tests/test_oracle_fixtures.py" tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_r2l.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_r2l.log
for dbg in 0 2; do
BAOFIT_FUSED_DEBUG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2l_$dbg.json 2> $OUT/bench_r2l_$dbg.err; echo "bench dbg=$dbg exit $?"; tail -3 $OUT/bench_r2l_$dbg.err
python - <<PY
import json
d=json.load(open("$OUT/bench_r2l_$dbg.json"))
print("dbg $dbg ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"], "checksum", d["chi2_checksum"], "e2e", d["e2e"]["value"])
PY
done
