#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest "tests/test_oracle_fixtures.py::test_gpu_baseline_configs_at_full_size_256_vectors[cfg4]" tests/test_gpu_parity.py -m gpu -q > $OUT/pytest_r2f.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_r2f.log
for dbg in 0 2; do
BAOFIT_FUSED_DEBUG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2f_$dbg.json 2> $OUT/bench_r2f_$dbg.err; echo "bench dbg=$dbg exit $?"; tail -3 $OUT/bench_r2f_$dbg.err
python - <<PY
import json
d=json.load(open("$OUT/bench_r2f_$dbg.json"))
print("dbg $dbg ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"], "checksum", d["chi2_checksum"], "e2e", d["e2e"]["value"])
PY
done
