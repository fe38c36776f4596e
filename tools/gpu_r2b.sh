#!/bin/bash
# fused kernel: parity first (short timeouts: a hang must not take the box down), then timing
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest "tests/test_oracle_fixtures.py::test_gpu_baseline_configs_at_full_size_256_vectors[cfg4]" tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_r2b.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_r2b.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2b.json 2> $OUT/bench_r2b.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2b.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in d["kernels"].items(): print("  %-50s %.4f ms share %.3f" % (k, v["ms_per_launch"], v["share"]))
PY
BAOFIT_UNFUSED=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2b_unfused.json 2> $OUT/bench_r2b_unfused.err; echo "bench(unfused) exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2b_unfused.json"))
print("UNFUSED value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "checksum", d["chi2_checksum"])
PY
python - <<PY
import json
print("FUSED checksum", json.load(open("$OUT/bench_r2b.json"))["chi2_checksum"])
PY
