#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_r2q.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_r2q.log
python tools/batch_sweep.py > $OUT/batch_sweep_r2q.json 2> $OUT/batch_sweep_r2q.err; echo "sweep exit $?"; python - <<PY
import json
d=json.load(open("$OUT/batch_sweep_r2q.json"))
for k,v in d.items(): print(k, [(r["B"], round(r["us_per_call"],1)) for r in v])
PY
timeout 300 python bench.py --steps 50 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2q.json 2> $OUT/bench_r2q.err; echo "bench exit $?"; tail -3 $OUT/bench_r2q.err
python - <<PY
import json
d=json.load(open("$OUT/bench_r2q.json"))
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"])
PY
