#!/bin/bash
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_r2y.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_r2y.log
python tools/batch_sweep.py > $OUT/batch_sweep_r2y.json 2> $OUT/batch_sweep_r2y.err; python - <<PY
import json
d=json.load(open("$OUT/batch_sweep_r2y.json"))
for k,v in d.items(): print(k, [(r["B"], round(r["us_per_call"],1)) for r in v])
PY
timeout 300 python tools/fit_time.py 2>&1 | tail -5
