#!/bin/bash
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_r2u.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_r2u.log
echo "== k-space scan full"; DBG=1 timeout 300 python tools/scan_probe.py 2>&1 | grep -v "^[0-9]*$" | tail -4
echo "== k-space scan 263 fits"; DBG=1 COUNT=263 timeout 300 python tools/scan_probe.py 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_r2u.json 2> $OUT/bench_r2u.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2u.json"))
print(d["summary"]); print("traffic", d["roofline"]["traffic"])
PY
