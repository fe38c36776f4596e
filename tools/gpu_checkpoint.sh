#!/bin/bash
# Full GPU test suite + default bench (both arms), results under gpurun_out/.
TAG=${1:-ckpt}
OUT=gpurun_out
mkdir -p $OUT
[ -n "$SKIP_TESTS" ] || timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke_$TAG.log
SECONDS=0; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $? wall ${SECONDS}s"
[ -n "$SKIP_REF" ] || timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("scan", d["scan"]["points_per_s"], "scan_k", d["scan_kspace"]["points_per_s"], "fft", d["fft"]["evals_per_s"], "toys", d["fft"]["toymc"]["toys_per_s"])
print("clocks", d["clocks"])
for k,v in d["kernels"].items(): print("  %-50s %.4f ms share %.3f" % (k, v["ms_per_launch"], v["share"]))
PY
