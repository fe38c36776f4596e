#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_fft_model.py tests/test_oracle_fixtures.py tests/test_gpu_edges.py -m gpu -q -x > $OUT/pytest_r2t.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_r2t.log
for g in 0 1; do
if [ $g = 1 ]; then export BAOFIT_FFT_EVAL_GLOBAL=1; fi
timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-cpu-baseline > $OUT/bench_r2t_$g.json 2> $OUT/bench_r2t_$g.err; echo "bench global=$g exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_r2t_$g.json"))
f=d["fft"]
print("fft ms", f["ms_per_step"], "evals/s", f["evals_per_s"], {k:round(v["ms_per_launch"],3) for k,v in f["kernels"].items()})
print("toys", f["toymc"]["toys_per_s"])
PY
done
unset BAOFIT_FFT_EVAL_GLOBAL
python tools/batch_sweep.py > $OUT/batch_sweep_r2t.json 2> $OUT/batch_sweep_r2t.err; python - <<PY
import json
d=json.load(open("$OUT/batch_sweep_r2t.json"))
for k,v in d.items(): print(k, [(r["B"], round(r["us_per_call"],1)) for r in v])
PY
