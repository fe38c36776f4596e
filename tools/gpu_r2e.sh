#!/bin/bash
OUT=gpurun_out
for dbg in 0 2; do
BAOFIT_FUSED_PROF=1 BAOFIT_FUSED_DEBUG=$dbg timeout 300 python bench.py --steps 20 --warmup 3 --no-scan --no-fft --no-cpu-baseline > $OUT/bench_r2e_$dbg.json 2> $OUT/bench_r2e_$dbg.err; echo "bench dbg=$dbg exit $?"
grep "fused prof" $OUT/bench_r2e_$dbg.err | sed -n '1p;2p;9p;10p'
python - <<PY
import json
d=json.load(open("$OUT/bench_r2e_$dbg.json"))
print("dbg $dbg ms", d["ms_per_step"], "fused", d["kernels"]["fused_model_chi2_kernel"]["ms_per_launch"], "checksum", d["chi2_checksum"])
PY
done
