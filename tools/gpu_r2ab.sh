#!/bin/bash
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
$CMD > $OUT/plain_r2ab.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused3_model_chi2_kernel' -s 4 -c 1 -f -o $OUT/prof_fused3_tri $CMD > $OUT/ncu_fused3_tri.log 2>&1
echo "ncu exit $?"
