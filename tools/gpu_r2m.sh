#!/bin/bash
# ncu --set full capture (with source) of the fused kernel: normal and with the consumers' DMMAs switched off
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
$CMD > $OUT/plain_r2m.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_r2m.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused3_model_chi2_kernel' -s 4 -c 1 -f -o $OUT/prof_fused3_r2m $CMD > $OUT/ncu_fused3_r2m.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/ncu_fused3_r2m.log | cut -c1-200
BAOFIT_FUSED_DEBUG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused3_model_chi2_kernel' -s 4 -c 1 -f -o $OUT/prof_fused3_dbg2_r2m $CMD > $OUT/ncu_fused3_dbg2_r2m.log 2>&1
echo "ncu exit $?"
