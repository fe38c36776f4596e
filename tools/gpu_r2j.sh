#!/bin/bash
# ncu --set full capture (with source) of the fused model + chi2 kernel on the headline workload
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
$CMD > $OUT/plain_r2j.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_r2j.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused_model_chi2_kernel' -s 4 -c 2 -f -o $OUT/prof_fused_r2j $CMD > $OUT/ncu_fused_r2j.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_fused_r2j.log
python tools/batch_sweep.py > $OUT/batch_sweep_r2j.json 2> $OUT/batch_sweep_r2j.err; echo "sweep exit $?"; tail -c 1500 $OUT/batch_sweep_r2j.json
