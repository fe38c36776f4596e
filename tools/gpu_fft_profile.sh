#!/bin/bash
# ncu evidence for the FFT-model path (config/BOSSDR11LyaF_fft.ini shape): launch list + full capture.
TAG=${1:-r02jfft}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/fft_probe.py 18944 3"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "ncu launches exit $?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fft_plane_kernel|fft_eval|dgemm_chi2_kernel' -s 3 -c 3 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_f_$TAG.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/plain_$TAG.log
