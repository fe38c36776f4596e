#!/bin/bash
TAG=${1:-r01fft}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
CMD="python tools/fft_probe.py 4096 3"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fft_plane_kernel|fft_eval_kernel|fft_hsum_kernel' -c 5 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/plain_$TAG.log
