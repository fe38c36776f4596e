#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, smoke, default bench (both arms), ncu launch list of the headline step,
# ncu --set full capture (with source) of the fused kernel and of the FFT-model kernels.
# Usage: gpurun -- bash tools/gpu_profile_r02.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused3_model_chi2_kernel' -s 4 -c 1 -f -o $OUT/prof_fused3_$TAG $CMD > $OUT/ncu_fused3_$TAG.log 2>&1
echo "ncu fused3 exit $?"
python tools/batch_sweep.py > $OUT/batch_sweep_$TAG.json 2> $OUT/batch_sweep_$TAG.err; echo "sweep exit $?"
timeout 120 tools/fp64_mix > $OUT/fp64_mix_$TAG.json; echo "mix exit $?"
