#!/bin/bash
# Round-end evidence run on one B200: GPU tests, default bench (both arms), ncu launch list and
# one --set full capture of the kernels that dominate a step.  Usage: gpurun -- bash tools/gpu_profile.sh <tag>
# MODE (2nd argument): all (default) | evidence (tests, smoke, both bench arms, ncu launch list) | full (ncu --set full only)
TAG=${1:-r01}
MODE=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
if [ "$MODE" != "full" ]; then
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks_$TAG.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref exit $?"; cat $OUT/bench_ref_$TAG.json
kill $SMI
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
fi
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
if [ "$MODE" != "evidence" ]; then
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dgemm_chi2_kernel|kspace_fast_eval_kernel|metal_add_kernel' -s 9 -c 6 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
fi
ls -la $OUT
