#!/bin/bash
# Final evidence of round 2 on one B200: GPU tests, smoke, default bench (both arms), per-workload timings.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref exit $?"
for w in qso_rspace dr9_rspace dr9_kspace dr11_k; do timeout 200 python tools/model_time.py $w 2>&1 | tail -8; done > $OUT/model_time_$TAG.txt; tail -40 $OUT/model_time_$TAG.txt | grep "M evals"
timeout 300 python tools/fit_time.py > $OUT/fit_time_$TAG.txt 2>&1; tail -6 $OUT/fit_time_$TAG.txt
python tools/batch_sweep.py > $OUT/batch_sweep_$TAG.json 2> $OUT/batch_sweep_$TAG.err; echo "sweep exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-scan --no-cpu-baseline --no-fft"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused3_model_chi2_kernel' -s 4 -c 1 -f -o $OUT/prof_fused3_$TAG $CMD > $OUT/ncu_fused3_$TAG.log 2>&1
echo "ncu fused3 exit $?"
