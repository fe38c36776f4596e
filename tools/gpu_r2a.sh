#!/bin/bash
# Round 2, first GPU pass: the new parity tests (fixtures at scale, option matrix, metal branches, file round
# trips, toy noise, Xi scan, plan / thread edge cases) and the DMMA+DFMA overlap microbenchmark.
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_oracle_fixtures.py tests/test_host_files.py tests/test_toy_noise.py tests/test_multipole_data.py tests/test_gpu_edges.py -m gpu -q -s > $OUT/pytest_r2a.log 2>&1; echo "pytest exit $?"; tail -40 $OUT/pytest_r2a.log
./tools/fp64_peak > $OUT/fp64_peak_r2a.json 2>&1; cat $OUT/fp64_peak_r2a.json
