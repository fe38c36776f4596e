#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fit.py -m gpu -q -x > $OUT/pytest_r2r.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_r2r.log
echo "== k-space scan full"; DBG=1 timeout 300 python tools/scan_probe.py 2>&1 | grep -v "^[0-9]*$" | tail -6
echo "== k-space scan 263 fits"; DBG=1 COUNT=263 timeout 300 python tools/scan_probe.py 2>&1 | tail -6
