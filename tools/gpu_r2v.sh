#!/bin/bash
OUT=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_quasar_data.py tests/test_multipole_data.py tests/test_fft_model.py -m gpu -q > $OUT/pytest_r2v.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_r2v.log
echo "== k-space scan full"; DBG=1 timeout 300 python tools/scan_probe.py 2>&1 | grep -v "^[0-9]*$" | tail -4
echo "== k-space scan 263 fits"; DBG=1 COUNT=263 timeout 300 python tools/scan_probe.py 2>&1 | tail -4
echo "== QSO scan 263"; DBG=1 COUNT=263 timeout 300 python tools/scan_probe_qso.py 2>&1 | tail -4
