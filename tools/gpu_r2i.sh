#!/bin/bash
OUT=gpurun_out
echo "== k-space scan full"; DBG=1 timeout 300 python tools/scan_probe.py 2>&1 | grep -v "^[0-9]*$" | tail -8
echo "== k-space scan 263 fits"; DBG=1 COUNT=263 timeout 300 python tools/scan_probe.py 2>&1 | tail -8
echo "== k-space scan 263 fits again"; DBG=1 FIRST=1000 COUNT=263 timeout 300 python tools/scan_probe.py 2>&1 | tail -8
