#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/tap_probe.py > gpurun_out/r2_tap_probe4.log 2>&1; echo "tap probe rc=$?"; tail -20 gpurun_out/r2_tap_probe4.log
timeout 200 python tools/tap_probe.py --breakdown > gpurun_out/r2_tap_breakdown2.log 2>&1; echo "breakdown rc=$?"; tail -6 gpurun_out/r2_tap_breakdown2.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q -x > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_tests6.log
