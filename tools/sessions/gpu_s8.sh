#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tap_probe.py > gpurun_out/r2_tap_probe6.log 2>&1; echo "tap probe rc=$?"; grep wgrad gpurun_out/r2_tap_probe6.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q -x > gpurun_out/r2_tests8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests8.log
