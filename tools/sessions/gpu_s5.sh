#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/tap_probe.py --breakdown > gpurun_out/r2_tap_breakdown.log 2>&1; echo "breakdown rc=$?"; cat gpurun_out/r2_tap_breakdown.log | tail -8
timeout 300 python -m pytest tests/test_gpu_gems_sp.py -m gpu -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_tests5.log
