#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_d2.py -m gpu -q -x > gpurun_out/r2_tests12.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests12.log
timeout 200 python tools/halo_cost_probe.py 8 > gpurun_out/r2_halo_cost_n8_v2.log 2>&1; echo "halo cost rc=$?"; cat gpurun_out/r2_halo_cost_n8_v2.log
timeout 200 python tools/halo_cost_probe.py 4 > gpurun_out/r2_halo_cost_n4_v2.log 2>&1; echo "halo cost rc=$?"; cat gpurun_out/r2_halo_cost_n4_v2.log
