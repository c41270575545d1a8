#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/halo_cost_probe.py 8 > gpurun_out/r2_halo_cost_n8.log 2>&1; echo "halo cost rc=$?"; cat gpurun_out/r2_halo_cost_n8.log
timeout 200 python tools/halo_cost_probe.py 4 > gpurun_out/r2_halo_cost_n4.log 2>&1; echo "halo cost rc=$?"; cat gpurun_out/r2_halo_cost_n4.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_halo" > gpurun_out/r2_tests11.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests11.log
