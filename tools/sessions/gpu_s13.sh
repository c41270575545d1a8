#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests13.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_tests13.log
timeout 200 python tools/halo_cost_probe.py 8 > gpurun_out/r2_halo_cost_n8_v2.log 2>&1; echo "halo cost rc=$?"; tail -1 gpurun_out/r2_halo_cost_n8_v2.log
timeout 200 python tools/halo_cost_probe.py 4 > gpurun_out/r2_halo_cost_n4_v2.log 2>&1; tail -1 gpurun_out/r2_halo_cost_n4_v2.log
timeout 250 python tools/wgrad_probe.py --splits > gpurun_out/r2_wgrad_splits.log 2>&1; echo "splits rc=$?"; grep -v "^$" gpurun_out/r2_wgrad_splits.log | tail -60
