#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/wgrad_probe.py --wide > gpurun_out/r2f_wgrad_wide.log 2>&1; echo "wgrad probe rc=$?"; cat gpurun_out/r2f_wgrad_wide.log
