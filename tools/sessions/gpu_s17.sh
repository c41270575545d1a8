#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/stride_probe.py > gpurun_out/r2f_stride_probe.log 2>&1; echo "stride probe rc=$?"; cat gpurun_out/r2f_stride_probe.log
