#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/pw_probe.py --stationary > gpurun_out/r2h_pw_n256.log 2>&1; echo "pw probe rc=$?"; cat gpurun_out/r2h_pw_n256.log
