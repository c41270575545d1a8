#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/pw_probe.py --boxes > gpurun_out/r2f_pw_box5.log 2>&1; echo "pw probe rc=$?"; cat gpurun_out/r2f_pw_box5.log
