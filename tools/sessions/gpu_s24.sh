#!/bin/bash
mkdir -p gpurun_out
echo "== old (e087d82)"; timeout 100 python tools/ab_probe.py tools/_ab/libspconv_old.so 2>&1 | tail -6
echo "== new"; timeout 100 python tools/ab_probe.py 2>&1 | tail -6
echo "== old again"; timeout 100 python tools/ab_probe.py tools/_ab/libspconv_old.so 2>&1 | tail -6
