#!/bin/bash
mkdir -p gpurun_out
echo "== old (e087d82)"; timeout 100 python tools/ab_probe.py tools/_ab/libspconv_old.so 2>&1 | tail -5
echo "== new"; timeout 100 python tools/ab_probe.py 2>&1 | tail -5
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_tests_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2j_bench_n1.err
