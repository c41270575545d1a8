#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_tests_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"; tail -8 gpurun_out/r2i_bench_n1.err
