#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/wgrad_probe.py --splits --quick > gpurun_out/r2_wgrad_splits2.log 2>&1; grep -v "^$" gpurun_out/r2_wgrad_splits2.log | grep -E "==|default"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cudnn-baseline --no-cpu-baseline --no-model-stage > gpurun_out/r2_bench_core2.json 2> gpurun_out/r2_bench_core2.err; echo "bench core rc=$?"; tail -2 gpurun_out/r2_bench_core2.err
timeout 300 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -x > gpurun_out/r2_tests14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tests14.log
