#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tap_probe.py > gpurun_out/r2_tap_probe5.log 2>&1; echo "tap probe rc=$?"; grep wgrad gpurun_out/r2_tap_probe5.log; tail -3 gpurun_out/r2_tap_probe5.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q -x > gpurun_out/r2_tests7.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_tests7.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cudnn-baseline --no-cpu-baseline > gpurun_out/r2_bench_core.json 2> gpurun_out/r2_bench_core.err; echo "bench core rc=$?"; tail -4 gpurun_out/r2_bench_core.err
timeout 300 python bench.py --workload resnet --steps 5 --warmup 3 --no-cudnn-baseline --no-cpu-baseline --no-model-stage > gpurun_out/r2_bench_resnet.json 2> gpurun_out/r2_bench_resnet.err; echo "bench resnet rc=$?"; tail -2 gpurun_out/r2_bench_resnet.err
