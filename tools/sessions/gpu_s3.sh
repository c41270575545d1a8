#!/bin/bash
# Round-2 GPU call 3: conv_tap.cu with deeper rings, ncu --set full of three tap shapes, bench core (graph replay)
mkdir -p gpurun_out
timeout 200 python tools/tap_probe.py > gpurun_out/r2_tap_probe2.log 2>&1; echo "tap probe rc=$?"; tail -24 gpurun_out/r2_tap_probe2.log
for i in 1 0 4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tap_kernel -s 2 -c 1 -o gpurun_out/r2_tap_shape$i -f \
     python tools/tap_probe.py --only=$i > gpurun_out/r2_ncu_tap$i.log 2>&1; echo "ncu shape $i rc=$?"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_tests_tap.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests_tap.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cudnn-baseline --no-model-stage --no-cpu-baseline > gpurun_out/r2_bench_core.json 2> gpurun_out/r2_bench_core.err; echo "bench core rc=$?"; tail -12 gpurun_out/r2_bench_core.err
timeout 700 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -12 gpurun_out/r2_bench_n1.err
