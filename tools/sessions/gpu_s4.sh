#!/bin/bash
# Round-2 GPU call 4: conv_tap.cu with register-window shifter + interleaved producer; fused BN/ReLU; GEMS+SP test
mkdir -p gpurun_out
timeout 200 python tools/tap_probe.py > gpurun_out/r2_tap_probe3.log 2>&1; echo "tap probe rc=$?"; tail -24 gpurun_out/r2_tap_probe3.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_gems_sp.py tests/test_gpu_d2.py -m gpu -q > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_tests4.log
for i in 1 0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tap_kernel -s 2 -c 1 -o gpurun_out/r2b_tap_shape$i -f \
     python tools/tap_probe.py --only=$i > gpurun_out/r2b_ncu_tap$i.log 2>&1; echo "ncu shape $i rc=$?"
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cudnn-baseline --no-cpu-baseline > gpurun_out/r2_bench_core.json 2> gpurun_out/r2_bench_core.err; echo "bench core rc=$?"; tail -12 gpurun_out/r2_bench_core.err
timeout 300 python bench.py --workload resnet --steps 5 --warmup 3 --no-cudnn-baseline --no-cpu-baseline --no-model-stage > gpurun_out/r2_bench_resnet.json 2> gpurun_out/r2_bench_resnet.err; echo "bench resnet rc=$?"; tail -4 gpurun_out/r2_bench_resnet.err
