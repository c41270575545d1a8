#!/bin/bash
# Round-2 final validation call: the full GPU suite, the driver's bench command (both arms), the ncu launch list of the
# bench command, and ncu --set full captures of the round-2 tap kernels.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_tests_gpu.log
SECONDS=0
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$? in ${SECONDS}s"; tail -8 gpurun_out/r2e_bench_n1.err
SECONDS=0
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "ref rc=$? in ${SECONDS}s"; tail -c 600 gpurun_out/r2e_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2e_launches.csv \
   python bench.py --steps 1 --warmup 3 --graph off --no-cpu-baseline --no-cudnn-baseline --no-model-stage > gpurun_out/r2e_ncu_bench.log 2>&1; echo "launch list rc=$?"
for i in 0 1; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tap_kernel -s 2 -c 1 -o gpurun_out/r2e_tap_shape$i -f \
     python tools/tap_probe.py --only=$i > gpurun_out/r2e_ncu_tap$i.log 2>&1; echo "ncu tap shape $i rc=$?"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wgrad_tap_kernel -s 2 -c 1 -o gpurun_out/r2e_wgrad_tap -f \
     python tools/tap_probe.py --only=0 > gpurun_out/r2e_ncu_wgrad.log 2>&1; echo "ncu wgrad_tap rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pw_wgrad_pair_kernel -s 2 -c 1 -o gpurun_out/r2e_wgrad_pair -f \
     python tools/wgrad_probe.py --quick > gpurun_out/r2e_ncu_wgrad_pair.log 2>&1; echo "ncu wgrad pair rc=$?"
